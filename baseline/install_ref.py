"""Copy recipe for the reference's own hot-path modules.

    python baseline/install_ref.py            # needs /root/reference (the build container)

Copies /root/reference/src/models/{unet,ddpm}.py BYTE FOR BYTE into the git-ignored ``baseline/_ref/models/`` and
records their SHA-256 in ``baseline/_ref/MANIFEST.json``.  Nothing of the reference enters the repository's history;
``baseline/_ref`` is not gpurun-ignored, so it travels to the GPU box with the snapshot, where
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg time THESE modules on the host cores, and
  * tests/test_gpu_reference.py runs them on the GPU under the same seeds as the B200 path.
The reference is plain Python with no build step (no setup.py / pyproject: ``pip install /root/reference`` has nothing
to install), hence a copy, not a pip install.  ``__graft_entry__.build()`` runs this whenever /root/reference exists.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/models"
DST = os.path.join(HERE, "_ref", "models")
FILES = ("unet.py", "ddpm.py")


def install(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: nothing installed (baseline/_ref is built in the container that has the reference)")
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f"src/models/{f}"] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": "/root/reference (mo-rsa24/super-diff-disease), unmodified", "sha256": manifest}, fh, indent=1)
    if verbose:
        print("installed", ", ".join(manifest), "->", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
