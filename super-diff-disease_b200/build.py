"""Build libsdd_b200.so (hand-written sm_100a CUDA behind the C ABI in include/sdd_b200.h).

In-tree build with nvcc; cross-compiles without a GPU.  The .so is git-ignored but ships to the
GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsdd_b200.so")
SOURCES = ["sdd_api.cu"]
HEADERS = ["common.cuh", "conv_common.cuh", "conv_tc4.cuh", "unet_kernels.cuh", "update.cuh", "train_eval.cuh", "attention.cuh",
           "attention_block.cuh", os.path.join("..", "..", "include", "sdd_b200.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, out=None, flags=()):
    """out / flags: build a VARIANT of the library (extra -D flags) to another path for same-box A/B runs."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if verbose else "-O3",
           "-o", out or LIB] + list(flags) + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libsdd_b200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return out or LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
