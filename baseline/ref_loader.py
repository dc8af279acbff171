"""Loader for the reference's own ``UNet`` / ``DDPM`` classes from ``baseline/_ref`` (see install_ref.py).

TEST / BENCH-BASELINE INFRASTRUCTURE ONLY, like ``oracle/``: imported by ``tests/`` and by ``bench.py``'s CPU legs, never
by the product package.  ``load()`` returns ``(UNet, DDPM)`` or ``None`` when ``baseline/_ref`` has not been installed.
"""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = None


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def available():
    return all(os.path.exists(os.path.join(HERE, "_ref", "models", f)) for f in ("unet.py", "ddpm.py"))


def load():
    global _CACHE
    if _CACHE is None and available():
        d = os.path.join(HERE, "_ref", "models")
        _CACHE = (_load_module("sdd_reference_unet", os.path.join(d, "unet.py")).UNet,
                  _load_module("sdd_reference_ddpm", os.path.join(d, "ddpm.py")).DDPM)
    return _CACHE
