"""Import the UNMODIFIED reference from /root/reference (read-only) through the mmcv shim.

Only usable in the build container; raises ``ReferenceUnavailable`` elsewhere (e.g. on the GPU box)."""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("E2FGVI_REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mmcv_shim")


class ReferenceUnavailable(RuntimeError):
    pass


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "e2fgvi.py"))


def import_reference(name):
    """``import_reference('model.e2fgvi')`` -> the reference module, isolated from this repo's ``model`` alias."""
    if not available():
        raise ReferenceUnavailable(f"{REFERENCE_ROOT} is not present")
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.") or
                  k == "mmcv" or k.startswith("mmcv.")}
    for k in saved_mods:
        del sys.modules[k]
    # the reference's ``model`` has no __init__.py (namespace package): any regular ``model`` package on the
    # path (this repo's alias) would win regardless of order, so the repo root is dropped for the import
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:] = [SHIM, REFERENCE_ROOT] + [p for p in saved_path
                                            if os.path.abspath(p or os.getcwd()) != repo_root]
    try:
        mod = importlib.import_module(name)
        ref_mods = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.") or
                    k == "mmcv" or k.startswith("mmcv.")}
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "model" or k.startswith("model.") or k == "mmcv" or k.startswith("mmcv."):
                del sys.modules[k]
        sys.modules.update(saved_mods)
    mod.__reference_modules__ = ref_mods  # keep them alive
    assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mod.__file__
    return mod


def reference_generator(hq=False):
    """The reference ``InpaintGenerator()`` (eval mode, CPU)."""
    mod = import_reference("model.e2fgvi_hq" if hq else "model.e2fgvi")
    return mod.InpaintGenerator().eval()
