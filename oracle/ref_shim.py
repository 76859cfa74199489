"""Import the real reference (``/root/reference/deepliif``) in THIS container only.

The reference imports eight third-party packages at module import time that are absent here
(deepliif/util/__init__.py:13-30, models/__init__.py:37, util/visualizer.py, util/html.py):
dask, skimage, bioformats, javabridge, tifffile, zarr, visdom, dominate.  None of them touches
generator arithmetic; they are replaced by inert stand-ins (dask.delayed/compute are made
functional so ``compute(x)[0]`` at models/__init__.py:327,334 keeps working).

Never imported by the product, never available on the GPU box.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DEEPLIIF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deepliif"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shim():
    if "dask" not in sys.modules:
        _mod("dask", delayed=lambda f: f, compute=lambda *a: a)
    if "skimage" not in sys.modules:
        sk = _mod("skimage")
        sk.filters = _mod("skimage.filters", threshold_multiotsu=lambda *a, **k: None)
        sk.metrics = _mod("skimage.metrics", structural_similarity=lambda *a, **k: None)
    if "bioformats" not in sys.modules:
        bf = _mod("bioformats", JARS=[])
        bf.omexml = _mod("bioformats.omexml")
    for name in ("javabridge", "zarr", "visdom"):
        if name not in sys.modules:
            _mod(name)
    if "tifffile" not in sys.modules:
        _mod("tifffile", TiffFile=object)
    if "dominate" not in sys.modules:
        d = _mod("dominate")
        tags = {t: (lambda *a, **k: None) for t in
                ("meta", "h3", "table", "tr", "td", "p", "a", "img", "br")}
        d.tags = _mod("dominate.tags", **tags)


def import_reference():
    """Returns the reference ``deepliif`` package (models.networks etc. importable after this)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT} (expected on the GPU box)")
    install_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("deepliif")


def reference_networks():
    import_reference()
    return importlib.import_module("deepliif.models.networks")
