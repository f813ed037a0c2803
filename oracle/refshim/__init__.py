"""
oracle/refshim -- run the reference's own, unmodified ``xcontour/core.py`` in an image
that has neither xarray nor xhistogram (TEST INFRASTRUCTURE; never imported by the
product).

    from oracle import refshim
    ref = refshim.load_reference("/root/reference")     # the `xcontour` package
    ref.Contour2D(...)

``load_reference`` registers stand-in modules for the third-party imports of
``xcontour/core.py:8-13`` and ``xcontour/utils.py:8-12`` (xarray, xhistogram.xarray,
skimage.measure, xgcm, xgcm.autogenerate -- the last three are only touched by code
outside the hot path and are empty stubs) and imports the package from the given
directory without copying it.  /root/reference does not exist on the GPU box: only
the fixture generator (tests/golden/make_reference_golden.py) and a CPU test that is
skipped when the directory is missing call this.
"""
import importlib
import importlib.util
import os
import sys
import types

from . import xarray_shim, xhistogram_shim


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


BACKEND = None        # "refshim" or "xarray <version> + xhistogram <version>" once install() has run


def _importable(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install():
    """Make `import xarray`, `from xhistogram.xarray import histogram`, `from skimage import measure` and
    `from xgcm import Grid` work.  When the REAL xarray and xhistogram are both installed they are used (the
    golden vectors then come from the true stack); otherwise both are replaced by the stand-ins -- never a mix."""
    global BACKEND

    def unavailable(*a, **k):
        raise NotImplementedError("outside the hot path: not provided by oracle/refshim")
    if _importable("xarray") and _importable("xhistogram.xarray"):
        import xarray
        import xhistogram
        BACKEND = "xarray %s + xhistogram %s" % (xarray.__version__, getattr(xhistogram, "__version__", "?"))
    else:
        sys.modules["xarray"] = xarray_shim
        xh = _module("xhistogram")
        xh.xarray = _module("xhistogram.xarray", histogram=xhistogram_shim.histogram)
        sys.modules["xhistogram"] = xh
        sys.modules["xhistogram.xarray"] = xh.xarray
        BACKEND = "refshim"
    if not _importable("skimage.measure"):
        sk = _module("skimage")
        sk.measure = _module("skimage.measure", find_contours=unavailable)
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = sk.measure
    if not _importable("xgcm.autogenerate"):
        xg = _module("xgcm", Grid=unavailable)
        xg.autogenerate = _module("xgcm.autogenerate", generate_grid_ds=unavailable)
        sys.modules["xgcm"] = xg
        sys.modules["xgcm.autogenerate"] = xg.autogenerate


def load_reference(root="/root/reference"):
    """Import the reference's `xcontour` package from `root` on top of the stand-ins."""
    if not os.path.isfile(os.path.join(root, "xcontour", "core.py")):
        raise FileNotFoundError("no reference checkout at %s" % root)
    install()
    if "xcontour" in sys.modules:
        return sys.modules["xcontour"]
    pkg = os.path.join(root, "xcontour")
    spec = importlib.util.spec_from_file_location("xcontour", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["xcontour"] = mod
    nobytecode, sys.dont_write_bytecode = sys.dont_write_bytecode, True     # the mount is read-only
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop("xcontour", None)
        raise
    finally:
        sys.dont_write_bytecode = nobytecode
    return mod
