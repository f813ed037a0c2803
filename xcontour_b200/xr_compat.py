"""
xarray when it is importable, the in-tree stand-in (labeled.py) otherwise.
Contour2D / Table only use the helpers below, so they are agnostic of which
container carries the labels.
"""
import numpy as np

try:                                   # pragma: no cover - xarray is absent in this image
    import xarray as _xr
    DataArray, Dataset, merge, where = _xr.DataArray, _xr.Dataset, _xr.merge, _xr.where
    HAVE_XARRAY = True
except ImportError:
    from .labeled import DataArray, Dataset, merge, where
    HAVE_XARRAY = False


def is_labeled(x):
    return hasattr(x, "dims") and hasattr(x, "values") and hasattr(x, "coords")


def coord(da, dim):
    """1-D coordinate values of ``dim`` (or None when the dim has no coordinate)."""
    try:
        c = da.coords[dim]
    except (KeyError, AttributeError):
        return None
    return np.asarray(getattr(c, "values", c))


def coords_for(da, dims):
    """{dim: values} for the dims of ``da`` that carry coordinates."""
    out = {}
    for d in dims:
        c = coord(da, d)
        if c is not None and np.ndim(c) == 1:
            out[d] = c
    return out


def make(data, dims, coords=None, name=None):
    return DataArray(data, dims=tuple(dims), coords=coords or {}, name=name)
