"""
Sphere helpers of xcontour/utils.py:491-534 over the GPU element-wise kernels.
"""
import numpy as np

from . import ops
from . import xr_compat as xc

Rearth = 6371200.0   # xcontour/utils.py:19


def _apply(fn, x):
    vals = np.asarray(getattr(x, 'values', x))
    out = fn(ops.to_dev(np.ascontiguousarray(vals, dtype=np.float64)))
    res = out.cpu().numpy().reshape(vals.shape)
    if vals.dtype in (np.float32, np.float64):
        res = res.astype(vals.dtype)                      # .astype(x.dtype) in the reference
    if xc.is_labeled(x):
        return xc.make(res, x.dims, xc.coords_for(x, x.dims), x.name)
    return res


def equivalent_latitudes(areas, Rearth=Rearth):
    """2*pi*a^2*[sin(latEq) + sin(90)] = area  ->  latEq (utils.py:491-515)."""
    if Rearth != globals()['Rearth']:
        raise Exception('only the default Earth radius is supported')
    return _apply(ops.eqlat, areas)


def latitude_lengths_at(lats, Rearth=Rearth):
    """Minimum possible contour length 2*pi*a*cos(lat) (utils.py:518-534)."""
    if Rearth != globals()['Rearth']:
        raise Exception('only the default Earth radius is supported')
    return _apply(ops.lmin, lats)
