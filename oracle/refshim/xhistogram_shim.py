"""
Stand-in for ``xhistogram.xarray.histogram`` (TEST INFRASTRUCTURE, see xarray_shim.py).

xhistogram is a third-party dependency of the reference that is absent from
/root/reference and from this image (setup.py:40-45 lists the bare name; README.md:26
states 0.3.0).  This restates its published algorithm for the one call shape the
reference uses (core.py:1284, 1307: one variable, one explicit edge array, weights,
reduction over `dim`):

* operands are broadcast by dim name, the reduced axes are moved last IN THE ORDER OF
  `dim` and flattened (so the accumulation order follows ``dim``);
* the last edge is nudged by ``+1e-8`` in the edge array's own dtype, values are
  binned with ``np.digitize`` (right=False), under/overflow -- which also collects
  NaN -- is dropped;
* ``np.bincount(..., weights=...)`` accumulates in float64, sequentially.
"""
import numpy as np

from . import xarray_shim as xr


def histogram(*args, bins=None, dim=None, weights=None, density=False, block_size="auto",
              keep_coords=False, bin_dim_suffix="_bin"):
    if len(args) != 1 or density:
        raise NotImplementedError("the reference histograms one variable without density")
    a = args[0]
    b = np.asarray(bins[0] if isinstance(bins, (list, tuple)) else bins)
    dim = [dim] if isinstance(dim, str) else list(dim)
    if weights is not None:
        a, weights = xr.broadcast(a, weights)
    keep = [d for d in a.dims if d not in dim]
    axis = [a.dims.index(d) for d in dim]

    def to2d(x):
        c = np.moveaxis(x, axis, tuple(range(-len(axis), 0)))
        split = c.ndim - len(axis)
        return c.reshape(int(np.prod(c.shape[:split], dtype=np.int64)), -1), c.shape[:split]

    x2, kshape = to2d(a.values)
    w2 = to2d(weights.values)[0] if weights is not None else None
    e = np.concatenate((b[:-1], b[-1:] + 1e-8))                 # in b's dtype
    nb = len(e) + 1                                             # with under/overflow slots
    idx = np.digitize(x2, e)
    idx = idx + (np.arange(x2.shape[0]) * nb)[:, None]
    cnt = np.bincount(idx.ravel(), weights=None if w2 is None else w2.ravel(),
                      minlength=nb * x2.shape[0]).reshape(x2.shape[0], nb)[:, 1:-1]
    name = a.name + bin_dim_suffix
    out = cnt.reshape(tuple(kshape) + (len(b) - 1,))
    coords = {k: v for k, v in a._coords.items() if k in keep}
    coords[name] = 0.5 * (b[:-1] + b[1:])
    return xr.DataArray(out, coords, tuple(keep) + (name,), "histogram_" + a.name)
