"""
Sphere helpers of xcontour/utils.py:491-534 over the GPU element-wise kernels.
"""
import os

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


def latlon_cell_area(lat_deg, lon_deg, Rearth=Rearth):
    """Cell areas dA[j, i] = R^2 (sin(phi_{j+1/2}) - sin(phi_{j-1/2})) dlambda of a
    regular lat-lon grid (cell edges midway between grid latitudes, clipped at the
    poles).  Host-side setup helper -- the reference builds the same metric with
    xgcm in add_latlon_metrics (utils.py:43-259, rA), which is out of scope here;
    Contour2D itself takes dA as an argument."""
    lat = np.asarray(lat_deg, dtype=np.float64)
    lon = np.asarray(lon_deg, dtype=np.float64)
    asc = lat[-1] > lat[0]
    la = lat if asc else lat[::-1]
    edges = np.empty(len(la) + 1)
    edges[1:-1] = 0.5 * (la[1:] + la[:-1])
    edges[0] = max(-90.0, la[0] - 0.5 * (la[1] - la[0]))
    edges[-1] = min(90.0, la[-1] + 0.5 * (la[-1] - la[-2]))
    band = Rearth ** 2 * np.diff(np.sin(np.deg2rad(edges)))
    if not asc:
        band = band[::-1]
    dlam = np.deg2rad(abs(lon[1] - lon[0]))
    return np.repeat((band * dlam)[:, None], len(lon), axis=1)


# Which NumPy scalar-promotion regime of the reference is reproduced where the two differ (hazard H4):
# under the NumPy-1.x rules the reference was written for, `step` of _histogram (core.py:1277) is fp64 and the
# per-'time' edge array is fp64 (core.py:1278), so xhistogram's +1e-8 closes the last bin over the maximum cell; under
# NEP 50 (NumPy >= 2) the edges keep the contour dtype.  The default is the regime of the NumPy that is installed --
# what the reference itself would compute on this machine; XCB200_NUMPY_RULES=numpy1|numpy2 (or assigning this
# variable) selects one explicitly.
def _installed_numpy_rules():
    try:
        return "numpy2" if int(np.__version__.split(".")[0]) >= 2 else "numpy1"
    except ValueError:
        return "numpy1"


NUMPY_SCALAR_RULES = os.environ.get("XCB200_NUMPY_RULES") or _installed_numpy_rules()


def scalar_rules(rules=None):
    """the regime in force: an explicit argument, else the module variable"""
    r = NUMPY_SCALAR_RULES if rules is None else rules
    if r not in ("numpy1", "numpy2"):
        raise Exception("scalar rules should be 'numpy1' or 'numpy2', got %r" % (r,))
    return r


BOUNDARY = {"periodic": 0, "extend": 1, "reflect": 2, "fill": 3}   # XC_BC_* of include/xcb200.h


def row_metrics_latlon(lat_deg, lon_deg, Rearth=Rearth):
    """Row metrics (cx, cy) of the centred-difference |grad q|^2 stencil on a regular lat-lon grid:
        dq/dx = (q[j,i+1] - q[j,i-1]) * cx[j],  cx = 1 / ((2 dlambda) (R cos phi_j))
        dq/dy = (q[j+1,i] - q[j-1,i]) * cy[j],  cy = 1 / ((phi_{j+1} - phi_{j-1}) R)   (one-sided at the ends)
    Host-side setup (ny values), handed to the kernels as fp64 arrays.  The reference's callers get this field
    from xinvert / GeoApps (tests/test_Keff_ocean.py:26-32); the definition is stated in DESIGN.md."""
    phi = np.deg2rad(np.asarray(lat_deg, dtype=np.float64))
    lam = np.deg2rad(np.asarray(lon_deg, dtype=np.float64))
    ny = phi.shape[0]
    jm = np.maximum(np.arange(ny) - 1, 0)
    jp = np.minimum(np.arange(ny) + 1, ny - 1)
    with np.errstate(divide="ignore"):
        cx = 1.0 / ((2.0 * (lam[1] - lam[0])) * (Rearth * np.cos(phi)))
        cy = 1.0 / ((phi[jp] - phi[jm]) * Rearth)
    return cx, cy


def row_metrics_cartesian(y, x):
    """Row metrics of the same stencil on a rectilinear Cartesian / X-Z grid with coordinates y[ny] (rows, may be
    non-uniform or descending) and uniform x[nx]: cx = 1 / (2 dx), cy[j] = 1 / (y[j+1] - y[j-1]) with the
    interior spacing continued at the first / last row (so that every ghost-cell rule of BOUNDARY sees 2*dy)."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    ny = y.shape[0]
    d = np.empty(ny)
    d[1:-1] = y[2:] - y[:-2]
    d[0] = 2.0 * (y[1] - y[0])
    d[-1] = 2.0 * (y[-1] - y[-2])
    with np.errstate(divide="ignore"):
        return np.full(ny, 1.0 / (2.0 * (x[1] - x[0]))), 1.0 / d
