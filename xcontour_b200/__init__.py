"""
xcontour_b200 -- the contour-coordinate hot path of miniufo/xcontour on B200.

Same public surface as ``xcontour`` for this path (xcontour/__init__.py:2-6):
Contour2D, Table, equivalent_latitudes, latitude_lengths_at.  Everything
numerical runs in hand-written sm_100a kernels behind the C ABI of
``include/xcb200.h`` (libxcb200.so, built in-tree by ``xcontour_b200.build``).
"""
from .core import Contour2D, Table
from .utils import equivalent_latitudes, latitude_lengths_at, latlon_cell_area
from .xr_compat import DataArray, Dataset, merge, HAVE_XARRAY

__version__ = "0.1.0"
__all__ = ["Contour2D", "Table", "equivalent_latitudes", "latitude_lengths_at", "latlon_cell_area",
           "DataArray", "Dataset", "merge"]
