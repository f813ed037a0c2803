"""cProfile of the Contour2D-API end-to-end leg of bench.py (dev tool): where the host time of the drop-in API goes.
usage: python scripts/profile_api.py [slices]"""
import cProfile, io, os, pstats, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xcontour_b200.utils import latlon_cell_area
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lat, lon = bench.grid()
dA = latlon_cell_area(lat, lon).astype(np.float32)
q = np.stack([bench.synth_slice_np(s, lat, lon) for s in range(S)])
grd = np.abs(np.random.default_rng(0).standard_normal(q.shape)).astype(np.float32)
bench.contour2d_e2e(lat, lon, dA, q, grd)            # warm-up (library load, allocator)
pr = cProfile.Profile()
pr.enable()
dt, h2d, d2h = bench.contour2d_e2e(lat, lon, dA, q, grd)
pr.disable()
print("timed call: %.1f ms for %d slices (%.0f slices/s), h2d %.0f MB, d2h %.0f MB" % (dt * 1e3, S, S / dt, h2d / 1e6, d2h / 1e6))
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(45)
print(out.getvalue()[:9000])
