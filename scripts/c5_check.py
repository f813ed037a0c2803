"""BASELINE config 5 (high-res ocean tracer 4096x8192, 2048 contour levels; Keff
part only = histogram/scan stress): run one slice through the C ABI, check the
CDFs against the oracle, time the binning kernel."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xcontour_oracle as O
from xcontour_b200 import ops
from xcontour_b200._lib import SCAN_PREFIX
ny, nx, N = 4096, 8192, 2048
rng = np.random.default_rng(4321)
y = (np.arange(ny) + 0.5) / ny; x = (np.arange(nx) + 0.5) / nx
q = (y[:, None] + 0.2 * np.sin(8 * np.pi * x)[None, :] * np.sin(4 * np.pi * y)[:, None]
     + 0.01 * rng.standard_normal((ny, nx))).astype(np.float32)[None]
dA = np.full((ny, nx), 1.0 / (ny * nx))
g = rng.random((1, ny, nx)).astype(np.float32)
qd, dAd, gd = ops.to_dev(q.reshape(1, -1)), ops.to_dev(dA.reshape(-1)), ops.to_dev(g.reshape(1, -1))
lv, _ = ops.minmax_levels(qd, N, True, 0)
e, d = ops.hist_edges(lv, 0, True)
for _ in range(2):
    cdf, _, _ = ops.bin_accumulate(qd, e, dAd, acc_area=True, integrands=[gd], decreasing=d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    lv, _ = ops.minmax_levels(qd, N, True, 0)
    e, d = ops.hist_edges(lv, 0, True)
    cdf, _, _ = ops.bin_accumulate(qd, e, dAd, acc_area=True, integrands=[gd], decreasing=d)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ctr = O.cal_contours(q, N, True)
assert np.array_equal(lv.cpu().numpy().astype(np.float32), ctr)
t = time.time()
ref_a = O.cal_integral_within_contours_hist(q, ctr, dA, True)
ref_g = O.cal_integral_within_contours_hist(q, ctr, dA, True, integrand=g)
tcpu = time.time() - t
c = cdf.cpu().numpy()
ea = np.abs(c[:, 0] - ref_a).max() / ref_a.max(); eg = np.abs(c[:, 1] - ref_g).max() / ref_g.max()
print("C5 1 slice 4096x8192 N=2048: levels bit-exact, area relerr %.2e, intg relerr %.2e; GPU %.2f ms/slice "
      "(%.1f slices/s, %.0f GB/s of the 2x134 MB algorithmic traffic); oracle (1 core) %.1f s"
      % (ea, eg, ms, 1e3 / ms, 2 * q.nbytes / ms / 1e6, tcpu))
assert ea < 1e-12 and eg < 1e-12
