"""Quick per-stage timing of xc_keff_lwa_batch on synthetic C4 slices (dev tool).
usage: [XCB200_LIB=...] python scripts/time_stages.py [batch] [sub]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xcontour_b200.utils import latlon_cell_area
from xcontour_b200._lib import N_STAGES, STAGE_NAMES
from xcontour_b200.pipeline import KeffLwaPlan
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lat, lon = bench.grid()
dA = latlon_cell_area(lat, lon).astype(np.float32)
plan = KeffLwaPlan(lat, lon, dA, bench.NLEV, sub_batch=sub)
g = torch.Generator(device="cuda"); g.manual_seed(1)
phi = torch.deg2rad(torch.tensor(lat, dtype=torch.float64, device="cuda"))[:, None]
lam = torch.deg2rad(torch.tensor(lon, dtype=torch.float64, device="cuda"))[None, :]
q = torch.empty((B, bench.NY, bench.NX), dtype=torch.float32, device="cuda")
for s in range(B):
    q[s] = (torch.sin(phi) + 0.3 * torch.cos(phi) ** 2 * torch.sin(6 * lam + 3 * phi + s)).float() \
        + float(os.environ.get("XC_NOISE", "0.02")) * torch.randn((bench.NY, bench.NX), generator=g, device="cuda")
if os.environ.get("XC_QUANT"):      # pathological: large patches of identical values
    k = float(os.environ["XC_QUANT"]); q = torch.round(q * k) / k
out = plan.alloc_outputs(B, lwa=not os.environ.get("XC_NO_LWA"))
for _ in range(3):
    plan.run(q, out=out)
torch.cuda.synchronize()
st = (ctypes.c_float * N_STAGES)(); acc = np.zeros(N_STAGES)
n = 5
for _ in range(n):
    plan.run(q, out=out, stage_ms=st); acc += np.array(list(st))
acc /= n
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    plan.run(q, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("%-10s total %.3f ms/%d slices = %.1f us/slice (%.0f slices/s) | " % (
    os.path.basename(os.environ.get("XCB200_LIB", "default")).replace("libxcb200_", "").replace(".so", ""),
    ms, B, 1e3 * ms / B, B / ms * 1e3) + "  ".join("%s %.3f" % (k[:6], v) for k, v in zip(STAGE_NAMES, acc)))
