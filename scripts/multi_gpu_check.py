"""N-GPU check of the sharded path (torchrun, NCCL): every rank processes its
slice_range of a common synthetic stack, the contour-space results are
all-gathered, and rank 0 verifies them bit-for-bit against a single-GPU pass over
the whole stack (partitioning must not change any result).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multi_gpu_check.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcontour_b200.utils import latlon_cell_area
from xcontour_b200 import ops
from xcontour_b200.pipeline import KeffLwaPlan, slice_range, gather_contour_space, CONTOUR_VARS
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ny, nx, N, S = 181, 360, 91, 13
lat = np.linspace(-90, 90, ny).astype(np.float32); lon = (np.arange(nx) * (360.0 / nx)).astype(np.float32)
rng = np.random.default_rng(5)
phi, lam = np.deg2rad(lat)[:, None], np.deg2rad(lon)[None, :]
q = np.stack([np.sin(phi) + 0.3 * np.cos(phi) ** 2 * np.sin(6 * lam + s) + 0.02 * rng.standard_normal((ny, nx))
              for s in range(S)]).astype(np.float32)
dA = latlon_cell_area(lat, lon).astype(np.float32)
plan = KeffLwaPlan(lat, lon, dA, N)
lo, hi = slice_range(S, rank, world)
out = plan.run(ops.to_dev(q[lo:hi])) if hi > lo else {k: torch.empty((0, N), dtype=torch.float64, device="cuda") for k in CONTOUR_VARS}
full = gather_contour_space({k: out[k] for k in CONTOUR_VARS}, S)
if rank == 0:
    ref = plan.run(ops.to_dev(q))
    torch.cuda.synchronize()
    for k in CONTOUR_VARS:
        assert torch.equal(full[k].nan_to_num(), ref[k].nan_to_num()), k
    print("multi-GPU check ok: world=%d, %d slices, %d contour-space arrays identical to the single-GPU pass" % (world, S, len(CONTOUR_VARS)))
dist.barrier(); dist.destroy_process_group()
