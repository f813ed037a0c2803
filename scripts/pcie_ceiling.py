"""Raw pinned-memory copy ceilings of the box (no kernels): H2D, D2H, both at once, and D2H split over
two streams.  Under torchrun every rank measures its own GPU at the same time (barrier first), so the
N-GPU run gives the shared host-path ceiling that bounds the end-to-end leg of bench.py.
usage: python scripts/pcie_ceiling.py [MB]      (or torchrun --nproc-per-node N scripts/pcie_ceiling.py)"""
import json, os, sys, time
import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = mb << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=6):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps

def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
def d2h_split():
    half = n // 2
    with torch.cuda.stream(s2): h_out[:half].copy_(d_out[:half], non_blocking=True)
    with torch.cuda.stream(s3): h_out[half:].copy_(d_out[half:], non_blocking=True)
def all3(): h2d(); d2h_split()

res = {"rank": rank, "world": world, "MB": mb}
for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("h2d+d2h", both, 2 * n),
                         ("d2h_2streams", d2h_split, n), ("h2d+d2h_2streams", all3, 2 * n)):
    res[name + "_GBps"] = round(nbytes / timed(fn) / 1e9, 2)
print(json.dumps(res), flush=True)
if world > 1: dist.destroy_process_group()
