"""The north-star run: Keff + LWA of BASELINE.json config 4 for >= 8760 slices of 721x1440 on the GPUs of one box,
streamed -- every batch is generated on the device from the global slice index (no input ever exists on the host),
run through the fused batch, and its contour-space results all-gathered on a side stream (the only collective);
LWA fields stay sharded and are reduced to checksums.  Replaces the reference's per-slice Python loops
(xcontour/core.py:1262-1287, :752-794).

    python scripts/run_c4.py [--slices 8760] [--batch 64]                    # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_c4.py --slices 8760

Prints one JSON line (rank 0): slices/s with and without the generation of the inputs, invariants that hold for
every slice (area[-1] = sum dA, CDFs and Q monotone, LWA >= 0), and order-independent checksums that are the same
for any number of GPUs.  Parity of the same streamed path against the oracle: tests/test_gpu_bench_configs.py.
"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xcontour_b200 import ops
from xcontour_b200.pipeline import ContourGather, KeffLwaPlan, slice_range
from xcontour_b200.utils import latlon_cell_area


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slices", type=int, default=8760)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev, pg_options=bench.nccl_options())
    lat, lon = bench.grid()
    dA = latlon_cell_area(lat, lon).astype(np.float32)
    plan = KeffLwaPlan(lat, lon, dA, bench.NLEV, increase=True, lt=True)
    B, S = args.batch, args.slices
    lo, hi = slice_range(S, rank, world)
    per = (S + world - 1) // world                               # every rank runs the same number of batches
    nb = (per + B - 1) // B
    phi = torch.deg2rad(torch.tensor(lat, dtype=torch.float64, device=dev))[:, None]
    lam = torch.deg2rad(torch.tensor(lon, dtype=torch.float64, device=dev))[None, :]
    base = (torch.sin(phi)).float(); amp = (0.3 * torch.cos(phi) ** 2)
    g = torch.Generator(device=dev)

    def generate(qb, first):
        """q[s] = sin(phi) + 0.3 cos^2(phi) sin(6 lam + 3 phi + phase_s) + 0.02 N(0,1), seeded by the GLOBAL slice index."""
        for k in range(B):
            g.manual_seed(1234 + first + k)
            phase = 2 * np.pi * torch.rand((), generator=g, device=dev, dtype=torch.float64)
            noise = torch.randn((bench.NY, bench.NX), generator=g, device=dev, dtype=torch.float32)
            qb[k] = (amp * torch.sin(6 * lam + 3 * phi + phase)).float() + base + 0.02 * noise
    qs = [torch.empty((B, bench.NY, bench.NX), dtype=torch.float32, device=dev) for _ in range(2)]
    outs = [plan.alloc_outputs(B) for _ in range(2)]
    ws = torch.empty(plan.workspace_bytes(B), dtype=torch.uint8, device=dev)
    gather = ContourGather(B, bench.NLEV, dev, nbuf=2) if world > 1 else None
    tot_area = float(dA.astype(np.float64).sum()); max_cell = float(dA.max())
    acc = torch.zeros(8, dtype=torch.float64, device=dev)       # area_last, sum nkeff, sum lwa, min lwa, violations ...
    acc[3] = float("inf")
    t_gen = t_run = 0.0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for b in range(nb):
        i = b % 2
        first = lo + b * B
        n = max(0, min(B, hi - first))                           # valid slices of this batch (ragged tail: padded)
        e[0].record()
        generate(qs[i], min(first, max(S - B, 0)) if n < B else first)
        e[1].record()
        if gather is not None and b >= 2:
            torch.cuda.current_stream().wait_event(gather.event(i))
        plan.run(qs[i], out=outs[i], ws=ws)
        if gather is not None:
            gather.launch(outs[i].packed)
        e[2].record()
        if n > 0:
            o = outs[i]
            acc[0] += o["area"][:n, -1].sum()
            acc[1] += o["nkeff"][:n].nan_to_num().sum()
            acc[2] += o["lwa"][:n].sum()
            acc[3] = torch.minimum(acc[3], o["lwa"][:n].min())
            acc[4] += (o["area"][:n].diff(dim=1) < 0).sum() + (o["intgrdS"][:n].diff(dim=1) < 0).sum() \
                + (o["Qref"][:n].diff(dim=1) < 0).sum()
            # area[-1] = sum dA, up to the cells that hold the slice maximum: the last level is the fp32-rounded
            # end of the linspace and can sit one ulp below the maximum, which then falls outside
            acc[5] += ((o["area"][:n, -1] - tot_area).abs() > 4.0 * max_cell).sum()
            acc[6] = torch.maximum(acc[6], o["lwa"][:n].max())
        e[3].record()
        torch.cuda.synchronize()
        t_gen += e[0].elapsed_time(e[1]); t_run += e[1].elapsed_time(e[2])
    if gather is not None:
        gather.wait()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    red = torch.stack([acc[0], acc[1], acc[2], acc[4], acc[5]])
    mn, mx = acc[3].clone(), acc[6].clone()
    tt = torch.tensor([wall, t_gen * 1e-3, t_run * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        red = red.tolist(); tt = tt.tolist()
        print(json.dumps({
            "run": "C4 streamed", "slices": S, "n_gpus": world, "batch": B, "batches_per_gpu": nb,
            "wall_s": tt[0], "slices_per_s_wall_incl_generation_and_checks": S / tt[0],
            "device_s_fused_batch_plus_gather": tt[2], "slices_per_s_fused_batch_plus_gather": S / tt[2],
            "device_s_generation": tt[1],
            "invariants": {"non_monotone_cdf_or_Q_entries": red[3], "slices_with_area_last_off_by_more_than_4_cells": red[4],
                           "lwa_min": float(mn), "lwa_max": float(mx)},
            "checksums": {"sum_area_last": red[0], "sum_nkeff": red[1], "sum_lwa": red[2]},
            "gather": None if gather is None else {"bytes_per_batch_per_gpu": int(outs[0].packed.nbytes), "batches": nb},
        }), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
