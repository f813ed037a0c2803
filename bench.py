#!/usr/bin/env python
"""
bench.py -- slices/sec for Keff + LWA at 721x1440 (BASELINE.json config 4:
ERA5-scale tracer, 361 equally spaced contours) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (levels -> binning/CDFs with in-flight
|grad q|^2 -> Keff epilogue -> Q(lat) -> LWA, one xc_keff_lwa_batch call) over a
batch of synthetic slices per GPU.  Independent slices shard across ranks with no
data-path collective ("weak" scaling: per-GPU work is fixed).

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how every
field is obtained.  `--impl reference` times the restated reference algorithm
(oracle/, NumPy, all host cores) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

NY, NX, NLEV = 721, 1440, 361
P = NY * NX
ALG_BYTES_PER_SLICE = P * 4 + P * 8          # one fp32 read of q + the fp64 LWA store
STAGE_ALG_BYTES = {                          # per slice, compulsory traffic of each stage
    "minmax_levels": P * 4,                  # read q
    "bin_accumulate": P * 4,                 # read q (dA / edges are L2-resident, shared)
    "lwa": P * 4 + P * 8,                    # read q, write LWA (fp64)
}
METRIC = "keff_lwa_slices_per_sec_721x1440"
# the SAME string in both arms (the driver compares config.workload of `ours` and `reference`)
WORKLOAD = "C4: Keff+LWA, 721x1440 fp32 tracer, 361 contours, increase&lt, |grad q|^2 by centred differences, fp64 LWA out"
# BASELINE.json config 5 (--config c5): histogram / scan stress, Keff part only
C5_NY, C5_NX, C5_NLEV = 4096, 8192, 2048
C5_P = C5_NY * C5_NX
C5_METRIC = "keff_slices_per_sec_4096x8192_N2048"
C5_WORKLOAD = "C5: Keff part (levels, area and |grad q|^2 dA CDFs, d/dA, Keff), 4096x8192 fp32 Cartesian tracer, 2048 contours"


def grid():
    lat = np.linspace(-90.0, 90.0, NY).astype(np.float32)
    lon = (np.arange(NX) * (360.0 / NX)).astype(np.float32)
    return lat, lon


def synth_slice_np(idx, lat, lon):
    """q = sin(phi) + 0.3 cos^2(phi) sin(6 lam + 3 phi + phase) + 0.02 N(0,1)."""
    rng = np.random.default_rng(1234 + idx)
    phi, lam = np.deg2rad(lat.astype(np.float64))[:, None], np.deg2rad(lon.astype(np.float64))[None, :]
    phase = 2 * np.pi * rng.random()
    return (np.sin(phi) + 0.3 * np.cos(phi) ** 2 * np.sin(6 * lam + 3 * phi + phase)
            + 0.02 * rng.standard_normal((NY, NX))).astype(np.float32)


def nccl_options():
    """Default NCCL configuration.  (A one-CTA configuration was tried for the gather and measured at N = 8: the
    collective then holds one SM for ~1.5 ms, and the persistent LWA grid -- one CTA per SM, static tile partition --
    waits for that SM: 330 k instead of 660 k slices/s.  What the gather needs is to be short and rare.)"""
    return None


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ----------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (NumPy restatement), all host cores
# ----------------------------------------------------------------------------
def _cpu_worker(args):
    idx, nrows = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import xcontour_oracle as O
    lat, lon = grid()
    q = synth_slice_np(idx, lat, lon)[None]
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    # static, once per run in the reference workflow too (core.py:156): outside the timed section, like
    # KeffLwaPlan.__init__ on our side
    tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones((NY, NX), np.float32), dA, 0, True, True)
    t0 = time.perf_counter()
    ctr = O.cal_contours(q, NLEV, True)
    grd = O.squared_gradient_latlon(q, lat, lon)
    area = O.cal_integral_within_contours_hist(q, ctr, dA, True)
    intg = O.cal_integral_within_contours_hist(q, ctr, dA, True, integrand=grd)
    latEq = O.table_lookup_coordinates(area, tbl, c)
    with np.errstate(all="ignore"):
        Lmin = O.latitude_lengths_at(latEq)
        Leq2 = O.cal_sqared_equivalent_length(O.cal_gradient_wrt_area(intg, area),
                                              O.cal_gradient_wrt_area(ctr, area))
        O.cal_normalized_Keff(Leq2, Lmin)
    Q = O.interp_to_coords(lat, latEq, ctr)
    t1 = time.perf_counter()
    rows = np.linspace(0, NY - 1, nrows).astype(int).tolist()
    O.cal_local_wave_activity(q, Q, dA, lat, True, rows=rows)      # the reference's j-loop
    t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) * NY / float(nrows)


def cpu_sample(nrows=96, per_core=1, cores=None):
    """One bounded sample: `per_core` slices per core, Keff in full, the LWA j-loop
    on `nrows` of the 721 rows (cost is uniform in j) extrapolated to 721."""
    cores = cores or os.cpu_count() or 1
    n = cores * per_core
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(i, nrows) for i in range(n)])
    wall = time.perf_counter() - t0
    keff = float(np.mean([a for a, _ in res])); lwa = float(np.mean([b for _, b in res]))
    value = cores / (keff + lwa)                                  # all cores busy concurrently
    sample = ("%d slices (%d per core x %d cores), Keff hist path measured in full (%.3f s per slice per core) + "
              "reference LWA j-loop measured on %d of %d rows and scaled x%.2f (%.2f s per slice per core); the static "
              "A(Yeq) table is built once outside the timed section in both arms; sample wall %.1f s"
              % (n, per_core, cores, keff, nrows, NY, NY / float(nrows), lwa, wall))
    return value, cores, sample, {"keff_s_per_slice_per_core": keff, "lwa_s_per_slice_per_core": lwa,
                                  "lwa_rows_measured": nrows, "lwa_rows_total": NY}


def _cpu_worker_c5(idx):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import xcontour_oracle as O
    y, x, q = c5_field_np(idx)
    dA = np.full((C5_NY, C5_NX), 1.0 / C5_P)
    cx, cy = O.row_metrics_cartesian(y, x)
    t0 = time.perf_counter()
    ctr = O.cal_contours(q[None], C5_NLEV, True)
    grd = O.squared_gradient(q[None], cx, cy, "periodic", "extend")
    area = O.cal_integral_within_contours_hist(q[None], ctr, dA, True)
    intg = O.cal_integral_within_contours_hist(q[None], ctr, dA, True, integrand=grd)
    with np.errstate(all="ignore"):
        O.cal_sqared_equivalent_length(O.cal_gradient_wrt_area(intg, area), O.cal_gradient_wrt_area(ctr, area))
    return time.perf_counter() - t0


def cpu_sample_c5(cores=None):
    cores = min(cores or os.cpu_count() or 1, 8)                  # ~1.5 GB of temporaries per worker
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker_c5, list(range(cores)))
    wall = time.perf_counter() - t0
    per = float(np.mean(res))
    return cores / per, cores, ("%d slices (1 per core x %d cores), Keff hist path in full: %.1f s per slice per core; "
                                "sample wall %.1f s" % (cores, cores, per, wall)), {"keff_s_per_slice_per_core": per}


def c5_field_np(idx):
    """q = y + 0.2 sin(8 pi x) sin(4 pi y) + 0.01 N(0,1) on the unit square, seed 4321 + slice."""
    y = (np.arange(C5_NY) + 0.5) / C5_NY
    x = (np.arange(C5_NX) + 0.5) / C5_NX
    rng = np.random.default_rng(4321 + idx)
    q = (y[:, None] + 0.2 * np.sin(8 * np.pi * x)[None, :] * np.sin(4 * np.pi * y)[:, None]
         + 0.01 * rng.standard_normal((C5_NY, C5_NX))).astype(np.float32)
    return y, x, q


def run_reference(args, rank):
    if rank != 0:
        return
    c5 = args.config == "c5"
    vals, t_all = [], time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, cores, sample, legs = cpu_sample_c5() if c5 else cpu_sample(nrows=96, per_core=1)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": C5_METRIC if c5 else METRIC, "value": value, "unit": "slices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(1, args.warmup + args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": C5_WORKLOAD if c5 else WORKLOAD, "impl_note":
                   "NumPy restatement of xcontour's hist path + LWA j-loop (oracle/, bit-identical to the reference's "
                   "own code on the refshim stand-ins; the reference itself is not importable here: no xarray/xhistogram)"},
        "cpu_baseline": dict({"value": value, "unit": "slices/s", "cores": cores, "kind": "port", "sample": sample}, **legs),
        "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.p = gpu_index, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        load = [x for x in sm if smax and x > 0.5 * smax] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def contour2d_e2e(lat, lon, dA, q_np, grd_np):
    """The same work through the drop-in Contour2D / Table API (what a user of the reference calls), host arrays in,
    host arrays out: tests/test_Keff_atmos.py:76-92 + tests/test_LWA.py:57-77 on a (time, lat, lon) stack with the
    squared gradient given as a field (as the reference's callers have it).  Returns (seconds, h2d bytes, d2h bytes)."""
    import torch
    import xcontour_b200 as xb
    coords2 = {"latitude": lat, "longitude": lon}
    coords3 = dict(coords2, time=np.arange(q_np.shape[0]))
    dims3 = ("time", "latitude", "longitude")
    dAx = xb.DataArray(dA, dims=("latitude", "longitude"), coords=coords2, name="dA")
    mask = xb.DataArray(np.ones_like(dA), dims=("latitude", "longitude"), coords=coords2)
    latx = xb.DataArray(lat, dims=("latitude",), coords={"latitude": lat})

    def once():
        tr = xb.DataArray(q_np, dims=dims3, coords=coords3, name="pv")
        gx = xb.DataArray(grd_np, dims=dims3, coords=coords3, name="grdS")
        an = xb.Contour2D(tr, dAx, dims={"X": "longitude", "Y": "latitude"}, dimEq={"Y": "latitude"},
                          increase=True, lt=True)
        table = an.cal_area_eqCoord_table_hist(mask)
        ctr = an.cal_contours(NLEV)
        area = an.cal_integral_within_contours_hist(ctr)
        intg = an.cal_integral_within_contours_hist(ctr, integrand=gx)
        latEq = table.lookup_coordinates(area)
        Lmin = xb.latitude_lengths_at(latEq)
        dint = an.cal_gradient_wrt_area(intg, area)
        dq = an.cal_gradient_wrt_area(ctr, area)
        nk = an.cal_normalized_Keff(an.cal_sqared_equivalent_length(dint, dq), Lmin)
        Q = an.interp_to_coords(latx, latEq, ctr)
        lwa = an.cal_local_wave_activity(tr, Q)
        return [ctr, area, intg, latEq, Lmin, dint, dq, nk, Q, lwa]
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, q_np.nbytes + grd_np.nbytes, int(sum(np.asarray(r.values).nbytes for r in res))


def run_ours_c5(args, rank, world, local_rank):
    """--config c5: Keff part of BASELINE.json config 5 (4096x8192 Cartesian tracer, 2048 contours)."""
    import torch
    import torch.distributed as dist
    from xcontour_b200 import ops
    from xcontour_b200._lib import N_STAGES, STAGE_NAMES
    from xcontour_b200.pipeline import KeffLwaPlan
    from xcontour_b200.utils import row_metrics_cartesian
    ops.require_cuda()                       # no CUDA device or no built library: stop here, there is no CPU path
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B = args.batch if args.batch != 64 else 4
    y = (np.arange(C5_NY) + 0.5) / C5_NY
    x = (np.arange(C5_NX) + 0.5) / C5_NX
    dA = np.full((C5_NY, C5_NX), 1.0 / C5_P)
    plan = KeffLwaPlan(y, x, dA, C5_NLEV, increase=True, lt=True, metrics=row_metrics_cartesian(y, x),
                       boundary=("periodic", "extend"), sub_batch=args.sub_batch)
    yt = torch.tensor(y, dtype=torch.float64, device=dev)[:, None]
    xt = torch.tensor(x, dtype=torch.float64, device=dev)[None, :]
    q = torch.empty((B, C5_NY, C5_NX), dtype=torch.float32, device=dev)
    for s in range(B):
        g = torch.Generator(device=dev); g.manual_seed(4321 + rank * B + s)
        q[s] = (yt + 0.2 * torch.sin(8 * np.pi * xt) * torch.sin(4 * np.pi * yt)).float() \
            + 0.01 * torch.randn((C5_NY, C5_NX), generator=g, device=dev, dtype=torch.float32)
    out = plan.alloc_outputs(B, lwa=False)
    ws = torch.empty(plan.workspace_bytes(B), dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_w, n_w = time.perf_counter(), 0
    while n_w < args.warmup or time.perf_counter() - t_w < 0.6:
        plan.run(q, out=out, ws=ws); torch.cuda.synchronize(); n_w += 1
    barrier()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.run(q, out=out, ws=ws)
    e1.record()
    barrier()
    launches = ops.launch_count()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_max * 1e-3)
    stage = (ctypes.c_float * N_STAGES)(); acc = np.zeros(N_STAGES)
    for _ in range(args.steps):
        plan.run(q, out=out, ws=ws, stage_ms=stage); acc += np.array(list(stage))
    stages = {n: float(v) / args.steps for n, v in zip(STAGE_NAMES, acc)}
    # end to end: pinned host slices in, contour-space results out
    q_host = torch.empty((B, C5_NY, C5_NX), dtype=torch.float32).pin_memory(); q_host.copy_(q)
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
    qd = torch.empty_like(q)

    def e2e_once():
        qd.copy_(q_host, non_blocking=True)
        plan.run(qd, out=out, ws=ws)
        for k, v in host_out.items():
            v.copy_(out[k], non_blocking=True)
    e2e_once(); barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e = max(1, min(args.steps, 3))
    g0.record()
    for _ in range(n_e):
        e2e_once()
    g1.record(); torch.cuda.synchronize()
    chk = float(host_out["area"][:, -1].sum())
    te = torch.tensor([g0.elapsed_time(g1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = peak_hbm()
        alg = {"minmax_levels": C5_P * 4, "bin_accumulate": C5_P * 4}
        dom = max(alg, key=lambda k: stages[k])
        ach = alg[dom] * B / (stages[dom] * 1e-3) / 1e9
        line = {
            "metric": C5_METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": C5_WORKLOAD, "slices_per_step_per_gpu": B,
                       "l2": "inputs larger than L2 (%.0f MB of q per step)" % (B * C5_P * 4 / 1e6),
                       "parallelism": "slices sharded over %d GPU(s), no data-path collective" % world},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "peak_source": peak_src, "stage_ms_per_step": stages,
                         "alg_bytes_per_launch": alg[dom] * min(B, plan.sub_batch or 1),
                         "pipeline": {"alg_bytes_per_slice": 2 * C5_P * 4, "achieved": 2 * C5_P * 4 * value / world / 1e9,
                                      "frac": 2 * C5_P * 4 * value / world / 1e9 / peak}},
            "clocks": clocks,
            "e2e": {"value": world * B * n_e / float(te.item()), "unit": "slices/s", "h2d_bytes_per_step": int(q_host.nbytes),
                    "d2h_bytes_per_step": int(sum(v.nbytes for v in host_out.values())), "checksum": chk},
            "gpu_launches": int(launches) * world,
        }
        if not args.no_cpu:
            v, cores, sample, legs = cpu_sample_c5()
            line["cpu_baseline"] = dict({"value": v, "unit": "slices/s", "cores": cores, "kind": "port", "sample": sample}, **legs)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from xcontour_b200 import ops
    from xcontour_b200._lib import N_STAGES, STAGE_NAMES
    from xcontour_b200.pipeline import ContourGather, HostStreamer, KeffLwaPlan, bind_host_thread_to_gpu
    from xcontour_b200.utils import latlon_cell_area

    ops.require_cuda()                       # no CUDA device or no built library: stop here, there is no CPU path
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=nccl_options())
    dev = torch.device("cuda", local_rank)
    lat, lon = grid()
    dA = latlon_cell_area(lat, lon).astype(np.float32)
    plan = KeffLwaPlan(lat, lon, dA, NLEV, increase=True, lt=True, sub_batch=args.sub_batch)
    B = args.batch

    # synthetic slices of this rank (global index = rank*B + s), generated on the device
    phi = torch.deg2rad(torch.tensor(lat, dtype=torch.float64, device=dev))[:, None]
    lam = torch.deg2rad(torch.tensor(lon, dtype=torch.float64, device=dev))[None, :]
    q = torch.empty((B, NY, NX), dtype=torch.float32, device=dev)
    for s in range(B):
        g = torch.Generator(device=dev); g.manual_seed(1234 + rank * B + s)
        phase = 2 * np.pi * torch.rand((), generator=g, device=dev, dtype=torch.float64)
        noise = torch.randn((NY, NX), generator=g, device=dev, dtype=torch.float32)
        q[s] = (torch.sin(phi) + 0.3 * torch.cos(phi) ** 2 * torch.sin(6 * lam + 3 * phi + phase)).float() + 0.02 * noise
    # Contour-space results are gathered across ranks every G steps (the only collective of the path):
    # G consecutive steps write their [9, B, N] results into slabs of one packed buffer [9, G*B, N]; two such buffers
    # alternate, so the all-gather of one (NCCL, side stream) overlaps the next G steps.  Default G = 1, every step's
    # results are gathered as soon as they exist.  The NCCL kernel needs SMs, and the persistent LWA grid (one CTA
    # per SM, all registers) cannot start on an SM that is busy with it, so a collective costs about its own duration
    # once per G steps; measured (profiles/r2_gather_cadence.txt): G = 1 and G = 5 are within 1 % of each other at
    # N = 4 (0.785 / 0.790 ms per step) and G = 1 is the better one at N = 8 (0.784 / 0.803), where the larger
    # trailing collective of G = 5 is exposed at the end of the timed region.
    G = max(1, args.gather_every) if world > 1 else 1
    nset = 2 if world > 1 else 1
    packs = [torch.empty((9, G * B, NLEV), dtype=torch.float64, device=dev) for _ in range(nset)]
    lwa_buf = torch.empty((B, NY, NX), dtype=torch.float64, device=dev)
    qref_buf = torch.empty((B, NY), dtype=torch.float64, device=dev)

    def slab_outputs(pack, k):
        from xcontour_b200.pipeline import CONTOUR_VARS, Outputs
        o = Outputs((name, pack[i, k * B:(k + 1) * B]) for i, name in enumerate(CONTOUR_VARS))
        o["Qref"], o["lwa"] = qref_buf, lwa_buf
        return o
    outs = [[slab_outputs(p, k) for k in range(G)] for p in packs]
    out = outs[0][0]
    ws = torch.empty(plan.workspace_bytes(B), dtype=torch.uint8, device=dev)
    gather = ContourGather(G * B, NLEV, dev, nbuf=2) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pack_ev = [None] * nset                                  # completion of the last gather that read pack g
    last = {"recv": None, "pack": 0}

    def launch_gather(g):
        recv, idx = gather.launch(packs[g])
        pack_ev[g] = gather.event(idx)
        last["recv"], last["pack"] = recv, g

    def step(i):
        g, k = (i // G) % nset, i % G
        if gather is not None and k == 0 and pack_ev[g] is not None:     # the gather that last read this buffer must be done
            torch.cuda.current_stream().wait_event(pack_ev[g])
        plan.run(q, out=outs[g][k], ws=ws)
        if gather is not None and k == G - 1:
            launch_gather(g)

    def timed_region():
        """W warm-up steps, then exactly K timed steps between barriers; returns
        (max-over-ranks ms, launches, clock record of rank 0)."""
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()                  # keeps sampling through warm-up and the timed region
        t_w = time.perf_counter()
        n_w = 0
        while n_w < args.warmup or (world == 1 and time.perf_counter() - t_w < 0.6):    # >= W steps, and nvidia-smi gets samples
            step(n_w)
            torch.cuda.synchronize()
            n_w += 1
        if gather is not None:               # the collective is warmed up too (NCCL connects lazily on first use)
            launch_gather(0)
            gather.wait()
            torch.cuda.synchronize()
        barrier()
        ops.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        if gather is not None:
            if args.steps % G:               # a trailing partial group is gathered too
                launch_gather(((args.steps - 1) // G) % nset)
            gather.wait()                    # the timed region ends when the last gather has landed
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ops.launch_count()
        if rank == 0 and world == 1:         # a few more steps so the 100 ms sampler sees the loaded clocks
            t_c = time.perf_counter()
            while time.perf_counter() - t_c < 0.5:
                plan.run(q, out=out, ws=ws)
                torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, clocks

    def clocks_bad(c):
        if not c or c.get("sm_mhz") is None:
            return False
        slow = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
        pinned = c["sm_max_mhz"] and c["sm_mhz"] < 0.7 * c["sm_max_mhz"] and not c["reasons"]
        return bool(slow) or bool(pinned)

    ms_max, launches, clocks = timed_region()
    redo = torch.tensor([1 if (rank == 0 and clocks_bad(clocks)) else 0], device=dev)
    if world > 1:
        dist.all_reduce(redo, op=dist.ReduceOp.MAX)
    if int(redo.item()):                     # throttled or clock-locked sample: measure once more
        first = clocks
        ms_max, launches, clocks = timed_region()
        if rank == 0:
            clocks["remeasured_after"] = first
    value = world * B * args.steps / (ms_max * 1e-3)
    gather_check = None
    if gather is not None:                   # the gathered block of this rank is what it computed
        torch.cuda.synchronize()
        gather_check = bool(torch.equal(last["recv"][rank].nan_to_num(), packs[last["pack"]].nan_to_num()))

    # per-stage device time, CUDA events on the launching stream inside the same call
    stage = (ctypes.c_float * N_STAGES)()
    acc = np.zeros(N_STAGES)
    for _ in range(args.steps):
        plan.run(q, out=out, ws=ws, stage_ms=stage)
        acc += np.array(list(stage))
    acc /= args.steps                                             # ms per step, per stage
    stages = {n: float(v) for n, v in zip(STAGE_NAMES, acc)}

    # end to end: pinned host slices -> H2D -> fused batch -> D2H of every result
    eb = min(args.e2e_batch, B)
    # pinned buffers are first-touched on the GPU's own NUMA node (undone after this leg)
    prev_aff, numa_note = (None, "off") if args.no_numa else bind_host_thread_to_gpu(local_rank)
    streamer = HostStreamer(plan, eb, copy_lwa=True, nbuf=args.e2e_nbuf)
    q_host = torch.empty((B, NY, NX), dtype=torch.float32).pin_memory()
    q_host.copy_(q)
    sink = {"n": 0, "chk": 0.0}

    def consume(s0, s1, host):
        sink["n"] += s1 - s0
        sink["chk"] += float(host["area"][:, -1].sum())           # a result actually read on the host

    streamer.run(q_host, consume)                                  # warm-up
    barrier()
    streamer.h2d_bytes = streamer.d2h_bytes = 0
    e2e_steps = max(1, min(args.steps, 3))
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    g0.record(cur)
    for st in streamer.streams:              # the copy/compute streams start after g0 ...
        st.wait_event(g0)
    for _ in range(e2e_steps):
        streamer.run(q_host, consume)
    for st in streamer.streams:              # ... and g1 is recorded after both have drained
        cur.wait_stream(st)
    g1.record(cur)
    torch.cuda.synchronize()
    te = torch.tensor([g0.elapsed_time(g1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(te.item())
    if prev_aff is not None:
        os.sched_setaffinity(0, prev_aff)                        # the cpu_baseline leg uses every core
    # the same through the drop-in Contour2D API (rank 0; a user-facing number, not the headline)
    c2d = None
    if rank == 0 and not args.no_api:
        nb = min(8, B)
        try:                                 # a reporting leg beside the headline: its failure is reported, not fatal
            qn = q_host[:nb].numpy()
            from xcontour_b200 import ops as _ops
            grd = _ops.grad2_latlon(q[:nb].contiguous(), plan.lat_rad, plan.dlambda, out_dtype=torch.float32).cpu().numpy()
            dt, hb, db = contour2d_e2e(lat, lon, dA, qn, grd)
            c2d = {"value": nb / dt, "unit": "slices/s", "slices": nb, "h2d_bytes": int(hb), "d2h_bytes": int(db),
                   "note": "Contour2D/Table API of the reference, numpy in / labelled numpy out, |grad q|^2 given as an fp32 field"}
        except Exception as e:
            c2d = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        peak, peak_src = peak_hbm()
        dom = max(STAGE_ALG_BYTES, key=lambda k: stages[k])
        ach = STAGE_ALG_BYTES[dom] * B / (stages[dom] * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        sub = plan.sub_batch or 32
        line = {
            "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "slices_per_step_per_gpu": B, "sub_batch": args.sub_batch or "auto",
                       "l2": "inputs larger than L2 (%.0f MB of q + %.0f MB of LWA per step)"
                             % (B * P * 4 / 1e6, B * P * 8 / 1e6),
                       "parallelism": ("slices sharded over %d GPU(s); " % world) +
                                      ("no collective at N=1" if world == 1 else
                                       "the [9, B, N] contour-space results of every step are all-gathered (NCCL, side stream, "
                                       "one collective per %s) INSIDE the timed region" % ("step" if G == 1 else "%d steps" % G))},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic,
                         "traffic_note": "ncu dram__bytes_read+write of one launch = one pass of 32 slices "
                                         "(profiles/traffic.json); algorithmic bytes of the same launch in alg_bytes_per_launch",
                         "alg_bytes_per_launch": STAGE_ALG_BYTES[dom] * min(B, sub),
                         "peak_source": peak_src,
                         "stage_ms_per_step": stages,
                         "stage_frac": {k: STAGE_ALG_BYTES[k] * B / (stages[k] * 1e-3) / 1e9 / peak for k in STAGE_ALG_BYTES},
                         "pipeline": {"alg_bytes_per_slice": ALG_BYTES_PER_SLICE,
                                      "achieved": ALG_BYTES_PER_SLICE * value / world / 1e9,
                                      "frac": ALG_BYTES_PER_SLICE * value / world / 1e9 / peak}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "slices/s",
                    "h2d_bytes_per_step": streamer.h2d_bytes // e2e_steps,
                    "d2h_bytes_per_step": streamer.d2h_bytes // e2e_steps,
                    "timing": "CUDA events spanning pinned H2D + kernels + D2H on both streams, max over ranks",
                    "batch": eb, "buffers_in_flight": args.e2e_nbuf, "host_numa": numa_note,
                    "checksum": sink["chk"]},
            "gpu_launches": int(launches) * world,
        }
        if c2d is not None:
            line["e2e_contour2d"] = c2d
        if gather is not None:
            line["gather"] = {"in_timed_region": True, "bytes_per_step_per_gpu": int(packs[0].nbytes) // G,
                              "steps_per_collective": G, "collectives_in_timed_region": (args.steps + G - 1) // G,
                              "own_block_bit_identical": gather_check}
        if not args.no_cpu:
            try:                             # the CPU baseline is reported beside the measurement: it must not lose the line
                v, cores, sample, legs = cpu_sample(nrows=96, per_core=3 if world == 1 else 1)
                line["cpu_baseline"] = dict({"value": v, "unit": "slices/s", "cores": cores, "kind": "port",
                                             "sample": sample}, **legs)
            except Exception as e:
                line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="slices per step per GPU")
    ap.add_argument("--sub-batch", type=int, default=0, help="slices per internal pass (0 = auto)")
    ap.add_argument("--e2e-batch", type=int, default=8)
    ap.add_argument("--e2e-nbuf", type=int, default=2, help="batches in flight in the end-to-end leg")
    ap.add_argument("--gather-every", type=int, default=1, help="steps per all-gather of the contour-space results (N > 1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-api", action="store_true", help="skip the Contour2D-API end-to-end leg")
    ap.add_argument("--config", default="c4", choices=["c4", "c5"], help="BASELINE.json config 4 (default, the headline) or 5")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the end-to-end leg's host thread to the GPU's NUMA node")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.config == "c5":
        run_ours_c5(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
