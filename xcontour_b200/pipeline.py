"""
Batched Keff + LWA over many independent slices -- the workload of
BASELINE.json (config 4: 721x1440 tracer, 361 contours) -- and its sharding over
the GPUs of one box.

The per-slice call chain is exactly the reference workflow of
tests/test_Keff_atmos.py:76-92 + tests/test_LWA.py:57-77:

    ctr   = cal_contours(N)                              core.py:205
    table = cal_area_eqCoord_table_hist(mask)            core.py:150   (once)
    area, intgrdS = cal_integral_within_contours_hist    core.py:412
    latEq = table.lookup_coordinates(area)               core.py:1136
    Lmin, dintSdA, dqdA, Leq2, nkeff                     utils.py:518, core.py:463-966
    Q     = interp_to_coords(lat, latEq, ctr)            core.py:1050
    LWA   = cal_local_wave_activity(q, Q)                core.py:696

run on the device for a whole batch with one C-ABI call (xc_keff_lwa_batch).
Slices are independent, so multi-GPU = contiguous slice ranges per rank and one
gather of the small contour-space results; the LWA fields stay
sharded.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from .utils import BOUNDARY, row_metrics_latlon, scalar_rules as _scalar_rules
from ._lib import KeffLwaArgs, PART, SCAN_PREFIX, SCAN_TOTAL_MINUS, XC_F32, XC_F64, check

CONTOUR_VARS = ("ctr", "area", "intgrdS", "latEq", "Lmin", "dintSdA", "dqdA", "Leq2", "nkeff")


class Outputs(dict):
    """{name: device tensor} of one batch; ``packed`` is the [9, S, N] buffer the contour-space entries are slabs of."""
    packed = None


def slice_range(S, rank, world):
    """Contiguous block of ceil(S/world) slice indices owned by ``rank``."""
    per = (S + world - 1) // world
    lo = min(S, rank * per)
    return lo, min(S, lo + per)


class KeffLwaPlan(object):
    """Static (time-independent) part of the workflow: grid metrics, the A(Yeq)
    table, LWA weights -- built once, reused for every batch."""

    def __init__(self, lat_deg, lon_deg, dA, N, increase=True, lt=True,
                 dtype=np.float32, keff_mask=1e5, part="all", mask=None, sub_batch=0,
                 metrics=None, boundary=("periodic", "extend"), fill_value=0.0, scalar_rules=None):
        """lat_deg / lon_deg: the equivalent (row) coordinate and the column coordinate of the plane.
        metrics: (cx[ny], cy[ny]) row metrics of the |grad q|^2 stencil (utils.row_metrics_cartesian for
        Cartesian / X-Z planes); default: the lat-lon metrics of utils.row_metrics_latlon.
        boundary: ghost-cell rule along (x, y), each one of utils.BOUNDARY; fill_value for 'fill'.
        scalar_rules: 'numpy1' | 'numpy2', the NumPy regime of the reference's per-'time' bin edges that is
        reproduced (utils.NUMPY_SCALAR_RULES: the installed NumPy's by default)."""
        ops.require_cuda()
        lat = np.asarray(lat_deg)
        lon = np.asarray(lon_deg)
        self.ny, self.nx, self.N = lat.shape[0], lon.shape[0], int(N)
        self.increase, self.lt = bool(increase), bool(lt)
        self.ctr_dtype = XC_F32 if np.dtype(dtype) == np.float32 else XC_F64
        self.keff_mask, self.part = float(keff_mask), PART[part]
        self.sub_batch = int(sub_batch)
        self.numpy2_rules = _scalar_rules(scalar_rules) == "numpy2"
        dA = np.ascontiguousarray(np.broadcast_to(np.asarray(dA), (self.ny, self.nx)))
        self.dA = ops.to_dev(dA)
        self.ww = ops.lwa_weights(self.dA.reshape(-1))
        self.ww_row = ops.row_constant(self.ww, self.ny, self.nx)
        # A(Yeq) table exactly as Contour2D.cal_area_eqCoord_table_hist builds it
        fdt = lat.dtype if lat.dtype in (np.float32, np.float64) else np.float64
        ctrVar = np.ascontiguousarray(np.broadcast_to(lat.astype(fdt)[:, None], (self.ny, self.nx)))
        yIncre = not (lat[-1] < lat[0])
        ylt = self.lt if self.increase == yIncre else (not self.lt)
        edges, _ = ops.hist_edges(ops.to_dev(lat.astype(np.float64).reshape(1, -1)),
                                  XC_F32 if fdt == np.float32 else XC_F64, time_branch=False)
        m = np.ones((self.ny, self.nx), np.uint8) if mask is None else (np.asarray(mask) == 1).astype(np.uint8)
        cdf, _, _ = ops.bin_accumulate(ops.to_dev(ctrVar).reshape(1, -1), edges[0], self.dA.reshape(-1),
                                       acc_area=True, q_mask=ops.to_dev(m.reshape(-1)),
                                       scan_mode=SCAN_PREFIX if ylt else SCAN_TOTAL_MINUS)
        self.table = cdf[0, 0].contiguous()
        self.table_coord = ops.to_dev((lat if yIncre else lat[::-1]).astype(np.float64))
        # coordinates Q is interpolated to: tracer.latitude.astype(dtype) (tests/test_LWA.py:72)
        self.eq_coord = ops.to_dev(lat.astype(dtype).astype(np.float64))
        self.lat_rad = ops.to_dev(np.deg2rad(lat.astype(np.float64)))
        lam = np.deg2rad(lon.astype(np.float64))
        self.dlambda = float(lam[1] - lam[0])
        # stencil metrics on the host (ny values each) and the hints that select the fast kernels
        cx, cy = row_metrics_latlon(lat, lon) if metrics is None else (np.asarray(m, np.float64) for m in metrics)
        self.cx, self.cy = ops.to_dev(np.ascontiguousarray(cx)), ops.to_dev(np.ascontiguousarray(cy))
        self.bcx, self.bcy = BOUNDARY[boundary[0]], BOUNDARY[boundary[1]]
        self.fill_value = float(fill_value)
        with np.errstate(invalid="ignore", over="ignore"):
            self.any_degenerate = bool(np.any(~np.isfinite(cx) | ~np.isfinite(cy) | ((cx * cx > 2.0 ** 30 * cy * cy) & (cy != 0))))
        row_const = bool(np.all((dA == dA[:, :1]) | np.isnan(dA)))
        self.dA_row = ops.to_dev(np.ascontiguousarray(dA[:, 0], dtype=np.float64)) if row_const else None
        self.uniform_dA = bool(row_const and np.all(dA[:, 0] == dA[0, 0]))

    def workspace_bytes(self, S):
        return _lib.load().xc_keff_lwa_batch_workspace_bytes(S, self.ny, self.nx, self.N)

    def alloc_outputs(self, S, lwa=True, lwa_dtype=torch.float64):
        """Output tensors of one batch.  The nine contour-space results are slabs of ONE buffer
        ``out.packed`` [9, S, N] (CONTOUR_VARS order), so that gathering them across ranks is a single
        collective on a single buffer with no packing copies (ContourGather)."""
        dev = self.dA.device
        packed = torch.empty((len(CONTOUR_VARS), S, self.N), dtype=torch.float64, device=dev)
        out = Outputs((k, packed[i]) for i, k in enumerate(CONTOUR_VARS))
        out.packed = packed
        out["Qref"] = torch.empty((S, self.ny), dtype=torch.float64, device=dev)
        if lwa:
            # lwa_dtype=torch.float32 is an opt-in that is NOT a drop-in result (the fp64 field rounded once)
            out["lwa"] = torch.empty((S, self.ny, self.nx), dtype=lwa_dtype, device=dev)
        return out

    def run(self, q, grdS=None, out=None, ws=None, stage_ms=None):
        """q[S, ny, nx] (fp32/fp64, on the GPU) -> dict of device tensors.
        grdS=None computes |grad q|^2 on the fly with the lat-lon stencil.
        stage_ms: optional ctypes float array of N_STAGES entries that receives the
        per-stage device time (makes the call synchronous)."""
        lib = ops.require_cuda()
        S = q.shape[0]
        assert q.is_cuda and q.is_contiguous() and tuple(q.shape[1:]) == (self.ny, self.nx)
        if out is None:
            out = self.alloc_outputs(S)
        nb = self.workspace_bytes(S)
        if ws is None:
            ws = ops.workspace(nb, tag="fused")
        a = KeffLwaArgs()
        a.q, a.q_dtype = q.data_ptr(), ops.fdtype(q)
        a.S, a.n_y, a.n_x, a.N = S, self.ny, self.nx, self.N
        a.increase, a.lt, a.ctr_dtype = int(self.increase), int(self.lt), self.ctr_dtype
        a.dA, a.dA_dtype = self.dA.data_ptr(), ops.fdtype(self.dA)
        if grdS is not None:
            a.grdS, a.grdS_dtype = grdS.data_ptr(), ops.fdtype(grdS)
        else:
            a.grdS, a.grdS_dtype = None, XC_F64
        a.lat_rad, a.dlambda = self.lat_rad.data_ptr(), self.dlambda
        a.cx, a.cy, a.bcx, a.bcy, a.fill_value = self.cx.data_ptr(), self.cy.data_ptr(), self.bcx, self.bcy, self.fill_value
        a.dA_row = self.dA_row.data_ptr() if self.dA_row is not None else None
        a.uniform_dA, a.any_degenerate = int(self.uniform_dA), int(self.any_degenerate)
        a.ww_row = self.ww_row.data_ptr() if self.ww_row is not None else None
        a.table, a.table_coord, a.n_table = self.table.data_ptr(), self.table_coord.data_ptr(), self.ny
        a.eq_coord, a.ww = self.eq_coord.data_ptr(), self.ww.data_ptr()
        a.keff_mask, a.part, a.sub_batch = self.keff_mask, self.part, self.sub_batch
        for k in CONTOUR_VARS:
            assert k not in out or out[k].is_contiguous()
            setattr(a, k, out[k].data_ptr() if k in out else None)
        a.Qref = out["Qref"].data_ptr() if "Qref" in out else None
        a.lwa = out["lwa"].data_ptr() if "lwa" in out else None
        a.lwa_f32 = int("lwa" in out and out["lwa"].dtype == torch.float32)
        a.numpy2_rules = int(self.numpy2_rules)
        a.stage_ms = ctypes.cast(stage_ms, ctypes.c_void_p) if stage_ms is not None else None
        check(lib.xc_keff_lwa_batch(ctypes.byref(a), ctypes.c_void_p(ws.data_ptr()), nb, ops.stream_ptr()))
        return out


def bind_host_thread_to_gpu(device_index=None):
    """Restrict the calling host thread to the CPUs NVML reports as local to the GPU
    (its NUMA node), so that pinned host buffers allocated afterwards are first-touched
    in the memory that GPU's PCIe root complex reaches without crossing the socket
    interconnect -- on an 8-GPU box that is what the end-to-end (host-buffer) path is
    bound by once every rank streams at PCIe rate.  Best effort: returns
    ``(previous_affinity, description)`` and never raises; ``previous_affinity`` is None
    when nothing was changed.  Undo with ``os.sched_setaffinity(0, previous_affinity)``."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        props = torch.cuda.get_device_properties(idx)
        h = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(uuid))
            except Exception:
                h = None
        if h is None and hasattr(props, "pci_bus_id"):
            try:
                h = pynvml.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (
                    getattr(props, "pci_domain_id", 0), props.pci_bus_id, getattr(props, "pci_device_id", 0)))
            except Exception:
                h = None
        if h is None:
            return None, "no NVML handle for cuda:%d" % idx
        allowed = os.sched_getaffinity(0)
        ncpu = max(allowed) + 1 if allowed else (os.cpu_count() or 1)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(ncpu, os.cpu_count() or 1) + 63) // 64)
        local = set(64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1)
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            node = None
        target = local & allowed
        if not target or target == allowed:
            return None, "numa node %s: %d local cpus, %d allowed, affinity unchanged" % (node, len(local), len(allowed))
        os.sched_setaffinity(0, target)
        return allowed, "numa node %s: bound to %d of %d allowed cpus" % (node, len(target), len(allowed))
    except Exception as e:                                      # no NVML, no permission, ...
        return None, "unavailable (%s)" % type(e).__name__


class HostStreamer(object):
    """End-to-end driver for HOST-resident tracers: pinned host slices ->
    (H2D, fused batch, D2H of every result) with ``nbuf`` buffers in flight so the
    copies of neighbouring batches overlap the kernels of the current one (H2D,
    kernels and D2H of a batch are ordered on that batch's own stream)."""

    def __init__(self, plan, batch, q_dtype=torch.float32, copy_lwa=True, nbuf=3, lwa_dtype=torch.float64):
        self.plan, self.batch, self.copy_lwa, self.nbuf = plan, int(batch), copy_lwa, int(nbuf)
        dev = plan.dA.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.nbuf)]
        self.qdev = [torch.empty((batch, plan.ny, plan.nx), dtype=q_dtype, device=dev) for _ in range(self.nbuf)]
        self.outs = [plan.alloc_outputs(batch, lwa_dtype=lwa_dtype) for _ in range(self.nbuf)]
        self.ws = [torch.empty(plan.workspace_bytes(batch), dtype=torch.uint8, device=dev) for _ in range(self.nbuf)]
        self.host = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in o.items()
                      if copy_lwa or k != "lwa"} for o in self.outs]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, q_host, consume=None):
        """q_host: pinned CPU tensor [S, ny, nx].  ``consume(s0, s1, host_dict)`` is
        called for every finished batch (host buffers are reused afterwards)."""
        S = q_host.shape[0]
        pending = [None] * self.nbuf
        nb = (S + self.batch - 1) // self.batch
        for b in range(nb + self.nbuf):
            i = b % self.nbuf
            if pending[i] is not None:                    # drain the batch that used buffer i
                s0, s1, ev = pending[i]
                ev.synchronize()
                if consume is not None:
                    consume(s0, s1, {k: v[:s1 - s0] for k, v in self.host[i].items()})
                pending[i] = None
            if b >= nb:
                continue
            s0, s1 = b * self.batch, min(S, (b + 1) * self.batch)
            n = s1 - s0
            with torch.cuda.stream(self.streams[i]):
                self.qdev[i][:n].copy_(q_host[s0:s1], non_blocking=True)
                self.h2d_bytes += q_host[s0:s1].numel() * q_host.element_size()
                out = {k: v[:n] for k, v in self.outs[i].items()}
                self.plan.run(self.qdev[i][:n], out=out, ws=self.ws[i])
                for k, hv in self.host[i].items():
                    hv[:n].copy_(out[k], non_blocking=True)
                    self.d2h_bytes += out[k].numel() * out[k].element_size()
                ev = torch.cuda.Event()
                ev.record(self.streams[i])
            pending[i] = (s0, s1, ev)


class ContourGather(object):
    """Asynchronous all-gather of the packed contour-space results ([9, S_local, N] per rank, see
    KeffLwaPlan.alloc_outputs) on a side stream: the collective of batch b overlaps the kernels of batch b+1.
    The only communication of the whole path; LWA fields stay sharded.  NCCL with CUDA tensors;
    with CPU tensors (gloo: the host-logic tests) the same buffers and layout, the collective issued in place.
    Every rank contributes the same S_local (the last batch of a run is padded by the caller).
    ``nbuf`` receive buffers rotate; ``wait()`` makes the current stream wait for every gather in flight."""

    def __init__(self, S_local, N, device, group=None, nbuf=2):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.recv = [torch.empty((self.world, len(CONTOUR_VARS), S_local, N), dtype=torch.float64, device=device)
                     for _ in range(nbuf)]
        self.on_gpu = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.on_gpu else None
        self.done = [None] * nbuf
        self.n = 0

    def launch(self, packed):
        """Enqueue the gather of ``packed`` (produced on the current stream); returns the receive buffer
        [world, 9, S_local, N] it will land in, and the index to pass to ``event()``."""
        i = self.n % len(self.recv)
        self.n += 1
        if tuple(packed.shape) != tuple(self.recv[i].shape[1:]):
            raise Exception("ContourGather: packed buffer %s, expected %s" % (tuple(packed.shape), tuple(self.recv[i].shape[1:])))
        if not self.on_gpu:
            self.dist.all_gather_into_tensor(self.recv[i].view(-1), packed.contiguous().view(-1), group=self.group)
            return self.recv[i], i
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.dist.all_gather_into_tensor(self.recv[i].view(-1), packed.view(-1), group=self.group)
            self.done[i] = torch.cuda.Event()
            self.done[i].record(self.stream)
        return self.recv[i], i

    def event(self, i):
        return self.done[i]

    def wait(self):
        if self.on_gpu:
            torch.cuda.current_stream().wait_stream(self.stream)

    @staticmethod
    def unpack(recv, S_total=None):
        """[world, 9, S_local, N] -> {name: [world * S_local (or S_total), N]} (views where possible)."""
        w, k, s, n = recv.shape
        out = {}
        for i, name in enumerate(CONTOUR_VARS):
            t = recv[:, i].reshape(w * s, n)
            out[name] = t if S_total is None else t[:S_total]
        return out


def gather_contour_space(local, S_total, group=None):
    """All-gather the contour-space results ([S_local, N] each) of every rank into
    [S_total, N] tensors.  Works on NCCL (CUDA tensors) and gloo (CPU tensors);
    ranks own contiguous ``slice_range`` blocks, padded to equal length for the
    collective.  This is the only communication of the whole path."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (S_total + world - 1) // world
    out = {}
    for k, v in local.items():
        pad = torch.zeros((per,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        pad[:v.shape[0]] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat(parts, dim=0)[:S_total]
    return out
