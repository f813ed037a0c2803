"""
Thin functional wrappers over the C ABI (include/xcb200.h) on torch CUDA
tensors.  torch is plumbing only: it owns device memory and the current stream;
every numerical step is a hand-written sm_100a kernel inside libxcb200.so.

Nothing here falls back to the CPU: without a CUDA device, or without the
built library, every function raises.
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib
from ._lib import (MAX_INTEGRANDS, PART, SCAN_PREFIX, SCAN_SUFFIX,
                   SCAN_TOTAL_MINUS, XC_F32, XC_F32_AS_F64, XC_F64, check)

_WS = {}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("xcontour_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    return _lib.load()


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes, tag="default"):
    """A cached scratch buffer per (device, tag), grown on demand."""
    key = (torch.cuda.current_device(), tag)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        _WS[key] = buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device())
    return buf


def to_dev(x, dtype=None):
    """numpy / torch / any DLPack exporter -> contiguous CUDA tensor."""
    dev = device()
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray) or np.isscalar(x) or isinstance(x, (list, tuple)):
        a = np.ascontiguousarray(x)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("="))
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    t = t.to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


_STAGE = {}
_CHUNK = 8 << 20
_NP_OF = {torch.float64: np.float64, torch.float32: np.float32, torch.int64: np.int64, torch.int32: np.int32,
          torch.int16: np.int16, torch.int8: np.int8, torch.uint8: np.uint8, torch.bool: np.bool_}


def _staging(dev_index):
    """two pinned 8 MB buffers + their events per (host thread, device); idle on return"""
    key = (threading.get_ident(), dev_index)
    st = _STAGE.get(key)
    if st is None:
        st = _STAGE[key] = ([torch.empty(_CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(2)],
                            [torch.cuda.Event() for _ in range(2)])
    for e in st[1]:
        e.synchronize()                                  # a DMA of an earlier transfer may still be reading them
    return st


def to_host(t):
    """CUDA tensor -> a fresh NumPy array.  Large results (an LWA field is 8.3 MB per slice) go through two pinned
    staging buffers: the DMA of chunk k+1 runs at the PCIe rate while chunk k is copied into the pageable result
    (a plain ``.cpu()`` of a 66 MB field ran at 2.2 GB/s on the B200 boxes, 30 of the 42 ms of an 8-slice
    Contour2D workflow).  The staging buffers are 2 x 8 MB per host thread and device; the result owns its memory."""
    if not t.is_cuda:
        return t.detach().numpy()
    nbytes = t.numel() * t.element_size()
    chunk_bytes = _CHUNK
    if nbytes < 4 * chunk_bytes or t.dtype not in _NP_OF:
        return t.detach().cpu().numpy()
    t = t.detach().contiguous()
    out = np.empty(tuple(t.shape), dtype=_NP_OF[t.dtype])
    src = t.reshape(-1).view(torch.uint8)
    dst = torch.from_numpy(out.reshape(-1).view(np.uint8))
    n = (nbytes + chunk_bytes - 1) // chunk_bytes
    with torch.cuda.device(t.device):
        bufs, evs = _staging(t.device.index)
        for k in range(n + 1):
            if k < n:                                   # DMA of chunk k into the buffer chunk k-2 has left
                lo = k * chunk_bytes
                hi = min(nbytes, lo + chunk_bytes)
                bufs[k & 1][:hi - lo].copy_(src[lo:hi], non_blocking=True)
                evs[k & 1].record()
            if k >= 1:                                  # chunk k-1 has landed: into the result
                lo = (k - 1) * chunk_bytes
                hi = min(nbytes, lo + chunk_bytes)
                evs[(k - 1) & 1].synchronize()
                dst[lo:hi].copy_(bufs[(k - 1) & 1][:hi - lo])
    return out


def fdtype(t):
    """XC dtype code of a float32/float64 tensor."""
    if t.dtype == torch.float32:
        return XC_F32
    if t.dtype == torch.float64:
        return XC_F64
    raise TypeError("expected a float32 or float64 tensor, got %s" % t.dtype)


def as_float(t):
    """Widen anything that is not fp32/fp64 to fp64 (ints, bools, halves)."""
    return t if t.dtype in (torch.float32, torch.float64) else t.to(torch.float64)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------------------
def minmax_levels(q, N, increase, out_dtype=XC_F32):
    """q[S, P] -> (levels[S, N] fp64 holding out_dtype-rounded values, minmax[S, 2])."""
    lib = require_cuda()
    S, P = q.shape
    levels = torch.empty((S, N), dtype=torch.float64, device=q.device)
    mm = torch.empty((S, 2), dtype=torch.float64, device=q.device)
    nb = lib.xc_minmax_levels_workspace_bytes(S, P)
    ws = workspace(nb)
    check(lib.xc_minmax_levels(_p(q), fdtype(q), S, P, int(N), int(bool(increase)), out_dtype,
                               _p(levels), _p(mm), _p(ws), nb, stream_ptr()))
    return levels, mm


def equal_area_levels(q, dA, N, refine, increase, lt, out_dtype, numpy2_rules):
    """q: [S, P] device tensor (fp32 / fp64), dA: [P] -> [S, N] fp64 levels rounded to out_dtype (xc_equal_area_levels)."""
    lib = require_cuda()
    S, P = q.shape
    out = torch.empty((S, int(N)), dtype=torch.float64, device=q.device)
    nb = lib.xc_equal_area_levels_workspace_bytes(S, P, int(N), int(refine))
    ws = workspace(nb)
    check(lib.xc_equal_area_levels(_p(q), fdtype(q), S, P, _p(dA), fdtype(dA), int(N), int(refine),
                                   int(bool(increase)), int(bool(lt)), int(out_dtype), int(bool(numpy2_rules)),
                                   _p(out), _p(ws), nb, stream_ptr()))
    return out


def hist_edges(levels, ctr_dtype, time_branch):
    lib = require_cuda()
    S, N = levels.shape
    edges = torch.empty((S, N + 1), dtype=torch.float64, device=levels.device)
    decr = torch.empty((S,), dtype=torch.int32, device=levels.device)
    check(lib.xc_hist_edges(_p(levels), S, N, ctr_dtype, int(bool(time_branch)),
                            _p(edges), _p(decr), stream_ptr()))
    return edges, decr


def bin_accumulate(q, edges, dA, acc_area=True, integrands=(), closed_right=False,
                   scan_mode=SCAN_PREFIX, decreasing=None, q_mask=None,
                   want_pdf=False, want_idx=False):
    """q[S, P]; edges[S, N+1] or [N+1]; dA[P]; integrands: list of [S, P].
    Returns (cdf[S, K, N], pdf or None, bin_idx or None)."""
    lib = require_cuda()
    S, P = q.shape
    N = edges.shape[-1] - 1
    stride = 0 if edges.dim() == 1 or edges.shape[0] == 1 and S > 1 else N + 1
    n_int = len(integrands)
    if n_int > MAX_INTEGRANDS:
        raise Exception("at most %d integrands per call" % MAX_INTEGRANDS)
    K = (1 if acc_area else 0) + n_int
    cdf = torch.empty((S, K, N), dtype=torch.float64, device=q.device)
    pdf = torch.empty((S, K, N), dtype=torch.float64, device=q.device) if want_pdf else None
    idx = torch.empty((S, P), dtype=torch.int32, device=q.device) if want_idx else None
    ptrs = (ctypes.c_void_p * max(n_int, 1))(*[g.data_ptr() for g in integrands])
    dts = (ctypes.c_int * max(n_int, 1))(*[fdtype(g) for g in integrands])
    nb = lib.xc_bin_accumulate_workspace_bytes(S, P, N, K)
    ws = workspace(nb)
    check(lib.xc_bin_accumulate(_p(q), fdtype(q), S, P, _p(edges), stride, N, int(bool(closed_right)),
                                _p(dA), fdtype(dA), int(bool(acc_area)), ptrs, dts, n_int,
                                _p(q_mask), scan_mode, _p(decreasing),
                                _p(pdf), _p(cdf), _p(idx), _p(ws), nb, stream_ptr()))
    return cdf, pdf, idx


def interp(x, xp, fp, reverse=-1):
    """np.interp per slice.  x: [S, M] or [M]; xp, fp: [S, n] or [n] (fp64)."""
    lib = require_cuda()
    M, n = x.shape[-1], xp.shape[-1]
    S = max(x.shape[0] if x.dim() == 2 else 1, xp.shape[0] if xp.dim() == 2 else 1,
            fp.shape[0] if fp.dim() == 2 else 1)

    def st(t, last):
        if t.dim() == 2 and t.shape[0] not in (1, S):
            raise Exception("interp: operands with %d and %d slices cannot be paired" % (t.shape[0], S))
        return last if (t.dim() == 2 and t.shape[0] == S) else 0
    out = torch.empty((S, M), dtype=torch.float64, device=x.device)
    check(lib.xc_interp(_p(x), st(x, M), M, _p(xp), st(xp, n), _p(fp), st(fp, n), n, int(reverse),
                        S, _p(out), stream_ptr()))
    return out


def gradient_wrt_area(var, var_kind, area, area_kind):
    lib = require_cuda()
    S, N = var.shape
    out = torch.empty((S, N), dtype=torch.float64, device=var.device)
    check(lib.xc_gradient_wrt_area(_p(var), var_kind, _p(area), area_kind, S, N, _p(out), stream_ptr()))
    return out


def gradient_coefficients(coord):
    """NumPy's own np.gradient(edge_order=1) difference coefficients for a 1-D coordinate, in the coordinate's
    dtype (numpy/lib/_function_base_impl.py): -> (coef fp64 device tensor, uniform flag, holds-fp32 flag)."""
    x = np.asanyarray(coord)
    if np.issubdtype(x.dtype, np.integer):
        x = x.astype(np.float64)
    d = np.diff(x)
    if (d == d[0]).all():
        coef = np.array([2. * d[0], d[0]])
        uniform = 1
    else:
        dx1, dx2 = d[0:-1], d[1:]
        coef = np.concatenate([-(dx2) / (dx1 * (dx1 + dx2)), (dx2 - dx1) / (dx1 * dx2), dx1 / (dx2 * (dx1 + dx2)),
                               d[:1], d[-1:]])
        uniform = 0
    return to_dev(np.ascontiguousarray(coef, dtype=np.float64)), uniform, int(coef.dtype == np.float32)


def gradient_wrt_area_coord(var, var_coord, area, area_coord):
    """np.gradient(var, var_coord) / np.gradient(area, area_coord) along the last axis of [S, N] fp32/fp64 tensors."""
    lib = require_cuda()
    S, N = var.shape
    out = torch.empty((S, N), dtype=torch.float64, device=var.device)
    vc, vu, v32 = gradient_coefficients(var_coord)
    ac, au, a32 = gradient_coefficients(area_coord)
    check(lib.xc_gradient_wrt_area_coord(_p(var), fdtype(var), _p(vc), vu, v32, _p(area), fdtype(area), _p(ac), au, a32,
                                         S, N, _p(out), stream_ptr()))
    return out


def leq2(dgrdSdA, dqdA):
    lib = require_cuda()
    out = torch.empty_like(dgrdSdA)
    check(lib.xc_leq2(_p(dgrdSdA), _p(dqdA), dgrdSdA.numel(), _p(out), stream_ptr()))
    return out


def lmin(lat):
    lib = require_cuda()
    out = torch.empty_like(lat)
    check(lib.xc_lmin(_p(lat), lat.numel(), _p(out), stream_ptr()))
    return out


def nkeff(Leq2, Lmin, mask):
    lib = require_cuda()
    out = torch.empty_like(Leq2)
    check(lib.xc_nkeff(_p(Leq2), _p(Lmin), float(mask), Leq2.numel(), _p(out), stream_ptr()))
    return out


def eqlat(area):
    lib = require_cuda()
    out = torch.empty_like(area)
    check(lib.xc_eqlat(_p(area), area.numel(), _p(out), stream_ptr()))
    return out


def lwa_weights(dA):
    """dA[P] (fp32/fp64) -> ww[P] fp64 = (dA/max dA) * dA."""
    lib = require_cuda()
    P = dA.numel()
    ww = torch.empty((P,), dtype=torch.float64, device=dA.device)
    nb = lib.xc_lwa_weights_workspace_bytes(P)
    ws = workspace(nb)
    check(lib.xc_lwa_weights(_p(dA), fdtype(dA), P, _p(ww), _p(ws), nb, stream_ptr()))
    return ww


def row_constant(w, ny, nx):
    """w[ny*nx] on the device -> its first column [ny] (fp64) when every row holds one value (NaN rows
    included), else None.  One reduction + one host sync; callers do it once per weight array."""
    m = w.reshape(ny, nx)
    first = m[:, :1]
    same = (m == first) | (m.isnan() & first.isnan())
    return first.reshape(ny).to(torch.float64).contiguous() if bool(same.all()) else None


def lwa(q, Q, ww, increase, part="all", variant=1, out=None, ww_row=None):
    """q[S, n_eq, n_x], Q[S, n_eq] fp64, ww[n_eq*n_x] fp64 -> LWA[S, n_eq, n_x] fp64.
    ww_row: row_constant(ww, n_eq, n_x) when known (selects the column-tile kernel)."""
    lib = require_cuda()
    if part not in PART:
        raise Exception("invalid part, should be in ['all', 'upper', 'lower']")
    S, ny, nx = q.shape
    if out is None:
        out = torch.empty((S, ny, nx), dtype=torch.float64, device=q.device)
    nb = lib.xc_lwa_workspace_bytes(S)
    ws = workspace(nb)
    check(lib.xc_lwa_ex(_p(q), fdtype(q), S, ny, nx, _p(Q), _p(ww), _p(ww_row), int(bool(increase)), PART[part],
                        int(variant), _p(out), _p(ws), nb, stream_ptr()))
    return out


def lwa_mask(q, Q, j, increase, variant=1):
    lib = require_cuda()
    S, ny, nx = q.shape
    out = torch.empty((S, ny, nx), dtype=torch.int8, device=q.device)
    check(lib.xc_lwa_mask(_p(q), fdtype(q), S, ny, nx, _p(Q), int(j), int(bool(increase)),
                          int(variant), _p(out), stream_ptr()))
    return out


def grad2_latlon(q, lat_rad, dlambda, out_dtype=torch.float64):
    lib = require_cuda()
    S, ny, nx = q.shape
    out = torch.empty((S, ny, nx), dtype=out_dtype, device=q.device)
    check(lib.xc_grad2_latlon(_p(q), fdtype(q), S, ny, nx, _p(lat_rad), float(dlambda),
                              _p(out), fdtype(out), stream_ptr()))
    return out


def latlon_cell_area(lat_deg, n_x, dlambda_deg, out_dtype=torch.float64):
    """lat_deg[n_y] (fp64, on the GPU) -> dA[n_y, n_x] built on the device."""
    lib = require_cuda()
    ny = lat_deg.numel()
    out = torch.empty((ny, n_x), dtype=out_dtype, device=lat_deg.device)
    check(lib.xc_latlon_cell_area(_p(lat_deg), ny, int(n_x), float(dlambda_deg), _p(out), fdtype(out), stream_ptr()))
    return out


def launch_count():
    return _lib.load().xc_launch_count()


def reset_launch_count():
    _lib.load().xc_reset_launch_count()
