"""
Host-side mirror of xcontour's public contour-analysis API (Contour2D, Table --
/root/reference/xcontour/core.py:16-1195) over the B200 kernels.

Same class names, method names, argument meaning, return conventions (labelled
arrays: dims, coords, names) and error behaviour (bare ``Exception`` with the
reference's messages).  What differs is what runs underneath: each method hands
device buffers (torch tensors, DLPack at the edge) to one or two entry points of
libxcb200.so instead of orchestrating xarray / xhistogram / np.vectorize.  There
is no CPU path: without the CUDA library every compute method raises.

Deliberate, documented deviations from the literal reference (SURVEY.md §8a
hazards H1-H6; details in DESIGN.md):
  * H1  the strict-integration and LWA methods reduce over the 2-D plane / the
        equivalent dimension given by ``dims`` / ``dimEq`` *values* (the
        reference passes the dict keys to ``.sum`` there, which only works when
        keys == dimension names);
  * the histogram path accepts contours that vary along ANY leading dimension
    (the reference loops only over a dim literally named 'time',
    core.py:1242-1287); such contours take that per-'time' branch's arithmetic;
  * sums are accumulated in fp64 on the GPU (the reference's strict path sums in
    the promoted dtype of integrand*dA, e.g. fp32).
"""
import os

import numpy as np
import torch

from . import ops
from . import utils
from ._lib import (SCAN_PREFIX, SCAN_SUFFIX, SCAN_TOTAL_MINUS, XC_F32,
                   XC_F32_AS_F64, XC_F64)
from . import xr_compat as xc
from .xr_compat import DataArray, Dataset, merge

# The NumPy scalar-promotion regime of the reference that is reproduced (numpy1 | numpy2) lives in
# utils.NUMPY_SCALAR_RULES: by default the regime of the installed NumPy (hazard H4).


def _np_dtype_code(dt):
    return XC_F32 if np.dtype(dt) == np.float32 else XC_F64


class Contour2D(object):
    """
    This class is designed for performing the 2D contour analysis.
    (drop-in for xcontour.Contour2D, core.py:16-70)
    """

    def __init__(self, trcr, dA, dims, dimEq, arakawa='A',
                 increase=True, lt=False, check_mono=False, dtype=np.float32):
        if len(dimEq) != 1:
            raise Exception('dimEq should be one dimension e.g., {"Y","lat"}')

        if len(dims) != 2:
            raise Exception('dims should be a 2D plane')

        self.dA      = dA
        self.arakawa = arakawa
        self.tracer  = trcr
        self.dims    = dims
        self.dimNs   = list(dims.keys())       # dim names,  ['X', 'Y', 'Z']
        self.dimVs   = list(dims.values())     # dim values, ['lon', 'lat', 'Z']
        self.dimEqN  = list(dimEq.keys())[0]   # equiv. dim name
        self.dimEqV  = list(dimEq.values())[0] # equiv. dim value
        self.lt      = lt
        self.dtype   = dtype
        self.check_mono = check_mono
        self.increase   = increase
        self._cache = {}

    # ------------------------------------------------------------------ layout
    def _layout(self, da):
        """Split the dims of a tracer-like array into (leading dims, plane dims in
        array order) and check that the plane is the trailing two dims."""
        plane = [d for d in da.dims if d in self.dimVs]
        if len(plane) != 2:
            raise Exception('tracer does not contain the 2D plane %s' % (self.dimVs,))
        lead = [d for d in da.dims if d not in self.dimVs]
        return lead, plane

    def _dev_field(self, da, key=None):
        """(tensor [S, n0, n1] on the GPU, lead dims, plane dims).  The upload of
        ``self.tracer`` / ``self.dA`` is cached for as long as the attribute refers to
        the same array object (the attributes are public and may be reassigned, as in
        the reference).  Values that already live in torch / behind DLPack (the
        stand-in DataArray keeps them as they are) go to the kernels without a host
        round trip."""
        if key is not None:
            hit = self._cache.get(key)
            if hit is not None and hit[0] is da:
                return hit[1]
        lead, plane = self._layout(da)
        perm = [da.dims.index(d) for d in lead + plane]
        raw = getattr(da, '_raw', None)
        if raw is not None:
            t = ops.to_dev(raw)
            if perm != list(range(len(perm))):
                t = t.permute(perm)
        else:
            vals = da.values
            if perm != list(range(len(perm))):
                vals = np.transpose(vals, perm)
            t = ops.to_dev(vals)
        t = ops.as_float(t)
        t = t.reshape((-1,) + tuple(t.shape[-2:])).contiguous()
        res = (t, lead, plane)
        if key is not None:
            self._cache[key] = (da, res)
        return res

    def _tracer_dev(self, tracer=None):
        if tracer is None or tracer is self.tracer:
            return self._dev_field(self.tracer, 'tracer')
        return self._dev_field(tracer)

    def _dA_plane(self, plane, like=None):
        """dA broadcast to the plane (host, its own dtype) and uploaded once per dA object.
        ``like``: the tracer whose plane shape applies (default: self.tracer).  Cell areas
        that vary along a non-plane dimension (time- or level-dependent, partial cells)
        are not supported by the kernels: they raise instead of failing inside NumPy."""
        trc = self.tracer if like is None else like
        n0 = trc.shape[trc.dims.index(plane[0])]
        n1 = trc.shape[trc.dims.index(plane[1])]
        key = ('dA',) + tuple(plane) + (n0, n1)
        hit = self._cache.get(key)
        if hit is not None and hit[0] is self.dA:
            return hit[1]
        dA = self.dA
        if xc.is_labeled(dA):
            vals = np.asarray(dA.values)
            ddims = list(dA.dims)
            keep = [i for i, d in enumerate(ddims) if d in plane]
            if len(keep) != len(ddims):          # drop singleton non-plane dims
                extra = [ddims[i] for i in range(len(ddims)) if i not in keep and vals.shape[i] != 1]
                if extra:
                    raise Exception('dA varies along %s: cell areas must be constant along every '
                                    'dimension outside the 2D plane %s' % (extra, list(plane)))
                vals = vals.reshape([vals.shape[i] for i in keep])
                ddims = [ddims[i] for i in keep]
            order = [d for d in plane if d in ddims]
            vals = np.transpose(vals, [ddims.index(d) for d in order])
            shp = [vals.shape[order.index(d)] if d in order else 1 for d in plane]
            vals = vals.reshape(shp)
        else:
            vals = np.asarray(dA)
        vals = np.ascontiguousarray(np.broadcast_to(vals, (n0, n1)))
        if vals.dtype not in (np.float32, np.float64):
            vals = vals.astype(np.float64)
        res = (ops.to_dev(vals), vals)
        self._cache[key] = (dA, res)
        return res

    def _lead_coords(self, da, lead):
        return xc.coords_for(da, lead)

    @staticmethod
    def _contour_coord(N, dtype=np.float32):
        return np.linspace(0.0, N - 1.0, N, dtype=dtype)

    # ---------------------------------------------------------------- contours
    def cal_contours(self, levels=10):
        """
        Establishing contour levels (space) of the tracer from its minimum
        to maximum values (core.py:205-266).  int -> equally spaced levels,
        array -> the prescribed levels broadcast to every slice.
        """
        q, lead, plane = self._tracer_dev()
        S = q.shape[0]
        lead_shape = tuple(self.tracer.shape[self.tracer.dims.index(d)] for d in lead)
        coords = self._lead_coords(self.tracer, lead)
        if type(levels) is int:
            N = levels
            lv, _ = ops.minmax_levels(q.reshape(S, -1), N, self.increase,
                                      _np_dtype_code(self.dtype))
            ctr = lv.cpu().numpy().astype(self.dtype).reshape(lead_shape + (N,))
            coords['contour'] = np.linspace(0.0, levels - 1.0, levels, dtype=self.dtype)
        else:
            levs = np.asarray(levels)
            N = levs.shape[0]
            ctr = np.broadcast_to(levs.astype(self.dtype), lead_shape + (N,)).copy()
            coords['contour'] = levs
        return xc.make(ctr, lead + ['contour'], coords, self.tracer.name)

    def cal_contours_at(self, predef, table):
        """core.py:269-313 (conditional-integration variant)."""
        return self._cal_contours_at(predef, table, hist=False)

    def cal_contours_at_hist(self, predef, table):
        """core.py:316-360 (histogram variant)."""
        return self._cal_contours_at(predef, table, hist=True)

    def _cal_contours_at(self, predef, table, hist):
        if len(predef.shape) != 1:
            raise Exception('predef should be a 1D array')
        if not xc.is_labeled(predef):
            predef = xc.make(np.asarray(predef), ['new'], {'new': np.asarray(predef)})
        N = predef.size
        ctr = self.cal_contours(N)
        area = self.cal_integral_within_contours_hist(ctr) if hist else \
            self.cal_integral_within_contours(ctr)
        dimEq = table.lookup_coordinates(area).rename('Z')
        qIntp = self.interp_to_coords(predef.squeeze(), dimEq, ctr.squeeze()) \
                    .rename({predef.dims[0]: 'contour'}).rename(ctr.name)
        qIntp = qIntp.assign_coords({'contour': np.linspace(0, N - 1, N, dtype=self.dtype)})
        return qIntp

    def cal_contours_equal_area(self, levels=10, refine=8):
        """
        Equal-area contour levels by a weighted-quantile histogram (an extension:
        the reference only offers equally *spaced* levels, core.py:205-266, and
        levels at prescribed equivalent coordinates, core.py:316-360).

        ``(levels-1)*refine+1`` equally spaced levels are binned with cell-area
        weights (cal_integral_within_contours_hist), the resulting A(q) relation
        is inverted by linear interpolation at ``levels`` equally spaced areas.
        Adjacent returned levels therefore enclose (to within one fine bin) the
        same area increment.  One library call (xc_equal_area_levels) chains the
        kernels of the methods it composes on the device; only the levels come back.
        """
        N = int(levels)
        q, lead, plane = self._tracer_dev(None)
        S = q.shape[0]
        dA_dev, _ = self._dA_plane(plane, self.tracer)
        # one call on the device (xc_equal_area_levels): levels -> edges -> area CDF -> inverse interpolation
        out = ops.equal_area_levels(q.reshape(S, -1), dA_dev.reshape(-1), N, int(refine), self.increase, self.lt,
                                    _np_dtype_code(self.dtype), utils.scalar_rules() == "numpy2")
        lshape = tuple(self.tracer.shape[self.tracer.dims.index(d)] for d in lead)
        res = out.cpu().numpy().astype(self.dtype).reshape(lshape + (N,))
        coords = xc.coords_for(self.tracer, lead)
        coords['contour'] = np.linspace(0.0, N - 1.0, N, dtype=self.dtype)
        return xc.make(res, lead + ['contour'], coords, self.tracer.name)

    # --------------------------------------------------------------- integrals
    def _contour_array(self, contour, lead, lead_shape):
        """-> (levels ndarray [S or 1, N], per_slice flag, contour coord values)."""
        if not xc.is_labeled(contour):
            contour = np.asarray(contour)
            contour = xc.make(contour, ['contour'], {'contour': contour})
        if 'contour' not in contour.dims:
            raise Exception('contour should have a dimension named contour')
        cdims = [d for d in contour.dims if d != 'contour']
        vals = np.asarray(contour.values)
        vals = np.transpose(vals, [contour.dims.index(d) for d in cdims + ['contour']])
        N = vals.shape[-1]
        ccoord = xc.coord(contour, 'contour')
        keep = [i for i in range(len(cdims)) if vals.shape[i] != 1]   # bins.squeeze()
        kd = [cdims[i] for i in keep]
        vals = vals.reshape([vals.shape[i] for i in keep] + [N])
        if not kd:                                  # static bins (core.py:1296-1313)
            return vals.reshape(1, N), False, ccoord
        for d in kd:
            if d not in lead:
                raise Exception('contour has a dimension (%s) that the tracer does not have' % d)
        order = [d for d in lead if d in kd]
        vals = np.transpose(vals, [kd.index(d) for d in order] + [len(kd)])
        shp = [vals.shape[order.index(d)] if d in order else 1 for d in lead]
        vals = np.broadcast_to(vals.reshape(shp + [N]), tuple(lead_shape) + (N,))
        return np.ascontiguousarray(vals).reshape(-1, N), True, ccoord

    def cal_integral_within_contours(self, contour, tracer=None, integrand=None):
        """
        Conditional integral of a (masked) variable within each tracer contour by
        strict comparison, ``tracer < contour`` (lt) or ``tracer > contour``
        (core.py:363-409).  One pass over the tracer: cells are binned against the
        sorted levels and a prefix (lt) / suffix (gt) scan over bins gives the
        same sums the reference obtains from a 4-D broadcast.
        """
        trc = self.tracer if tracer is None else tracer
        q, lead, plane = self._tracer_dev(tracer)
        S = q.shape[0]
        lead_shape = tuple(trc.shape[trc.dims.index(d)] for d in lead)
        levels, per_slice, ccoord = self._contour_array(contour, lead, lead_shape)
        N = levels.shape[-1]
        lev64 = levels.astype(np.float64)
        decreasing = lev64[:, 0] > lev64[:, -1]
        asc = np.where(decreasing[:, None], lev64[:, ::-1], lev64)
        if self.lt:
            edges = np.concatenate([np.full((asc.shape[0], 1), -np.inf), asc], axis=1)
            closed_right, mode = False, SCAN_PREFIX
        else:
            edges = np.concatenate([asc, np.full((asc.shape[0], 1), np.inf)], axis=1)
            closed_right, mode = True, SCAN_SUFFIX
        if edges.shape[0] == 1 and S > 1:
            edges = np.broadcast_to(edges, (S, N + 1)).copy()
            decreasing = np.broadcast_to(decreasing, (S,)).copy()
        dA_dev, dA_np = self._dA_plane(plane, trc)
        integ = []
        res_dtype = np.result_type(trc.dtype, dA_np.dtype)
        if integrand is not None:
            g, _, _ = self._dev_field(integrand)
            if g.shape[0] != S:
                g = g.expand(S, -1, -1).contiguous()
            integ = [g.reshape(S, -1)]
            res_dtype = np.result_type(integrand.dtype, dA_np.dtype)
        cdf, _, _ = ops.bin_accumulate(
            q.reshape(S, -1), ops.to_dev(edges), dA_dev.reshape(-1),
            acc_area=integrand is None, integrands=integ, closed_right=closed_right,
            scan_mode=mode, decreasing=ops.to_dev(decreasing.astype(np.int32)))
        out = cdf[:, 0, :].cpu().numpy().astype(res_dtype).reshape(lead_shape + (N,))
        coords = self._lead_coords(trc, lead)
        if ccoord is not None:
            coords['contour'] = ccoord
        intVar = xc.make(out, lead + ['contour'], coords, None)
        if self.check_mono:
            _check_monotonicity(intVar, 'contour')
        return intVar

    def cal_integral_within_contours_hist(self, contour, tracer=None, integrand=None):
        """
        Integral of a masked variable within pre-calculated tracer contours by the
        histogram method (core.py:412-460 + _histogram, core.py:1202-1325): bin
        every cell against the levels, accumulate integrand*dA per bin in fp64,
        scan over bins.
        """
        trc = self.tracer if tracer is None else tracer
        q, lead, plane = self._tracer_dev(tracer)
        S = q.shape[0]
        lead_shape = tuple(trc.shape[trc.dims.index(d)] for d in lead)
        levels, per_slice, _ = self._contour_array(contour, lead, lead_shape)
        N = levels.shape[-1]
        if not np.diff(levels, axis=-1).all():
            raise Exception('non monotonic bins')                 # core.py:1233-1240
        ctr_code = _np_dtype_code(levels.dtype) if levels.dtype in (np.float32, np.float64) \
            else XC_F64
        edges, decr = ops.hist_edges(ops.to_dev(levels.astype(np.float64)), ctr_code,
                                     time_branch=per_slice and utils.scalar_rules() != "numpy2")
        dA_dev, dA_np = self._dA_plane(plane, trc)
        integ = []
        if integrand is not None:
            g, _, _ = self._dev_field(integrand)
            if g.shape[0] != S:
                g = g.expand(S, -1, -1).contiguous()
            integ = [g.reshape(S, -1)]
        if not per_slice and S > 1:
            decr = decr.expand(S).contiguous()
        cdf, _, _ = ops.bin_accumulate(
            q.reshape(S, -1), edges if per_slice else edges[0], dA_dev.reshape(-1),
            acc_area=integrand is None, integrands=integ, closed_right=False,
            scan_mode=SCAN_PREFIX if self.lt else SCAN_TOTAL_MINUS, decreasing=decr)
        out = cdf[:, 0, :].cpu().numpy().reshape(lead_shape + (N,))
        coords = self._lead_coords(trc, lead)
        coords['contour'] = np.arange(N).astype(np.float32)       # core.py:1255-1257
        CDF = xc.make(out, lead + ['contour'], coords, None)
        if self.check_mono:
            _check_monotonicity(CDF, 'contour')
        return CDF

    # ------------------------------------------------------------------ tables
    def _eq_field(self, mask):
        """Coordinate vector of the equivalent dim and its axis in the plane."""
        plane = [d for d in mask.dims if d in self.dimVs]
        if len(plane) != 2:
            raise Exception('mask does not contain the 2D plane %s' % (self.dimVs,))
        ctr = xc.coord(mask, self.dimEqV)
        if ctr is None:
            raise Exception('mask has no coordinate %s' % self.dimEqV)
        lead = [d for d in mask.dims if d not in plane]
        mvals = np.asarray(mask.values)
        mvals = np.transpose(mvals, [mask.dims.index(d) for d in lead + plane])
        mvals = mvals.reshape((-1,) + mvals.shape[-2:])[0]     # mask is static in time
        return ctr, plane, plane.index(self.dimEqV), mvals

    def cal_area_eqCoord_table_hist(self, mask):
        """
        A(Yeq) relation table by the histogram method (core.py:150-203): the
        equivalent coordinate itself is binned against the coordinate vector with
        weights dA.
        """
        ctr, plane, eq_axis, mvals = self._eq_field(mask)
        n0, n1 = mvals.shape
        shp = [1, 1]
        shp[eq_axis] = ctr.shape[0]
        fdt = ctr.dtype if ctr.dtype in (np.float32, np.float64) else np.float64
        ctrVar = np.ascontiguousarray(np.broadcast_to(ctr.astype(fdt).reshape(shp), (n0, n1)))
        yIncre = not (ctr[-1] < ctr[0])
        ylt = self.lt if self.increase == yIncre else (not self.lt)
        edges, _ = ops.hist_edges(ops.to_dev(ctr.astype(np.float64).reshape(1, -1)),
                                  _np_dtype_code(fdt), time_branch=False)
        dA_dev, _ = self._dA_plane(plane)
        qmask = ops.to_dev((mvals == 1).astype(np.uint8).reshape(-1))
        cdf, _, _ = ops.bin_accumulate(
            ops.to_dev(ctrVar).reshape(1, -1), edges[0], dA_dev.reshape(-1), acc_area=True,
            scan_mode=SCAN_PREFIX if ylt else SCAN_TOTAL_MINUS, q_mask=qmask)
        tbl = cdf[0, 0].cpu().numpy()
        cvals = ctr if yIncre else ctr[::-1]
        tbl = xc.make(tbl, [self.dimEqV], {self.dimEqV: np.array(cvals)}, 'AeqCTbl')
        if self.check_mono:
            _check_monotonicity(tbl, self.dimEqV)
        return Table(tbl, self.dimEqV)

    def cal_area_eqCoord_table(self, mask):
        """
        A(Yeq) relation table by strict conditional integration of the mask
        (core.py:73-147), endpoint replaced by the total masked area.
        """
        ctr, plane, eq_axis, mvals = self._eq_field(mask)
        n0, n1 = mvals.shape
        shp = [1, 1]
        shp[eq_axis] = ctr.shape[0]
        c64 = ctr.astype(np.float64)
        ctrVar = np.ascontiguousarray(np.broadcast_to(c64.reshape(shp), (n0, n1)))
        eqDimIncre = bool(ctr[-1] > ctr[0])
        use_lt = (eqDimIncre == self.increase) if self.lt else (eqDimIncre != self.increase)
        asc = c64 if eqDimIncre else c64[::-1]
        if use_lt:
            edges = np.concatenate([[-np.inf], asc]); closed_right, mode = False, SCAN_PREFIX
        else:
            edges = np.concatenate([asc, [np.inf]]); closed_right, mode = True, SCAN_SUFFIX
        dA_dev, dA_np = self._dA_plane(plane)
        mfl = mvals.astype(np.float32 if mvals.dtype != np.float64 else np.float64)
        cdf, _, _ = ops.bin_accumulate(
            ops.to_dev(ctrVar).reshape(1, -1), ops.to_dev(edges), dA_dev.reshape(-1),
            acc_area=False, integrands=[ops.to_dev(mfl).reshape(1, -1)],
            closed_right=closed_right, scan_mode=mode,
            decreasing=ops.to_dev(np.array([0 if eqDimIncre else 1], dtype=np.int32)))
        tbl = np.abs(cdf[0, 0].cpu().numpy())                  # integral of mask*dA
        # total masked area: every cell satisfies exactly one of <, ==, > -- take it
        # from a dedicated pass with a single all-embracing bin
        tot, _, _ = ops.bin_accumulate(
            ops.to_dev(ctrVar).reshape(1, -1), ops.to_dev(np.array([-np.inf, np.inf])),
            dA_dev.reshape(-1), acc_area=False,
            integrands=[ops.to_dev(mfl).reshape(1, -1)], closed_right=False, scan_mode=SCAN_PREFIX)
        maxArea = abs(float(tot[0, 0, 0].item()))
        if tbl[-1] > tbl[0]:
            tbl[-1] = maxArea
        else:
            tbl[0] = maxArea
        tbl = xc.make(tbl, [self.dimEqV], {self.dimEqV: np.array(ctr)}, 'AeqCTbl')
        if self.check_mono:
            _check_monotonicity(tbl, self.dimEqV)
        return Table(tbl, self.dimEqV)

    # ---------------------------------------------------------------- gradients
    @staticmethod
    def _flat2(da, interp_dim='contour'):
        """values as [S, N] with the named dim last; returns (array, lead dims)."""
        lead = [d for d in da.dims if d != interp_dim]
        vals = np.asarray(da.values)
        vals = np.transpose(vals, [da.dims.index(d) for d in lead + [interp_dim]])
        return vals.reshape(-1, vals.shape[-1]), lead, vals.shape[:-1]

    @staticmethod
    def _align2(a, b, interp_dim='contour'):
        """
        Two labelled operands as [S, n_a] and [S, n_b] arrays with ``interp_dim`` last and the other dims paired BY
        NAME, as xarray broadcasts them (the reference runs these steps through DataArray arithmetic / apply_ufunc):
        the lead dims are the union in order of first appearance, an operand that lacks one is repeated along it,
        and a dim present in both with different lengths is an error.  Returns (a2, b2, lead dims, lead shape).
        """
        la = [d for d in a.dims if d != interp_dim]
        lb = [d for d in b.dims if d != interp_dim]
        lead = la + [d for d in lb if d not in la]
        size = {}
        for arr in (a, b):
            for d, n in zip(arr.dims, arr.shape):
                if d == interp_dim:
                    continue
                if size.setdefault(d, n) != n:
                    raise Exception('cannot align dimension %r: sizes %d and %d' % (d, size[d], n))
        lshape = tuple(size[d] for d in lead)

        def flat(arr, own):
            vals = np.asarray(arr.values)
            order = [d for d in lead if d in own] + [interp_dim]
            vals = np.transpose(vals, [arr.dims.index(d) for d in order])
            n = vals.shape[-1]
            vals = vals.reshape(tuple(size[d] if d in own else 1 for d in lead) + (n,))
            if vals.shape[:-1] != lshape:
                vals = np.broadcast_to(vals, lshape + (n,))
            return np.ascontiguousarray(vals).reshape(-1, n)
        return flat(a, la), flat(b, lb), lead, lshape

    def cal_gradient_wrt_area(self, var, area):
        """d(var)/dA by centred differences along the contour axis (core.py:463-488)."""
        v, a, lead, lshape = self._align2(var, area)
        if a.shape != v.shape:
            raise Exception('var and area differ along contour: %d and %d' % (v.shape[-1], a.shape[-1]))
        vt, at = ops.as_float(ops.to_dev(v)), ops.as_float(ops.to_dev(a))
        # differentiate('contour') runs against each array's own 'contour' coordinate: 0..N-1 from cal_contours(int),
        # the level values themselves after cal_contours(array) (core.py:253-264), possibly non-uniform
        N = v.shape[-1]
        vcoord, acoord = xc.coord(var, 'contour'), xc.coord(area, 'contour')
        unit = np.arange(N)
        if (vcoord is None or np.array_equal(vcoord, unit)) and (acoord is None or np.array_equal(acoord, unit)):
            out = ops.gradient_wrt_area(vt, ops.fdtype(vt), at, ops.fdtype(at))
        else:
            out = ops.gradient_wrt_area_coord(vt, unit.astype(np.float32) if vcoord is None else vcoord,
                                              at, unit.astype(np.float32) if acoord is None else acoord)
        rdt = np.result_type(np.asarray(var.values).dtype, np.asarray(area.values).dtype)
        if rdt not in (np.float32, np.float64):
            rdt = np.float64
        res = out.cpu().numpy().astype(rdt).reshape(tuple(lshape) + (v.shape[-1],))
        coords = xc.coords_for(area, [d for d in lead if d in area.dims])
        coords.update(xc.coords_for(var, [d for d in lead + ['contour'] if d in var.dims]))
        name = 'dvardA' if var.name is None else 'd' + var.name + 'dA'
        return xc.make(res, lead + ['contour'], coords, name)

    def _ew(self, fn, a, b, name, *extra):
        """element-wise fp64 kernel on two equally shaped labelled arrays."""
        av = np.asarray(a.values, dtype=np.float64)
        bv = np.broadcast_to(np.asarray(b.values, dtype=np.float64), av.shape)
        out = fn(ops.to_dev(av), ops.to_dev(np.ascontiguousarray(bv)), *extra)
        return xc.make(out.cpu().numpy().reshape(av.shape), a.dims, xc.coords_for(a, a.dims), name)

    def cal_sqared_equivalent_length(self, dgrdSdA, dqdA):
        """Leq2 = dgrdSdA / dqdA**2 (core.py:619-637)."""
        return self._ew(ops.leq2, dgrdSdA, dqdA, 'Leq2')

    def cal_normalized_Keff(self, Leq2, Lmin, mask=1e5):
        """nkeff = Leq2/Lmin/Lmin, NaN where it is not below ``mask`` (core.py:945-966)."""
        return self._ew(ops.nkeff, Leq2, Lmin, 'nkeff', mask)

    # ------------------------------------------------------- along-contour means
    def cal_contour_weigh_mean(self, contour, integrand, area=None):
        """core.py:491-520."""
        intA = self.cal_integral_within_contours(contour, integrand=integrand)
        if area is None:
            area = self.cal_integral_within_contours(contour)
        lmA = self.cal_gradient_wrt_area(intA, area)
        return lmA.rename('lwm' if integrand.name is None else 'lwm' + integrand.name)

    def cal_contour_weigh_mean_hist(self, contour, integrand, area=None):
        """core.py:523-552."""
        intA = self.cal_integral_within_contours_hist(contour, integrand=integrand)
        if area is None:
            area = self.cal_integral_within_contours_hist(contour)
        lmA = self.cal_gradient_wrt_area(intA, area)
        return lmA.rename('lwm' if integrand.name is None else 'lwm' + integrand.name)

    def cal_contour_mean(self, contour, integrand, grdm, area=None):
        """core.py:555-583."""
        upper = self.cal_contour_weigh_mean(contour, integrand * grdm, area=area)
        lower = self.cal_contour_weigh_mean(contour, grdm, area=area)
        lmA = upper / lower
        return lmA.rename('cm' if integrand.name is None else 'cm' + integrand.name)

    def cal_contour_mean_hist(self, contour, integrand, grdm, area=None):
        """core.py:586-616."""
        upper = self.cal_contour_weigh_mean_hist(contour, integrand * grdm, area=area)
        lower = self.cal_contour_weigh_mean_hist(contour, grdm, area=area)
        lmA = upper / lower
        return lmA.rename('cm' if integrand.name is None else 'cm' + integrand.name)

    # ---------------------------------------------------------------------- LWA
    def _lwa(self, q, Q, mask_idx, part, variant):
        part = part.lower()
        if part not in ['all', 'upper', 'lower']:
            raise Exception('invalid part, should be in [\'all\', \'upper\', \'lower\']')
        eqDim = xc.coord(q, self.dimEqV)
        n_eq = q.shape[q.dims.index(self.dimEqV)]
        if mask_idx is not None and max(mask_idx) >= n_eq:
            raise Exception('indices in mask_idx out of boundary')
        qt, lead, plane = self._tracer_dev(q)
        S = qt.shape[0]
        eq_first = plane[0] == self.dimEqV
        dA_src, _ = self._dA_plane(plane, q)
        dA_dev = dA_src
        if not eq_first:                      # kernels want the equivalent dim first
            qt = qt.transpose(1, 2).contiguous()
            dA_dev = dA_dev.transpose(0, 1).contiguous()
        key = ('ww', eq_first) + tuple(plane) + tuple(dA_src.shape)
        hit = self._cache.get(key)
        if hit is None or hit[0] is not dA_src:          # rebuilt whenever the dA upload was
            w = ops.lwa_weights(dA_dev.reshape(-1))
            hit = self._cache[key] = (dA_src, w, ops.row_constant(w, dA_dev.shape[0], dA_dev.shape[1]))
        ww, ww_row = hit[1], hit[2]
        # the sorted profile, [S, n_eq] fp64
        Qlead = [d for d in Q.dims if d != self.dimEqV]
        Qv = np.asarray(Q.values, dtype=np.float64)
        Qv = np.transpose(Qv, [Q.dims.index(d) for d in Qlead + [self.dimEqV]]).reshape(-1, n_eq)
        if Qv.shape[0] != S:
            Qv = np.broadcast_to(Qv, (S, n_eq))
        Qt = ops.to_dev(np.ascontiguousarray(Qv))
        out = ops.lwa(qt, Qt, ww, self.increase, part, variant, ww_row=ww_row)

        def back(t, dtype=None):
            if not eq_first:
                t = t.transpose(1, 2)
            arr = ops.to_host(t)
            if dtype is not None:
                arr = arr.astype(dtype)
            lead_shape = tuple(q.shape[q.dims.index(d)] for d in lead)
            arr = arr.reshape(lead_shape + arr.shape[-2:])
            order = lead + plane
            if list(q.dims) != order:
                arr = np.transpose(arr, [order.index(d) for d in q.dims])
            return arr
        coords = xc.coords_for(q, q.dims)
        if eqDim is not None:
            coords[self.dimEqV] = eqDim
        LWA = xc.make(back(out), q.dims, coords, 'LWA')
        if mask_idx is None:
            return LWA
        contours, masks = [], []
        for j in range(n_eq):
            if j in mask_idx:
                contours.append(Q.isel({self.dimEqV: j}))
                m = ops.lwa_mask(qt, Qt, j, self.increase, variant)
                masks.append(xc.make(back(m, np.int64), q.dims, coords, None))
        return LWA, contours, masks

    def cal_local_wave_activity(self, q, Q, mask_idx=None, part='all'):
        """
        Local finite-amplitude wave activity density (Huang and Nakamura 2016),
        core.py:696-799.  Returns LWA, or (LWA, contours, masks) with mask_idx.
        """
        return self._lwa(q, Q, mask_idx, part, 1)

    def cal_local_wave_activity2(self, q, Q, mask_idx=None, part='all'):
        """Impulse-Casimir variant (point fixed, profile varies), core.py:802-905."""
        return self._lwa(q, Q, mask_idx, part, 2)

    def cal_local_APE(self, q, Q, mask_idx=None, part='all'):
        """Local APE density = LWA renamed 'LAPE' (core.py:908-942)."""
        if mask_idx is not None:
            LWA, contours, masks = self.cal_local_wave_activity(q, Q, mask_idx, part=part)
            return LWA.rename('LAPE'), contours, masks
        return self.cal_local_wave_activity(q, Q, None, part).rename('LAPE')

    # ------------------------------------------------------------- interpolation
    def interp_to_dataset(self, predef, dimEq, vs):
        """core.py:1017-1047."""
        re = []
        if isinstance(vs, Dataset):
            for var in vs:
                re.append(self.interp_to_coords(predef, dimEq, vs[var]).rename(var))
        else:
            for var in vs:
                re.append(self.interp_to_coords(predef, dimEq, var).rename(var.name))
        return merge(re)

    def interp_to_coords(self, predef, eqCoords, var, interpDim='contour'):
        """
        Interpolate a variable from the contour dimension to predefined
        coordinates along the equivalent dimension (core.py:1050-1100).
        """
        dimTmp = 'new'
        if isinstance(predef, (np.ndarray, list)):
            pvals = np.asarray(predef)
        else:
            dimTmp = predef.dims[0]
            pvals = np.asarray(predef.values)
        # slices are paired by dimension name (apply_ufunc's broadcasting), eqCoords' dims first
        e, v, lead, lshape = self._align2(eqCoords, var, interpDim)
        increasing = bool(e[0, 0] < e[0, -1])                   # core.py:1080-1088
        et = ops.to_dev(e.astype(np.float64))
        vt = ops.to_dev(v.astype(np.float64))
        out = ops.interp(ops.to_dev(pvals.astype(np.float64)), et if e.shape[0] > 1 else et[0],
                         vt if v.shape[0] > 1 else vt[0], reverse=0 if increasing else 1)
        res = out.cpu().numpy().reshape(tuple(lshape) + (pvals.shape[0],))
        coords = xc.coords_for(var, [d for d in lead if d in var.dims])
        coords.update(xc.coords_for(eqCoords, [d for d in lead if d in eqCoords.dims]))
        coords[dimTmp] = pvals
        return xc.make(res, lead + [dimTmp], coords, var.name)


class Table(object):
    """
    One-to-one mapping table between two monotonic quantities, y = F(x)
    (core.py:1103-1195).
    """

    def __init__(self, table, dimEq):
        tv = np.asarray(table.values)
        tmp = tv[..., -1] > tv[..., 0]
        if np.all(tmp):
            areaInc = True
        elif not np.any(tmp):
            areaInc = False
        else:
            raise Exception('not every time or level is increasing/decreasing')
        self._table = table
        self._coord = table[dimEq]
        self._dimEq = dimEq
        self._incVl = areaInc
        cv = np.asarray(self._coord.values)
        self._incCd = bool(cv[-1] > cv[0])

    def _interp(self, x, xf, yf, inc):
        xv = np.asarray(getattr(x, 'values', x), dtype=np.float64)
        flat = xv.reshape(-1, xv.shape[-1]) if xv.ndim >= 1 else xv.reshape(1, 1)
        out = ops.interp(ops.to_dev(np.ascontiguousarray(flat)),
                         ops.to_dev(np.asarray(xf, dtype=np.float64)),
                         ops.to_dev(np.asarray(yf, dtype=np.float64)), reverse=0 if inc else 1)
        return out.cpu().numpy().reshape(xv.shape)

    def lookup_coordinates(self, values):
        """For y = F(x), get coordinates (x) given values (y) (core.py:1136-1174)."""
        res = self._interp(values, self._table.values, self._coord.values, self._incVl)
        if xc.is_labeled(values):
            return xc.make(res, values.dims, xc.coords_for(values, values.dims), values.name)
        return res

    def lookup_values(self, coords):
        """For y = F(x), get values (y) given coordinates (x) (core.py:1176-1195;
        the reference reads an undefined attribute there -- this is the evident
        intent)."""
        res = self._interp(coords, self._coord.values, self._table.values, self._incCd)
        if xc.is_labeled(coords):
            return xc.make(res, coords.dims, xc.coords_for(coords, coords.dims), coords.name)
        return res


def _check_monotonicity(var, dim):
    """core.py:1328-1355: raise when any first difference along ``dim`` is zero."""
    ax = var.dims.index(dim)
    d = np.diff(np.asarray(var.values), axis=ax)
    if not d.all():
        pos = np.argwhere(d == 0)[0]
        raise Exception('not monotonic var at\n' + str(dict(zip(var.dims, pos.tolist()))))
