"""
A minimal stand-in for the slice of ``xarray`` that ``xcontour/core.py`` touches.

TEST INFRASTRUCTURE (lives under oracle/): it exists so that the reference's
*unmodified* source (``/root/reference/xcontour/core.py``) can be imported and
executed in this image, where xarray itself is absent, in order to generate golden
vectors for the oracle (``tests/golden/make_reference_golden.py``).  It is written
from xarray's documented behaviour, not from its source:

* dimensions are matched by NAME; binary operations broadcast to the ordered union
  of the operands' dims (left operand first) and keep the dim coordinates;
* ``sum`` / ``min`` / ``max`` / ``cumsum`` skip NaN for floating dtypes
  (``np.nansum`` & co.), ``where`` fills with NaN (integers are promoted to
  float64, float32 stays float32), ``differentiate`` is ``np.gradient`` against the
  coordinate with ``edge_order=1``;
* ``apply_ufunc(vectorize=True)`` moves the core dims last and wraps the function in
  ``np.vectorize(func, otypes=output_dtypes, signature=...)`` -- so the wrapped
  function sees NumPy scalars / 1-D core slices exactly as under xarray;
* ``concat`` along a new dim stacks it in front; ``broadcast`` orders dims by first
  appearance.

Anything the reference does not use is absent.  What this pins is the reference's
own logic (edge construction, flips, CDF direction, mask algebra, integration loops,
table end points, interpolation direction), executed verbatim; it cannot pin
xarray's internal summation order, which the oracle tolerances cover.
"""
import functools

import numpy as np


def _values(x):
    return x.values if isinstance(x, DataArray) else np.asarray(x)


class DataArray(object):
    __array_priority__ = 70

    def __init__(self, data, coords=None, dims=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            coords = data._coords if coords is None else coords
            dims = data.dims if dims is None else dims
            name = data.name if name is None else name
            data = data.values
        data = np.asarray(data)
        if dims is None:
            dims = tuple("dim_%d" % i for i in range(data.ndim))
        if isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        if len(dims) != data.ndim:
            raise ValueError("dims %r do not match shape %r" % (dims, data.shape))
        self._data = data
        self.dims = dims
        self.name = name
        self.attrs = dict(attrs or {})
        self._coords = {}
        for k, v in dict(coords or {}).items():
            v = _values(v)
            if k in dims:
                if v.shape != (data.shape[dims.index(k)],):
                    raise ValueError("coordinate %r has shape %r" % (k, v.shape))
                self._coords[k] = v
            elif v.ndim == 0:
                self._coords[k] = v

    # ---- protocol ------------------------------------------------------------
    values = property(lambda self: self._data)
    data = property(lambda self: self._data)
    shape = property(lambda self: self._data.shape)
    dtype = property(lambda self: self._data.dtype)
    ndim = property(lambda self: self._data.ndim)
    size = property(lambda self: self._data.size)
    sizes = property(lambda self: dict(zip(self.dims, self._data.shape)))

    @property
    def coords(self):
        return _Coords(self)

    def __len__(self):
        return self._data.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    def __bool__(self):
        return bool(self._data)

    def __float__(self):
        return float(self._data)

    def __int__(self):
        return int(self._data)

    def __index__(self):
        return int(self._data)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __repr__(self):
        return "<refshim.DataArray %r (%s)>\n%r" % (
            self.name, ", ".join("%s: %d" % kv for kv in zip(self.dims, self.shape)), self._data)

    def item(self):
        return self._data.item()

    def __getattr__(self, k):                     # arr.time, arr.latitude ...
        if k.startswith("_"):
            raise AttributeError(k)
        try:
            return self.__getitem__(k)
        except KeyError:
            raise AttributeError(k)

    def _replace(self, data, dims=None, coords=None, name="__keep__"):
        dims = self.dims if dims is None else tuple(dims)
        if coords is None:
            coords = self._coords
        coords = {k: v for k, v in coords.items() if k in dims or np.ndim(v) == 0}
        return DataArray(data, coords, dims, self.name if name == "__keep__" else name, self.attrs)

    def copy(self, deep=True):
        return DataArray(self._data.copy(), {k: np.copy(v) for k, v in self._coords.items()},
                         self.dims, self.name, self.attrs)

    def load(self):
        return self

    compute = load

    def astype(self, dtype):
        return self._replace(self._data.astype(dtype))

    # ---- indexing ------------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self._coords:
                c = self._coords[key]
                return DataArray(c, {key: c}, (key,) if c.ndim == 1 else (), key)
            if key in self.dims:
                n = self.shape[self.dims.index(key)]
                return DataArray(np.arange(n), None, (key,), key)
            raise KeyError(key)
        if isinstance(key, dict):
            return self.isel(key)
        if not isinstance(key, tuple):
            key = (key,)
        return self.isel(dict(zip(self.dims, key)))

    def __setitem__(self, key, value):
        if isinstance(key, str):
            self._coords[key] = _values(value)
            return
        if isinstance(key, dict):
            idx = tuple(key.get(d, slice(None)) for d in self.dims)
        else:
            idx = key
        self._data[idx] = _values(value)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        idx, dims, coords = [], [], {}
        for d in self.dims:
            k = indexers.get(d, slice(None))
            if isinstance(k, DataArray):
                k = k.values
            if isinstance(k, np.ndarray) and k.ndim == 0:
                k = int(k)
            idx.append(k)
            if not isinstance(k, (int, np.integer)):
                dims.append(d)
            if d in self._coords:
                coords[d] = self._coords[d][k]        # an integer leaves a scalar coordinate
        for k, v in self._coords.items():
            if v.ndim == 0:
                coords.setdefault(k, v)
        return DataArray(self._data[tuple(idx)], coords, dims, self.name, self.attrs)

    def squeeze(self, dim=None, drop=False):
        keep = [i for i, n in enumerate(self.shape) if n != 1]
        dims = tuple(self.dims[i] for i in keep)
        coords = {}
        for k, v in self._coords.items():
            if k in dims or v.ndim == 0:
                coords[k] = v
            elif v.size == 1:
                coords[k] = v.reshape(())
        return DataArray(self._data.reshape([self.shape[i] for i in keep]), coords, dims,
                         self.name, self.attrs)

    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        return self._replace(self._data.transpose([self.dims.index(d) for d in dims]), dims)

    def rename(self, new_name_or_name_dict=None, **names):
        new = new_name_or_name_dict
        if new is None or isinstance(new, dict):
            mp = dict(new or {}, **names)
            return DataArray(self._data, {mp.get(k, k): v for k, v in self._coords.items()},
                             tuple(mp.get(d, d) for d in self.dims), self.name, self.attrs)
        return self._replace(self._data, name=new)

    def assign_coords(self, coords=None, **kw):
        out = self._replace(self._data)
        for k, v in dict(coords or {}, **kw).items():
            v = _values(v)
            if k in out.dims and v.shape != (out.shape[out.dims.index(k)],):
                raise ValueError("coordinate %r has the wrong length" % k)
            out._coords[k] = v
        return out

    def broadcast_like(self, other):
        return broadcast(other, self)[1].transpose(*broadcast(other, self)[0].dims)

    # ---- arithmetic: broadcasting by dimension name --------------------------
    def _binary(self, other, op, reflexive=False):
        if isinstance(other, DataArray):
            dims, (x, y) = _align_by_name([self, other])
            coords = _merge_coords([self, other], dims)
            name = self.name if self.name == other.name else None
        else:
            dims, x, y, coords, name = self.dims, self._data, other, dict(self._coords), self.name
        with np.errstate(all="ignore"):
            res = op(y, x) if reflexive else op(x, y)
        return DataArray(res, coords, dims, name)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __rsub__(self, o): return self._binary(o, np.subtract, True)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __truediv__(self, o): return self._binary(o, np.true_divide)
    def __rtruediv__(self, o): return self._binary(o, np.true_divide, True)
    def __pow__(self, o): return self._binary(o, np.power)
    def __lt__(self, o): return self._binary(o, np.less)
    def __le__(self, o): return self._binary(o, np.less_equal)
    def __gt__(self, o): return self._binary(o, np.greater)
    def __ge__(self, o): return self._binary(o, np.greater_equal)
    def __eq__(self, o): return self._binary(o, np.equal)
    def __ne__(self, o): return self._binary(o, np.not_equal)
    def __and__(self, o): return self._binary(o, np.logical_and)
    def __or__(self, o): return self._binary(o, np.logical_or)
    __hash__ = None
    def __neg__(self): return self._replace(-self._data)
    def __abs__(self): return self._replace(np.abs(self._data))
    def __invert__(self): return self._replace(~self._data)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        arrs = [a for a in inputs if isinstance(a, DataArray)]
        dims, datas = _align_by_name(arrs)
        it = iter(datas)
        args = [next(it) if isinstance(a, DataArray) else a for a in inputs]
        with np.errstate(all="ignore"):
            res = ufunc(*args, **kwargs)
        names = set(a.name for a in arrs)
        return DataArray(res, _merge_coords(arrs, dims), dims, names.pop() if len(names) == 1 else None)

    # ---- masking / reductions ------------------------------------------------
    def where(self, cond, other=None):
        return _where3(cond, self, np.nan if other is None else other, dims_first=self)

    def fillna(self, value):
        return self._replace(np.where(np.isnan(self._data), value, self._data).astype(self._data.dtype))

    def isnull(self):
        return self._replace(np.isnan(self._data))

    def notnull(self):
        return self._replace(~np.isnan(self._data))

    def _axes(self, dim):
        if dim is None:
            return tuple(range(self.ndim))
        dim = [dim] if isinstance(dim, str) else list(dim)
        for d in dim:
            if d not in self.dims:
                raise ValueError("%r not found in array dimensions %r" % (d, self.dims))
        return tuple(self.dims.index(d) for d in dim)

    def _reduce(self, fn_nan, fn_plain, dim):
        axes = self._axes(dim)
        dims = tuple(d for i, d in enumerate(self.dims) if i not in axes)
        fn = fn_nan if self._data.dtype.kind in "fc" else fn_plain      # skipna only for floats
        with np.errstate(all="ignore"):
            return self._replace(fn(self._data, axis=axes), dims)

    def sum(self, dim=None, **kw): return self._reduce(np.nansum, np.sum, dim)
    def min(self, dim=None, **kw): return self._reduce(np.nanmin, np.min, dim)
    def max(self, dim=None, **kw): return self._reduce(np.nanmax, np.max, dim)
    def mean(self, dim=None, **kw): return self._reduce(np.nanmean, np.mean, dim)

    def all(self, dim=None):
        return self._reduce(np.all, np.all, dim)

    def any(self, dim=None):
        return self._reduce(np.any, np.any, dim)

    def cumsum(self, dim=None, **kw):
        ax = self.dims.index(dim)
        fn = np.nancumsum if self._data.dtype.kind in "fc" else np.cumsum
        return self._replace(fn(self._data, axis=ax))

    def diff(self, dim, n=1):
        ax = self.dims.index(dim)
        coords = dict(self._coords)
        if dim in coords:
            coords[dim] = coords[dim][1:]
        return DataArray(np.diff(self._data, axis=ax), coords, self.dims, self.name, self.attrs)

    def differentiate(self, coord, edge_order=1):
        ax = self.dims.index(coord)
        return self._replace(np.gradient(self._data, self._coords[coord], edge_order=edge_order, axis=ax))

    def argmax(self, dim=None):
        raise NotImplementedError("only reached when the reference raises 'not monotonic'")


class _Coords(object):
    def __init__(self, arr):
        self._a = arr

    def __getitem__(self, k):
        return self._a[k]

    def __setitem__(self, k, v):
        self._a._coords[k] = _values(v)

    def __contains__(self, k):
        return k in self._a._coords

    def __iter__(self):
        return iter(self._a._coords)

    def keys(self):
        return self._a._coords.keys()


def _align_by_name(arrs):
    """Ordered union of dims (first appearance) and the operands reshaped to it."""
    dims = []
    for a in arrs:
        for d in a.dims:
            if d not in dims:
                dims.append(d)
    out = []
    for a in arrs:
        order = [d for d in dims if d in a.dims]
        x = a.values.transpose([a.dims.index(d) for d in order])
        out.append(x.reshape([x.shape[order.index(d)] if d in order else 1 for d in dims]))
    sizes = {}
    for a in arrs:
        for d, n in zip(a.dims, a.shape):
            if sizes.setdefault(d, n) != n:
                raise ValueError("size mismatch along %r" % d)
    return tuple(dims), out


def _merge_coords(arrs, dims):
    coords = {}
    for a in arrs:
        for k, v in a._coords.items():
            if k in dims and np.ndim(v) == 1:       # a scalar coordinate never shadows a dim coordinate
                coords.setdefault(k, v)
    return coords


def _where3(cond, x, y, dims_first=None):
    arrs = [a for a in ((dims_first,) if dims_first is not None else ()) + (cond, x, y)
            if isinstance(a, DataArray)]
    dims, datas = _align_by_name(arrs)
    lut = {id(a): d for a, d in zip(arrs, datas)}
    c, xx, yy = [lut[id(a)] if isinstance(a, DataArray) else a for a in (cond, x, y)]
    res = np.where(c, xx, yy)           # NaN fill: integers -> float64, float32 stays float32
    name = x.name if isinstance(x, DataArray) else None
    return DataArray(res, _merge_coords(arrs, dims), dims, name)


def where(cond, x, y):
    return _where3(cond, x, y)


def broadcast(*args):
    dims, datas = _align_by_name(args)
    sizes = {}
    for a in args:
        sizes.update(a.sizes)
    shape = [sizes[d] for d in dims]
    coords = _merge_coords(args, dims)
    return tuple(DataArray(np.broadcast_to(d, shape).copy(), coords, dims, a.name) for a, d in zip(args, datas))


def concat(objs, dim):
    objs = list(objs)
    first = objs[0]
    if dim in first.dims:
        ax = first.dims.index(dim)
        data = np.concatenate([o.values for o in objs], axis=ax)
        coords = dict(first._coords)
        if dim in coords:
            coords[dim] = np.concatenate([o._coords[dim] for o in objs])
        return DataArray(data, coords, first.dims, first.name)
    data = np.stack([o.transpose(*first.dims).values for o in objs], axis=0)
    coords = {k: v for k, v in first._coords.items() if k in first.dims}
    if all(dim in o._coords for o in objs):
        coords[dim] = np.array([o._coords[dim] for o in objs])
    return DataArray(data, coords, (dim,) + first.dims, first.name)


class Dataset(object):
    def __init__(self, data_vars=None):
        self.data_vars = {}
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        self.data_vars[k] = v.rename(k)

    def __getitem__(self, k):
        if k in self.data_vars:
            return self.data_vars[k]
        for v in self.data_vars.values():
            if k in v._coords:
                return v[k]
        raise KeyError(k)

    def __getattr__(self, k):
        try:
            return self.__getitem__(k)
        except KeyError:
            raise AttributeError(k)

    def __iter__(self):
        return iter(self.data_vars)


def merge(objs):
    ds = Dataset()
    for o in objs:
        if isinstance(o, Dataset):
            for k in o:
                ds[k] = o[k]
        else:
            ds[o.name] = o
    return ds


def apply_ufunc(func, *args, input_core_dims=None, output_core_dims=((),), exclude_dims=frozenset(),
                vectorize=False, kwargs=None, dask="forbidden", output_dtypes=None, keep_attrs=None,
                join="exact", **unused):
    """The subset of xarray.apply_ufunc the reference uses: one output, core dims moved
    last, optional np.vectorize with a gufunc signature."""
    if input_core_dims is None:
        input_core_dims = [[] for _ in args]
    input_core_dims = [list(c) for c in input_core_dims]
    out_core = list(list(output_core_dims)[0])
    arrs = [a for a in args if isinstance(a, DataArray)]
    bdims, sizes = [], {}
    for a, core in zip(args, input_core_dims):
        if not isinstance(a, DataArray):
            continue
        for d, n in zip(a.dims, a.shape):
            if d in core:
                continue
            if d not in bdims:
                bdims.append(d)
            if sizes.setdefault(d, n) != n:
                raise ValueError("size mismatch along %r" % d)
    datas = []
    for a, core in zip(args, input_core_dims):
        if not isinstance(a, DataArray):
            datas.append(a)
            continue
        missing = [d for d in core if d not in a.dims]
        if missing:
            raise ValueError("core dims %r missing on an operand with dims %r" % (missing, a.dims))
        order = [d for d in bdims if d in a.dims] + core
        x = a.values.transpose([a.dims.index(d) for d in order])
        lead = [d for d in bdims if d in a.dims]
        shape = [x.shape[lead.index(d)] if d in lead else 1 for d in bdims] + list(x.shape[len(lead):])
        datas.append(x.reshape(shape))
    f = functools.partial(func, **kwargs) if kwargs else func
    if vectorize:
        ids = {}

        def tag(d, k):
            key = (d, k) if d in exclude_dims else d                 # excluded dims may differ in size
            return ids.setdefault(key, "d%d" % len(ids))
        sig = ",".join("(" + ",".join(tag(d, k) for d in core) + ")" for k, core in enumerate(input_core_dims))
        sig += "->(" + ",".join(tag(d, -1) for d in out_core) + ")"
        f = np.vectorize(f, otypes=output_dtypes, signature=sig)
    res = np.asarray(f(*datas))
    dims = tuple(bdims) + tuple(out_core)
    coords = {k: v for k, v in _merge_coords(arrs, dims).items() if k not in exclude_dims}
    names = set(a.name for a in arrs)
    return DataArray(res, coords, dims, names.pop() if len(names) == 1 else None)
