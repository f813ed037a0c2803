"""
A minimal labelled-array container with the slice of the xarray.DataArray /
Dataset interface that the Contour2D workflow touches (dims, coords, name,
.values, isel, rename, squeeze, transpose, where, broadcasting arithmetic ...).

xarray is not installed in this image, and the reference's return convention is
"xarray DataArrays" (SURVEY.md §8b).  ``xcontour_b200.xr_compat`` therefore uses
real xarray when it is importable and this stand-in otherwise; Contour2D only
talks to the small helper API in xr_compat, so both behave the same.  This
module does host-side bookkeeping only -- no hot-path arithmetic lives here.
"""
import numpy as np


def _as_np(x):
    if isinstance(x, DataArray):
        return x.values
    return np.asarray(x)


def _device_array(x):
    """torch.Tensor for data that lives in (or is only reachable through) torch / DLPack, else None.
    Such data is kept as it is -- Contour2D hands it to the kernels without a host round trip -- and
    is only copied to a NumPy array when somebody asks for ``.values``."""
    if isinstance(x, (np.ndarray, np.generic, list, tuple, float, int)):
        return None
    if type(x).__module__.split(".")[0] == "torch":
        return x
    if hasattr(x, "__dlpack__") and not hasattr(x, "__array__") and not hasattr(x, "__array_interface__"):
        import torch
        return torch.from_dlpack(x)
    return None


class DataArray(object):
    def __init__(self, data, dims=None, coords=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            dims = data.dims if dims is None else dims
            coords = dict(data.coords) if coords is None else coords
            name = data.name if name is None else name
            data = data.values if data._raw is None else data._raw
        self._raw = _device_array(data)            # torch tensor (any device) or None
        self._np = None if self._raw is not None else np.asarray(data)
        data = self._raw if self._raw is not None else self._np
        if dims is None:
            dims = tuple("dim_%d" % i for i in range(data.ndim))
        if isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        if len(dims) != len(data.shape):
            raise ValueError("dims %s do not match data of shape %s" % (dims, tuple(data.shape)))
        self.dims = dims
        self.name = name
        self.attrs = dict(attrs or {})
        self.coords = {}
        for k, v in (coords or {}).items():
            v = _as_np(v)
            if v.ndim == 1 and k in dims and v.shape[0] != data.shape[dims.index(k)]:
                raise ValueError("coordinate %s has the wrong length" % k)
            if v.ndim == 0 or k in dims:
                self.coords[k] = v

    # ---- basic protocol ----------------------------------------------------
    @property
    def _data(self):
        if self._np is None:                       # device-backed: materialise on first host access
            from . import ops
            self._np = ops.to_host(self._raw)
        return self._np

    @property
    def values(self):
        return self._data

    @property
    def data(self):
        return self._data if self._raw is None else self._raw

    @property
    def shape(self):
        return tuple(self._raw.shape) if self._raw is not None else self._np.shape

    @property
    def dtype(self):
        if self._raw is not None and self._np is None:
            return np.dtype(str(self._raw.dtype).replace("torch.", ""))
        return self._data.dtype

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    def __repr__(self):
        dims = ", ".join("%s: %d" % (d, n) for d, n in zip(self.dims, self.shape))
        return "<xcontour_b200.DataArray %r (%s)>\n%r" % (self.name, dims, self._data)

    def _new(self, data, dims=None, coords=None, name="__keep__"):
        dims = self.dims if dims is None else dims
        if coords is None:
            coords = {k: v for k, v in self.coords.items() if k in dims or np.ndim(v) == 0}
        return DataArray(data, dims, coords, self.name if name == "__keep__" else name, self.attrs)

    def copy(self):
        return self._new(self._data.copy(), coords={k: np.copy(v) for k, v in self.coords.items()})

    def load(self):
        return self

    def astype(self, dtype):
        return self._new(self._data.astype(dtype))

    def item(self):
        return self._data.item()

    def __bool__(self):
        return bool(self._data)

    def __float__(self):
        return float(self._data)

    # ---- indexing ----------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self.coords:
                c = self.coords[key]
                return DataArray(c, (key,) if np.ndim(c) == 1 else (), {key: c}, key)
            if key in self.dims:
                n = self.shape[self.dims.index(key)]
                return DataArray(np.arange(n), (key,), {}, key)
            raise KeyError(key)
        if isinstance(key, dict):
            return self.isel(key)
        if not isinstance(key, tuple):
            key = (key,)
        key = key + (slice(None),) * (self.ndim - len(key))
        return self.isel({d: k for d, k in zip(self.dims, key)})

    def __setitem__(self, key, value):
        if isinstance(key, str):
            self.coords[key] = _as_np(value)
            return
        if isinstance(key, dict):
            idx = tuple(key.get(d, slice(None)) for d in self.dims)
        else:
            idx = key
        data = self._data
        self._raw = None                           # a host-side edit ends the device-backed life of the array
        data[idx] = _as_np(value)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        idx, dims, coords = [], [], {}
        for d in self.dims:
            k = indexers.get(d, slice(None))
            idx.append(k)
            scalar = isinstance(k, (int, np.integer))
            if not scalar:
                dims.append(d)
            if d in self.coords:
                coords[d] = self.coords[d][k]
        for k, v in self.coords.items():
            if np.ndim(v) == 0:
                coords.setdefault(k, v)
        return DataArray(self._data[tuple(idx)], dims, coords, self.name, self.attrs)

    def squeeze(self):
        keep = [i for i, n in enumerate(self.shape) if n != 1]
        dims = tuple(self.dims[i] for i in keep)
        coords = {}
        for k, v in self.coords.items():
            if k in dims or np.ndim(v) == 0:
                coords[k] = v
            elif np.size(v) == 1:
                coords[k] = np.asarray(v).reshape(())
        return DataArray(self._data.reshape([self.shape[i] for i in keep]), dims, coords,
                         self.name, self.attrs)

    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        return self._new(self._data.transpose([self.dims.index(d) for d in dims]), tuple(dims))

    def rename(self, new=None, **kw):
        if new is None or isinstance(new, dict):
            mp = dict(new or {}, **kw)
            dims = tuple(mp.get(d, d) for d in self.dims)
            coords = {mp.get(k, k): v for k, v in self.coords.items()}
            return DataArray(self._data, dims, coords, self.name, self.attrs)
        return self._new(self._data, name=new)

    def assign_coords(self, coords=None, **kw):
        out = self._new(self._data)
        for k, v in dict(coords or {}, **kw).items():
            out.coords[k] = _as_np(v)
        return out

    # ---- arithmetic with broadcasting by dimension name --------------------
    @staticmethod
    def _align(a, b):
        if not isinstance(b, DataArray):
            return a.dims, a.coords, a._data, np.asarray(b)
        dims = list(a.dims) + [d for d in b.dims if d not in a.dims]

        def expand(x):
            order = [d for d in dims if d in x.dims]
            arr = x._data.transpose([x.dims.index(d) for d in order])
            return arr.reshape([arr.shape[order.index(d)] if d in order else 1 for d in dims])
        coords = dict(b.coords)
        coords.update(a.coords)
        return tuple(dims), coords, expand(a), expand(b)

    def _binary(self, other, op, reflexive=False):
        dims, coords, x, y = DataArray._align(self, other)
        with np.errstate(invalid="ignore", divide="ignore"):
            res = op(y, x) if reflexive else op(x, y)
        return DataArray(res, dims, {k: v for k, v in coords.items() if k in dims or np.ndim(v) == 0},
                         self.name, self.attrs)

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
    __hash__ = None
    def __neg__(self): return self._new(-self._data)
    def __abs__(self): return self._new(np.abs(self._data))

    def where(self, cond, other=np.nan):
        dims, coords, x, c = DataArray._align(self, cond)
        return DataArray(np.where(c, x, other), dims, coords, self.name, self.attrs)

    def fillna(self, value):
        return self._new(np.where(np.isnan(self._data), value, self._data))

    def isnull(self):
        return self._new(np.isnan(self._data))

    def _reduce(self, fn, dim):
        if dim is None:
            axes = tuple(range(self.ndim))
        else:
            dim = [dim] if isinstance(dim, str) else list(dim)
            axes = tuple(self.dims.index(d) for d in dim)
        dims = tuple(d for i, d in enumerate(self.dims) if i not in axes)
        with np.errstate(invalid="ignore"):
            return self._new(fn(self._data, axis=axes), dims)

    def min(self, dim=None): return self._reduce(np.nanmin, dim)
    def max(self, dim=None): return self._reduce(np.nanmax, dim)
    def sum(self, dim=None): return self._reduce(np.nansum, dim)
    def mean(self, dim=None): return self._reduce(np.nanmean, dim)
    def all(self): return bool(np.all(self._data))
    def any(self): return bool(np.any(self._data))

    def diff(self, dim):
        ax = self.dims.index(dim)
        coords = dict(self.coords)
        if dim in coords:
            coords[dim] = coords[dim][1:]
        return DataArray(np.diff(self._data, axis=ax), self.dims, coords, self.name, self.attrs)


class Dataset(object):
    """An ordered name -> DataArray mapping with attribute access."""

    def __init__(self, data_vars=None):
        self.data_vars = {}
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        self.data_vars[k] = v if v.name == k else v.rename(k)

    def __getitem__(self, k):
        if k in self.data_vars:
            return self.data_vars[k]
        for v in self.data_vars.values():
            if k in v.coords:
                return v[k]
        raise KeyError(k)

    def __getattr__(self, k):
        try:
            return self.__getitem__(k)
        except KeyError:
            raise AttributeError(k)

    def __iter__(self):
        return iter(self.data_vars)

    def __contains__(self, k):
        return k in self.data_vars

    def __len__(self):
        return len(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def __repr__(self):
        return "<xcontour_b200.Dataset %s>" % ", ".join(self.data_vars)


def merge(arrays):
    ds = Dataset()
    for a in arrays:
        if isinstance(a, Dataset):
            for k in a:
                ds[k] = a[k]
        else:
            if a.name is None:
                raise ValueError("cannot merge an unnamed DataArray")
            ds[a.name] = a
    return ds


def where(cond, x, y):
    if isinstance(cond, DataArray):
        dims, coords, c, xx = DataArray._align(cond, x)
        yy = _as_np(y)
        return DataArray(np.where(c, xx, yy), dims, coords, getattr(x, "name", None))
    return np.where(cond, _as_np(x), _as_np(y))
