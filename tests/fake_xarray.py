"""A minimal stand-in for the slice of xarray's API that discontinuum_b200's optional adapter touches
(DataArray: values / attrs / name / dims / ndim / coords / assign_coords; Dataset: coords / data_vars / __getitem__).

xarray cannot be installed in this environment (no index access), so the adapter branches of data.py / engine.py are
exercised against this stand-in: it checks that those branches run and put values, coordinates and attributes where the
reference puts them (src/discontinuum/engines/gpytorch.py:496-499,583-591; data_manager.py:101-103), not xarray itself."""
import numpy as np


class DataArray:
    def __init__(self, data, coords=None, dims=None, attrs=None, name=None):
        self.values = np.asarray(data)
        if self.values.dtype == object and self.values.size and hasattr(self.values.flat[0], "to_datetime64"):
            self.values = np.array([v.to_datetime64() for v in self.values.flat], dtype="datetime64[ns]").reshape(self.values.shape)
        self.attrs = dict(attrs or {})
        self.name = name
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(self.values.ndim))
        self.dims = tuple(dims)
        if isinstance(coords, (list, tuple)):
            coords = {d: c for d, c in zip(self.dims, coords)}
        self.coords = {k: (v if isinstance(v, DataArray) else DataArray(v, dims=(k,))) for k, v in (coords or {}).items()}

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def data(self):
        return self.values

    @property
    def shape(self):
        return self.values.shape

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    # -- what the reference's plot mixins touch beyond the adapter: element-wise arithmetic, da["coord"], da.plot.<kind>(...)
    def _binary(self, other, op):
        o = other.values if isinstance(other, DataArray) else other
        return DataArray(op(self.values, o), coords=dict(self.coords), dims=self.dims, attrs=self.attrs, name=self.name)

    def __add__(self, o): return self._binary(o, np.add)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __truediv__(self, o): return self._binary(o, np.divide)
    def __pow__(self, o): return self._binary(o, np.power)
    def __radd__(self, o): return self._binary(o, lambda a, b: b + a)
    def __rmul__(self, o): return self._binary(o, lambda a, b: b * a)
    def __rsub__(self, o): return self._binary(o, lambda a, b: b - a)
    def __rtruediv__(self, o): return self._binary(o, lambda a, b: b / a)

    def __getitem__(self, key):
        return self.coords[key]

    def __getattr__(self, name):  # da.time: attribute-style access to a coordinate
        coords = self.__dict__.get("coords", {})
        if name in coords:
            return coords[name]
        raise AttributeError(name)

    def min(self):
        import types
        return types.SimpleNamespace(values=self.values.min())   # (.values of a reduced array: a numpy scalar)

    def max(self):
        import types
        return types.SimpleNamespace(values=self.values.max())

    def __len__(self):
        return len(self.values)

    @property
    def plot(self):
        from unittest.mock import MagicMock
        if self.__dict__.get("_plot") is None:
            self.__dict__["_plot"] = MagicMock(name="DataArray.plot")
        return self.__dict__["_plot"]

    def assign_coords(self, coords):
        out = DataArray(self.values, coords=dict(self.coords), dims=self.dims, attrs=self.attrs, name=self.name)
        out.coords.update({k: (v if isinstance(v, DataArray) else DataArray(v, dims=(k,))) for k, v in dict(coords).items()})
        return out


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.coords = {k: (v if isinstance(v, DataArray) else DataArray(v, dims=(k,))) for k, v in (coords or {}).items()}
        self.data_vars = {}
        for k, v in (data_vars or {}).items():
            if isinstance(v, tuple):  # (dims, values) as in xarray
                v = DataArray(v[1], dims=v[0] if isinstance(v[0], (list, tuple)) else (v[0],))
            self.data_vars[k] = v if isinstance(v, DataArray) else DataArray(v)
        self.attrs = dict(attrs or {})

    def __getitem__(self, key):
        return self.data_vars[key] if key in self.data_vars else self.coords[key]

    def __getattr__(self, name):  # ds.time, ds.stage: attribute-style access to coordinates and variables
        d = self.__dict__
        for table in ("data_vars", "coords"):
            if table in d and name in d[table]:
                return d[table][name]
        raise AttributeError(name)

    @property
    def plot(self):
        from unittest.mock import MagicMock
        if self.__dict__.get("_plot") is None:
            self.__dict__["_plot"] = MagicMock(name="Dataset.plot")
        return self.__dict__["_plot"]

    def __iter__(self):          # a Dataset iterates over the names of its data variables
        return iter(self.data_vars)

    def __len__(self):
        return len(self.data_vars)
