"""A minimal stand-in for the slice of xarray's API that discontinuum_b200's optional adapter touches
(DataArray: values / attrs / name / dims / ndim / coords / assign_coords; Dataset: coords / data_vars / __getitem__).

xarray cannot be installed in this environment (no index access), so the adapter branches of data.py / engine.py are
exercised against this stand-in: it checks that those branches run and put values, coordinates and attributes where the
reference puts them (src/discontinuum/engines/gpytorch.py:496-499,583-591; data_manager.py:101-103), not xarray itself."""
import numpy as np


class DataArray:
    def __init__(self, data, coords=None, dims=None, attrs=None, name=None):
        self.values = np.asarray(data)
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

    def __iter__(self):          # a Dataset iterates over the names of its data variables
        return iter(self.data_vars)

    def __len__(self):
        return len(self.data_vars)
