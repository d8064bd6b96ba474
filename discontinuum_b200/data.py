"""Array-level mirror of the reference's DataManager and pipelines (host side, O(n), float64).

The reference moves between raw and model space with sklearn pipelines wrapped around xarray objects
(src/discontinuum/data_manager.py:28-120, src/discontinuum/pipeline.py:14-403).  xarray is optional
here: covariates may be a dict of numpy arrays (or an xarray Dataset), targets numpy arrays (or
DataArrays); when xarray objects come in, xarray objects go out with coords/attrs restored.
The arithmetic of each pipeline follows the cited reference lines exactly.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

try:  # optional adapter
    import xarray as _xr
except Exception:  # noqa: BLE001
    _xr = None


def datetime_to_decimal_year(x: np.ndarray) -> np.ndarray:
    """src/discontinuum/pipeline.py:14-39 (year + day-of-year fraction with leap years), numpy only."""
    x = np.asarray(x)
    if not np.issubdtype(x.dtype, np.datetime64):
        raise ValueError("Array must contain numpy datetime64 objects.")
    x = x.reshape(-1).astype("datetime64[ns]")
    years = x.astype("datetime64[Y]")
    start = years.astype("datetime64[ns]")
    nxt = (years + np.timedelta64(1, "Y")).astype("datetime64[ns]")
    frac = (x - start) / (nxt - start)
    return years.astype(int) + 1970 + frac.astype(np.float64)


def decimal_year_to_datetime(x: np.ndarray) -> np.ndarray:
    """src/discontinuum/pipeline.py:42-66."""
    x = np.asarray(x, dtype=np.float64)
    year = np.floor(x).astype(int)
    rem = x - year
    start = (year - 1970).astype("datetime64[Y]").astype("datetime64[ns]")
    nxt = (year - 1969).astype("datetime64[Y]").astype("datetime64[ns]")
    ns = (rem * (nxt - start).astype(np.int64)).astype(np.int64)
    out = start + ns.astype("timedelta64[ns]")
    return out.astype("datetime64[s]").astype("datetime64[ns]")


class _Pipe:
    attrs: dict = {}
    name = None

    def fit(self, x):
        return self


class TimePipeline(_Pipe):
    """decimal year, mean removed (pipeline.py:290-299)."""

    def fit(self, x):
        self.mean_ = datetime_to_decimal_year(np.asarray(x)).mean()
        return self

    def transform(self, x):
        return datetime_to_decimal_year(np.asarray(x)) - self.mean_

    def inverse_transform(self, z):
        return decimal_year_to_datetime(np.asarray(z, dtype=np.float64) + self.mean_)


class LogStandardPipeline(_Pipe):
    """clip(1e-6) -> log -> standardise (pipeline.py:242-252)."""

    def fit(self, x):
        z = np.log(np.clip(np.asarray(x, dtype=np.float64), 1e-6, None))
        self.mean_, self.scale_ = z.mean(), z.std()
        return self

    def transform(self, x):
        return (np.log(np.clip(np.asarray(x, dtype=np.float64), 1e-6, None)) - self.mean_) / self.scale_

    def inverse_transform(self, z):
        return np.clip(np.exp(np.asarray(z, dtype=np.float64) * self.scale_ + self.mean_), 1e-6, None)


class StandardPipeline(_Pipe):
    """clip(0) -> standardise (pipeline.py:266-275)."""

    def fit(self, x):
        z = np.clip(np.asarray(x, dtype=np.float64), 0, None)
        self.mean_, self.scale_ = z.mean(), z.std()
        return self

    def transform(self, x):
        return (np.clip(np.asarray(x, dtype=np.float64), 0, None) - self.mean_) / self.scale_

    def inverse_transform(self, z):
        return np.clip(np.asarray(z, dtype=np.float64) * self.scale_ + self.mean_, 0, None)


class UnitPipeline(_Pipe):
    """clip(0) -> rescale to [1, 2] (pipeline.py:278-287, UnitScaler zero_value=1)."""

    def fit(self, x):
        z = np.clip(np.asarray(x, dtype=np.float64), 0, None)
        self.min_, self.max_ = z.min(), z.max()
        return self

    def transform(self, x):
        return 1.0 + (np.clip(np.asarray(x, dtype=np.float64), 0, None) - self.min_) / (self.max_ - self.min_)

    def inverse_transform(self, z):
        return np.clip(self.min_ + (np.asarray(z, dtype=np.float64) - 1.0) * (self.max_ - self.min_), 0, None)


class LogErrorPipeline(_Pipe):
    """GSE <-> model-space variance (pipeline.py:364-403): fit on the TARGET; transform(gse) = clip((log gse / s)^2, 1e-6);
    inverse_transform(var) = exp(sqrt(clip(var, 1e-6)) * s)."""

    def fit(self, target):
        self.scale_ = np.log(np.asarray(target, dtype=np.float64)).std()
        return self

    def transform(self, gse):
        return np.clip((np.log(np.asarray(gse, dtype=np.float64)) / self.scale_) ** 2, 1e-6, None)

    def inverse_transform(self, var):
        return np.exp(np.sqrt(np.clip(np.asarray(var, dtype=np.float64), 1e-6, None)) * self.scale_)

    def ci(self, mean, se, ci=0.95):
        from scipy.stats import norm

        cb = se ** norm.ppf(1 - (1 - ci) / 2)
        return mean / cb, mean * cb


class StandardErrorPipeline(_Pipe):
    """SE <-> model-space variance (pipeline.py:323-361)."""

    def fit(self, target):
        self.scale_ = np.asarray(target, dtype=np.float64).std()
        return self

    def transform(self, se):
        return np.clip((np.asarray(se, dtype=np.float64) / self.scale_) ** 2, 0, None)

    def inverse_transform(self, var):
        return np.sqrt(np.clip(np.asarray(var, dtype=np.float64), 0, None)) * self.scale_

    def ci(self, mean, se, ci=0.95):
        from scipy.stats import norm

        cb = se * norm.ppf(1 - (1 - ci) / 2)
        return mean - cb, mean + cb


def _values(obj):
    """numpy view of an array-like / xarray object."""
    if _xr is not None and isinstance(obj, (_xr.DataArray,)):
        return obj.values
    return np.asarray(obj)


def _cov_dict(covariates) -> Dict[str, np.ndarray]:
    """{name: 1-D array}, coordinates first then data variables (data_manager.py:105-120 ordering)."""
    if _xr is not None and isinstance(covariates, _xr.Dataset):
        out = {c: covariates.coords[c].values for c in covariates.coords}
        out.update({v: covariates[v].values for v in covariates.data_vars})
        return out
    return {k: np.asarray(v) for k, v in covariates.items()}


@dataclass
class Data:
    target: object
    covariates: object
    target_unc: object = None


class DataManager:
    """fit / X / y / y_unc / Xnew / y_t / get_dim with the reference's meaning (data_manager.py:28-120)."""

    def __init__(self, target_pipeline, error_pipeline, covariate_pipelines: Dict[str, type]):
        self.target_pipeline = target_pipeline
        self.error_pipeline = error_pipeline
        self.covariate_pipelines = dict(covariate_pipelines)
        self.data: Optional[Data] = None
        self._cache = {}

    def fit(self, target, covariates, target_unc=None):
        self.data = Data(target, covariates, target_unc)
        self._cache = {}
        cov = _cov_dict(covariates)
        if isinstance(self.target_pipeline, type):
            self.target_pipeline = self.target_pipeline().fit(_values(target))
        if isinstance(self.error_pipeline, type):
            self.error_pipeline = self.error_pipeline().fit(_values(target))
        for key, value in self.covariate_pipelines.items():
            if isinstance(value, type):
                self.covariate_pipelines[key] = value().fit(cov[key])

    def transform_covariates(self, covariates) -> np.ndarray:
        cov = _cov_dict(covariates)
        cols = [np.asarray(p.transform(cov[k]), dtype=np.float64).reshape(-1) for k, p in self.covariate_pipelines.items()]
        return np.ascontiguousarray(np.stack(cols, axis=-1))

    @property
    def X(self) -> np.ndarray:
        if "X" not in self._cache:
            self._cache["X"] = self.transform_covariates(self.data.covariates)
        return self._cache["X"]

    @property
    def y(self) -> np.ndarray:
        if "y" not in self._cache:
            self._cache["y"] = np.asarray(self.target_pipeline.transform(_values(self.data.target)), dtype=np.float64).reshape(-1)
        return self._cache["y"]

    @property
    def y_unc(self) -> np.ndarray:
        if "y_unc" not in self._cache:
            self._cache["y_unc"] = np.asarray(self.error_pipeline.transform(_values(self.data.target_unc)), dtype=np.float64).reshape(-1)
        return self._cache["y_unc"]

    def Xnew(self, covariates) -> np.ndarray:
        return self.transform_covariates(covariates)

    def y_t(self, y):
        """model space -> original units; DataArray when the training target was one."""
        out = self.target_pipeline.inverse_transform(np.asarray(y, dtype=np.float64))
        return self._wrap(out)

    def se_t(self, var):
        return self._wrap(self.error_pipeline.inverse_transform(np.asarray(var, dtype=np.float64)))

    def _wrap(self, arr):
        t = self.data.target if self.data is not None else None
        if _xr is not None and isinstance(t, _xr.DataArray):
            return _xr.DataArray(np.asarray(arr).squeeze(), attrs=t.attrs, name=t.name, dims=t.dims if np.ndim(arr) == t.ndim else None)
        return np.asarray(arr)

    def get_dim(self, dim: str) -> int:
        return list(_cov_dict(self.data.covariates)).index(dim)
