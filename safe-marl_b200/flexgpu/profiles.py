"""Profile datasets: load / PV / price time series at the env's sample interval.

`load_csv_profiles` follows the reference's loaders (flexibility_provision_env.py:431-471:
read_csv -> scale -> resample(sample_interval).mean() -> linear interpolate).  The bundled
CSVs are Git-LFS payloads, so benchmarks and tests use `synthetic_profiles`, which has the
bundled data's shape (~3 years at 15 min = 105 216 rows; 32 load buses, 5 PV
plants, one price column).
"""
import os

import numpy as np


class Profiles:
    """P[T, nl], Q[T, nl] (non-slack buses, bus order), PV[T, na], price[T]; all fp64, C order."""

    def __init__(self, P, Q, PV, price, time_delta_min=15):
        self.P = np.ascontiguousarray(P, dtype=np.float64)
        self.Q = np.ascontiguousarray(Q, dtype=np.float64)
        self.PV = np.ascontiguousarray(PV, dtype=np.float64)
        self.price = np.ascontiguousarray(np.asarray(price).reshape(-1), dtype=np.float64)
        self.time_delta = int(time_delta_min)
        T = self.P.shape[0]
        if not (self.Q.shape[0] == T and self.PV.shape[0] == T and self.price.shape[0] == T):
            raise ValueError("profile arrays must have the same number of rows")
        if self.P.shape != self.Q.shape:
            raise ValueError("P and Q must have the same shape")

    @property
    def T(self):
        return self.P.shape[0]

    def n_days(self):
        """(index[-1] - index[0]).days of a uniform index (flexibility_provision_env.py:421)."""
        return ((self.T - 1) * self.time_delta) // (24 * 60)

    def as_dict(self):
        return dict(P=self.P, Q=self.Q, PV=self.PV, price=self.price)


def synthetic_profiles(network, n_agents=5, T=105216, seed=0, pv_scale=0.15):
    """Synthetic data of the bundled data's shape."""
    rng = np.random.default_rng(seed)
    nl = network.n_bus - 1
    t = np.arange(T)
    hour = (t % 96) / 4.0
    # double-peak diurnal demand curve in [0, 1]
    d = 0.55 * np.exp(-0.5 * ((hour - 8.0) / 2.0) ** 2) + 1.0 * np.exp(-0.5 * ((hour - 19.0) / 2.5) ** 2)
    d = np.clip(d, 0.0, 1.0)
    level = (0.35 + 0.65 * d)[:, None]
    P = network.base_p[None, 1:] * level * (1.0 + 0.1 * rng.standard_normal((T, nl)))
    Q = network.base_q[None, 1:] * level * (1.0 + 0.1 * rng.standard_normal((T, nl)))
    sun = np.maximum(0.0, np.sin(np.pi * (hour - 6.0) / 12.0))[:, None]
    PV = pv_scale * sun * rng.uniform(0.3, 1.0, size=(T, n_agents))
    price = 0.05 + 0.25 * rng.uniform(0.0, 1.0, size=T)
    return Profiles(P, Q, PV, price, 15)


CSV_NAMES = ('pv_active.csv', 'load_active.csv', 'load_reactive.csv', 'prices.csv')


def _is_lfs_pointer(path):
    try:
        with open(path, 'rb') as f:
            return f.read(40).startswith(b'version https://git-lfs')
    except OSError:
        return False


def csv_profiles_available(data_path):
    """True when the four profile CSVs the reference loads (:431-465) exist under data_path as real files
    (the bundled ones are Git-LFS pointers until `git lfs pull` has run)."""
    if not data_path:
        return False
    paths = [os.path.join(data_path, n) for n in CSV_NAMES]
    return all(os.path.isfile(p) and not _is_lfs_pointer(p) for p in paths)


def load_csv_profiles(data_path, args):
    """The reference's four loaders (:431-465) + resample_data (:467-471): read_csv, first column -> time index,
    scale, resample(sample_interval).mean(), linear interpolate.  The reference keeps four frames of possibly
    different lengths and slices all of them with the same row offsets (:473-547); here they are cut to the
    shortest one, which is every row the reference can address in all four."""
    import pandas as pd

    def load(name, scale):
        df = pd.read_csv(os.path.join(data_path, name), index_col=None)
        df.index = pd.to_datetime(df.iloc[:, 0])
        df.index.name = 'time'
        df = df.iloc[::1, 1:] * scale
        df = df.resample(args["sample_interval"]).mean()
        return df.interpolate(method='linear')

    pv = load('pv_active.csv', args["pv_scale"])
    p = load('load_active.csv', args["demand_scale"])
    q = load('load_reactive.csv', args["reactive_scale"])
    price = load('prices.csv', 1.0)
    delta = (pv.index[1] - pv.index[0]).seconds // 60          # :422
    T = min(len(pv), len(p), len(q), len(price))
    return Profiles(p.values[:T], q.values[:T], pv.values[:T], price.values[:T, 0], delta)
