"""TEST INFRASTRUCTURE ONLY — golden vectors for the task statistics (tests/golden/stats/series_statistics.npz).

Feeds seeded series to the UNMODIFIED reference estimators — SampleMoments::NumberStatistics
(src/extra/sample-moments/NumberStatistics.h) and TimeAverage (src/tasks/auxiliary/time-average-aux.h:27-67),
through oracle/_ref/libljmd_ref.so (ref_tasks_shim.cpp: ljref_series_statistics) — and stores series + outputs.
Run in the build container (needs /root/reference):  python oracle/make_golden_stats.py
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "stats")


def series(seed):
    """Occupancy-like integer series with exponential memory, and real-valued ones (u*, p* like)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 4000))
    kind = seed % 4
    if kind == 0:      # binomial occupancy, uncorrelated
        x = rng.binomial(400, rng.uniform(0.05, 0.95), size=n).astype(np.float64)
    elif kind == 1:    # AR(1) occupancy, rounded
        rho, mu, sig = rng.uniform(0.3, 0.98), rng.uniform(20, 300), rng.uniform(2, 12)
        e = rng.normal(size=n)
        y = np.empty(n)
        y[0] = e[0]
        for k in range(1, n):
            y[k] = rho * y[k - 1] + np.sqrt(1 - rho * rho) * e[k]
        x = np.rint(mu + sig * y)
    elif kind == 2:    # energy per particle: small fluctuations around a negative mean
        x = -3.2 + 0.01 * np.cumsum(rng.normal(size=n)) / np.sqrt(np.arange(1, n + 1)) + 0.02 * rng.normal(size=n)
    else:              # anticorrelated series (negative lag-one covariance: the s = NaN / s = 1 branches)
        x = 5.0 + np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * rng.uniform(0.5, 1.5, size=n)
    return np.ascontiguousarray(x, dtype=np.float64)


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libljmd_ref.so"))
    lib.ljref_series_statistics.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    os.makedirs(OUT, exist_ok=True)
    xs, outs = [], []
    for seed in range(24):
        x = series(seed)
        out = np.zeros(12)
        lib.ljref_series_statistics(x.ctypes.data, len(x), out.ctypes.data)
        xs.append(x)
        outs.append(out)
    np.savez_compressed(os.path.join(OUT, "series_statistics.npz"), n=np.array([len(x) for x in xs]),
                        x=np.concatenate(xs), out=np.array(outs))
    print("wrote", len(xs), "series;", sum(len(x) for x in xs), "values")


if __name__ == "__main__":
    main()
