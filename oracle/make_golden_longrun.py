#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — generates tests/golden/stats/c1_longrun_reference.json.

BASELINE config C1 (N = 400, T* = 1.4, rho* = 0.05, periodic, TVN, dt* = 0.004; input/N400.Tst1.4.rhost0.05 of the
reference) run through the UNMODIFIED reference CPU path (oracle/_ref/libljmd_ref.so): per replica a seeded start
snapshot, teq = 50 time units of equilibration, then `NPROD` production steps.  Stored per replica: the means of
u* = U/N, T*, Z = P/(rho* T*) over the production phase and their errors by the reference's own error model
(TimeAverage::GetMeanError, src/tasks/auxiliary/time-average-aux.h:38-66).  The GPU long-run test compares against
these.  Usage: python oracle/make_golden_longrun.py   (about 5 minutes: the replicas run in parallel processes)."""
import json
import math
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, T0, RHO, DT = 400, 1.4, 0.05, 0.004
NEQ, NPROD = 12500, 40000
SEEDS = [101, 102, 103, 104]


def time_average(x):
    """(mean, error, s) as TimeAverage::GetMean / GetMeanError / GetS (time-average-aux.h:38-66)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    mean = x.mean()
    var = x.var()
    c1 = (x[:-1] * x[1:]).sum() / (n - 1) - mean * mean
    s = 2.0 / math.log(var / c1) if (c1 > 0 and var > c1) else 1.0
    if s < 0:
        s = 1.0
    return mean, math.sqrt(var / n) * math.sqrt(s), s


def replica(seed):
    import ljpkg
    from oracle.oracle import Reference
    pkg = ljpkg.load()
    pos = pkg.snapshots.lattice(N, RHO, jitter=0.05, seed=seed)
    vel = pkg.snapshots.velocities(N, T0, seed=seed)
    ref = Reference(N, T0, RHO, 1, 0)
    ref.set_state(pos, vel)
    ref.integrate(DT, NEQ)
    u, T, Z = np.empty(NPROD), np.empty(NPROD), np.empty(NPROD)
    for k in range(NPROD):
        ref.integrate(DT, 1)
        sc = ref.scalars()
        u[k], T[k], Z[k] = sc["U"] / N, sc["T"], sc["P"] / (RHO * sc["T"])
    out = {"seed": seed}
    for name, x in (("u", u), ("T", T), ("Z", Z)):
        m, e, s = time_average(x)
        out[name] = {"mean": m, "error": e, "inefficiency": s, "std": float(x.std())}
    return out


def main():
    with mp.Pool(len(SEEDS)) as pool:
        reps = pool.map(replica, SEEDS)
    doc = {"config": {"N": N, "T0": T0, "rho": RHO, "dt": DT, "canonical": 1, "bc": 0, "neq": NEQ, "nprod": NPROD},
           "start": "pkg.snapshots.lattice(N, rho, jitter=0.05, seed) + pkg.snapshots.velocities(N, T0, seed)",
           "source": "unmodified reference CPU path, oracle/_ref/libljmd_ref.so (-O2 -ffp-contract=off)",
           "replicas": reps, "combined": {}}
    for name in ("u", "T", "Z"):
        means = np.array([r[name]["mean"] for r in reps])
        doc["combined"][name] = {"mean": float(means.mean()),
                                 "error_replicas": float(means.std(ddof=1) / math.sqrt(len(means))),
                                 "error_model": float(math.sqrt(sum(r[name]["error"] ** 2 for r in reps)) / len(reps))}
    path = os.path.join(ROOT, "tests", "golden", "stats", "c1_longrun_reference.json")
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=1)
    print(json.dumps(doc["combined"], indent=1))


if __name__ == "__main__":
    main()
