"""One-off check beyond the tested sizes: one force evaluation at N = 4 194 304 and N = 8 388 608 (periodic,
rho* = 0.3) and the FP64 subsample arbiter on 128 particles; prints the kernel family the library chose, the time
of the evaluation and the worst force error.  usage: python tests/large_n_check.py [N ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ljpkg import load  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402  (test infrastructure: the arbiter is the checker here)

pkg, o = load(), Oracle()
for N in [int(a) for a in sys.argv[1:]] or [4194304, 8388608]:
    rho = 0.3
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=3)
    vel = pkg.snapshots.velocities(N, 1.0, seed=3)
    try:
        with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=True, bc=0) as s:
            info = s.launch_info()
            t0 = time.perf_counter()
            s.set_state(pos, vel)
            dt = time.perf_counter() - t0
            _, _, frc = s.get_state()
            idx = np.sort(np.random.default_rng(1).choice(N, 128, replace=False)).astype(np.int32)
            f64, fterm, _, _ = o.forces_f64_subset(pos, s.L, 0, idx)
            err = (np.abs(frc[idx, :3].astype(np.float64) - f64).max(axis=1) / fterm).max()
            f = frc[:, :3].astype(np.float64)
            print(f"N={N}: newton3={info['newton3']} ctas={info['force_ctas']} set_state {dt:.2f} s "
                  f"({N * (N - 1.0) / dt:.3e} pairs/s incl. upload) force err {err:.2e} "
                  f"|sum f|/sum|f| {np.abs(f.sum(axis=0)).max() / np.abs(f).sum():.1e}", flush=True)
    except Exception as e:   # report, do not hide: which size stops working and why
        print(f"N={N}: FAILED {type(e).__name__}: {e}", flush=True)
