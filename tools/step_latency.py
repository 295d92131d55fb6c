"""Step latency of small systems with and without the CUDA-graph replay (gpurun -- python tools/step_latency.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ljpkg import load  # noqa: E402

pkg = load()
for N, rho in ((400, 0.05), (1500, 0.3), (4096, 0.3), (16384, 0.85)):
    pos, vel = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=3), pkg.snapshots.velocities(N, 1.0, seed=3)
    for canonical in (True, False):
        out = []
        for g in ("1", "0"):
            os.environ["LJMD_GRAPH"] = g
            with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=0) as s:
                s.set_state(pos, vel)
                s.step(0.004, 200)
                n = 4000 if N <= 4096 else 1000
                t0 = time.perf_counter()
                s.step(0.004, n)
                out.append((time.perf_counter() - t0) / n * 1e6)
        print(f"N={N:6d} {'TVN' if canonical else 'EVN'}: {out[0]:7.2f} us/step with graph replay, {out[1]:7.2f} plain launches")
