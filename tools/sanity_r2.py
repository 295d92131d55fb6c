"""compute-sanitizer workload for the round-2 code paths: record sorting + warp frames (periodic Newton-3 runs), the
RDF build of the FRAMES kernel, super-tiles with ragged tails, the observation trace with graph replay, the
single-process multi-rank handle on one device (LJMD_SHARE_DEVICES=1), device-side initial conditions, P_xy."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ljpkg import load  # noqa: E402

pkg = load()
for N, rho, bc, canonical in ((5003, 0.8, 0, True), (9000, 0.3, 0, False), (6000, 0.05, 1, False)):
    pos, vel = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=5), pkg.snapshots.velocities(N, 1.0, seed=5)
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=bc) as s:
        s.set_state(pos, vel)
        s.trace_begin([(0, 0.05), (3, 0.05), (6, 0.05, 3.0)], 64)
        s.step(0.004, 40, rdf_every=3)
        tr = s.trace_read()
        s.trace_end()
        s.step(0.004, 140, rdf_every=5)          # graph replay, a re-sort inside the batch, RDF launches
        print(N, bc, tr["counts"].shape, int(s.rdf_counts().sum()), s.scalars()["T"], s.pshear())
if os.environ.get("LJMD_SHARE_DEVICES") == "1":
    N, rho = 9000, 0.5
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=True, bc=0, devices=[0, 0, 0]) as s:
        s.init_state(5)
        s.step(0.004, 20, rdf_every=4)
        p, v, f = s.get_state()
        print("shared x3", int(s.rdf_counts().sum()), s.scalars()["T"], float(np.abs(f[:, :3]).max()))
