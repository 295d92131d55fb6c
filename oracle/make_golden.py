"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the UNMODIFIED reference CPU path.

Run in the build container (needs /root/reference, via oracle/_ref/libljmd_ref.so):
    python oracle/make_golden.py
The reference holds no golden vectors of its own; these fixtures are outputs of the
reference itself on seeded snapshots, committed so that CPU tests and the GPU box (which has no
/root/reference) can check both the restatement and the CUDA path against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ljpkg  # noqa: E402
from oracle.oracle import Reference  # noqa: E402

snap = ljpkg.load().snapshots
OUT = os.path.join(ROOT, "tests", "golden")

# name: N, T, rho, canonical, bc, equilibration steps (TVN, from the jittered lattice), dt, steps recorded
CASES = {
    "c1_gas_tvn_periodic":     dict(N=400, T=1.4, rho=0.05, canonical=1, bc=0, eq=300, dt=0.004, steps=4),
    "liquid_evn_periodic":     dict(N=500, T=1.0, rho=0.85, canonical=0, bc=0, eq=300, dt=0.004, steps=4),
    "solid_tvn_periodic":      dict(N=432, T=1.0, rho=1.1, canonical=1, bc=0, eq=100, dt=0.004, steps=4),
    "gas_evn_hardwall":        dict(N=400, T=1.0, rho=0.01, canonical=0, bc=1, eq=0, dt=0.004, steps=4),
    "mixed_tvn_periodic":      dict(N=600, T=1.0, rho=0.3, canonical=1, bc=0, eq=300, dt=0.004, steps=4),
    "expansion_evn_none":      dict(N=343, T=1.5, rho=0.6, canonical=0, bc=2, eq=0, dt=0.004, steps=4),
    "ragged_tvn_hardwall":     dict(N=131, T=2.0, rho=0.2, canonical=1, bc=1, eq=50, dt=0.005, steps=3),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, c in CASES.items():
        N = c["N"]
        pos, vel = snap.lattice(N, c["rho"], jitter=0.05, seed=7), snap.velocities(N, c["T"], seed=7)
        r = Reference(N, c["T"], c["rho"], 1, c["bc"])
        r.set_state(pos, vel)
        if c["eq"]:
            r.integrate(c["dt"], c["eq"])           # equilibrate in TVN so g(r) is not a lattice comb
        pos0, vel0, _ = r.get_state()
        r.set_canonical(c["canonical"])
        r.set_state(pos0, vel0)                      # the recorded snapshot: forces/parameters on it
        _, _, f0 = r.get_state()
        s0 = r.scalars()
        rdf0, dr2 = r.rdf_counts()
        vx, vd = r.velocity_histogram(12.0, 0.12)
        r.integrate(c["dt"], c["steps"])
        pos1, vel1, f1 = r.get_state()
        s1 = r.scalars()
        rdf1, _ = r.rdf_counts()
        keys = ["U", "T", "K", "V", "P", "Pshear", "t", "L", "av_U_tot", "av_T_tot", "av_p_tot", "av_iters"]
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            cfg=np.array([N, c["T"], c["rho"], c["canonical"], c["bc"], c["dt"], c["steps"]], dtype=np.float64),
            dr2=np.float32(dr2), pos0=pos0, vel0=vel0, force0=f0, rdf0=rdf0,
            scal0=np.array([s0[k] for k in keys]), velhist0=np.rint(vd * 0.12 * N).astype(np.int32),
            pos1=pos1, vel1=vel1, force1=f1, rdf1=rdf1, scal1=np.array([s1[k] for k in keys]))
        print(f"{name}: N={N} U/N={s0['U'] / N:.4f} T={s0['T']:.4f} P={s0['P']:.4f} rdf_sum={rdf0.sum()} "
              f"-> after {c['steps']} steps U/N={s1['U'] / N:.4f} T={s1['T']:.4f}")
        r.close()


if __name__ == "__main__":
    main()
