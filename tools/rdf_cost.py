"""Force-kernel time of plain steps and of RDF steps, on one GPU (gpurun -- python tools/rdf_cost.py [config])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ljpkg import load
pkg = load()
ljmd, snapshots = pkg.ljmd, pkg.snapshots

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    cfg = snapshots.CONFIGS[name]
    pos, vel = snapshots.make(name)
    N = pos.shape[0]
    with ljmd.LJSystem(N, T0=cfg["T"], rho=cfg["rho"], canonical=cfg["canonical"], bc=cfg["bc"]) as s:
        s.set_state(pos, vel)
        s.step(0.004, 20, 0)
        s.set_event_timing(True)
        for label, every in (("plain", 0), ("rdf every step", 1)):
            for rep in range(2):
                s.step(0.004, 10, every)
                t = s.last_step_timing()
            print(f"{name} N={N} {label:16s}: force kernel {t['force_ms'] / max(1, t['force_launches']):.4f} ms/launch, "
                  f"step {t['total_ms'] / 10:.4f} ms, rdf pairs {int(np.sum(s.rdf_counts())) if every else 0}")
        # scrambled particle order: block bounding boxes lose their meaning, pruning cannot help
        perm = np.random.default_rng(1).permutation(N)
        p, v, _ = s.get_state()
        s.set_state(np.ascontiguousarray(p[perm]), np.ascontiguousarray(v[perm]))
        for rep in range(2):
            s.step(0.004, 10, 1)
            t = s.last_step_timing()
        print(f"{name} N={N} rdf, scrambled order: force kernel {t['force_ms'] / max(1, t['force_launches']):.4f} ms/launch")

if __name__ == "__main__":
    main()
