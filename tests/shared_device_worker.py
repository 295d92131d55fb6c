"""Worker of tests/test_multigpu.py::test_ranks_sharing_one_device_match_single_gpu: `world` ranks of the
single-process multi-device handle on ONE GPU (LJMD_SHARE_DEVICES=1), the observation sequence of the multi-GPU
tests, results to an .npz.  Run in its own process: the environment must be set before CUDA starts
(CUDA_DEVICE_MAX_CONNECTIONS: every rank's stream needs its own hardware queue, or a rank's spinning barrier could
sit in front of the kernel its peer has to run first)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out, world, canonical, bc, N, rho = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
    import ljpkg
    from test_multigpu import observe
    pkg = ljpkg.load()
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=21)
    vel = pkg.snapshots.velocities(N, 1.0, seed=21)
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=bc, devices=[0] * world) as s:
        info = s.launch_info()
        assert info["world"] == world, info
        got = observe(s, pos, vel)
        p, v, _ = s.get_state()
        hp, hv, hf = p.copy(), v.copy(), np.zeros_like(p)
        s.integrate_host(0.004, hp, hv, hf)
        p2, v2, f2 = s.get_state()
        assert np.array_equal(hp, p2) and np.array_equal(hv, v2) and np.array_equal(hf, f2)
        # seeded initial conditions are counted by the global particle index: any rank count gives the same state
        s.init_state(77)
        ip, iv, _ = s.get_state()
    np.savez(out, init_pos=ip, init_vel=iv, newton3=np.array(int(info["newton3"])), **got)


if __name__ == "__main__":
    main()
