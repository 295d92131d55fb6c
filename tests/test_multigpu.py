"""GPU, 2 ranks (one process per GPU, NCCL): the i-sharded step against the single-GPU step on the same
snapshot — forces equal up to the order of the j-split partial sums, RDF and speed histograms bit-identical,
positions after 5 steps equal to float rounding.  Skipped with fewer than 2 devices."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("name,canonical,bc,N,rho", [("tvn_periodic", 1, 0, 6000, 0.8), ("evn_hardwall", 0, 1, 5003, 0.05),
                                                     ("tvn_periodic_sym", 1, 0, 20000, 0.5)])
def test_two_gpu_step_matches_single_gpu(pkg, gpu_lib, tmp_path, name, canonical, bc, N, rho, comm):
    """comm = p2p: per-step exchange over the CUDA-IPC peer windows (fabric); nccl: the NCCL transport."""
    if gpu_lib.ljmd_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", LJMD_COMM=comm)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29544", os.path.join(ROOT, "tests", "dist_worker.py"), "gpu", str(tmp_path),
           name, str(canonical), str(bc), str(N), str(rho)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    meta = json.load(open(os.path.join(tmp_path, "gpu_rank0.json")))
    assert meta["fabric"] == (comm == "p2p")
    r0 = np.load(os.path.join(tmp_path, f"{name}_rank0.npz"))
    r1 = np.load(os.path.join(tmp_path, f"{name}_rank1.npz"))
    for k in r0.files:                       # every rank reports the same global state
        assert np.array_equal(r0[k], r1[k]), k
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=21)
    vel = pkg.snapshots.velocities(N, 1.0, seed=21)
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=bc) as s:
        s.set_state(pos, vel)
        _, _, f0 = s.get_state()
        sc0 = s.scalars()
        rdf0 = s.rdf_counts()
        s.step(0.004, 5, rdf_every=5)
        p1, v1, f1 = s.get_state()
        sc1 = s.scalars()
        rdf1, _ = s.rdf_accum()
        vh = s.velocity_histogram(0.12, 101)
        sub = s.subvolume_counts(3, 0.05)
        s.trace_begin([(0, 0.05), (3, 0.1), (5, 0.05, 3.0)], 8)
        s.step(0.004, 3)
        tr = s.trace_read()
        s.trace_end()
    # sharded counters and trace rows: the same occupancies up to a particle within 1e-6 of a sub-volume face
    assert r0["sub"].shape == sub.shape and np.abs(r0["sub"] - sub).max() <= 2
    assert r0["tr_counts"].shape == tr["counts"].shape == (3, 19 + 9 + 20)
    assert np.abs(r0["tr_counts"] - tr["counts"]).max() <= 2
    assert (r0["tr_counts"] <= N).all() and (r0["tr_counts"][:, -1] > 0.99 * N).all()      # |vy| < 3 sigma: 99.7 %
    assert np.allclose(r0["tr_scal"], tr["scalars"], rtol=1e-5, atol=1e-7)
    assert np.allclose(r0["tr_mv"], tr["mean_velocity"], rtol=0, atol=1e-6)
    fscale = np.abs(f0[:, :3]).max()
    assert np.abs(r0["f0"][:, :3] - f0[:, :3]).max() <= 2e-6 * fscale
    assert np.array_equal(r0["rdf0"], rdf0)
    assert np.abs(r0["p1"][:, :3] - p1[:, :3]).max() <= 1e-6 * max(1.0, s.L)
    assert np.abs(r0["v1"][:, :3] - v1[:, :3]).max() <= 1e-5 * np.abs(v1[:, :3]).max()
    assert np.abs(r0["rdf1"] - rdf1).sum() <= max(4, 1e-5 * rdf1.sum())   # a pair may cross a bin edge after 5 steps
    assert np.abs(r0["vh"] - vh).sum() <= 2
    a0 = np.array([sc0[k] for k in sorted(sc0)])
    a1 = np.array([sc1[k] for k in sorted(sc1)])
    assert np.allclose(r0["sc0"], a0, rtol=1e-6, atol=1e-9)
    assert np.allclose(r0["sc1"], a1, rtol=1e-5, atol=1e-7)
