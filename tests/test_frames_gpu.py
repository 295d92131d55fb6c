"""GPU tests of the record order and the warp frames (periodic Newton-3 runs; ljmd_sort.cuh, ljmd_force_sym.cuh):
the sorted / float-frame path against the FP64 arbiter and against the same library with LJMD_FRAMES=0 (identity
record order, fixed-point minimum image everywhere), on particle orders and geometries chosen to hurt:
shuffled input, small boxes where half a box is a few sigma, boundary switches, re-sorts inside a batch.

Tolerances are the ones of test_parity_gpu.py (north star: <= 1e-5 relative in FP32, RDF bit-exact)."""
import numpy as np
import pytest

from test_parity_gpu import FORCE_TOL, SCALAR_TOL, subsample_force_error

pytestmark = pytest.mark.gpu


def evaluate(pkg, N, rho, pos, vel, bc=0, canonical=0, T0=1.0):
    with pkg.ljmd.LJSystem(N, T0=T0, rho=rho, canonical=canonical, bc=bc) as s:
        s.set_state(pos, vel)
        p, v, f = s.get_state()
        return p, v, f, s.scalars(), s.rdf_counts(), s.L


@pytest.mark.parametrize("N,rho", [(16384, 0.85), (32768, 0.3), (65536, 1.1)])
def test_frames_and_fixed_point_agree(pkg, oracle, gpu_lib, monkeypatch, N, rho):
    """Shuffled particle order in, the caller's order out: positions untouched, forces of the sorted float-frame
    path and of the unsorted fixed-point path both within 1e-5 of the FP64 arbiter and within 2e-6 of each other,
    V and virial alike, RDF bins identical."""
    pos = pkg.snapshots.lattice(N, rho, jitter=0.08, seed=5)
    vel = pkg.snapshots.velocities(N, 1.0, seed=5)
    perm = np.random.default_rng(9).permutation(N)
    pos, vel = np.ascontiguousarray(pos[perm]), np.ascontiguousarray(vel[perm])
    monkeypatch.setenv("LJMD_KERNEL", "sym")
    monkeypatch.delenv("LJMD_FRAMES", raising=False)
    p1, v1, f1, sc1, rdf1, L = evaluate(pkg, N, rho, pos, vel)
    monkeypatch.setenv("LJMD_FRAMES", "0")
    p0, v0, f0, sc0, rdf0, _ = evaluate(pkg, N, rho, pos, vel)
    assert np.array_equal(p1, pos) and np.array_equal(v1, vel), "the state arrays keep the caller's particle order"
    for name, f in (("frames", f1), ("fixed-point", f0)):
        err_f, _, n = subsample_force_error(oracle, pos, L, 0, f)
        assert err_f <= FORCE_TOL, f"{name}: force error {err_f:.3e} on {n} particles"
    idx = np.arange(0, N, 7)
    _, fterm, _, _ = oracle.forces_f64_subset(pos, L, 0, idx.astype(np.int32))
    diff = np.abs(f1[idx, :3].astype(np.float64) - f0[idx, :3].astype(np.float64)).max(axis=1) / fterm
    assert diff.max() <= 2e-6, f"frames vs fixed-point: {diff.max():.3e}"
    vabs = 2.0 * np.abs(f0[:, 3]).astype(np.float64).sum() * 4.0 / 2.0
    assert abs(sc1["V"] - sc0["V"]) <= 3e-6 * vabs   # two float summation orders of 65 536-term per-particle sums
    assert abs(sc1["Pvirial"] - sc0["Pvirial"]) <= 3e-6 * max(abs(sc0["Pvirial"]), vabs)
    assert np.array_equal(rdf1, rdf0), "RDF bins must not depend on the record order"


@pytest.mark.parametrize("N,rho", [(4096, 1.2), (8192, 0.59)])
def test_small_box_where_half_a_box_is_close(pkg, oracle, gpu_lib, monkeypatch, N, rho):
    """Small boxes (L/2 = 7.5 and 12 sigma): pairs whose separation sits at half a box on some axis still carry a
    visible share of the potential, and every (warp, chunk) combination is close to the wrap.  A frame that took one
    translation where the minimum image needs two would move V and the virial: full FP64 arbiter, half the usual
    tolerance on the scalars."""
    pos = pkg.snapshots.lattice(N, rho, jitter=0.06, seed=13)
    vel = pkg.snapshots.velocities(N, 1.0, seed=13)
    perm = np.random.default_rng(5).permutation(N)
    pos = np.ascontiguousarray(pos[perm])
    monkeypatch.setenv("LJMD_KERNEL", "sym")
    monkeypatch.delenv("LJMD_FRAMES", raising=False)
    p, v, frc, sc, rdf, L = evaluate(pkg, N, rho, pos, vel)
    f64, fabs_sum, sc64 = oracle.forces_f64(pos, L, 0)
    err = np.abs(frc[:, :3].astype(np.float64) - f64).max(axis=1) / sc64["fterm_sum"]
    assert err.max() <= FORCE_TOL, f"force error {err.max():.3e}"
    assert abs(sc["V"] - sc64["V"]) <= 0.5 * SCALAR_TOL * sc64["Vabs"]
    assert abs(sc["Pvirial"] - sc64["Pvirial"]) <= 0.5 * SCALAR_TOL * sc64["Pabs"]
    _, _, rdf_ref = oracle.forces(pos, L, 0, oracle.rdf_dr2(N))
    assert np.array_equal(rdf, rdf_ref)


@pytest.mark.parametrize("canonical", [0, 1])
def test_resorts_inside_a_batch_keep_batched_and_single_steps_identical(pkg, gpu_lib, monkeypatch, canonical):
    """LJMD_SORT_INTERVAL=8: a 40-step batch is cut into segments at every re-sort; the same 40 steps one call at a
    time sort at the same steps (handle-wide count).  Bit-identical state and scalars; RDF cadence preserved."""
    N, rho = 8192, 0.6
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=8)
    vel = pkg.snapshots.velocities(N, 1.0, seed=8)
    monkeypatch.setenv("LJMD_KERNEL", "sym")
    monkeypatch.setenv("LJMD_SORT_INTERVAL", "8")
    out = []
    for single in (False, True):
        with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=0) as s:
            s.set_state(pos, vel)
            if single:
                for k in range(40):
                    s.step(0.004, 1, rdf_every=(1 if (k + 1) % 5 == 0 else 0))
            else:
                s.step(0.004, 40, rdf_every=5)
            p, v, f = s.get_state()
            out.append((p, v, f, s.scalars(), s.rdf_accum()[0] if hasattr(s, "rdf_accum") else None))
    (p0, v0, f0, sc0, r0), (p1, v1, f1, sc1, r1) = out
    assert np.array_equal(p0, p1) and np.array_equal(v0, v1) and np.array_equal(f0, f1)
    for k in ("U", "T", "P", "K", "V"):
        assert sc0[k] == sc1[k], k
    if r0 is not None:
        assert np.array_equal(r0, r1)


def test_sorted_run_tracks_the_unsorted_run(pkg, oracle, gpu_lib, monkeypatch):
    """60 EVN steps at liquid density with a re-sort every 16: energy conserved like the fixed-point run, positions
    still within rounding-driven divergence of it, forces of the final state within tolerance of the arbiter."""
    N, rho = 16384, 0.85
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=4)
    vel = pkg.snapshots.velocities(N, 1.0, seed=4)
    monkeypatch.setenv("LJMD_KERNEL", "sym")
    res = {}
    for frames in ("1", "0"):
        monkeypatch.setenv("LJMD_FRAMES", frames)
        monkeypatch.setenv("LJMD_SORT_INTERVAL", "16")
        with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=0, bc=0) as s:
            s.set_state(pos, vel)
            u0 = s.scalars()["U"]
            s.step(0.004, 60)
            p, v, f = s.get_state()
            res[frames] = (p, f, s.scalars()["U"], u0, s.L)
    p1, f1, u1, u01, L = res["1"]
    p0, f0, u0, u00, _ = res["0"]
    assert abs(u01 - u00) <= 1e-6 * abs(u00) + 1e-3
    assert abs((u1 - u01) - (u0 - u00)) / N <= 2e-5, "energy drift must not depend on the record order"
    d = np.abs(p1[:, :3] - p0[:, :3])
    d = np.minimum(d, L - d)
    assert d.max() <= 5e-3, f"trajectories diverged by {d.max():.2e} after 60 steps"
    err_f, _, n = subsample_force_error(oracle, p1, L, 0, f1, nsample=256, seed=1, interior=3.0)
    assert err_f <= 2 * FORCE_TOL


def test_boundary_switch_and_back_keeps_a_valid_order(pkg, oracle, gpu_lib, monkeypatch):
    """periodic -> hard wall -> periodic with steps in between: the record order is a permutation at every point
    (forces stay right), and the periodic box is re-sorted when it comes back."""
    N, rho = 8192, 0.5
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=2)
    vel = pkg.snapshots.velocities(N, 1.0, seed=2)
    monkeypatch.setenv("LJMD_KERNEL", "sym")
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=0, bc=0) as s:
        s.set_state(pos, vel)
        s.step(0.004, 5)
        s.set_boundary(1)
        s.step(0.004, 5)
        p, _, _ = s.get_state()
        s.compute_forces()
        _, _, f = s.get_state()
        err_f, _, _ = subsample_force_error(oracle, p, s.L, 1, f, nsample=200, seed=3)
        assert err_f <= FORCE_TOL
        s.set_boundary(0)
        s.step(0.004, 5)
        p, _, _ = s.get_state()
        s.compute_forces()
        _, _, f = s.get_state()
        err_f, _, _ = subsample_force_error(oracle, p, s.L, 0, f, nsample=200, seed=3)
        assert err_f <= FORCE_TOL
