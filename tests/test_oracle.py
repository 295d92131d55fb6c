"""CPU: the C restatement (oracle/ljmd_oracle.c) against (a) the golden fixtures generated from the
unmodified reference CPU path and (b) that reference library itself when it is present."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle.oracle import Reference, reference_available


@pytest.mark.parametrize("name", golden_names())
def test_forces_match_golden_bitwise(oracle, name):
    g = load_golden(name)
    L = oracle.box_length(g["N"], g["rho"])
    assert L == g["s0"]["L"]
    assert oracle.rdf_dr2(g["N"]) == g["dr2"]
    frc, sc, rdf = oracle.forces(g["pos0"], L, g["bc"], g["dr2"])
    assert np.array_equal(frc[:, :3], g["force0"][:, :3])
    assert np.array_equal(rdf, g["rdf0"])
    par = oracle.parameters(g["N"], g["rho"], g["vel0"], sc["V"], sc["Pvirial"], sc["Pshear_conf"])
    for k in ["U", "T", "K", "V", "P", "Pshear"]:
        assert par[k] == g["s0"][k], k


@pytest.mark.parametrize("name", golden_names())
def test_steps_match_golden_bitwise(oracle, name):
    g = load_golden(name)
    pos, vel, frc, sc, rdf = oracle.integrate(g["N"], g["rho"], g["T0"], g["canonical"], g["bc"], g["dt"], g["pos0"],
                                              g["vel0"], g["force0"], nsteps=g["steps"], dr2=g["dr2"])
    assert np.array_equal(pos[:, :3], g["pos1"][:, :3])
    assert np.array_equal(vel[:, :3], g["vel1"][:, :3])
    assert np.array_equal(frc[:, :3], g["force1"][:, :3])
    assert np.array_equal(rdf, g["rdf1"])
    for k in ["U", "T", "K", "V", "P", "Pshear"]:
        assert sc[k] == g["s1"][k], k


@pytest.mark.parametrize("name", golden_names())
def test_velocity_histogram_matches_golden(oracle, name):
    g = load_golden(name)
    assert np.array_equal(oracle.velocity_histogram(g["vel0"], 0.12, 101), g["velhist0"])


def test_rdf_curve_and_lattice(oracle, pkg):
    g = load_golden("liquid_evn_periodic")
    L = oracle.box_length(g["N"], g["rho"])
    r, gr = oracle.rdf_curve(g["N"], L, g["dr2"], g["rdf0"])
    r2, gr2 = pkg.ljmd.rdf_curve(g["N"], L, g["dr2"], g["rdf0"])
    assert np.array_equal(r, r2) and np.array_equal(gr, gr2)
    # liquid structure: first peak of g(r) near r ~ 1.1 and g -> 1 at large r
    assert 1.0 < r[np.argmax(gr)] < 1.25 and gr.max() > 2.0
    for N, rho in [(400, 0.05), (500, 0.85), (131, 0.2), (4096, 1.1)]:
        assert np.array_equal(oracle.lattice(N, oracle.box_length(N, rho)), pkg.snapshots.lattice(N, rho))


def test_fp64_arbiter_agrees_with_restatement(oracle):
    g = load_golden("liquid_evn_periodic")
    L = oracle.box_length(g["N"], g["rho"])
    f64, fabs_sum, sc64 = oracle.forces_f64(g["pos0"], L, g["bc"])
    frc, sc, _ = oracle.forces(g["pos0"], L, g["bc"], g["dr2"])
    err = np.abs(frc[:, :3].astype(np.float64) - f64).max(axis=1) / fabs_sum
    assert err.max() < 5e-6          # the reference's own float rounding on this metric (dense liquid)
    assert abs(sc["V"] - sc64["V"]) < 1e-6 * sc64["Vabs"]
    assert abs(sc["Pvirial"] - sc64["Pvirial"]) < 1e-6 * sc64["Pabs"]


def test_numpy_rdf_restatement_equals_c_oracle(oracle, pkg):
    from oracle.rdf_numpy import rdf_counts
    for N, rho, bc, jit in [(500, 0.85, 0, 0.3), (400, 0.05, 0, 0.4), (600, 1.1, 0, 0.1), (400, 0.3, 1, 0.3)]:
        pos = pkg.snapshots.lattice(N, rho, jitter=jit, seed=3)
        L, dr2 = oracle.box_length(N, rho), oracle.rdf_dr2(N)
        pos[:, :3] += np.float32(0.3 * L) * (np.arange(N) % 3 == 0)[:, None]   # some outside the box
        _, _, rdf = oracle.forces(pos, L, bc, dr2)
        assert np.array_equal(rdf, rdf_counts(pos, L, bc, dr2))


@pytest.mark.skipif(not reference_available(), reason="oracle/_ref/libljmd_ref.so not built")
@pytest.mark.parametrize("N,T,rho,canonical,bc,seed", [
    (257, 1.3, 0.4, 1, 0, 1), (300, 0.8, 0.9, 0, 0, 2), (222, 1.0, 0.02, 0, 1, 3), (200, 2.0, 0.5, 1, 2, 4),
    (64, 1.0, 0.7, 1, 0, 5),
])
def test_restatement_equals_reference_live(oracle, pkg, N, T, rho, canonical, bc, seed):
    # dense lattices only tolerate a small jitter before particles overlap and the energy explodes
    pos = pkg.snapshots.lattice(N, rho, jitter=0.2 if rho < 0.1 else 0.05, seed=seed)
    vel = pkg.snapshots.velocities(N, T, seed=seed)
    ref = Reference(N, T, rho, canonical, bc)
    ref.set_state(pos, vel)
    _, _, rf = ref.get_state()
    L, dr2 = oracle.box_length(N, rho), oracle.rdf_dr2(N)
    frc, sc, rdf = oracle.forces(pos, L, bc, dr2)
    rrdf, rdr2 = ref.rdf_counts()
    assert rdr2 == dr2 and np.array_equal(rdf, rrdf)
    assert np.array_equal(frc[:, :3], rf[:, :3])
    ref.integrate(0.004, 7)
    p2, v2, f2, s2, rdf2 = oracle.integrate(N, rho, T, canonical, bc, 0.004, pos, vel, frc, nsteps=7)
    rp, rv, rf = ref.get_state()
    rs = ref.scalars()
    assert np.array_equal(p2[:, :3], rp[:, :3]) and np.array_equal(v2[:, :3], rv[:, :3])
    assert np.array_equal(f2[:, :3], rf[:, :3])
    assert np.array_equal(rdf2, ref.rdf_counts()[0])
    for k in ["U", "T", "K", "V", "P", "Pshear"]:
        assert s2[k] == rs[k], k
    rx, rg = ref.rdf_curve()
    ox, og = oracle.rdf_curve(N, L, dr2, rdf2)
    assert np.array_equal(rx, ox) and np.array_equal(rg, og)
    vx, vd = ref.velocity_histogram(12.0, 0.12)
    assert np.array_equal(np.rint(vd * 0.12 * N).astype(np.int32), oracle.velocity_histogram(v2, 0.12, len(vx)))


def test_subset_arbiter_equals_full_arbiter(oracle):
    """ljo_forces_f64_subset (the full-size checker: O(nsub * N), threaded) is the FP64 arbiter restricted to
    the sampled particles, bit for bit, for periodic and open boxes and any thread count."""
    import ljpkg
    pkg = ljpkg.load()
    N, rho = 1500, 0.6
    pos = pkg.snapshots.lattice(N, rho, jitter=0.08, seed=11)
    L = oracle.box_length(N, rho)
    idx = np.array([0, 1, 17, 511, 512, 733, N - 1], dtype=np.int32)
    for bc in (0, 1):
        f64, _, sc = oracle.forces_f64(pos, L, bc)
        for threads in (1, 3, 16):
            f, fterm, pe, peabs = oracle.forces_f64_subset(pos, L, bc, idx, threads=threads)
            assert np.array_equal(f, f64[idx])
            assert np.array_equal(fterm, sc["fterm_sum"][idx])
            assert (peabs >= np.abs(pe)).all()
        # the per-particle potentials add up to V = 4/2 * sum_i pe_i (MDSystem.cpp:297,308)
        _, _, pe_all, _ = oracle.forces_f64_subset(pos, L, bc, np.arange(N, dtype=np.int32))
        assert abs(2.0 * pe_all.sum() - sc["V"]) <= 1e-12 * sc["Vabs"]
