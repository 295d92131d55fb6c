"""GPU tests of the product's host class (lennard-jones-cuda_b200/host/MDSystem.{h,cpp} -> libljmd_host.so) next
to the reference class, both driven through the same extern "C" shim (oracle/ref_shim.cpp compiled once against
the reference header, once against the product's source-compatible header): SURVEY.md §8 rows a-10
(KineticTemperature / Renormalize* / CorrectTotalMomentum), a-11 (SampleInitialConditions) and the velocity
histogram's running mean (a-9)."""
import os

import numpy as np
import pytest

from oracle.oracle import HOST_SHIM_LIB, Reference, reference_available

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(HOST_SHIM_LIB) and reference_available()),
                                 reason="host shim / reference library not built")]

CASES = [(400, 1.4, 0.05, 1, 0), (1000, 0.9, 0.7, 0, 0), (729, 2.0, 0.3, 0, 1)]


def both(N, T0, rho, canonical, bc):
    return Reference(N, T0, rho, canonical, bc), Reference(N, T0, rho, canonical, bc, legacy="product")


@pytest.mark.parametrize("N,T0,rho,canonical,bc", CASES)
def test_sample_initial_conditions(pkg, gpu_lib, monkeypatch, N, T0, rho, canonical, bc):
    """MDSystem.cpp:147-181: the lattice is bit-equal to the reference's; the velocities (the reference's generator
    is time-seeded, ours is a seeded mt19937_64) have zero total momentum, kinetic temperature T0 and Maxwell speeds."""
    from scipy import stats
    monkeypatch.setenv("LJMD_SEED", "31337")
    ref, our = both(N, T0, rho, canonical, bc)
    pr, vr, _ = ref.get_state()
    po, vo, _ = our.get_state()
    assert np.array_equal(pr, po), "start lattice differs from MDSystem.cpp:150-168"
    assert np.all(po[:, 3] == np.float32(ref.scalars()["L"] / np.float32(150.0)))
    for v in (vr, vo):
        v3 = v[:, :3].astype(np.float64)
        assert np.abs(v3.sum(axis=0)).max() <= 2e-5 * np.sqrt(N) * np.sqrt(T0)            # CorrectTotalMomentum, float rounding
        assert abs((v3 * v3).sum() / (3 * N) - T0) <= 2e-6 * T0                               # RenormalizeVelocities(true)
        speed = np.sqrt((v3 * v3).sum(axis=1))
        ks = stats.kstest(speed, stats.maxwell(scale=np.sqrt(T0)).cdf)
        assert ks.pvalue > 1e-3, ks                                                           # Maxwell(v) at T0, :36-53
        # isotropy: every component is N(0, T0)
        for a in range(3):
            assert stats.kstest(v3[:, a], stats.norm(scale=np.sqrt(T0)).cdf).pvalue > 1e-3
    assert abs(our.kinetic_temperature() - T0) <= 2e-6 * T0
    so, sr = our.scalars(), ref.scalars()
    assert so["L"] == sr["L"] and so["t"] == 0 and so["av_iters"] == 0
    # same lattice -> same potential energy (the GPU evaluation against the reference's CPU evaluation)
    assert abs(so["V"] - sr["V"]) <= 1e-5 * max(1.0, abs(sr["V"]))
    # a second system with the same seed reproduces the state; another seed does not
    our2 = Reference(N, T0, rho, canonical, bc, legacy="product")
    assert np.array_equal(our2.get_state()[1], vo)
    monkeypatch.setenv("LJMD_SEED", "31338")
    our3 = Reference(N, T0, rho, canonical, bc, legacy="product")
    assert not np.array_equal(our3.get_state()[1], vo)
    for s in (ref, our, our2, our3):
        s.close()


@pytest.mark.parametrize("N,T0,rho,canonical,bc", CASES)
def test_velocity_helpers_match_reference_class(pkg, gpu_lib, N, T0, rho, canonical, bc):
    """KineticTemperature, CorrectTotalMomentum, RenormalizeVelocities, RenormalizeVelocitiesToEnergy
    (MDSystem.cpp:183-216,361-404) on identical host arrays: the loops are float/double host arithmetic and must
    give the same bits; K, U, T follow to the accuracy of V (GPU vs CPU evaluation)."""
    ref, our = both(N, T0, rho, canonical, bc)
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=4)
    vel = pkg.snapshots.velocities(N, 1.3 * T0, seed=4)
    vel[:, :3] += np.float32(0.05)                       # a net momentum to remove
    for s in (ref, our):
        s.set_state(pos, vel)
    assert our.kinetic_temperature() == ref.kinetic_temperature()
    for s in (ref, our):
        s.correct_total_momentum()
    vr, vo = ref.get_state()[1], our.get_state()[1]
    assert np.array_equal(vr, vo), "CorrectTotalMomentum differs"
    assert np.abs(vo[:, :3].astype(np.float64).sum(axis=0)).max() <= 1e-4
    for s in (ref, our):
        s.renormalize_velocities(True)
    vr, vo = ref.get_state()[1], our.get_state()[1]
    assert np.array_equal(vr, vo), "RenormalizeVelocities(true) differs"
    sr, so = ref.scalars(), our.scalars()
    assert so["T"] == sr["T"] == T0
    assert abs(so["K"] - sr["K"]) <= 1e-6 * sr["K"] and abs(so["U"] - sr["U"]) <= 1e-5 * (abs(sr["K"]) + abs(sr["V"]))
    # to a target energy per particle (the fluctuation tasks' start-up, run-fluctuations.cpp:62-66)
    ust = (sr["U"] / N) + 0.2
    for s in (ref, our):
        s.renormalize_to_energy(ust)
    vr, vo = ref.get_state()[1], our.get_state()[1]
    assert np.abs(vr[:, :3] - vo[:, :3]).max() <= 2e-6 * np.abs(vr[:, :3]).max()      # the factor carries V: GPU vs CPU sum
    sr, so = ref.scalars(), our.scalars()
    assert abs(so["U"] / N - ust) <= 1e-12 and abs(sr["U"] / N - ust) <= 1e-12
    assert abs(so["K"] - sr["K"]) <= 1e-5 * (abs(sr["K"]) + abs(sr["V"]))
    # one step from the renormalised state: the class uploads its host arrays (dirty velocities included)
    for s in (ref, our):
        s.integrate(0.004, 2)
    sr, so = ref.scalars(), our.scalars()
    assert abs(so["U"] - sr["U"]) <= 1e-4 * (abs(sr["K"]) + abs(sr["V"]))
    assert abs(so["T"] - sr["T"]) <= 1e-4 * sr["T"]
    assert so["av_iters"] == sr["av_iters"] == 2
    for s in (ref, our):
        s.close()


@pytest.mark.parametrize("N,T0,rho,canonical,bc", CASES[:2])
def test_velocity_histogram_running_mean(pkg, gpu_lib, N, T0, rho, canonical, bc):
    """initvelo / updatevelo / getvelo (MDSystem.cpp:651-694): counts come from the device histogram kernel, the
    running mean over veloIters is the reference's; identical velocities give identical curves, update after update."""
    ref, our = both(N, T0, rho, canonical, bc)
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=9)
    for s in (ref, our):
        s.set_state(pos, pkg.snapshots.velocities(N, T0, seed=9))
    xr, dr = ref.velocity_histogram(12.0, 0.12)
    xo, do = our.velocity_histogram(12.0, 0.12)
    assert len(xr) == len(xo) == 101 and np.array_equal(xr, xo) and np.array_equal(dr, do)
    assert abs(do.sum() * 0.12 - 1.0) <= 1e-12
    for k in range(3):
        v = pkg.snapshots.velocities(N, T0 * (1.0 + 0.2 * k), seed=20 + k)
        for s in (ref, our):
            s.poke_velocities(v)          # a caller editing h_Vel in place, then asking for the histogram update
            s.updatevelo()
        (xr, dr), (xo, do) = (ref._getvelo(), our._getvelo())
        assert np.array_equal(dr, do), f"running mean differs after update {k + 1}"
    for s in (ref, our):
        s.close()


@pytest.mark.parametrize("N,T0,rho", [(400, 1.4, 0.05), (4096, 1.0, 0.85), (100003, 0.7, 0.3)])
def test_device_side_initial_conditions(pkg, oracle, gpu_lib, N, T0, rho):
    """ljmd_init_state: the reference's start lattice bit for bit, Philox velocities with zero total
    momentum, T = T0 and Maxwell statistics; the seed fixes the state; the state is evaluated (forces, V, K)."""
    from scipy import stats
    with pkg.ljmd.LJSystem(N, T0=T0, rho=rho, canonical=True, bc=0) as s:
        s.init_state(2024)
        pos, vel, frc = s.get_state()
        assert np.array_equal(pos, oracle.lattice(N, s.L)), "start lattice differs from MDSystem.cpp:150-168"
        v3 = vel[:, :3].astype(np.float64)
        assert np.abs(v3.sum(axis=0)).max() <= 2e-5 * np.sqrt(N * T0)
        assert abs((v3 * v3).sum() / (3 * N) - T0) <= 2e-6 * T0
        assert stats.kstest(np.sqrt((v3 * v3).sum(axis=1)), stats.maxwell(scale=np.sqrt(T0)).cdf).pvalue > 1e-3
        for a in range(3):
            assert stats.kstest(v3[:, a], stats.norm(scale=np.sqrt(T0)).cdf).pvalue > 1e-3
        assert abs(np.corrcoef(v3[:, 0], v3[:, 1])[0, 1]) < 5.0 / np.sqrt(N)          # components are independent
        sc = s.scalars()
        assert abs(sc["T"] - T0) <= 2e-6 * T0 and sc["t"] == 0.0 and sc["av_iters"] == 0
        assert abs(2.0 * frc[:, 3].astype(np.float64).sum() - sc["V"]) <= 1e-6 * max(1.0, abs(sc["V"]))
        s.init_state(2024)
        p2, v2, _ = s.get_state()
        assert np.array_equal(p2, pos) and np.array_equal(v2, vel)
        s.init_state(2025)
        assert not np.array_equal(s.get_state()[1], vel)
        s.step(0.004, 3)                                      # and the state steps
        assert abs(s.scalars()["T"] - T0) <= 2e-2 * T0
