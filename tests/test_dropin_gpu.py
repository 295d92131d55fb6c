"""GPU drop-in tests: UNMODIFIED reference code on top of the product.

  * oracle/_ref/libljmd_ref_legacy.so — the reference host layer (MDSystem.cpp, -DUSE_CUDA_TOOLKIT) linked
    against the product's legacy C seam (include/ljmd.h section B): its GPU branch runs the sm_100a kernels.
  * oracle/_ref/run-fluctuations, semiGCEfluctuations — the reference task drivers compiled unmodified against
    the product's source-compatible MDSystem.h (lennard-jones-cuda_b200/host) and linked to libljmd_host.so.

Both are built in the container where /root/reference exists (make -C oracle ref_legacy dropin) and travel
to the GPU box as binaries; the tests skip when they are absent.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle.oracle import REF_LEGACY_LIB, Reference, reference_available

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not (os.path.exists(REF_LEGACY_LIB) and reference_available()), reason="legacy rig not built")
@pytest.mark.parametrize("name", ["liquid_evn_periodic", "c1_gas_tvn_periodic", "gas_evn_hardwall"])
def test_reference_host_layer_drives_the_new_kernels(oracle, gpu_lib, name):
    g = load_golden(name)
    N = g["N"]
    cpu = Reference(N, g["T0"], g["rho"], g["canonical"], g["bc"])
    gpu = Reference(N, g["T0"], g["rho"], g["canonical"], g["bc"], legacy=True)
    cpu.set_state(g["pos0"], g["vel0"])
    gpu.set_state(g["pos0"], g["vel0"])
    _, _, fc = cpu.get_state()
    _, _, fg = gpu.get_state()
    L = cpu.scalars()["L"]
    f64, fabs_sum, sc64 = oracle.forces_f64(g["pos0"], float(np.float32(L)), g["bc"])
    err = np.abs(fg[:, :3].astype(np.float64) - f64).max(axis=1) / fabs_sum
    assert err.max() <= 2e-5     # the seam carries L as float (MDSystem.cpp:246): ulp(L)/2 on wrapped separations
    sc_c, sc_g = cpu.scalars(), gpu.scalars()
    assert abs(sc_g["V"] - sc_c["V"]) <= 2e-5 * sc64["Vabs"]
    assert abs(sc_g["K"] - sc_c["K"]) <= 1e-9 * sc_c["K"]
    vol = N / g["rho"]
    assert abs(sc_g["P"] - sc_c["P"]) <= 2e-5 * (sc64["Pabs"] + N * sc_c["T"]) / vol
    if g["bc"] != 0:   # with a float L only the un-imaged histogram is guaranteed identical
        assert np.array_equal(gpu.rdf_counts()[0], cpu.rdf_counts()[0])
    cpu.integrate(g["dt"], 3)
    gpu.integrate(g["dt"], 3)
    pc, vc, _ = cpu.get_state()
    pg, vg, _ = gpu.get_state()
    assert np.abs(pg[:, :3] - pc[:, :3]).max() <= 5e-6 * max(1.0, L)
    assert np.abs(vg[:, :3] - vc[:, :3]).max() <= 5e-5 * np.abs(vc[:, :3]).max()
    assert abs(gpu.scalars()["U"] - cpu.scalars()["U"]) <= 1e-4 * (sc64["Vabs"] + sc_c["K"])


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "run-fluctuations")), reason="drop-in tasks not built")
def test_run_fluctuations_unmodified(tmp_path, gpu_lib):
    """BASELINE config #1 (N=400, T*=1.4, rho*=0.05, periodic, TVN) through the reference's own driver.
    The reference CPU build of the same driver on the same input gives <T*> = 1.40007, <u*> = 1.728,
    <Z> = 0.869 (run in the build container; the repo's input file quotes u* = 1.708 for this state)."""
    exe = os.path.join(REF_DIR, "run-fluctuations")
    inp = os.path.join(ROOT, "tests", "data", "N400.short.input")
    env = dict(os.environ, LJMD_SEED="2024")
    out = subprocess.run([exe, inp], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [ln.split() for ln in out.stdout.splitlines() if len(ln.split()) == 8 and ln.split()[0][0].isdigit()]
    assert len(rows) == 2, out.stdout[-2000:]
    t, u, T, Z, uav, Tav, Zav, w = map(float, rows[-1])
    assert abs(t - 28.0) < 0.01
    assert abs(Tav - 1.4) < 3e-3
    assert abs(uav - 1.728) < 0.06
    assert abs(Zav - 0.869) < 0.06
    assert 0.0 < w < 1.0
    files = os.listdir(tmp_path)
    for suffix in (".TimeDep.txt", ".RDF.dat", ".flucsX.dat", ".flucsCube.dat", ".flucsVz.dat"):
        assert any(f.endswith(suffix) for f in files), suffix
    rdf = np.loadtxt(os.path.join(tmp_path, [f for f in files if f.endswith(".RDF.dat")][0]), skiprows=1)
    assert rdf.shape[0] > 50 and np.isfinite(rdf).all()


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "semiGCEfluctuations")), reason="drop-in tasks not built")
def test_semigce_driver_starts_and_reports(gpu_lib):
    """semiGCEfluctuations hard-codes N=512, 10 000 TVN steps then 10 000 x 200 EVN steps (minutes of GPU
    time in full); run it until its first report (10 events = 12 000 steps) and check the table it prints:
    20 sub-volume fractions, <N> = alpha*N, scaled variance over the binomial value of order one."""
    exe = os.path.join(REF_DIR, "semiGCEfluctuations")
    env = dict(os.environ, LJMD_SEED="7")
    proc = subprocess.Popen([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    lines = []
    try:
        for _ in range(25):
            ln = proc.stdout.readline()
            if not ln.strip():          # the report ends with a blank line
                break
            lines.append(ln.split())
    finally:
        proc.kill()
        proc.wait()
    assert len(lines) in (19, 20), lines      # 0.05 accumulated in double stops at 0.95 or 1.0
    for k, f in enumerate(lines):
        frac = 0.05 * (k + 1)
        assert int(f[0]) == 10
        mean = float(f[1])
        assert abs(mean - frac * 512) < 0.2 * frac * 512 + 6.0
    mid = lines[9]                       # alpha = 0.5
    assert 0.05 < float(mid[7]) < 12.0   # omega / (1 - alpha) after only 10 events, at the LJ critical point
