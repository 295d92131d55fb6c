"""GPU tests of the task drivers rebuilt on the library: lennard-jones-cuda_b200/tasks/bin/*.

Each driver is run on a short input next to the reference's own driver source compiled unmodified against the
product's MDSystem class (oracle/_ref/*, `make -C oracle dropin`).  With the same LJMD_SEED both start from the
same state, and batched device-resident stepping is bit-identical to single Integrate calls, so the two programs
must print the same numbers although one reads h_Pos / h_Vel after every step and the other reads a trace once
per thousand steps."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "lennard-jones-cuda_b200", "tasks", "bin")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
INPUT = os.path.join(ROOT, "tests", "data", "N400.short.input")


def need(path):
    if not os.path.exists(path):
        pytest.fail(f"{path} not built: run __graft_entry__.build()")
    return path


def table_rows(text, ncols):
    rows = []
    for ln in text.splitlines():
        f = ln.split()
        if len(f) == ncols:
            try:
                rows.append([float(v) for v in f])
            except ValueError:
                pass
    return np.array(rows)


def run(exe, args, cwd, seed="2024", timeout=900):
    env = dict(os.environ, LJMD_SEED=seed)
    out = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, timeout=timeout, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def output_file(cwd, suffix):
    names = [f for f in os.listdir(cwd) if f.endswith(suffix)]
    assert len(names) == 1, (suffix, os.listdir(cwd))
    return os.path.join(cwd, names[0])


def test_run_fluctuations_cli(tmp_path, gpu_lib):
    """BASELINE config #1 (N = 400, T* = 1.4, rho* = 0.05, periodic, TVN), 5000 equilibration + 2000 production
    steps.  Reference CPU build of the reference driver on this input: <T*> = 1.40007, <u*> = 1.728, <Z> = 0.869."""
    mine = tmp_path / "mine"
    mine.mkdir()
    text = run(need(os.path.join(BIN, "run-fluctuations")), [INPUT], mine)
    rows = table_rows(text, 8)
    assert rows.shape == (2, 8), text[-2000:]
    t, u, T, Z, uav, Tav, Zav, w = rows[-1]
    assert abs(t - 28.0) < 0.01 and abs(Tav - 1.4) < 3e-3 and abs(uav - 1.728) < 0.06 and abs(Zav - 0.869) < 0.06
    assert 0.0 < w < 1.0
    td = np.loadtxt(output_file(mine, ".TimeDep.txt"), skiprows=1)
    assert td.shape == (2, 13) and np.allclose(td[:, :7], rows[:, :7], rtol=1e-5)
    assert np.abs(td[:, 10:13]).max() < 1e-4                     # total momentum stays ~0
    rdf = np.loadtxt(output_file(mine, ".RDF.dat"), skiprows=1)
    assert rdf.shape == (256, 2) and np.isfinite(rdf).all()
    assert rdf[rdf[:, 0] < 0.8, 1].max() < 1e-3 and abs(rdf[rdf[:, 0] > 4.0, 1].mean() - 1.0) < 0.05
    fx = np.loadtxt(output_file(mine, ".flucsX.dat"), skiprows=1)
    assert fx.shape == (19, 6) and np.allclose(fx[:, 1], fx[:, 0] * 400, rtol=0.2, atol=8.0)   # <N> ~ alpha N (8 time units of a dilute gas)
    fv = np.loadtxt(output_file(mine, ".flucsVz.dat"), skiprows=1)
    assert fv.shape == (20, 7) and fv[-1, 2] > 0.99                                      # |vz| < 3 sigma: all

    exe_ref = os.path.join(REF_DIR, "run-fluctuations")
    if not os.path.exists(exe_ref):
        pytest.skip("reference driver (oracle/_ref/run-fluctuations) not built: compared against physics only")
    theirs = tmp_path / "theirs"
    theirs.mkdir()
    rows_ref = table_rows(run(exe_ref, [INPUT], theirs), 8)
    assert rows_ref.shape == rows.shape
    assert np.allclose(rows, rows_ref, rtol=1e-5, atol=1e-12), (rows, rows_ref)
    for suffix, tol in ((".TimeDep.txt", 1e-5), (".flucsX.dat", 1e-5), (".flucsY.dat", 1e-5), (".flucsZ.dat", 1e-5),
                        (".flucsCube.dat", 1e-5), (".flucsVz.dat", 1e-5), (".RDF.dat", 1e-5)):
        a = np.loadtxt(output_file(mine, suffix), skiprows=1)
        b = np.loadtxt(output_file(theirs, suffix), skiprows=1)
        assert a.shape == b.shape, suffix
        atol = 1e-7 if suffix == ".TimeDep.txt" else 1e-12      # mean velocities are ~1e-9 numbers
        assert np.allclose(a, b, rtol=tol, atol=atol, equal_nan=True), suffix


def test_run_isotherm_cli(tmp_path, gpu_lib):
    inp = tmp_path / "iso.input"
    inp.write_text("N 256\nT* 1.4\nrho*_min 0.1\nrho*_max 0.31\ndrho* 0.1\nteq 2.\ntfin 10.\ndt* 0.004\nuseCUDA 1\n")
    mine = tmp_path / "mine"
    mine.mkdir()
    text = run(need(os.path.join(BIN, "run-isotherm")), [str(inp)], mine)
    con = table_rows(text, 11)
    assert con.shape[0] == 3 * 2 and np.allclose(con[::2, 0], [0.1, 0.2, 0.3])          # 2000 observations per density
    dat = np.loadtxt(output_file(mine, ".dat"), skiprows=1)
    assert dat.shape == (3, 13)
    assert np.allclose(dat[:, 0], [0.1, 0.2, 0.3]) and np.allclose(dat[:, 2], 1.4, atol=0.01)
    assert (np.diff(dat[:, 3]) < 0).all()             # u* falls with density on this isotherm
    assert (dat[:, 9] < 1.0).all() and (dat[:, 9] > 0.3).all()      # Z below the ideal gas value, T* = 1.4 < T_Boyle
    exe_ref = os.path.join(REF_DIR, "run-isotherm")
    if not os.path.exists(exe_ref):
        pytest.skip("reference driver (oracle/_ref/run-isotherm) not built")
    theirs = tmp_path / "theirs"
    theirs.mkdir()
    con_ref = table_rows(run(exe_ref, [str(inp)], theirs), 11)
    assert np.allclose(con, con_ref, rtol=1e-5, atol=1e-12, equal_nan=True)
    dat_ref = np.loadtxt(output_file(theirs, ".dat"), skiprows=1)
    assert np.allclose(dat, dat_ref, rtol=1e-5, atol=1e-12, equal_nan=True)


def test_semigce_cli(tmp_path, gpu_lib):
    """10 events of 200 EVN steps after 10 000 TVN steps at the driver's hard-coded state point."""
    text = run(need(os.path.join(BIN, "semiGCEfluctuations")), ["10"], tmp_path, seed="7")
    lines = [ln.replace("+-", " ").split() for ln in text.splitlines() if ln.strip()]
    assert len(lines) in (19, 20)
    tab = np.array([[float(v) for v in f] for f in lines])
    assert (tab[:, 0] == 10).all()
    frac = 0.05 * np.arange(1, len(lines) + 1)
    assert np.allclose(tab[:, 1], frac * 512, rtol=0.2, atol=6.0)
    assert 0.05 < tab[9, 5] < 12.0        # 10 events at the LJ critical point (T* = 1.312, rho* = 0.316)
    exe_ref = os.path.join(REF_DIR, "semiGCEfluctuations")
    if not os.path.exists(exe_ref):
        pytest.skip("reference driver (oracle/_ref/semiGCEfluctuations) not built")
    env = dict(os.environ, LJMD_SEED="7")
    proc = subprocess.Popen([exe_ref], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    ref_lines = []
    try:
        for _ in range(25):
            ln = proc.stdout.readline()
            if not ln.strip():
                break
            ref_lines.append(ln.replace("+-", " ").split())
    finally:
        proc.kill()
        proc.wait()
    ref = np.array([[float(v) for v in f] for f in ref_lines])
    assert ref.shape == tab.shape
    assert np.allclose(tab, ref, rtol=1e-5, atol=2e-6, equal_nan=True)       # %lf prints six decimals
