"""CPU tests of the task-driver helpers: statistics estimators against golden vectors produced by
the reference's NumberStatistics / TimeAverage (oracle/make_golden_stats.py), and live against those classes
when oracle/_ref is built; parameter-file reader, output prefix and fraction grids against the reference's
documented behaviour (run-fluctuations-aux.h:26-121, 345-353, 446-456)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle.oracle import REF_LIB, reference_available

TASKS = os.path.join(ROOT, "lennard-jones-cuda_b200", "tasks")
LIB = os.path.join(TASKS, "libljmd_taskstats.so")
NAMES = ["mean", "mean_error", "variance", "variance_error", "scaled_variance", "scaled_variance_error", "skewness",
         "skewness_error", "kurtosis", "kurtosis_error", "inefficiency", "correlated_mean_error"]


@pytest.fixture(scope="module")
def tasklib():
    if not os.path.exists(LIB):
        pytest.fail(f"{LIB} not built: run __graft_entry__.build()")
    lib = C.CDLL(LIB)
    lib.ljtasks_series_statistics.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.ljtasks_parameter.restype = C.c_double
    lib.ljtasks_parameter.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int)]
    lib.ljtasks_prefix_tail.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    lib.ljtasks_coordinate_fractions.argtypes = [C.c_double, C.c_void_p, C.c_int]
    lib.ljtasks_momentum_cuts.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int]
    return lib


def stats(lib, fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(12)
    getattr(lib, fn)(x.ctypes.data, len(x), out.ctypes.data)
    return out


def assert_same(out, ref, rtol):
    assert np.array_equal(np.isnan(out), np.isnan(ref)), (out, ref)     # s = NaN when the lag-one covariance is <= 0
    ok = ~np.isnan(ref)
    err = np.abs(out[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
    assert err.max() <= rtol, dict(zip(np.array(NAMES)[ok], err))


def test_statistics_match_reference_golden_vectors(tasklib):
    z = np.load(os.path.join(ROOT, "tests", "golden", "stats", "series_statistics.npz"))
    off = 0
    for k, n in enumerate(z["n"]):
        x = z["x"][off:off + n]
        off += n
        assert_same(stats(tasklib, "ljtasks_series_statistics", x), z["out"][k], 1e-12)
    assert off == len(z["x"]) and len(z["n"]) >= 20


@pytest.mark.skipif(not reference_available(), reason="oracle/_ref/libljmd_ref.so not built")
def test_statistics_match_reference_live(tasklib):
    ref = C.CDLL(REF_LIB)
    if not hasattr(ref, "ljref_series_statistics"):
        pytest.skip("reference shim predates ljref_series_statistics")
    ref.ljref_series_statistics.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(99)
    for trial in range(30):
        n = int(rng.integers(3, 3000))
        if trial % 3 == 0:
            x = rng.poisson(rng.uniform(1, 200), size=n).astype(np.float64)
        elif trial % 3 == 1:
            x = np.rint(100 + 10 * np.convolve(rng.normal(size=n + 9), np.ones(10) / 10, mode="valid"))
        else:
            x = rng.normal(rng.uniform(-5, 5), rng.uniform(0.01, 2), size=n)
        assert_same(stats(tasklib, "ljtasks_series_statistics", x), stats(ref, "ljref_series_statistics", x), 1e-10)


def test_known_answers(tasklib):
    """Hand-checkable series: 1..5 has mean 3, population variance 2, mu3 = 0, mu4 = 6.8."""
    out = dict(zip(NAMES, stats(tasklib, "ljtasks_series_statistics", [1, 2, 3, 4, 5])))
    assert out["mean"] == 3.0 and abs(out["variance"] - 2.0) < 1e-14
    assert abs(out["mean_error"] - np.sqrt(2.0 / 5)) < 1e-14
    assert abs(out["scaled_variance"] - 2.0 / 3.0) < 1e-14
    assert abs(out["skewness"]) < 1e-13
    assert abs(out["kurtosis"] - (6.8 - 3 * 4.0) / 2.0) < 1e-13
    assert abs(out["variance_error"] - np.sqrt((6.8 - 4.0) / 5)) < 1e-13
    # lag-one covariance of 1..5: (2+6+12+20)/4 - 9 = 1 -> s = 2/ln(2)
    assert abs(out["inefficiency"] - 2.0 / np.log(2.0)) < 1e-13


def test_parameter_file_reader(tasklib, tmp_path):
    found = C.c_int(0)
    get = lambda path, kind, key: tasklib.ljtasks_parameter(path.encode(), kind, key.encode(), C.byref(found))
    # defaults (run-fluctuations-aux.h:35-48, run-isotherm-aux.h:35-46)
    assert get("", 0, "N") == 400 and get("", 0, "rho*") == 0.60 and get("", 0, "u*") == 1.708
    assert get("", 0, "tfin") == 1000. and get("", 0, "subvolume_spacing") == 0.05
    assert get("", 1, "rho*_min") == 0.60 and get("", 1, "drho*") == 0.01 and get("", 1, "tfin") == 5000.
    get("", 1, "canonical")
    assert found.value == 0
    # the repo's short copy of the reference's sample input
    inp = os.path.join(ROOT, "tests", "data", "N400.short.input")
    assert get(inp, 0, "rho*") == 0.05 and get(inp, 0, "teq") == 20. and get(inp, 0, "T*") == 1.4
    # comments, a commented-out key, unknown keys and output_prefix
    f = tmp_path / "p.input"
    f.write_text("# header line\nN 1000   # trailing comment\n# u* 3.0\nrho* 0.3\noutput_prefix myrun\nnewkey 7\ncanonical 0\nu* 2.5\n")
    assert get(str(f), 0, "N") == 1000 and get(str(f), 0, "rho*") == 0.3 and get(str(f), 0, "u*") == 2.5
    assert get(str(f), 0, "newkey") == 7 and found.value == 1
    buf = C.create_string_buffer(256)
    tasklib.ljtasks_prefix_tail(str(f).encode(), 0, buf, 256)
    assert buf.value.decode() == "myrun|.N1000.ust2.5.rhost0.3"          # microcanonical: labelled by u*
    tasklib.ljtasks_prefix_tail(inp.encode(), 0, buf, 256)
    assert buf.value.decode() == "run|.N400.Tst1.4.rhost0.05"
    tasklib.ljtasks_prefix_tail(b"", 1, buf, 256)
    assert buf.value.decode() == "isotherm.run|.N400.Tst1.4"


def test_fraction_grids(tasklib):
    out = np.zeros(256)
    n = tasklib.ljtasks_coordinate_fractions(0.05, out.ctypes.data, 256)
    grid = []
    a = 0.05
    while a <= 1. - 1.e-9:
        grid.append(a)
        a += 0.05
    assert n == len(grid) == 19 and np.array_equal(out[:n], grid)
    n = tasklib.ljtasks_momentum_cuts(1.4, 0.05, 3.0, out.ctypes.data, 256)
    assert n == 20 and abs(out[n - 1] - 3.0 * np.sqrt(1.4)) < 1e-12
    assert tasklib.ljtasks_coordinate_fractions(0.5, out.ctypes.data, 256) == 1 and out[0] == 0.5


@pytest.mark.parametrize("exe,args", [("run-fluctuations", ["tests/data/N400.short.input"]), ("run-isotherm", []),
                                      ("semiGCEfluctuations", ["1"])])
def test_task_drivers_fail_loudly_without_a_gpu(pkg, exe, args):
    """No CPU fallback: on a machine without a CUDA device the drivers say so and exit non-zero before any work."""
    import subprocess
    if pkg.ljmd.load_library().ljmd_device_count() > 0:
        pytest.skip("a GPU is present: the drivers run (tests/test_tasks_gpu.py)")
    path = os.path.join(TASKS, "bin", exe)
    if not os.path.exists(path):
        pytest.fail(f"{path} not built: run __graft_entry__.build()")
    out = subprocess.run([path] + [os.path.join(ROOT, a) if a.endswith(".input") else a for a in args],
                         capture_output=True, text=True, timeout=60, cwd=os.path.join(ROOT, "tests"))
    assert out.returncode != 0
    assert "CUDA device" in out.stderr
