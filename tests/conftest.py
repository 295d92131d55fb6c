import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ljpkg  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SCAL_KEYS = ["U", "T", "K", "V", "P", "Pshear", "t", "L", "av_U_tot", "av_T_tot", "av_p_tot", "av_iters"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return ljpkg.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    N, T, rho, canonical, bc, dt, steps = g["cfg"]
    g.update(N=int(N), T0=float(T), rho=float(rho), canonical=int(canonical), bc=int(bc), dt=float(dt),
             steps=int(steps), dr2=float(g["dr2"]))
    g["s0"] = dict(zip(SCAL_KEYS, g["scal0"]))
    g["s1"] = dict(zip(SCAL_KEYS, g["scal1"]))
    return g


@pytest.fixture(params=["default", "ordered", "sym"])
def kernel(request, monkeypatch):
    """Force-kernel choice for systems created inside the test: the library's own (ordered below 8 blocks of 512 particles,
    Newton-3 above), or one of the two forced through LJMD_KERNEL (read at ljmd_create)."""
    if request.param == "default":
        monkeypatch.delenv("LJMD_KERNEL", raising=False)
    else:
        monkeypatch.setenv("LJMD_KERNEL", request.param)
    return request.param


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    """The CUDA library with a device behind it — GPU tests must never run on a fallback."""
    lib = pkg.ljmd.load_library()
    assert lib.ljmd_device_count() > 0, "no CUDA device visible: -m gpu tests need the B200 box"
    return lib
