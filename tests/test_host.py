"""CPU: host-side logic of the product — the C-ABI library loads and exports everything the header
declares, fails loudly without a device, and its launch plan / exact-RDF constants are right.
No compute is launched here (there is no GPU in the build container)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle.rdf_numpy import reference_r2


def header_functions():
    src = open(os.path.join(ROOT, "include", "ljmd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.ljmd.load_library()
    declared = header_functions()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ljmd.h but not exported"
    assert sorted(pkg.ljmd.API_SYMBOLS) == declared


def test_no_cpu_fallback(pkg):
    """Without a device the product refuses to construct; with a missing library it refuses to load."""
    lib = pkg.ljmd.load_library()
    if lib.ljmd_device_count() == 0:
        with pytest.raises(pkg.ljmd.LJMDError, match="no CUDA device"):
            pkg.ljmd.LJSystem(64)
    with pytest.raises(pkg.ljmd.LJMDError, match="no CPU fallback"):
        pkg.ljmd.load_library(os.path.join(ROOT, "does_not_exist.so"))


def test_product_never_imports_oracle():
    pkgdir = os.path.join(ROOT, "lennard-jones-cuda_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "ljmd_oracle" not in text and "ljref_" not in text and "libljmd_ref" not in text, f


def test_argument_errors(pkg):
    lib = pkg.ljmd.load_library()
    h = C.c_void_p()
    assert lib.ljmd_create(C.byref(h), 1, 0.5, 1.0, 0, 0, C.c_float(0.1), 0) == 2      # N < 2
    assert b"N must be" in lib.ljmd_last_error()
    assert lib.ljmd_create(C.byref(h), 100, -1.0, 1.0, 0, 0, C.c_float(0.1), 0) == 2   # rho <= 0
    assert lib.ljmd_create(C.byref(h), 100, 0.5, 1.0, 0, 7, C.c_float(0.1), 0) == 2    # bad bc
    assert lib.ljmd_step(None, 0.004, 1, 0) == 2                                        # NULL handle
    buf = (C.c_int * 6)()
    assert lib.ljmd_plan(100, 3, 2, 148, buf) == 2


def test_rdf_dr2_rule(pkg, oracle):
    lib = pkg.ljmd.load_library()
    for N in [2, 50, 100, 128, 399, 400, 401, 512, 4096, 65536, 1048576]:
        assert float(lib.ljmd_rdf_dr2(N)) == oracle.rdf_dr2(N) == pkg.snapshots.rdf_dr2(N)
    assert float(lib.ljmd_rdf_dr2(400)) == float(np.float32(0.1))
    assert float(lib.ljmd_rdf_dr2(100)) == float(np.float32(0.2))


@pytest.mark.parametrize("N,world", [(400, 1), (16384, 1), (65536, 1), (262144, 8), (1048576, 8), (1000003, 8),
                                     (777, 4), (65536, 2)])
def test_launch_plan_tiles_the_problem(pkg, N, world):
    covered = 0
    for r in range(world):
        p = pkg.ljmd.plan(N, r, world, 148)
        assert p["i_begin"] == covered
        covered = p["i_end"]
        n_loc = p["i_end"] - p["i_begin"]
        if n_loc == 0:                                  # more ranks than blocks: trailing ranks are empty
            assert p["j_splits"] == 0
            continue
        assert p["i_begin"] % p["i_tile"] == 0          # shards are whole blocks
        assert p["i_tiles"] * p["i_tile"] >= n_loc > (p["i_tiles"] - 1) * p["i_tile"]
        assert 1 <= p["j_splits"] <= max(1, N // 8)
        assert p["newton3"] == (N >= 8 * 512 - 511)
        if p["newton3"]:
            g = pkg.ljmd.plan_newton3(N, r, world, 148)
            assert p["j_splits"] == g["nwin"] and p["force_ctas"] == g["n_super"] * g["nwin"]
            assert g["n_super"] == -(-p["i_tiles"] // g["mi"]) and 1 <= g["mi"] <= max(1, g["nblk"] // 2)
        else:
            assert p["force_ctas"] == p["i_tiles"] * p["j_splits"]
        if p["newton3"]:
            # many small CTAs (measured optimum, profiles/r02_tune_force_sym_split_scan.log): at least ~4.5 per
            # resident slot so the hardware scheduler can balance the SMs, at most ~60 so the partial-force rows
            # k_gather reads back stay a small fraction of the step
            slots = 148 * 3
            assert 4.4 * slots <= p["force_ctas"] <= 61 * slots
            # partial-force rows + reaction blocks of the super-tiles: a few GB at N = 1M on one GPU (round 1: 17 GB)
            cnt = -(-g["nblk"] // world) * 512
            assert (g["nwin"] * cnt + g["n_super"] * g["nwin"] * g["mju"] * g["bj"]) * 16 <= 4.0e9
    assert covered == N


def _partner_count(g, n):
    return (n - 1) // 2 if n & 1 else n // 2 - 1 + (1 if g < n // 2 else 0)


@pytest.mark.parametrize("N,world,sms", [(4096, 1, 148), (5000, 1, 148), (16384, 1, 148), (65536, 1, 148), (65536, 2, 148),
                                         (100003, 3, 148), (262144, 8, 148), (1048576, 1, 148), (1048576, 8, 148),
                                         (524288, 1, 16)])
def test_newton3_super_tiles_cover_every_block_pair_once(pkg, N, world, sms):
    """Python model of the work list of k_force_sym (csrc/ljmd_force_sym.cuh) and of reaction_sum
    (csrc/ljmd_step.cuh) for the planner's geometry: every unordered pair of 512-blocks is evaluated by exactly one
    CTA unit-by-unit, every block's diagonal once, and the gather finds, for every (block, chunk), exactly the
    reaction entries of the blocks that own it — all of them inside blocks the force kernel writes."""
    B = 512
    n = -(-N // B)
    done = {}                 # (I, J) -> set of chunks of J processed with I's particles in registers
    racc = {}                 # (rank, super, window, slot) -> set of (I, J, chunk)
    geo = []
    for r in range(world):
        g = pkg.ljmd.plan_newton3(N, r, world, sms)
        p = pkg.ljmd.plan(N, r, world, sms)
        geo.append((g, p))
        if p["i_end"] == p["i_begin"]:
            continue
        assert g["bj"] in (64, 128, 256) and g["nblk"] == n
        cpb = B // g["bj"]
        for a in range(g["n_super"]):
            I0 = g["blk0"] + a * g["mi"]
            ibase0 = p["i_begin"] + a * g["mi"] * B
            ntiles = min(g["mi"], -(-(p["i_end"] - ibase0) // B))
            for by in range(g["nwin"]):
                win = (by + g["win_shift"]) % g["nwin"]
                for slot in range(g["mju"]):
                    racc[(r, a, win, slot)] = set()
                for t in range(ntiles):
                    gI = (I0 + t) % n
                    ub = max(win * g["mju"], t * cpb)
                    ue = min((win + 1) * g["mju"], (t + _partner_count(gI, n) + 1) * cpb)
                    for u in range(ub, ue):
                        q = u // cpb
                        assert q < n
                        J = (I0 + q) % n
                        c = u % cpb
                        if min(g["bj"], min(N, (J + 1) * B) - (J * B + c * g["bj"])) <= 0:
                            continue
                        assert (J == gI) == (q == t)
                        key = (gI, J)
                        assert c not in done.setdefault(key, set()), "a unit is evaluated twice"
                        done[key].add(c)
                        if J != gI:
                            racc[(r, a, win, u - win * g["mju"])].add((gI, J, c))
    # coverage: the diagonal of every block, and every unordered pair under exactly one of its two orders
    bj = geo[0][0]["bj"]
    cpb = B // bj

    def chunks(J):
        return {c for c in range(cpb) if min(N, (J + 1) * B) - (J * B + c * bj) > 0}
    for I in range(n):
        assert done.get((I, I)) == chunks(I)
        for J in range(I + 1, n):
            a_, b_ = done.get((I, J)), done.get((J, I))
            assert (a_ is None) != (b_ is None), (I, J)
            assert (a_ == chunks(J)) if a_ is not None else (b_ == chunks(I))
    # the gather's index arithmetic (reaction_sum)
    hmax = (n - 1) // 2 if n & 1 else n // 2
    for J in range(0, n, max(1, n // 61)):
        for c in chunks(J):
            found = set()
            for r, (g, p) in enumerate(geo):
                if p["i_end"] == p["i_begin"]:
                    continue
                qmax = min(g["mi"] - 1 + hmax, n - 1)
                covering = [(t, (J - (g["blk0"] + t * g["mi"])) % n) for t in range(g["n_super"])]
                covering = [(t, q) for t, q in covering if q <= qmax]
                # reaction_sum walks two runs of t instead of testing every t: the same (t, q), in the same order,
                # also when 2, 4 or 8 lanes split a particle's super-tiles (t = first mod stride)
                for stride in (1, 2, 4, 8):
                    walked = []
                    for first in range(stride):
                        walked.append(_reaction_walk(J, g["blk0"], n, g["mi"], qmax, g["n_super"], first, stride))
                        assert walked[-1] == [(t, q) for t, q in covering if t % stride == first]
                for t, q in covering:
                    u = q * cpb + c
                    w = u // g["mju"]
                    # the record offset without the window: windows of a super-tile lie back to back
                    assert (t * g["nwin"] + w) * g["mju"] + (u - w * g["mju"]) == t * g["nwin"] * g["mju"] + u
                    found |= racc[(r, t, w, u - w * g["mju"])]       # KeyError = the kernel never writes it
            assert found == {(I, J, c) for I in range(n) if I != J and (I, J) in done}


def _reaction_walk(J, blk0, n, mi, qmax, n_super, first, stride):
    """The (t, q) sequence of reaction_sum() in csrc/ljmd_step.cuh, line by line."""
    D = J - blk0
    if D < 0:
        D += n
    out = []

    def walk(lo, hi, Dq):
        t = lo + ((first - lo) & (stride - 1))
        while t <= hi:
            out.append((t, Dq - t * mi))
            t += stride
    tD = D // mi
    walk((D - qmax + mi - 1) // mi if D > qmax else 0, min(n_super - 1, tD), D)
    walk(max(tD + 1, (D + n - qmax + mi - 1) // mi), n_super - 1, D + n)
    return out


def _emulate_image(d, L, thr1, thr2):
    """numpy model of image_exact() in csrc/ljmd_force.cuh."""
    a = np.abs(d)
    n = np.zeros(d.shape, dtype=np.float64)
    n[a >= thr1] = 1.0
    far = a >= thr2
    qf = (a[far].astype(np.float64) / L).astype(np.float32)
    n[far] = np.trunc(qf + np.float32(0.5)).astype(np.float64)
    n = np.where(d < 0, -n, n)
    out = (d.astype(np.float64) - L * n).astype(np.float32)
    return np.where(a < thr1, d, out)


def _emulate_bin(r2, dr2):
    """numpy model of rdf_bin_exact(): approximate quotient + exact-sign FMA residual fix-ups."""
    dr2 = np.float32(dr2)
    inv = np.float32(1.0) / dr2
    b = np.floor(r2 * inv).astype(np.float64)
    res = b * np.float64(dr2) - r2.astype(np.float64)             # exact in double (24x24-bit product)
    too_big = res > 0
    res2 = (b + 1.0) * np.float64(dr2) - r2.astype(np.float64)
    too_small = (~too_big) & (res2 <= 0)
    return b - too_big + too_small


@pytest.mark.parametrize("N,rho", [(400, 0.05), (16384, 0.85), (65536, 1.1), (1048576, 0.3), (512, 0.316)])
def test_exact_rdf_device_algorithm_matches_reference_semantics(pkg, N, rho):
    L = math.pow(N / rho, 1.0 / 3.0)
    thr1, thr2 = np.float32(pkg.ljmd.image_threshold(L, 1)), np.float32(pkg.ljmd.image_threshold(L, 2))
    rng = np.random.Generator(np.random.PCG64(99))
    n = 400_000
    xi = (rng.random((n, 3)) * L * 1.2 - 0.1 * L).astype(np.float32)
    xj = (rng.random((n, 3)) * L * 1.2 - 0.1 * L).astype(np.float32)
    # a quarter of the samples sit within a few ulp of the image threshold, a quarter far outside
    k = n // 4
    xj[:k] = xi[:k] - np.float32(0.5 * L) + (rng.integers(-8, 9, (k, 3)) * np.spacing(np.float32(L))).astype(np.float32)
    xj[k:2 * k] += (rng.integers(-3, 4, (k, 3)) * L).astype(np.float32)
    d = xi - xj
    img = _emulate_image(d, L, thr1, thr2)
    ref = d.astype(np.float32)
    nref = np.where(((d.astype(np.float64) / L).astype(np.float32)) > 0,
                    np.trunc((d.astype(np.float64) / L).astype(np.float32) + np.float32(0.5)),
                    np.trunc((d.astype(np.float64) / L).astype(np.float32) - np.float32(0.5)))
    ref = (d.astype(np.float64) - L * nref.astype(np.float64)).astype(np.float32)
    assert np.array_equal(img, ref)
    # bins: random r^2 plus values clustered on bin edges
    dr2 = pkg.snapshots.rdf_dr2(N)
    r2 = reference_r2(xi, xj, L, 0)
    edges = (np.arange(1, 300, dtype=np.float64) * np.float64(np.float32(dr2))).astype(np.float32)
    near = np.concatenate([np.nextafter(edges, np.float32(0)), edges, np.nextafter(edges, np.float32(1e9))])
    r2 = np.concatenate([r2, near, (rng.random(200_000) * 30).astype(np.float32)])
    want = np.floor(r2.astype(np.float64) / np.float64(np.float32(dr2)))
    assert np.array_equal(_emulate_bin(r2, dr2), want)


def test_image_threshold_is_the_first_float_that_rounds_up(pkg):
    for L in [7.3231346668217085, 19.999999999999996, 26.812275377187092, 151.76078099157598, 297.0]:
        for k in (1, 2):
            t = np.float32(pkg.ljmd.image_threshold(L, k))
            below = np.nextafter(t, np.float32(0))
            for x, want_ge in ((t, True), (below, False)):
                qf = np.float32(np.float64(x) / L)
                n = int(np.trunc(qf + np.float32(0.5)))
                assert (n >= k) == want_ge


def test_snapshots_are_deterministic_and_exact(pkg):
    snap = pkg.snapshots
    p1, v1 = snap.make("C2")
    p2, v2 = snap.make("C2")
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2)
    assert p1.shape == (16384, 4) and p1.dtype == np.float32
    L = snap.box_length(16384, 0.85)
    assert p1[:, :3].min() >= 0 and p1[:, :3].max() < L
    v = v1[:, :3].astype(np.float64)
    assert abs((v * v).sum() / (3 * 16384) - 1.0) < 1e-6 and np.abs(v.mean(axis=0)).max() < 1e-6
    g = snap.random_gas(2000, 0.3, min_sep=0.9, seed=5)
    from scipy.spatial import cKDTree
    assert len(cKDTree(g[:, :3].astype(np.float64), boxsize=snap.box_length(2000, 0.3) + 1e-6).query_pairs(0.899)) == 0


def test_host_class_is_source_compatible_with_the_reference_header():
    """The callers' view of MDSystem (tests/data/api_probe.cpp) compiles against the product's header, and —
    where the reference tree exists — against the reference's own header too, so the probe is a fair witness."""
    import subprocess
    probe = os.path.join(ROOT, "tests", "data", "api_probe.cpp")
    mine = os.path.join(ROOT, "lennard-jones-cuda_b200", "host")
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-I", mine, probe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = "/root/reference/src/library"
    if os.path.exists(os.path.join(ref, "MDSystem.h")):
        r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-w", "-I", ref, "-I", "/root/reference/thirdparty", probe],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]


def test_launch_plan_is_near_the_measured_optimum(pkg):
    """The Newton-3 launch planner against the split scan it was calibrated on (profiles/, measured on a B200):
    at every scanned size the number of splits it picks is one of the measured configurations, and that
    configuration ran within 4 % of the best one found for the size."""
    path = os.path.join(ROOT, "profiles", "r02_tune_force_sym_split_scan.log")
    table, N = {}, None
    for ln in open(path):
        m = re.match(r"== N=(\d+)", ln)
        if m:
            N = int(m.group(1))
            continue
        m = re.search(r"scan bj(\d+) S(\d+) per_cta(\d+) ctas(\d+).*?best\s+([\d.]+) ms", ln)
        if m and N is not None:
            table.setdefault(N, []).append((int(m.group(2)), float(m.group(5))))
    sizes = [n for n in sorted(table) if n >= 8192]
    assert len(sizes) >= 5
    for n in sizes:
        best = min(t for _, t in table[n])
        splits = pkg.ljmd.plan(n, 0, 1, 148)["j_splits"]
        mine = [t for s_, t in table[n] if s_ == splits]
        assert mine, (n, splits)
        assert min(mine) <= 1.04 * best, (n, splits, min(mine), best)


def test_library_and_torch_import_in_either_order():
    """libljmd.so and PyTorch both need `libnccl.so.2`; loading the library first must not break `import torch`
    (the binding preloads torch's bundled copy), and the other order is what bench.py does."""
    import subprocess
    import sys
    for first in ("lib", "torch"):
        code = ("import sys; sys.path.insert(0, %r); import ljpkg; pkg = ljpkg.load();\n" % ROOT +
                ("pkg.ljmd.load_library(); import torch\n" if first == "lib" else "import torch; pkg.ljmd.load_library()\n") +
                "print('ok', torch.__version__)")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "ok" in r.stdout, (first, r.stderr[-1500:])
