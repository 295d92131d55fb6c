"""Warp frames of the Newton-3 kernel (ljmd_force_sym.cuh, "Warp frames") and the Hilbert record order
(ljmd_sort.cuh / ljmd_hilbert.cuh): CPU-side checks of the integer logic the kernels rely on.

The float path of a (warp, chunk) combination is only exact if NO pair of it can wrap: the 32-bit wrapped
difference u_i - u_j must equal (u_i - c) - (u_j - c) computed without wrap.  The model below restates the device
rule line by line (make_warp_frame + the per-record test of LJMD_SYM_PARTNER_CHUNKS) in numpy with 64-bit
integers, and checks that claim — for sorted, shuffled and adversarial clouds (half a box apart, across the seam).
"""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lennard-jones-cuda_b200", "csrc")
M32 = 1 << 32


def wrap32(x):
    """two's-complement int32 of an int64 array (what a 32-bit subtract leaves)"""
    return ((x + (1 << 31)) % M32) - (1 << 31)


def make_warp_frame(ui):
    """ui: [128, 3] uint32 coordinates of a warp's i-particles.  Returns (ok, c, h) as the device does."""
    u0 = ui[0].astype(np.int64)
    r = wrap32(ui.astype(np.int64) - u0)                       # pi.ax - u0x (32-bit wrap)
    lo, hi = r.min(0), r.max(0)
    small = bool(np.all(lo > -(1 << 30)) and np.all(hi < (1 << 30)))
    c = (u0 + ((lo + hi) >> 1)) % M32                          # wrapping add
    h = ((hi - lo) >> 1) + 2 if small else np.zeros(3, np.int64)
    return small, c, h


def chunk_eligible(uj, c, h, far2_units):
    """uj: [32, 3] uint32.  The per-record test and the warp vote."""
    r = wrap32(uj.astype(np.int64) - c)                        # (int)rec.x - fr.cx
    a = np.abs(r)                                              # abs(INT_MIN) stays 2^31 as unsigned
    lim = (1 << 31) - h
    ok = np.all(a < lim, axis=1)
    g = np.maximum(0.0, (a - h).astype(np.float64))
    far = (g * g).sum(1) >= far2_units
    return bool(np.all(ok & far)), r


def clouds(rng, L, centre, width, n):
    x = (np.asarray(centre) + (rng.random((n, 3)) - 0.5) * width) % L
    return (np.rint(x * (M32 / L)).astype(np.int64) % M32).astype(np.uint32)


@pytest.mark.parametrize("case", ["random", "half_box_apart", "across_seam", "wide_warp", "touching"])
def test_eligible_chunks_never_wrap(case):
    rng = np.random.default_rng(11)
    L = 39.06
    far2 = (2.5 * M32 / L) ** 2
    n_elig = n_tot = 0
    for trial in range(400):
        wi = rng.uniform(0.5, 8.0)
        wj = rng.uniform(0.5, 6.0)
        ci = rng.random(3) * L
        if case == "random":
            cj = rng.random(3) * L
        elif case == "half_box_apart":       # the clouds' separation sits right at L/2 on one or more axes
            cj = ci + np.where(rng.random(3) < 0.6, L / 2 + rng.uniform(-6, 6, 3), rng.uniform(-L / 2, L / 2, 3))
        elif case == "across_seam":          # both clouds hug the periodic seam
            ci = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.random() * L])
            cj = ci + rng.uniform(-12, 12, 3)
        elif case == "wide_warp":            # an unsorted warp: spans most of the box, must be refused
            wi = L * rng.uniform(0.4, 1.0)
            cj = rng.random(3) * L
        else:                                # touching clouds: never FAR
            cj = ci + rng.uniform(-1, 1, 3)
        ui = clouds(rng, L, ci, wi, 128)
        uj = clouds(rng, L, cj, wj, 32)
        ok, c, h = make_warp_frame(ui)
        n_tot += 1
        if not ok:
            continue
        elig, rj = chunk_eligible(uj, c, h, far2)
        if not elig:
            continue
        n_elig += 1
        ri = wrap32(ui.astype(np.int64) - c)
        assert np.all(np.abs(ri) <= h), "the frame's box must hold every i-particle"
        d_frame = ri[:, None, :] - rj[None, :, :]                          # what the float path subtracts
        d_image = wrap32(ui.astype(np.int64)[:, None, :] - uj.astype(np.int64)[None, :, :])   # the minimum image
        assert np.array_equal(d_frame, d_image), "an eligible chunk holds a pair whose image the frame gets wrong"
        # and it is FAR: every pair at least R_far apart
        r2 = ((d_image * (L / M32)) ** 2).sum(2)
        assert r2.min() >= 2.5 ** 2 * (1 - 1e-6)
    if case in ("random", "half_box_apart", "across_seam"):
        assert n_elig > 20, f"the test must exercise eligible chunks ({n_elig} of {n_tot})"
    if case in ("wide_warp", "touching"):
        assert n_elig == 0


def test_hilbert_index_is_a_bijection_of_adjacent_cells(tmp_path):
    """ljmd_hilbert.cuh on the host: every cell exactly once, consecutive indices are face neighbours."""
    src = tmp_path / "h.cpp"
    src.write_text(r'''
#include "ljmd_hilbert.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
int main() {
  using namespace ljmd;
  for (int bits = 1; bits <= 6; ++bits) {
    const int n = 1 << bits, m = n * n * n;
    std::vector<int> cx(m, -1), cy(m), cz(m);
    for (int x = 0; x < n; ++x) for (int y = 0; y < n; ++y) for (int z = 0; z < n; ++z) {
      const uint32_t k = hilbert3(x, y, z, bits);
      if (k >= (uint32_t)m || cx[k] != -1) { printf("collision at bits=%d\n", bits); return 1; }
      cx[k] = x; cy[k] = y; cz[k] = z;
    }
    for (int k = 1; k < m; ++k)
      if (abs(cx[k] - cx[k - 1]) + abs(cy[k] - cy[k - 1]) + abs(cz[k] - cz[k - 1]) != 1) { printf("jump at bits=%d k=%d\n", bits, k); return 1; }
  }
  printf("ok\n");
  return 0;
}
''')
    exe = tmp_path / "h"
    subprocess.run(["g++", "-O2", "-I", CSRC, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout
