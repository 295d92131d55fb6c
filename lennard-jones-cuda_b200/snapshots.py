"""Seeded synthetic snapshots for parity tests and benchmarks.

Positions follow the reference's start lattice (/root/reference/src/library/MDSystem.cpp:147-168);
velocities are seeded Gaussians with the mean removed and rescaled to exactly T*
(mirroring CorrectTotalMomentum / RenormalizeVelocities, MDSystem.cpp:183-216,375-389) because the
reference's own generator is time-seeded and cannot be reproduced.  Arrays use the reference layout:
float32 [N, 4] = (x, y, z, w), w = L/150 for positions (MDSystem.cpp:168), 0 for velocities.
"""
import math

import numpy as np


def box_length(N, rho):
    """MDSystem.cpp:70."""
    return math.pow(N / rho, 1.0 / 3.0)


def rdf_dr2(N):
    """MDSystem.cpp:93-95 (float member)."""
    dr2 = np.float32(max(0.2 * math.sqrt(100.0 / N), 0.05))
    if np.float32(250) * dr2 < 25.0:
        dr2 = np.float32(25.0 / 250)
    return float(dr2)


def lattice(N, rho, jitter=0.0, seed=12345):
    """Simple-cubic start lattice, optionally with a uniform jitter of +-jitter*dL per coordinate."""
    L = box_length(N, rho)
    ns = int(math.ceil(math.pow(N, 1.0 / 3.0)))
    dL = L / ns
    idx = np.arange(N, dtype=np.int64)
    pos = np.zeros((N, 4), dtype=np.float64)
    pos[:, 0] = ((idx % ns) + 0.5) * dL
    pos[:, 1] = (((idx // ns) % ns) + 0.5) * dL
    pos[:, 2] = ((idx // (ns * ns)) + 0.5) * dL
    if jitter > 0.0:
        rng = np.random.Generator(np.random.PCG64(seed + 1))
        pos[:, :3] += (rng.random((N, 3)) * 2.0 - 1.0) * jitter * dL
    out = pos.astype(np.float32)
    out[:, 3] = np.float32(L / np.float32(150.0))
    return out


def random_gas(N, rho, min_sep=0.9, seed=12345, periodic=True):
    """Uniform random positions with a minimum separation (rejection, re-drawing offenders)."""
    from scipy.spatial import cKDTree

    L = box_length(N, rho)
    rng = np.random.Generator(np.random.PCG64(seed + 2))
    x = rng.random((N, 3)) * L
    for _ in range(200):
        tree = cKDTree(x, boxsize=L if periodic else None)
        pairs = tree.query_pairs(min_sep, output_type="ndarray")
        if len(pairs) == 0:
            break
        bad = np.unique(pairs[:, 1])
        x[bad] = rng.random((len(bad), 3)) * L
    else:
        raise RuntimeError("could not place particles with the requested minimum separation")
    out = np.empty((N, 4), dtype=np.float32)
    out[:, :3] = x.astype(np.float32)
    # float32 rounding can land exactly on L; keep strictly inside the box
    out[:, :3] = np.minimum(out[:, :3], np.nextafter(np.float32(L), np.float32(0)))
    out[:, 3] = np.float32(L / np.float32(150.0))
    return out


def velocities(N, T, seed=12345):
    """Gaussian N(0, T) per component, zero total momentum, kinetic temperature exactly T."""
    rng = np.random.Generator(np.random.PCG64(seed))
    v = rng.standard_normal((N, 3)) * math.sqrt(T)
    v -= v.mean(axis=0, keepdims=True)
    tkin = (v * v).sum() / (3.0 * N)
    v *= math.sqrt(T / tkin)
    out = np.zeros((N, 4), dtype=np.float32)
    out[:, :3] = v.astype(np.float32)
    return out


# The five BASELINE.json configurations.  `kind` picks the position generator.
CONFIGS = {
    "C1": dict(N=400, T=1.4, rho=0.05, bc=0, canonical=True, kind="lattice", rdf_every=0),
    "C2": dict(N=16384, T=1.0, rho=0.85, bc=0, canonical=False, kind="lattice", rdf_every=0),
    "C3": dict(N=65536, T=1.0, rho=1.1, bc=0, canonical=True, kind="lattice", rdf_every=10),
    "C4": dict(N=262144, T=1.0, rho=0.01, bc=1, canonical=False, kind="lattice", rdf_every=0),
    "C5": dict(N=1048576, T=1.0, rho=0.3, bc=0, canonical=True, kind="lattice", rdf_every=0),
}


def make(name_or_cfg, seed=12345, jitter=0.05):
    cfg = CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
    if cfg.get("kind", "lattice") == "gas":
        pos = random_gas(cfg["N"], cfg["rho"], seed=seed, periodic=cfg["bc"] == 0)
    else:
        pos = lattice(cfg["N"], cfg["rho"], jitter=jitter, seed=seed)
    vel = velocities(cfg["N"], cfg["T"], seed=seed)
    return pos, vel
