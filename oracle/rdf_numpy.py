"""TEST INFRASTRUCTURE ONLY — vectorised numpy restatement of the reference CPU path's RDF binning
(/root/reference/src/library/MDSystem.cpp:269-285, fast_round :732-739) for sizes where the O(N^2)
C oracle takes minutes.  Candidate pairs come from a k-d tree with a safety margin; every candidate
is then pushed through the reference's exact float32/float64 sequence (numpy float32 arithmetic is
IEEE and never fused), so the histogram equals the all-pairs one bit for bit.
tests/test_oracle.py pins it against oracle/ljmd_oracle.c at small N.
"""
import numpy as np

RDF_BINS = 256


def _fast_round(qf):
    """MDSystem.cpp:732-739 on float32 arrays."""
    half = np.float32(0.5)
    return np.where(qf > 0, np.trunc(qf + half), np.trunc(qf - half)).astype(np.int32)


def reference_r2(pos_i, pos_j, L, bc):
    """float32 r^2 of the reference for paired rows of xyz float32 arrays."""
    d = pos_i.astype(np.float32) - pos_j.astype(np.float32)            # :269-271 (float32)
    if bc == 0:
        n = _fast_round((d.astype(np.float64) / L).astype(np.float32))  # :274 (double divide, float arg)
        d = (d.astype(np.float64) - L * n.astype(np.float64)).astype(np.float32)
    sq = d * d                                                          # float32 products
    return (sq[:, 0] + sq[:, 1]) + sq[:, 2]                             # :279 left to right


def rdf_counts(pos4, L, bc, dr2, chunk=4_000_000):
    from scipy.spatial import cKDTree

    x = np.ascontiguousarray(pos4, dtype=np.float32).reshape(-1, 4)[:, :3]
    rmax = np.sqrt(RDF_BINS * float(np.float32(dr2))) * 1.001 + 1e-3
    xd = x.astype(np.float64)
    if bc == 0:
        xw = np.mod(xd, L)
        xw[xw >= L] = 0.0
        tree = cKDTree(xw, boxsize=L)
    else:
        tree = cKDTree(xd)
    pairs = tree.query_pairs(rmax, output_type="ndarray")
    counts = np.zeros(RDF_BINS, dtype=np.int64)
    dr2d = np.float64(np.float32(dr2))
    for s in range(0, len(pairs), chunk):
        p = pairs[s:s + chunk]
        for a, b in ((p[:, 0], p[:, 1]), (p[:, 1], p[:, 0])):            # ordered pairs: both (i,j) and (j,i)
            r2 = reference_r2(x[a], x[b], L, bc)
            k = np.floor(r2.astype(np.float64) / dr2d)                  # :282
            k = k[(k >= 0) & (k < RDF_BINS)].astype(np.int64)
            counts += np.bincount(k, minlength=RDF_BINS)
    return counts
