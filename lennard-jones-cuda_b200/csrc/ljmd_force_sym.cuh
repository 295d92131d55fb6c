// Newton's-third-law variant of the all-pairs force kernel: every unordered pair {i,j} is evaluated ONCE and
// applied to both particles, which halves the pair evaluations of the reference's ordered double loop
// (/root/reference/src/library/MDSystem.cpp:260-302, MDSystem.cu:60-111 evaluate (i,j) and (j,i)).
//
// Work decomposition: particles are cut in blocks of B = 2*NPAIR*THREADS.  Block I interacts
// with itself (ordered loop, self pair excluded) and with the next h blocks cyclically, h = (n-1)/2 (+ the
// antipodal block for the lower half when the block count n is even): every unordered block pair appears
// exactly once and every i-tile has the same amount of work.
//
// Super-tiles.  A CTA owns a SUPER-TILE: `mi` consecutive i-tiles of this rank (first global block I0) times a
// WINDOW of `mju` consecutive j-units (chunks of bj records) counted from block I0: unit u covers records
// [((I0 + u / cpb) mod n) * B + (u % cpb) * bj, +bj), cpb = B / bj.  Tile t of the super-tile (block I0 + t) owns
// the units of blocks q = u / cpb with q == t (its diagonal) or 1 <= q - t <= partner_count(I0 + t).  The CTA
// walks its tiles one after the other (i-particles in registers, as in k_force) and, inside a tile, the owned
// units of its window; the reaction forces of the window's records accumulate in shared memory ACROSS the
// tiles and leave the CTA once, as one block of rpart; the direct forces of a tile leave it once per window, as
// one row segment of fpart.  Output per CTA is (mi*B + mju*bj) records for mi*B*mju*bj pair evaluations, so the
// partial-force traffic and footprint fall as 1/mi + 1/mju (round 1 had mi = 1 and one reaction row per unit:
// 17 GB at N = 1M; mi = 16, mju = 16: 3.3 GB).
//
// Reaction forces without atomics: inside a warp the 32 j-records of a chunk ROTATE through the lanes, so at
// every step each lane works on a different j: lane l handles record (l + k) mod 32 at step k.  The record is
// read from a per-warp staging copy of the chunk that is stored twice back to back, so the read address is
// base + (l + k) with no wrap arithmetic and does not depend on the previous step; the three reaction
// accumulators travel with their j through one shuffle each per step.  After 32 steps a packet is home and has
// met all 32 lanes x 2*NPAIR i-particles; the warps' packets go to per-warp shared-memory slices, are summed in
// warp order and added to the window's accumulator in tile order; the gather kernel adds the blocks of the
// super-tiles in a fixed order.  No floating-point atomics anywhere: runs are bit-reproducible.
// The CTA's (tile, unit) work list is built once in shared memory and the main loop just walks it: with the
// enumeration inlined in the loop (round-2 first attempt) ptxas kept ~12 more live scalars, stopped coalescing
// the accumulator registers across the back-edge of the pair loop and re-built the packed i-coordinate pairs
// inside it: +6 % instructions in the periodic loop, +27 % in the open one (5 % / 12 % of the kernel time).
// Measured alternatives at N = 65 536 periodic (profiles/r01_tune_force_sym_*.log): ordered kernel 2.82 ms;
// rotating positions and accumulators by shuffle (6 SHFL per step) 1.92 ms; broadcast j + 5-level butterfly
// sum of the reaction (15 SHFL + 15 FADD per j) 2.37 ms.
#pragma once
#include <type_traits>

#include "ljmd_force.cuh"

namespace ljmd {

struct SymParams {
#ifdef LJMD_PERTURB   // tools/ only: shifts every parameter offset to sample ptxas' schedules
  char perturb_[LJMD_PERTURB];
#endif
  ForceParams f;   // jrec, posf, fpart, blockW, rdf, N, i_begin (multiple of B), i_end, ilocal_cap, constants
  float4* rpart;   // [super-tiles of this rank][nwin][mju * bj] reaction sums (fx,fy,fz,0) of the window's records
  int ncols;       // unused (kept: the byte offsets of the fields below steer ptxas' schedule of the hot loop)
  int nblk;        // global number of blocks = ceil(N / B)
  int bj;          // j-records per unit (multiple of 32, divides B)
  // RDF pruning.  Appended here (not to ForceParams) so that every parameter offset the plain kernels read is
  // what it was before: ptxas' schedule of the hot loop shifts with them, worth 1-2 % of the kernel either way.
  float bbox_cut2;    // squared histogram range (real units, with margin) for the box-gap test
  const uint4* bbox;  // [nblk][2] per-block bounding boxes (lo.xyz, hi.xyz) or nullptr; periodic: fixed-point
                      // coordinates, open: float bits
  // super-tile geometry (see the header comment)
  int mi;          // i-tiles per super-tile
  int mju;         // j-units per window
  int nwin;        // windows per super-tile = gridDim.y
  int win_shift;   // blockIdx.y -> window (blockIdx.y + win_shift) mod nwin: the partial windows at both ends of
                   // the band are launched last, so the tail of the launch is made of the short CTAs
  // warp frames (FRAMES kernels, periodic boxes; see "Warp frames" below)
  int frames;      // 0: every chunk takes the fixed-point path (particle order not spatially sorted)
  float kunit;     // L / 2^32: fixed-point units -> sigma
  float far2;      // (R_far / kunit)^2: squared gap, in fixed-point units, from which a chunk may take the float path
  float rdf2;      // (histogram range / kunit)^2 with margin: a chunk whose every record is farther than this from
                   // the warp's box cannot hold an in-range pair and skips the RDF loop (FRAMES + RDF kernels)
};

// Units of a super-tile's band: tile t = 0..mi-1 owns blocks q in [t, t + h_t] counted from the super-tile's
// first block; q never reaches n (the planner keeps mi <= n/2), so a record appears at most once in the band.
__host__ __device__ inline int sym_band_units(int mi, int hmax, int n, int cpb) {
  int qmax = mi - 1 + hmax;
  if (qmax > n - 1) qmax = n - 1;
  return (qmax + 1) * cpb;
}

// number of partner offsets of global block g among n blocks
__host__ __device__ inline int sym_partner_count(int g, int n) {
  if (n & 1) return (n - 1) / 2;
  return n / 2 - 1 + (g < n / 2 ? 1 : 0);
}
inline int sym_max_partner_count(int n) { return (n & 1) ? (n - 1) / 2 : n / 2; }

// ---- Warp frames: the periodic minimum image without per-pair integer work -----------------------------------
// The fixed-point minimum image costs an integer subtract and an I2FP per axis and pair (and the c2 multiply that
// brings r^-2 back to sigma units): 7 of the ~25 issue slots of a pair.  When the particle order is spatially
// sorted (ljmd_core.cu re-sorts the records along a Hilbert curve every few hundred steps) the 128 i-particles a
// warp holds and the 32 j-records of a chunk are two small clouds, and for most (warp, chunk) combinations NO pair
// can come near half a box length on any axis.  For those the image is one translation: the warp keeps its
// i-particles as floats relative to the centre c of its own cloud, converts each staged j-record once to
// (float)((int)(u_j - c)) * L/2^32 — the 32-bit wrap of that one subtract is the image — and the pair loop is the
// open-box loop (one FADD2 per axis, no c2).  The decision is exact and per (warp, chunk): every lane checks its
// own record against the warp's box on the ring of 2^32 units, |u_j - c| + h < 2^31 per axis (no pair can wrap),
// one vote.  Chunks that fail, partial chunks and ragged tiles take the fixed-point loop, so the result never
// depends on the order being sorted — only the speed does.  Accuracy: a frame coordinate is a float of up to L/2,
// i.e. coarser than the fixed-point records near pairs need; so a chunk also has to be FAR — the gap between the
// warp's box and each record at least R_far = 2.5 sigma, where the force gradient is 1e-4 of the contact value —
// and the near chunks (a few per cent) keep the fixed-point path.
struct WarpFrame {
  int cx, cy, cz;              // centre of the warp's i-particles, fixed-point units (wrapping)
  int hx, hy, hz;              // half extents (+2 units of slack), < 2^30
  unsigned limx, limy, limz;   // 2^31 - h: a record with |u_j - c| below it cannot wrap against any i of the warp
  bool ok;                     // the warp's particles span less than half the box on every axis, none is a clamped duplicate
};

// Load the thread's 2*NPAIR i-particles, WARP-CONTIGUOUS: warp w holds particles [ibase + w*64*NPAIR, +64*NPAIR),
// lane l the ones at offsets l, 32 + l, 64 + l, ...; pair q = offsets (2q, 2q+1) * 32.  (k_force strides by
// THREADS instead; here a warp's particles must be neighbours in the sorted order so that their box is small.)
template <typename V, int THREADS, int NPAIR>
__device__ __forceinline__ int sym_i_offset(int m) {
  return ((int)threadIdx.x >> 5) * (64 * NPAIR) + m * 32 + ((int)threadIdx.x & 31);
}
template <typename V, bool PERIODIC, int THREADS, int NPAIR>
__device__ __forceinline__ bool load_i_particles_warp(const ForceParams& p, int ibase, PairI<V> (&pi)[NPAIR]) {
  bool all_valid = true;
#pragma unroll
  for (int q = 0; q < NPAIR; ++q) {
    int i0 = ibase + sym_i_offset<V, THREADS, NPAIR>(2 * q), i1 = i0 + 32;
    pi[q].v_lo = i0 < p.i_end;
    pi[q].v_hi = i1 < p.i_end;
    all_valid = all_valid && pi[q].v_lo && pi[q].v_hi;
    if (!pi[q].v_lo) i0 = p.i_end - 1;
    if (!pi[q].v_hi) i1 = p.i_end - 1;
    pi[q].i_lo = (unsigned)i0;
    pi[q].i_hi = (unsigned)i1;
    const uint4 r0 = p.jrec[i0], r1 = p.jrec[i1];
    pi[q].ax = (int)r0.x; pi[q].ay = (int)r0.y; pi[q].az = (int)r0.z;
    pi[q].bx = (int)r1.x; pi[q].by = (int)r1.y; pi[q].bz = (int)r1.z;
    if (!PERIODIC) {
      const float4 f0 = p.posf[i0], f1 = p.posf[i1];
      pi[q].x2 = mk2<V>(f0.x, f1.x); pi[q].y2 = mk2<V>(f0.y, f1.y); pi[q].z2 = mk2<V>(f0.z, f1.z);
    } else {
      pi[q].x2 = pi[q].y2 = pi[q].z2 = bc2<V>(0.f);
    }
  }
  return all_valid;
}

// The warp's frame from the fixed-point coordinates just loaded, and the float copies of the i-particles in it.
template <typename V, int NPAIR>
__device__ __forceinline__ void make_warp_frame(PairI<V> (&pi)[NPAIR], float kunit, bool warp_all_valid, WarpFrame& fr) {
  const unsigned full = 0xffffffffu;
  const int u0x = __shfl_sync(full, pi[0].ax, 0), u0y = __shfl_sync(full, pi[0].ay, 0), u0z = __shfl_sync(full, pi[0].az, 0);
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
#pragma unroll
  for (int q = 0; q < NPAIR; ++q) {
    const int r[6] = {pi[q].ax - u0x, pi[q].bx - u0x, pi[q].ay - u0y, pi[q].by - u0y, pi[q].az - u0z, pi[q].bz - u0z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = min(lo[a], min(r[2 * a], r[2 * a + 1]));
      hi[a] = max(hi[a], max(r[2 * a], r[2 * a + 1]));
    }
  }
  bool small = warp_all_valid;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(full, lo[a]);
    hi[a] = __reduce_max_sync(full, hi[a]);
    // every particle within a quarter box of the first one: the spread is below half a box and lo/hi did not wrap
    small = small && lo[a] > -(1 << 30) && hi[a] < (1 << 30);
  }
  fr.ok = small;
  fr.cx = (int)((unsigned)u0x + (unsigned)((lo[0] + hi[0]) >> 1));
  fr.cy = (int)((unsigned)u0y + (unsigned)((lo[1] + hi[1]) >> 1));
  fr.cz = (int)((unsigned)u0z + (unsigned)((lo[2] + hi[2]) >> 1));
  fr.hx = small ? ((hi[0] - lo[0]) >> 1) + 2 : 0;
  fr.hy = small ? ((hi[1] - lo[1]) >> 1) + 2 : 0;
  fr.hz = small ? ((hi[2] - lo[2]) >> 1) + 2 : 0;
  fr.limx = 0x80000000u - (unsigned)fr.hx;
  fr.limy = 0x80000000u - (unsigned)fr.hy;
  fr.limz = 0x80000000u - (unsigned)fr.hz;
#pragma unroll
  for (int q = 0; q < NPAIR; ++q) {
    pi[q].x2 = mk2<V>(__int2float_rn(pi[q].ax - fr.cx) * kunit, __int2float_rn(pi[q].bx - fr.cx) * kunit);
    pi[q].y2 = mk2<V>(__int2float_rn(pi[q].ay - fr.cy) * kunit, __int2float_rn(pi[q].by - fr.cy) * kunit);
    pi[q].z2 = mk2<V>(__int2float_rn(pi[q].az - fr.cz) * kunit, __int2float_rn(pi[q].bz - fr.cz) * kunit);
  }
}

// ---- RDF pruning by block bounding boxes ----------------------------------------------------------------------
// One CTA per block of B particles: component-wise min / max of the j-records (fixed-point for periodic boxes,
// floats otherwise).  A block that straddles the periodic wrap simply gets a box spanning the axis.
template <bool PERIODIC, int B>
__global__ void __launch_bounds__(128) k_bbox(const uint4* __restrict__ jrec, int N, uint4* __restrict__ bbox) {
  __shared__ unsigned int slo[4][3], shi[4][3];
  const int base = blockIdx.x * B;
  unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  float flo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, fhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int k = threadIdx.x; k < B; k += 128) {
    const int i = base + k;
    if (i < N) {
      const uint4 r = jrec[i];
      const unsigned int c[3] = {r.x, r.y, r.z};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (PERIODIC) { lo[a] = min(lo[a], c[a]); hi[a] = max(hi[a], c[a]); }
        else { flo[a] = fminf(flo[a], __uint_as_float(c[a])); fhi[a] = fmaxf(fhi[a], __uint_as_float(c[a])); }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (PERIODIC) {
        lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
      } else {
        flo[a] = fminf(flo[a], __shfl_xor_sync(0xffffffffu, flo[a], o));
        fhi[a] = fmaxf(fhi[a], __shfl_xor_sync(0xffffffffu, fhi[a], o));
      }
    }
    if ((threadIdx.x & 31) == 0) {
      slo[threadIdx.x >> 5][a] = PERIODIC ? lo[a] : __float_as_uint(flo[a]);
      shi[threadIdx.x >> 5][a] = PERIODIC ? hi[a] : __float_as_uint(fhi[a]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int L3[3], H3[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (PERIODIC) {
        L3[a] = min(min(slo[0][a], slo[1][a]), min(slo[2][a], slo[3][a]));
        H3[a] = max(max(shi[0][a], shi[1][a]), max(shi[2][a], shi[3][a]));
      } else {
        L3[a] = __float_as_uint(fminf(fminf(__uint_as_float(slo[0][a]), __uint_as_float(slo[1][a])),
                                      fminf(__uint_as_float(slo[2][a]), __uint_as_float(slo[3][a]))));
        H3[a] = __float_as_uint(fmaxf(fmaxf(__uint_as_float(shi[0][a]), __uint_as_float(shi[1][a])),
                                      fmaxf(__uint_as_float(shi[2][a]), __uint_as_float(shi[3][a]))));
      }
    }
    bbox[2 * blockIdx.x] = make_uint4(L3[0], L3[1], L3[2], 0u);
    bbox[2 * blockIdx.x + 1] = make_uint4(H3[0], H3[1], H3[2], 0u);
  }
}

// Can any particle of box a be within the histogram range of any particle of box b?  Per axis the gap between
// the two intervals (on the ring of 2^32 fixed-point units for periodic boxes), summed in quadrature.
template <bool PERIODIC>
__device__ __forceinline__ bool boxes_in_range(const uint4& alo, const uint4& ahi, const uint4& blo, const uint4& bhi,
                                               double L, float cut2) {
  const unsigned int al[3] = {alo.x, alo.y, alo.z}, ah[3] = {ahi.x, ahi.y, ahi.z};
  const unsigned int bl[3] = {blo.x, blo.y, blo.z}, bh[3] = {bhi.x, bhi.y, bhi.z};
  float g2 = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float gap;
    if (PERIODIC) {
      const bool overlap = (bl[a] <= ah[a]) && (al[a] <= bh[a]);
      const unsigned int d1 = bl[a] - ah[a], d2 = al[a] - bh[a];   // modulo 2^32: the two ways round the ring
      gap = overlap ? 0.f : (float)min(d1, d2) * (float)(L * (1.0 / 4294967296.0));
    } else {
      const float d1 = __uint_as_float(bl[a]) - __uint_as_float(ah[a]);
      const float d2 = __uint_as_float(al[a]) - __uint_as_float(bh[a]);
      gap = fmaxf(0.f, fmaxf(d1, d2));
    }
    g2 = fmaf(gap, gap, g2);
  }
  return g2 <= cut2;
}

// One pair of i-particles (two lanes of V) against the broadcast j-record; also accumulates this lane's
// reaction on j.  KILL: per-lane flags zero the interaction (clamped duplicate i's of a ragged last tile).
template <typename V, bool PERIODIC, bool KILL, bool RDF>
__device__ __forceinline__ void pair_sym(const uint4& uj, const PairI<V>& pi, PairAcc<V>& acc, bool kill_lo,
                                         bool kill_hi, float& rjx, float& rjy, float& rjz, const ForceParams& p,
                                         unsigned jglobal, RdfCtx& R) {
  V dx, dy, dz;
  if (PERIODIC) {
    dx = mk2<V>(__int2float_rn(pi.ax - (int)uj.x), __int2float_rn(pi.bx - (int)uj.x));
    dy = mk2<V>(__int2float_rn(pi.ay - (int)uj.y), __int2float_rn(pi.by - (int)uj.y));
    dz = mk2<V>(__int2float_rn(pi.az - (int)uj.z), __int2float_rn(pi.bz - (int)uj.z));
  } else {
    dx = sub2(pi.x2, bc2<V>(__uint_as_float(uj.x)));
    dy = sub2(pi.y2, bc2<V>(__uint_as_float(uj.y)));
    dz = sub2(pi.z2, bc2<V>(__uint_as_float(uj.z)));
  }
  const V r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
  const float2 r2s = upk(r2);
  float inv0 = rcp_approx(r2s.x), inv1 = rcp_approx(r2s.y);
  if (KILL) {
    if (kill_lo) inv0 = 0.f;
    if (kill_hi) inv1 = 0.f;
  }
  V x = mk2<V>(inv0, inv1);
  if (PERIODIC) x = mul2(x, bc2<V>(p.c2));
  const V x2 = mul2(x, x);
  const V r6 = mul2(x2, x);
  const V t = fma2(r6, bc2<V>(12.f), bc2<V>(-6.f));
  const V u = mul2(r6, t);
  const V s = mul2(u, x);
  acc.fx = fma2(dx, s, acc.fx);
  acc.fy = fma2(dy, s, acc.fy);
  acc.fz = fma2(dz, s, acc.fz);
  acc.s6 = add2(acc.s6, r6);
  acc.w = add2(acc.w, u);
  // reaction on j: -(d*s) of both lanes, scalar FMAs into this lane's partial
  const float2 sx = upk(s), ddx = upk(dx), ddy = upk(dy), ddz = upk(dz);
  rjx = __fmaf_rn(-ddx.x, sx.x, rjx); rjx = __fmaf_rn(-ddx.y, sx.y, rjx);
  rjy = __fmaf_rn(-ddy.x, sx.x, rjy); rjy = __fmaf_rn(-ddy.y, sx.y, rjy);
  rjz = __fmaf_rn(-ddz.x, sx.x, rjz); rjz = __fmaf_rn(-ddz.y, sx.y, rjz);
  if (RDF) {
    // the reference counts (i,j) and (j,i); its float sequence is odd in the separation, so both land in the
    // same bin: one queue entry with the "counts twice" bit (31) set
    rdf_push<PERIODIC, false>(R, r2s.x < p.cut_fast && !(KILL && kill_lo), r2s.y < p.cut_fast && !(KILL && kill_hi),
                       (pi.i_lo - R.ioff) | 0x80000000u, (pi.i_hi - R.ioff) | 0x80000000u, jglobal - R.joff, p);
  }
}

// grid: (i-tiles of this rank, splits).  block: THREADS.  dyn smem: force_sym_smem_bytes().
// The chunk loop of one partner unit, RU = build the RDF for it.  A macro, not a lambda or a helper template:
// the plain kernel must compile to exactly the loop it had before the pruned path existed (wrapping it in a
// generic lambda cost 2.3 % of the kernel through a different instruction schedule, same instruction mix).
#define LJMD_SYM_PARTNER_CHUNKS(RU)                                                                              \
  {                                                                                                              \
    const int nchunk = (nj + 31) >> 5;                                                                           \
    for (int c = 0; c < nchunk; ++c) {                                                                           \
      const int jl = (c << 5) + lane;                                                                            \
      /* stage the chunk twice back to back: step k reads entry lane + k, no wrap arithmetic */                  \
      uint4 rec = tu[min(jl, nj - 1)];                                                                           \
      const bool full = warp_all_valid && ((c << 5) + 32 <= nj);                                                 \
      bool fl = false;                                                                                           \
      bool ru_chunk = (RU);   /* build the RDF for this chunk? */                                                \
      if (FRAMES && fr.ok && full) {                                                                             \
        /* my record against the warp's box on the ring: can any pair of this chunk wrap?  is it far enough? */ \
        const int rx = (int)rec.x - fr.cx, ry = (int)rec.y - fr.cy, rz = (int)rec.z - fr.cz;                     \
        const unsigned axu = (unsigned)abs(rx), ayu = (unsigned)abs(ry), azu = (unsigned)abs(rz);                \
        const float gx = fmaxf(0.f, __int2float_rn((int)axu - fr.hx));                                           \
        const float gy = fmaxf(0.f, __int2float_rn((int)ayu - fr.hy));                                           \
        const float gz = fmaxf(0.f, __int2float_rn((int)azu - fr.hz));                                           \
        const float g2 = __fmaf_rn(gz, gz, __fmaf_rn(gy, gy, gx * gx));   /* squared gap record <-> box */        \
        /* RDF: no record of the chunk within histogram range of the warp's box -> no pair can count */          \
        if (RU) ru_chunk = __any_sync(0xffffffffu, g2 <= sp.rdf2);                                               \
        if (!ru_chunk) {                                                                                         \
          const bool okj = axu < fr.limx && ayu < fr.limy && azu < fr.limz && g2 >= sp.far2;                     \
          fl = __all_sync(0xffffffffu, okj);                                                                     \
          if (fl)                                                                                                \
            rec = make_uint4(__float_as_uint(__int2float_rn(rx) * sp.kunit), __float_as_uint(__int2float_rn(ry) * sp.kunit), \
                             __float_as_uint(__int2float_rn(rz) * sp.kunit), 0u);                                \
        }                                                                                                        \
      }                                                                                                          \
      if (FRAMES && fl != cur_float) {                                                                           \
        fold_forces(cur_float ? 4.f : p.fscale);   /* the accumulators change units with the path */             \
        cur_float = fl;                                                                                          \
      }                                                                                                          \
      __syncwarp();                                                                                              \
      mystage[lane] = rec;                                                                                       \
      mystage[lane + 32] = rec;                                                                                  \
      __syncwarp();                                                                                              \
      const uint4* sp_l = mystage + lane;                                                                        \
      float rjx = 0.f, rjy = 0.f, rjz = 0.f;                                                                     \
      const unsigned nxt_lane = (lane + 1) & 31;                                                                 \
      if (FRAMES && fl) {                                                                                        \
        _Pragma("unroll UNROLLK")                                                                                \
        for (int k = 0; k < 32; ++k) {                                                                           \
          const uint4 uj = sp_l[k];                                                                              \
          _Pragma("unroll")                                                                                      \
          for (int q = 0; q < NPAIR; ++q)                                                                        \
            pair_sym<V, false, false, false>(uj, pi[q], acc[q], false, false, rjx, rjy, rjz, p, 0u, R);          \
          rjx = __shfl_sync(0xffffffffu, rjx, nxt_lane);                                                         \
          rjy = __shfl_sync(0xffffffffu, rjy, nxt_lane);                                                         \
          rjz = __shfl_sync(0xffffffffu, rjz, nxt_lane);                                                         \
        }                                                                                                        \
      } else if (full && (RU) && ru_chunk) {                                                                     \
        /* two rotation steps per trip, ONE drain check: the loop body must stay inside the instruction cache */ \
        _Pragma("unroll 1")                                                                                      \
        for (int k = 0; k < 32; k += 2) {                                                                        \
          _Pragma("unroll")                                                                                      \
          for (int kk = 0; kk < 2; ++kk) {                                                                       \
            const uint4 uj = sp_l[k + kk];                                                                       \
            const unsigned jg = (unsigned)(j0 + (c << 5) + ((lane + k + kk) & 31));                              \
            _Pragma("unroll")                                                                                    \
            for (int q = 0; q < NPAIR; ++q)                                                                      \
              pair_sym<V, PERIODIC, false, RU>(uj, pi[q], acc[q], false, false, rjx, rjy, rjz, p, jg, R);        \
            rjx = __shfl_sync(0xffffffffu, rjx, nxt_lane);                                                       \
            rjy = __shfl_sync(0xffffffffu, rjy, nxt_lane);                                                       \
            rjz = __shfl_sync(0xffffffffu, rjz, nxt_lane);                                                       \
          }                                                                                                      \
          if (R.n >= 32) rdf_drain<PERIODIC>(R, p, false);                                                       \
        }                                                                                                        \
      } else if (full) {                                                                                         \
        _Pragma("unroll UNROLLK")                                                                                \
        for (int k = 0; k < 32; ++k) {                                                                           \
          const uint4 uj = sp_l[k];                                                                              \
          _Pragma("unroll")                                                                                      \
          for (int q = 0; q < NPAIR; ++q)                                                                        \
            pair_sym<V, PERIODIC, false, false>(uj, pi[q], acc[q], false, false, rjx, rjy, rjz, p, 0u, R);       \
          rjx = __shfl_sync(0xffffffffu, rjx, nxt_lane);                                                         \
          rjy = __shfl_sync(0xffffffffu, rjy, nxt_lane);                                                         \
          rjz = __shfl_sync(0xffffffffu, rjz, nxt_lane);                                                         \
        }                                                                                                        \
      } else {                                                                                                   \
        for (int k = 0; k < 32; ++k) {                                                                           \
          const uint4 uj = sp_l[k];                                                                              \
          const int hl = (lane + k) & 31; /* home lane of the j I work on now */                                 \
          const bool jdead = ((c << 5) + hl) >= nj;                                                              \
          const unsigned jg = (unsigned)(j0 + min((c << 5) + hl, nj - 1));                                       \
          _Pragma("unroll")                                                                                      \
          for (int q = 0; q < NPAIR; ++q)                                                                        \
            pair_sym<V, PERIODIC, true, RU>(uj, pi[q], acc[q], jdead || !pi[q].v_lo, jdead || !pi[q].v_hi, rjx,  \
                                            rjy, rjz, p, jg, R);                                                 \
          rjx = __shfl_sync(0xffffffffu, rjx, nxt_lane);                                                         \
          rjy = __shfl_sync(0xffffffffu, rjy, nxt_lane);                                                         \
          rjz = __shfl_sync(0xffffffffu, rjz, nxt_lane);                                                         \
          if (RU) { if (R.n >= 32) rdf_drain<PERIODIC>(R, p, false); }                                           \
        }                                                                                                        \
      }                                                                                                          \
      /* the accumulators are home again; jl < BJ always.  12-byte entries: stride 3 words, conflict-free.       \
         FRAMES kernels scale here (the two paths accumulate in different units), the others at the very end */  \
      if (FRAMES) {                                                                                              \
        const float rs = fl ? 4.f : p.fscale;                                                                    \
        rjx *= rs; rjy *= rs; rjz *= rs;                                                                         \
      }                                                                                                          \
      myslice[3 * jl] = rjx; myslice[3 * jl + 1] = rjy; myslice[3 * jl + 2] = rjz;                                \
    }                                                                                                            \
  }

// Work items of a CTA, built once in shared memory: (first record, record count, tile, window slot, diagonal?)
// packed in an int2: x = j0, y = nj | slot << 10 | tile << 16 | diag << 24 | first-of-tile << 25.
constexpr int kSymMaxItems = 256;   // mi * mju <= 16 * 16

// FRAMES (periodic only): chunks that qualify take the float loop in the warp's frame (see "Warp frames" above).
template <typename V, bool PERIODIC, bool RDF, int THREADS, int MINB, int NPAIR, int UNROLLK = 4, bool FRAMES = false>
__global__ void __launch_bounds__(THREADS, MINB) k_force_sym(const SymParams sp) {
  static_assert(!FRAMES || PERIODIC, "warp frames replace the periodic minimum image only");
  pdl_trigger();
  pdl_wait();
  constexpr int IPT = 2 * NPAIR;
  constexpr int B = THREADS * IPT;
  constexpr int NW = THREADS / 32;
  const ForceParams& p = sp.f;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int BJ = sp.bj;
  uint4* tile_u = reinterpret_cast<uint4*>(smem_raw);  // [2][BJ]
  unsigned char* after_tiles = smem_raw + (size_t)2 * BJ * 16;
  uint4* stage = reinterpret_cast<uint4*>(after_tiles);                          // [NW][64]
  float* slices = reinterpret_cast<float*>(after_tiles + (size_t)NW * 64 * 16);  // [NW][BJ][3]
  float* racc = slices + (size_t)NW * BJ * 3;                                    // [mju][BJ][3] window accumulator
  unsigned char* tail = reinterpret_cast<unsigned char*>(racc + (size_t)sp.mju * BJ * 3);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);          // [2]
  double* red = reinterpret_cast<double*>(tail + 16);          // [NW]
  int2* items = reinterpret_cast<int2*>(tail + 16 + 8 * NW);   // [kSymMaxItems]
  int* nitems_s = reinterpret_cast<int*>(tail + 16 + 8 * NW + kSymMaxItems * 8);   // [4]: item count
  unsigned int* hist = reinterpret_cast<unsigned int*>(tail + 32 + 8 * NW + kSymMaxItems * 8);         // [NW][256] (RDF)
  uint2* queues = reinterpret_cast<uint2*>(tail + 32 + 8 * NW + kSymMaxItems * 8 + NW * kRdfBins * 4);  // [NW][cap] (RDF)
  // RDF: float positions of the unit's j-records (double-buffered like tile_u) and of the tile in registers
  float4* tile_f = reinterpret_cast<float4*>(tail + 32 + 8 * NW + kSymMaxItems * 8 + NW * kRdfBins * 4 + NW * kRdfQueueCap * 8);  // [2][BJ]
  float4* itile_f = tile_f + (size_t)2 * BJ;                                                                                       // [B]

  const int ibase0 = p.i_begin + blockIdx.x * sp.mi * B;   // first particle of the super-tile
  const int ntiles = min(sp.mi, (p.i_end - ibase0 + B - 1) / B);   // the rank's last super-tile may be short
  int win = (int)blockIdx.y + sp.win_shift;
  if (win >= sp.nwin) win -= sp.nwin;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    // the CTA's work list: tile-major, empty units (beyond N: ragged last block) skipped
    const int n = sp.nblk, cpb = B / BJ, I0 = ibase0 / B;   // i_begin is a multiple of B
    const int u0 = win * sp.mju, u1 = u0 + sp.mju;
    int cnt = 0;
    for (int t = 0; t < ntiles; ++t) {
      int g = I0 + t;
      if (g >= n) g -= n;
      const int ub = max(u0, t * cpb), ue = min(u1, (t + sym_partner_count(g, n) + 1) * cpb);   // blocks q in [t, t + h_t]
      bool first = true;
      for (int u = ub; u < ue; ++u) {
        const int q = u / cpb;
        int J = I0 + q;
        if (J >= n) J -= n;
        const int j0 = J * B + (u - q * cpb) * BJ;
        const int nj = min(BJ, min(p.N, (J + 1) * B) - j0);
        if (nj <= 0) continue;
        items[cnt++] = make_int2(j0, nj | ((u - u0) << 10) | (t << 16) | ((q == t) ? (1 << 24) : 0) | (first ? (1 << 25) : 0));
        first = false;
      }
    }
    nitems_s[0] = cnt;
  }
  if (RDF) {
    for (int k = tid; k < NW * kRdfBins; k += THREADS) hist[k] = 0u;
  }
  for (int k = tid; k < sp.mju * BJ * 3; k += THREADS) racc[k] = 0.f;
  __syncthreads();
  const int nitems = nitems_s[0];

  auto issue = [&](int it, int st) {
    const int2 d = items[it];
    const uint32_t bytes = (uint32_t)(d.y & 1023) * 16u;
    mbar_expect_tx(&bars[st], RDF ? 2u * bytes : bytes);
    bulk_g2s(tile_u + (size_t)st * BJ, p.jrec + d.x, bytes, &bars[st]);
    if (RDF) bulk_g2s(tile_f + (size_t)st * BJ, p.posf + d.x, bytes, &bars[st]);
  };
  if (nitems > 0 && tid == 0) issue(0, 0);

  PairI<V> pi[NPAIR];
  PairAcc<V> acc[NPAIR];
  V s6run[NPAIR], wrun[NPAIR], fxrun[NPAIR], fyrun[NPAIR], fzrun[NPAIR];
  const V zero2 = bc2<V>(0.f);
#pragma unroll
  for (int q = 0; q < NPAIR; ++q) {
    acc[q].fx = acc[q].fy = acc[q].fz = acc[q].s6 = acc[q].w = zero2;
    s6run[q] = wrun[q] = fxrun[q] = fyrun[q] = fzrun[q] = zero2;
    pi[q].ax = pi[q].ay = pi[q].az = pi[q].bx = pi[q].by = pi[q].bz = 0;
    pi[q].x2 = pi[q].y2 = pi[q].z2 = zero2;
    pi[q].i_lo = pi[q].i_hi = 0u;
    pi[q].v_lo = pi[q].v_hi = false;
  }
  RdfCtx R;
  R.q = queues + warp * kRdfQueueCap;
  R.hist = hist + warp * kRdfBins;
  R.n = 0;
  R.pa = itile_f; R.pb = tile_f;
  R.ioff = R.joff = 0u;
  float* myslice = slices + (size_t)warp * BJ * 3;
  uint4* mystage = stage + warp * 64;
  double wsum = 0.;
  float4* out = p.fpart + (size_t)blockIdx.y * p.ilocal_cap;
  unsigned int stored = 0u;      // tiles whose row segment has been written
  int ibase = ibase0;            // first particle of the tile in registers
  int cur_tile = -1;
  bool warp_all_valid = false;
  uint4 my_lo = make_uint4(0u, 0u, 0u, 0u), my_hi = my_lo;   // bounding box of my own block (RDF pruning)
  WarpFrame fr;                  // FRAMES: this warp's frame for the tile in registers
  fr.cx = fr.cy = fr.cz = fr.hx = fr.hy = fr.hz = 0; fr.limx = fr.limy = fr.limz = 0u; fr.ok = false;
  bool cur_float = false;        // FRAMES: the tile-level force accumulators hold float-path (sigma) units
  // FRAMES: tile-level force accumulators -> run-level ones, in final units (the two paths differ by L/2^32)
  auto fold_forces = [&](float sc) {
    const V s2 = bc2<V>(sc);
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) {
      fxrun[q] = fma2(acc[q].fx, s2, fxrun[q]);
      fyrun[q] = fma2(acc[q].fy, s2, fyrun[q]);
      fzrun[q] = fma2(acc[q].fz, s2, fzrun[q]);
      acc[q].fx = acc[q].fy = acc[q].fz = zero2;
    }
  };

  // the tile in registers is done for this window: its partial forces leave the CTA
  auto store_tile = [&]() {
    const float fs = FRAMES ? 1.f : p.fscale;   // FRAMES: the run-level sums are already in final units
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) {
      const float2 fx = upk(fxrun[q]), fy = upk(fyrun[q]), fz = upk(fzrun[q]);
      const float2 s6 = upk(s6run[q]), w = upk(wrun[q]);
      const int il = (ibase - p.i_begin) + sym_i_offset<V, THREADS, NPAIR>(2 * q);
      // r^-12 - r^-6 = u/12 - r^-6/2
      if (pi[q].v_lo) {
        out[il] = make_float4(fx.x * fs, fy.x * fs, fz.x * fs, w.x * (1.f / 12.f) - 0.5f * s6.x);
        wsum += (double)w.x;
      }
      if (pi[q].v_hi) {
        out[il + 32] = make_float4(fx.y * fs, fy.y * fs, fz.y * fs, w.y * (1.f / 12.f) - 0.5f * s6.y);
        wsum += (double)w.y;
      }
      s6run[q] = wrun[q] = fxrun[q] = fyrun[q] = fzrun[q] = zero2;
    }
    stored |= 1u << cur_tile;
  };

  for (int it = 0; it < nitems; ++it) {
    if (it + 1 < nitems && tid == 0) issue(it + 1, (it + 1) & 1);
    const int2 d = items[it];
    const int j0 = d.x, nj = d.y & 1023;
    const bool diag = (d.y >> 24) & 1;
    if ((d.y >> 25) & 1) {
      // ---- next tile: the finished one leaves, its successor's i-particles come into the registers
      if (cur_tile >= 0) store_tile();
      cur_tile = (d.y >> 16) & 255;
      ibase = ibase0 + cur_tile * B;
      const bool all_valid = load_i_particles_warp<V, PERIODIC, THREADS, NPAIR>(p, ibase, pi);
      // is every lane of this warp holding real particles? (warp-uniform choice of the unmasked fast path)
      warp_all_valid = __all_sync(0xffffffffu, all_valid);
      if (FRAMES) {
        if (sp.frames) make_warp_frame<V, NPAIR>(pi, sp.kunit, warp_all_valid, fr);
        else fr.ok = false;
      }
      if (RDF) {
        // the drains read the float positions of this warp's own i-particles: a copy in shared memory
        // (the queue is empty here: it is flushed at the end of every unit)
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) {
          // (clamped duplicates of a ragged last tile never queue a pair; they must not touch the entry of the
          // particle they duplicate either: it belongs to another warp, which may be reading it)
          if (pi[q].v_lo) itile_f[pi[q].i_lo - (unsigned)ibase] = p.posf[pi[q].i_lo];
          if (pi[q].v_hi) itile_f[pi[q].i_hi - (unsigned)ibase] = p.posf[pi[q].i_hi];
        }
        __syncwarp();
        R.ioff = (unsigned)ibase;
      }
      if (RDF && sp.bbox != nullptr) { const int gI = ibase / B; my_lo = sp.bbox[2 * gI]; my_hi = sp.bbox[2 * gI + 1]; }
    }
    const int st = it & 1;
    mbar_wait(&bars[st], (uint32_t)((it >> 1) & 1));
    const uint4* tu = tile_u + (size_t)st * BJ;
    if (RDF) { R.pb = tile_f + (size_t)st * BJ; R.joff = (unsigned)j0; }
    float wgt;
    if (diag) {
      // ---- diagonal block: ordered loop over its own particles, self pair excluded ----
      wgt = 1.f;
      if (FRAMES && cur_float) { fold_forces(4.f); cur_float = false; }   // the ordered loop is fixed-point
      const int jrel0 = j0 - ibase - sym_i_offset<V, THREADS, NPAIR>(0);   // j relative to my particle m = 0
#pragma unroll 2
      for (int j = 0; j < nj; ++j) {
        const uint4 uj = tu[j];
        const int jr = jrel0 + j;
#pragma unroll
        for (int q = 0; q < NPAIR; ++q)
          pair_body<V, PERIODIC, true, RDF>(uj, pi[q], acc[q], jr == (2 * q) * 32, jr == (2 * q + 1) * 32,
                                            p, (unsigned)(j0 + j), R);
      }
    } else {
      // ---- partner block: each unordered pair once, reaction accumulators travel with the rotating j ----
      wgt = 2.f;
      if (RDF && FRAMES) {
        // sorted records: the per-chunk test against the warp's own box (tighter than two 512-block boxes) decides
        LJMD_SYM_PARTNER_CHUNKS(true)
      } else if (RDF) {
        // bounding boxes of the two blocks farther apart than the histogram range: no pair of this unit can
        // count, run the plain loop (uniform: every thread of the CTA sees the same two boxes)
        bool near = true;
        if (sp.bbox != nullptr) {
          const int J = j0 / B;
          near = boxes_in_range<PERIODIC>(my_lo, my_hi, sp.bbox[2 * J], sp.bbox[2 * J + 1], p.L, sp.bbox_cut2);
        }
        if (near) { LJMD_SYM_PARTNER_CHUNKS(true) } else { LJMD_SYM_PARTNER_CHUNKS(false) }
      } else {
        LJMD_SYM_PARTNER_CHUNKS(false)
      }
    }
    // fold the unit's tile-level accumulators into the run-level ones (two-level float summation);
    // an unordered pair of a partner block stands for two ordered pairs in the potential / virial sums
    const V w2 = bc2<V>(wgt);
    if (FRAMES) fold_forces(cur_float ? 4.f : p.fscale);
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) {
      s6run[q] = fma2(acc[q].s6, w2, s6run[q]);
      wrun[q] = fma2(acc[q].w, w2, wrun[q]);
      if (!FRAMES) {
        fxrun[q] = add2(fxrun[q], acc[q].fx);
        fyrun[q] = add2(fyrun[q], acc[q].fy);
        fzrun[q] = add2(fzrun[q], acc[q].fz);
      }
      acc[q].s6 = acc[q].w = acc[q].fx = acc[q].fy = acc[q].fz = zero2;
    }
    if (RDF) rdf_drain<PERIODIC>(R, p, true);   // queued pairs point into this unit's stage: flush before it is refilled
    __syncthreads();  // slices complete; stage st free for the load after next
    if (!diag) {
      // reaction of this unit: sum the warps' slices in warp order, add to the window accumulator
      float* ra = racc + (size_t)((d.y >> 10) & 63) * BJ * 3;
      for (int k = tid; k < nj * 3; k += THREADS) {
        float a = slices[k];
#pragma unroll
        for (int w = 1; w < NW; ++w) a += slices[(size_t)w * BJ * 3 + k];
        ra[k] += a;
      }
      __syncthreads();  // slices are rewritten by the next partner unit
    }
  }
  if (cur_tile >= 0) store_tile();
  // tiles that own no unit of this window (triangles at the ends of the band) still owe their zero rows
  for (int t = 0; t < ntiles; ++t) {
    if ((stored >> t) & 1u) continue;
#pragma unroll
    for (int m = 0; m < IPT; ++m) {
      const int i = ibase0 + t * B + m * THREADS + tid;
      if (i < p.i_end) out[i - p.i_begin] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // ---- the window's reaction sums leave the CTA: one block of rpart, every entry written (zeros included),
  // so the gather kernels never read stale data and nothing has to be cleared between steps
  {
    float4* dst = sp.rpart + ((size_t)blockIdx.x * sp.nwin + win) * ((size_t)sp.mju * BJ);
    const float fs = FRAMES ? 1.f : p.fscale;   // FRAMES: the slices were scaled chunk by chunk
    for (int k = tid; k < sp.mju * BJ; k += THREADS)
      dst[k] = make_float4(racc[3 * k] * fs, racc[3 * k + 1] * fs, racc[3 * k + 2] * fs, 0.f);
  }
  const double wtot = block_sum<THREADS>(wsum, red);
  if (tid == 0) p.blockW[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = wtot;
  if (RDF) {
    rdf_drain<PERIODIC>(R, p, true);
    __syncthreads();
    for (int b = tid; b < kRdfBins; b += THREADS) {
      unsigned int c = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) c += hist[w * kRdfBins + b];
      if (c) atomicAdd(&p.rdf[b], (unsigned long long)c);
    }
  }
}
#undef LJMD_SYM_PARTNER_CHUNKS

inline size_t force_sym_smem_bytes(bool rdf, int bj, int threads, int mju) {
  const int nw = threads / 32;
  // j-chunk double buffer, per-warp rotation stages, per-warp reaction slices (12 B), window accumulator (12 B),
  // barriers, block-sum scratch, work list + its length, RDF scratch
  return (size_t)2 * bj * 16 + (size_t)nw * 64 * 16 + (size_t)nw * bj * 12 + (size_t)mju * bj * 12 + 16 + 8 * nw +
         kSymMaxItems * 8 + 16 + (rdf ? rdf_smem_bytes(threads) + (size_t)2 * bj * 16 + (size_t)threads * 4 * 16 : 0);
}

}  // namespace ljmd
