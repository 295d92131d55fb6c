// All-pairs Lennard-Jones force / potential / virial (/ RDF) kernel for sm_100a.
//
// Replaces the reference's `calculateForces` + `RDFmerger` kernels
// (/root/reference/src/library/MDSystem.cu:21-153) and its CPU double loop
// (/root/reference/src/library/MDSystem.cpp:252-310).  Not a port: see DESIGN.md.
//
//  * i-particles live in registers, 2*NPAIR per thread, as packed f32x2 pairs so the
//    FP32-pipe work issues as FFMA2/FMUL2/FADD2 (one issue slot, two lanes).
//  * j-records stream through shared memory in double-buffered tiles filled by
//    1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier.
//  * periodic boxes: j-records are 32-bit fixed-point box fractions, so the
//    minimum image is the wrap of a 32-bit integer subtract (IADD3 on the ALU
//    pipe) and I2FP turns the difference into a float.  No division, no rounding
//    instruction, no rsqrt; one MUFU.RCP per pair.
//  * open / hard-wall boxes: j-records are the float positions, FADD2 differences.
//  * potential and virial are two packed accumulators per pair (sum r^-6 and
//    sum u, u = 12 r^-12 - 6 r^-6), reduced by warp shuffles to one double per CTA.
//  * RDF (on request): pairs whose fast r^2 is within the histogram range are pushed
//    (ballot + popc compaction) on a per-warp shared-memory queue; whenever 32 are
//    waiting, all 32 lanes re-derive r^2 with the reference CPU path's exact
//    float/double sequence and hit a per-warp shared-memory histogram; bins are
//    bit-exact with MDSystem.cpp:269-285.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ljmd {

typedef unsigned long long u64;

// ---- two-lane FP32 values -------------------------------------------------------------------------
// P2: one 64-bit register pair driven by the packed f32x2 instructions of sm_100 (PTX ISA 8.6:
//     fma/mul/add/sub.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2): one issue slot for two lanes.
// S2: the same two lanes as two scalar instructions (kept for A/B measurements in tools/tune_force.cu).
struct P2 { u64 v; };
struct S2 { float lo, hi; };

__device__ __forceinline__ void pk(P2& r, float lo, float hi) {
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
}
__device__ __forceinline__ float2 upk(const P2& a) {
  float2 o; asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(a.v)); return o;
}
__device__ __forceinline__ P2 fma2(const P2& a, const P2& b, const P2& c) {
  P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
}
__device__ __forceinline__ P2 mul2(const P2& a, const P2& b) {
  P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
}
__device__ __forceinline__ P2 add2(const P2& a, const P2& b) {
  P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
}
__device__ __forceinline__ P2 sub2(const P2& a, const P2& b) {
  P2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
}

__device__ __forceinline__ void pk(S2& r, float lo, float hi) { r.lo = lo; r.hi = hi; }
__device__ __forceinline__ float2 upk(const S2& a) { return make_float2(a.lo, a.hi); }
__device__ __forceinline__ S2 fma2(const S2& a, const S2& b, const S2& c) {
  S2 r; r.lo = __fmaf_rn(a.lo, b.lo, c.lo); r.hi = __fmaf_rn(a.hi, b.hi, c.hi); return r;
}
__device__ __forceinline__ S2 mul2(const S2& a, const S2& b) {
  S2 r; r.lo = __fmul_rn(a.lo, b.lo); r.hi = __fmul_rn(a.hi, b.hi); return r;
}
__device__ __forceinline__ S2 add2(const S2& a, const S2& b) {
  S2 r; r.lo = __fadd_rn(a.lo, b.lo); r.hi = __fadd_rn(a.hi, b.hi); return r;
}
__device__ __forceinline__ S2 sub2(const S2& a, const S2& b) {
  S2 r; r.lo = __fsub_rn(a.lo, b.lo); r.hi = __fsub_rn(a.hi, b.hi); return r;
}
template <typename V>
__device__ __forceinline__ V mk2(float lo, float hi) { V r; pk(r, lo, hi); return r; }
template <typename V>
__device__ __forceinline__ V bc2(float a) { V r; pk(r, a, a); return r; }

__device__ __forceinline__ float rcp_approx(float x) {
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in
// the stream still runs: pdl_trigger() lets the NEXT kernel begin launching, pdl_wait() blocks until the
// PREVIOUS one has completed and its writes are visible.  Both are no-ops for a plain launch.  Used by the
// small-system step chain (force -> gather -> finish), where launch latency is most of the step.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier + 1-D bulk copy ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  int spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    // a bulk copy that never lands must fail the launch, not hang the device
    if (!done && ++spins > (1 << 22)) __trap();
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- kernel parameters -------------------------------------------------------------------------
constexpr int kRdfBins = 256;  // MDSystem.cpp:97

struct ForceParams {
  const uint4* jrec;     // [N] j-records: periodic -> fixed-point {ux,uy,uz,0}; open -> float4 bits {x,y,z,w}
  const float4* posf;    // [N] float positions (always valid; RDF slow path + open boxes)
  float4* fpart;         // [nsplit][ilocal_cap] partial forces (fx,fy,fz, sum_j(r^-12 - r^-6))
  double* blockW;        // [gridDim.y * gridDim.x] per-CTA sum of u = r.f_ij/4 (virial, un-prefactored)
  unsigned long long* rdf;  // [256] global RDF counters (RDF variants only)
  int N;                 // total particles (j range is always 0..N)
  int i_begin, i_end;    // this rank's i-shard [i_begin, i_end)
  int ilocal_cap;        // row stride of fpart
  int tile_j;            // j's per smem tile
  float c2;              // periodic: (2^32/L)^2 : k-units r^-2 -> real r^-2
  float fscale;          // periodic: 4*L/2^32 ; open: 4
  float cut_fast;        // RDF prefilter on the fast r^2 (k-units when periodic), with margin
  // exact-RDF constants (MDSystem.cpp:273-285)
  double L;              // box edge, double as in the reference
  float thr1, thr2;      // smallest float d with fast_round((float)(d/L)) >= 1, >= 2 (host bisection)
  float dr2;             // rdf_dr2
  float inv_dr2;         // 1/dr2 (float, approximate: fixed up exactly in-kernel)
};

// Exact bin of the reference CPU path: floor((double)r2 / (double)dr2), MDSystem.cpp:282.
// IEEE double division of a 24-bit by a 24-bit significand cannot round across an integer
// (DESIGN.md §RDF), so this equals the exact rational floor, which is what is computed:
// approximate quotient, then sign tests of single-rounded FMA residuals (exact in sign).
__device__ __forceinline__ int rdf_bin_exact(float r2, float dr2, float inv_dr2) {
  float b = floorf(r2 * inv_dr2);
  // b*dr2 - r2 > 0  -> b too large
  if (__fmaf_rn(b, dr2, -r2) > 0.f) b -= 1.f;
  // (b+1)*dr2 - r2 <= 0 -> b too small
  else if (__fmaf_rn(b + 1.f, dr2, -r2) <= 0.f) b += 1.f;
  return (b < 1.0e9f) ? (int)b : 0x7fffffff;
}

// Reference image of one component: (float)(d - L * fast_round((float)(d / L))), MDSystem.cpp:274-276,
// 732-739.  fast_round((float)(d/L)) is monotone and odd in d, so |n| = 0 / 1 come from one host-bisected
// float threshold; anything farther out (never the case for positions wrapped once per step) takes the
// literal double-division route.  L*n and the subtraction are separate double roundings, as on the CPU.
__device__ __forceinline__ float image_exact(float d, const ForceParams& p) {
  const float a = fabsf(d);
  if (a < p.thr1) return d;
  double nn = 1.;
  if (a >= p.thr2) {
    const float qf = (float)__ddiv_rn((double)a, p.L);
    nn = (double)(int)__fadd_rn(qf, 0.5f);
  }
  if (d < 0.f) nn = -nn;
  return (float)__dadd_rn((double)d, -__dmul_rn(p.L, nn));
}

// ---- RDF: per-warp compaction queue -----------------------------------------------------------------
// In-range pairs are rare (543 rho* of N neighbours), so evaluating the exact sequence inside the pair loop
// leaves most lanes idle.  Instead each warp queues (i, j) index pairs in shared memory and drains them 32 at
// a time with every lane busy.  Bit 31 of the i index marks an unordered pair of the Newton-3 kernel, which
// stands for (i,j) and (j,i): both fall in the same bin (the reference sequence is odd in the separation).
constexpr int kRdfQueueCap = 288;   // < 32 waiting + at most 64 pushed by one call, 4 calls between two drain checks
                                    // of the Newton-3 kernel's RDF loop (ljmd_force_sym.cuh)
struct RdfCtx {
  uint2* q;            // this warp's queue
  unsigned int* hist;  // this warp's 256-bin histogram
  int n;               // queue length (warp-uniform)
  // where the drain finds the float positions of a queued pair: entry (i - ioff, j - joff) reads pa[.x], pb[.y].
  // k_force: the global array, offsets 0.  k_force_sym: shared-memory copies of the tile in registers and of the
  // unit's j-records (a drain is two dependent loads per pair in the middle of the pair loop: from L2 they stall
  // the warp for most of a microsecond, and dense units drain every few steps).
  const float4* pa;
  const float4* pb;
  unsigned ioff, joff;
};

template <bool PERIODIC>
__device__ __forceinline__ void rdf_exact_one(const uint2 e, const ForceParams& p, const RdfCtx& R) {
  unsigned int* hist = R.hist;
  const float4 a = R.pa[e.x & 0x7fffffffu], b = R.pb[e.y];
  float rx = __fsub_rn(a.x, b.x), ry = __fsub_rn(a.y, b.y), rz = __fsub_rn(a.z, b.z);   // MDSystem.cpp:269-271
  if (PERIODIC) { rx = image_exact(rx, p); ry = image_exact(ry, p); rz = image_exact(rz, p); }
  // MDSystem.cpp:279 in float, un-fused, left to right
  const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
  const int bin = rdf_bin_exact(r2, p.dr2, p.inv_dr2);
  if ((unsigned)bin < (unsigned)kRdfBins) atomicAdd(&hist[bin], 1u + (e.x >> 31));
}

// drain whole groups of 32 (all == false) or everything (all == true)
template <bool PERIODIC>
__device__ __forceinline__ void rdf_drain(RdfCtx& R, const ForceParams& p, bool all) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  while (R.n >= 32 || (all && R.n > 0)) {
    const int cnt = min(32, R.n);
    if (lane < cnt) rdf_exact_one<PERIODIC>(R.q[R.n - cnt + lane], p, R);
    R.n -= cnt;
  }
  __syncwarp();
}

// c_lo / c_hi: this lane's lo / hi pair is inside the histogram range (by the fast r^2, with margin).
// DRAIN = false: the caller checks the queue itself, once per several calls (every inlined drain is ~150
// instructions: eight copies in an unrolled loop body overflowed the instruction cache, ncu `no_instruction` 0.98
// stalled warps per issue).
template <bool PERIODIC, bool DRAIN = true>
__device__ __forceinline__ void rdf_push(RdfCtx& R, bool c_lo, bool c_hi, unsigned i_lo, unsigned i_hi, unsigned j,
                                         const ForceParams& p) {
  const unsigned m_lo = __ballot_sync(0xffffffffu, c_lo), m_hi = __ballot_sync(0xffffffffu, c_hi);
  if ((m_lo | m_hi) == 0u) return;   // warp-uniform: the common case costs two votes and a branch
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  if (c_lo) R.q[R.n + __popc(m_lo & lt)] = make_uint2(i_lo, j);
  R.n += __popc(m_lo);
  if (c_hi) R.q[R.n + __popc(m_hi & lt)] = make_uint2(i_hi, j);
  R.n += __popc(m_hi);
  if (DRAIN && R.n >= 32) rdf_drain<PERIODIC>(R, p, false);
}

// One pair of i-particles (two lanes of V).
template <typename V>
struct PairAcc {
  V fx, fy, fz;  // force accumulators (k-units when periodic)
  V s6, w;       // tile-level sum r^-6, sum u
};
template <typename V>
struct PairI {
  int ax, ay, az, bx, by, bz;  // fixed-point coordinates of the two particles (periodic)
  V x2, y2, z2;                // float coordinates (open boxes)
  unsigned i_lo, i_hi;         // global indices (RDF queue entries)
  bool v_lo, v_hi;             // lane holds a real particle (not a clamped duplicate)
};

// ORDERED evaluation of (i_lo, j) and (i_hi, j): force, potential and virial land on the i side only.
// DIAG: the tile overlaps this CTA's own particles, self_* flags exclude the self pair.
template <typename V, bool PERIODIC, bool DIAG, bool RDF>
__device__ __forceinline__ void pair_body(const uint4& uj, const PairI<V>& pi, PairAcc<V>& acc, bool self_lo,
                                          bool self_hi, const ForceParams& p, unsigned jglobal, RdfCtx& R) {
  V dx, dy, dz;
  if (PERIODIC) {
    // the wrap of the 32-bit subtract IS the minimum image
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
  if (DIAG) {
    if (self_lo) inv0 = 0.f;
    if (self_hi) inv1 = 0.f;
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
  if (RDF) {
    // clamped duplicate lanes (v_* false) and the self pair never count
    rdf_push<PERIODIC>(R, r2s.x < p.cut_fast && pi.v_lo && !(DIAG && self_lo),
                       r2s.y < p.cut_fast && pi.v_hi && !(DIAG && self_hi), pi.i_lo - R.ioff, pi.i_hi - R.ioff,
                       jglobal - R.joff, p);
  }
}

template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < THREADS / 32; ++k) t += red[k];
  }
  __syncthreads();
  return t;  // valid in thread 0
}

// Load the thread's 2*NPAIR i-particles: particle m is ibase + m*THREADS + tid, pair q = (2q, 2q+1).
template <typename V, bool PERIODIC, int THREADS, int NPAIR>
__device__ __forceinline__ bool load_i_particles(const ForceParams& p, int ibase, PairI<V> (&pi)[NPAIR]) {
  bool all_valid = true;
#pragma unroll
  for (int q = 0; q < NPAIR; ++q) {
    int i0 = ibase + (2 * q) * THREADS + (int)threadIdx.x, i1 = i0 + THREADS;
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

// Shared tail of both force kernels: scale and store the direct partial forces + per-particle potential,
// reduce the virial sum to one double per CTA, merge the per-warp RDF histograms.
template <typename V, bool PERIODIC, bool RDF, int THREADS, int NPAIR>
__device__ __forceinline__ void force_epilogue(const ForceParams& p, int ibase, const PairI<V> (&pi)[NPAIR],
                                               const V (&fxrun)[NPAIR], const V (&fyrun)[NPAIR],
                                               const V (&fzrun)[NPAIR], const V (&s6run)[NPAIR],
                                               const V (&wrun)[NPAIR], double* red, unsigned int* hist, RdfCtx& R) {
  const int tid = threadIdx.x;
  const float fs = p.fscale;
  double wsum = 0.;
  float4* out = p.fpart + (size_t)blockIdx.y * p.ilocal_cap;
#pragma unroll
  for (int q = 0; q < NPAIR; ++q) {
    const float2 fx = upk(fxrun[q]), fy = upk(fyrun[q]), fz = upk(fzrun[q]);
    const float2 s6 = upk(s6run[q]), w = upk(wrun[q]);
    const int il = (ibase - p.i_begin) + (2 * q) * THREADS + tid;
    // r^-12 - r^-6 = u/12 - r^-6/2
    if (pi[q].v_lo) {
      out[il] = make_float4(fx.x * fs, fy.x * fs, fz.x * fs, w.x * (1.f / 12.f) - 0.5f * s6.x);
      wsum += (double)w.x;
    }
    if (pi[q].v_hi) {
      out[il + THREADS] = make_float4(fx.y * fs, fy.y * fs, fz.y * fs, w.y * (1.f / 12.f) - 0.5f * s6.y);
      wsum += (double)w.y;
    }
  }
  const double wtot = block_sum<THREADS>(wsum, red);
  if (tid == 0) p.blockW[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = wtot;
  if (RDF) {
    rdf_drain<PERIODIC>(R, p, true);
    __syncthreads();
    for (int b = tid; b < kRdfBins; b += THREADS) {
      unsigned int c = 0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) c += hist[w * kRdfBins + b];
      if (c) atomicAdd(&p.rdf[b], (unsigned long long)c);
    }
  }
}

// RDF scratch per CTA: per-warp histograms + per-warp queues
inline size_t rdf_smem_bytes(int threads) {
  return (size_t)(threads / 32) * (kRdfBins * 4 + kRdfQueueCap * 8);
}

// grid: (i-tiles, j-splits).  block: THREADS.  dyn smem: force_smem_bytes().
// UNROLL: j-loop unroll.  (A software-pipelined main loop — stage A of batch b+1 beside stage B of batch b —
// was measured 2-10 % slower and removed: profiles/r01_tune_force_65536_pipelined.log.)
template <typename V, bool PERIODIC, bool RDF, int THREADS, int MINB, int NPAIR, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINB) k_force(const ForceParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int IPT = 2 * NPAIR;
  constexpr int NW = THREADS / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int TJ = p.tile_j;
  uint4* tile_u = reinterpret_cast<uint4*>(smem_raw);  // [2][TJ]
  unsigned char* tail = smem_raw + (size_t)2 * TJ * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // [2]
  double* red = reinterpret_cast<double*>(tail + 16);  // [NW]
  unsigned int* hist = reinterpret_cast<unsigned int*>(tail + 16 + 8 * NW);         // [NW][256]   (RDF)
  uint2* queues = reinterpret_cast<uint2*>(tail + 16 + 8 * NW + NW * kRdfBins * 4);  // [NW][cap]   (RDF)

  const int ibase = p.i_begin + blockIdx.x * (THREADS * IPT);
  // j-split: near-equal contiguous chunks
  const int ns = gridDim.y;
  const int jb = (int)(((long long)p.N * blockIdx.y) / ns);
  const int je = (int)(((long long)p.N * (blockIdx.y + 1)) / ns);
  const int ntiles = (je - jb + TJ - 1) / TJ;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (RDF) {
    for (int k = tid; k < NW * kRdfBins; k += THREADS) hist[k] = 0u;
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int j0 = jb + t * TJ;
    const int nj = min(TJ, je - j0);
    const int st = t & 1;
    const uint32_t bytes = (uint32_t)nj * 16u;
    mbar_expect_tx(&bars[st], bytes);
    bulk_g2s(tile_u + (size_t)st * TJ, p.jrec + j0, bytes, &bars[st]);
  };
  if (tid == 0 && ntiles > 0) issue(0);

  PairI<V> pi[NPAIR];
  PairAcc<V> acc[NPAIR];
  // run-level sums: every accumulator is folded per tile (two-level float summation), which keeps the
  // rounding error of an N-term float sum at ~sqrt(tile)+sqrt(N/tile) ulps instead of sqrt(N)
  V s6run[NPAIR], wrun[NPAIR], fxrun[NPAIR], fyrun[NPAIR], fzrun[NPAIR];
  const V zero2 = bc2<V>(0.f);
  load_i_particles<V, PERIODIC, THREADS, NPAIR>(p, ibase, pi);
#pragma unroll
  for (int q = 0; q < NPAIR; ++q) {
    acc[q].fx = acc[q].fy = acc[q].fz = acc[q].s6 = acc[q].w = zero2;
    s6run[q] = wrun[q] = fxrun[q] = fyrun[q] = fzrun[q] = zero2;
  }
  RdfCtx R;
  R.q = queues + (tid >> 5) * kRdfQueueCap;
  R.hist = hist + (tid >> 5) * kRdfBins;
  R.n = 0;
  R.pa = R.pb = p.posf;
  R.ioff = R.joff = 0u;

  for (int t = 0; t < ntiles; ++t) {
    if (tid == 0 && t + 1 < ntiles) issue(t + 1);
    const int st = t & 1;
    mbar_wait(&bars[st], (uint32_t)((t >> 1) & 1));
    const int j0 = jb + t * TJ;
    const int nj = min(TJ, je - j0);
    const uint4* tu = tile_u + (size_t)st * TJ;
    // does this tile contain any of this CTA's own particles?
    const bool diag = (j0 < ibase + THREADS * IPT) && (j0 + nj > ibase);
    if (!diag) {
#pragma unroll UNROLL
      for (int j = 0; j < nj; ++j) {
        const uint4 uj = tu[j];
#pragma unroll
        for (int q = 0; q < NPAIR; ++q)
          pair_body<V, PERIODIC, false, RDF>(uj, pi[q], acc[q], false, false, p, (unsigned)(j0 + j), R);
      }
    } else {
      const int jrel0 = j0 - ibase - tid;  // j-index relative to my particle m = 0
#pragma unroll 2
      for (int j = 0; j < nj; ++j) {
        const uint4 uj = tu[j];
        const int jr = jrel0 + j;
#pragma unroll
        for (int q = 0; q < NPAIR; ++q)
          pair_body<V, PERIODIC, true, RDF>(uj, pi[q], acc[q], jr == (2 * q) * THREADS, jr == (2 * q + 1) * THREADS,
                                            p, (unsigned)(j0 + j), R);
      }
    }
#pragma unroll
    for (int q = 0; q < NPAIR; ++q) {
      s6run[q] = add2(s6run[q], acc[q].s6);
      wrun[q] = add2(wrun[q], acc[q].w);
      fxrun[q] = add2(fxrun[q], acc[q].fx);
      fyrun[q] = add2(fyrun[q], acc[q].fy);
      fzrun[q] = add2(fzrun[q], acc[q].fz);
      acc[q].s6 = acc[q].w = acc[q].fx = acc[q].fy = acc[q].fz = zero2;
    }
    __syncthreads();  // everyone is done with stage st before it is refilled
  }
  force_epilogue<V, PERIODIC, RDF, THREADS, NPAIR>(p, ibase, pi, fxrun, fyrun, fzrun, s6run, wrun, red, hist, R);
}

inline size_t force_smem_bytes(bool rdf, int tile_j, int threads) {
  return (size_t)2 * tile_j * 16 + 16 + 8 * (threads / 32) + (rdf ? rdf_smem_bytes(threads) : 0);
}

}  // namespace ljmd
