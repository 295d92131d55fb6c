// All-pairs Lennard-Jones force / potential / virial (/ RDF) kernel for sm_100a.
//
// Replaces the reference's `calculateForces` + `RDFmerger` kernels
// (/root/reference/src/library/MDSystem.cu:21-153) and its CPU double loop
// (/root/reference/src/library/MDSystem.cpp:252-310).  Not a port: see DESIGN.md.
//
//  * i-particles live in registers, 4 per thread, as two packed f32x2 pairs so the
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
//  * RDF (on request): pairs whose fast r^2 is within the histogram range re-derive
//    r^2 with the reference CPU path's exact float/double sequence and hit a
//    per-warp shared-memory histogram; bins are bit-exact with MDSystem.cpp:269-285.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ljmd {

typedef unsigned long long u64;

// ---- packed f32x2 helpers (PTX ISA 8.6, sm_100+) ------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ float2 upk(u64 v) {
  float2 o; asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(v)); return o;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}

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
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
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
constexpr int kIPT = 4;        // i-particles per thread (two packed pairs)

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

template <bool PERIODIC>
__device__ __forceinline__ void rdf_slow(float xi, float yi, float zi, const float4& pj, const ForceParams& p,
                                         unsigned int* hist) {
  float rx = __fsub_rn(xi, pj.x), ry = __fsub_rn(yi, pj.y), rz = __fsub_rn(zi, pj.z);
  if (PERIODIC) { rx = image_exact(rx, p); ry = image_exact(ry, p); rz = image_exact(rz, p); }
  // MDSystem.cpp:279 in float, un-fused, left to right
  float r2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
  int b = rdf_bin_exact(r2, p.dr2, p.inv_dr2);
  if ((unsigned)b < (unsigned)kRdfBins) atomicAdd(&hist[b], 1u);
}

// One packed pair of i-particles against one j-record.
struct PairAcc {
  u64 fx, fy, fz;  // force accumulators (k-units when periodic)
  u64 s6, w;       // sum r^-6, sum u
};

template <bool PERIODIC, bool DIAG, bool RDF>
__device__ __forceinline__ void pair_body(const uint4& uj, const int4& ia, const int4& ib,  // fixed-point i (periodic)
                                          u64 xi2, u64 yi2, u64 zi2,                        // packed float i (open)
                                          PairAcc& acc, bool self_lo, bool self_hi, bool v_lo, bool v_hi,
                                          const ForceParams& p, const float4* pjf, unsigned int* hist) {
  u64 dx, dy, dz;
  if (PERIODIC) {
    // wrap of the 32-bit subtract IS the minimum image
    int ax = ia.x - (int)uj.x, ay = ia.y - (int)uj.y, az = ia.z - (int)uj.z;
    int bx = ib.x - (int)uj.x, by = ib.y - (int)uj.y, bz = ib.z - (int)uj.z;
    dx = pk(__int2float_rn(ax), __int2float_rn(bx));
    dy = pk(__int2float_rn(ay), __int2float_rn(by));
    dz = pk(__int2float_rn(az), __int2float_rn(bz));
  } else {
    float xj = __uint_as_float(uj.x), yj = __uint_as_float(uj.y), zj = __uint_as_float(uj.z);
    dx = sub2(xi2, pk(xj, xj));
    dy = sub2(yi2, pk(yj, yj));
    dz = sub2(zi2, pk(zj, zj));
  }
  u64 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
  float2 r2s = upk(r2);
  float inv0 = rcp_approx(r2s.x), inv1 = rcp_approx(r2s.y);
  if (DIAG) {
    if (self_lo) inv0 = 0.f;
    if (self_hi) inv1 = 0.f;
  }
  u64 x = pk(inv0, inv1);
  if (PERIODIC) x = mul2(x, pk(p.c2, p.c2));
  u64 x2 = mul2(x, x);
  u64 r6 = mul2(x2, x);
  u64 t = fma2(r6, pk(12.f, 12.f), pk(-6.f, -6.f));
  u64 u = mul2(r6, t);
  u64 s = mul2(u, x);
  acc.fx = fma2(dx, s, acc.fx);
  acc.fy = fma2(dy, s, acc.fy);
  acc.fz = fma2(dz, s, acc.fz);
  acc.s6 = add2(acc.s6, r6);
  acc.w = add2(acc.w, u);
  if (RDF) {
    float2 xs = upk(xi2), ys = upk(yi2), zs = upk(zi2);
    // clamped duplicate lanes (v_* false) and the self pair never count
    if (r2s.x < p.cut_fast && v_lo && !(DIAG && self_lo)) rdf_slow<PERIODIC>(xs.x, ys.x, zs.x, *pjf, p, hist);
    if (r2s.y < p.cut_fast && v_hi && !(DIAG && self_hi)) rdf_slow<PERIODIC>(xs.y, ys.y, zs.y, *pjf, p, hist);
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

// grid: (i-tiles, j-splits).  block: THREADS.  dyn smem: see force_smem_bytes().
template <bool PERIODIC, bool RDF, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_force(const ForceParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int TJ = p.tile_j;
  uint4* tile_u = reinterpret_cast<uint4*>(smem_raw);                          // [2][TJ]
  float4* tile_f = reinterpret_cast<float4*>(smem_raw + (size_t)2 * TJ * 16);  // [2][TJ] (RDF && PERIODIC)
  unsigned char* tail = smem_raw + (size_t)((RDF && PERIODIC) ? 4 : 2) * TJ * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);          // [2]
  double* red = reinterpret_cast<double*>(tail + 16);          // [THREADS/32]
  unsigned int* hist = reinterpret_cast<unsigned int*>(tail + 16 + 8 * (THREADS / 32));  // [THREADS/32][256] (RDF)

  const int ibase = p.i_begin + blockIdx.x * (THREADS * kIPT);
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
    for (int k = tid; k < (THREADS / 32) * kRdfBins; k += THREADS) hist[k] = 0u;
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int j0 = jb + t * TJ;
    const int nj = min(TJ, je - j0);
    const int st = t & 1;
    const uint32_t bytes = (uint32_t)nj * 16u;
    mbar_expect_tx(&bars[st], (RDF && PERIODIC) ? 2u * bytes : bytes);
    bulk_g2s(tile_u + (size_t)st * TJ, p.jrec + j0, bytes, &bars[st]);
    if (RDF && PERIODIC) bulk_g2s(tile_f + (size_t)st * TJ, p.posf + j0, bytes, &bars[st]);
  };
  if (tid == 0 && ntiles > 0) issue(0);

  // my four i's: m-th is ibase + m*THREADS + tid; pairs are (m=0,1) and (m=2,3)
  int4 ui[kIPT];
  float4 xf[kIPT];
  bool valid[kIPT];
#pragma unroll
  for (int m = 0; m < kIPT; ++m) {
    int i = ibase + m * THREADS + tid;
    valid[m] = i < p.i_end;
    if (!valid[m]) i = p.i_end - 1;
    uint4 r = p.jrec[i];
    ui[m] = make_int4((int)r.x, (int)r.y, (int)r.z, 0);
    xf[m] = (PERIODIC && !RDF) ? make_float4(0.f, 0.f, 0.f, 0.f) : p.posf[i];
  }
  u64 xA = pk(xf[0].x, xf[1].x), yA = pk(xf[0].y, xf[1].y), zA = pk(xf[0].z, xf[1].z);
  u64 xB = pk(xf[2].x, xf[3].x), yB = pk(xf[2].y, xf[3].y), zB = pk(xf[2].z, xf[3].z);

  const u64 zero2 = pk(0.f, 0.f);
  PairAcc A = {zero2, zero2, zero2, zero2, zero2}, B = {zero2, zero2, zero2, zero2, zero2};
  u64 s6A = zero2, wA = zero2, s6B = zero2, wB = zero2;  // run-level (two-level float summation)
  unsigned int* myhist = hist + (tid >> 5) * kRdfBins;

  for (int t = 0; t < ntiles; ++t) {
    if (tid == 0 && t + 1 < ntiles) issue(t + 1);
    const int st = t & 1;
    mbar_wait(&bars[st], (uint32_t)((t >> 1) & 1));
    const int j0 = jb + t * TJ;
    const int nj = min(TJ, je - j0);
    const uint4* tu = tile_u + (size_t)st * TJ;
    const float4* tf = (RDF && PERIODIC) ? (tile_f + (size_t)st * TJ) : reinterpret_cast<const float4*>(tu);
    // does this tile contain any of this CTA's own particles?
    const bool diag = (j0 < ibase + THREADS * kIPT) && (j0 + nj > ibase);
    if (!diag) {
#pragma unroll 4
      for (int j = 0; j < nj; ++j) {
        const uint4 uj = tu[j];
        pair_body<PERIODIC, false, RDF>(uj, ui[0], ui[1], xA, yA, zA, A, false, false, valid[0], valid[1], p, tf + j,
                                        myhist);
        pair_body<PERIODIC, false, RDF>(uj, ui[2], ui[3], xB, yB, zB, B, false, false, valid[2], valid[3], p, tf + j,
                                        myhist);
      }
    } else {
      const int jrel0 = j0 - ibase - tid;  // j-index relative to my m=0 particle
#pragma unroll 2
      for (int j = 0; j < nj; ++j) {
        const uint4 uj = tu[j];
        const int jr = jrel0 + j;
        pair_body<PERIODIC, true, RDF>(uj, ui[0], ui[1], xA, yA, zA, A, jr == 0, jr == THREADS, valid[0], valid[1], p,
                                       tf + j, myhist);
        pair_body<PERIODIC, true, RDF>(uj, ui[2], ui[3], xB, yB, zB, B, jr == 2 * THREADS, jr == 3 * THREADS, valid[2],
                                       valid[3], p, tf + j, myhist);
      }
    }
    // two-level summation of the scalar sums: tile-level floats folded into run-level floats
    s6A = add2(s6A, A.s6); wA = add2(wA, A.w); A.s6 = zero2; A.w = zero2;
    s6B = add2(s6B, B.s6); wB = add2(wB, B.w); B.s6 = zero2; B.w = zero2;
    __syncthreads();  // everyone is done with stage st before it is refilled
  }

  // ---- epilogue ----
  const float fs = p.fscale;
  float2 fxA = upk(A.fx), fyA = upk(A.fy), fzA = upk(A.fz), fxB = upk(B.fx), fyB = upk(B.fy), fzB = upk(B.fz);
  float2 a6 = upk(s6A), aw = upk(wA), b6 = upk(s6B), bw = upk(wB);
  const float fxs[kIPT] = {fxA.x, fxA.y, fxB.x, fxB.y};
  const float fys[kIPT] = {fyA.x, fyA.y, fyB.x, fyB.y};
  const float fzs[kIPT] = {fzA.x, fzA.y, fzB.x, fzB.y};
  const float s6s[kIPT] = {a6.x, a6.y, b6.x, b6.y};
  const float ws[kIPT] = {aw.x, aw.y, bw.x, bw.y};
  double wsum = 0.;
  float4* out = p.fpart + (size_t)blockIdx.y * p.ilocal_cap;
#pragma unroll
  for (int m = 0; m < kIPT; ++m) {
    if (valid[m]) {
      const int il = (ibase - p.i_begin) + m * THREADS + tid;
      // r^-12 - r^-6 = u/12 - r^-6/2
      const float pe = ws[m] * (1.f / 12.f) - 0.5f * s6s[m];
      out[il] = make_float4(fxs[m] * fs, fys[m] * fs, fzs[m] * fs, pe);
      wsum += (double)ws[m];
    }
  }
  const double wtot = block_sum<THREADS>(wsum, red);
  if (tid == 0) p.blockW[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = wtot;
  if (RDF) {
    __syncthreads();
    for (int b = tid; b < kRdfBins; b += THREADS) {
      unsigned int c = 0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) c += hist[w * kRdfBins + b];
      if (c) atomicAdd(&p.rdf[b], (unsigned long long)c);
    }
  }
}

inline size_t force_smem_bytes(bool periodic, bool rdf, int tile_j, int threads) {
  size_t b = (size_t)((rdf && periodic) ? 4 : 2) * tile_j * 16 + 16 + 8 * (threads / 32);
  if (rdf) b += (size_t)(threads / 32) * kRdfBins * 4;
  return b;
}

}  // namespace ljmd
