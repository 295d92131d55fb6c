// O(N) kernels of the MD step: drift, gather (+EVN finish), TVN finish, parameters,
// velocity histogram, fixed-point conversion.  All HBM-bound; float4 AoS, one thread per
// particle, coalesced 16-byte accesses.
//
// Arithmetic follows /root/reference/src/library/MDSystem.cpp literally where the reference
// evaluates in double and stores to float (drift :447-449/:473-475, kicks :451-453,:460-462,
// TVN :477-505, boundaries :414-435, parameters :325-359), using explicit round-to-nearest
// intrinsics so nvcc cannot contract a*b+c into an FMA the CPU build (-ffp-contract=off)
// does not have.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ljmd {

constexpr int kStepThreads = 256;

enum { SUM_PE = 0, SUM_W = 1, SUM_TV2 = 2, SUM_K = 3, SUM_COUNT = 4 };

// Device-resident scalar block (mirrors the public members of MDSystem, MDSystem.h:63-66,80,98-99).
struct DevScalars {
  double U, T, K, V, P, Pvirial, t;
  double av_U_tot, av_T_tot, av_p_tot;
  long long av_iters;
  double chi, Tkin_trial;
  double sums[SUM_COUNT];  // raw sums of the current step (local, then all-reduced in place)
  int fabric_timeout;      // set when a peer never arrived at a fabric barrier
  int pad_;
};

// ---- intra-node fabric: peers' windows mapped through CUDA IPC (NVLink / NVSwitch peer memory) ----------
// Each rank owns one window allocation [posA | upos | rsum | slots | flags]; base[r] is rank r's window as
// seen from this device (base[me] is the local one).  n == 0: fabric off.
constexpr int kMaxPeers = 8;
struct Fabric {
  char* base[kMaxPeers];
  int n, me;
  unsigned long long off_posA, off_upos, off_rsum, off_slots, off_flags;
};
constexpr int kSlotDoubles = 8;   // per (parity, source rank): up to 8 doubles

struct StepParams {
  int N;            // all particles
  int nloc;         // particles of this rank
  int i_begin;      // first global index of this rank
  int bc;           // 0 periodic, 1 hard wall, 2 none
  int nsplit;       // rows of fpart
  int ilocal_cap;   // row stride of fpart
  int nforce_blocks;  // entries of blockW
  int world;        // number of ranks
  int gather_shift; // k_gather: 2^gather_shift adjacent lanes share one particle's rows (0: one thread each)
  double dt, dt2;   // dt, dt*dt
  double L;         // box edge
  double rho;       // N / rho is the volume the reference divides by (MDSystem.cpp:350)
  double T0;
  double fix_scale;  // 2^32 / L
  float4* pos;      // [nloc]   wrapped positions (what h_Pos shows after Integrate)
  float4* posA;     // [Npad]   positions at force-evaluation time (all ranks' shards)
  uint4* upos;      // [Npad]   fixed-point box fractions of posA (periodic)
  float4* vel;      // [nloc]
  float4* force;    // [nloc]   f(t) (xyz), w = per-particle sum_j(r^-12 - r^-6)
  float4* tforce;   // [nloc]   TVN t_Force
  const float4* fpart;   // [nsplit][ilocal_cap]
  const double* blockW;  // [nforce_blocks]
  double* part;     // [2][gridDim.x] per-block partial sums
  unsigned int* counter;  // last-block ticket
  DevScalars* sc;
  // Newton-3 kernel: reaction rows (ljmd_force_sym.cuh)
  int use_sym;      // reaction rows are in use
  int nblk;         // global number of 512-particle blocks
  int blk0;         // global index of this rank's first block
  int n_itiles;     // i-tiles (rows) of this rank
  int rp_stride;    // records per rpart block = sym_mju * sym_bj
  int sym_bj;       // j-records per unit of the Newton-3 kernel
  int sym_mi;       // i-tiles per super-tile
  int sym_mju;      // units per window
  int sym_nwin;     // windows per super-tile
  int n_super;      // super-tiles of this rank
  int npad;         // padded particle count (world * shard capacity)
  const float4* rpart;    // [n_super][sym_nwin][rp_stride]
  float4* rsum;           // [npad] rank-local column sums (world > 1)
  const float4* rshard;   // [nloc] reaction totals after the reduce-scatter (world > 1, NCCL path)
  Fabric fab;             // peer windows (world > 1, fabric path)
  // Record order (periodic Newton-3 runs): slot[il] is the global RECORD index of local particle il — where its
  // evaluation position lives in posA / upos and which row entries of fpart / rpart / rsum are its.  The state
  // arrays (pos, vel, force, tforce) stay in the caller's particle order; only the records are kept sorted along
  // a Hilbert curve so that the force kernel's warps and chunks are compact clouds (ljmd_force_sym.cuh, "Warp
  // frames").  A permutation of [i_begin, i_begin + nloc); nullptr = identity.  ANY permutation is correct: the
  // records are rewritten through it before every evaluation, so a stale sort only costs speed.
  const int* slot;
  // the inverse: order[k] is the local particle whose record is i_begin + k (nullptr = identity).  k_gather runs
  // record-major — its ~100 row and reaction reads per particle stay coalesced, only the handful of state-array
  // accesses scatter — while the drift and finish kernels, which touch one record each, run particle-major.
  const int* order;
};
__device__ __forceinline__ int record_of(const StepParams& p, int il) { return p.slot ? p.slot[il] : p.i_begin + il; }

constexpr int kBlockParticles = 512;   // = kITile of ljmd_core.cu / B of k_force_sym

__host__ __device__ inline int partner_count(int g, int n) {
  if (n & 1) return (n - 1) / 2;
  return n / 2 - 1 + (g < n / 2 ? 1 : 0);
}

// Sum of the reaction blocks that hold contributions for global particle j (block J = j / 512).  Super-tile a of
// this rank (first global block I0 = blk0 + a * mi) accumulated the reaction on J in band unit
// u = ((J - I0) mod n) * cpb + jj / bj when u lies inside its band, q = (J - I0) mod n <= qmax
// (ljmd_force_sym.cuh); every entry of a band is written on every launch (zeros included).  With D = (J - blk0)
// mod n the covering super-tiles are two runs of t: t * mi <= D with q = D - t * mi <= qmax, and t * mi > D with
// q = D + n - t * mi <= qmax.  Both are walked in ascending t (fixed order: deterministic; the same order and the
// same bits as a predicated walk over every t), four independent loads in flight: with the coverage test inside
// the loop every load sat behind its own branch and its add behind the load — 32 serial L2/DRAM latencies per
// thread, a third of k_gather's stall samples at N = 65 536 (profiles/r02_gather_kernel_ncu_C3.md).
// A super-tile's windows lie back to back, rp_stride = mju * bj records each, so band unit u of super-tile t
// starts at record (t * nwin * mju + u) * bj whatever window it is in: no division in the loop.
__device__ __forceinline__ float4 reaction_sum(const StepParams& p, int j, int first = 0, int stride = 1) {
  const int J = j / kBlockParticles, jj = j - J * kBlockParticles;
  const int n = p.nblk, mi = p.sym_mi;
  const int hmax = (n & 1) ? (n - 1) / 2 : n / 2;
  const int cpb = kBlockParticles / p.sym_bj;
  int qmax = mi - 1 + hmax;
  if (qmax > n - 1) qmax = n - 1;
  const int c = jj / p.sym_bj, jr = jj - c * p.sym_bj;
  const size_t band = (size_t)p.sym_nwin * p.sym_mju;
  const float4* base = p.rpart + (size_t)c * p.sym_bj + jr;
  int D = J - p.blk0;
  if (D < 0) D += n;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  // t in [lo, hi] with t = first (mod stride) — this lane's share when several lanes split a particle (stride is a
  // power of two); q = Dq - t * mi
  auto walk = [&](int lo, int hi, int Dq) {
    // (a short last batch loads under a predicate and adds zeros: with 8 lanes per particle a lane's whole share
    // is one partial batch, which must not fall back to one load at a time)
    for (int t = lo + ((first - lo) & (stride - 1)); t <= hi; t += 4 * stride) {
      float4 g[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        // load unconditionally from a clamped (valid) address, select afterwards: a guarded load makes ptxas
        // put every load behind a branch again
        const int tt = t + m * stride, tc = min(tt, hi);
        const float4 v = base[((size_t)tc * band + (size_t)((Dq - tc * mi) * cpb)) * p.sym_bj];
        const bool on = tt <= hi;
        g[m] = make_float4(on ? v.x : 0.f, on ? v.y : 0.f, on ? v.z : 0.f, 0.f);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) { a.x += g[m].x; a.y += g[m].y; a.z += g[m].z; }
    }
  };
  const int tD = D / mi;
  walk(D > qmax ? (D - qmax + mi - 1) / mi : 0, min(p.n_super - 1, tD), D);
  walk(max(tD + 1, (D + n - qmax + mi - 1) / mi), p.n_super - 1, D + n);
  return a;
}

// world > 1: column sums over this rank's rows for every particle (input of the reduce-scatter)
__global__ void __launch_bounds__(256) k_reduce_reaction(const StepParams p) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= p.npad) return;
  p.rsum[j] = (j < p.N) ? reaction_sum(p, j) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ uint32_t to_fixed(float x, double fix_scale) {
  // box fraction in 32-bit fixed point; the cast to 32 bits is the periodic wrap
  return (uint32_t)(unsigned long long)__double2ll_rn(__dmul_rn((double)x, fix_scale));
}

// Publish one particle's evaluation position (and fixed-point record).  Fabric on: the store goes straight
// into every rank's window over NVLink — the all-gather is fused into the producing kernel.
__device__ __forceinline__ void publish_position(const StepParams& p, int ig, const float4& x) {
  const uint4 u = make_uint4(to_fixed(x.x, p.fix_scale), to_fixed(x.y, p.fix_scale), to_fixed(x.z, p.fix_scale), 0u);
  if (p.fab.n > 0) {
    for (int r = 0; r < p.fab.n; ++r) {
      reinterpret_cast<float4*>(p.fab.base[r] + p.fab.off_posA)[ig] = x;
      if (p.bc == 0) reinterpret_cast<uint4*>(p.fab.base[r] + p.fab.off_upos)[ig] = u;
    }
  } else {
    p.posA[ig] = x;
    if (p.bc == 0) p.upos[ig] = u;
  }
}

// pos += dt*v + dt*dt*f/2  in double, stored to float (MDSystem.cpp:447-449, :473-475)
__device__ __forceinline__ float drift1(float x, float v, float f, double dt, double dt2) {
  const double a = __dmul_rn(dt, (double)v);
  const double b = __dmul_rn(__dmul_rn(dt2, (double)f), 0.5);
  return (float)__dadd_rn((double)x, __dadd_rn(a, b));
}
// v += dt*f/2 (MDSystem.cpp:451-453, :460-462)
__device__ __forceinline__ float kick1(float v, float f, double dt) {
  return (float)__dadd_rn((double)v, __dmul_rn(__dmul_rn(dt, (double)f), 0.5));
}
// float expression vx*vx+vy*vy+vz*vz, un-fused, left to right (MDSystem.cpp:334,367,660)
__device__ __forceinline__ float sq3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// MDSystem::ApplyBoundaryConditions for one particle (MDSystem.cpp:406-436).
__device__ __forceinline__ void apply_bc(float4& p, float4& v, double L, int bc) {
  if (bc == 0) {
    if ((double)p.x < 0.) p.x = (float)__dadd_rn((double)p.x, L);
    if ((double)p.x > L) p.x = (float)__dadd_rn((double)p.x, -L);
    if ((double)p.y < 0.) p.y = (float)__dadd_rn((double)p.y, L);
    if ((double)p.y > L) p.y = (float)__dadd_rn((double)p.y, -L);
    if ((double)p.z < 0.) p.z = (float)__dadd_rn((double)p.z, L);
    if ((double)p.z > L) p.z = (float)__dadd_rn((double)p.z, -L);
  } else if (bc == 1) {
    if ((double)p.x < 0. && v.x < 0.f) v.x = -v.x;
    if ((double)p.x > L && v.x > 0.f) v.x = -v.x;
    if ((double)p.y < 0. && v.y < 0.f) v.y = -v.y;
    if ((double)p.y > L && v.y > 0.f) v.y = -v.y;
    if ((double)p.z < 0. && v.z < 0.f) v.z = -v.z;
    if ((double)p.z > L && v.z > 0.f) v.z = -v.z;
  }
}

// MDSystem::CalculateParameters (MDSystem.cpp:348-358) + t += dt (:582) from the reduced sums.
// accumulate: 1 inside Integrate, 0 for a bare evaluation (set_state resets av_* anyway).
__device__ __forceinline__ void finalize_values(DevScalars* sc, int N, double rho, double dt, int accumulate) {
  const double K = sc->sums[SUM_K];
  const double V = sc->sums[SUM_PE] * (4. / 2.);            // :308
  const double Pvir = sc->sums[SUM_W] * (4. / 3. / 2.);     // :307
  const double T = 2. * K / 3. / N;                          // :348
  double P = Pvir + N * T;                                   // :349
  P /= (N / rho);                                            // :350
  sc->K = K; sc->V = V; sc->T = T; sc->P = P; sc->Pvirial = Pvir;
  sc->U = K + V;                                             // :351
  if (accumulate) {
    sc->av_iters += 1;                                       // :355-358
    sc->av_U_tot += sc->U;
    sc->av_p_tot += P;
    sc->av_T_tot += T;
    sc->t += dt;                                             // :582
  }
}
__device__ __forceinline__ void finalize_params(const StepParams& p, int accumulate) {
  finalize_values(p.sc, p.N, p.rho, p.dt, accumulate);
}

template <bool CANON>
__global__ void __launch_bounds__(kStepThreads) k_drift(const StepParams p) {
  pdl_trigger();
  pdl_wait();
  const int il = blockIdx.x * kStepThreads + threadIdx.x;
  if (il >= p.nloc) return;
  float4 x = p.pos[il];
  float4 v = p.vel[il];
  const float4 f = p.force[il];
  x.x = drift1(x.x, v.x, f.x, p.dt, p.dt2);
  x.y = drift1(x.y, v.y, f.y, p.dt, p.dt2);
  x.z = drift1(x.z, v.z, f.z, p.dt, p.dt2);
  publish_position(p, record_of(p, il), x);
  if (!CANON) {
    v.x = kick1(v.x, f.x, p.dt);
    v.y = kick1(v.y, f.y, p.dt);
    v.z = kick1(v.z, f.z, p.dt);
    p.vel[il] = v;
  }
}

// posA/upos from pos for a freshly uploaded snapshot (local shard -> global slots).
__global__ void __launch_bounds__(kStepThreads) k_prepare(const StepParams p) {
  const int il = blockIdx.x * kStepThreads + threadIdx.x;
  if (il >= p.nloc) return;
  publish_position(p, record_of(p, il), p.pos[il]);
}

// sum_k a[tid + k * kStepThreads] in ascending k, loads batched eight deep (L2: ld.global.cg)
__device__ __forceinline__ double strided_sum_l2(const double* a, int n) {
  double t = 0.;
  int k = threadIdx.x;
  for (; k + 7 * kStepThreads < n; k += 8 * kStepThreads) {
    double v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = __ldcg(a + k + m * kStepThreads);
#pragma unroll
    for (int m = 0; m < 8; ++m) t += v[m];
  }
  for (; k < n; k += kStepThreads) t += __ldcg(a + k);
  return t;
}

// Deterministic two-value block reduction + last-block final sum over all blocks.
// Returns true in thread 0 of the last block after sc->sums[ia], sc->sums[ib] are written.
__device__ __forceinline__ bool reduce_two(double a, double b, const StepParams& p, int ia, int ib,
                                           bool also_force_blocks) {
  __shared__ double sa[kStepThreads / 32], sb[kStepThreads / 32];
  __shared__ bool is_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0., tb = 0.;
#pragma unroll
    for (int k = 0; k < kStepThreads / 32; ++k) { ta += sa[k]; tb += sb[k]; }
    p.part[blockIdx.x] = ta;
    p.part[gridDim.x + blockIdx.x] = tb;
    __threadfence();
    const unsigned int ticket = atomicAdd(p.counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // last block: fixed-order sums (deterministic run to run)
  // (the partials were written by other CTAs of this launch / by the force kernel: L2 loads, never L1.  Eight
  // loads go out before the first add: a strided walk of up to 16 640 force-block sums with one dependent
  // load per add cost the last block 0.6 us per trip, 20 us of a 73 us k_gather at N = 65 536; the order of the
  // additions — and with it every bit of the sums — is what it was.)
  double ta = 0., tb = 0., tw = 0.;
  ta = strided_sum_l2(p.part, (int)gridDim.x);
  tb = strided_sum_l2(p.part + gridDim.x, (int)gridDim.x);
  if (also_force_blocks) tw = strided_sum_l2(p.blockW, p.nforce_blocks);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ta += __shfl_xor_sync(0xffffffffu, ta, o);
    tb += __shfl_xor_sync(0xffffffffu, tb, o);
    tw += __shfl_xor_sync(0xffffffffu, tw, o);
  }
  __shared__ double sw[kStepThreads / 32];
  __syncthreads();
  if (l == 0) { sa[w] = ta; sb[w] = tb; sw[w] = tw; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ta = tb = tw = 0.;
#pragma unroll
    for (int k = 0; k < kStepThreads / 32; ++k) { ta += sa[k]; tb += sb[k]; tw += sw[k]; }
    p.sc->sums[ia] = ta;
    p.sc->sums[ib] = tb;
    if (also_force_blocks) p.sc->sums[SUM_W] = tw;
    *p.counter = 0u;
    return true;
  }
  return false;
}

enum { GATHER_EVAL = 0, GATHER_EVN = 1, GATHER_TVN = 2 };

// Kick-drift-wrap fusion: when another step follows, the kernel that finishes step n (final velocities,
// boundary wrap, kinetic energy) goes straight on with the drift of step n+1 for its particle — the same
// arithmetic k_drift would do on the values it just produced — so a batched ljmd_step runs one O(N) kernel
// per step beside the force kernel instead of two.  `canon`: step n+1 is TVN (no first half-kick).
__device__ __forceinline__ void fused_next_drift(const StepParams& p, int rec, float4 x, float4& v, const float4& f,
                                                 bool canon) {
  x.x = drift1(x.x, v.x, f.x, p.dt, p.dt2);
  x.y = drift1(x.y, v.y, f.y, p.dt, p.dt2);
  x.z = drift1(x.z, v.z, f.z, p.dt, p.dt2);
  publish_position(p, rec, x);
  if (!canon) {
    v.x = kick1(v.x, f.x, p.dt);
    v.y = kick1(v.y, f.y, p.dt);
    v.z = kick1(v.z, f.z, p.dt);
  }
}

// Sum the j-split partial forces; then, per mode,
//  EVAL: K from the current velocities (a bare CalculateForces + CalculateParameters)
//  EVN : second half-kick, boundary conditions, K            (MDSystem.cpp:458-463,578-579)
//  TVN : t_Force, t_Vel and sum t_Vel^2                       (MDSystem.cpp:484-498)
// FUSE (EVN only here; TVN fuses in k_finish_tvn): also perform the drift of the next step.
template <int MODE, bool FUSE = false>
__global__ void __launch_bounds__(kStepThreads) k_gather(const StepParams p, int finalize, int accumulate) {
  pdl_trigger();
  pdl_wait();
  // R = 2^gather_shift adjacent lanes share a particle: each sums every R-th row of partial forces (and of
  // reaction rows), a butterfly over the R lanes adds the shares (every lane ends with the same bits), lane 0 of
  // the group goes on.  Mid-size systems have too few particles to hide the latency of S = 30-130 dependent-
  // address row loads with one thread each (N = 65 536: 76 -> 46 us; N = 1 500: the TVN step 23.8 -> 15.7 us).
  const int R = 1 << p.gather_shift;
  const int gt = blockIdx.x * kStepThreads + threadIdx.x;
  const int rl = gt >> p.gather_shift, r = gt & (R - 1);   // rl: my record (local index): rows and reaction entries
  const int rec = p.i_begin + rl;
  int il = rl;                                             // il: the particle it belongs to (state arrays)
  double pe = 0., q = 0.;
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f), rr = f;
  // the state-array reads of the tail (lane 0 of a group) go out first: two dependent latencies (order -> vel,
  // force, posA) that would otherwise follow the row sums
  float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), fo0 = v0, x0 = v0;
  if (rl < p.nloc) {
    if (p.order) il = p.order[rl];
    if (r == 0) {
      v0 = p.vel[il];
      if (MODE == GATHER_TVN) fo0 = p.force[il];
      if (MODE == GATHER_EVN) x0 = p.posA[rec];
    }
    // rows are independent loads (up to ~130 rows with fine splits): eight in flight, added in row order
    {
      const float4* row = p.fpart + rl;
      const size_t cap = (size_t)p.ilocal_cap;
      for (int s = r; s < p.nsplit; s += 8 * R) {   // (a short last batch: predicated loads, zeros added)
        float4 g[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int sm = s + m * R;
          g[m] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (sm < p.nsplit) g[m] = row[(size_t)sm * cap];
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) { f.x += g[m].x; f.y += g[m].y; f.z += g[m].z; f.w += g[m].w; }
      }
    }
    // Newton-3 kernel, one GPU: the reaction of every pair this particle was the j of
    if (p.use_sym && p.world == 1) rr = reaction_sum(p, rec, r, R);
  }
  for (int o = 1; o < R; o <<= 1) {
    f.x += __shfl_xor_sync(0xffffffffu, f.x, o); f.y += __shfl_xor_sync(0xffffffffu, f.y, o);
    f.z += __shfl_xor_sync(0xffffffffu, f.z, o); f.w += __shfl_xor_sync(0xffffffffu, f.w, o);
    rr.x += __shfl_xor_sync(0xffffffffu, rr.x, o); rr.y += __shfl_xor_sync(0xffffffffu, rr.y, o);
    rr.z += __shfl_xor_sync(0xffffffffu, rr.z, o);
  }
  if (rl < p.nloc && r == 0) {
    if (p.use_sym) {
      if (p.world > 1) {
        if (p.fab.n > 0) {
          // pull every rank's column sum for my particle straight from its window, fixed rank order
          rr = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = 0; q < p.fab.n; ++q) {
            const float4 g = reinterpret_cast<const float4*>(p.fab.base[q] + p.fab.off_rsum)[rec];
            rr.x += g.x; rr.y += g.y; rr.z += g.z;
          }
        } else {
          rr = p.rshard[rl];
        }
      }
      f.x += rr.x; f.y += rr.y; f.z += rr.z;
    }
    pe = (double)f.w;
    float4 v = v0;
    if (MODE == GATHER_EVAL) {
      p.force[il] = f;
      q = (double)sq3(v.x, v.y, v.z) * 0.5;                 // :334
    } else if (MODE == GATHER_EVN) {
      p.force[il] = f;
      v.x = kick1(v.x, f.x, p.dt);
      v.y = kick1(v.y, f.y, p.dt);
      v.z = kick1(v.z, f.z, p.dt);
      float4 x = x0;
      apply_bc(x, v, p.L, p.bc);
      p.pos[il] = x;
      q = (double)sq3(v.x, v.y, v.z) * 0.5;
      if (FUSE) fused_next_drift(p, rec, x, v, f, false);
      p.vel[il] = v;
    } else {
      const float4 fo = fo0;
      p.force[il] = f;
      float4 tf;
      tf.x = __fadd_rn(__fmul_rn(0.5f, fo.x), __fmul_rn(0.5f, f.x));   // :477-479,486-488
      tf.y = __fadd_rn(__fmul_rn(0.5f, fo.y), __fmul_rn(0.5f, f.y));
      tf.z = __fadd_rn(__fmul_rn(0.5f, fo.z), __fmul_rn(0.5f, f.z));
      tf.w = 0.f;
      p.tforce[il] = tf;
      const float tx = kick1(v.x, tf.x, p.dt);                           // :493-495
      const float ty = kick1(v.y, tf.y, p.dt);
      const float tz = kick1(v.z, tf.z, p.dt);
      q = (double)sq3(tx, ty, tz);                                       // :367
    }
  }
  const bool last = reduce_two(pe, q, p, SUM_PE, (MODE == GATHER_TVN) ? SUM_TV2 : SUM_K, true);
  if (last && finalize && MODE != GATHER_TVN) finalize_params(p, accumulate);
}

// TVN velocity update with chi = sqrt(T0 / Tkin(t_Vel)), boundaries, K  (MDSystem.cpp:498-506,578-579)
template <bool FUSE = false>
__global__ void __launch_bounds__(kStepThreads) k_finish_tvn(const StepParams p, int finalize) {
  pdl_trigger();
  pdl_wait();
  const int il = blockIdx.x * kStepThreads + threadIdx.x;
  double Tkin = p.sc->sums[SUM_TV2];
  Tkin *= 1. / 3. / p.N;                                     // :370
  const double chi = sqrt(p.T0 / Tkin);                      // :499
  double q = 0.;
  if (il < p.nloc) {
    float4 v = p.vel[il];
    const float4 tf = p.tforce[il];
    const double a = 2. * chi - 1.;
    const double b = chi * p.dt;
    v.x = (float)__dadd_rn(__dmul_rn(a, (double)v.x), __dmul_rn(b, (double)tf.x));   // :503-505
    v.y = (float)__dadd_rn(__dmul_rn(a, (double)v.y), __dmul_rn(b, (double)tf.y));
    v.z = (float)__dadd_rn(__dmul_rn(a, (double)v.z), __dmul_rn(b, (double)tf.z));
    const int rec = record_of(p, il);
    float4 x = p.posA[rec];
    apply_bc(x, v, p.L, p.bc);
    p.pos[il] = x;
    q = (double)sq3(v.x, v.y, v.z) * 0.5;
    if (FUSE) fused_next_drift(p, rec, x, v, p.force[il], true);
    p.vel[il] = v;
  }
  if (il == 0) { p.sc->chi = chi; p.sc->Tkin_trial = Tkin; }
  const bool last = reduce_two(q, 0., p, SUM_K, SUM_TV2, false);
  if (last && finalize) finalize_params(p, 1);
}

// K only (after a velocity upload).
__global__ void __launch_bounds__(kStepThreads) k_kinetic(const StepParams p) {
  const int il = blockIdx.x * kStepThreads + threadIdx.x;
  double q = 0.;
  if (il < p.nloc) {
    const float4 v = p.vel[il];
    q = (double)sq3(v.x, v.y, v.z) * 0.5;
  }
  reduce_two(q, 0., p, SUM_K, SUM_TV2, false);
}

// Multi-rank: parameters from the all-reduced sums.
__global__ void k_params(const StepParams p, int accumulate) {
  if (threadIdx.x == 0 && blockIdx.x == 0) finalize_params(p, accumulate);
}

// Speed histogram, shared-memory privatised (MDSystem.cpp:651-662,676-688).
// bin = (int)(sqrt((double)(float)(vx^2+vy^2+vz^2)) / step); IEEE double sqrt and divide are
// correctly rounded on the device, so bins are bit-exact with the CPU path.
__global__ void __launch_bounds__(kStepThreads) k_velhist(const float4* __restrict__ vel, int n, double step,
                                                           int nbins, unsigned int* __restrict__ out) {
  extern __shared__ unsigned int h[];
  for (int k = threadIdx.x; k < nbins; k += kStepThreads) h[k] = 0u;
  __syncthreads();
  for (int i = blockIdx.x * kStepThreads + threadIdx.x; i < n; i += gridDim.x * kStepThreads) {
    const float4 v = vel[i];
    const double s = sqrt((double)sq3(v.x, v.y, v.z));
    const double b = __ddiv_rn(s, step);
    if (b < (double)nbins) atomicAdd(&h[(int)b], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nbins; k += kStepThreads)
    if (h[k]) atomicAdd(&out[k], h[k]);
}

// Fabric barrier with an optional all-reduce of sc->sums[first .. first+count): every rank stores its partial
// sums into its slot of every peer's window, raises its flag there to `epoch`, waits until all its own flags
// reached `epoch`, and adds the slots up in rank order (bit-identical totals on every rank).  Slots are
// double-buffered on the epoch parity: a fast rank can be one sync ahead of a slow one, never two.
// Release / acquire at system scope (PTX memory model): the flag store publishes every earlier store of this
// rank's kernels to the peers (.release orders them before the flag becomes visible), the flag load makes
// the peers' earlier stores visible to this rank (.acquire orders the following loads after it).
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kFabricTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;   // 10 s of wall time

// fin: 0 barrier / all-reduce only; 1 also CalculateParameters from the reduced sums (the step's last barrier:
// saves the separate k_params launch); 2 the same with the av_* accumulation and t += dt of an Integrate.
__global__ void k_fabric_sync(const Fabric f, unsigned long long epoch, int first, int count, DevScalars* sc, int fin,
                              int N, double rho, double dt) {
  const int lane = threadIdx.x;
  const int par = (int)(epoch & 1ull);
  if (lane < f.n) {
    if (count > 0) {
      double* dst = reinterpret_cast<double*>(f.base[lane] + f.off_slots) + ((size_t)par * kMaxPeers + f.me) * kSlotDoubles;
      for (int k = 0; k < count; ++k) dst[k] = sc->sums[first + k];
    }
    // everything this rank's earlier kernels in the stream stored into peer windows (positions, reaction sums)
    // and the slot above become visible to rank `lane` before it sees the flag
    st_release_sys(reinterpret_cast<unsigned long long*>(f.base[lane] + f.off_flags) + f.me, epoch);
    // wait for rank `lane` to arrive: acquire load, exponential back-off, wall-clock watchdog
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(f.base[f.me] + f.off_flags) + lane;
    const unsigned long long t0 = global_timer_ns();
    unsigned int ns = 32;
    while (ld_acquire_sys(mine) < epoch) {
      __nanosleep(ns);
      if (ns < 1024) ns <<= 1;
      if (global_timer_ns() - t0 > kFabricTimeoutNs) {   // a peer died: fail the step instead of hanging the GPU
        sc->fabric_timeout = 1;
        break;
      }
    }
  }
  __syncwarp();
  if (count > 0 && lane == 0) {
    // lane 0 reads slots written by peers whose flags OTHER lanes acquired: order those loads after the
    // warp-level rendezvous with a system-scope fence
    __threadfence_system();
    const double* src = reinterpret_cast<const double*>(f.base[f.me] + f.off_slots) + (size_t)par * kMaxPeers * kSlotDoubles;
    for (int k = 0; k < count; ++k) {
      double t = 0.;
      for (int r = 0; r < f.n; ++r) t += __ldcv(src + (size_t)r * kSlotDoubles + k);
      sc->sums[first + k] = t;
    }
    if (fin) finalize_values(sc, N, rho, dt, fin == 2);
  }
  // the kernels that follow in the stream read peer-written data with plain loads: make this kernel's acquires
  // cover them (kernel boundary orders; the fence makes the visibility explicit for every lane)
  __threadfence_system();
}

// Sub-volume occupancy histograms for the fluctuation tasks
// (src/tasks/run-fluctuations/include/run-fluctuations-aux.h:188-278 GetNSubsystemBatch / GetNsubVzBatch):
//  type 0..2  slab along x/y/z: bin = (int)(((double)coord / L) / alpha_step)
//  type 3     cube about the centre: bin = max over axes of upper_bound(tLs, |coord - L/2| / 0.5)
//  type 4..6  |v_x|,|v_y|,|v_z|:   bin = (int)((|(double)v| / vcut_max) / alpha_step)
// Particles whose bin is >= nbins are dropped; the host turns the histogram into cumulative counts.
// IEEE double divisions on the device are correctly rounded, so the bins equal the CPU's bit for bit.
constexpr int kMaxSubBins = 128;
struct SubvolSpec {
  int type, nbins;     // 0-2 slab in x/y/z, 3 centred cube, 4-6 |vx|,|vy|,|vz|
  double L, alpha_step, vcut_max;
  double tLs[kMaxSubBins];   // type 3: L * alpha_k^(1/3), ascending
};
struct SubvolParams {
  const float4* arr;   // pos (types 0-3) or vel (types 4-6), local shard
  int n;
  SubvolSpec s;
};
// Bin of one particle: the literal arithmetic of GetNSubsystemBatch / GetNsubVzBatch
// (run-fluctuations-aux.h:188-278); q.nbins = "counted nowhere".
__device__ __forceinline__ int subvol_bin(const SubvolSpec& q, const float4 a) {
  int bin;
  if (q.type <= 2) {
    double c = (q.type == 0) ? (double)a.x : (q.type == 1) ? (double)a.y : (double)a.z;
    c = __ddiv_rn(c, q.L);
    const double b = __ddiv_rn(c, q.alpha_step);
    bin = (b < (double)q.nbins) ? (int)b : q.nbins;   // (int) truncates toward zero like the CPU cast
  } else if (q.type == 3) {
    bin = 0;
    const double half = __dmul_rn(0.5, q.L);
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const double c = (ax == 0) ? (double)a.x : (ax == 1) ? (double)a.y : (double)a.z;
      const double v = __ddiv_rn(fabs(__dadd_rn(c, -half)), 0.5);
      int lo = 0, hi = q.nbins;                        // upper_bound: first tLs[k] > v
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (q.tLs[mid] > v) hi = mid; else lo = mid + 1;
      }
      bin = max(bin, lo);
    }
  } else {
    double v = (q.type == 4) ? (double)a.x : (q.type == 5) ? (double)a.y : (double)a.z;
    v = __ddiv_rn(v, q.vcut_max);
    const double b = __ddiv_rn(fabs(v), q.alpha_step);
    bin = (b < (double)q.nbins) ? (int)b : q.nbins;
  }
  return bin;
}
__global__ void __launch_bounds__(kStepThreads) k_subvolume(const SubvolParams q, unsigned int* __restrict__ out) {
  __shared__ unsigned int h[kMaxSubBins];
  for (int k = threadIdx.x; k < q.s.nbins; k += kStepThreads) h[k] = 0u;
  __syncthreads();
  for (int i = blockIdx.x * kStepThreads + threadIdx.x; i < q.n; i += gridDim.x * kStepThreads) {
    const int bin = subvol_bin(q.s, q.arr[i]);
    if (bin >= 0 && bin < q.s.nbins) atomicAdd(&h[bin], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < q.s.nbins; k += kStepThreads)
    if (h[k]) atomicAdd(&out[k], h[k]);
}

// ---- observation trace: what the fluctuation tasks read after every step, recorded on the device ----------
// One row per step: the bins of every configured counter (per-bin, not yet cumulative), then the three velocity
// sums in 2^-32 fixed point (integers: order-free, identical for any grid and any number of ranks), and the
// step's scalars {t, U, T, P, K, V, Pvirial}.  Replaces the per-step D2H of h_Pos / h_Vel in
// run-fluctuations.cpp:124-135,177-180 and the per-step reads of U, P in run-isotherm.cpp:104-106.
constexpr int kMaxTraceCounters = 8;
constexpr int kTraceScalars = 8;
struct TraceParams {
  const float4* pos;
  const float4* vel;
  int n;                       // particles of this rank
  int ncounters, row;          // row = sum of nbins + 3
  const SubvolSpec* specs;     // [ncounters], device memory
  const DevScalars* sc;
  unsigned long long* counts;  // [capacity][row]
  double* scal;                // [capacity][kTraceScalars]
  int* row_idx;                // device-resident index of the row this launch fills: the launch arguments are the
  unsigned int* ticket;        // same for every step, so a traced batch can be replayed from a CUDA graph
};
__global__ void __launch_bounds__(kStepThreads) k_trace(const TraceParams q) {
  // dynamic smem: the counter specifications (the cube counter binary-searches its table of edges: 21 dependent
  // loads per particle, 13 us per launch at N = 400 when they came from global memory), then row - 3 bins
  extern __shared__ __align__(8) unsigned char trace_smem[];
  SubvolSpec* sspec = reinterpret_cast<SubvolSpec*>(trace_smem);
  unsigned int* th = reinterpret_cast<unsigned int*>(trace_smem + (size_t)q.ncounters * sizeof(SubvolSpec));
  const int nb = q.row - 3;
  const int r = *q.row_idx;              // every CTA reads it before the last one to finish advances it
  unsigned long long* counts = q.counts + (size_t)r * q.row;
  {
    const int nw = q.ncounters * (int)(sizeof(SubvolSpec) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(q.specs);
    int* dst = reinterpret_cast<int*>(sspec);
    for (int k = threadIdx.x; k < nw; k += kStepThreads) dst[k] = src[k];
  }
  for (int k = threadIdx.x; k < nb; k += kStepThreads) th[k] = 0u;
  __syncthreads();
  long long sx = 0, sy = 0, sz = 0;
  for (int i = blockIdx.x * kStepThreads + threadIdx.x; i < q.n; i += gridDim.x * kStepThreads) {
    const float4 x = q.pos[i], v = q.vel[i];
    int off = 0;
    for (int c = 0; c < q.ncounters; ++c) {
      const SubvolSpec& sp = sspec[c];
      const int bin = subvol_bin(sp, sp.type <= 3 ? x : v);
      if (bin >= 0 && bin < sp.nbins) atomicAdd(&th[off + bin], 1u);
      off += sp.nbins;
    }
    sx += __double2ll_rn((double)v.x * 4294967296.0);
    sy += __double2ll_rn((double)v.y * 4294967296.0);
    sz += __double2ll_rn((double)v.z * 4294967296.0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counts[nb + 0], (unsigned long long)sx);
    atomicAdd(&counts[nb + 1], (unsigned long long)sy);
    atomicAdd(&counts[nb + 2], (unsigned long long)sz);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nb; k += kStepThreads)
    if (th[k]) atomicAdd(&counts[k], (unsigned long long)th[k]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double* scal = q.scal + (size_t)r * kTraceScalars;
    scal[0] = q.sc->t; scal[1] = q.sc->U; scal[2] = q.sc->T; scal[3] = q.sc->P;
    scal[4] = q.sc->K; scal[5] = q.sc->V; scal[6] = q.sc->Pvirial; scal[7] = 0.;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(q.ticket, 1u) == gridDim.x - 1) {   // last CTA: the row is complete
      *q.ticket = 0u;
      *q.row_idx = r + 1;
    }
  }
}

// ---- device-side seeded initial conditions (MDSystem::SampleInitialConditions, MDSystem.cpp:147-181)
// Positions: the reference's simple-cubic start lattice, same double arithmetic rounded to float.  Velocities: the
// reference draws a Maxwell speed and an isotropic direction from a time-seeded Mersenne twister (not
// reproducible); here three N(0, T) components per particle from Philox4x32-10 keyed by the seed and COUNTED by the
// global particle index, so a particle gets the same draw on any grid shape and any number of GPUs.  The total-
// momentum correction and the rescale to T0 (:179-180) need two global sums; they are accumulated as 2^-32
// fixed-point integers (associative: bit-identical for any launch shape or GPU count).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct InitParams {
  float4* pos;        // [nloc]
  float4* vel;        // [nloc]
  int nloc, i_begin, N;
  int Ns;             // lattice sites per edge = ceil(N^(1/3)) (:150)
  double dL;          // L / Ns (:151)
  double L, T0;
  unsigned long long seed;
  long long* sums;    // [4]: sum vx, vy, vz (2^-32 fixed point), then sum v^2 (2^-32 fixed point)
  double mean[3];     // phase 2: total momentum / N
  double factor;      // phase 3: sqrt(T0 / Tkin)
};

// phase 1: lattice + raw velocities + momentum sums
__global__ void __launch_bounds__(kStepThreads) k_init_sample(const InitParams q) {
  const int il = blockIdx.x * kStepThreads + threadIdx.x;
  long long sx = 0, sy = 0, sz = 0;
  if (il < q.nloc) {
    const int iN = q.i_begin + il;
    const int Ns = q.Ns;                                                    // :150, evaluated on the host (libm's pow)
    const double dL = q.dL;                                                 // :151
    float4 x;
    x.x = (float)__dmul_rn(__dadd_rn((double)(iN % Ns), 0.5), dL);          // :161-167
    x.y = (float)__dmul_rn(__dadd_rn((double)((iN / Ns) % Ns), 0.5), dL);
    x.z = (float)__dmul_rn(__dadd_rn((double)(iN / (Ns * Ns)), 0.5), dL);
    x.w = (float)__ddiv_rn(q.L, 150.);                                      // :168 `L / 150.f`: a double division
    q.pos[il] = x;
    uint32_t r[4];
    philox4x32_10((uint32_t)iN, 0u, 0u, 0u, (uint32_t)q.seed, (uint32_t)(q.seed >> 32), r);
    // Box-Muller on uniforms in (0, 1): three of the four normals
    const double u1 = ((double)r[0] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)r[1] + 0.5) * (1.0 / 4294967296.0);
    const double u3 = ((double)r[2] + 0.5) * (1.0 / 4294967296.0), u4 = ((double)r[3] + 0.5) * (1.0 / 4294967296.0);
    const double ra = sqrt(-2. * log(u1)), rb = sqrt(-2. * log(u3));
    const double sT = sqrt(q.T0);
    float4 v;
    v.x = (float)(sT * ra * cospi(2. * u2));
    v.y = (float)(sT * ra * sinpi(2. * u2));
    v.z = (float)(sT * rb * cospi(2. * u4));
    v.w = 0.f;
    q.vel[il] = v;
    sx = __double2ll_rn((double)v.x * 4294967296.0);
    sy = __double2ll_rn((double)v.y * 4294967296.0);
    sz = __double2ll_rn((double)v.z * 4294967296.0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(q.sums), (unsigned long long)sx);
    atomicAdd(reinterpret_cast<unsigned long long*>(q.sums) + 1, (unsigned long long)sy);
    atomicAdd(reinterpret_cast<unsigned long long*>(q.sums) + 2, (unsigned long long)sz);
  }
}
// phase 2: CorrectTotalMomentum (:183-216) + sum of v^2
__global__ void __launch_bounds__(kStepThreads) k_init_center(const InitParams q) {
  const int il = blockIdx.x * kStepThreads + threadIdx.x;
  long long s2 = 0;
  if (il < q.nloc) {
    float4 v = q.vel[il];
    v.x = (float)__dadd_rn((double)v.x, -q.mean[0]);     // float += double: formed in double, rounded once (:199-201)
    v.y = (float)__dadd_rn((double)v.y, -q.mean[1]);
    v.z = (float)__dadd_rn((double)v.z, -q.mean[2]);
    q.vel[il] = v;
    s2 = __double2ll_rn((double)sq3(v.x, v.y, v.z) * 4294967296.0);   // :367 float expression
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(q.sums) + 3, (unsigned long long)s2);
}
// phase 3: RenormalizeVelocities(true) (:375-389)
__global__ void __launch_bounds__(kStepThreads) k_init_scale(const InitParams q) {
  const int il = blockIdx.x * kStepThreads + threadIdx.x;
  if (il >= q.nloc) return;
  float4 v = q.vel[il];
  v.x = (float)__dmul_rn((double)v.x, q.factor);
  v.y = (float)__dmul_rn((double)v.y, q.factor);
  v.z = (float)__dmul_rn((double)v.z, q.factor);
  q.vel[il] = v;
}

__global__ void k_rdf_accum(const unsigned long long* cur, unsigned long long* acc) {
  acc[threadIdx.x] += cur[threadIdx.x];
}

}  // namespace ljmd
