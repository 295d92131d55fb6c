// Developer tool (not part of the product library): A/B timings of force-kernel variants and a few
// pipe-throughput microbenchmarks on the B200.  Build: make -C tools.  Run on the GPU box:
//   tools/tune_force [N] [reps]
// Prints one line per variant: ms per launch, ordered pairs/s, and the algorithmic TFLOP/s
// (37 flop/pair periodic, 25 open).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../lennard-jones-cuda_b200/csrc/ljmd_force_sym.cuh"
#include "../lennard-jones-cuda_b200/csrc/ljmd_hilbert.cuh"
using namespace ljmd;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

// ------------------------------------------------------------------ pipe microbenchmarks
// 8 independent chains per thread; `iters` loop trips; results kept live through a final store.
template <int MODE>
__global__ void __launch_bounds__(256) k_micro(float* out, int iters, float seed) {
  float a[8], b[8];
  int ia[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = seed + k + threadIdx.x; b[k] = seed * 0.5f + k; ia[k] = (int)threadIdx.x * 7 + k; }
  const float c = seed * 1.0001f, d = seed * 0.9999f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) {            // FFMA (3 distinct source registers)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(c), "f"(d));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[k]) : "f"(c), "f"(d));
      } else if (MODE == 1) {     // FFMA2: a[k],b[k] as one packed pair
        unsigned long long v, cc, dd;
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(a[k]), "f"(b[k]));
        asm volatile("mov.b64 %0, {%1,%1};" : "=l"(cc) : "f"(c));
        asm volatile("mov.b64 %0, {%1,%1};" : "=l"(dd) : "f"(d));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(cc), "l"(dd));
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a[k]), "=f"(b[k]) : "l"(v));
      } else if (MODE == 2) {     // I2FP + IADD (the conversion needs a changing integer)
        asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[k]) : "r"(it));
        float t;
        asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(ia[k]));
        a[k] += t;                // FADD to keep the result live
      } else if (MODE == 3) {     // MUFU.RCP
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
      } else if (MODE == 4) {     // IADD3 only
        asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[k]) : "r"(it));
        asm volatile("sub.s32 %0, %0, %1;" : "+r"(ia[k]) : "r"(k));
      } else if (MODE == 5) {     // LJ-like mix per 2 lanes: 7 FFMA2-class + 3*2 IADD + 3*2 I2FP + 2 MUFU (approx.)
        unsigned long long v, cc;
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(a[k]), "f"(b[k]));
        asm volatile("mov.b64 %0, {%1,%1};" : "=l"(cc) : "f"(c));
#pragma unroll
        for (int r = 0; r < 7; ++r) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(cc));
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a[k]), "=f"(b[k]) : "l"(v));
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          asm volatile("sub.s32 %0, %0, %1;" : "+r"(ia[k]) : "r"(it + r));
          float t;
          asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(ia[k]));
          if (r & 1) a[k] += t; else b[k] += t;
        }
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b[k]));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k] + b[k] + (float)ia[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run_micro(const char* name, double ops_per_iter_per_thread, int sms, float* d_out) {
  const int iters = 4096, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_micro<MODE><<<blocks, threads>>>(d_out, 64, 1.0f);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_micro<MODE><<<blocks, threads>>>(d_out, iters, 1.0f);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double ops = ops_per_iter_per_thread * iters * (double)blocks * threads;
  printf("micro %-34s %8.3f ms  %8.2f Gop/s/SM  (= %6.1f thread-ops/clk/SM at 1.9 GHz)\n", name, ms,
         ops / (ms * 1e-3) / sms / 1e9, ops / (ms * 1e-3) / sms / 1.9e9);
}

// ------------------------------------------------------------------ force-kernel variants
struct Problem {
  int N;
  double L;
  uint4* upos;
  float4* posf;
  float4* fpart;
  double* blockW;
  unsigned long long* rdf;
  float4* rpart;   // [n][hmax*B] reaction rows for the symmetric kernel (B = 512)
  uint4* bbox;     // [n][2] block bounding boxes
  int hmax;
  int sms;
  size_t fpart_elems, rpart_elems;   // capacity of the scratch buffers, in float4 records
};

static int pick_split(int n_it, int N, int sms, int minb) {
  const double ovh = 128.;
  const long long slots = (long long)sms * minb;
  int best = 1;
  double bc = 1e300;
  for (int s = 1; s <= std::max(1, N / 64) && s <= 8 * sms; ++s) {
    const long long waves = ((long long)n_it * s + slots - 1) / slots;
    const double cost = (double)waves * ((double)N / s + ovh);
    if (cost < bc * 0.999) { bc = cost; best = s; }
  }
  return best;
}

static int g_force_split = 0;   // scan_ordered: overrides the split of run_variant
static int g_shard = 1;         // emulate rank 0 of a g_shard-way sharded run: i-particles [0, N / g_shard)
template <typename V, bool PERIODIC, bool RDF, int THREADS, int MINB, int NPAIR, int UNROLL>
static void run_variant(const Problem& pb, const char* tag, int reps, int tile_j, std::vector<float4>* keep) {
  auto kern = k_force<V, PERIODIC, RDF, THREADS, MINB, NPAIR, UNROLL>;
  const size_t smem = force_smem_bytes(RDF, tile_j, THREADS);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  const int itile = THREADS * 2 * NPAIR;
  const int n_it = (pb.N + itile - 1) / itile;
  const int S = g_force_split > 0 ? g_force_split : pick_split(n_it, pb.N, pb.sms, occ > 0 ? occ : MINB);
  ForceParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.jrec = PERIODIC ? pb.upos : reinterpret_cast<const uint4*>(pb.posf);
  fp.posf = pb.posf; fp.fpart = pb.fpart; fp.blockW = pb.blockW; fp.rdf = pb.rdf;
  fp.N = pb.N; fp.i_begin = 0; fp.i_end = pb.N; fp.ilocal_cap = pb.N; fp.tile_j = tile_j;
  const double k2 = 4294967296.0 / pb.L;
  fp.c2 = PERIODIC ? (float)(k2 * k2) : 1.f;
  fp.fscale = PERIODIC ? (float)(4.0 * pb.L / 4294967296.0) : 4.f;
  fp.cut_fast = (float)(PERIODIC ? 25.6 * 1.001 * k2 * k2 : 25.6 * 1.001);
  fp.L = pb.L; fp.thr1 = (float)(0.5 * pb.L); fp.thr2 = (float)(1.5 * pb.L); fp.dr2 = 0.1f; fp.inv_dr2 = 10.f;
  dim3 grid(n_it, S);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  kern<<<grid, THREADS, smem>>>(fp);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float best = 1e30f, sum = 0.f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    kern<<<grid, THREADS, smem>>>(fp);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
    sum += ms;
  }
  const double pairs = (double)pb.N * (pb.N - 1);
  const double flop = PERIODIC ? 37. : 25.;
  // checksum of the summed partial forces of particle 12345 against the first variant run
  std::vector<float4> h((size_t)S * pb.N);
  CK(cudaMemcpy(h.data(), pb.fpart, h.size() * 16, cudaMemcpyDeviceToHost));
  std::vector<float4> f(pb.N);
  for (int i = 0; i < pb.N; ++i) {
    float4 a = h[i];
    for (int s = 1; s < S; ++s) { float4 g = h[(size_t)s * pb.N + i]; a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w; }
    f[i] = a;
  }
  double maxdiff = 0., scale = 0.;
  if (keep->empty()) *keep = f;
  for (int i = 0; i < pb.N; ++i) {
    maxdiff = std::max(maxdiff, (double)fabsf(f[i].x - (*keep)[i].x));
    scale = std::max(scale, (double)fabsf((*keep)[i].x));
  }
  printf("force %-44s regs %3d occ %d grid %4dx%-3d smem %6zu | best %8.4f ms avg %8.4f ms | %7.3f Gpairs/s %6.2f TFLOP/s "
         "| maxdiff %.2e/%.2e\n",
         tag, fa.numRegs, occ, n_it, S, smem, best, sum / reps, pairs / (best * 1e-3) / 1e9,
         pairs * flop / (best * 1e-3) / 1e12, maxdiff, scale);
  fflush(stdout);
}

// mju: units per window, mi: i-tiles per super-tile (0: pick like the library's planner for this N)
static int g_win_shift = -1;    // >= 0: override of SymParams::win_shift
static int g_frames = 1;        // FRAMES kernels: SymParams::frames
static double g_rfar = 2.5;     // FRAMES kernels: R_far
template <typename V, bool PERIODIC, bool RDF, int THREADS, int MINB, int NPAIR, int UNROLLK = 4, bool FRAMES = false>
static void run_sym(const Problem& pb, const char* tag, int reps, int bj, std::vector<float4>* keep,
                    bool prune = true, int mju = 0, int mi = 1) {
  auto kern = k_force_sym<V, PERIODIC, RDF, THREADS, MINB, NPAIR, UNROLLK, FRAMES>;
  const int B = THREADS * 2 * NPAIR;
  const int n = (pb.N + B - 1) / B;
  const int cpb = B / bj;
  const int hmax = std::max(1, sym_max_partner_count(n));
  const int units = (hmax + 1) * cpb;
  if (mju <= 0) {   // ~10 CTAs per slot, as the library
    const long long slots = (long long)pb.sms * MINB;
    const long long total = (long long)n * units;
    int per = (int)std::max<long long>(1, (total + 10 * slots - 1) / (10 * slots));
    if (per >= 8) per = (int)std::max<long long>(8, (total + 60 * slots - 1) / (60 * slots));
    mju = std::min(per, 16);
  }
  mi = std::max(1, std::min(mi, n / 2));
  const int band = sym_band_units(mi, hmax, n, cpb);
  const int nwin = (band + mju - 1) / mju;
  const int n_loc = (n + g_shard - 1) / g_shard;   // i-tiles of the emulated rank
  const int nsup = (n_loc + mi - 1) / mi;
  const size_t smem = force_sym_smem_bytes(RDF, bj, THREADS, mju);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  SymParams sp;
  memset(&sp, 0, sizeof(sp));
  ForceParams& fp = sp.f;
  fp.jrec = PERIODIC ? pb.upos : reinterpret_cast<const uint4*>(pb.posf);
  fp.posf = pb.posf; fp.fpart = pb.fpart; fp.blockW = pb.blockW; fp.rdf = pb.rdf;
  fp.N = pb.N; fp.i_begin = 0; fp.i_end = std::min(pb.N, n_loc * B); fp.ilocal_cap = pb.N; fp.tile_j = bj;
  const double k2 = 4294967296.0 / pb.L;
  fp.c2 = PERIODIC ? (float)(k2 * k2) : 1.f;
  fp.fscale = PERIODIC ? (float)(4.0 * pb.L / 4294967296.0) : 4.f;
  fp.cut_fast = (float)(PERIODIC ? 25.6 * 1.001 * k2 * k2 : 25.6 * 1.001);
  fp.L = pb.L; fp.thr1 = (float)(0.5 * pb.L); fp.thr2 = (float)(1.5 * pb.L); fp.dr2 = 0.1f; fp.inv_dr2 = 10.f;
  sp.bbox = nullptr; sp.bbox_cut2 = 25.7f;
  if (RDF && prune) {   // block bounding boxes for the RDF pruning test
    if (PERIODIC) k_bbox<true, THREADS * 2 * NPAIR><<<n, 128>>>(fp.jrec, pb.N, pb.bbox);
    else k_bbox<false, THREADS * 2 * NPAIR><<<n, 128>>>(fp.jrec, pb.N, pb.bbox);
    CK(cudaGetLastError());
    sp.bbox = pb.bbox;
  }
  sp.rpart = pb.rpart; sp.ncols = 0; sp.nblk = n; sp.bj = bj;
  sp.mi = mi; sp.mju = mju; sp.nwin = nwin; sp.win_shift = std::min(nwin - 1, ((mi - 1) * cpb + mju - 1) / mju);
  if (g_win_shift >= 0) sp.win_shift = std::min(nwin - 1, g_win_shift);
  sp.frames = FRAMES ? g_frames : 0; sp.kunit = (float)(pb.L / 4294967296.0);
  sp.far2 = (float)((g_rfar * 4294967296.0 / pb.L) * (g_rfar * 4294967296.0 / pb.L));
  sp.rdf2 = (float)((25.6 * 1.002 + 1e-3) * k2 * k2);
  const size_t rp_elems = (size_t)nsup * nwin * mju * bj;
  if ((size_t)nwin * pb.N > pb.fpart_elems || rp_elems > pb.rpart_elems) {
    printf("sym   %-44s skipped: needs %zu + %zu records of scratch\n", tag, (size_t)nwin * pb.N, rp_elems);
    return;
  }
  dim3 grid(nsup, nwin);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaMemset(pb.rpart, 0xff, rp_elems * 16));   // NaN fill: the kernel must write every entry the gather reads
  CK(cudaMemset(pb.fpart, 0xff, (size_t)nwin * pb.N * 16));
  CK(cudaMemset(pb.rdf, 0, 256 * 8));
  kern<<<grid, THREADS, smem>>>(sp);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  unsigned long long hist[256], hsum = 0, hmix = 0;   // the histogram of the first launch: total and a hash
  CK(cudaMemcpy(hist, pb.rdf, sizeof(hist), cudaMemcpyDeviceToHost));
  for (int b = 0; b < 256; ++b) { hsum += hist[b]; hmix = hmix * 1000003ull + hist[b]; }
  float best = 1e30f, sum = 0.f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    kern<<<grid, THREADS, smem>>>(sp);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
    sum += ms;
  }
  const double pairs = (double)pb.N * (pb.N - 1) / g_shard;
  const double flop = PERIODIC ? 37. : 25.;
  double maxdiff = 0., scale = 0., maxw = 0.;
  if (pb.N <= 262144 && g_shard == 1) {   // host-side gather of the partial rows and reaction blocks (the library's index rule)
    std::vector<float4> h((size_t)nwin * pb.N), hr(rp_elems);
    CK(cudaMemcpy(h.data(), pb.fpart, h.size() * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), pb.rpart, hr.size() * 16, cudaMemcpyDeviceToHost));
    int qmax = mi - 1 + hmax;
    if (qmax > n - 1) qmax = n - 1;
    for (int i = 0; i < pb.N; ++i) {
      float4 a = h[i];
      for (int s = 1; s < nwin; ++s) { float4 g = h[(size_t)s * pb.N + i]; a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w; }
      const int J = i / B, jj = i % B, c = jj / bj, jr = jj % bj;
      for (int t = 0; t < nsup; ++t) {
        int q = J - t * mi;
        if (q < 0) q += n;
        if (q <= qmax) {
          const int u = q * cpb + c, w = u / mju;
          float4 g = hr[((size_t)t * nwin + w) * ((size_t)mju * bj) + (size_t)(u - w * mju) * bj + jr];
          a.x += g.x; a.y += g.y; a.z += g.z;
        }
      }
      if (!keep->empty()) {
        double d = std::max(std::max(fabs((double)a.x - (*keep)[i].x), fabs((double)a.y - (*keep)[i].y)), fabs((double)a.z - (*keep)[i].z));
        if (!(d == d)) d = 1e30;   // NaN: an entry nobody wrote
        maxdiff = std::max(maxdiff, d);
        scale = std::max(scale, (double)fabsf((*keep)[i].x));
      }
      maxw += a.w;
    }
  }
  double wref = 0.;
  for (int i = 0; i < pb.N && !keep->empty(); ++i) wref += (*keep)[i].w;
  printf("sym   %-44s regs %3d occ %d grid %4dx%-3d mi %2d mju %2d smem %6zu | best %8.4f ms avg %8.4f ms | %7.3f Gpairs/s %6.2f TFLOP/s(alg) "
         "| out %.3f GB | maxdiff %.2e/%.2e sum_pe %.6e vs %.6e",
         tag, fa.numRegs, occ, nsup, nwin, mi, mju, smem, best, sum / reps, pairs / (best * 1e-3) / 1e9,
         pairs * flop / (best * 1e-3) / 1e12, ((double)nwin * pb.N + (double)rp_elems) * 16 / 1e9, maxdiff, scale, maxw, wref);
  if (RDF) printf(" | rdf total %llu hash %016llx", hsum, hmix);
  printf("\n");
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 65536;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  const double rho = getenv("TUNE_RHO") ? atof(getenv("TUNE_RHO")) : 1.1;
  const double L = pow(N / rho, 1. / 3.);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %.0f MHz; N=%d L=%.4f\n", prop.name, sms, prop.clockRate / 1e3, N, L);


  // lattice + jitter, deterministic LCG
  std::vector<float4> hp(N);
  std::vector<uint4> hu(N);
  const int ns = (int)ceil(pow((double)N, 1. / 3.));
  const double dL = L / ns;
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)(st >> 11) / 9007199254740992.0; };
  for (int i = 0; i < N; ++i) {
    double x = ((i % ns) + 0.5 + 0.1 * (rnd() - 0.5)) * dL, y = (((i / ns) % ns) + 0.5 + 0.1 * (rnd() - 0.5)) * dL,
           z = ((i / (ns * ns)) + 0.5 + 0.1 * (rnd() - 0.5)) * dL;
    hp[i] = make_float4((float)x, (float)y, (float)z, 0.f);
    const double sc = 4294967296.0 / L;
    hu[i] = make_uint4((unsigned)(unsigned long long)llrint(hp[i].x * sc), (unsigned)(unsigned long long)llrint(hp[i].y * sc),
                       (unsigned)(unsigned long long)llrint(hp[i].z * sc), 0u);
  }
  // TUNE_ORDER=hilbert: the particles sorted along a Hilbert curve (what the library does for periodic boxes);
  // TUNE_ORDER=random: shuffled (no locality at all); default: lattice order
  if (const char* ord = getenv("TUNE_ORDER")) {
    std::vector<std::pair<unsigned long long, int>> keyed(N);
    int bits = 1;
    while ((1LL << (3 * bits)) < 2LL * N && bits < 10) ++bits;
    for (int i = 0; i < N; ++i) {
      unsigned long long k;
      if (!strcmp(ord, "hilbert")) {
        auto cell = [&](float x) { int c = (int)(x / L * (1 << bits)); return (uint32_t)std::min(std::max(c, 0), (1 << bits) - 1); };
        k = hilbert3(cell(hp[i].x), cell(hp[i].y), cell(hp[i].z), bits);
      } else {
        k = (unsigned long long)(rnd() * 1e15);
      }
      keyed[i] = {k, i};
    }
    std::sort(keyed.begin(), keyed.end());
    std::vector<float4> hp2(N);
    std::vector<uint4> hu2(N);
    for (int i = 0; i < N; ++i) { hp2[i] = hp[keyed[i].second]; hu2[i] = hu[keyed[i].second]; }
    hp.swap(hp2); hu.swap(hu2);
    printf("particle order: %s (%d bits per axis)\n", ord, bits);
  }
  Problem pb;
  pb.N = N; pb.L = L; pb.sms = sms;
  CK(cudaMalloc(&pb.upos, (size_t)N * 16));
  CK(cudaMalloc(&pb.posf, (size_t)N * 16));
  const size_t smax = std::min<size_t>(8 * sms, std::max(1, N / 8));
  pb.fpart_elems = std::max<size_t>(smax * (size_t)N, (size_t)3 << 27);   // at least 6 GB: 140 rows at N = 1M
  pb.fpart_elems = std::min<size_t>(pb.fpart_elems, (size_t)1 << 29);
  CK(cudaMalloc(&pb.fpart, pb.fpart_elems * 16));
  CK(cudaMalloc(&pb.blockW, smax * (N / 64 + 1) * sizeof(double)));
  CK(cudaMalloc(&pb.rdf, 256 * 8));
  pb.hmax = sym_max_partner_count((N + 511) / 512);
  if (pb.hmax < 1) pb.hmax = 1;
  pb.rpart_elems = std::min<size_t>((size_t)N * N / 256 + 4096, (size_t)1 << 29);
  CK(cudaMalloc(&pb.rpart, pb.rpart_elems * 16));
  CK(cudaMalloc(&pb.bbox, (size_t)(N / 256 + 2) * 2 * sizeof(uint4)));   // enough for blocks of 256 and up
  CK(cudaMemset(pb.rdf, 0, 256 * 8));
  CK(cudaMemcpy(pb.upos, hu.data(), (size_t)N * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pb.posf, hp.data(), (size_t)N * 16, cudaMemcpyHostToDevice));

  std::vector<float4> keepP, keepO;
  if (argc > 3 && !strcmp(argv[3], "shipped")) {   // only the shipped Newton-3 variants (A/B of builds)
    run_sym<P2, true, false, 128, 3, 2, 4>(pb, "periodic sym t128 b3 np2 uk4 bj256", reps, 256, &keepP);
    run_sym<P2, false, false, 128, 3, 2, 4>(pb, "open sym t128 b3 np2 uk4 bj256", reps, 256, &keepO);
    return 0;
  }
  if (argc > 3 && !strcmp(argv[3], "frames")) {
    // warp frames (periodic boxes): the shipped fixed-point kernel against the FRAMES kernel, frames on and off;
    // run it with TUNE_ORDER=hilbert|random and TUNE_RHO to see what the particle order is worth
    const int mju = argc > 4 ? atoi(argv[4]) : 0, mi = argc > 5 ? atoi(argv[5]) : 1;
    if (N <= 131072) run_variant<P2, true, false, 128, 4, 2, 4>(pb, "periodic ordered P2 t128 b4 np2 u4", reps, 1024, &keepP);
    run_sym<P2, true, false, 128, 3, 2, 4, false>(pb, "periodic sym fixed-point (shipped r02)", reps, 256, &keepP, true, mju, mi);
    g_frames = 0;
    run_sym<P2, true, false, 128, 3, 2, 4, true>(pb, "periodic sym FRAMES kernel, frames off", reps, 256, &keepP, true, mju, mi);
    g_frames = 1;
    run_sym<P2, true, false, 128, 3, 2, 4, true>(pb, "periodic sym FRAMES kernel, frames on", reps, 256, &keepP, true, mju, mi);
    g_rfar = 0.;
    run_sym<P2, true, false, 128, 3, 2, 4, true>(pb, "periodic sym FRAMES kernel, R_far = 0", reps, 256, &keepP, true, mju, mi);
    g_rfar = 2.5;
    run_sym<P2, true, true, 128, 3, 2, 4, false>(pb, "periodic+RDF sym fixed-point", reps, 256, &keepP, true, mju, mi);
    g_win_shift = 0;
    run_sym<P2, true, true, 128, 3, 2, 4, false>(pb, "periodic+RDF sym fixed-point, diagonal windows first", reps, 256, &keepP, true, mju, mi);
    g_win_shift = -1;
    run_sym<P2, true, true, 128, 3, 2, 4, true>(pb, "periodic+RDF sym FRAMES kernel, frames on", reps, 256, &keepP, true, mju, mi);
    return 0;
  }
  if (argc > 3 && !strcmp(argv[3], "scan_ordered")) {   // ordered kernel, small N: how fine should the j-split be?
    for (int per = 256; per >= 8; per >>= 1) {
      g_force_split = (N + per - 1) / per;
      if ((size_t)g_force_split > smax) continue;
      char tag[96];
      snprintf(tag, sizeof(tag), "scan_ordered j_per_cta%d S%d", per, g_force_split);
      run_variant<P2, true, false, 128, 4, 2, 4>(pb, tag, reps, 1024, &keepP);
    }
    return 0;
  }
  if (argc > 3 && !strcmp(argv[3], "scan")) {
    // every distinct "units per CTA" for three unit sizes: the data the launch planner is calibrated on
    run_variant<P2, true, false, 128, 4, 2, 4>(pb, "periodic ordered P2 t128 b4 np2 u4", reps, 1024, &keepP);
    const int n = (N + 511) / 512;
    for (int bj = 256; bj >= 64; bj >>= 1) {
      const int units = (sym_max_partner_count(n) + 1) * (512 / bj);
      int lastS = -1;
      for (int per = std::min(units, 16); per >= 1; --per) {
        const int S = (units + per - 1) / per;
        if (S == lastS || (long long)n * S > 24LL * sms * 3) continue;
        lastS = S;
        char tag[96];
        snprintf(tag, sizeof(tag), "scan bj%d S%d per_cta%d ctas%d", bj, S, per, n * S);
        run_sym<P2, true, false, 128, 3, 2, 4>(pb, tag, reps, bj, &keepP, true, per, 1);
      }
    }
    return 0;
  }
  if (argc > 4 && !strcmp(argv[3], "shardscan")) {
    // the force kernel of rank 0 of an argv[4]-way sharded run, every window size: what the planner should pick
    g_shard = atoi(argv[4]);
    for (int mju = 16; mju >= 1; --mju) {
      if (mju > 8 && (mju & 1)) continue;
      char tag[96];
      snprintf(tag, sizeof(tag), "shard%d periodic mi1 mju%d", g_shard, mju);
      run_sym<P2, true, false, 128, 3, 2, 4>(pb, tag, reps, 256, &keepP, true, mju, 1);
      snprintf(tag, sizeof(tag), "shard%d open     mi1 mju%d", g_shard, mju);
      run_sym<P2, false, false, 128, 3, 2, 4>(pb, tag, reps, 256, &keepO, true, mju, 1);
    }
    return 0;
  }
  if (argc > 3 && !strcmp(argv[3], "super")) {
    // super-tile shapes (mi x mju): kernel time and bytes of partial-force + reaction output per launch
    if (N <= 262144) run_variant<P2, true, false, 128, 4, 2, 4>(pb, "periodic ordered P2 t128 b4 np2 u4", reps, 1024, &keepP);
    const bool quick = argc > 4 && !strcmp(argv[4], "quick");
    const std::vector<int> mis = quick ? std::vector<int>{1, 4, 16} : std::vector<int>{1, 2, 4, 8, 16};
    const std::vector<int> mjus = quick ? std::vector<int>{8, 16} : std::vector<int>{4, 8, 12, 16};
    for (int mju : mjus)
      for (int mi : mis) {
        char tag[96];
        snprintf(tag, sizeof(tag), "super periodic mi%d mju%d", mi, mju);
        run_sym<P2, true, false, 128, 3, 2, 4>(pb, tag, reps, 256, &keepP, true, mju, mi);
      }
    if (N <= 262144) run_variant<P2, false, false, 128, 4, 2, 4>(pb, "open ordered P2 t128 b4 np2 u4", reps, 1024, &keepO);
    for (int mi : mis) {
      char tag[96];
      snprintf(tag, sizeof(tag), "super open mi%d mju16", mi);
      run_sym<P2, false, false, 128, 3, 2, 4>(pb, tag, reps, 256, &keepO, true, 16, mi);
    }
    run_sym<P2, true, true, 128, 3, 2, 4>(pb, "super periodic+RDF mi1 mju8", reps, 256, &keepP, true, 8, 1);
    run_sym<P2, true, true, 128, 3, 2, 4>(pb, "super periodic+RDF mi4 mju16", reps, 256, &keepP, true, 16, 4);
    return 0;
  }
  //                 V   PER    RDF   THR MINB NPAIR UNROLL
  run_variant<P2, true, false, 128, 4, 2, 4>(pb, "periodic ordered P2 t128 b4 np2 u4", reps, 1024, &keepP);
  run_variant<S2, true, false, 128, 4, 2, 4>(pb, "periodic ordered S2(scalar) t128 b4 np2 u4", reps, 1024, &keepP);
  run_sym<P2, true, false, 128, 3, 2, 4>(pb, "periodic sym t128 b3 np2 uk4 bj256", reps, 256, &keepP);
  run_sym<P2, true, false, 128, 4, 2, 4>(pb, "periodic sym t128 b4 np2 uk4 bj256", reps, 256, &keepP);
  run_sym<P2, true, false, 128, 3, 2, 4>(pb, "periodic sym t128 b3 np2 uk4 bj128", reps, 128, &keepP);
  run_sym<P2, true, false, 128, 3, 2, 4>(pb, "periodic sym t128 b3 np2 uk4 bj64", reps, 64, &keepP);
  run_sym<P2, true, false, 128, 3, 2, 4>(pb, "periodic sym t128 b3 np2 uk4 bj32", reps, 32, &keepP);
  run_variant<P2, true, true, 128, 3, 2, 4>(pb, "periodic+RDF ordered t128 b3 np2 u4", reps, 1024, &keepP);
  run_sym<P2, true, true, 128, 3, 2, 4>(pb, "periodic+RDF sym t128 b3 np2 uk4 bj256", reps, 256, &keepP);
  run_sym<P2, true, true, 128, 3, 2, 4>(pb, "periodic+RDF sym, no box pruning", reps, 256, &keepP, false);
  run_variant<P2, false, false, 128, 4, 2, 4>(pb, "open ordered P2 t128 b4 np2 u4", reps, 1024, &keepO);
  run_sym<P2, false, false, 128, 3, 2, 4>(pb, "open sym t128 b3 np2 uk4 bj256", reps, 256, &keepO);
  run_variant<P2, false, true, 128, 3, 2, 4>(pb, "open+RDF ordered t128 b3 np2 u4", reps, 1024, &keepO);
  run_sym<P2, false, true, 128, 3, 2, 4>(pb, "open+RDF sym t128 b3 np2 uk4 bj256", reps, 256, &keepO);
  run_sym<P2, false, true, 128, 3, 2, 4>(pb, "open+RDF sym, no box pruning", reps, 256, &keepO, false);
  return 0;
}
