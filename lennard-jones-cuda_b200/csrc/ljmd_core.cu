// Host side of the C ABI in include/ljmd.h: handle, launch logic, collectives.
// Product code: no CPU fallback anywhere — every entry point needs a CUDA device.
#include "../../include/ljmd.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#ifdef LJMD_WITH_NCCL
#include <nccl.h>
#endif
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges around force / gather / exchange / step

#include "ljmd_force_sym.cuh"
#include "ljmd_sort.cuh"
#include "ljmd_step.cuh"

using namespace ljmd;

// ------------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char* ljmd_last_error(void) { return g_err; }

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return set_err(LJMD_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#ifdef LJMD_WITH_NCCL
#define NC(call)                                                                                  \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess)                                                                        \
      return set_err(LJMD_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
  } while (0)
#endif

// NVTX range for the enclosing scope (host-side: the launches are asynchronous, the range marks their submission;
// a timeline tool attributes the kernels launched inside it).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// --------------------------------------------------------------------------------- the handle
constexpr int kForceThreads = 128;
constexpr int kNPair = 2;                     // packed i-pairs per thread
constexpr int kUnroll = 4;                    // j-loop unroll
constexpr int kITile = kForceThreads * 2 * kNPair;  // i-particles per CTA
constexpr int kTileJ = 1024;                  // j-records per smem stage
constexpr int kMinBlocks = 4;                 // resident CTAs/SM the ordered non-RDF kernel is built for
constexpr int kSymMinBlocks = 3;              // Newton-3 kernel: 139 registers, 3 CTAs/SM measured 4 % faster than 4 (128 regs)
constexpr int kMinBlocksRdf = 3;
constexpr int kSymBJ = 256;                   // largest j-chunk (work unit) of the Newton-3 kernel; 128 or 64 for small N
constexpr int kSymMaxMJU = 16;                // units per window: 16 x 256 records x 12 B = 48 KB of shared memory, 3 CTAs/SM
constexpr int kSymMaxMI = 16;                 // i-tiles per super-tile
constexpr int kSymMinBlocksN = 8;             // Newton-3 kernel from this many 512-particle blocks on (N = 4 096: 20.2 vs 22.0 us ordered)
constexpr double kFrameRfar = 2.5;            // warp frames: a chunk takes the float path from this gap (sigma) on

struct MultiCtl;   // single-process multi-GPU front handle (end of this file)

struct ljmd_system {
  // A handle made by ljmd_create_multi is a FRONT: it owns one sub-handle per device (rank r of world G, each
  // driven by its own worker thread) and forwards every call.  Sub-handles have inproc = 1: no NCCL communicator,
  // the fabric is wired with plain peer pointers, read-outs return the rank's own contribution and the front
  // combines them, state downloads write the rank's own shard of the caller's array.
  MultiCtl* multi = nullptr;
  int inproc = 0;
  int N = 0, bc = 0, canonical = 0;
  double rho = 0., L = 0., T0 = 0.;
  float dr2 = 0.1f;
  int device = 0, rank = 0, world = 1;
  int cnt = 0;      // shard capacity = ceil(N/world)
  int i_begin = 0, i_end = 0, nloc = 0;
  int npad = 0;     // world*cnt
  int num_sms = 148;
  int nsplit = 1, n_itiles = 1;
  int nblk = 1;     // global number of kITile-blocks
  int use_sym = 0;  // Newton-3 kernel (k_force_sym) instead of the ordered one (k_force)
  int hmax = 0;     // partner offsets per i-tile
  int sym_bj = 256; // j-records per work unit of the Newton-3 kernel
  int sym_mi = 1, sym_mju = 1, sym_nwin = 1, n_super = 1, sym_win_shift = 0;   // super-tile geometry (ljmd_force_sym.cuh)
  int gather_shift = 0;  // k_gather: 2^shift lanes per particle
  int pdl = 0;           // launch the step chain with programmatic dependent launch (one GPU)
  float4* rpart = nullptr;   // [n_super][sym_nwin][sym_mju*sym_bj] reaction sums of the super-tiles' windows
  float4* rsum = nullptr;    // [npad] rank-local column sums of the reaction rows (world > 1)
  float4* rshard = nullptr;  // [cnt]  reaction totals of this rank's particles after the reduce-scatter
  uint4* bbox = nullptr;     // [nblk][2] block bounding boxes (RDF pruning in the Newton-3 kernel)
  // record order (Newton-3 kernel): slot[il] = record of local particle il; periodic boxes keep the records sorted
  // along a Hilbert curve (ljmd_sort.cuh) and run the FRAMES kernel (ljmd_force_sym.cuh, "Warp frames")
  int frames = 0;            // sorting + FRAMES kernel enabled (Newton-3 kernel; LJMD_FRAMES=0 turns it off)
  int sorted = 0;            // slot[] currently holds a Hilbert order (not the identity)
  int sort_interval = 128;   // steps between re-sorts
  int sort_bits = 1, sort_ncell = 8;
  long long steps_since_sort = 0;
  int* slot = nullptr;       // [cnt]
  int* sort_order = nullptr; // [cnt]
  unsigned int *sort_key = nullptr, *sort_count = nullptr, *sort_offs = nullptr, *sort_bsum = nullptr;
  // CUDA graph of `graph_period` steady-state steps of a batched ljmd_step (single GPU, no instrumentation)
  cudaGraphExec_t graph_exec = nullptr;
  int graph_period = 0, graph_rdf_every = -1, graph_canonical = -1, graph_bc = -1;
  double graph_dt = 0., graph_T0 = 0.;
  long long graph_launches = 0;   // kernel launches / RDF accumulations one replay stands for
  int graph_rdf_nacc = 0;
  // observation trace (ljmd_trace_*)
  int trace_on = 0, trace_cap = 0, trace_n = 0, trace_row = 0, trace_ncounters = 0;
  int trace_nbins[kMaxTraceCounters] = {0};
  SubvolSpec* trace_specs = nullptr;            // device copy of the counter specifications
  unsigned long long* trace_counts = nullptr;   // [trace_cap][trace_row]
  double* trace_scal = nullptr;                 // [trace_cap][kTraceScalars]
  int* trace_idx = nullptr;                     // device: {next row, ticket}
  int graph_trace = -1, graph_trace_rows = 0;   // trace state the cached graph was captured with / rows per replay
  // fabric (world > 1): one window allocation holding posA | upos | rsum | slots | flags, exported through
  // CUDA IPC; once the peers' windows are mapped the per-step collectives run over peer memory, not NCCL
  char* win = nullptr;
  size_t win_bytes = 0;
  Fabric fab;                 // fab.n == 0 until ljmd_fabric_connect succeeded
  void* peer_map[kMaxPeers] = {nullptr};   // cudaIpcOpenMemHandle results (to close on destroy)
  unsigned long long epoch = 0;
  float thr1 = 0.f, thr2 = 0.f;
  cudaStream_t stream = nullptr;
  float4 *pos = nullptr, *posA = nullptr, *vel = nullptr, *force = nullptr, *tforce = nullptr, *fpart = nullptr;
  float4* gath = nullptr;  // [npad] scratch for on-demand all-gathers of sharded arrays
  uint4* upos = nullptr;
  double *blockW = nullptr, *part = nullptr;
  unsigned int* counter = nullptr;
  unsigned int* velh = nullptr;
  DevScalars* sc = nullptr;
  DevScalars* h_sc = nullptr;  // pinned mirror
  unsigned long long *rdf_cur = nullptr, *rdf_acc = nullptr;
  unsigned long long* h_rdf = nullptr;  // pinned [256]
  void* flush_buf = nullptr;  // optional L2-flush scratch (benchmark hygiene)
  size_t flush_bytes = 0;
  int rdf_valid = 0;  // rdf_cur matches the latest force evaluation
  int rdf_nacc = 0;
  long long launches = 0;
  // event timing of the force kernel
  int timing = 0;
  std::vector<cudaEvent_t> ev;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  double last_force_ms = 0., last_total_ms = 0., last_steps_ms = 0.;
  std::vector<cudaEvent_t> step_ev;   // per-step (begin, end) pairs: step time without the L2-flush write
  std::vector<cudaEvent_t> gath_ev;   // (begin, end) pairs around k_gather: the dominant HBM-bound kernel
  std::vector<cudaEvent_t> red_ev;    // (begin, end) pairs around k_reduce_reaction (sharded Newton-3 runs)
  double last_gather_ms = 0., last_reduce_ms = 0.;
  int last_gather_launches = 0, last_reduce_launches = 0;
  int last_force_launches = 0;
#ifdef LJMD_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
};

// MDSystem.cpp:732-739 on the host, for the threshold bisection only.
static int fast_round_host(float x) { return x > 0 ? (int)(x + 0.5f) : (int)(x - 0.5f); }

// Smallest positive float d with fast_round((float)((double)d / L)) >= k.  The map is monotone
// in d, so bisect over the (ordered) bit patterns of positive floats.
static float image_threshold(double L, int k) {
  uint32_t lo = 0u, hi = 0x7f7fffffu;  // fast_round(0)=0 < k ; FLT_MAX/L rounds huge
  while (hi - lo > 1u) {
    uint32_t mid = lo + (hi - lo) / 2u;
    float d;
    memcpy(&d, &mid, 4);
    double q = (double)d / L;
    float qf = (float)q;
    int n = (qf < 1.0e9f) ? fast_round_host(qf) : 0x7fffffff;
    if (n >= k) hi = mid; else lo = mid;
  }
  float d;
  memcpy(&d, &hi, 4);
  return d;
}

// Split heuristic (DESIGN.md §grid sizing).  The grid is n_itiles x S CTAs of equal work and the SM holds
// `resident` of them, so the launch runs in ceil(n_itiles*S / (SMs*resident)) waves of (work/S + overhead)
// j-iterations each; pick the S that minimises that product, smallest S on ties (fewer partial rows).
// `work` is the number of j-iterations of one i-tile (N for the ordered kernel, its units x BJ for Newton-3),
// `smax` the finest split that still leaves whole work units.
static int choose_split(int n_itiles, long long work, int smax, int num_sms, int resident, int nloc) {
  const double ovh = 128.;  // per-CTA fixed cost (prologue, first tile latency, epilogue) in j-iterations
  const long long slots = (long long)num_sms * resident;
  int best = 1;
  double best_cost = 1e300;
  smax = std::max(1, std::min(smax, 8 * num_sms));
  for (int s = 1; s <= smax; ++s) {
    if ((double)s * nloc * 16. > 1.5e9) break;  // partial-force buffer cap
    const long long waves = ((long long)n_itiles * s + slots - 1) / slots;
    const double cost = (double)waves * ((double)work / s + ovh);
    if (cost < best_cost * 0.999) { best_cost = cost; best = s; }
  }
  return best;
}

struct Plan {
  int nblk, bpr, cnt, i_begin, i_end, nloc, n_itiles, use_sym, hmax, nsplit, bj;
  int mi, mju, nwin, n_super, win_shift;   // Newton-3 kernel: super-tile geometry (nsplit == nwin)
};

// Newton-3 kernel: an i-tile's work is `units` chunks of bj j-records that cannot be cut further.  Measured on
// the B200 (tools/tune_force N reps scan, profiles/r01_tune_force_sym_split_scan.log): at every size from 8 192
// to 131 072 particles the kernel gets faster the MORE and SMALLER its CTAs are — down to one or two units per
// CTA — because the hardware block scheduler then balances the SMs and the tail in which an SM runs a single CTA
// (one warp per scheduler, well under half the issue rate) shrinks to one unit.  Few large CTAs sized to "whole
// waves" lost 5 % (65 536), 16 % (32 768) and 15 % (16 384) against that.  So: aim for ~10 CTAs per slot, never
// fewer than ~4.5 per slot (finer units instead: bj 128 or 64).  Large systems have units to spare: there a CTA
// takes a super-tile of mi i-tiles x mju units (up to 16 x 16), which divides the partial-force and reaction
// traffic by mju and mi, as long as ~32 CTAs per slot remain (tail under ~1.5 %).
static void choose_sym_plan(int n_itiles, int nblk, int num_sms, Plan* pl) {
  const long long slots = (long long)num_sms * kSymMinBlocks;
  const long long want_lo = (9 * slots) / 2, want_hi = 10 * slots;
  const int hmax = sym_max_partner_count(nblk);
  const int partners = hmax + 1;
  int bj = kSymBJ;
  while (bj > 64 && (long long)n_itiles * partners * (kITile / bj) < want_lo) bj >>= 1;
  const int cpb = kITile / bj;
  const int units = partners * cpb;
  const long long total = (long long)n_itiles * units;
  int per_cta = (int)std::max<long long>(1, (total + want_hi - 1) / want_hi);
  // CTAs of 8+ units amortise their prologue and partial-force row: go on to ~60 CTAs per slot, which keeps the
  // tail (about half a CTA duration at low residency) under 1 % of the launch
  if (per_cta >= 8) per_cta = (int)std::max<long long>(8, (total + 60 * slots - 1) / (60 * slots));
  // Between the two regimes (about 10 CTAs of 3-8 units per slot) the launch is a handful of "waves" and an
  // unlucky count leaves most slots idle for a whole CTA duration: the 1/8 shard of C4 (32 896 units) ran 9.3 waves
  // of 8-unit CTAs in 3.02 ms where 2.77 ms is a perfect split.  Halve the CTAs until ~16 per slot are in flight.
  while (per_cta > 2 && total / per_cta < 16 * slots) per_cta = (per_cta + 1) / 2;
  int mju = std::min(per_cta, kSymMaxMJU);
  int mi = 1;
  if (per_cta > mju) {
    mi = (int)std::min<long long>(kSymMaxMI, total / ((long long)mju * 32 * slots));
    mi = std::max(1, std::min(mi, std::min(n_itiles, nblk / 2)));
  }
  const int band = sym_band_units(mi, hmax, nblk, cpb);
  pl->bj = bj; pl->mi = mi; pl->mju = mju;
  pl->nwin = (band + mju - 1) / mju;
  pl->n_super = (n_itiles + mi - 1) / mi;
  // the first ceil((mi-1)*cpb / mju) windows are triangles (tile t starts at its own diagonal): launch them last
  pl->win_shift = std::min(pl->nwin - 1, ((mi - 1) * cpb + mju - 1) / mju);
  pl->nsplit = pl->nwin;
}
// Shards are whole kITile-blocks so that an i-tile never straddles two ranks (the Newton-3 block pairing
// needs global block indices).  LJMD_KERNEL=ordered|sym overrides the kernel choice (force_ordered: internal).
static Plan make_plan(int N, int rank, int world, int num_sms, bool force_ordered = false) {
  Plan pl;
  pl.nblk = (N + kITile - 1) / kITile;
  pl.bpr = (pl.nblk + world - 1) / world;
  pl.cnt = pl.bpr * kITile;
  pl.i_begin = std::min(N, rank * pl.cnt);
  pl.i_end = std::min(N, (rank + 1) * pl.cnt);
  pl.nloc = pl.i_end - pl.i_begin;
  pl.n_itiles = (pl.nloc + kITile - 1) / kITile;
  pl.use_sym = pl.nblk >= kSymMinBlocksN ? 1 : 0;
  if (const char* e = getenv("LJMD_KERNEL")) {
    if (!strcmp(e, "ordered")) pl.use_sym = 0;
    if (!strcmp(e, "sym")) pl.use_sym = 1;
  }
  if (force_ordered) pl.use_sym = 0;
  pl.hmax = std::max(1, sym_max_partner_count(pl.nblk));
  if (pl.nloc <= 0) { pl.nsplit = 0; return pl; }
  pl.bj = kSymBJ; pl.mi = 1; pl.mju = 1; pl.nwin = 1; pl.n_super = pl.n_itiles; pl.win_shift = 0;
  if (pl.use_sym) {
    choose_sym_plan(pl.n_itiles, pl.nblk, num_sms, &pl);
  } else if (N <= 2048) {
    // small systems are latency-bound: a CTA's fixed cost is worth ~16-32 j-iterations, not the 128 of the wave
    // model, and the kernel keeps getting faster down to 16 j-records per CTA (tools/tune_force N reps
    // scan_ordered: N = 400 11.3 -> 6.8 us, N = 1 024 10.8 -> 7.2 us) — but every split is one more row for
    // k_gather (~0.08 us each at this size: 94 splits at N = 1 500 cost more than they gained), hence the cap
    pl.nsplit = std::max(1, std::min((N + 15) / 16, 32));
  } else {
    pl.nsplit = choose_split(pl.n_itiles, N, N / 64, num_sms, kMinBlocks, pl.cnt);
  }
  return pl;
}

static StepParams make_step_params(ljmd_system* s, double dt) {
  StepParams p;
  memset(&p, 0, sizeof(p));
  p.N = s->N; p.nloc = s->nloc; p.i_begin = s->i_begin; p.bc = s->bc;
  p.nsplit = s->nsplit; p.ilocal_cap = s->cnt; p.world = s->world;
  p.nforce_blocks = (s->use_sym ? s->n_super : s->n_itiles) * s->nsplit;
  p.gather_shift = s->gather_shift;
  p.dt = dt; p.dt2 = dt * dt; p.L = s->L; p.rho = s->rho; p.T0 = s->T0;
  p.fix_scale = 4294967296.0 / s->L;
  p.pos = s->pos; p.posA = s->posA; p.upos = s->upos; p.vel = s->vel; p.force = s->force; p.tforce = s->tforce;
  p.fpart = s->fpart; p.blockW = s->blockW; p.part = s->part; p.counter = s->counter; p.sc = s->sc;
  p.use_sym = s->use_sym; p.nblk = s->nblk; p.blk0 = s->i_begin / kITile; p.n_itiles = s->n_itiles;
  p.rp_stride = s->sym_mju * s->sym_bj; p.sym_bj = s->sym_bj; p.sym_mi = s->sym_mi; p.sym_mju = s->sym_mju;
  p.sym_nwin = s->sym_nwin; p.n_super = s->n_super; p.rpart = s->rpart; p.rsum = s->rsum; p.rshard = s->rshard; p.npad = s->npad;
  p.fab = s->fab;
  p.slot = s->slot;
  p.order = s->sort_order;   // identity until the first sort (nullptr with LJMD_FRAMES=0)
  return p;
}

static int step_grid(const ljmd_system* s) { return (s->nloc + kStepThreads - 1) / kStepThreads; }
// k_gather runs 2^gather_shift lanes per particle
static int gather_grid(const ljmd_system* s) {
  return (int)((((long long)s->nloc << s->gather_shift) + kStepThreads - 1) / kStepThreads);
}

// Launch with programmatic stream serialization when `pdl` (the kernel calls pdl_trigger / pdl_wait first).
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static void trace_free(ljmd_system* s);
static int trace_record(ljmd_system* s);

template <bool PERIODIC, bool RDF>
static cudaError_t launch_force_t(ljmd_system* s, const ForceParams& fp) {
  constexpr int MINB = RDF ? kMinBlocksRdf : kMinBlocks;
  auto kern = k_force<P2, PERIODIC, RDF, kForceThreads, MINB, kNPair, kUnroll>;
  const size_t smem = force_smem_bytes(RDF, kTileJ, kForceThreads);
  static_assert(2 * kTileJ * 16 + 16 + 64 + (kForceThreads / 32) * (kRdfBins * 4 + kRdfQueueCap * 8) <= 48 * 1024,
                "dynamic shared memory must stay under the 48 KB default (no per-device opt-in needed)");
  dim3 grid(s->n_itiles, s->nsplit);
  return launch_k(s->pdl != 0, kern, grid, dim3(kForceThreads), smem, s->stream, fp);
}

template <bool PERIODIC, bool RDF, bool FRAMES = false>
static cudaError_t launch_force_sym_t(ljmd_system* s, const SymParams& sp) {
  constexpr int MINB = RDF ? kMinBlocksRdf : kSymMinBlocks;
  auto kern = k_force_sym<P2, PERIODIC, RDF, kForceThreads, MINB, kNPair, kUnroll, FRAMES>;
  const size_t smem = force_sym_smem_bytes(RDF, s->sym_bj, kForceThreads, s->sym_mju);
  dim3 grid(s->n_super, s->sym_nwin);
  return launch_k(s->pdl != 0, kern, grid, dim3(kForceThreads), smem, s->stream, sp);
}
// The window accumulator takes the Newton-3 kernel past the 48 KB default of dynamic shared memory: opt in once
// per device, for the largest window the planner can choose.
template <bool PERIODIC, bool RDF, bool FRAMES = false>
static cudaError_t sym_smem_opt_in() {
  constexpr int MINB = RDF ? kMinBlocksRdf : kSymMinBlocks;
  auto kern = k_force_sym<P2, PERIODIC, RDF, kForceThreads, MINB, kNPair, kUnroll, FRAMES>;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)force_sym_smem_bytes(RDF, kSymBJ, kForceThreads, kSymMaxMJU));
}

static int launch_force(ljmd_system* s, bool rdf) {
  NvtxRange nvtx_(rdf ? "ljmd:force+rdf" : "ljmd:force");
  ForceParams fp;
  memset(&fp, 0, sizeof(fp));
  const bool periodic = (s->bc == LJMD_BC_PERIODIC);
  fp.jrec = periodic ? s->upos : reinterpret_cast<const uint4*>(s->posA);
  fp.posf = s->posA;
  fp.fpart = s->fpart;
  fp.blockW = s->blockW;
  fp.rdf = s->rdf_cur;
  fp.N = s->N; fp.i_begin = s->i_begin; fp.i_end = s->i_end; fp.ilocal_cap = s->cnt; fp.tile_j = kTileJ;
  const double k2 = 4294967296.0 / s->L;
  fp.c2 = periodic ? (float)(k2 * k2) : 1.f;
  fp.fscale = periodic ? (float)(4.0 * s->L / 4294967296.0) : 4.f;
  const double cut = (double)kRdfBins * (double)s->dr2 * 1.001;
  fp.cut_fast = (float)(periodic ? cut * k2 * k2 : cut);
  fp.L = s->L; fp.thr1 = s->thr1; fp.thr2 = s->thr2; fp.dr2 = s->dr2; fp.inv_dr2 = 1.0f / s->dr2;
  const uint4* bbox = nullptr;
  // RDF evaluations of sorted records run the FRAMES build, which prunes chunk by chunk against the warp's own box
  // (N = 65 536, L = 39: 2.21 -> 2.07 ms; N = 262 144, L = 96: 29.2 -> 27.0 ms) — except in small boxes, where
  // nearly every chunk is within histogram range of every warp and the block-box kernel is as good (L = 27: 0.257 vs
  // 0.262 ms).
  const bool frames_rdf = rdf && s->use_sym && periodic && s->frames && s->sorted && s->L >= 32.;
  if (rdf && s->use_sym && !frames_rdf) {
    // block bounding boxes: units whose two boxes are out of histogram range skip the RDF test altogether
    if (periodic) k_bbox<true, kITile><<<s->nblk, 128, 0, s->stream>>>(fp.jrec, s->N, s->bbox);
    else k_bbox<false, kITile><<<s->nblk, 128, 0, s->stream>>>(fp.jrec, s->N, s->bbox);
    CU(cudaGetLastError());
    s->launches += 1;
    bbox = s->bbox;
  }
  if (rdf) CU(cudaMemsetAsync(s->rdf_cur, 0, kRdfBins * sizeof(unsigned long long), s->stream));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (s->timing) {
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, s->stream));
  }
  cudaError_t e;
  if (s->use_sym) {
    SymParams sp;
    sp.f = fp;
    sp.f.tile_j = s->sym_bj;
    sp.rpart = s->rpart; sp.ncols = 0; sp.nblk = s->nblk; sp.bj = s->sym_bj;
    sp.mi = s->sym_mi; sp.mju = s->sym_mju; sp.nwin = s->sym_nwin; sp.win_shift = s->sym_win_shift;
    sp.bbox = bbox; sp.bbox_cut2 = (float)(cut * 1.002 + 1e-3);
    // warp frames: only worth testing for when the records are spatially sorted
    sp.frames = (periodic && s->frames && s->sorted) ? 1 : 0;
    sp.kunit = (float)(s->L / 4294967296.0);
    sp.far2 = (float)((kFrameRfar * k2) * (kFrameRfar * k2));
    sp.rdf2 = (float)((cut * 1.002 + 1e-3) * k2 * k2);
    // RDF evaluations of sorted records: the FRAMES build prunes chunk by chunk against the warp's own box (no
    // block boxes needed); unsorted records keep the fixed-point RDF kernel with its block-box pruning
    if (periodic && s->frames && (!rdf || frames_rdf)) e = rdf ? launch_force_sym_t<true, true, true>(s, sp) : launch_force_sym_t<true, false, true>(s, sp);
    else if (periodic) e = rdf ? launch_force_sym_t<true, true>(s, sp) : launch_force_sym_t<true, false>(s, sp);
    else e = rdf ? launch_force_sym_t<false, true>(s, sp) : launch_force_sym_t<false, false>(s, sp);
  } else if (periodic) e = rdf ? launch_force_t<true, true>(s, fp) : launch_force_t<true, false>(s, fp);
  else e = rdf ? launch_force_t<false, true>(s, fp) : launch_force_t<false, false>(s, fp);
  if (e != cudaSuccess) return set_err(LJMD_ERR_CUDA, "force kernel launch: %s", cudaGetErrorString(e));
  if (s->timing) {
    CU(cudaEventRecord(e1, s->stream));
    s->ev.push_back(e0);
    s->ev.push_back(e1);
  }
  s->launches += 1;
  s->rdf_valid = rdf ? 1 : 0;
  return LJMD_OK;
}

// ---- collectives (no-ops for world == 1) -------------------------------------------------------
// Two transports: the fabric (peer windows over NVLink, ljmd_fabric_connect) and NCCL (always available,
// and still used for the rare read-out collectives).
static int fabric_sync(ljmd_system* s, int first, int count, int fin = 0, double dt = 0.) {
  NvtxRange nvtx_("ljmd:exchange(fabric barrier)");
  s->epoch += 1;
  k_fabric_sync<<<1, 32, 0, s->stream>>>(s->fab, s->epoch, first, count, s->sc, fin, s->N, s->rho, dt);
  CU(cudaGetLastError());
  s->launches += 1;
  return LJMD_OK;
}
// after k_drift / k_prepare: every rank's evaluation positions must be in every window
static int allgather_positions(ljmd_system* s) {
  if (s->world == 1) return LJMD_OK;
  if (s->fab.n > 0) return fabric_sync(s, 0, 0);   // the kernels pushed the records themselves: barrier only
  NvtxRange nvtx_("ljmd:exchange(nccl all-gather)");
#ifdef LJMD_WITH_NCCL
  const size_t bytes = (size_t)s->cnt * 16;
  NC(ncclAllGather((const char*)s->posA + (size_t)s->rank * bytes, s->posA, bytes, ncclChar, s->comm, s->stream));
  if (s->bc == LJMD_BC_PERIODIC)
    NC(ncclAllGather((const char*)s->upos + (size_t)s->rank * bytes, s->upos, bytes, ncclChar, s->comm, s->stream));
#endif
  return LJMD_OK;
}
static int allreduce_sums(ljmd_system* s, int first, int count) {
  if (s->world == 1) return LJMD_OK;
  if (s->fab.n > 0) return fabric_sync(s, first, count);
  NvtxRange nvtx_("ljmd:exchange(nccl all-reduce)");
#ifdef LJMD_WITH_NCCL
  double* ptr = s->sc->sums + first;
  NC(ncclAllReduce(ptr, ptr, count, ncclDouble, ncclSum, s->comm, s->stream));
#endif
  return LJMD_OK;
}

// Re-sort this rank's records along the Hilbert curve (periodic Newton-3 runs).  Call only where every record is
// rewritten before it is read again: right before k_drift / k_prepare.  `force`: sort even if not yet due.
static int maybe_sort_records(ljmd_system* s, bool force) {
  if (!s->frames || !s->slot) return LJMD_OK;
  if (s->bc != LJMD_BC_PERIODIC) return LJMD_OK;   // open boxes have no image to save; keep whatever order there is
  if (!force && s->sorted && s->steps_since_sort < s->sort_interval) return LJMD_OK;
  NvtxRange nvtx_("ljmd:sort records");
  SortParams q;
  q.pos = s->pos; q.nloc = s->nloc; q.i_begin = s->i_begin; q.bits = s->sort_bits; q.ncell = s->sort_ncell;
  q.fix_scale = 4294967296.0 / s->L;
  q.key = s->sort_key; q.count = s->sort_count; q.offs = s->sort_offs; q.bsum = s->sort_bsum;
  q.order = s->sort_order; q.slot = s->slot;
  const int gp = (s->nloc + kSortThreads - 1) / kSortThreads;
  const int gc = (s->sort_ncell + kSortThreads - 1) / kSortThreads;
  const int nb = (s->sort_ncell + kScanBlock - 1) / kScanBlock;
  k_sort_keys<<<gp, kSortThreads, 0, s->stream>>>(q);
  k_scan_local<<<nb, kSortThreads, 0, s->stream>>>(q.count, q.offs, q.bsum, q.ncell);
  k_scan_top<<<1, kSortThreads, 0, s->stream>>>(q.bsum, nb);
  k_scan_add<<<gc, kSortThreads, 0, s->stream>>>(q.offs, q.bsum, q.ncell, (unsigned int)s->nloc);
  k_sort_scatter<<<gp, kSortThreads, 0, s->stream>>>(q);
  k_sort_slots<<<gc, kSortThreads, 0, s->stream>>>(q);
  CU(cudaGetLastError());
  s->launches += 6;
  s->sorted = 1;
  s->steps_since_sort = 0;
  return LJMD_OK;
}

// force evaluation at posA/upos + gather in the given mode
// fuse_next: another Integrate follows immediately, so the finishing kernel also does its drift
static int evaluate(ljmd_system* s, const StepParams& p, int mode, bool rdf, int accumulate, bool fuse_next = false) {
  int rc = launch_force(s, rdf);
  if (rc) return rc;
  const int g = step_grid(s);
  const int fin = (s->world == 1) ? 1 : 0;
  if (s->use_sym && s->world > 1) {
    // reaction forces land on particles of every rank: column sums of the local rows, then either a barrier
    // (fabric: k_gather pulls the peers' sums for its own particles) or a reduce-scatter (NCCL)
    cudaEvent_t r0 = nullptr, r1 = nullptr;
    if (s->timing) {
      CU(cudaEventCreate(&r0));
      CU(cudaEventCreate(&r1));
      CU(cudaEventRecord(r0, s->stream));
    }
    k_reduce_reaction<<<(s->npad + kStepThreads - 1) / kStepThreads, kStepThreads, 0, s->stream>>>(p);
    CU(cudaGetLastError());
    s->launches += 1;
    if (s->timing) {
      CU(cudaEventRecord(r1, s->stream));
      s->red_ev.push_back(r0);
      s->red_ev.push_back(r1);
    }
    if (s->fab.n > 0) {
      if ((rc = fabric_sync(s, 0, 0))) return rc;
    } else {
#ifdef LJMD_WITH_NCCL
      NC(ncclReduceScatter(s->rsum, s->rshard, (size_t)s->cnt * 4, ncclFloat, ncclSum, s->comm, s->stream));
#endif
    }
  }
  cudaEvent_t g0 = nullptr, g1 = nullptr;
  if (s->timing) {
    CU(cudaEventCreate(&g0));
    CU(cudaEventCreate(&g1));
    CU(cudaEventRecord(g0, s->stream));
  }
  NvtxRange nvtx_("ljmd:gather+finish");
  const int gg = gather_grid(s);
  const bool pdl = s->pdl != 0;
  const dim3 gb(kStepThreads);
  if (mode == GATHER_EVAL) CU(launch_k(pdl, k_gather<GATHER_EVAL, false>, dim3(gg), gb, 0, s->stream, p, fin, accumulate));
  else if (mode == GATHER_EVN && fuse_next) CU(launch_k(pdl, k_gather<GATHER_EVN, true>, dim3(gg), gb, 0, s->stream, p, fin, accumulate));
  else if (mode == GATHER_EVN) CU(launch_k(pdl, k_gather<GATHER_EVN, false>, dim3(gg), gb, 0, s->stream, p, fin, accumulate));
  else CU(launch_k(pdl, k_gather<GATHER_TVN, false>, dim3(gg), gb, 0, s->stream, p, fin, accumulate));
  s->launches += 1;
  if (s->timing) {
    CU(cudaEventRecord(g1, s->stream));
    s->gath_ev.push_back(g0);
    s->gath_ev.push_back(g1);
  }
  // Fabric: the step's LAST barrier all-reduces the remaining sums, evaluates CalculateParameters in the same
  // one-warp kernel (no k_params launch) and — every rank having pushed its next evaluation positions in the
  // finishing kernel just before — is also the barrier those positions need (no separate one below).
  const bool fab = s->world > 1 && s->fab.n > 0;
  bool positions_synced = false;
  if (mode == GATHER_TVN) {
    if ((rc = allreduce_sums(s, SUM_PE, 3))) return rc;  // PE, W, TV2
    if (fuse_next) CU(launch_k(pdl, k_finish_tvn<true>, dim3(g), gb, 0, s->stream, p, fin));
    else CU(launch_k(pdl, k_finish_tvn<false>, dim3(g), gb, 0, s->stream, p, fin));
    s->launches += 1;
    if (fab) {
      if ((rc = fabric_sync(s, SUM_K, 1, accumulate ? 2 : 1, p.dt))) return rc;
      positions_synced = true;
    } else if (s->world > 1) {
      if ((rc = allreduce_sums(s, SUM_K, 1))) return rc;
      k_params<<<1, 32, 0, s->stream>>>(p, accumulate);
      CU(cudaGetLastError());
      s->launches += 1;
    }
  } else if (fab) {
    if ((rc = fabric_sync(s, SUM_PE, SUM_COUNT, accumulate ? 2 : 1, p.dt))) return rc;
    positions_synced = true;
  } else if (s->world > 1) {
    if ((rc = allreduce_sums(s, SUM_PE, SUM_COUNT))) return rc;
    k_params<<<1, 32, 0, s->stream>>>(p, accumulate);
    CU(cudaGetLastError());
    s->launches += 1;
  }
  if (rdf && mode != GATHER_EVAL) {
    k_rdf_accum<<<1, kRdfBins, 0, s->stream>>>(s->rdf_cur, s->rdf_acc);
    CU(cudaGetLastError());
    s->launches += 1;
    s->rdf_nacc += 1;
  }
  // the fused kernel published the next step's evaluation positions: make them visible on every rank
  if (fuse_next && !positions_synced && (rc = allgather_positions(s))) return rc;
  return LJMD_OK;
}

// One Integrate.  drifted: the previous step's finishing kernel already did this step's drift (fusion);
// fuse_next: do the next step's drift in this step's finishing kernel.
static int one_step(ljmd_system* s, const StepParams& p, bool rdf, bool drifted = false, bool fuse_next = false) {
  NvtxRange nvtx_("ljmd:integrate");
  const int g = step_grid(s);
  int rc;
  if (!drifted) {
    NvtxRange nvtx_d("ljmd:drift");
    if (s->canonical) CU(launch_k(s->pdl != 0, k_drift<true>, dim3(g), dim3(kStepThreads), 0, s->stream, p));
    else CU(launch_k(s->pdl != 0, k_drift<false>, dim3(g), dim3(kStepThreads), 0, s->stream, p));
    s->launches += 1;
    if ((rc = allgather_positions(s))) return rc;
  }
  return evaluate(s, p, s->canonical ? GATHER_TVN : GATHER_EVN, rdf, 1, fuse_next);
}

static int sync_scalars(ljmd_system* s) {
  CU(cudaMemcpyAsync(s->h_sc, s->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (s->h_sc->fabric_timeout)
    return set_err(LJMD_ERR_NCCL, "fabric barrier timed out: a peer rank never arrived (rank %d of %d)", s->rank,
                   s->world);
  return LJMD_OK;
}

static void collect_timing(ljmd_system* s) {
  s->last_force_ms = 0.;
  s->last_force_launches = 0;
  for (size_t k = 0; k + 1 < s->ev.size(); k += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev[k], s->ev[k + 1]) == cudaSuccess) {
      s->last_force_ms += ms;
      s->last_force_launches += 1;
    }
    cudaEventDestroy(s->ev[k]);
    cudaEventDestroy(s->ev[k + 1]);
  }
  s->ev.clear();
  s->last_gather_ms = 0.;
  s->last_gather_launches = 0;
  for (size_t k = 0; k + 1 < s->gath_ev.size(); k += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->gath_ev[k], s->gath_ev[k + 1]) == cudaSuccess) {
      s->last_gather_ms += ms;
      s->last_gather_launches += 1;
    }
    cudaEventDestroy(s->gath_ev[k]);
    cudaEventDestroy(s->gath_ev[k + 1]);
  }
  s->gath_ev.clear();
  s->last_reduce_ms = 0.;
  s->last_reduce_launches = 0;
  for (size_t k = 0; k + 1 < s->red_ev.size(); k += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->red_ev[k], s->red_ev[k + 1]) == cudaSuccess) {
      s->last_reduce_ms += ms;
      s->last_reduce_launches += 1;
    }
    cudaEventDestroy(s->red_ev[k]);
    cudaEventDestroy(s->red_ev[k + 1]);
  }
  s->red_ev.clear();
  for (cudaEvent_t e : s->step_ev) cudaEventDestroy(e);   // per-step pairs are consumed by ljmd_step before this
  s->step_ev.clear();
}

// Run f(sub-handle, rank) on every device of a front handle, each on its own worker thread; first error wins.
static int multi_all(ljmd_system* front, const std::function<int(ljmd_system*, int)>& f);
static ljmd_system* multi_sub(ljmd_system* front, int r);
static int multi_count(const ljmd_system* front);
#define MULTI_ALL(s, expr)                                                                   \
  do {                                                                                       \
    if ((s) && (s)->multi) return multi_all((s), [&](ljmd_system* sub, int r_) -> int { (void)r_; return (expr); }); \
  } while (0)

// ------------------------------------------------------------------------------------ C ABI: A
extern "C" int ljmd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" float ljmd_rdf_dr2(int N) {
  float dr2 = (float)std::max(0.2 * sqrt(100. / N), 0.05);   // MDSystem.cpp:93
  if (250 * dr2 < 25.0) dr2 = (float)(25.0 / 250);            // :94-95
  return dr2;
}

extern "C" int ljmd_nccl_unique_id(void* out128) {
#ifdef LJMD_WITH_NCCL
  ncclUniqueId id;
  NC(ncclGetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId size");
  memcpy(out128, &id, 128);
  return LJMD_OK;
#else
  (void)out128;
  return set_err(LJMD_ERR_NCCL, "library built without NCCL");
#endif
}

static Plan make_plan_ordered(int N, int rank, int world, int num_sms) { return make_plan(N, rank, world, num_sms, true); }

extern "C" float ljmd_image_threshold(double L, int k) { return image_threshold(L, k); }

extern "C" int ljmd_plan(int N, int rank, int world, int num_sms, int* out8) {
  if (!out8 || N < 2 || world < 1 || rank < 0 || rank >= world || num_sms < 1)
    return set_err(LJMD_ERR_ARG, "bad arguments to ljmd_plan");
  const Plan pl = make_plan(N, rank, world, num_sms);
  out8[0] = pl.i_begin; out8[1] = pl.i_end; out8[2] = pl.n_itiles; out8[3] = pl.nsplit;
  out8[4] = (pl.use_sym ? pl.n_super : pl.n_itiles) * pl.nsplit; out8[5] = kITile; out8[6] = pl.use_sym;
  out8[7] = pl.use_sym ? pl.hmax : 0;
  return LJMD_OK;
}

extern "C" int ljmd_plan_newton3(int N, int rank, int world, int num_sms, int* out8) {
  if (!out8 || N < 2 || world < 1 || rank < 0 || rank >= world || num_sms < 1)
    return set_err(LJMD_ERR_ARG, "bad arguments to ljmd_plan_newton3");
  const Plan pl = make_plan(N, rank, world, num_sms);
  out8[0] = pl.use_sym ? pl.bj : 0; out8[1] = pl.mi; out8[2] = pl.mju; out8[3] = pl.nwin; out8[4] = pl.n_super;
  out8[5] = pl.win_shift; out8[6] = pl.nblk; out8[7] = pl.i_begin / kITile;
  return LJMD_OK;
}

static int destroy_impl(ljmd_system* s) {
  if (!s) return LJMD_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
#ifdef LJMD_WITH_NCCL
  if (s->comm) ncclCommDestroy(s->comm);
#endif
  for (int r = 0; r < kMaxPeers; ++r)
    if (s->peer_map[r]) cudaIpcCloseMemHandle(s->peer_map[r]);
  if (s->win) {
    cudaFree(s->win);   // posA, upos and rsum live inside the window
  } else {
    cudaFree(s->posA); cudaFree(s->upos);
  }
  cudaFree(s->pos); cudaFree(s->vel); cudaFree(s->force); cudaFree(s->tforce);
  cudaFree(s->fpart); cudaFree(s->gath); cudaFree(s->blockW); cudaFree(s->part);
  cudaFree(s->counter); cudaFree(s->velh); cudaFree(s->sc); cudaFree(s->rdf_cur); cudaFree(s->rdf_acc);
  cudaFree(s->flush_buf);
  cudaFree(s->rpart); cudaFree(s->rshard); cudaFree(s->bbox);
  cudaFree(s->slot); cudaFree(s->sort_order); cudaFree(s->sort_key); cudaFree(s->sort_count); cudaFree(s->sort_offs);
  cudaFree(s->sort_bsum);
  trace_free(s);   // also destroys the cached graph
  cudaFreeHost(s->h_sc); cudaFreeHost(s->h_rdf);
  for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : s->step_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : s->gath_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : s->red_ev) cudaEventDestroy(e);
  if (s->ev_begin) cudaEventDestroy(s->ev_begin);
  if (s->ev_end) cudaEventDestroy(s->ev_end);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return LJMD_OK;
}

static int create_impl(ljmd_system** out, int N, double rho_or_negL, double T0, int canonical, int bc, float rdf_dr2,
                       int device, int rank, int world, const void* uid, int inproc = 0) {
  if (!out) return set_err(LJMD_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (N < 2) return set_err(LJMD_ERR_ARG, "N must be >= 2 (got %d)", N);
  if (bc < 0 || bc > 2) return set_err(LJMD_ERR_ARG, "boundary condition must be 0, 1 or 2 (got %d)", bc);
  if (world < 1 || rank < 0 || rank >= world) return set_err(LJMD_ERR_ARG, "bad rank/world %d/%d", rank, world);
  if (!(rdf_dr2 > 0.f)) return set_err(LJMD_ERR_ARG, "rdf_dr2 must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(LJMD_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return set_err(LJMD_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  ljmd_system* s = new (std::nothrow) ljmd_system();
  if (!s) return set_err(LJMD_ERR_ARG, "out of host memory");
  s->N = N; s->bc = bc; s->canonical = canonical ? 1 : 0; s->T0 = T0; s->dr2 = rdf_dr2;
  s->device = device; s->rank = rank; s->world = world; s->inproc = inproc;
  if (rho_or_negL > 0.) { s->rho = rho_or_negL; s->L = pow(N / s->rho, 1. / 3.); }   // MDSystem.cpp:70
  else { s->L = -rho_or_negL; s->rho = N / (s->L * s->L * s->L); }
  // (the device properties are needed for the plan; queried again below for the arch check)
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const Plan pl = make_plan(N, rank, world, sms);
    s->nblk = pl.nblk; s->cnt = pl.cnt; s->npad = pl.cnt * world; s->i_begin = pl.i_begin; s->i_end = pl.i_end;
    s->nloc = pl.nloc; s->n_itiles = pl.n_itiles; s->use_sym = pl.use_sym; s->hmax = pl.hmax; s->nsplit = pl.nsplit; s->sym_bj = pl.bj;
    s->sym_mi = pl.mi; s->sym_mju = pl.mju; s->sym_nwin = pl.nwin; s->n_super = pl.n_super; s->sym_win_shift = pl.win_shift;
    s->num_sms = sms;
  }
  if (s->nloc < 1) { delete s; return set_err(LJMD_ERR_ARG, "rank %d of %d has no particles for N=%d", rank, world, N); }
  cudaDeviceProp prop;
  cudaError_t pe = cudaGetDeviceProperties(&prop, device);
  if (pe != cudaSuccess) { delete s; return set_err(LJMD_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(pe)); }
  if (prop.major < 10) {
    delete s;
    return set_err(LJMD_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                   prop.minor);
  }
  s->num_sms = prop.multiProcessorCount;
  s->thr1 = image_threshold(s->L, 1);
  s->thr2 = image_threshold(s->L, 2);

#define CUC(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      set_err(LJMD_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));     \
      destroy_impl(s);                                                                                  \
      return LJMD_ERR_CUDA;                                                                             \
    }                                                                                                   \
  } while (0)
  CUC(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  const size_t b16 = 16;
  CUC(cudaMalloc(&s->pos, (size_t)s->cnt * b16));
  memset(&s->fab, 0, sizeof(s->fab));
  if (world > 1) {
    // one allocation so that a single IPC handle exposes everything the peers touch
    const size_t arr = (size_t)s->npad * b16;
    s->fab.off_posA = 0; s->fab.off_upos = arr; s->fab.off_rsum = 2 * arr;
    s->fab.off_slots = 3 * arr;
    s->fab.off_flags = s->fab.off_slots + (size_t)2 * kMaxPeers * kSlotDoubles * sizeof(double);
    s->win_bytes = s->fab.off_flags + kMaxPeers * sizeof(unsigned long long) + 256;
    CUC(cudaMalloc(&s->win, s->win_bytes));
    CUC(cudaMemsetAsync(s->win, 0, s->win_bytes, s->stream));
    s->posA = reinterpret_cast<float4*>(s->win + s->fab.off_posA);
    s->upos = reinterpret_cast<uint4*>(s->win + s->fab.off_upos);
    s->rsum = reinterpret_cast<float4*>(s->win + s->fab.off_rsum);
  } else {
    CUC(cudaMalloc(&s->posA, (size_t)s->npad * b16));
    CUC(cudaMalloc(&s->upos, (size_t)s->npad * b16));
  }
  CUC(cudaMalloc(&s->gath, (size_t)s->npad * b16));
  CUC(cudaMalloc(&s->vel, (size_t)s->cnt * b16));
  CUC(cudaMalloc(&s->force, (size_t)s->cnt * b16));
  CUC(cudaMalloc(&s->tforce, (size_t)s->cnt * b16));
  if (s->use_sym) {
    // Partial-force rows and reaction blocks grow as N^2 / G (divided by mju and mi): fall back to the ordered
    // kernel when they would take more than 45 % of the device's memory.  The test depends on N, the rank count and
    // the device model only — never on the momentary free memory — so every rank of a sharded system takes the
    // same decision (a rank on its own kernel would miss the reaction exchange of its peers).
    const size_t need = ((size_t)s->nsplit * s->cnt + (size_t)s->n_super * s->sym_nwin * s->sym_mju * s->sym_bj) * b16;
    if ((double)need > 0.45 * (double)prop.totalGlobalMem) {
      s->use_sym = 0;
      const Plan po = make_plan_ordered(N, rank, world, s->num_sms);
      s->nsplit = po.nsplit;
    }
  }
  CUC(cudaMalloc(&s->fpart, (size_t)s->nsplit * s->cnt * b16));
  CUC(cudaMalloc(&s->blockW, (size_t)std::max(s->n_itiles, s->n_super) * s->nsplit * sizeof(double)));
  if (s->use_sym) {
    // every entry of rpart is rewritten by every launch (ljmd_force_sym.cuh): no clearing needed, ever
    const size_t rp = (size_t)s->n_super * s->sym_nwin * s->sym_mju * s->sym_bj * b16;
    CUC(cudaMalloc(&s->rpart, rp));
    if (world > 1) CUC(cudaMalloc(&s->rshard, (size_t)s->cnt * b16));
    CUC(cudaMalloc(&s->bbox, (size_t)s->nblk * 2 * sizeof(uint4)));
    CUC((sym_smem_opt_in<true, false>()));
    CUC((sym_smem_opt_in<true, true>()));
    CUC((sym_smem_opt_in<false, false>()));
    CUC((sym_smem_opt_in<false, true>()));
    CUC((sym_smem_opt_in<true, false, true>()));
    CUC((sym_smem_opt_in<true, true, true>()));
    // record order: identity until the first sort (LJMD_FRAMES=0: stays the identity, fixed-point kernel only)
    {
      const char* e = getenv("LJMD_FRAMES");
      s->frames = (e && e[0] == '0') ? 0 : 1;
      // a sort is six small launches (~20-100 us): every 128 steps where the step itself takes ~0.1 ms, more
      // often where a step takes milliseconds and the clouds should stay tight
      s->sort_interval = s->N >= 262144 ? 16 : (s->N >= 65536 ? 32 : 128);
      if (const char* iv = getenv("LJMD_SORT_INTERVAL")) s->sort_interval = std::max(1, atoi(iv));
    }
    s->sort_bits = 1;
    while (s->sort_bits < 8 && (1LL << (3 * s->sort_bits)) < 2LL * s->nloc) s->sort_bits += 1;
    s->sort_ncell = 1 << (3 * s->sort_bits);
    CUC(cudaMalloc(&s->slot, (size_t)s->cnt * sizeof(int)));
    k_slot_identity<<<(s->nloc + kSortThreads - 1) / kSortThreads, kSortThreads, 0, s->stream>>>(s->slot, s->nloc, s->i_begin);
    CUC(cudaGetLastError());
    if (s->frames) {
      CUC(cudaMalloc(&s->sort_order, (size_t)s->cnt * sizeof(int)));
      k_slot_identity<<<(s->nloc + kSortThreads - 1) / kSortThreads, kSortThreads, 0, s->stream>>>(s->sort_order, s->nloc, 0);
      CUC(cudaGetLastError());
      CUC(cudaMalloc(&s->sort_key, (size_t)s->cnt * sizeof(unsigned int)));
      CUC(cudaMalloc(&s->sort_count, (size_t)s->sort_ncell * sizeof(unsigned int)));
      CUC(cudaMalloc(&s->sort_offs, ((size_t)s->sort_ncell + 1) * sizeof(unsigned int)));
      CUC(cudaMalloc(&s->sort_bsum, ((size_t)(s->sort_ncell + kScanBlock - 1) / kScanBlock + 1) * sizeof(unsigned int)));
      CUC(cudaMemsetAsync(s->sort_count, 0, (size_t)s->sort_ncell * sizeof(unsigned int), s->stream));
    }
  }
  // lanes per particle in k_gather: enough threads to keep ~2 CTAs of 256 on every SM, never more lanes than
  // half the rows they share
  {
    const char* e = getenv("LJMD_PDL");
    s->pdl = (world == 1 && !(e && e[0] == '0')) ? 1 : 0;
  }
  s->gather_shift = 0;
  while (s->gather_shift < 3 && ((long long)s->nloc << s->gather_shift) < 2LL * kStepThreads * s->num_sms &&
         (2 << s->gather_shift) * 2 <= s->nsplit + (s->use_sym ? (s->n_super + 1) / 2 : 0))
    s->gather_shift += 1;
  if (const char* e = getenv("LJMD_GATHER_SHIFT")) {   // tuning override: 2^shift lanes per particle, 0..3
    const int v = atoi(e);
    if (v >= 0 && v <= 3) s->gather_shift = v;
  }
  CUC(cudaMalloc(&s->part, (size_t)2 * (gather_grid(s) + 1) * sizeof(double)));
  CUC(cudaMalloc(&s->counter, sizeof(unsigned int)));
  CUC(cudaMalloc(&s->velh, 65536 * sizeof(unsigned int)));
  CUC(cudaMalloc(&s->sc, sizeof(DevScalars)));
  CUC(cudaMalloc(&s->rdf_cur, kRdfBins * sizeof(unsigned long long)));
  CUC(cudaMalloc(&s->rdf_acc, kRdfBins * sizeof(unsigned long long)));
  CUC(cudaMallocHost(&s->h_sc, sizeof(DevScalars)));
  CUC(cudaMallocHost(&s->h_rdf, kRdfBins * sizeof(unsigned long long)));
  if (world == 1) {
    CUC(cudaMemsetAsync(s->posA, 0, (size_t)s->npad * b16, s->stream));
    CUC(cudaMemsetAsync(s->upos, 0, (size_t)s->npad * b16, s->stream));
  }
  CUC(cudaMemsetAsync(s->pos, 0, (size_t)s->cnt * b16, s->stream));
  CUC(cudaMemsetAsync(s->vel, 0, (size_t)s->cnt * b16, s->stream));
  CUC(cudaMemsetAsync(s->force, 0, (size_t)s->cnt * b16, s->stream));
  CUC(cudaMemsetAsync(s->tforce, 0, (size_t)s->cnt * b16, s->stream));
  CUC(cudaMemsetAsync(s->counter, 0, sizeof(unsigned int), s->stream));
  CUC(cudaMemsetAsync(s->sc, 0, sizeof(DevScalars), s->stream));
  CUC(cudaMemsetAsync(s->rdf_cur, 0, kRdfBins * sizeof(unsigned long long), s->stream));
  CUC(cudaMemsetAsync(s->rdf_acc, 0, kRdfBins * sizeof(unsigned long long), s->stream));
  CUC(cudaEventCreate(&s->ev_begin));
  CUC(cudaEventCreate(&s->ev_end));
  CUC(cudaStreamSynchronize(s->stream));
  memset(s->h_sc, 0, sizeof(DevScalars));
#undef CUC
  if (world > 1 && !inproc) {
#ifdef LJMD_WITH_NCCL
    if (!uid) { destroy_impl(s); return set_err(LJMD_ERR_ARG, "nccl_unique_id is NULL"); }
    ncclUniqueId id;
    memcpy(&id, uid, 128);
    ncclResult_t r = ncclCommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) {
      set_err(LJMD_ERR_NCCL, "ncclCommInitRank: %s", ncclGetErrorString(r));
      s->comm = nullptr;
      destroy_impl(s);
      return LJMD_ERR_NCCL;
    }
#else
    destroy_impl(s);
    return set_err(LJMD_ERR_NCCL, "library built without NCCL; world must be 1");
#endif
  }
  *out = s;
  return LJMD_OK;
}

extern "C" int ljmd_create(ljmd_system** out, int N, double rho, double T0, int canonical, int bc, float rdf_dr2,
                           int device) {
  if (!(rho > 0.)) return set_err(LJMD_ERR_ARG, "rho must be positive");
  return create_impl(out, N, rho, T0, canonical, bc, rdf_dr2, device, 0, 1, nullptr);
}
extern "C" int ljmd_create_distributed(ljmd_system** out, int N, double rho, double T0, int canonical, int bc,
                                       float rdf_dr2, int device, int rank, int world, const void* uid) {
  if (!(rho > 0.)) return set_err(LJMD_ERR_ARG, "rho must be positive");
  return create_impl(out, N, rho, T0, canonical, bc, rdf_dr2, device, rank, world, uid);
}
// Internal (legacy seam): explicit box edge instead of density.
int ljmd_create_with_L(ljmd_system** out, int N, double L, int bc, float rdf_dr2, int device) {
  if (!(L > 0.)) return set_err(LJMD_ERR_ARG, "L must be positive");
  return create_impl(out, N, -L, 1.0, 0, bc, rdf_dr2, device, 0, 1, nullptr);
}

static int multi_destroy(ljmd_system* front);
extern "C" int ljmd_destroy(ljmd_system* s) { return (s && s->multi) ? multi_destroy(s) : destroy_impl(s); }

extern "C" int ljmd_fabric_export(ljmd_system* s, void* out64) {
  if (!s || !out64) return set_err(LJMD_ERR_ARG, "NULL argument");
  if (s->multi || s->inproc) return set_err(LJMD_ERR_ARG, "a single-process multi-GPU system wires its fabric itself");
  CU(cudaSetDevice(s->device));
  if (s->world == 1 || !s->win) return set_err(LJMD_ERR_ARG, "fabric needs a distributed system (world > 1)");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
  CU(cudaIpcGetMemHandle(&h, s->win));
  memcpy(out64, &h, 64);
  return LJMD_OK;
}

extern "C" int ljmd_fabric_connect(ljmd_system* s, const void* handles) {
  if (!s || !handles) return set_err(LJMD_ERR_ARG, "NULL argument");
  if (s->multi || s->inproc) return set_err(LJMD_ERR_ARG, "a single-process multi-GPU system wires its fabric itself");
  CU(cudaSetDevice(s->device));
  if (s->world == 1 || !s->win) return set_err(LJMD_ERR_ARG, "fabric needs a distributed system (world > 1)");
  if (s->world > kMaxPeers) return set_err(LJMD_ERR_ARG, "fabric supports up to %d ranks", kMaxPeers);
  if (s->fab.n > 0) return LJMD_OK;
  CU(cudaStreamSynchronize(s->stream));
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) { s->fab.base[r] = s->win; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (s->peer_map[q]) { cudaIpcCloseMemHandle(s->peer_map[q]); s->peer_map[q] = nullptr; }
      return set_err(LJMD_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (staying on NCCL)", r, cudaGetErrorString(e));
    }
    s->peer_map[r] = ptr;
    s->fab.base[r] = (char*)ptr;
  }
  s->fab.me = s->rank;
  s->fab.n = s->world;   // from now on the step collectives run over the peer windows
  return LJMD_OK;
}

#define CHECK_S(s)                                              \
  do {                                                          \
    if (!(s)) return set_err(LJMD_ERR_ARG, "system is NULL");   \
    CU(cudaSetDevice((s)->device));                             \
  } while (0)

extern "C" int ljmd_set_canonical(ljmd_system* s, int canonical) {
  MULTI_ALL(s, ljmd_set_canonical(sub, canonical));
  CHECK_S(s);
  s->canonical = canonical ? 1 : 0;
  return LJMD_OK;
}
extern "C" int ljmd_set_T0(ljmd_system* s, double T0) {
  MULTI_ALL(s, ljmd_set_T0(sub, T0));
  CHECK_S(s);
  if (!(T0 > 0.)) return set_err(LJMD_ERR_ARG, "T0 must be positive");
  s->T0 = T0;
  return LJMD_OK;
}
extern "C" int ljmd_set_boundary(ljmd_system* s, int bc) {
  MULTI_ALL(s, ljmd_set_boundary(sub, bc));
  CHECK_S(s);
  if (bc < 0 || bc > 2) return set_err(LJMD_ERR_ARG, "boundary condition must be 0, 1 or 2 (got %d)", bc);
  if (bc != s->bc) {
    s->bc = bc;
    s->rdf_valid = 0;
    // Re-anchor the evaluation positions on what the caller sees (the wrapped positions) and,
    // for a periodic box, rebuild the fixed-point records the other modes do not maintain.
    int rc = maybe_sort_records(s, true);   // a box that just became periodic wants its records sorted
    if (rc) return rc;
    StepParams p = make_step_params(s, 0.);
    k_prepare<<<step_grid(s), kStepThreads, 0, s->stream>>>(p);
    CU(cudaGetLastError());
    s->launches += 1;
    if ((rc = allgather_positions(s))) return rc;
    CU(cudaStreamSynchronize(s->stream));
  }
  return LJMD_OK;
}

static int check_domain(const ljmd_system* s, const float* pos4) {
  for (int i = 0; i < s->N; ++i)
    for (int k = 0; k < 3; ++k) {
      const float x = pos4[4 * (size_t)i + k];
      if (!(fabsf(x) <= 3.0e38f))
        return set_err(LJMD_ERR_DOMAIN, "particle %d coordinate %d is not finite", i, k);
    }
  return LJMD_OK;
}

static int upload_state(ljmd_system* s, const float* pos4, const float* vel4) {
  const size_t off = (size_t)s->i_begin * 4;
  if (pos4) CU(cudaMemcpyAsync(s->pos, pos4 + off, (size_t)s->nloc * 16, cudaMemcpyHostToDevice, s->stream));
  if (vel4) CU(cudaMemcpyAsync(s->vel, vel4 + off, (size_t)s->nloc * 16, cudaMemcpyHostToDevice, s->stream));
  return LJMD_OK;
}

extern "C" int ljmd_set_state(ljmd_system* s, const float* pos4, const float* vel4) {
  MULTI_ALL(s, ljmd_set_state(sub, pos4, vel4));
  CHECK_S(s);
  if (!pos4 || !vel4) return set_err(LJMD_ERR_ARG, "pos4/vel4 must not be NULL");
  int rc = check_domain(s, pos4);
  if (rc) return rc;
  if ((rc = upload_state(s, pos4, vel4))) return rc;
  if ((rc = maybe_sort_records(s, true))) return rc;
  StepParams p = make_step_params(s, 0.);
  k_prepare<<<step_grid(s), kStepThreads, 0, s->stream>>>(p);
  CU(cudaGetLastError());
  s->launches += 1;
  if ((rc = allgather_positions(s))) return rc;
  CU(cudaMemsetAsync(s->sc, 0, sizeof(DevScalars), s->stream));  // t = 0, av_* = 0
  if ((rc = evaluate(s, p, GATHER_EVAL, false, 0))) return rc;
  s->rdf_nacc = 0;
  CU(cudaMemsetAsync(s->rdf_acc, 0, kRdfBins * sizeof(unsigned long long), s->stream));
  return sync_scalars(s);
}

extern "C" int ljmd_upload(ljmd_system* s, const float* pos4, const float* vel4) {
  MULTI_ALL(s, ljmd_upload(sub, pos4, vel4));
  CHECK_S(s);
  int rc = upload_state(s, pos4, vel4);
  if (rc) return rc;
  if (pos4) s->rdf_valid = 0;
  CU(cudaStreamSynchronize(s->stream));
  return LJMD_OK;
}

extern "C" int ljmd_set_velocities(ljmd_system* s, const float* vel4) {
  MULTI_ALL(s, ljmd_set_velocities(sub, vel4));
  CHECK_S(s);
  if (!vel4) return set_err(LJMD_ERR_ARG, "vel4 must not be NULL");
  int rc = upload_state(s, nullptr, vel4);
  if (rc) return rc;
  StepParams p = make_step_params(s, 0.);
  k_kinetic<<<step_grid(s), kStepThreads, 0, s->stream>>>(p);
  CU(cudaGetLastError());
  if ((rc = allreduce_sums(s, SUM_K, 1))) return rc;
  k_params<<<1, 32, 0, s->stream>>>(p, 0);
  CU(cudaGetLastError());
  s->launches += 2;
  return sync_scalars(s);
}

// this rank's shard of a sharded array into its place in a full-length host array
static int download_shard(ljmd_system* s, const float4* dev_local, float* host4) {
  CU(cudaMemcpyAsync(host4 + (size_t)s->i_begin * 4, dev_local, (size_t)s->nloc * 16, cudaMemcpyDeviceToHost, s->stream));
  return LJMD_OK;
}
// full-length copy of a sharded array into host memory (one process per GPU: every rank gets everything)
static int download_sharded(ljmd_system* s, const float4* dev_local, float* host4) {
  if (s->world == 1 || s->inproc) return download_shard(s, dev_local, host4);
#ifdef LJMD_WITH_NCCL
  const size_t bytes = (size_t)s->cnt * 16;
  CU(cudaMemcpyAsync((char*)s->gath + (size_t)s->rank * bytes, dev_local, (size_t)s->nloc * 16,
                     cudaMemcpyDeviceToDevice, s->stream));
  NC(ncclAllGather((const char*)s->gath + (size_t)s->rank * bytes, s->gath, bytes, ncclChar, s->comm, s->stream));
  CU(cudaMemcpyAsync(host4, s->gath, (size_t)s->N * 16, cudaMemcpyDeviceToHost, s->stream));
  return LJMD_OK;
#else
  return set_err(LJMD_ERR_NCCL, "built without NCCL");
#endif
}

extern "C" int ljmd_get_state(ljmd_system* s, float* pos4, float* vel4, float* force4) {
  MULTI_ALL(s, ljmd_get_state(sub, pos4, vel4, force4));   // every device fills its own shard of the arrays
  CHECK_S(s);
  int rc;
  if (pos4 && (rc = download_sharded(s, s->pos, pos4))) return rc;
  if (vel4 && (rc = download_sharded(s, s->vel, vel4))) return rc;
  if (force4 && (rc = download_sharded(s, s->force, force4))) return rc;
  CU(cudaStreamSynchronize(s->stream));
  return LJMD_OK;
}

extern "C" int ljmd_step(ljmd_system* s, double dt, int nsteps, int rdf_every) {
  MULTI_ALL(s, ljmd_step(sub, dt, nsteps, rdf_every));
  CHECK_S(s);
  if (nsteps < 0) return set_err(LJMD_ERR_ARG, "nsteps must be >= 0");
  if (s->trace_on && s->trace_n + nsteps > s->trace_cap)
    return set_err(LJMD_ERR_ARG, "trace holds %d of %d steps: read it before %d more", s->trace_n, s->trace_cap, nsteps);
  StepParams p = make_step_params(s, dt);
  if (s->timing) CU(cudaEventRecord(s->ev_begin, s->stream));
  // Periodic Newton-3 runs re-sort their records every sort_interval steps (handle-wide count, so a batch and the
  // same steps one by one sort at the same steps and stay bit-identical).  A sort needs a step that starts with a
  // plain drift, so the batch is cut there into segments, each run by the loop below.
  const int total_steps = nsteps;
  int seg_done = 0;
  do {
  nsteps = total_steps - seg_done;
  {
    int rc = maybe_sort_records(s, false);
    if (rc) return rc;
    if (s->frames && s->slot && s->bc == LJMD_BC_PERIODIC) {
      const long long room = std::max<long long>(1, s->sort_interval - s->steps_since_sort);
      if (room < nsteps) nsteps = (int)room;
    }
  }
  bool drifted = false;
  // Steady-state steps of a batch (previous step fused this one's drift, this one fuses the next) are the same
  // launches with the same arguments, RDF cadence included: capture `period` of them once in a CUDA graph and
  // replay it.  For small systems the step is launch-bound (N = 400: three ~5 us kernels), and a graph launch
  // replaces 3-6 kernel launches per step by one submission per period.  Off with instrumentation (event
  // timing, L2 flush, trace), on more than one GPU, and with LJMD_GRAPH=0.
  const int period = rdf_every > 0 ? rdf_every : 16;
  const char* genv = getenv("LJMD_GRAPH");
  // (the captured RDF cadence counts from the segment's first step: only segments that start on it replay)
  const bool use_graph = s->world == 1 && !s->timing && s->flush_bytes == 0 && period <= 64 &&
                         nsteps - 2 >= 2 * period && !(genv && genv[0] == '0') &&
                         (rdf_every <= 0 || seg_done % rdf_every == 0);
  // Kick-drift-wrap fusion inside a batch.  With the fabric an EVN step of the ordered kernel has no barrier
  // between a peer's force kernel and this rank's finishing kernel, so its position pushes must not be fused.
  // A trace row needs the end-of-step velocities: the fused EVN half-kick of the next step would be in them
  // (TVN has no first half-kick, so traced TVN batches keep the fusion).
  const bool steady_fuse = ((s->world == 1) || s->fab.n == 0 || s->use_sym || s->canonical) &&
                           !(s->trace_on && !s->canonical);
  for (int k = 0; k < nsteps; ++k) {
    if (use_graph && k == 1) {
      const bool fresh = s->graph_exec && s->graph_period == period && s->graph_rdf_every == rdf_every &&
                         s->graph_canonical == s->canonical && s->graph_bc == s->bc && s->graph_dt == dt &&
                         s->graph_T0 == s->T0 && s->graph_trace == s->trace_on;
      if (!fresh) {
        if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
        const long long l0 = s->launches;
        const int r0 = s->rdf_nacc, t0 = s->trace_n;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        int rc = LJMD_OK;
        for (int j = 0; j < period && rc == LJMD_OK; ++j) {
          const int kk = 1 + j;
          rc = one_step(s, p, rdf_every > 0 && ((kk + 1) % rdf_every == 0), steady_fuse, steady_fuse);
          if (rc == LJMD_OK && s->trace_on) rc = trace_record(s);
        }
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
        s->graph_launches = s->launches - l0;
        s->graph_rdf_nacc = s->rdf_nacc - r0;
        s->graph_trace_rows = s->trace_n - t0;
        s->launches = l0;      // nothing ran yet
        s->rdf_nacc = r0;
        s->trace_n = t0;
        if (rc != LJMD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return set_err(LJMD_ERR_CUDA, "graph capture: %s", cudaGetErrorString(ce));
        const cudaError_t ie = cudaGraphInstantiate(&s->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { s->graph_exec = nullptr; return set_err(LJMD_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(ie)); }
        s->graph_period = period; s->graph_rdf_every = rdf_every; s->graph_canonical = s->canonical;
        s->graph_bc = s->bc; s->graph_dt = dt; s->graph_T0 = s->T0; s->graph_trace = s->trace_on;
      }
      const int reps = (nsteps - 2) / period;
      for (int r = 0; r < reps; ++r) {
        CU(cudaGraphLaunch(s->graph_exec, s->stream));
        s->launches += s->graph_launches;
        s->rdf_nacc += s->graph_rdf_nacc;
        s->trace_n += s->graph_trace_rows;
      }
      k += reps * period;    // the steps left (at least the last one) run below, with the same cadence
    }
    const bool rdf = rdf_every > 0 && ((seg_done + k + 1) % rdf_every == 0);
    if (s->flush_bytes) CU(cudaMemsetAsync(s->flush_buf, k & 0xff, s->flush_bytes, s->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing) {
      CU(cudaEventCreate(&e0));
      CU(cudaEventCreate(&e1));
      CU(cudaEventRecord(e0, s->stream));
    }
    const bool fuse_next = steady_fuse && (k + 1 < nsteps);
    int rc = one_step(s, p, rdf, drifted, fuse_next);
    if (rc) return rc;
    if (s->trace_on && (rc = trace_record(s))) return rc;
    drifted = fuse_next;
    if (s->timing) {
      CU(cudaEventRecord(e1, s->stream));
      s->step_ev.push_back(e0);
      s->step_ev.push_back(e1);
    }
  }
  seg_done += nsteps;
  s->steps_since_sort += nsteps;
  } while (seg_done < total_steps);
  if (s->timing) CU(cudaEventRecord(s->ev_end, s->stream));
  int rc = sync_scalars(s);
  if (rc) return rc;
  if (s->timing) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev_begin, s->ev_end));
    s->last_total_ms = ms;
    s->last_steps_ms = 0.;
    for (size_t k = 0; k + 1 < s->step_ev.size(); k += 2) {
      if (cudaEventElapsedTime(&ms, s->step_ev[k], s->step_ev[k + 1]) == cudaSuccess) s->last_steps_ms += ms;
      cudaEventDestroy(s->step_ev[k]);
      cudaEventDestroy(s->step_ev[k + 1]);
    }
    s->step_ev.clear();
    collect_timing(s);
  }
  return LJMD_OK;
}

extern "C" int ljmd_integrate_host(ljmd_system* s, double dt, float* pos4, float* vel4, float* force4) {
  MULTI_ALL(s, ljmd_integrate_host(sub, dt, pos4, vel4, force4));
  CHECK_S(s);
  if (!pos4 || !vel4) return set_err(LJMD_ERR_ARG, "pos4/vel4 must not be NULL");
  int rc = upload_state(s, pos4, vel4);
  if (rc) return rc;
  if ((rc = maybe_sort_records(s, false))) return rc;
  s->steps_since_sort += 1;
  StepParams p = make_step_params(s, dt);
  if ((rc = one_step(s, p, false))) return rc;
  // a rank moves ITS shard both ways: up from the caller's arrays, back into the same places.  On one GPU (and
  // behind a multi-GPU front handle, where the devices share the arrays) that is the whole state; one process
  // per GPU gets the other shards through ljmd_get_state when it wants them.
  if ((rc = download_shard(s, s->pos, pos4))) return rc;
  if ((rc = download_shard(s, s->vel, vel4))) return rc;
  if (force4 && (rc = download_shard(s, s->force, force4))) return rc;
  rc = sync_scalars(s);
  if (s->timing) collect_timing(s);
  return rc;
}

extern "C" int ljmd_compute_forces(ljmd_system* s, int with_rdf) {
  MULTI_ALL(s, ljmd_compute_forces(sub, with_rdf));
  CHECK_S(s);
  // positions the caller sees are the wrapped ones: evaluate there (CalculateForces reads h_Pos)
  int rc = maybe_sort_records(s, false);
  if (rc) return rc;
  StepParams p = make_step_params(s, 0.);
  k_prepare<<<step_grid(s), kStepThreads, 0, s->stream>>>(p);
  CU(cudaGetLastError());
  s->launches += 1;
  if ((rc = allgather_positions(s))) return rc;
  if ((rc = evaluate(s, p, GATHER_EVAL, with_rdf != 0, 0))) return rc;
  rc = sync_scalars(s);
  if (s->timing) collect_timing(s);
  return rc;
}

extern "C" int ljmd_get_scalars(ljmd_system* s, double* out) {
  if (s && s->multi) return ljmd_get_scalars(multi_sub(s, 0), out);   // identical on every device (fixed-order sums)
  CHECK_S(s);
  if (!out) return set_err(LJMD_ERR_ARG, "out is NULL");
  const DevScalars& h = *s->h_sc;
  for (int k = 0; k < LJMD_S_COUNT; ++k) out[k] = 0.;
  out[LJMD_S_U] = h.U; out[LJMD_S_T] = h.T; out[LJMD_S_K] = h.K; out[LJMD_S_V] = h.V; out[LJMD_S_P] = h.P;
  out[LJMD_S_PVIRIAL] = h.Pvirial; out[LJMD_S_TIME] = h.t; out[LJMD_S_L] = s->L;
  out[LJMD_S_AV_U_TOT] = h.av_U_tot; out[LJMD_S_AV_T_TOT] = h.av_T_tot; out[LJMD_S_AV_P_TOT] = h.av_p_tot;
  out[LJMD_S_AV_ITERS] = (double)h.av_iters; out[LJMD_S_CHI] = h.chi; out[LJMD_S_TKIN_TRIAL] = h.Tkin_trial;
  return LJMD_OK;
}

extern "C" int ljmd_reset_averaging(ljmd_system* s) {
  MULTI_ALL(s, ljmd_reset_averaging(sub));
  CHECK_S(s);
  DevScalars z;
  memset(&z, 0, sizeof(z));
  const size_t off = offsetof(DevScalars, av_U_tot);
  const size_t len = offsetof(DevScalars, chi) - off;
  CU(cudaMemsetAsync((char*)s->sc + off, 0, len, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->h_sc->av_U_tot = s->h_sc->av_T_tot = s->h_sc->av_p_tot = 0.;
  s->h_sc->av_iters = 0;
  return LJMD_OK;
}

static int fetch_rdf(ljmd_system* s, const unsigned long long* dev, unsigned long long* host256) {
#ifdef LJMD_WITH_NCCL
  if (s->world > 1 && !s->inproc) {   // in-process sub-handles return their own counts: the front adds them
    unsigned long long* tmp = reinterpret_cast<unsigned long long*>(s->gath);
    CU(cudaMemcpyAsync(tmp, dev, kRdfBins * 8, cudaMemcpyDeviceToDevice, s->stream));
    NC(ncclAllReduce(tmp, tmp, kRdfBins, ncclUint64, ncclSum, s->comm, s->stream));
    dev = tmp;
  }
#endif
  CU(cudaMemcpyAsync(host256, dev, kRdfBins * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return LJMD_OK;
}

extern "C" int ljmd_get_rdf(ljmd_system* s, int* out256) {
  if (s && s->multi) {
    if (!out256) return set_err(LJMD_ERR_ARG, "out256 is NULL");
    const int G = multi_count(s);
    std::vector<int> part((size_t)G * kRdfBins);
    const int rc = multi_all(s, [&](ljmd_system* sub, int r) { return ljmd_get_rdf(sub, part.data() + (size_t)r * kRdfBins); });
    if (rc) return rc;
    for (int k = 0; k < kRdfBins; ++k) {
      long long t = 0;
      for (int r = 0; r < G; ++r) t += part[(size_t)r * kRdfBins + k];
      out256[k] = (int)t;
    }
    return LJMD_OK;
  }
  CHECK_S(s);
  if (!out256) return set_err(LJMD_ERR_ARG, "out256 is NULL");
  int rc;
  if (!s->rdf_valid) {
    // lazy: rebuild from the saved evaluation positions (posA/upos); forces are recomputed into
    // scratch (fpart/blockW) and discarded, the gathered force array is untouched.
    if ((rc = launch_force(s, true))) return rc;
  }
  if ((rc = fetch_rdf(s, s->rdf_cur, s->h_rdf))) return rc;
  for (int k = 0; k < kRdfBins; ++k) out256[k] = (int)s->h_rdf[k];
  return LJMD_OK;
}

extern "C" int ljmd_get_rdf_accum(ljmd_system* s, long long* out256, int* nsamples, int reset) {
  if (s && s->multi) {
    if (!out256) return set_err(LJMD_ERR_ARG, "out256 is NULL");
    const int G = multi_count(s);
    std::vector<long long> part((size_t)G * kRdfBins);
    std::vector<int> ns(G, 0);
    const int rc = multi_all(s, [&](ljmd_system* sub, int r) {
      return ljmd_get_rdf_accum(sub, part.data() + (size_t)r * kRdfBins, &ns[r], reset);
    });
    if (rc) return rc;
    for (int k = 0; k < kRdfBins; ++k) {
      out256[k] = 0;
      for (int r = 0; r < G; ++r) out256[k] += part[(size_t)r * kRdfBins + k];
    }
    if (nsamples) *nsamples = ns[0];
    return LJMD_OK;
  }
  CHECK_S(s);
  if (!out256) return set_err(LJMD_ERR_ARG, "out256 is NULL");
  int rc = fetch_rdf(s, s->rdf_acc, s->h_rdf);
  if (rc) return rc;
  for (int k = 0; k < kRdfBins; ++k) out256[k] = (long long)s->h_rdf[k];
  if (nsamples) *nsamples = s->rdf_nacc;
  if (reset) {
    CU(cudaMemsetAsync(s->rdf_acc, 0, kRdfBins * 8, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->rdf_nacc = 0;
  }
  return LJMD_OK;
}

extern "C" int ljmd_velocity_histogram(ljmd_system* s, double step, int nbins, int* out) {
  if (s && s->multi) {
    if (!out || nbins < 1 || nbins > 65536 / 2) return set_err(LJMD_ERR_ARG, "bad histogram arguments");
    const int G = multi_count(s);
    std::vector<int> part((size_t)G * nbins);
    const int rc = multi_all(s, [&](ljmd_system* sub, int r) {
      return ljmd_velocity_histogram(sub, step, nbins, part.data() + (size_t)r * nbins);
    });
    if (rc) return rc;
    for (int k = 0; k < nbins; ++k) {
      out[k] = 0;
      for (int r = 0; r < G; ++r) out[k] += part[(size_t)r * nbins + k];
    }
    return LJMD_OK;
  }
  CHECK_S(s);
  if (!out || nbins < 1 || nbins > 65536 / 2 || !(step > 0.)) return set_err(LJMD_ERR_ARG, "bad histogram arguments");
  CU(cudaMemsetAsync(s->velh, 0, (size_t)nbins * 4, s->stream));
  const int g = std::min(step_grid(s), 4 * s->num_sms);
  // up to 32 768 bins = 128 KB of dynamic shared memory: beyond the 48 KB default a per-function opt-in is needed
  if ((size_t)nbins * 4 > 48 * 1024)
    CU(cudaFuncSetAttribute(k_velhist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)nbins * 4)));
  k_velhist<<<g, kStepThreads, (size_t)nbins * 4, s->stream>>>(s->vel, s->nloc, step, nbins, s->velh);
  CU(cudaGetLastError());
  s->launches += 1;
#ifdef LJMD_WITH_NCCL
  if (s->world > 1 && !s->inproc) NC(ncclAllReduce(s->velh, s->velh, nbins, ncclUint32, ncclSum, s->comm, s->stream));
#endif
  std::vector<unsigned int> h((size_t)nbins);
  CU(cudaMemcpyAsync(h.data(), s->velh, (size_t)nbins * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  for (int k = 0; k < nbins; ++k) out[k] = (int)h[k];
  return LJMD_OK;
}

extern "C" int ljmd_set_l2_flush(ljmd_system* s, long long bytes) {
  MULTI_ALL(s, ljmd_set_l2_flush(sub, bytes));
  CHECK_S(s);
  if (bytes < 0) return set_err(LJMD_ERR_ARG, "bytes must be >= 0");
  if ((size_t)bytes != s->flush_bytes) {
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaFree(s->flush_buf));
    s->flush_buf = nullptr;
    s->flush_bytes = 0;
    if (bytes > 0) {
      CU(cudaMalloc(&s->flush_buf, (size_t)bytes));
      s->flush_bytes = (size_t)bytes;
    }
  }
  return LJMD_OK;
}

// Sub-volume occupancy: the per-step consumers of run-fluctuations read h_Pos / h_Vel only
// to count particles in nested sub-volumes; counting on the device returns ~20 integers instead of 32 B/particle.
// The fraction grid of one counter.  The reference builds it by repeated addition in double: restate the loops
// literally (run-fluctuations-aux.h:198-203 for coordinates, :251-253 for velocities).
static int build_subvol_spec(ljmd_system* s, int type, double alpha_step, double vcut_max, SubvolSpec* q) {
  if (!(alpha_step > 0.) || type < 0 || type > 6) return set_err(LJMD_ERR_ARG, "bad sub-volume arguments");
  memset(q, 0, sizeof(*q));
  q->type = type; q->L = s->L; q->alpha_step = alpha_step; q->vcut_max = vcut_max;
  int nb = 0;
  if (type <= 3) {
    for (double alpha = alpha_step; alpha < 1. - 1.e-9; alpha += alpha_step) {
      if (nb >= kMaxSubBins) return set_err(LJMD_ERR_ARG, "alpha_step too small (more than %d fractions)", kMaxSubBins);
      q->tLs[nb++] = s->L * pow(alpha, 1. / 3.);
    }
  } else {
    if (!(vcut_max > 0.)) return set_err(LJMD_ERR_ARG, "vcut_max must be positive");
    for (double alpha = alpha_step; alpha < 1. + 1.e-9; alpha += alpha_step) {
      if (nb >= kMaxSubBins) return set_err(LJMD_ERR_ARG, "alpha_step too small (more than %d fractions)", kMaxSubBins);
      ++nb;
    }
  }
  q->nbins = nb;
  return LJMD_OK;
}

static int subvolume_impl(ljmd_system* s, int type, double alpha_step, double vcut_max, int* out, int cap, int* nout) {
  if (s && s->multi) {
    if (!out || !nout || cap < 0) return set_err(LJMD_ERR_ARG, "bad sub-volume arguments");
    const int G = multi_count(s);
    std::vector<int> part((size_t)G * std::max(cap, 1), 0), n(G, 0);
    const int rc = multi_all(s, [&](ljmd_system* sub, int r) {
      return subvolume_impl(sub, type, alpha_step, vcut_max, part.data() + (size_t)r * std::max(cap, 1), cap, &n[r]);
    });
    if (rc) return rc;
    *nout = n[0];
    for (int k = 0; k < n[0]; ++k) {   // cumulative counts of disjoint shards add up
      out[k] = 0;
      for (int r = 0; r < G; ++r) out[k] += part[(size_t)r * std::max(cap, 1) + k];
    }
    return LJMD_OK;
  }
  CHECK_S(s);
  if (!out || !nout) return set_err(LJMD_ERR_ARG, "bad sub-volume arguments");
  SubvolParams q;
  int rc = build_subvol_spec(s, type, alpha_step, vcut_max, &q.s);
  if (rc) return rc;
  q.arr = (type <= 3) ? s->pos : s->vel;
  q.n = s->nloc;
  const int nb = q.s.nbins;
  *nout = nb;
  if (nb == 0) return LJMD_OK;
  if (nb > cap) return set_err(LJMD_ERR_ARG, "output holds %d counts, %d needed", cap, nb);
  CU(cudaMemsetAsync(s->velh, 0, (size_t)nb * 4, s->stream));
  const int g = std::max(1, std::min(step_grid(s), 4 * s->num_sms));
  k_subvolume<<<g, kStepThreads, 0, s->stream>>>(q, s->velh);
  CU(cudaGetLastError());
  s->launches += 1;
#ifdef LJMD_WITH_NCCL
  if (s->world > 1 && !s->inproc) NC(ncclAllReduce(s->velh, s->velh, nb, ncclUint32, ncclSum, s->comm, s->stream));
#endif
  unsigned int h[kMaxSubBins];
  CU(cudaMemcpyAsync(h, s->velh, (size_t)nb * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  long long run = 0;
  for (int k = 0; k < nb; ++k) { run += h[k]; out[k] = (int)run; }   // cumulative, as :236-237 / :273-274
  return LJMD_OK;
}

// ---- observation trace ---------------------------------------------------------------------------
static void trace_free(ljmd_system* s) {
  cudaFree(s->trace_specs); cudaFree(s->trace_counts); cudaFree(s->trace_scal); cudaFree(s->trace_idx);
  s->trace_specs = nullptr; s->trace_counts = nullptr; s->trace_scal = nullptr; s->trace_idx = nullptr;
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }   // it holds the old buffers
  s->trace_on = 0; s->trace_cap = s->trace_n = s->trace_row = s->trace_ncounters = 0;
}

extern "C" int ljmd_trace_begin(ljmd_system* s, int ncounters, const int* kinds, const double* alpha_steps,
                                const double* vcut_max, int capacity_steps) {
  MULTI_ALL(s, ljmd_trace_begin(sub, ncounters, kinds, alpha_steps, vcut_max, capacity_steps));
  CHECK_S(s);
  if (ncounters < 0 || ncounters > kMaxTraceCounters) return set_err(LJMD_ERR_ARG, "0..%d counters", kMaxTraceCounters);
  if (capacity_steps < 1) return set_err(LJMD_ERR_ARG, "capacity_steps must be >= 1");
  if (ncounters > 0 && (!kinds || !alpha_steps)) return set_err(LJMD_ERR_ARG, "kinds/alpha_steps must not be NULL");
  trace_free(s);
  std::vector<SubvolSpec> specs((size_t)std::max(1, ncounters));
  int row = 3;
  for (int c = 0; c < ncounters; ++c) {
    const double vc = (kinds[c] >= 4) ? (vcut_max ? vcut_max[c] : 0.) : 1.;
    int rc = build_subvol_spec(s, kinds[c], alpha_steps[c], vc, &specs[c]);
    if (rc) return rc;
    s->trace_nbins[c] = specs[c].nbins;
    row += specs[c].nbins;
  }
  CU(cudaMalloc(&s->trace_specs, specs.size() * sizeof(SubvolSpec)));
  CU(cudaMemcpyAsync(s->trace_specs, specs.data(), specs.size() * sizeof(SubvolSpec), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));   // `specs` is a local
  CU(cudaMalloc(&s->trace_counts, (size_t)capacity_steps * row * sizeof(unsigned long long)));
  CU(cudaMalloc(&s->trace_scal, (size_t)capacity_steps * kTraceScalars * sizeof(double)));
  CU(cudaMemsetAsync(s->trace_counts, 0, (size_t)capacity_steps * row * sizeof(unsigned long long), s->stream));
  CU(cudaMalloc(&s->trace_idx, 2 * sizeof(int)));
  CU(cudaMemsetAsync(s->trace_idx, 0, 2 * sizeof(int), s->stream));
  s->trace_on = 1; s->trace_cap = capacity_steps; s->trace_n = 0; s->trace_row = row; s->trace_ncounters = ncounters;
  return LJMD_OK;
}

extern "C" int ljmd_trace_row_length(ljmd_system* s, int* counts_per_step) {
  if (s && s->multi) return ljmd_trace_row_length(multi_sub(s, 0), counts_per_step);
  CHECK_S(s);
  if (!counts_per_step) return set_err(LJMD_ERR_ARG, "counts_per_step must not be NULL");
  if (!s->trace_on) return set_err(LJMD_ERR_ARG, "no trace is active");
  *counts_per_step = s->trace_row - 3;
  return LJMD_OK;
}

// one row: called in stream order right after a step finished
static int trace_record(ljmd_system* s) {
  TraceParams q;
  q.pos = s->pos; q.vel = s->vel; q.n = s->nloc;
  q.ncounters = s->trace_ncounters; q.row = s->trace_row; q.specs = s->trace_specs; q.sc = s->sc;
  q.counts = s->trace_counts;
  q.scal = s->trace_scal;
  q.row_idx = s->trace_idx;
  q.ticket = reinterpret_cast<unsigned int*>(s->trace_idx + 1);
  const int g = std::max(1, std::min(step_grid(s), 4 * s->num_sms));
  const size_t smem = (size_t)s->trace_ncounters * sizeof(SubvolSpec) + (size_t)(s->trace_row - 3 + 1) * sizeof(unsigned int);
  k_trace<<<g, kStepThreads, smem, s->stream>>>(q);
  CU(cudaGetLastError());
  s->launches += 1;
  s->trace_n += 1;
  return LJMD_OK;
}

// vel_int: [nsteps][3] integer velocity sums in 2^-32 fixed point (this rank's particles when in-process)
static int trace_read_impl(ljmd_system* s, int max_steps, int* nsteps, double* scalars, long long* counts,
                           long long* vel_int) {
  CHECK_S(s);
  if (!nsteps) return set_err(LJMD_ERR_ARG, "nsteps must not be NULL");
  if (!s->trace_on) return set_err(LJMD_ERR_ARG, "no trace is active");
  const int n = s->trace_n, row = s->trace_row, nb = row - 3;
  if (n > max_steps) return set_err(LJMD_ERR_ARG, "trace holds %d steps, output has room for %d", n, max_steps);
  *nsteps = n;
  if (n == 0) return LJMD_OK;
#ifdef LJMD_WITH_NCCL
  if (s->world > 1 && !s->inproc)
    NC(ncclAllReduce(s->trace_counts, s->trace_counts, (size_t)n * row, ncclUint64, ncclSum, s->comm, s->stream));
#endif
  std::vector<unsigned long long> h((size_t)n * row);
  CU(cudaMemcpyAsync(h.data(), s->trace_counts, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
  if (scalars)
    CU(cudaMemcpyAsync(scalars, s->trace_scal, (size_t)n * kTraceScalars * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemsetAsync(s->trace_counts, 0, (size_t)n * row * sizeof(unsigned long long), s->stream));
  CU(cudaMemsetAsync(s->trace_idx, 0, 2 * sizeof(int), s->stream));
  CU(cudaStreamSynchronize(s->stream));
  for (int k = 0; k < n; ++k) {
    const unsigned long long* r = h.data() + (size_t)k * row;
    if (counts) {
      int off = 0;
      for (int c = 0; c < s->trace_ncounters; ++c) {   // cumulative per counter, as :236-237 / :273-274
        long long run = 0;
        for (int b = 0; b < s->trace_nbins[c]; ++b) { run += (long long)r[off + b]; counts[(size_t)k * nb + off + b] = run; }
        off += s->trace_nbins[c];
      }
    }
    if (vel_int)
      for (int a = 0; a < 3; ++a) vel_int[3 * k + a] = (long long)r[nb + a];
  }
  s->trace_n = 0;
  return LJMD_OK;
}

extern "C" int ljmd_trace_read(ljmd_system* s, int max_steps, int* nsteps, double* scalars, long long* counts,
                               double* mean_velocity) {
  if (!s) return set_err(LJMD_ERR_ARG, "system is NULL");
  if (!nsteps || max_steps < 0) return set_err(LJMD_ERR_ARG, "nsteps must not be NULL");
  std::vector<long long> vint(mean_velocity ? (size_t)3 * std::max(max_steps, 1) : 0, 0);
  int rc;
  if (s->multi) {
    // every device returns the rows of its own particles: integer counts and integer velocity sums add up exactly
    const int G = multi_count(s);
    int nb = 0;
    if ((rc = ljmd_trace_row_length(multi_sub(s, 0), &nb))) return rc;
    const size_t cstride = (size_t)std::max(max_steps, 1) * std::max(nb, 1), vstride = (size_t)3 * std::max(max_steps, 1);
    std::vector<long long> cpart(counts ? cstride * G : 0, 0), vpart(mean_velocity ? vstride * G : 0, 0);
    std::vector<int> ns(G, 0);
    rc = multi_all(s, [&](ljmd_system* sub, int r) {
      return trace_read_impl(sub, max_steps, &ns[r], r == 0 ? scalars : nullptr, counts ? cpart.data() + cstride * r : nullptr,
                             mean_velocity ? vpart.data() + vstride * r : nullptr);
    });
    if (rc) return rc;
    *nsteps = ns[0];
    for (int k = 0; k < ns[0]; ++k) {
      if (counts)
        for (int b = 0; b < nb; ++b) {
          long long t = 0;
          for (int r = 0; r < G; ++r) t += cpart[cstride * r + (size_t)k * nb + b];
          counts[(size_t)k * nb + b] = t;
        }
      if (mean_velocity)
        for (int a = 0; a < 3; ++a) {
          long long t = 0;
          for (int r = 0; r < G; ++r) t += vpart[vstride * r + 3 * k + a];
          vint[3 * k + a] = t;
        }
    }
  } else {
    if ((rc = trace_read_impl(s, max_steps, nsteps, scalars, counts, mean_velocity ? vint.data() : nullptr))) return rc;
  }
  if (mean_velocity)
    for (int k = 0; k < *nsteps; ++k)
      for (int a = 0; a < 3; ++a) mean_velocity[3 * k + a] = (double)vint[3 * k + a] / 4294967296.0 / (double)s->N;
  return LJMD_OK;
}

extern "C" int ljmd_trace_end(ljmd_system* s) {
  MULTI_ALL(s, ljmd_trace_end(sub));
  CHECK_S(s);
  trace_free(s);
  return LJMD_OK;
}

extern "C" int ljmd_subvolume_counts(ljmd_system* s, int type, double alpha_step, int* out, int cap, int* nout) {
  if (type < 0 || type > 3) return set_err(LJMD_ERR_ARG, "type must be 0 (x slab), 1 (y), 2 (z) or 3 (cube)");
  return subvolume_impl(s, type, alpha_step, 1., out, cap, nout);
}
extern "C" int ljmd_velocity_subvolume_counts(ljmd_system* s, int type, double vcut_max, double alpha_step, int* out,
                                              int cap, int* nout) {
  if (type < 0 || type > 2) return set_err(LJMD_ERR_ARG, "type must be 0 (vx), 1 (vy) or 2 (vz)");
  return subvolume_impl(s, 4 + type, alpha_step, vcut_max, out, cap, nout);
}

extern "C" long long ljmd_launch_count(ljmd_system* s) {
  if (s && s->multi) {
    long long t = 0;
    for (int r = 0; r < multi_count(s); ++r) t += multi_sub(s, r)->launches;
    return t;
  }
  return s ? s->launches : 0;
}

extern "C" int ljmd_set_event_timing(ljmd_system* s, int on) {
  MULTI_ALL(s, ljmd_set_event_timing(sub, on));
  CHECK_S(s);
  s->timing = on ? 1 : 0;
  return LJMD_OK;
}
extern "C" int ljmd_last_step_timing(ljmd_system* s, double* force_ms, double* total_ms, int* force_launches) {
  if (s && s->multi) {   // the slowest device defines the step
    double f = 0., t = 0.;
    for (int r = 0; r < multi_count(s); ++r) {
      double fr = 0., tr = 0.;
      const int rc = ljmd_last_step_timing(multi_sub(s, r), &fr, &tr, force_launches);
      if (rc) return rc;
      f = std::max(f, fr); t = std::max(t, tr);
    }
    if (force_ms) *force_ms = f;
    if (total_ms) *total_ms = t;
    return LJMD_OK;
  }
  CHECK_S(s);
  if (force_ms) *force_ms = s->last_force_ms;
  // the steps themselves (sum of per-step intervals); the L2-flush writes between them are not step work
  if (total_ms) *total_ms = s->flush_bytes ? s->last_steps_ms : s->last_total_ms;
  if (force_launches) *force_launches = s->last_force_launches;
  return LJMD_OK;
}
// bytes of all reaction blocks of this rank (what one pass over every band reads)
static double reaction_band_bytes(const ljmd_system* s) {
  if (!s->use_sym) return 0.;
  const int hmax = sym_max_partner_count(s->nblk);
  int qmax = s->sym_mi - 1 + hmax;
  if (qmax > s->nblk - 1) qmax = s->nblk - 1;
  return 16. * (double)s->n_super * (qmax + 1) * kITile;
}

// k_reduce_reaction (sharded Newton-3 runs): column sums of this rank's reaction blocks for all N particles —
// reads every band once, writes one record per particle.  The HBM-bound kernel of the sharded step.
extern "C" int ljmd_last_reduce_timing(ljmd_system* s, double* reduce_ms, int* launches, double* bytes_per_launch) {
  if (s && s->multi) {
    double g = 0.;
    for (int r = 0; r < multi_count(s); ++r) {
      double gr = 0.;
      const int rc = ljmd_last_reduce_timing(multi_sub(s, r), &gr, launches, bytes_per_launch);
      if (rc) return rc;
      g = std::max(g, gr);
    }
    if (reduce_ms) *reduce_ms = g;
    return LJMD_OK;
  }
  CHECK_S(s);
  if (reduce_ms) *reduce_ms = s->last_reduce_ms;
  if (launches) *launches = s->last_reduce_launches;
  if (bytes_per_launch) *bytes_per_launch = reaction_band_bytes(s) + 16. * (double)s->npad;
  return LJMD_OK;
}

extern "C" int ljmd_last_gather_timing(ljmd_system* s, double* gather_ms, int* launches, double* bytes_per_launch) {
  if (s && s->multi) {
    double g = 0.;
    for (int r = 0; r < multi_count(s); ++r) {
      double gr = 0.;
      const int rc = ljmd_last_gather_timing(multi_sub(s, r), &gr, launches, bytes_per_launch);   // bytes: per device
      if (rc) return rc;
      g = std::max(g, gr);
    }
    if (gather_ms) *gather_ms = g;
    return LJMD_OK;
  }
  CHECK_S(s);
  if (gather_ms) *gather_ms = s->last_gather_ms;
  if (launches) *launches = s->last_gather_launches;
  if (bytes_per_launch) {
    // algorithmic bytes of one k_gather launch: per local particle the nsplit partial-force rows,
    // velocity and old force in; force plus t_Force (TVN) or velocity and position out; plus the reaction
    // records: on one GPU every super-tile's band (each entry is read exactly once), sharded one pre-reduced
    // record per rank (pulled from the peers' windows over NVLink: not HBM traffic of this device)
    *bytes_per_launch = (16. * s->nsplit + 32. + 48.) * (double)s->nloc + reaction_band_bytes(s) * (s->world == 1 ? 1. : 0.) +
                        (s->use_sym && s->world > 1 ? 16. * s->world * (double)s->nloc : 0.);
  }
  return LJMD_OK;
}

extern "C" int ljmd_get_launch_info(ljmd_system* s, int* out8) {
  if (s && s->multi) return ljmd_get_launch_info(multi_sub(s, 0), out8);
  CHECK_S(s);
  if (!out8) return set_err(LJMD_ERR_ARG, "out8 is NULL");
  out8[0] = s->num_sms; out8[1] = kITile; out8[2] = s->nsplit; out8[3] = (s->use_sym ? s->n_super : s->n_itiles) * s->nsplit;
  out8[4] = s->world; out8[5] = s->nloc; out8[6] = s->use_sym; out8[7] = s->use_sym ? s->sym_bj : kTileJ;
  return LJMD_OK;
}

// ------------------------------------------------------------------------------------ C ABI: B
// Worker for the legacy seam (ljmd_legacy.cu): forces on caller-owned device arrays.
int ljmd_legacy_forces(ljmd_system* s, const float* d_pos, float* d_force, float* pressure, int* rdf256) {
  if (s && s->multi) return set_err(LJMD_ERR_ARG, "the legacy seam is single-device");
  CHECK_S(s);
  // The caller's copyArrayToDevice (a pageable-memory cudaMemcpy on the legacy default stream) may return before
  // its DMA has landed, and this handle's stream is non-blocking: order explicitly after the legacy stream.
  CU(cudaStreamSynchronize(cudaStreamLegacy));
  CU(cudaMemcpyAsync(s->pos, d_pos, (size_t)s->N * 16, cudaMemcpyDeviceToDevice, s->stream));
  int rc = maybe_sort_records(s, false);   // the caller integrates on its side: one call = one step
  if (rc) return rc;
  s->steps_since_sort += 1;
  StepParams p = make_step_params(s, 0.);
  k_prepare<<<step_grid(s), kStepThreads, 0, s->stream>>>(p);
  CU(cudaGetLastError());
  s->launches += 1;
  if ((rc = evaluate(s, p, GATHER_EVAL, rdf256 != nullptr, 0))) return rc;
  CU(cudaMemcpyAsync(d_force, s->force, (size_t)s->N * 16, cudaMemcpyDeviceToDevice, s->stream));
  if ((rc = sync_scalars(s))) return rc;
  if (pressure) *pressure = (float)s->h_sc->Pvirial;
  if (rdf256) {
    if ((rc = fetch_rdf(s, s->rdf_cur, s->h_rdf))) return rc;
    for (int k = 0; k < kRdfBins; ++k) rdf256[k] = (int)s->h_rdf[k];
  }
  return LJMD_OK;
}

// ------------------------------------------------------------------- device-side initial conditions (§8 f-4)
// Three phases with two global integer sums in between (total momentum, then kinetic temperature); `sums` are this
// handle's particles only — sharded systems add them up (NCCL when one process per GPU, on the host behind a front).
static int init_phase(ljmd_system* s, int phase, unsigned long long seed, const long long* total4, long long* own4) {
  CHECK_S(s);
  InitParams q;
  memset(&q, 0, sizeof(q));
  q.pos = s->pos; q.vel = s->vel; q.nloc = s->nloc; q.i_begin = s->i_begin; q.N = s->N; q.L = s->L; q.T0 = s->T0; q.seed = seed;
  q.Ns = (int)ceil(pow((double)s->N, 1. / 3.));   // MDSystem.cpp:150
  q.dL = s->L / q.Ns;                             // :151
  long long* dsum = reinterpret_cast<long long*>(s->velh);   // scratch: 4 x int64
  q.sums = dsum;
  const int g = step_grid(s);
  if (phase == 1) {
    CU(cudaMemsetAsync(dsum, 0, 4 * sizeof(long long), s->stream));
    k_init_sample<<<g, kStepThreads, 0, s->stream>>>(q);
  } else if (phase == 2) {
    for (int a = 0; a < 3; ++a) q.mean[a] = (double)total4[a] / 4294967296.0 / (double)s->N;
    k_init_center<<<g, kStepThreads, 0, s->stream>>>(q);
  } else {
    const double Tkin = (double)total4[3] / 4294967296.0 * (1. / 3. / (double)s->N);   // MDSystem.cpp:370
    q.factor = sqrt(s->T0 / Tkin);                                                      // :382
    k_init_scale<<<g, kStepThreads, 0, s->stream>>>(q);
  }
  CU(cudaGetLastError());
  s->launches += 1;
  if (own4) {
#ifdef LJMD_WITH_NCCL
    if (s->world > 1 && !s->inproc) NC(ncclAllReduce(dsum, dsum, 4, ncclInt64, ncclSum, s->comm, s->stream));
#endif
    CU(cudaMemcpyAsync(own4, dsum, 4 * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }
  return LJMD_OK;
}

// after the three phases: what ljmd_set_state does once the arrays are on the device
static int init_finish(ljmd_system* s) {
  CHECK_S(s);
  int rc = maybe_sort_records(s, true);
  if (rc) return rc;
  StepParams p = make_step_params(s, 0.);
  k_prepare<<<step_grid(s), kStepThreads, 0, s->stream>>>(p);
  CU(cudaGetLastError());
  s->launches += 1;
  if ((rc = allgather_positions(s))) return rc;
  CU(cudaMemsetAsync(s->sc, 0, sizeof(DevScalars), s->stream));
  if ((rc = evaluate(s, p, GATHER_EVAL, false, 0))) return rc;
  s->rdf_nacc = 0;
  CU(cudaMemsetAsync(s->rdf_acc, 0, kRdfBins * sizeof(unsigned long long), s->stream));
  return sync_scalars(s);
}

extern "C" int ljmd_init_state(ljmd_system* s, unsigned long long seed) {
  if (!s) return set_err(LJMD_ERR_ARG, "system is NULL");
  long long tot[4] = {0, 0, 0, 0};
  if (s->multi) {
    const int G = multi_count(s);
    std::vector<long long> part((size_t)4 * G, 0);
    auto add_up = [&]() { for (int k = 0; k < 4; ++k) { tot[k] = 0; for (int r = 0; r < G; ++r) tot[k] += part[4 * r + k]; } };
    int rc = multi_all(s, [&](ljmd_system* sub, int r) { return init_phase(sub, 1, seed, nullptr, &part[4 * r]); });
    if (rc) return rc;
    add_up();
    if ((rc = multi_all(s, [&](ljmd_system* sub, int r) { return init_phase(sub, 2, seed, tot, &part[4 * r]); }))) return rc;
    add_up();
    if ((rc = multi_all(s, [&](ljmd_system* sub, int) { return init_phase(sub, 3, seed, tot, nullptr); }))) return rc;
    return multi_all(s, [&](ljmd_system* sub, int) { return init_finish(sub); });
  }
  int rc = init_phase(s, 1, seed, nullptr, tot);
  if (rc) return rc;
  if ((rc = init_phase(s, 2, seed, tot, tot))) return rc;
  if ((rc = init_phase(s, 3, seed, tot, nullptr))) return rc;
  return init_finish(s);
}

// ------------------------------------------------------------------- shear stress, on demand
// Pshear = (4/2 * sum_i sum_{j != i} (-r_x,ij * f_y,ij / 4) + sum_i (-v_x v_y)) / (N / rho), MDSystem.cpp:299,309,335,353.
// The reference computes it on its CPU path only and nothing reads it (its GPU path leaves the member stale), so
// it is kept OUT of the force kernels' hot loop: one more accumulator there costs ~4 % of every step.  This kernel
// is a plain ordered all-pairs pass over the saved evaluation positions: one i per thread, j-records staged in
// shared memory, float products, double accumulation per thread.
template <bool PERIODIC>
__global__ void __launch_bounds__(256) k_shear(const uint4* __restrict__ jrec, const float4* __restrict__ vel, int N,
                                               int i_begin, int i_end, float c2, double kunit, double* __restrict__ out2,
                                               const int* __restrict__ slot) {
  __shared__ uint4 tile[256];
  __shared__ double red[2][8];
  const int ip = i_begin + blockIdx.x * 256 + threadIdx.x;   // particle (the velocity's index)
  const bool live = ip < i_end;
  const int i = live ? (slot ? slot[ip - i_begin] : ip) : i_begin;   // its record
  const uint4 me = jrec[i];
  double acc = 0.;
  for (int j0 = 0; j0 < N; j0 += 256) {
    __syncthreads();
    tile[threadIdx.x] = jrec[min(j0 + (int)threadIdx.x, N - 1)];
    __syncthreads();
    const int nj = min(256, N - j0);
    float part = 0.f;
    for (int j = 0; j < nj; ++j) {
      const uint4 u = tile[j];
      float dx, dy, dz;
      if (PERIODIC) {
        dx = __int2float_rn((int)me.x - (int)u.x); dy = __int2float_rn((int)me.y - (int)u.y); dz = __int2float_rn((int)me.z - (int)u.z);
      } else {
        dx = __uint_as_float(me.x) - __uint_as_float(u.x); dy = __uint_as_float(me.y) - __uint_as_float(u.y);
        dz = __uint_as_float(me.z) - __uint_as_float(u.z);
      }
      const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      float x = (j0 + j == i) ? 0.f : rcp_approx(r2);
      if (PERIODIC) x *= c2;
      const float r6 = x * x * x;
      const float sfac = r6 * fmaf(r6, 12.f, -6.f) * x;   // r^-2 (12 r^-12 - 6 r^-6)
      part = fmaf(-dx * dy, sfac, part);                   // -r_x * f_y / 4 (coordinates in k-units when periodic)
    }
    acc += (double)part;
  }
  if (!live) acc = 0.;
  acc *= kunit * kunit;                                    // k-units^2 -> sigma^2 (1 for open boxes)
  double kin = 0.;
  if (live) { const float4 v = vel[ip - i_begin]; kin = (double)(-v.x * v.y); }   // :335 float product, double sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); kin += __shfl_xor_sync(0xffffffffu, kin, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc; red[1][threadIdx.x >> 5] = kin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0., k = 0.;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; k += red[1][w]; }
    out2[2 * blockIdx.x] = a;
    out2[2 * blockIdx.x + 1] = k;
  }
}

// conf: sum_i sum_{j != i} (-r_x f_y) * 4/2 over this handle's i-particles; kin: sum_i (-v_x v_y)
static int shear_parts(ljmd_system* s, double* conf, double* kin) {
  CHECK_S(s);
  const bool periodic = (s->bc == LJMD_BC_PERIODIC);
  const int nb = (s->nloc + 255) / 256;
  double* d = nullptr;
  CU(cudaMalloc(&d, (size_t)2 * nb * sizeof(double)));
  const double k2 = 4294967296.0 / s->L;
  const uint4* jrec = periodic ? s->upos : reinterpret_cast<const uint4*>(s->posA);
  if (periodic) k_shear<true><<<nb, 256, 0, s->stream>>>(jrec, s->vel, s->N, s->i_begin, s->i_end, (float)(k2 * k2), 1. / k2, d, s->slot);
  else k_shear<false><<<nb, 256, 0, s->stream>>>(jrec, s->vel, s->N, s->i_begin, s->i_end, 1.f, 1., d, s->slot);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { cudaFree(d); return set_err(LJMD_ERR_CUDA, "k_shear launch: %s", cudaGetErrorString(e)); }
  s->launches += 1;
  std::vector<double> h((size_t)2 * nb);
  e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  if (e != cudaSuccess) return set_err(LJMD_ERR_CUDA, "k_shear: %s", cudaGetErrorString(e));
  double a = 0., k = 0.;
  for (int b = 0; b < nb; ++b) { a += h[2 * b]; k += h[2 * b + 1]; }   // fixed order: deterministic
  *conf = a * (4. / 2.);                                                // :309
  *kin = k;
  return LJMD_OK;
}

extern "C" int ljmd_get_pshear(ljmd_system* s, double* pshear) {
  if (!s || !pshear) return set_err(LJMD_ERR_ARG, "NULL argument");
  double conf = 0., kin = 0.;
  if (s->multi) {
    const int G = multi_count(s);
    std::vector<double> c(G, 0.), k(G, 0.);
    const int rc = multi_all(s, [&](ljmd_system* sub, int r) { return shear_parts(sub, &c[r], &k[r]); });
    if (rc) return rc;
    for (int r = 0; r < G; ++r) { conf += c[r]; kin += k[r]; }
  } else {
    int rc = shear_parts(s, &conf, &kin);
    if (rc) return rc;
#ifdef LJMD_WITH_NCCL
    if (s->world > 1) {   // rare read-out: host-staged all-reduce of two doubles through the scratch buffer
      double two[2] = {conf, kin};
      CU(cudaMemcpyAsync(s->gath, two, sizeof(two), cudaMemcpyHostToDevice, s->stream));
      NC(ncclAllReduce(s->gath, s->gath, 2, ncclDouble, ncclSum, s->comm, s->stream));
      CU(cudaMemcpyAsync(two, s->gath, sizeof(two), cudaMemcpyDeviceToHost, s->stream));
      CU(cudaStreamSynchronize(s->stream));
      conf = two[0]; kin = two[1];
    }
#endif
  }
  *pshear = (conf + kin) / ((double)s->N / s->rho);   // :353
  return LJMD_OK;
}

// ------------------------------------------------------------------- device pointers
// The arrays a renderer or a downstream CUDA consumer reads every frame, without the D2H copy the reference's GUI
// pays (MDSystemGL.cpp:150-151 hands h_Pos to glVertexPointer; its registerGLBufferObject hooks are stubs,
// MDSystem.cu:199-226).  A GL client maps its own buffer with cudaGraphicsGLRegisterBuffer and copies device to
// device from these pointers on the stream returned here, or reads them in its own kernels.  Single-device
// handles only (a sharded system has no one device holding the velocities).  Valid until ljmd_destroy; positions
// are float4 (x,y,z,w = L/150), exactly what h_Pos holds after Integrate.
extern "C" int ljmd_device_arrays(ljmd_system* s, const void** pos4, const void** vel4, const void** force4, void** stream) {
  if (s && s->multi) return set_err(LJMD_ERR_ARG, "device arrays are exported by single-device handles only");
  CHECK_S(s);
  if (s->world > 1) return set_err(LJMD_ERR_ARG, "device arrays are exported by single-device handles only");
  if (pos4) *pos4 = s->pos;
  if (vel4) *vel4 = s->vel;
  if (force4) *force4 = s->force;
  if (stream) *stream = (void*)s->stream;
  return LJMD_OK;
}

// ------------------------------------------------------------------- FP32 peak probe (roofline denominator)
// A stream of independent packed FMAs (FFMA2, 8 chains per thread, 16 warps per SM sub-partition): the FP32
// CUDA-core rate this device sustains, measured with CUDA events.  bench.py reports the force kernel against this
// figure next to the nominal SMs x 128 lanes x 2 x clock (MEASURED_PEAKS.json carries no FP32 CUDA-core number).
__global__ void __launch_bounds__(256) k_fp32_probe(float* out, int iters, float seed) {
  u64 v[8], ca, cb;
  asm volatile("mov.b64 %0, {%1,%1};" : "=l"(ca) : "f"(seed * 1.0001f));
  asm volatile("mov.b64 %0, {%1,%1};" : "=l"(cb) : "f"(seed * 0.0001f));
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float lo = seed + c + threadIdx.x, hi = seed - c;
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v[c]) : "f"(lo), "f"(hi));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(ca), "l"(cb));
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float lo, hi;
    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int ljmd_fp32_peak_probe(int device, double* tflops) {
  if (!tflops) return set_err(LJMD_ERR_ARG, "tflops is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(LJMD_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return set_err(LJMD_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 8192;
  float* d = nullptr;
  CU(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  k_fp32_probe<<<blocks, threads>>>(d, 64, 1.0f);   // warm-up
  double best = 0.;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0));
    k_fp32_probe<<<blocks, threads>>>(d, iters, 1.0f);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)blocks * threads * iters * 8 /*chains*/ * 2 /*lanes*/ * 2 /*fma*/;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return LJMD_OK;
}

// ------------------------------------------------------------------- single-process multi-GPU (front handle)
// ljmd_create_multi: the caller keeps ONE handle and ONE thread; the library drives G devices of the node.
// Each device gets a sub-handle (rank r of world G) and a persistent worker thread that owns every CUDA call on
// it; a call on the front handle is forwarded to all workers and returns when the slowest is done.  The per-step
// exchange is the fabric of ljmd_step.cuh over plain peer pointers (cudaDeviceEnablePeerAccess: no IPC, no NCCL,
// no launcher); read-outs are combined on the host from the devices' own integer counts.
struct MultiCtl {
  int n = 0;
  std::vector<ljmd_system*> sub;
  std::vector<int> dev;
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  const std::function<int(ljmd_system*, int)>* job = nullptr;
  unsigned long long gen = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;
  std::vector<std::string> err;
};

static void multi_worker(MultiCtl* c, int r) {
  cudaSetDevice(c->dev[r]);
  unsigned long long seen = 0;
  for (;;) {
    const std::function<int(ljmd_system*, int)>* job;
    {
      std::unique_lock<std::mutex> lk(c->m);
      c->cv_go.wait(lk, [&] { return c->stop || c->gen != seen; });
      if (c->stop) return;
      seen = c->gen;
      job = c->job;
    }
    g_err[0] = 0;
    const int rc = (*job)(c->sub[r], r);
    {
      std::lock_guard<std::mutex> lk(c->m);
      c->rc[r] = rc;
      c->err[r] = rc ? g_err : "";
      if (--c->pending == 0) c->cv_done.notify_all();
    }
  }
}

static int multi_run(MultiCtl* c, const std::function<int(ljmd_system*, int)>& f) {
  {
    std::unique_lock<std::mutex> lk(c->m);
    c->job = &f;
    c->pending = c->n;
    c->gen += 1;
    c->cv_go.notify_all();
    c->cv_done.wait(lk, [&] { return c->pending == 0; });
    c->job = nullptr;
  }
  for (int r = 0; r < c->n; ++r)
    if (c->rc[r]) return set_err(c->rc[r], "device %d (rank %d of %d): %s", c->dev[r], r, c->n, c->err[r].c_str());
  return LJMD_OK;
}

static int multi_all(ljmd_system* front, const std::function<int(ljmd_system*, int)>& f) { return multi_run(front->multi, f); }
static ljmd_system* multi_sub(ljmd_system* front, int r) { return front->multi->sub[r]; }
static int multi_count(const ljmd_system* front) { return front->multi->n; }

static int multi_destroy(ljmd_system* front) {
  MultiCtl* c = front->multi;
  multi_run(c, [&](ljmd_system* sub, int r) {
    const int rc = destroy_impl(sub);
    c->sub[r] = nullptr;
    return rc;
  });
  {
    std::lock_guard<std::mutex> lk(c->m);
    c->stop = true;
    c->cv_go.notify_all();
  }
  for (std::thread& t : c->workers) t.join();
  delete c;
  delete front;
  return LJMD_OK;
}

extern "C" int ljmd_create_multi(ljmd_system** out, int N, double rho, double T0, int canonical, int bc, float rdf_dr2,
                                 const int* devices, int ndev) {
  if (!out) return set_err(LJMD_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (!(rho > 0.)) return set_err(LJMD_ERR_ARG, "rho must be positive");
  if (!devices || ndev < 1 || ndev > kMaxPeers) return set_err(LJMD_ERR_ARG, "1..%d devices", kMaxPeers);
  if (ndev == 1) return create_impl(out, N, rho, T0, canonical, bc, rdf_dr2, devices[0], 0, 1, nullptr);
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) {
    cudaGetLastError();
    return set_err(LJMD_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  // LJMD_SHARE_DEVICES=1: a device may be listed more than once — several ranks of the sharded system then run on
  // one GPU, each on its own stream (a diagnostic mode: the whole multi-rank data path — shard plan, windows,
  // barriers, reaction exchange — on a box with fewer GPUs than ranks; no speed-up, the ranks share the SMs).
  const char* share_env = std::getenv("LJMD_SHARE_DEVICES");
  const bool share = share_env && share_env[0] == '1';
  for (int a = 0; a < ndev; ++a) {
    if (devices[a] < 0 || devices[a] >= visible) return set_err(LJMD_ERR_ARG, "device %d out of range (%d visible)", devices[a], visible);
    for (int b = 0; b < a; ++b)
      if (devices[a] == devices[b] && !share)
        return set_err(LJMD_ERR_ARG, "device %d listed twice (LJMD_SHARE_DEVICES=1 allows it)", devices[a]);
  }
  bool shared_any = false;
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < a; ++b) shared_any = shared_any || devices[a] == devices[b];
  if (shared_any) {
    // ranks on one device wait for each other inside kernels: a lazily loaded kernel's first launch would
    // synchronise the context behind a peer's spinning barrier and the step would end in the barrier's watchdog
    const char* ml = std::getenv("CUDA_MODULE_LOADING");
    if (!ml || strcmp(ml, "EAGER") != 0)
      return set_err(LJMD_ERR_ARG, "ranks sharing a device need CUDA_MODULE_LOADING=EAGER in the environment before "
                                   "CUDA starts (and CUDA_DEVICE_MAX_CONNECTIONS >= %d)", ndev);
  }
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < ndev; ++b) {
      int ok = 1;
      if (devices[a] != devices[b]) CU(cudaDeviceCanAccessPeer(&ok, devices[a], devices[b]));
      if (!ok) return set_err(LJMD_ERR_CUDA, "device %d cannot map the memory of device %d (no NVLink / PCIe peer access)", devices[a], devices[b]);
    }
  MultiCtl* c = new (std::nothrow) MultiCtl();
  ljmd_system* front = new (std::nothrow) ljmd_system();
  if (!c || !front) { delete c; delete front; return set_err(LJMD_ERR_ARG, "out of host memory"); }
  c->n = ndev;
  c->sub.assign(ndev, nullptr);
  c->dev.assign(devices, devices + ndev);
  c->rc.assign(ndev, 0);
  c->err.assign(ndev, "");
  for (int r = 0; r < ndev; ++r) c->workers.emplace_back(multi_worker, c, r);
  front->multi = c;
  front->N = N; front->bc = bc; front->canonical = canonical ? 1 : 0; front->T0 = T0; front->rho = rho;
  front->L = pow(N / rho, 1. / 3.); front->dr2 = rdf_dr2; front->world = ndev; front->device = devices[0];
  int rc = multi_run(c, [&](ljmd_system*, int r) {
    return create_impl(&c->sub[r], N, rho, T0, canonical, bc, rdf_dr2, c->dev[r], r, ndev, nullptr, /*inproc=*/1);
  });
  if (rc == LJMD_OK) {
    // wire the fabric: every device maps every peer, the windows are plain pointers in this address space
    rc = multi_run(c, [&](ljmd_system* sub, int r) {
      for (int q = 0; q < ndev; ++q) {
        if (c->dev[q] == c->dev[r]) continue;   // itself, or a rank sharing this device
        const cudaError_t e = cudaDeviceEnablePeerAccess(c->dev[q], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return set_err(LJMD_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", c->dev[q], cudaGetErrorString(e));
        cudaGetLastError();
      }
      for (int q = 0; q < ndev; ++q) sub->fab.base[q] = c->sub[q]->win;
      sub->fab.me = r;
      sub->fab.n = ndev;
      return (int)LJMD_OK;
    });
  }
  if (rc != LJMD_OK) {
    std::string keep = g_err;
    multi_destroy(front);
    return set_err(rc, "%s", keep.c_str());
  }
  *out = front;
  return LJMD_OK;
}
