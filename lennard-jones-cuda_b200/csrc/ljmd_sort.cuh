// Record order for periodic Newton-3 runs: a rank's particles sorted along a Hilbert curve through a grid of
// ~2 cells per particle, so that consecutive records are neighbours in space (ljmd_force_sym.cuh, "Warp frames";
// the RDF box pruning profits too).  Only the RECORDS (posA / upos and the force kernel's row entries) follow this
// order, through StepParams::slot; the state arrays keep the caller's particle order.
//
// A counting sort, hand-written and deterministic: (1) cell key per particle + cell populations (integer atomics:
// order-free), (2) exclusive scan of the populations, (3) scatter into the cells (the order INSIDE a cell comes out
// of an atomic and is arbitrary), (4) one thread per cell sorts its few entries by particle index — so the final
// order is a pure function of the positions — and writes slot[].  O(N), a handful of launches every few hundred
// steps; no library sort.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ljmd_hilbert.cuh"

namespace ljmd {

struct SortParams {
  const float4* pos;     // [nloc] wrapped positions (what h_Pos shows)
  int nloc, i_begin;
  int bits;              // 2^bits cells per axis
  int ncell;             // 8^bits
  double fix_scale;      // 2^32 / L
  unsigned int* key;     // [nloc] cell of every particle
  unsigned int* count;   // [ncell] populations; counted down to zero again by the scatter
  unsigned int* offs;    // [ncell + 1] first entry of every cell
  unsigned int* bsum;    // [blocks of the scan]
  int* order;            // [nloc] particles in cell order
  int* slot;             // [nloc] out: global record index of every local particle
};

constexpr int kSortThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanBlock = kSortThreads * kScanPer;
constexpr int kSortCellCap = 64;   // a cell with more entries keeps its arbitrary (still valid) order

__global__ void __launch_bounds__(kSortThreads) k_sort_keys(const SortParams q) {
  const int il = blockIdx.x * kSortThreads + threadIdx.x;
  if (il >= q.nloc) return;
  const float4 x = q.pos[il];
  // the 32-bit box fraction wraps like the records do: positions boxes away land in their image's cell
  const int sh = 32 - q.bits;
  const uint32_t cx = (uint32_t)(unsigned long long)__double2ll_rn(__dmul_rn((double)x.x, q.fix_scale)) >> sh;
  const uint32_t cy = (uint32_t)(unsigned long long)__double2ll_rn(__dmul_rn((double)x.y, q.fix_scale)) >> sh;
  const uint32_t cz = (uint32_t)(unsigned long long)__double2ll_rn(__dmul_rn((double)x.z, q.fix_scale)) >> sh;
  const unsigned int k = hilbert3(cx, cy, cz, q.bits);
  q.key[il] = k;
  atomicAdd(&q.count[k], 1u);
}

// exclusive scan of 256 per-thread values; every thread also gets the block total
__device__ __forceinline__ unsigned int block_scan_256(unsigned int v, unsigned int* sh, unsigned int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[w] = x;
  __syncthreads();
  unsigned int base = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < kSortThreads / 32; ++k) {
    const unsigned int t = sh[k];
    if (k < w) base += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return base + x - v;
}

// phase 1: scan inside blocks of kScanBlock cells, block totals aside
__global__ void __launch_bounds__(kSortThreads) k_scan_local(const unsigned int* __restrict__ in, unsigned int* __restrict__ out,
                                                             unsigned int* __restrict__ bsum, int n) {
  __shared__ unsigned int sh[kSortThreads / 32];
  const int base = blockIdx.x * kScanBlock + threadIdx.x * kScanPer;
  unsigned int v[kScanPer], s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) { v[k] = (base + k < n) ? in[base + k] : 0u; s += v[k]; }
  unsigned int total;
  unsigned int run = block_scan_256(s, sh, total);
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}
// phase 2: one block turns the block totals into exclusive prefixes
__global__ void __launch_bounds__(kSortThreads) k_scan_top(unsigned int* __restrict__ bsum, int nb) {
  __shared__ unsigned int sh[kSortThreads / 32];
  const int per = (nb + kSortThreads - 1) / kSortThreads;
  const int b0 = threadIdx.x * per;
  unsigned int s = 0;
  for (int k = 0; k < per; ++k) s += (b0 + k < nb) ? bsum[b0 + k] : 0u;
  unsigned int total;
  unsigned int run = block_scan_256(s, sh, total);
  for (int k = 0; k < per; ++k) {
    if (b0 + k < nb) { const unsigned int t = bsum[b0 + k]; bsum[b0 + k] = run; run += t; }
  }
}
// phase 3: add the prefixes; the entry one past the last cell closes the last cell
__global__ void __launch_bounds__(kSortThreads) k_scan_add(unsigned int* __restrict__ out, const unsigned int* __restrict__ bsum,
                                                           int n, unsigned int total) {
  const int i = blockIdx.x * kSortThreads + threadIdx.x;
  if (i < n) out[i] += bsum[i / kScanBlock];
  if (i == 0) out[n] = total;
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const SortParams q) {
  const int il = blockIdx.x * kSortThreads + threadIdx.x;
  if (il >= q.nloc) return;
  const unsigned int k = q.key[il];
  const unsigned int r = atomicSub(&q.count[k], 1u) - 1u;   // counts the cell back down to zero: ready for the next sort
  q.order[q.offs[k] + r] = il;
}

// one thread per cell: ascending particle index inside the cell (deterministic), then the slots
__global__ void __launch_bounds__(kSortThreads) k_sort_slots(const SortParams q) {
  const int c = blockIdx.x * kSortThreads + threadIdx.x;
  if (c >= q.ncell) return;
  const int s = (int)q.offs[c], e = (int)q.offs[c + 1];
  if (e - s > 1 && e - s <= kSortCellCap) {
    for (int a = s + 1; a < e; ++a) {
      const int v = q.order[a];
      int b = a - 1;
      while (b >= s && q.order[b] > v) { q.order[b + 1] = q.order[b]; --b; }
      q.order[b + 1] = v;
    }
  }
  for (int a = s; a < e; ++a) q.slot[q.order[a]] = q.i_begin + a;
}

__global__ void __launch_bounds__(kSortThreads) k_slot_identity(int* slot, int nloc, int i_begin) {
  const int il = blockIdx.x * kSortThreads + threadIdx.x;
  if (il < nloc) slot[il] = i_begin + il;
}

}  // namespace ljmd
