// Section B of include/ljmd.h: the `extern "C"` seam the reference host layer links against
// (/root/reference/src/library/MDSystem.cpp:9-25; definitions replaced: MDSystem.cu:155-299).
// The unmodified reference MDSystem.cpp, compiled with -DUSE_CUDA_TOOLKIT, resolves these
// symbols here, so its GPU branch (MDSystem.cpp:240-251) runs the sm_100a force kernel.
//
// Differences from MDSystem.cu, on purpose:
//  * no per-call cudaMalloc/cudaFree or constant-symbol copies: a cached handle keyed on
//    (numBodies, L, periodic, dr2) owns all scratch;
//  * host_RDF carries the CPU-path semantics (MDSystem.cpp:279-285) — far pairs dropped instead of
//    clamped into bin 255 (which overflows int32 for N >~ 46 341 on the reference GPU path);
//  * errors print to stderr and zero the outputs instead of exit()-ing the process.
#include "../../include/ljmd.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <vector>

int ljmd_create_with_L(ljmd_system** out, int N, double L, int bc, float rdf_dr2, int device);
int ljmd_legacy_forces(ljmd_system* s, const float* d_pos, float* d_force, float* pressure, int* rdf256);

namespace {
struct Cache {
  ljmd_system* sys = nullptr;
  int N = 0, periodic = -1;
  float L = 0.f, dr2 = 0.f;
};
Cache g_cache;

void report(const char* what) { fprintf(stderr, "ljmd legacy seam: %s: %s\n", what, ljmd_last_error()); }
void report_cuda(const char* what, cudaError_t e) {
  fprintf(stderr, "ljmd legacy seam: %s: %s\n", what, cudaGetErrorString(e));
}
}  // namespace

extern "C" {

void allocateArray(float** dest, int number) {   // MDSystem.cu:167-173: 4 floats per body
  *dest = nullptr;
  cudaError_t e = cudaMalloc((void**)dest, sizeof(float) * 4 * (size_t)number);
  if (e != cudaSuccess) report_cuda("allocateArray", e);
}

void deleteArray(float* arr) {                   // MDSystem.cu:181-184
  cudaError_t e = cudaFree(arr);
  if (e != cudaSuccess) report_cuda("deleteArray", e);
}

void allocateNBodyArrays(float* vel[2], int numBodies) {   // MDSystem.cu:157-165
  allocateArray(&vel[0], numBodies);
  allocateArray(&vel[1], numBodies);
}
void deleteNBodyArrays(float* vel[2]) {                     // MDSystem.cu:175-179
  deleteArray(vel[0]);
  deleteArray(vel[1]);
}

void copyArrayToDevice(float* device, const float* host, int numBodies) {   // MDSystem.cu:212-216
  cudaError_t e = cudaMemcpy(device, host, (size_t)numBodies * 4 * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) report_cuda("copyArrayToDevice", e);
}

void copyArrayFromDevice(float* host, const float* device, unsigned int pbo, int numBodies) {   // :199-210
  (void)pbo;
  cudaError_t e = cudaMemcpy(host, device, (size_t)numBodies * 4 * sizeof(float), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) report_cuda("copyArrayFromDevice", e);
}

void registerGLBufferObject(unsigned int pbo) { (void)pbo; }     // stubs in the reference too (:218-226)
void unregisterGLBufferObject(unsigned int pbo) { (void)pbo; }
void threadSync(void) { cudaDeviceSynchronize(); }                // :228

// MDSystem.cu:230-291.  Pos/Force: device float4[numBodies]; Force.w = per-particle
// sum_j (r^-12 - r^-6) (consumed at MDSystem.cpp:340-346); *host_pressure = (4/3/2) * sum r.f/4
// (MDSystem.cu:108,136, consumed at MDSystem.cpp:248); host_RDF[256].  p (block size) and q are
// accepted and ignored: the launch shape is the library's business.
void calculateNForces(float* Pos, float* Force, float* host_pressure, int numBodies, float host_L, int Lperiodic,
                      int* host_RDF, float host_dr2, int p, int q) {
  (void)p; (void)q;
  if (host_pressure) *host_pressure = 0.f;
  if (host_RDF) memset(host_RDF, 0, LJMD_RDF_BINS * sizeof(int));
  Cache& c = g_cache;
  if (!c.sys || c.N != numBodies || c.L != host_L || c.periodic != (Lperiodic != 0) || c.dr2 != host_dr2) {
    if (c.sys) ljmd_destroy(c.sys);
    c.sys = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (ljmd_create_with_L(&c.sys, numBodies, (double)host_L, Lperiodic ? LJMD_BC_PERIODIC : LJMD_BC_NONE, host_dr2,
                           dev) != LJMD_OK) {
      report("create");
      return;
    }
    c.N = numBodies; c.L = host_L; c.periodic = (Lperiodic != 0); c.dr2 = host_dr2;
  }
  if (ljmd_legacy_forces(c.sys, Pos, Force, host_pressure, host_RDF) != LJMD_OK) report("calculateNForces");
}

void threadExit(void) {                           // MDSystem.cu:294-297 (cudaThreadExit)
  if (g_cache.sys) ljmd_destroy(g_cache.sys);
  g_cache = Cache();
}

}  // extern "C"
