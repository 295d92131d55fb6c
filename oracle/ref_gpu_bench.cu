// TEST INFRASTRUCTURE ONLY.  Times the reference's own CUDA force path — calculateNForces of
// /root/reference/src/library/MDSystem.cu, compiled UNMODIFIED for sm_100a and linked here — on the
// reference start lattice: "the existing GPU kernel on the same box" the new kernel is measured against
// (BASELINE.md §1).  usage: ref_gpu_bench N rho reps     -> one line: N, ms per call, ordered pairs/s
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <vector>

extern "C" {
void allocateArray(float** dest, int number);
void deleteArray(float* arr);
void copyArrayToDevice(float* device, const float* host, int numBodies);
void copyArrayFromDevice(float* host, const float* device, unsigned int pbo, int numBodies);
void calculateNForces(float* Pos, float* Force, float* host_pressure, int numBodies, float host_L, int Lperiodic,
                      int* host_RDF, float host_dr2, int p, int q);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 16384;
  const double rho = argc > 2 ? atof(argv[2]) : 0.85;
  const int reps = argc > 3 ? atoi(argv[3]) : 3;
  const int periodic = argc > 4 ? atoi(argv[4]) : 1;
  const double L = pow(N / rho, 1. / 3.);
  const int ns = (int)ceil(pow((double)N, 1. / 3.));
  std::vector<float> pos(4 * (size_t)N), frc(4 * (size_t)N);
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)(st >> 11) / 9007199254740992.0; };
  for (int i = 0; i < N; ++i) {
    pos[4 * i] = (float)(((i % ns) + 0.5 + 0.1 * (rnd() - 0.5)) * L / ns);
    pos[4 * i + 1] = (float)((((i / ns) % ns) + 0.5 + 0.1 * (rnd() - 0.5)) * L / ns);
    pos[4 * i + 2] = (float)(((i / (ns * ns)) + 0.5 + 0.1 * (rnd() - 0.5)) * L / ns);
    pos[4 * i + 3] = (float)(L / 150.);
  }
  float *d_pos, *d_frc;
  allocateArray(&d_pos, N);
  allocateArray(&d_frc, N);
  copyArrayToDevice(d_pos, pos.data(), N);
  float pressure = 0.f;
  int rdf[256];
  calculateNForces(d_pos, d_frc, &pressure, N, (float)L, periodic, rdf, 0.1f, 256, 1);   // warm-up
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r) calculateNForces(d_pos, d_frc, &pressure, N, (float)L, periodic, rdf, 0.1f, 256, 1);
  cudaDeviceSynchronize();
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
  copyArrayFromDevice(frc.data(), d_frc, 0, N);
  printf("{\"impl\": \"reference MDSystem.cu recompiled for sm_100a\", \"N\": %d, \"rho\": %g, \"periodic\": %d, "
         "\"ms_per_force_call\": %.4f, \"ordered_pairs_per_s\": %.4e, \"f0\": [%g, %g, %g], \"pressure\": %g}\n",
         N, rho, periodic, ms, (double)N * (N - 1) / (ms * 1e-3), frc[0], frc[1], frc[2], pressure);
  deleteArray(d_pos);
  deleteArray(d_frc);
  return 0;
}
