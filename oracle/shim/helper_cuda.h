// TEST INFRASTRUCTURE ONLY.  Stand-in for the CUDA-samples header the reference downloads at configure time
// (/root/reference/CMakeLists.txt:41-47; not under /root/reference): only the two macros MDSystem.cu uses.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define checkCudaErrors(call)                                                             \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(EXIT_FAILURE);                                                                 \
    }                                                                                     \
  } while (0)
#define getLastCudaError(msg)                                                             \
  do {                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess) {                                                              \
      fprintf(stderr, "%s: %s\n", msg, cudaGetErrorString(e_));                            \
      exit(EXIT_FAILURE);                                                                 \
    }                                                                                     \
  } while (0)
