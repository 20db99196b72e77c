// Lets g++ parse the __device__ helpers of tfy_common.cuh when a host-only translation unit (the DDP reducer) is
// built for the CPU-side unit test: the helpers are never called there, they only have to compile.
#pragma once
#include <cuda_runtime.h>
#ifndef __CUDACC__
struct tfy_test_dim3 { unsigned x, y, z; };
static tfy_test_dim3 threadIdx, blockIdx, blockDim, gridDim;
static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline long long clock64() { return 0; }
static inline float __uint_as_float(unsigned u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; __builtin_memcpy(&u, &f, 4); return u; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p += v; return o; }
#endif
