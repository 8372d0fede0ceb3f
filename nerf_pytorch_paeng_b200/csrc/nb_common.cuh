// Shared declarations for libnerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nerf_b200.h"

struct nb_handle_s {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  long long launches;   // kernels launched through this handle
  char err[512];
  // per-device lazily initialised state (function attributes are per device; a process may hold one handle per GPU)
  bool fwd_attr_done[9];
  bool bwd_attr_done;
};

#define NB_SET_ERR(h, ...) do { if (h) snprintf((h)->err, sizeof((h)->err), __VA_ARGS__); } while (0)

#define NB_REQUIRE(h, cond, ...)                                   \
  do { if (!(cond)) { NB_SET_ERR(h, __VA_ARGS__); return NB_ERR_INVALID; } } while (0)

#define NB_CUDA(h, call)                                                                   \
  do { cudaError_t e__ = (call);                                                           \
       if (e__ != cudaSuccess) {                                                           \
         NB_SET_ERR(h, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
         return NB_ERR_CUDA; } } while (0)

// every entry point: validate handle, bind the handle's device to this thread
#define NB_ENTER(h)                                                 \
  do { if (!(h)) return NB_ERR_INVALID;                             \
       NB_CUDA(h, cudaSetDevice((h)->device)); } while (0)

// after a launch: count it and surface launch-configuration errors
#define NB_LAUNCHED(h)                                              \
  do { (h)->launches++; NB_CUDA(h, cudaGetLastError()); } while (0)

static inline int nb_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Philox4x32-10 (Salmon et al. 2011): counter-based RNG for the perf path
// (the reference draws torch.rand, nerf_process.py:55,162; only the distribution is contractual).
__device__ __forceinline__ uint4 nb_philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
// uniform in [0,1) with 24 bits, like torch.rand's fp32 path
__device__ __forceinline__ float nb_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
