// Measurement only (not product code): how fast can ONE SM pull operand bytes into shared memory on sm_100a, and through which
// path?  The weight-gradient kernel (nb_mlp_tc_bwd.cu, mlp_wgrad_kernel) streams 64 KB stages of 8 KB half blobs with zero reuse,
// so its ceiling is min(HBM, per-SM ingest).  This probe runs the same ring (3 stages x 64 KB, one persistent CTA per SM, 192 threads)
// with NO tensor work and varies
//   * the size of one bulk copy (cp.async.bulk, UBLKCP): 1 KB .. 64 KB,
//   * the source: a stream larger than L2 (HBM) or a 64 MB window that stays in L2,
//   * the issuing path: one thread (TMA), two threads (TMA), TMA + cp.async (LDGSTS) by four otherwise idle warps, LDGSTS only.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ingest_probe scripts/ingest_probe.cu && /tmp/ingest_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

namespace {
constexpr int kStages = 3;
constexpr uint32_t kStageBytes = 65536;
constexpr int kThreads = 224;     // warp0 TMA producer, warp1 consumer, warps 2-5 LDGSTS producers, warp6 second TMA producer

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 4000000000LL) { printf("probe: mbarrier timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ldgsts16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldgsts_arrive(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

struct Params {
  const uint8_t* src;
  long long stages_per_cta;     // 64 KB stages each CTA pulls
  long long wrap_stages;        // > 0: the CTA's slice wraps after this many stages (L2-resident window)
  uint32_t copy_bytes;          // size of one bulk copy
  uint32_t tma_bytes;           // bytes of every stage that arrive by bulk copy (the rest by LDGSTS); multiple of copy_bytes
  int two_issuers;              // bulk copies split between warp 0 and warp 6
};

__global__ void __launch_bounds__(kThreads, 1) ingest_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_full = sbase + kStages * kStageBytes, b_empty = b_full + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ld_bytes = kStageBytes - p.tma_bytes;
  const uint32_t n_tma_issuers = p.tma_bytes ? (p.two_issuers ? 2u : 1u) : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(b_full + 8 * i, n_tma_issuers + (ld_bytes ? 128u : 0u)); mbar_init(b_empty + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long span = p.wrap_stages > 0 ? p.wrap_stages : p.stages_per_cta;
  const uint8_t* base = p.src + (size_t)blockIdx.x * (size_t)span * kStageBytes;
  if ((warp == 0 || (warp == 6 && p.two_issuers)) && lane == 0 && p.tma_bytes) {
    const uint32_t who = warp == 0 ? 0u : 1u;
    const uint32_t share = p.tma_bytes / n_tma_issuers;       // bytes of a stage this thread issues
    uint32_t stage = 0, phase = 0;
    for (long long s = 0; s < p.stages_per_cta; ++s) {
      const uint8_t* g = base + (size_t)(p.wrap_stages > 0 ? s % p.wrap_stages : s) * kStageBytes + who * share;
      mbar_wait(b_empty + 8 * stage, phase ^ 1);
      mbar_expect_tx(b_full + 8 * stage, share);
      const uint32_t dst = sbase + stage * kStageBytes + who * share;
      for (uint32_t o = 0; o < share; o += p.copy_bytes) bulk_g2s(dst + o, g + o, p.copy_bytes, b_full + 8 * stage);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long s = 0; s < p.stages_per_cta; ++s) {
        mbar_wait(b_full + 8 * stage, phase);
        mbar_arrive(b_empty + 8 * stage);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2 && warp <= 5 && ld_bytes) {
    const uint32_t t = threadIdx.x - 64;      // 0..127
    uint32_t stage = 0, phase = 0;
    for (long long s = 0; s < p.stages_per_cta; ++s) {
      const uint8_t* g = base + (size_t)(p.wrap_stages > 0 ? s % p.wrap_stages : s) * kStageBytes + p.tma_bytes;
      mbar_wait(b_empty + 8 * stage, phase ^ 1);
      const uint32_t dst = sbase + stage * kStageBytes + p.tma_bytes;
      for (uint32_t o = t * 16u; o < ld_bytes; o += 128u * 16u) ldgsts16(dst + o, g + o);
      ldgsts_arrive(b_full + 8 * stage);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  }
}

double run(const uint8_t* src, int grid, long long stages_per_cta, long long wrap, uint32_t copy, uint32_t tma_bytes, int two, int reps) {
  Params p{src, stages_per_cta, wrap, copy, tma_bytes, two};
  const size_t smem = kStages * kStageBytes + 2048;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    CK(cudaEventRecord(e0));
    ingest_kernel<<<grid, kThreads, smem>>>(p);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  return (double)grid * stages_per_cta * kStageBytes / (best * 1e-3) / 1e9;      // GB/s
}
}  // namespace

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int grid = prop.multiProcessorCount;
  const double ghz = prop.clockRate * 1e-6;
  CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes + 2048));
  const long long hbm_stages = 800;                           // 800 x 64 KB x 148 = 7.76 GB: the fine-pass wgrad volume
  const long long l2_wrap = 6;                                // 6 x 64 KB x 148 = 58 MB window
  uint8_t* buf; CK(cudaMalloc(&buf, (size_t)grid * hbm_stages * kStageBytes));
  CK(cudaMemset(buf, 1, (size_t)grid * hbm_stages * kStageBytes));
  printf("{\"sm_count\": %d, \"nominal_ghz\": %.3f, \"ring\": \"3 x 64 KB per CTA, 1 CTA per SM\"}\n", grid, ghz);
  auto report = [&](const char* what, const char* srcname, uint32_t copy, uint32_t tma, int two, double gbs) {
    printf("{\"path\": \"%s\", \"source\": \"%s\", \"copy_bytes\": %u, \"tma_bytes_per_stage\": %u, \"tma_issuers\": %d, \"GBps\": %.1f, \"GBps_per_sm\": %.2f}\n",
           what, srcname, copy, tma, two ? 2 : 1, gbs, gbs / grid);
    fflush(stdout);
  };
  for (int l2 = 0; l2 < 2; ++l2) {
    const long long wrap = l2 ? l2_wrap : 0;
    const char* sn = l2 ? "L2 window 58 MB" : "HBM stream 7.76 GB";
    for (uint32_t copy = 1024; copy <= 65536; copy *= 2) report("bulk copy, one issuing thread", sn, copy, kStageBytes, 0, run(buf, grid, hbm_stages, wrap, copy, kStageBytes, 0, 3));
    report("bulk copy, two issuing threads", sn, 8192, kStageBytes, 1, run(buf, grid, hbm_stages, wrap, 8192, kStageBytes, 1, 3));
    report("bulk copy 32 KB + LDGSTS 32 KB (4 warps)", sn, 8192, 32768, 0, run(buf, grid, hbm_stages, wrap, 8192, 32768, 0, 3));
    report("bulk copy 48 KB + LDGSTS 16 KB (4 warps)", sn, 8192, 49152, 0, run(buf, grid, hbm_stages, wrap, 8192, 49152, 0, 3));
    report("LDGSTS only (4 warps)", sn, 0, 0, 0, run(buf, grid, hbm_stages, wrap, 8192, 0, 0, 3));
  }
  CK(cudaFree(buf));
  return 0;
}
