// Measurement only (not product code): what bounds the weight-gradient kernel's inner loop?  Same ring as mlp_wgrad_kernel
// (nb_mlp_tc_bwd.cu): one persistent CTA per SM, 3 stages x 64 KB (4 dY + 4 X half blobs of 64 points), per stage
// 2 x 4 tcgen05.mma M=128 N=256 K=16 into the 512 TMEM columns.  The operand BYTES are arbitrary (results are not looked at); varied:
//   * operand descriptors: MN-major SWIZZLE_NONE (what wgrad uses), MN-major SWIZZLE_128B (round 1), K-major SWIZZLE_128B (forward chain)
//   * source: HBM stream or an L2-resident window;   * MMAs on / off;   * the bias warps' shared-memory column sums on / off
//   * one or two bulk-copy issuing threads, 8 KB or 32 KB copies.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nerf_pytorch_paeng_b200/csrc -o /tmp/wgrad_probe scripts/wgrad_probe.cu && /tmp/wgrad_probe
#include <cstdlib>
#include "nb_tc_common.cuh"

using namespace tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

namespace {
constexpr int kStages = 3;
constexpr uint32_t kStageBytes = 65536;
constexpr int kThreads = 224;     // warp0 producer, warp1 MMA, warps 2-5 column sums, warp6 second producer

struct Params {
  const uint8_t* src;
  long long stages_per_cta, wrap_stages;
  uint32_t copy_bytes;
  int two_issuers;
  int layout;        // 0 MN-major no swizzle, 1 MN-major SW128, 2 K-major SW128
  int mma;           // MMAs per stage: 0 none, 1 = as wgrad (8), 2 = half of them (4: M=128 only)
  int colsum;        // bias warps read the A operand from shared memory
  float* sink;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_probe_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_full = sbase + kStages * kStageBytes, b_empty = b_full + 64, s_tmem = b_full + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_iss = p.two_issuers ? 2u : 1u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(b_full + 8 * i, n_iss); mbar_init(b_empty + 8 * i, 1 + (p.colsum ? 4 : 0)); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));
  const long long span = p.wrap_stages > 0 ? p.wrap_stages : p.stages_per_cta;
  const uint8_t* base = p.src + (size_t)blockIdx.x * (size_t)span * kStageBytes;

  if ((warp == 0 || (warp == 6 && p.two_issuers)) && lane == 0) {
    const uint32_t who = warp == 0 ? 0u : 1u, share = kStageBytes / n_iss;
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
    uint32_t stage = 0, phase = 0;
    const uint32_t idesc = p.layout == 2 ? umma_idesc(128, 256, 0, 0) : umma_idesc(128, 256, 1, 1);
    for (long long s = 0; s < p.stages_per_cta; ++s) {
      mbar_wait(b_full + 8 * stage, phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = sbase + stage * kStageBytes, b_addr = a_addr + 32768u;
        const int halves = p.mma == 1 ? 2 : (p.mma == 2 ? 1 : 0);
        for (int mh = 0; mh < halves; ++mh) {
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            uint64_t ad, bd;
            if (p.layout == 0) {          // [feature/8][64 points][8 features]: atoms every 1024 B (SBO), 8-point groups every 128 B (LBO)
              ad = umma_desc_mn_noswz(a_addr + (uint32_t)mh * 16384u + k16 * 256u, 128, 1024);
              bd = umma_desc_mn_noswz(b_addr + k16 * 256u, 128, 1024);
            } else if (p.layout == 1) {   // [64-feature chunk][64 points][128 B swizzled]: K16 slice = 16 rows = 2048 B, next 64-feature chunk 8192 B (LBO)
              ad = umma_desc(a_addr + (uint32_t)mh * 16384u + k16 * 2048u, 8192, 1024);
              bd = umma_desc(b_addr + k16 * 2048u, 8192, 1024);
            } else {                      // K-major rows of 128 B (64 k): 128 / 256 rows, K16 slice 32 B further
              ad = umma_desc(a_addr + (uint32_t)mh * 16384u + k16 * 32u, 16, 1024);
              bd = umma_desc(b_addr + k16 * 32u, 16, 1024);
            }
            umma_ss(tmem_base + (uint32_t)mh * 256u, ad, bd, idesc, (s > 0 || k16 > 0) ? 1u : 0u);
          }
        }
        umma_commit(b_empty + 8 * stage);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 2 && warp <= 5 && p.colsum) {
    const int t = threadIdx.x - 64, j8 = t & 7, q16 = t >> 3;
    float bs[2][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bs[0][k] = bs[1][k] = 0.f;
    uint32_t stage = 0, phase = 0;
    for (long long s = 0; s < p.stages_per_cta; ++s) {
      mbar_wait(b_full + 8 * stage, phase);
      const uint32_t st_base = sbase + stage * kStageBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t b0 = st_base + (uint32_t)(q16 + 16 * h) * 1024u + (uint32_t)j8 * 16u;
#pragma unroll 4
        for (uint32_t i = 0; i < 8; ++i) {
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(b0 + i * 128u));
          bs[h][0] += __uint_as_float(w0 << 16); bs[h][1] += __uint_as_float(w0 & 0xFFFF0000u);
          bs[h][2] += __uint_as_float(w1 << 16); bs[h][3] += __uint_as_float(w1 & 0xFFFF0000u);
          bs[h][4] += __uint_as_float(w2 << 16); bs[h][5] += __uint_as_float(w2 & 0xFFFF0000u);
          bs[h][6] += __uint_as_float(w3 << 16); bs[h][7] += __uint_as_float(w3 & 0xFFFF0000u);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(b_empty + 8 * stage);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += bs[0][k] + bs[1][k];
    if (acc == 12345.678f) p.sink[threadIdx.x] = acc;
  }
  // drain: every MMA has retired once the last stage's empty barrier has completed its final phase
  if (warp == 1 && lane == 0 && p.mma) {
    const long long n = p.stages_per_cta;
    const uint32_t last = (uint32_t)((n - 1) % kStages);
    mbar_wait(b_empty + 8 * last, (uint32_t)(((n - 1) / kStages) & 1));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void fill_random_kernel(uint32_t* w, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + 12345u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    // two bf16: sign | exponent in [120, 127] | 7 mantissa bits; every other element zero
    const uint32_t lo = (x & 0x8000u) | ((120u + ((x >> 7) & 7u)) << 7) | (x & 0x7Fu);
    w[i] = (x & 0x10000u) ? lo : (lo << 16);
  }
}

double run(Params p, int grid, int reps, float* ms_out) {
  const size_t smem = kStages * kStageBytes + 2048;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f, sum = 0.f;
  for (int r = 0; r < reps + 1; ++r) {
    CK(cudaEventRecord(e0));
    wgrad_probe_kernel<<<grid, kThreads, smem>>>(p);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
    if (r > 0) sum += ms;
  }
  ms_out[0] = best; ms_out[1] = sum / reps;
  return (double)grid * p.stages_per_cta * kStageBytes / (best * 1e-3) / 1e9;
}
}  // namespace

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int grid = prop.multiProcessorCount;
  CK(cudaFuncSetAttribute(wgrad_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes + 2048));
  const long long stages = 800;      // x 64 KB x 148 = 7.76 GB, the fine-pass wgrad volume
  uint8_t* buf; CK(cudaMalloc(&buf, (size_t)grid * stages * kStageBytes));
  CK(cudaMemset(buf, 0, (size_t)grid * stages * kStageBytes));      // bf16 zeros first: the tensor cores switch nothing
  float* sink; CK(cudaMalloc(&sink, 4096));
  const char* lname[3] = {"MN-major SWIZZLE_NONE (wgrad)", "MN-major SWIZZLE_128B", "K-major SWIZZLE_128B"};
  for (int data = 0; data < 2; ++data) {
  if (data == 1) {      // bf16 values in (-2, 2), half of them zero (post-ReLU activations): real switching activity, real power draw
    fill_random_kernel<<<grid * 8, 256>>>(reinterpret_cast<uint32_t*>(buf), (size_t)grid * stages * kStageBytes / 4);
    CK(cudaDeviceSynchronize());
  }
  for (int l2 = 1; l2 >= 0; --l2) {
    for (int layout = 0; layout < 3; ++layout)
      for (int mma = 1; mma <= 2; ++mma)
        for (int colsum = 0; colsum < 2; ++colsum)
          for (int two = 0; two < 2; ++two) {
            if (mma == 2 && (colsum || two)) continue;
            if (layout > 0 && (colsum || two)) continue;
            if (data == 1 && (layout > 0 || two)) continue;
            Params p{buf, stages, l2 ? 6 : 0, 8192u, two, layout, mma, colsum, sink};
            float msv[2];
            const double gbs = run(p, grid, 8, msv);
            const float ms = msv[0];
            const double flop = (double)grid * stages * (mma == 1 ? 2 : 1) * 4 * 2.0 * 128 * 256 * 16;
            printf("{\"data\": \"%s\", \"source\": \"%s\", \"layout\": \"%s\", \"mma_per_stage\": %d, \"colsum_warps\": %d, \"tma_issuers\": %d, \"ms\": %.4f, \"ms_mean_of_8\": %.4f, \"GBps\": %.1f, \"TFLOPs\": %.1f}\n",
                   data ? "random bf16, half zeros" : "zeros", l2 ? "L2 window" : "HBM stream", lname[layout], mma == 1 ? 8 : 4, colsum, two ? 2 : 1, ms, msv[1], gbs, flop / (ms * 1e-3) / 1e12);
            fflush(stdout);
          }
  }
  }
  CK(cudaFree(buf));
  return 0;
}
