// Loss gradient and Adam on flat buffers.
//
// Replaces train.py:58-65 (nn.MSELoss on rgb_c / rgb_f and its backward) and main.py:79-80 /
// train.py:70 (torch.optim.Adam(betas=(0.9,0.999)), eps 1e-8, no weight decay).  Elementwise,
// HBM-bound: Adam touches 4 reads + 3 writes of 4 B per parameter.
#include "nb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
mse_grad_kernel(long long n3, const float* __restrict__ rgb, const float* __restrict__ target, float scale,
                float loss_scale, float* __restrict__ d_rgb, float* __restrict__ loss_out) {
  float local = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += stride) {
    const float diff = rgb[i] - target[i];
    if (d_rgb) d_rgb[i] = scale * diff;
    local += diff * diff;
  }
  if (loss_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      atomicAdd(loss_out, s * loss_scale);
    }
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            float beta1, float beta2, float eps, float step_size, float inv_bc2_sqrt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);          // exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * beta2 + (gi * gi) * (1.0f - beta2);    // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;            // sqrt(v)/sqrt(bc2) + eps
    p[i] = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// Data-parallel variant: the gradient all-reduce is folded into Adam's load.  Every rank holds the gradient
// buffers of all ranks (its own, plus the peers' copies that arrived over NVLink through the copy engines while the backward was
// still running); this kernel sums them IN RANK ORDER (so every rank computes bit-identical sums and the replicas never diverge),
// writes the sum back to g (p.grad then holds the all-reduced gradient) and applies the update.
struct GradSrcs { const float* src[NB_MAX_RANKS]; int n; };

__global__ void __launch_bounds__(256)
adam_sum_kernel(long long n, float* __restrict__ p, float* __restrict__ g, GradSrcs s, float* __restrict__ m, float* __restrict__ v,
                float beta1, float beta2, float eps, float step_size, float inv_bc2_sqrt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = s.src[0][i];
    for (int r = 1; r < s.n; ++r) gi += s.src[r][i];
    g[i] = gi;
    const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);
    const float vi = v[i] * beta2 + (gi * gi) * (1.0f - beta2);
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// frame output (test.py:50-61, utils.py:11): rgb8 = to8b(rgb), disp8 = to8b(disp / nanmax(disp))
__global__ void __launch_bounds__(256)
nanmax_kernel(long long n, const float* __restrict__ x, float* __restrict__ out) {
  float m = 0.f;   // disparities are >= 0; NaNs are skipped like np.nanmax
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    if (v == v) m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // valid for non-negative floats
}

__device__ __forceinline__ unsigned char to8b_dev(float x) {
  // (255*np.clip(x,0,1)).astype(np.uint8): truncation towards zero; NaN -> 0
  if (!(x == x)) return 0;
  return (unsigned char)(255.0f * fminf(fmaxf(x, 0.f), 1.f));
}

__global__ void __launch_bounds__(256)
frame8_kernel(long long n, const float* __restrict__ rgb, const float* __restrict__ disp, const float* __restrict__ disp_max,
              unsigned char* __restrict__ rgb8, unsigned char* __restrict__ disp8) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float dm = disp_max ? *disp_max : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * 3; i += stride) rgb8[i] = to8b_dev(rgb[i]);
  if (disp && disp8)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) disp8[i] = to8b_dev(disp[i] / dm);
}

}  // namespace

extern "C" int nb_frame_to8b(nb_handle_t h, int64_t N, const float* rgb, const float* disp, float* disp_max_scratch,
                             uint8_t* rgb8, uint8_t* disp8, void* stream) {
  NB_ENTER(h);
  if (N == 0) return NB_OK;
  NB_REQUIRE(h, N > 0 && rgb && rgb8 && (!disp8 || (disp && disp_max_scratch)), "nb_frame_to8b: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((N + 255) / 256);
  const int cap = h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (disp8) {
    NB_CUDA(h, cudaMemsetAsync(disp_max_scratch, 0, sizeof(float), st));
    nanmax_kernel<<<blocks, 256, 0, st>>>((long long)N, disp, disp_max_scratch);
    NB_LAUNCHED(h);
  }
  frame8_kernel<<<blocks, 256, 0, st>>>((long long)N, rgb, disp8 ? disp : nullptr, disp8 ? disp_max_scratch : nullptr, rgb8, disp8);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_mse_grad(nb_handle_t h, int64_t N, const float* rgb, const float* target, float scale, float loss_scale,
                           float* d_rgb, float* loss_out, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, N >= 0 && rgb && target && (d_rgb || loss_out), "nb_mse_grad: bad arguments");
  if (N == 0) return NB_OK;
  const long long n3 = (long long)N * 3;
  long long blocks = (n3 + 255) / 256;
  const long long cap = (long long)h->sm_count * 4;
  if (blocks > cap) blocks = cap;
  mse_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n3, rgb, target, scale, loss_scale, d_rgb, loss_out);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_adam_step(nb_handle_t h, int64_t n, float* p, const float* g, float* m, float* v, float lr, float beta1,
                            float beta2, float eps, int32_t step, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, n >= 0 && p && g && m && v && step >= 1, "nb_adam_step: bad arguments");
  if (n == 0) return NB_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((long long)n, p, g, m, v, beta1, beta2, eps, step_size, inv_bc2_sqrt);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_adam_step_sum(nb_handle_t h, int64_t n, float* p, float* g, const float* const* srcs, int32_t n_srcs, float* m,
                                float* v, float lr, float beta1, float beta2, float eps, int32_t step, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, n >= 0 && p && g && srcs && m && v && step >= 1 && n_srcs >= 1 && n_srcs <= NB_MAX_RANKS, "nb_adam_step_sum: bad arguments");
  if (n == 0) return NB_OK;
  GradSrcs s;
  s.n = n_srcs;
  for (int i = 0; i < NB_MAX_RANKS; ++i) s.src[i] = i < n_srcs ? srcs[i] : nullptr;
  for (int i = 0; i < n_srcs; ++i) NB_REQUIRE(h, srcs[i], "nb_adam_step_sum: NULL gradient source");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  adam_sum_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((long long)n, p, g, s, m, v, beta1, beta2, eps, step_size, inv_bc2_sqrt);
  NB_LAUNCHED(h);
  return NB_OK;
}
