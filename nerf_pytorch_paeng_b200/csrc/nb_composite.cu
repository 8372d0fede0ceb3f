// K5: alpha compositing (raw2outputs) forward and fused backward.
//
// Replaces nerf_process.py:89-140 (post_process) and its autograd.  One warp per ray; lane l owns
// samples l, l+32, ...; the transmittance T_s = prod_{j<s}(1-alpha_j+1e-10) is an exclusive
// multiplicative warp scan (shuffles) with a running carry across 32-sample chunks; the backward
// needs the suffix quantity R_s = sum_{k>s} dw_k alpha_k prod_{s<j<k} f_j, a reverse affine scan
// (no division, so fully-opaque samples with f ~ 1e-10 are safe).
// HBM-bound: forward 20 B/sample read + 4 B/sample (weights) + 24 B/ray written; raw is read as
// one float4 per sample (512 B per warp instruction, fully coalesced).  The training drivers use composite_train_kernel
// (forward + MSE gradient + backward in one pass over raw, per-sample values kept in registers).
#include "nb_common.cuh"
#include "nb_mlp.h"

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// alpha, f=(1-alpha)+1e-10 and e=exp(-relu(sigma)*dist) of one sample
__device__ __forceinline__ void sample_alpha(float sigma, float dist, float& alpha, float& f, float& e) {
  e = expf(-(fmaxf(sigma, 0.0f) * dist));
  alpha = 1.0f - e;
  f = (1.0f - alpha) + 1e-10f;
}

__global__ void __launch_bounds__(kWarps * 32)
composite_fwd_kernel(long long N, int S, const float4* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, float* __restrict__ rgb_out, float* __restrict__ disp_out,
                     float* __restrict__ acc_out, float* __restrict__ w_out, float* __restrict__ depth_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  for (long long ray = warp0; ray < N; ray += (long long)gridDim.x * kWarps) {
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    float carry = 1.0f;                       // T at the start of this 32-sample chunk
    float ar = 0.f, ag = 0.f, ab = 0.f, aw = 0.f, adep = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float4 r = ok ? __ldg(&raw[ray * S + s]) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float zs = ok ? z[ray * S + s] : 0.f;
      // dists[s] = z[s+1]-z[s], last = 1e10; * ||d||            (nerf_process.py:93-101)
      float zn = __shfl_down_sync(0xffffffffu, zs, 1);
      if (lane == 31 && s + 1 < S) zn = z[ray * S + s + 1];
      const float dist = ((s + 1 < S) ? (zn - zs) : 1e10f) * dnorm;
      float alpha, f, e;
      sample_alpha(r.w, dist, alpha, f, e);
      if (!ok) { alpha = 0.f; f = 1.f; }
      // exclusive product scan of f over the warp
      float incl = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      carry = carry * __shfl_sync(0xffffffffu, incl, 31);
      const float w = alpha * T;
      if (ok) {
        if (w_out) w_out[ray * S + s] = w;
        ar += w * sigmoidf_(r.x); ag += w * sigmoidf_(r.y); ab += w * sigmoidf_(r.z);
        aw += w; adep += w * zs;
      }
    }
    ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); aw = warp_sum(aw); adep = warp_sum(adep);
    if (lane == 0) {
      const float bg = 1.0f - aw;                                  // rgb_map + (1 - acc_map)  (:138)
      rgb_out[ray * 3] = ar + bg; rgb_out[ray * 3 + 1] = ag + bg; rgb_out[ray * 3 + 2] = ab + bg;
      if (acc_out) acc_out[ray] = aw;
      if (depth_out) depth_out[ray] = adep;
      if (disp_out) {
        const float q = adep / aw;                                 // NaN when acc == 0
        float disp = (q != q) ? 0.0f : 1.0f / fmaxf(1e-10f, q);    // torch.max(1e-10, nan)=nan -> where(isnan) -> 0
        if (disp != disp) disp = 0.0f;
        if (disp > 5.0f) disp = 5.0f;
        disp_out[ray] = disp;
      }
    }
  }
}

// backward for d_rgb only.  Two passes over the ray's samples held in registers (S <= 32*kMaxChunks).
constexpr int kMaxChunks = 24;   // S <= 768 (covers the 256+512 sweep point)

__global__ void __launch_bounds__(kWarps * 32)
composite_bwd_kernel(long long N, int S, const float4* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, const float* __restrict__ d_rgb, float4* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int nchunk = (S + 31) >> 5;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  for (long long ray = warp0; ray < N; ray += (long long)gridDim.x * kWarps) {
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float gr = d_rgb[ray * 3], gg = d_rgb[ray * 3 + 1], gb = d_rgb[ray * 3 + 2];
    const float gsum = gr + gg + gb;
    // ---- forward sweep: T_s at each sample; keep per-chunk carries for the reverse sweep
    float carry = 1.0f;
    // reverse sweep state: R at the first sample of the chunk to the right
    // (processed right-to-left, so recompute per chunk from global memory: raw/z are L1/L2 hot)
    float chunk_T0[kMaxChunks];
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
      chunk_T0[c] = carry;
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const float sig = ok ? __ldg(&raw[ray * S + s]).w : 0.f;
      const float zs = ok ? z[ray * S + s] : 0.f;
      float zn = __shfl_down_sync(0xffffffffu, zs, 1);
      if (lane == 31 && s + 1 < S) zn = z[ray * S + s + 1];
      const float dist = ((s + 1 < S) ? (zn - zs) : 1e10f) * dnorm;
      float alpha, f, e;
      sample_alpha(sig, dist, alpha, f, e);
      if (!ok) f = 1.f;
      float incl = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    // ---- reverse sweep
    float Rcarry = 0.0f;     // R_s for s = last sample of the current chunk (suffix beyond the chunk)
#pragma unroll 1
    for (int c = nchunk - 1; c >= 0; --c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const float4 r = ok ? __ldg(&raw[ray * S + s]) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float zs = ok ? z[ray * S + s] : 0.f;
      float zn = __shfl_down_sync(0xffffffffu, zs, 1);
      if (lane == 31 && s + 1 < S) zn = z[ray * S + s + 1];
      const float dist = ((s + 1 < S) ? (zn - zs) : 1e10f) * dnorm;
      float alpha, f, e;
      sample_alpha(r.w, dist, alpha, f, e);
      if (!ok) { alpha = 0.f; f = 1.f; }
      float incl = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = chunk_T0[c] * excl;
      const float cr = sigmoidf_(r.x), cg = sigmoidf_(r.y), cb = sigmoidf_(r.z);
      // dL/dw_s = g . c_s - sum(g)   (rgb_map = sum w c + 1 - sum w)
      const float dw = ok ? (gr * cr + gg * cg + gb * cb - gsum) : 0.f;
      // R_s = a_{s+1} + f_{s+1} * R_{s+1}, a_k = dw_k*alpha_k : reverse inclusive affine scan of
      // pairs (m,b) = (f_k, a_k) taken from the right neighbour.
      float m = f, b = dw * alpha;               // element k=s contributes to R_{s-1}
      // suffix-compose within the warp: after the scan lane l holds the composition of
      // elements l..31 applied to the incoming Rcarry:  R_{l-1} = B_l + M_l * Rcarry
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float m2 = __shfl_down_sync(0xffffffffu, m, o);
        float b2 = __shfl_down_sync(0xffffffffu, b, o);
        if (lane + o < 32) { b = b + m * b2; m = m * m2; }
      }
      const float Rprev = b + m * Rcarry;        // R_{s-1} for this lane's s
      float R = __shfl_down_sync(0xffffffffu, Rprev, 1);   // R_s = value computed by lane+1
      if (lane == 31) R = Rcarry;
      Rcarry = __shfl_sync(0xffffffffu, Rprev, 0);         // R_{first-1}: carry into the chunk on the left
      if (ok) {
        const float w = alpha * T;
        const float dalpha = dw * T - T * R;
        // alpha = 1 - exp(-relu(sigma)*dist): d alpha/d sigma = dist*e for sigma > 0
        const float dsig = (r.w > 0.0f) ? dalpha * dist * e : 0.0f;
        float4 o4;
        o4.x = w * gr * cr * (1.0f - cr);
        o4.y = w * gg * cg * (1.0f - cg);
        o4.z = w * gb * cb * (1.0f - cb);
        o4.w = dsig;
        d_raw[ray * S + s] = o4;
      }
    }
  }
}

// Training: post_process + the MSE gradient + the compositing backward of ONE ray in one pass (nb_train_rays).  Same arithmetic,
// operation for operation, as composite_fwd_kernel -> mse_grad_kernel -> composite_bwd_kernel (the stage-by-stage entries stay,
// and test_fused_drivers_equal_stepwise_calls compares the two bit for bit), but raw / z are read ONCE: the per-sample quantities of the
// forward sweep (colours, alpha, f, T, dist, e) stay in registers for the reverse sweep, so the backward's two extra passes over
// raw and two of the three launches disappear.  NCH = number of 32-sample chunks held (S <= 32 * NCH).
template <int NCH>
__global__ void __launch_bounds__(kWarps * 32)
composite_train_kernel(long long N, int S, const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                       const float* __restrict__ target, float scale, float loss_scale, float* __restrict__ rgb_out,
                       float* __restrict__ disp_out, float* __restrict__ w_out, float4* __restrict__ d_raw, float* __restrict__ loss_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  float loss_local = 0.f;
  for (long long ray = warp0; ray < N; ray += (long long)gridDim.x * kWarps) {
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    float carry = 1.0f;
    float ar = 0.f, ag = 0.f, ab = 0.f, aw = 0.f, adep = 0.f;
    float cr_[NCH], cg_[NCH], cb_[NCH], al_[NCH], f_[NCH], T_[NCH], di_[NCH], e_[NCH], sg_[NCH];
    // ---- forward sweep (composite_fwd_kernel)
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      float4 r = ok ? __ldg(&raw[ray * S + s]) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float zs = ok ? z[ray * S + s] : 0.f;
      float zn = __shfl_down_sync(0xffffffffu, zs, 1);
      if (lane == 31 && s + 1 < S) zn = z[ray * S + s + 1];
      const float dist = ((s + 1 < S) ? (zn - zs) : 1e10f) * dnorm;
      float alpha, f, e;
      sample_alpha(r.w, dist, alpha, f, e);
      if (!ok) { alpha = 0.f; f = 1.f; }
      float incl = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      carry = carry * __shfl_sync(0xffffffffu, incl, 31);
      const float w = alpha * T;
      const float cr = sigmoidf_(r.x), cg = sigmoidf_(r.y), cb = sigmoidf_(r.z);
      if (ok) {
        if (w_out) w_out[ray * S + s] = w;
        ar += w * cr; ag += w * cg; ab += w * cb;
        aw += w; adep += w * zs;
      }
      cr_[c] = cr; cg_[c] = cg; cb_[c] = cb; al_[c] = alpha; f_[c] = f; T_[c] = T; di_[c] = dist; e_[c] = e; sg_[c] = r.w;
    }
    ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); aw = warp_sum(aw); adep = warp_sum(adep);
    const float bg = 1.0f - aw;
    const float rgb0 = ar + bg, rgb1 = ag + bg, rgb2 = ab + bg;
    if (lane == 0) {
      rgb_out[ray * 3] = rgb0; rgb_out[ray * 3 + 1] = rgb1; rgb_out[ray * 3 + 2] = rgb2;
      if (disp_out) {
        const float q = adep / aw;
        float disp = (q != q) ? 0.0f : 1.0f / fmaxf(1e-10f, q);
        if (disp != disp) disp = 0.0f;
        if (disp > 5.0f) disp = 5.0f;
        disp_out[ray] = disp;
      }
    }
    // ---- MSE gradient (mse_grad_kernel): d_rgb = scale * (rgb - target); the loss sum is taken per ray by lane 0
    const float d0 = rgb0 - target[ray * 3], d1 = rgb1 - target[ray * 3 + 1], d2 = rgb2 - target[ray * 3 + 2];
    const float gr = scale * d0, gg = scale * d1, gb = scale * d2;
    if (lane == 0) loss_local += d0 * d0 + d1 * d1 + d2 * d2;
    const float gsum = gr + gg + gb;
    // ---- reverse sweep (composite_bwd_kernel)
    float Rcarry = 0.0f;
#pragma unroll
    for (int c = NCH - 1; c >= 0; --c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const float alpha = al_[c], f = f_[c], T = T_[c], cr = cr_[c], cg = cg_[c], cb = cb_[c];
      const float dw = ok ? (gr * cr + gg * cg + gb * cb - gsum) : 0.f;
      float m = f, b = dw * alpha;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float m2 = __shfl_down_sync(0xffffffffu, m, o);
        float b2 = __shfl_down_sync(0xffffffffu, b, o);
        if (lane + o < 32) { b = b + m * b2; m = m * m2; }
      }
      const float Rprev = b + m * Rcarry;
      float R = __shfl_down_sync(0xffffffffu, Rprev, 1);
      if (lane == 31) R = Rcarry;
      Rcarry = __shfl_sync(0xffffffffu, Rprev, 0);
      if (ok) {
        const float w = alpha * T;
        const float dalpha = dw * T - T * R;
        const float dsig = (sg_[c] > 0.0f) ? dalpha * di_[c] * e_[c] : 0.0f;
        float4 o4;
        o4.x = w * gr * cr * (1.0f - cr);
        o4.y = w * gg * cg * (1.0f - cg);
        o4.z = w * gb * cb * (1.0f - cb);
        o4.w = dsig;
        d_raw[ray * S + s] = o4;
      }
    }
  }
  if (loss_out) {
    __shared__ float red[kWarps];
    if (lane == 0) red[threadIdx.x >> 5] = loss_local;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < kWarps; ++w) t += red[w];
      atomicAdd(loss_out, t * loss_scale);
    }
  }
}

}  // namespace

// internal (nb_render.cu): returns NB_ERR_UNSUPPORTED when S does not fit the register-resident variants (the caller then runs the
// three stage-by-stage entries)
int nb_composite_train(nb_handle_t h, int64_t N, int32_t S, const float* raw, const float* z, const float* rays_d, const float* target,
                       float scale, float loss_scale, float* rgb, float* disp, float* weights, float* d_raw, float* loss_out,
                       cudaStream_t st) {
  if (S > 192 || N <= 0 || (((uintptr_t)raw | (uintptr_t)d_raw) & 15) != 0) return NB_ERR_UNSUPPORTED;
  long long blocks = (N + kWarps - 1) / kWarps;
  const long long cap = (long long)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (S <= 64)
    composite_train_kernel<2><<<(int)blocks, kWarps * 32, 0, st>>>((long long)N, S, (const float4*)raw, z, rays_d, target, scale, loss_scale,
                                                                   rgb, disp, weights, (float4*)d_raw, loss_out);
  else
    composite_train_kernel<6><<<(int)blocks, kWarps * 32, 0, st>>>((long long)N, S, (const float4*)raw, z, rays_d, target, scale, loss_scale,
                                                                   rgb, disp, weights, (float4*)d_raw, loss_out);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_composite_forward(nb_handle_t h, int64_t N, int32_t S, const float* raw, const float* z,
                                    const float* rays_d, float* rgb, float* disp, float* acc, float* weights,
                                    float* depth, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, N >= 0 && S > 0 && raw && z && rays_d && rgb, "nb_composite_forward: bad arguments");
  NB_REQUIRE(h, ((uintptr_t)raw & 15) == 0, "nb_composite_forward: raw must be 16-byte aligned");
  if (N == 0) return NB_OK;
  long long blocks = (N + kWarps - 1) / kWarps;
  const long long cap = (long long)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  composite_fwd_kernel<<<(int)blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(
      (long long)N, S, (const float4*)raw, z, rays_d, rgb, disp, acc, weights, depth);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_composite_backward(nb_handle_t h, int64_t N, int32_t S, const float* raw, const float* z,
                                     const float* rays_d, const float* d_rgb, float* d_raw, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, N >= 0 && S > 0 && S <= 32 * kMaxChunks && raw && z && rays_d && d_rgb && d_raw,
             "nb_composite_backward: bad arguments (S <= 768)");
  NB_REQUIRE(h, (((uintptr_t)raw | (uintptr_t)d_raw) & 15) == 0, "nb_composite_backward: raw/d_raw must be 16-byte aligned");
  if (N == 0) return NB_OK;
  long long blocks = (N + kWarps - 1) / kWarps;
  const long long cap = (long long)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  composite_bwd_kernel<<<(int)blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(
      (long long)N, S, (const float4*)raw, z, rays_d, d_rgb, (float4*)d_raw);
  NB_LAUNCHED(h);
  return NB_OK;
}
