// Fused drivers: render_rays (nerf_process.py:185-216) and the loss+backward half of train.py:53-70 as ONE
// C-ABI call each.  Pure host-side sequencing of the library's own entry points on the caller's stream --
// no allocation, no synchronisation; every intermediate lives in the caller's workspace.
#include "nb_common.cuh"
#include "nb_mlp.h"

namespace {

struct Carve {
  uint8_t* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = (off + 1023) & ~(size_t)1023;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Plan {
  float *z_c, *w_c, *z_f, *w_f, *raw, *d_raw, *d_rgb, *rays_d, *rgb_tmp, *disp_tmp;
  float *z_last, *raw_last;        // exact_last: depth and fp32 network output of every ray's last sample
  uint8_t *act, *mlp_ws, *fp32_ws;
  size_t act_bytes, mlp_ws_bytes, fp32_ws_bytes, total;
};

// exact_last (nb_render_cfg): the reference gives the LAST sample of a ray a 1e10-long interval (nerf_process.py:98), so its alpha is
// a step function of sign(sigma_last) and the bf16 path's ~1e-3 noise on sigma can land a ray on the other side of the step when
// |sigma_last| is tiny (random-init networks).  With the flag set the last sample of every ray is re-evaluated on the fp32 path
// (N points per network) and its network output replaces the bf16 one before compositing.
__global__ void gather_last_kernel(long long N, int S, const float* __restrict__ z, float* __restrict__ z_last) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) z_last[i] = z[i * S + S - 1];
}
__global__ void scatter_last_kernel(long long N, int S, const float4* __restrict__ raw_last, float4* __restrict__ raw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) raw[i * S + S - 1] = raw_last[i];
}

int make_plan(nb_handle_t h, const nb_mlp_desc* d, int64_t N, const nb_render_cfg* c, int train, void* ws, Plan* p) {
  const int64_t S = c->S_c + (c->S_f > 0 ? c->S_f : 0);
  const int64_t P = N * S;
  Carve cv{(uint8_t*)ws, 0};
  p->z_c = cv.take<float>(N * c->S_c);
  p->w_c = cv.take<float>(N * c->S_c);
  p->z_f = cv.take<float>(N * S);
  p->w_f = cv.take<float>(N * S);
  p->raw = cv.take<float>(P * 4);
  p->rays_d = cv.take<float>(N * 3);
  p->rgb_tmp = cv.take<float>(N * 3);
  p->disp_tmp = cv.take<float>(N);
  p->d_raw = train ? cv.take<float>(P * 4) : nullptr;
  p->d_rgb = train ? cv.take<float>(N * 3) : nullptr;
  size_t act = 0, wf = 0, wb = 0;
  int rc;
  if (train && (rc = nb_mlp_act_bytes(h, d, P, c->precision, &act))) return rc;
  if ((rc = nb_mlp_workspace_bytes(h, d, P, c->precision, 0, &wf))) return rc;
  if (train && (rc = nb_mlp_workspace_bytes(h, d, P, c->precision, 1, &wb))) return rc;
  p->act_bytes = act;
  p->mlp_ws_bytes = wf > wb ? wf : wb;
  p->act = cv.take<uint8_t>(act);
  p->mlp_ws = cv.take<uint8_t>(p->mlp_ws_bytes);
  p->z_last = p->raw_last = nullptr; p->fp32_ws = nullptr; p->fp32_ws_bytes = 0;
  if (c->exact_last && c->precision == NB_BF16) {
    if ((rc = nb_mlp_workspace_bytes(h, d, N, NB_FP32, 0, &p->fp32_ws_bytes))) return rc;
    p->z_last = cv.take<float>(N);
    p->raw_last = cv.take<float>(N * 4);
    p->fp32_ws = cv.take<uint8_t>(p->fp32_ws_bytes);
  }
  p->total = cv.off + 1024;
  return NB_OK;
}

int check_common(nb_handle_t h, const nb_mlp_desc* d, const nb_render_cfg* c, int64_t N) {
  NB_REQUIRE(h, d && c, "render: NULL desc/cfg");
  NB_REQUIRE(h, N >= 0 && c->S_c >= 2 && c->S_f >= 0, "render: bad sizes N=%lld S_c=%d S_f=%d", (long long)N, c->S_c, c->S_f);
  NB_REQUIRE(h, c->precision == NB_FP32 || c->precision == NB_BF16, "render: bad precision");
  NB_REQUIRE(h, c->u_mode >= 0 && c->u_mode <= 2, "render: bad u_mode");
  return NB_OK;
}

// one network: sampling -> MLP -> compositing [-> loss -> backward]
int run_net(nb_handle_t h, const nb_mlp_desc* d, const nb_render_cfg* c, const Plan& p, bool fine, bool train, const float* params,
            const void* packed, int64_t N, const float* rays, const float* target, void* target_ready, double n_global,
            const float* lower, const float* span, const float* t_rand, const float* u, float* grad, int accumulate, float* loss,
            float* rgb_out, float* disp_out, cudaStream_t st) {
  int rc;
  const int32_t S = fine ? c->S_c + c->S_f : c->S_c;
  float* z = fine ? p.z_f : p.z_c;
  float* w = fine ? p.w_f : p.w_c;
  if (!fine) rc = nb_stratified(h, N, c->S_c, lower, span, t_rand, c->seed, c->offset_c, c->ctr, z, st);
  else rc = nb_sample_pdf(h, N, c->S_c, c->S_f, p.z_c, p.w_c, u, c->u_mode, c->seed, c->offset_f, nullptr, nullptr, z, nullptr,
                          nullptr, nullptr, c->cdf_rows, c->ctr, st);
  if (rc) return rc;
  if ((rc = nb_mlp_forward_rays(h, d, params, packed, N, S, rays, z, p.raw, train ? p.act : nullptr, c->precision, p.mlp_ws,
                                p.mlp_ws_bytes, st))) return rc;
  if (p.z_last != nullptr && N > 0) {      // exact last-sample decision: fp32 re-evaluation of z[:, S-1] (see gather_last_kernel)
    const int blocks = nb_cdiv(N, 256);
    gather_last_kernel<<<blocks, 256, 0, st>>>((long long)N, S, z, p.z_last);
    NB_LAUNCHED(h);
    if ((rc = nb_mlp_forward_rays(h, d, params, nullptr, N, 1, rays, p.z_last, p.raw_last, nullptr, NB_FP32, p.fp32_ws, p.fp32_ws_bytes, st)))
      return rc;
    scatter_last_kernel<<<blocks, 256, 0, st>>>((long long)N, S, reinterpret_cast<const float4*>(p.raw_last), reinterpret_cast<float4*>(p.raw));
    NB_LAUNCHED(h);
  }
  float* rgb = rgb_out ? rgb_out : p.rgb_tmp;
  if (train) {      // compositing + loss gradient + compositing backward in one pass over raw (S <= 192), else stage by stage below
    if (target_ready) NB_CUDA(h, cudaStreamWaitEvent(st, (cudaEvent_t)target_ready, 0));
    const float scale = (float)(2.0 / (3.0 * n_global)), lscale = (float)(1.0 / (3.0 * n_global));
    rc = nb_composite_train(h, N, S, p.raw, z, p.rays_d, target, scale, lscale, rgb, disp_out ? disp_out : p.disp_tmp, fine ? nullptr : w,
                            p.d_raw, loss ? loss + (fine ? 1 : 0) : nullptr, st);
    if (rc == NB_OK)
      return nb_mlp_backward(h, d, params, packed, N * S, p.act, p.d_raw, grad, accumulate, c->precision, p.mlp_ws, p.mlp_ws_bytes, st);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
    target_ready = nullptr;      // already waited for
  }
  // the fine pass of render_rays drops weights/depth/acc (nerf_process.py:211-216); the coarse weights feed sample_pdf
  if ((rc = nb_composite_forward(h, N, S, p.raw, z, p.rays_d, rgb, disp_out ? disp_out : p.disp_tmp, nullptr, fine ? nullptr : w,
                                 nullptr, st))) return rc;
  if (!train) return NB_OK;
  if (target_ready) NB_CUDA(h, cudaStreamWaitEvent(st, (cudaEvent_t)target_ready, 0));
  const float scale = (float)(2.0 / (3.0 * n_global)), lscale = (float)(1.0 / (3.0 * n_global));
  if ((rc = nb_mse_grad(h, N, rgb, target, scale, lscale, p.d_rgb, loss ? loss + (fine ? 1 : 0) : nullptr, st))) return rc;
  if ((rc = nb_composite_backward(h, N, S, p.raw, z, p.rays_d, p.d_rgb, p.d_raw, st))) return rc;
  return nb_mlp_backward(h, d, params, packed, N * S, p.act, p.d_raw, grad, accumulate, c->precision, p.mlp_ws, p.mlp_ws_bytes, st);
}

int split_dirs(nb_handle_t h, const Plan& p, int64_t N, const float* rays, cudaStream_t st) {
  if (N > 0) NB_CUDA(h, cudaMemcpy2DAsync(p.rays_d, 12, rays + 3, 24, 12, (size_t)N, cudaMemcpyDeviceToDevice, st));
  return NB_OK;
}

}  // namespace

extern "C" int nb_render_workspace_bytes(nb_handle_t h, const nb_mlp_desc* d, int64_t N, const nb_render_cfg* cfg, int32_t train,
                                         size_t* out) {
  NB_ENTER(h);
  int rc = check_common(h, d, cfg, N);
  if (rc) return rc;
  NB_REQUIRE(h, out, "nb_render_workspace_bytes: NULL out");
  Plan p;
  if ((rc = make_plan(h, d, N, cfg, train, nullptr, &p))) return rc;
  *out = p.total;
  return NB_OK;
}

extern "C" int nb_render_rays(nb_handle_t h, const nb_mlp_desc* d, const nb_render_cfg* cfg, const float* params_c,
                              const void* packed_c, const float* params_f, const void* packed_f, int64_t N, const float* rays,
                              const float* lower, const float* span, const float* t_rand, const float* u, float* rgb_c,
                              float* disp_c, float* rgb_f, float* disp_f, void* ws, size_t ws_bytes, void* stream) {
  NB_ENTER(h);
  int rc = check_common(h, d, cfg, N);
  if (rc) return rc;
  if (N == 0) return NB_OK;
  NB_REQUIRE(h, params_c && rays && lower && span && ws, "nb_render_rays: NULL pointer");
  NB_REQUIRE(h, cfg->S_f == 0 || params_f, "nb_render_rays: fine network missing");
  NB_REQUIRE(h, cfg->u_mode == 2 || cfg->S_f == 0 || u, "nb_render_rays: u missing");
  Plan p;
  if ((rc = make_plan(h, d, N, cfg, 0, (void*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023), &p))) return rc;
  if (p.total > ws_bytes) { NB_SET_ERR(h, "nb_render_rays: workspace %zu < %zu", ws_bytes, p.total); return NB_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = split_dirs(h, p, N, rays, st))) return rc;
  if ((rc = run_net(h, d, cfg, p, false, false, params_c, packed_c, N, rays, nullptr, nullptr, 1.0, lower, span, t_rand, u, nullptr, 0,
                    nullptr, rgb_c, disp_c, st))) return rc;
  if (cfg->S_f > 0)
    rc = run_net(h, d, cfg, p, true, false, params_f, packed_f, N, rays, nullptr, nullptr, 1.0, lower, span, t_rand, u, nullptr, 0,
                 nullptr, rgb_f, disp_f, st);
  return rc;
}

extern "C" int nb_train_rays(nb_handle_t h, const nb_mlp_desc* d, const nb_render_cfg* cfg, const float* params_c,
                             const void* packed_c, const float* params_f, const void* packed_f, int64_t N, const float* rays,
                             const float* target, void* target_ready, int64_t n_global, const float* lower, const float* span,
                             const float* t_rand, const float* u, float* grad_c, float* grad_f, int32_t accumulate, float* loss,
                             float* rgb_c, float* disp_c, float* rgb_f, float* disp_f, int32_t nets, void* ws, size_t ws_bytes,
                             void* stream) {
  NB_ENTER(h);
  int rc = check_common(h, d, cfg, N);
  if (rc) return rc;
  NB_REQUIRE(h, (nets & 3) != 0 && (nets & ~3) == 0, "nb_train_rays: nets must be 1, 2 or 3");
  NB_REQUIRE(h, N > 0 && n_global > 0, "nb_train_rays: need N > 0 and n_global > 0");
  NB_REQUIRE(h, rays && target && lower && span && ws, "nb_train_rays: NULL pointer");
  NB_REQUIRE(h, !(nets & 1) || (params_c && grad_c), "nb_train_rays: coarse network missing");
  NB_REQUIRE(h, !(nets & 2) || (cfg->S_f > 0 && params_f && grad_f), "nb_train_rays: fine network missing");
  NB_REQUIRE(h, cfg->u_mode == 2 || !(nets & 2) || u, "nb_train_rays: u missing");
  Plan p;
  if ((rc = make_plan(h, d, N, cfg, 1, (void*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023), &p))) return rc;
  if (p.total > ws_bytes) { NB_SET_ERR(h, "nb_train_rays: workspace %zu < %zu", ws_bytes, p.total); return NB_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = split_dirs(h, p, N, rays, st))) return rc;
  if (nets & 1)
    if ((rc = run_net(h, d, cfg, p, false, true, params_c, packed_c, N, rays, target, target_ready, (double)n_global, lower, span,
                      t_rand, u, grad_c, accumulate, loss, rgb_c, disp_c, st))) return rc;
  if (nets & 2)   // a fine-only call continues from the z_c / weights a previous coarse call left in the workspace
    rc = run_net(h, d, cfg, p, true, true, params_f, packed_f, N, rays, target, (nets & 1) ? nullptr : target_ready, (double)n_global,
                 lower, span, t_rand, u, grad_f, accumulate, loss, rgb_f, disp_f, st);
  return rc;
}
