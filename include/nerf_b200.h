/*
 * nerf_b200.h -- C ABI of libnerf_b200.so, the B200 (sm_100a) engine behind the
 * ray-batch hot path of nuggy875/NeRF_pytorch_paeng.
 *
 * The reference has no FFI of its own: its "plugin boundary" is the Python call
 * surface of rays.py / nerf_process.py / model/*.py.  Every
 * entry point below names the reference interface (file:line under the
 * reference tree) whose body it replaces; nerf_pytorch_paeng_b200/ holds the
 * Python mirror of those interfaces, which binds this library with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.
 *  - All data pointers are DEVICE pointers on the handle's GPU unless a
 *    parameter is documented as "host".  Tensors are contiguous row-major
 *    fp32 unless stated; index tensors are int64.
 *  - Every call returns 0 (NB_OK) or a negative nb_status; the text of the
 *    last error of a handle is nb_last_error(h).  Nothing throws.
 *  - No call allocates or frees caller-visible memory and no call
 *    synchronises the device: work is enqueued on `stream` (a cudaStream_t
 *    passed as void*; NULL = legacy default stream).  Scratch memory comes
 *    from a caller-supplied workspace whose size is queried first.
 *  - One handle per GPU per process; calls on one handle are not re-entrant.
 *  - There is no CPU fallback: without a CUDA device nb_create fails.
 */
#ifndef NERF_B200_H_
#define NERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

#define NB_ABI_VERSION 2

typedef struct nb_handle_s* nb_handle_t;

typedef enum {
  NB_OK = 0,
  NB_ERR_INVALID = -1,    /* bad argument (shape, NULL pointer, unsupported size) */
  NB_ERR_CUDA = -2,       /* a CUDA runtime call failed; see nb_last_error */
  NB_ERR_WORKSPACE = -3,  /* workspace too small */
  NB_ERR_UNSUPPORTED = -4 /* valid request this build cannot serve (e.g. not sm_100) */
} nb_status;

typedef enum {
  NB_FP32 = 0, /* CUDA-core FFMA path: the <=1e-4 parity path */
  NB_BF16 = 1  /* tcgen05 tensor-core path: bf16 operands, fp32 accumulate in TMEM */
} nb_precision;

/* Topology of one NeRFModule (model/NeRF.py:10-30): D trunk layers of width W,
 * input_ch = in_x (63), input_ch_d = in_d (27), a single skip layer index
 * (skips=[4]) after which the trunk input becomes [x, h] (NeRF.py:40-41).
 * Parameters travel as ONE flat fp32 buffer in model.parameters() order:
 *   linear_x.0.weight [W,in_x], linear_x.0.bias [W], ... linear_x.{D-1}.*,
 *   linear_d.weight [W/2, W+in_d], linear_d.bias, linear_feat.weight [W,W], .bias,
 *   linear_density.weight [1,W], .bias, linear_color.weight [3,W/2], .bias
 * (nn.Linear layout weight[out,in]).  Gradients use the same layout. */
typedef struct {
  int32_t D;     /* 8   */
  int32_t W;     /* 256 (multiple of 64) */
  int32_t in_x;  /* 63 = 3 + 6*L_x */
  int32_t in_d;  /* 27 = 3 + 6*L_d */
  int32_t skip;  /* 4, or -1 for none */
  int32_t L_x;   /* 10 */
  int32_t L_d;   /* 4  */
} nb_mlp_desc;

/* ---- lifecycle ------------------------------------------------------------------------- */
int nb_abi_version(void);
/* device: CUDA ordinal.  flags: reserved, 0.  Fails with NB_ERR_CUDA if there is no GPU and
 * NB_ERR_UNSUPPORTED if the GPU is not compute capability 10.x.
 * Threading / streams: a handle is used by one host thread at a time.  All work is stream-ordered on the stream passed
 * to each call, and calls on different streams or handles are independent.  The bf16 MLP entries stage their per-network
 * constants (biases, head weights) in a __constant__ bank bound to the calling stream (four banks per device, shared by
 * the handles of that device): up to four streams run MLP kernels concurrently with no ordering between them; a further
 * stream takes over the least recently used bank after a stream-side wait for that bank's previous owner. */
int nb_create(nb_handle_t* out, int device, unsigned flags);
int nb_destroy(nb_handle_t h);
const char* nb_last_error(nb_handle_t h);
/* info[0]=SM count, [1]=cc major, [2]=cc minor, [3]=kernels launched by this handle so far (low 31 bits) */
int nb_device_info(nb_handle_t h, int32_t info[4]);
int64_t nb_launch_count(nb_handle_t h);

/* ---- K1: ray generation ---------------------------------------------------------------- */
/* rays.py:20-34 make_o_d (+ rays.py:59-60 gather when pix_idx != NULL, + nerf_process.py:8-28
 * when NB_RAYGEN_NDC).  pose: 3x4 row-major c2w with `pose_ld` floats between rows.
 * pix_idx: N flat pixel indices r*W+c, or NULL for the full image (then N must be H*W).
 * Arithmetic: x=(c-cx)/fx, y=-((r-cy)/fy) individually rounded, K demoted to fp32;
 * d_k = fma(-1,R[k][2], fma(y,R[k][1], x*R[k][0])) (the order MKL/cuBLAS K=3 produces). */
#define NB_RAYGEN_NDC 1u
int nb_raygen_pinhole(nb_handle_t h, int32_t H, int32_t W, double fx, double fy, double cx, double cy,
                      const float* pose, int64_t pose_ld, const int64_t* pix_idx, int64_t N,
                      float* rays_o, float* rays_d, unsigned flags, double ndc_focal, double ndc_near,
                      void* stream);
/* rays.py:7-17 get_rays_np, the global-batch precompute (main.py:95-101), bit-exact: under NumPy >= 2 the float64 entries of K
 * promote the whole computation to float64, so rays_d [H*W,3] is DOUBLE here: x=(c-cx)/fx, y=-((r-cy)/fy),
 * d_k = ((x*R[k][0]) + (y*R[k][1])) + (-1*R[k][2]), each operation rounded in fp64 (no FMA).  rays_o is pose[:,3] broadcast
 * (host side).  pose: 3x4 (or 3x3) row-major fp32 with pose_ld floats between rows. */
int nb_raygen_pinhole_f64(nb_handle_t h, int32_t H, int32_t W, double fx, double fy, double cx, double cy, const float* pose,
                          int64_t pose_ld, double* rays_d, void* stream);
/* nerf_process.py:8-28 ndc_rays on arbitrary rays [N,3] (the global-batch path). In-place allowed. */
int nb_ndc_rays(nb_handle_t h, int64_t N, int32_t H, int32_t W, double focal, double near,
                const float* rays_o, const float* rays_d, float* o_out, float* d_out, void* stream);
/* rays.py:62 target gather: out[n,:] = img[pix_idx[n],:], img is [H*W,C] fp32. */
int nb_gather_rows(nb_handle_t h, int64_t N, int32_t C, const int64_t* idx, const float* src, float* out,
                   void* stream);

/* rays.py:40-54 pixel selection on the device: N DISTINCT flat pixel indices r*W+c drawn
 * uniformly from the region rows [r0,r0+nr) x cols [c0,c0+nc) (the whole image, or the precrop window),
 * out[n] = perm_seed((offset+n) mod nr*nc) with perm a keyed bijection -- no permutation array, no host work. */
int nb_select_pixels(nb_handle_t h, int64_t N, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t nr, int32_t nc,
                     uint64_t seed, uint64_t offset, int64_t* out, void* stream);

/* ---- K2: sampling ---------------------------------------------------------------------- */
/* nerf_process.py:43-60: z[n,s] = lower[s] + span[s]*t_rand[n,s]  (span = upper-lower, both [S_c],
 * computed once by the caller from torch.linspace).  t_rand NULL => in-kernel Philox4x32-10
 * keyed by (seed, offset). */
int nb_stratified(nb_handle_t h, int64_t N, int32_t S_c, const float* lower, const float* span,
                  const float* t_rand, uint64_t seed, uint64_t offset, const uint64_t* ctr, float* z_out, void* stream);
/* Philox counters may live on the device: every sampling entry takes `ctr` (device pointer to one uint64, or NULL) whose value is
 * ADDED to `offset` inside the kernel; nb_counter_add advances it in stream order.  A captured CUDA graph of a train step can
 * then be replayed without any per-step host value. */
int nb_counter_add(nb_handle_t h, uint64_t* ctr, uint64_t delta, void* stream);
/* nerf_process.py:62-67 + 144-182: mids, pdf, cdf, inverse-CDF sampling, merge with z_c, sort.
 * u_mode 0: u is [S_f] shared by all rays (det=True: torch.linspace); 1: u is [N,S_f] (injected
 * torch.rand); 2: u NULL, Philox(seed, offset).  cdf_in (optional, [N,S_c-1]) overrides the
 * kernel's own cdf (parity tests against a reference cdf).  bins_in (optional, [N,S_c-1]) gives the
 * bin positions directly instead of mids(z_c) (the stand-alone sample_pdf(bins, weights) entry,
 * nerf_process.py:144; z_c may then be NULL and z_fine must be NULL).  Optional outputs (may be NULL):
 * z_samples [N,S_f] (unsorted, pre-merge), inds [N,S_f] int64 (torch.searchsorted right=True),
 * cdf_out [N,S_c-1].
 * Summation order of nerf_process.py:150-152 (torch.sum, torch.cumsum): the reference's result depends on the device and,
 * on CUDA, on the shape of the call.  cdf_rows >= 0: the order of torch's CUDA kernels (fp32; row sum = lane-strided
 * partial sums + shuffle-down tree, cumsum = Sklansky scan in blocks whose width ATen derives from [rows, S_c-2]) for a call
 * with cdf_rows rows (0 = N; pass the reference's chunk_rays when N is a larger chunk than the reference would use), so the
 * cdf and every bin index are bit-identical to the reference on the same device (needs S_c-2 < 128, else as below);
 * cdf_rows < 0: fp64 accumulation of the row sum and the cumsum (torch's CPU cumsum; the order the CPU fixtures hold). */
int nb_sample_pdf(nb_handle_t h, int64_t N, int32_t S_c, int32_t S_f, const float* z_c, const float* weights_c,
                  const float* u, int32_t u_mode, uint64_t seed, uint64_t offset, const float* cdf_in,
                  const float* bins_in, float* z_fine, float* z_samples, int64_t* inds, float* cdf_out, int64_t cdf_rows,
                  const uint64_t* ctr, void* stream);

/* ---- K3: positional encoding (materialised form) ---------------------------------------- */
/* model/PositionalEncoding.py:29-30: out[p] = [x, sin(2^k x), cos(2^k x)]_{k<L}; out is [P, 3+6L]. */
int nb_posenc(nb_handle_t h, int64_t P, int32_t L, const float* x, float* out, void* stream);
/* nerf_process.py:34-39,69-85: embedded[n*S+s, :] = [PE_Lx(o+d*z), PE_Ld(d/|d|)], row stride ld_out
 * floats (>= 6+6Lx+6Ld); columns beyond the 90 features are left untouched. */
int nb_embed_points(nb_handle_t h, int64_t N, int32_t S, int32_t L_x, int32_t L_d, const float* rays,
                    const float* z, float* out, int64_t ld_out, void* stream);

/* ---- K4: MLP ---------------------------------------------------------------------------- */
/* Bytes of the activation stash forward() fills for backward() (P points). */
int nb_mlp_act_bytes(nb_handle_t h, const nb_mlp_desc* d, int64_t P, int32_t precision, size_t* out);
/* Scratch bytes needed by forward/backward for P points. */
int nb_mlp_workspace_bytes(nb_handle_t h, const nb_mlp_desc* d, int64_t P, int32_t precision, int32_t backward,
                           size_t* out);
/* Bytes of the packed (tile-layout bf16) copy of one net's weights and the packer itself
 * (call after load_state_dict / every optimizer step).  NB_BF16 only.  The packer also folds the activation-free feature layer
 * into the view layer (model/NeRF.py:44,47-49: W' = Wd[:, :W] . Wf, b' = Wd[:, :W] . b_feat + b_d, in fp32), which the bf16
 * forward / backward kernels use instead of two separate GEMMs; the gradients of linear_feat and linear_d are recovered exactly
 * (up to fp32 rounding) by nb_mlp_backward. */
int nb_mlp_packed_bytes(nb_handle_t h, const nb_mlp_desc* d, size_t* out);
int nb_mlp_pack(nb_handle_t h, const nb_mlp_desc* d, const float* params, void* packed, void* stream);
/* model/NeRF.py:33-52 on a materialised embedding x[P, ld_x] (first in_x+in_d columns used):
 * raw_out[P,4] = [rgb, sigma].  act_save NULL in inference. */
int nb_mlp_forward_emb(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                       const float* x, int64_t ld_x, float* raw_out, void* act_save, int32_t precision,
                       void* ws, size_t ws_bytes, void* stream);
/* Fused form used by render_rays (nerf_process.py:187-194 / 202-209): points and both encodings
 * are generated from rays[N,6], z[N,S]; the [N*S,90] tensor is never materialised in NB_BF16. */
int nb_mlp_forward_rays(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t N,
                        int32_t S, const float* rays, const float* z, float* raw_out, void* act_save,
                        int32_t precision, void* ws, size_t ws_bytes, void* stream);
/* Autograd of the above wrt the parameters (train.py:69): grad (flat, params layout) = or += dL/dparams
 * given d_raw[P,4] and the stash written by forward. */
int nb_mlp_backward(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                    const void* act_save, const float* d_raw, float* grad, int32_t accumulate, int32_t precision,
                    void* ws, size_t ws_bytes, void* stream);

/* The same backward in two separately enqueueable stages (so a caller can time them, or start the gradient
 * all-reduce of one network while the other still runs): stage 1 = zero grad (unless accumulate) + input-gradient
 * chain (writes dY tiles into ws), stage 2 = weight/bias gradients from ws.  stage 1 then stage 2 == nb_mlp_backward.
 * NB_FP32 runs entirely in stage 2. */
int nb_mlp_backward_stage(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                          const void* act_save, const float* d_raw, float* grad, int32_t accumulate, int32_t precision,
                          void* ws, size_t ws_bytes, int32_t stage, void* stream);

/* Diagnostic (NB_BF16): run the forward chain on rays/z and dump the raw fp32 TMEM accumulators of chain
 * step `step` (0..8; before bias/activation; step 8 = the view layer with the feature layer folded in) to acc_out[N*S,256]; raw_out[N*S,4] as nb_mlp_forward_rays. */
int nb_mlp_tc_probe(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t N, int32_t S,
                    const float* rays, const float* z, int32_t step, float* acc_out, float* raw_out, void* stream);

/* ---- K5: compositing -------------------------------------------------------------------- */
/* nerf_process.py:89-140 post_process: raw[N,S,4], z[N,S], rays_d[N,3] -> rgb[N,3], disp[N], acc[N],
 * weights[N,S], depth[N].  Any output except rgb may be NULL. */
int nb_composite_forward(nb_handle_t h, int64_t N, int32_t S, const float* raw, const float* z, const float* rays_d,
                         float* rgb, float* disp, float* acc, float* weights, float* depth, void* stream);
/* Autograd of post_process wrt raw for an upstream gradient on rgb_map (train.py:60-66), fused:
 * d_raw[N,S,4].  d_rgb is [N,3]. */
int nb_composite_backward(nb_handle_t h, int64_t N, int32_t S, const float* raw, const float* z, const float* rays_d,
                          const float* d_rgb, float* d_raw, void* stream);

/* ---- loss / optimizer ----------------------------------------------------- */
/* train.py:58-65 nn.MSELoss: d_rgb = scale*(rgb-target) with scale = 2/(3*N_global);
 * loss_out (device scalar, optional) += sum((rgb-target)^2) * loss_scale. */
int nb_mse_grad(nb_handle_t h, int64_t N, const float* rgb, const float* target, float scale, float loss_scale,
                float* d_rgb, float* loss_out, void* stream);
/* main.py:79-80 torch.optim.Adam(betas, eps, no weight decay) on flat buffers; step is 1-based. */
int nb_adam_step(nb_handle_t h, int64_t n, float* p, const float* g, float* m, float* v, float lr, float beta1,
                 float beta2, float eps, int32_t step, void* stream);

/* Data-parallel form: the gradient all-reduce folded into Adam's load.  srcs: HOST array of n_srcs (<= NB_MAX_RANKS)
 * device pointers to the per-rank gradient buffers in RANK order (the caller's own buffer at its rank's position, the peers'
 * copies -- delivered into this GPU's memory over NVLink by the copy engines during the backward -- elsewhere); the kernel
 * sums them in that order (bit-identical on every rank), stores the sum in g, and applies nb_adam_step's update. */
#define NB_MAX_RANKS 16
int nb_adam_step_sum(nb_handle_t h, int64_t n, float* p, float* g, const float* const* srcs, int32_t n_srcs, float* m, float* v,
                     float lr, float beta1, float beta2, float eps, int32_t step, void* stream);

/* ---- fused drivers (nb_render_rays) ------------------------------------------ */
/* Sampling configuration shared by the two drivers below.  lower/span are nb_stratified's [S_c] arrays.
 * u_mode as in nb_sample_pdf (0: u[S_f] shared = perturb==0, 1: u[N,S_f] injected, 2: Philox);
 * t_rand NULL => Philox.  offset_c / offset_f are the Philox counters of the coarse / fine draws. */
typedef struct nb_render_cfg {
  int32_t S_c, S_f;   /* N_samples_c, N_samples_f (0 = coarse only) */
  int32_t precision;  /* NB_FP32 / NB_BF16 */
  int32_t u_mode;
  uint64_t seed, offset_c, offset_f;
  int64_t cdf_rows;   /* nb_sample_pdf's cdf_rows (summation order of the fine pdf/cdf) */
  const uint64_t* ctr; /* optional device-resident Philox counter added to offset_c / offset_f (nb_counter_add) */
  int32_t exact_last;  /* NB_BF16 only: re-evaluate the LAST sample of every ray on the fp32 path before compositing.  That sample's
                        * interval is 1e10 (nerf_process.py:98), so its alpha is a step function of sign(sigma): with the flag the
                        * decision is the fp32 path's (costs one fp32 MLP pass over N points per network; default 0) */
  int32_t reserved;
} nb_render_cfg;
/* Bytes of caller workspace the drivers need for N rays (train != 0: including the activation stash). */
int nb_render_workspace_bytes(nb_handle_t h, const nb_mlp_desc* d, int64_t N, const nb_render_cfg* cfg, int32_t train,
                              size_t* out);
/* nerf_process.py:185-216 render_rays as one call: stratified -> coarse MLP -> post_process -> sample_pdf -> fine MLP
 * -> post_process, rays [N,6] (already NDC-warped for llff).  Outputs [N,3]/[N]; any may be NULL.  *_f unused if S_f == 0. */
int nb_render_rays(nb_handle_t h, const nb_mlp_desc* d, const nb_render_cfg* cfg, const float* params_c, const void* packed_c,
                   const float* params_f, const void* packed_f, int64_t N, const float* rays, const float* lower,
                   const float* span, const float* t_rand, const float* u, float* rgb_c, float* disp_c, float* rgb_f,
                   float* disp_f, void* ws, size_t ws_bytes, void* stream);
/* train.py:53-69 (render, MSE_c + MSE_f with the mean over n_global*3 values, backward) as one call; the optimizer step
 * is nb_adam_step.  grad_c / grad_f: flat gradient buffers (= or += per `accumulate`); loss[2] += {MSE_c, MSE_f}.
 * target [N,3]; target_ready: optional cudaEvent_t the stream waits on right before target is first read (lets the
 * caller's host->device copy of the target overlap the coarse forward).  nets: bit 0 coarse, bit 1 fine -- a call with
 * nets=1 followed by one with nets=2 on the same untouched workspace equals nets=3 (the caller can start the coarse
 * gradient all-reduce in between). */
int nb_train_rays(nb_handle_t h, const nb_mlp_desc* d, const nb_render_cfg* cfg, const float* params_c, const void* packed_c,
                  const float* params_f, const void* packed_f, int64_t N, const float* rays, const float* target,
                  void* target_ready, int64_t n_global, const float* lower, const float* span, const float* t_rand,
                  const float* u, float* grad_c, float* grad_f, int32_t accumulate, float* loss, float* rgb_c, float* disp_c,
                  float* rgb_f, float* disp_f, int32_t nets, void* ws, size_t ws_bytes, void* stream);

/* ---- frame output ---------------------------------------------------------- */
/* test.py:50-61 / utils.py:11: rgb8[N,3] = to8b(rgb), disp8[N] = to8b(disp / nanmax(disp)) on the device.
 * disp8 may be NULL (rgb only); disp_max_scratch is one device float used for the nanmax reduction. */
int nb_frame_to8b(nb_handle_t h, int64_t N, const float* rgb, const float* disp, float* disp_max_scratch,
                  uint8_t* rgb8, uint8_t* disp8, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H_ */
