// Internal declarations shared by the MLP translation units.
#pragma once
#include "nb_common.cuh"

#define NB_MAX_D 16

// offsets (in floats) of each tensor inside the flat parameter / gradient buffer (nerf_b200.h, nb_mlp_desc)
struct NbParamLayout {
  size_t w[NB_MAX_D], b[NB_MAX_D];
  int in_dim[NB_MAX_D];
  size_t wd, bd, wf, bf, ws, bs, wc, bc, total;
};
NbParamLayout nb_param_layout(const nb_mlp_desc& d);
int nb_desc_check(nb_handle_t h, const nb_mlp_desc* d);

// NB_FP32 (nb_mlp_fp32.cu)
size_t nb_fp32_act_bytes(const nb_mlp_desc& d, long long P);
size_t nb_fp32_ws_bytes(const nb_mlp_desc& d, long long P, int backward);
int nb_fp32_forward(nb_handle_t h, const nb_mlp_desc* d, const float* params, int64_t P, const float* x, int64_t ld_x,
                    const float* rays, const float* z, int32_t S, float* raw_out, void* act_save, void* ws, size_t ws_bytes,
                    cudaStream_t st);
int nb_fp32_backward(nb_handle_t h, const nb_mlp_desc* d, const float* params, int64_t P, const void* act_save,
                     const float* d_raw, float* grad, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);

// NB_BF16 (nb_mlp_tc.cu)
bool nb_tc_supported(const nb_mlp_desc& d);
size_t nb_tc_packed_bytes(const nb_mlp_desc& d);
size_t nb_tc_act_bytes(const nb_mlp_desc& d, long long P);
size_t nb_tc_ws_bytes(const nb_mlp_desc& d, long long P, int backward);
int nb_tc_pack(nb_handle_t h, const nb_mlp_desc* d, const float* params, void* packed, cudaStream_t st);
int nb_tc_forward(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P, const float* x,
                  int64_t ld_x, const float* rays, const float* z, int32_t S, float* raw_out, void* act_save, void* ws,
                  size_t ws_bytes, cudaStream_t st);
// stages: bit 0 = (zero grad unless accumulate) + dgrad chain, bit 1 = wgrad
int nb_tc_backward(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                   const void* act_save, const float* d_raw, float* grad, int accumulate, void* ws, size_t ws_bytes,
                   cudaStream_t st, int stages = 3);

// fused compositing + MSE gradient + compositing backward of the training drivers (nb_composite.cu); NB_ERR_UNSUPPORTED for S > 192
int nb_composite_train(nb_handle_t h, int64_t N, int32_t S, const float* raw, const float* z, const float* rays_d, const float* target,
                       float scale, float loss_scale, float* rgb, float* disp, float* weights, float* d_raw, float* loss_out,
                       cudaStream_t st);
