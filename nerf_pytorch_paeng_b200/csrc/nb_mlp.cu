// K4 C-ABI entry points: argument checks and dispatch on precision (include/nerf_b200.h).
#include "nb_mlp.h"

#define NB_PREC_CHECK(h, precision)                                                              \
  NB_REQUIRE(h, (precision) == NB_FP32 || (precision) == NB_BF16, "mlp: unknown precision %d", (int)(precision))

extern "C" int nb_mlp_act_bytes(nb_handle_t h, const nb_mlp_desc* d, int64_t P, int32_t precision, size_t* out) {
  NB_ENTER(h);
  int rc = nb_desc_check(h, d);
  if (rc) return rc;
  NB_REQUIRE(h, out && P >= 0, "nb_mlp_act_bytes: bad arguments");
  NB_PREC_CHECK(h, precision);
  if (precision == NB_BF16) {
    if (!nb_tc_supported(*d)) { NB_SET_ERR(h, "mlp bf16: topology not supported by the tcgen05 kernels"); return NB_ERR_UNSUPPORTED; }
    *out = nb_tc_act_bytes(*d, P);
  } else {
    *out = nb_fp32_act_bytes(*d, P);
  }
  return NB_OK;
}

extern "C" int nb_mlp_workspace_bytes(nb_handle_t h, const nb_mlp_desc* d, int64_t P, int32_t precision, int32_t backward,
                                      size_t* out) {
  NB_ENTER(h);
  int rc = nb_desc_check(h, d);
  if (rc) return rc;
  NB_REQUIRE(h, out && P >= 0, "nb_mlp_workspace_bytes: bad arguments");
  NB_PREC_CHECK(h, precision);
  if (precision == NB_BF16) {
    if (!nb_tc_supported(*d)) { NB_SET_ERR(h, "mlp bf16: topology not supported by the tcgen05 kernels"); return NB_ERR_UNSUPPORTED; }
    *out = nb_tc_ws_bytes(*d, P, backward);
  } else {
    *out = nb_fp32_ws_bytes(*d, P, backward);
  }
  return NB_OK;
}

extern "C" int nb_mlp_packed_bytes(nb_handle_t h, const nb_mlp_desc* d, size_t* out) {
  NB_ENTER(h);
  int rc = nb_desc_check(h, d);
  if (rc) return rc;
  NB_REQUIRE(h, out, "nb_mlp_packed_bytes: null out");
  if (!nb_tc_supported(*d)) { NB_SET_ERR(h, "mlp bf16: topology not supported by the tcgen05 kernels"); return NB_ERR_UNSUPPORTED; }
  *out = nb_tc_packed_bytes(*d);
  return NB_OK;
}

extern "C" int nb_mlp_pack(nb_handle_t h, const nb_mlp_desc* d, const float* params, void* packed, void* stream) {
  NB_ENTER(h);
  int rc = nb_desc_check(h, d);
  if (rc) return rc;
  NB_REQUIRE(h, params && packed, "nb_mlp_pack: null buffer");
  if (!nb_tc_supported(*d)) { NB_SET_ERR(h, "mlp bf16: topology not supported by the tcgen05 kernels"); return NB_ERR_UNSUPPORTED; }
  return nb_tc_pack(h, d, params, packed, (cudaStream_t)stream);
}

static int fwd_common(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                      const float* x, int64_t ld_x, const float* rays, const float* z, int32_t S, float* raw_out,
                      void* act_save, int32_t precision, void* ws, size_t ws_bytes, void* stream) {
  int rc = nb_desc_check(h, d);
  if (rc) return rc;
  NB_PREC_CHECK(h, precision);
  NB_REQUIRE(h, P >= 0 && params && raw_out, "mlp forward: bad arguments");
  NB_REQUIRE(h, ((uintptr_t)raw_out & 15) == 0, "mlp forward: raw_out must be 16-byte aligned");
  if (P == 0) return NB_OK;
  if (precision == NB_BF16) {
    if (!nb_tc_supported(*d)) { NB_SET_ERR(h, "mlp bf16: topology not supported by the tcgen05 kernels"); return NB_ERR_UNSUPPORTED; }
    NB_REQUIRE(h, packed, "mlp bf16 forward: packed weights required (nb_mlp_pack)");
    return nb_tc_forward(h, d, params, packed, P, x, ld_x, rays, z, S, raw_out, act_save, ws, ws_bytes, (cudaStream_t)stream);
  }
  return nb_fp32_forward(h, d, params, P, x, ld_x, rays, z, S, raw_out, act_save, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int nb_mlp_forward_emb(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                                  const float* x, int64_t ld_x, float* raw_out, void* act_save, int32_t precision,
                                  void* ws, size_t ws_bytes, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, d && x && ld_x >= d->in_x + d->in_d, "nb_mlp_forward_emb: x must have >= in_x+in_d columns");
  return fwd_common(h, d, params, packed, P, x, ld_x, nullptr, nullptr, 0, raw_out, act_save, precision, ws, ws_bytes, stream);
}

extern "C" int nb_mlp_forward_rays(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t N,
                                   int32_t S, const float* rays, const float* z, float* raw_out, void* act_save,
                                   int32_t precision, void* ws, size_t ws_bytes, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, d && rays && z && S > 0 && N >= 0, "nb_mlp_forward_rays: bad arguments");
  NB_REQUIRE(h, d->in_x == 3 + 6 * d->L_x && d->in_d == 3 + 6 * d->L_d, "nb_mlp_forward_rays: in_x/in_d must equal 3+6L");
  return fwd_common(h, d, params, packed, N * S, nullptr, 0, rays, z, S, raw_out, act_save, precision, ws, ws_bytes, stream);
}

static int bwd_common(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                      const void* act_save, const float* d_raw, float* grad, int32_t accumulate, int32_t precision,
                      void* ws, size_t ws_bytes, void* stream, int stages);

extern "C" int nb_mlp_backward(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                               const void* act_save, const float* d_raw, float* grad, int32_t accumulate, int32_t precision,
                               void* ws, size_t ws_bytes, void* stream) {
  return bwd_common(h, d, params, packed, P, act_save, d_raw, grad, accumulate, precision, ws, ws_bytes, stream, 3);
}

extern "C" int nb_mlp_backward_stage(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                                     const void* act_save, const float* d_raw, float* grad, int32_t accumulate, int32_t precision,
                                     void* ws, size_t ws_bytes, int32_t stage, void* stream) {
  if (stage != 1 && stage != 2) return NB_ERR_INVALID;
  return bwd_common(h, d, params, packed, P, act_save, d_raw, grad, accumulate, precision, ws, ws_bytes, stream, stage);
}

static int bwd_common(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                      const void* act_save, const float* d_raw, float* grad, int32_t accumulate, int32_t precision,
                      void* ws, size_t ws_bytes, void* stream, int stages) {
  NB_ENTER(h);
  int rc = nb_desc_check(h, d);
  if (rc) return rc;
  NB_PREC_CHECK(h, precision);
  NB_REQUIRE(h, P >= 0 && params && act_save && d_raw && grad, "nb_mlp_backward: bad arguments");
  if (P == 0) {
    if (!accumulate) NB_CUDA(h, cudaMemsetAsync(grad, 0, nb_param_layout(*d).total * sizeof(float), (cudaStream_t)stream));
    return NB_OK;
  }
  if (precision == NB_BF16) {
    if (!nb_tc_supported(*d)) { NB_SET_ERR(h, "mlp bf16: topology not supported by the tcgen05 kernels"); return NB_ERR_UNSUPPORTED; }
    NB_REQUIRE(h, packed, "mlp bf16 backward: packed weights required (nb_mlp_pack)");
    return nb_tc_backward(h, d, params, packed, P, act_save, d_raw, grad, accumulate, ws, ws_bytes, (cudaStream_t)stream, stages);
  }
  if (stages == 1) return NB_OK;     // fp32 path: the whole backward runs as stage 2
  return nb_fp32_backward(h, d, params, P, act_save, d_raw, grad, accumulate, ws, ws_bytes, (cudaStream_t)stream);
}
