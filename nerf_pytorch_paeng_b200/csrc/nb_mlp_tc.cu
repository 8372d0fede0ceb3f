// K4, NB_BF16 precision: tcgen05 / TMEM tensor-core MLP.  (placeholder until the kernels land)
#include "nb_mlp.h"

bool nb_tc_supported(const nb_mlp_desc& d) { (void)d; return false; }
size_t nb_tc_packed_bytes(const nb_mlp_desc&) { return 0; }
size_t nb_tc_act_bytes(const nb_mlp_desc&, long long) { return 0; }
size_t nb_tc_ws_bytes(const nb_mlp_desc&, long long, int) { return 0; }
int nb_tc_pack(nb_handle_t h, const nb_mlp_desc*, const float*, void*, cudaStream_t) { NB_SET_ERR(h, "bf16 path not built"); return NB_ERR_UNSUPPORTED; }
int nb_tc_forward(nb_handle_t h, const nb_mlp_desc*, const float*, const void*, int64_t, const float*, int64_t, const float*,
                  const float*, int32_t, float*, void*, void*, size_t, cudaStream_t) { NB_SET_ERR(h, "bf16 path not built"); return NB_ERR_UNSUPPORTED; }
int nb_tc_backward(nb_handle_t h, const nb_mlp_desc*, const float*, const void*, int64_t, const void*, const float*, float*,
                   int, void*, size_t, cudaStream_t) { NB_SET_ERR(h, "bf16 path not built"); return NB_ERR_UNSUPPORTED; }
