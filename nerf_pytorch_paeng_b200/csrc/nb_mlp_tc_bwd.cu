// K4 backward, NB_BF16 precision (dgrad chain + wgrad).  (placeholder until the kernels land)
#include "nb_mlp_tc.h"

size_t nb_tc_bwd_packed_bytes() { return 0; }
size_t nb_tc_bwd_ws_bytes(const nb_mlp_desc&, long long) { return 256; }
void nb_tc_bwd_add_blobs(const NbParamLayout&, const std::function<void(size_t, int, int, int, int, int, int, int)>&) {}
int nb_tc_backward(nb_handle_t h, const nb_mlp_desc*, const float*, const void*, int64_t, const void*, const float*, float*,
                   int, void*, size_t, cudaStream_t) { NB_SET_ERR(h, "bf16 backward not built"); return NB_ERR_UNSUPPORTED; }
