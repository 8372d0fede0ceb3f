// K3 (materialised form): positional encoding and the [n_pts, 90] network input.
//
// Replaces model/PositionalEncoding.py:29-30 and nerf_process.py:34-39,69-85.  This is the form
// the reference API exposes (get_positional_encoder, pre_process); the tensor-core MLP generates
// the same features inside its operand producer and never materialises them (nb_mlp_tc.cu).
// One thread per OUTPUT element so stores are fully coalesced regardless of the 63/90-float row
// pitch; sinf/cosf are the accurate libdevice versions (arguments reach 2^9*|x|, SURVEY hard part 5).
#include "nb_common.cuh"

namespace {

// feature c of a (3+6L)-wide encoding of (x,y,z)
__device__ __forceinline__ float pe_feature(int c, float x, float y, float z) {
  if (c < 3) return c == 0 ? x : (c == 1 ? y : z);
  const int q = c - 3;
  const int k = q / 6, r = q - 6 * k;
  const int dim = r >= 3 ? r - 3 : r;
  const float v = dim == 0 ? x : (dim == 1 ? y : z);
  const float a = __fmul_rn(v, (float)(1 << k));     // x * freq, freq = 2^k exact
  return r >= 3 ? cosf(a) : sinf(a);
}

__global__ void __launch_bounds__(256)
posenc_kernel(long long total, int width, const float* __restrict__ x, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long p = i / width;
    const int c = (int)(i - p * width);
    out[i] = pe_feature(c, x[p * 3], x[p * 3 + 1], x[p * 3 + 2]);
  }
}

__global__ void __launch_bounds__(256)
embed_points_kernel(long long n_pts, int S, int wx, int wd, const float* __restrict__ rays,
                    const float* __restrict__ z, float* __restrict__ out, long long ld_out) {
  const int width = wx + wd;
  const long long total = n_pts * width;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long p = i / width;
    const int c = (int)(i - p * width);
    const long long n = p / S;
    const float* r = rays + n * 6;
    float v;
    if (c < wx) {
      const float zz = z[p];
      // pts = o + d*z  (un-normalised d; individually rounded, nerf_process.py:69)
      v = pe_feature(c, __fadd_rn(r[0], __fmul_rn(r[3], zz)), __fadd_rn(r[1], __fmul_rn(r[4], zz)),
                     __fadd_rn(r[2], __fmul_rn(r[5], zz)));
    } else {
      // viewdirs = d / ||d||   (nerf_process.py:38-39)
      const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
      v = pe_feature(c - wx, __fdiv_rn(r[3], nrm), __fdiv_rn(r[4], nrm), __fdiv_rn(r[5], nrm));
    }
    out[p * ld_out + c] = v;
  }
}

}  // namespace

extern "C" int nb_posenc(nb_handle_t h, int64_t P, int32_t L, const float* x, float* out, void* stream) {
  NB_ENTER(h);
  if (P == 0) return NB_OK;
  NB_REQUIRE(h, P >= 0 && L >= 0 && L <= 24 && x && out, "nb_posenc: bad arguments");
  const int width = 3 + 6 * L;
  const long long total = (long long)P * width;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)h->sm_count * 32;
  if (blocks > cap) blocks = cap;
  posenc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(total, width, x, out);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_embed_points(nb_handle_t h, int64_t N, int32_t S, int32_t L_x, int32_t L_d, const float* rays,
                               const float* z, float* out, int64_t ld_out, void* stream) {
  NB_ENTER(h);
  const int wx = 3 + 6 * L_x, wd = 3 + 6 * L_d;
  if (N == 0) return NB_OK;
  NB_REQUIRE(h, N >= 0 && S > 0 && L_x >= 0 && L_d >= 0 && L_x <= 24 && L_d <= 24 && rays && z && out && ld_out >= wx + wd,
             "nb_embed_points: bad arguments");
  const long long total = (long long)N * S * (wx + wd);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)h->sm_count * 32;
  if (blocks > cap) blocks = cap;
  embed_points_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((long long)N * S, S, wx, wd, rays, z, out,
                                                                     (long long)ld_out);
  NB_LAUNCHED(h);
  return NB_OK;
}
