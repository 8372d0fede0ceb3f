// K1: pinhole ray generation (+ optional pixel selection, + optional LLFF NDC warp), NDC on
// arbitrary rays, and the target-pixel row gather.
//
// Replaces rays.py:20-34 (make_o_d), rays.py:54-62 (the gathers of sample_rays_and_pixel) and
// nerf_process.py:8-28 (ndc_rays).  HBM-bound: 24 B written per ray (+8 B index read).  Every
// fp32 operation is individually rounded (__f*_rn intrinsics: no FMA contraction) because ray
// origins/directions are a bit-exact contract; the one place the reference itself produces an
// FMA chain (the K=3 GEMM) uses explicit fmaf.
#include "nb_common.cuh"

namespace {

constexpr int kRayThreads = 256;

struct NdcConst { float c_w, c_h, near, two_near, neg_two_near; };

__device__ __forceinline__ void ndc_warp(const NdcConst& c, float& ox, float& oy, float& oz,
                                         float& dx, float& dy, float& dz) {
  // t = -(near + o_z)/d_z ; o = o + t*d                      (nerf_process.py:11-12)
  float t = -__fdiv_rn(__fadd_rn(c.near, oz), dz);
  ox = __fadd_rn(ox, __fmul_rn(t, dx));
  oy = __fadd_rn(oy, __fmul_rn(t, dy));
  oz = __fadd_rn(oz, __fmul_rn(t, dz));
  // projection                                               (nerf_process.py:15-23)
  float o0 = __fdiv_rn(__fmul_rn(c.c_w, ox), oz);
  float o1 = __fdiv_rn(__fmul_rn(c.c_h, oy), oz);
  float o2 = __fadd_rn(1.0f, __fdiv_rn(c.two_near, oz));
  float d0 = __fmul_rn(c.c_w, __fsub_rn(__fdiv_rn(dx, dz), __fdiv_rn(ox, oz)));
  float d1 = __fmul_rn(c.c_h, __fsub_rn(__fdiv_rn(dy, dz), __fdiv_rn(oy, oz)));
  float d2 = __fdiv_rn(c.neg_two_near, oz);
  ox = o0; oy = o1; oz = o2; dx = d0; dy = d1; dz = d2;
}

// One ray per thread; the block's 256x3 floats of o and d are staged in shared memory and
// written back as fully coalesced 4-byte-per-lane rows (a [N,3] tensor has no 16-byte alignment
// per row, but a block's slab of 768 floats starts 16-byte aligned, so float4 stores are used).
__global__ void __launch_bounds__(kRayThreads)
raygen_kernel(int H, int W, float fx, float fy, float cx, float cy, const float* __restrict__ pose,
              long long pose_ld, const long long* __restrict__ pix_idx, long long N,
              float* __restrict__ rays_o, float* __restrict__ rays_d, int do_ndc, NdcConst ndc) {
  __shared__ __align__(16) float so[kRayThreads * 3];
  __shared__ __align__(16) float sd[kRayThreads * 3];
  __shared__ float sp[12];
  if (threadIdx.x < 12) sp[threadIdx.x] = pose[(threadIdx.x / 4) * pose_ld + (threadIdx.x % 4)];
  __syncthreads();
  const long long base = (long long)blockIdx.x * kRayThreads;
  const long long n = base + threadIdx.x;
  if (n < N) {
    long long p = pix_idx ? pix_idx[n] : n;
    int r = (int)(p / W), c = (int)(p - (long long)r * W);
    float x = __fdiv_rn(__fsub_rn((float)c, cx), fx);          // (i - K[0][2]) / K[0][0]
    float y = -__fdiv_rn(__fsub_rn((float)r, cy), fy);         // -(j - K[1][2]) / K[1][1]
    float ox = sp[3], oy = sp[7], oz = sp[11];
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)                                // dirs @ R^T, k-ordered FMA chain
      d[k] = fmaf(-1.0f, sp[k * 4 + 2], fmaf(y, sp[k * 4 + 1], __fmul_rn(x, sp[k * 4 + 0])));
    if (do_ndc) ndc_warp(ndc, ox, oy, oz, d[0], d[1], d[2]);
    so[threadIdx.x * 3 + 0] = ox; so[threadIdx.x * 3 + 1] = oy; so[threadIdx.x * 3 + 2] = oz;
    sd[threadIdx.x * 3 + 0] = d[0]; sd[threadIdx.x * 3 + 1] = d[1]; sd[threadIdx.x * 3 + 2] = d[2];
  }
  __syncthreads();
  const long long rem = N - base;
  const int nfl = (int)(rem < kRayThreads ? rem : kRayThreads) * 3;
  float* go = rays_o + base * 3;
  float* gd = rays_d + base * 3;
  if (nfl == kRayThreads * 3 && (((uintptr_t)go | (uintptr_t)gd) & 15) == 0) {
    for (int i = threadIdx.x; i < kRayThreads * 3 / 4; i += kRayThreads) {
      reinterpret_cast<float4*>(go)[i] = reinterpret_cast<const float4*>(so)[i];
      reinterpret_cast<float4*>(gd)[i] = reinterpret_cast<const float4*>(sd)[i];
    }
  } else {
    for (int i = threadIdx.x; i < nfl; i += kRayThreads) { go[i] = so[i]; gd[i] = sd[i]; }
  }
}

__global__ void __launch_bounds__(kRayThreads)
ndc_kernel(long long N, const float* __restrict__ o_in, const float* __restrict__ d_in,
           float* __restrict__ o_out, float* __restrict__ d_out, NdcConst ndc) {
  __shared__ float so[kRayThreads * 3];
  __shared__ float sd[kRayThreads * 3];
  const long long base = (long long)blockIdx.x * kRayThreads;
  const long long rem = N - base;
  const int nr = (int)(rem < kRayThreads ? rem : kRayThreads);
  for (int i = threadIdx.x; i < nr * 3; i += kRayThreads) { so[i] = o_in[base * 3 + i]; sd[i] = d_in[base * 3 + i]; }
  __syncthreads();
  float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
  if ((int)threadIdx.x < nr) {
    ox = so[threadIdx.x * 3]; oy = so[threadIdx.x * 3 + 1]; oz = so[threadIdx.x * 3 + 2];
    dx = sd[threadIdx.x * 3]; dy = sd[threadIdx.x * 3 + 1]; dz = sd[threadIdx.x * 3 + 2];
    ndc_warp(ndc, ox, oy, oz, dx, dy, dz);
  }
  __syncthreads();
  if ((int)threadIdx.x < nr) {
    so[threadIdx.x * 3] = ox; so[threadIdx.x * 3 + 1] = oy; so[threadIdx.x * 3 + 2] = oz;
    sd[threadIdx.x * 3] = dx; sd[threadIdx.x * 3 + 1] = dy; sd[threadIdx.x * 3 + 2] = dz;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nr * 3; i += kRayThreads) { o_out[base * 3 + i] = so[i]; d_out[base * 3 + i] = sd[i]; }
}

__global__ void gather_rows_kernel(long long N, int C, const long long* __restrict__ idx,
                                   const float* __restrict__ src, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  long long n = i / C;
  int c = (int)(i - n * C);
  out[i] = src[idx[n] * C + c];
}

// Random pixel selection without replacement (rays.py:40-54) with no permutation array:
// a keyed Feistel network is a bijection on [0, 2^bits); cycle-walking restricts it to [0, D), so distinct
// counters give distinct pixels.  (The reference draws np.random.choice on the host; only the distribution --
// a uniformly random subset -- is contractual.)
__device__ __forceinline__ uint32_t feistel_perm(uint32_t x, int half_bits, uint32_t k0, uint32_t k1) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t f = (r + k0) * 0x9E3779B1u + (k1 ^ (0x85EBCA6Bu * (uint32_t)(i + 1)));
    f ^= f >> 15; f *= 0x2C1B3C6Du; f ^= f >> 12;
    const uint32_t nl = r;
    r = (l ^ f) & mask;
    l = nl;
  }
  return (l << half_bits) | r;
}

__global__ void select_pixels_kernel(long long N, int W, int r0, int c0, int nc, uint32_t D, int half_bits, uint32_t k0,
                                     uint32_t k1, unsigned long long offset, long long* __restrict__ out) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  uint32_t y = (uint32_t)((offset + (unsigned long long)n) % D);
  do { y = feistel_perm(y, half_bits, k0, k1); } while (y >= D);
  const int rr = (int)(y / (uint32_t)nc), cc = (int)(y % (uint32_t)nc);
  out[n] = (long long)(r0 + rr) * W + (c0 + cc);
}

NdcConst make_ndc(int H, int W, double focal, double near) {
  NdcConst c;
  c.c_w = (float)(-1. / (W / (2. * focal)));   // evaluated in double, then rounded
  c.c_h = (float)(-1. / (H / (2. * focal)));
  c.near = (float)near;
  c.two_near = (float)(2. * near);
  c.neg_two_near = (float)(-2. * near);
  return c;
}

}  // namespace

extern "C" int nb_raygen_pinhole(nb_handle_t h, int32_t H, int32_t W, double fx, double fy, double cx, double cy,
                                 const float* pose, int64_t pose_ld, const int64_t* pix_idx, int64_t N,
                                 float* rays_o, float* rays_d, unsigned flags, double ndc_focal, double ndc_near,
                                 void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, H > 0 && W > 0 && pose && rays_o && rays_d && pose_ld >= 4, "nb_raygen_pinhole: bad arguments");
  NB_REQUIRE(h, N >= 0 && (pix_idx || N == (int64_t)H * W), "nb_raygen_pinhole: N must be H*W without pix_idx");
  if (N == 0) return NB_OK;
  NdcConst ndc = make_ndc(H, W, ndc_focal == 0. ? 1. : ndc_focal, ndc_near);
  raygen_kernel<<<nb_cdiv(N, kRayThreads), kRayThreads, 0, (cudaStream_t)stream>>>(
      H, W, (float)fx, (float)fy, (float)cx, (float)cy, pose, (long long)pose_ld, (const long long*)pix_idx,
      (long long)N, rays_o, rays_d, (flags & NB_RAYGEN_NDC) ? 1 : 0, ndc);
  NB_LAUNCHED(h);
  return NB_OK;
}

// rays.py:7-17 get_rays_np as NumPy >= 2 evaluates it: float32 pixel grids and pose, float64 K scalars (NEP 50) => every
// operation in float64, individually rounded: d_k = ((x*R[k][0]) + (y*R[k][1])) + (z*R[k][2]) with z = -1.
__global__ void __launch_bounds__(256)
raygen_f64_kernel(int W, double fx, double fy, double cx, double cy, const float* __restrict__ pose, long long pose_ld, long long N,
                  double* __restrict__ rays_d) {
  __shared__ double sr[9];
  if (threadIdx.x < 9) sr[threadIdx.x] = (double)pose[(threadIdx.x / 3) * pose_ld + (threadIdx.x % 3)];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one output component per thread: coalesced 8-byte stores
  if (i >= 3 * N) return;
  const long long p = i / 3;
  const int k = (int)(i - 3 * p);
  const int r = (int)(p / W), c = (int)(p - (long long)r * W);
  const double x = __ddiv_rn(__dsub_rn((double)c, cx), fx);
  const double y = -__ddiv_rn(__dsub_rn((double)r, cy), fy);
  rays_d[i] = __dadd_rn(__dadd_rn(__dmul_rn(x, sr[k * 3 + 0]), __dmul_rn(y, sr[k * 3 + 1])), __dmul_rn(-1.0, sr[k * 3 + 2]));
}

extern "C" int nb_raygen_pinhole_f64(nb_handle_t h, int32_t H, int32_t W, double fx, double fy, double cx, double cy, const float* pose,
                                     int64_t pose_ld, double* rays_d, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, H > 0 && W > 0 && pose && rays_d && pose_ld >= 3 && fx != 0. && fy != 0., "nb_raygen_pinhole_f64: bad arguments");
  const long long N = (long long)H * W;
  raygen_f64_kernel<<<nb_cdiv(3 * N, 256), 256, 0, (cudaStream_t)stream>>>(W, fx, fy, cx, cy, pose, (long long)pose_ld, N, rays_d);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_ndc_rays(nb_handle_t h, int64_t N, int32_t H, int32_t W, double focal, double near,
                           const float* rays_o, const float* rays_d, float* o_out, float* d_out, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, N >= 0 && H > 0 && W > 0 && focal != 0. && rays_o && rays_d && o_out && d_out, "nb_ndc_rays: bad arguments");
  if (N == 0) return NB_OK;
  ndc_kernel<<<nb_cdiv(N, kRayThreads), kRayThreads, 0, (cudaStream_t)stream>>>(
      (long long)N, rays_o, rays_d, o_out, d_out, make_ndc(H, W, focal, near));
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_gather_rows(nb_handle_t h, int64_t N, int32_t C, const int64_t* idx, const float* src, float* out,
                              void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, N >= 0 && C > 0 && idx && src && out, "nb_gather_rows: bad arguments");
  if (N == 0) return NB_OK;
  gather_rows_kernel<<<nb_cdiv(N * C, 256), 256, 0, (cudaStream_t)stream>>>((long long)N, C, (const long long*)idx, src, out);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_select_pixels(nb_handle_t h, int64_t N, int32_t H, int32_t W, int32_t r0, int32_t c0, int32_t nr, int32_t nc,
                                uint64_t seed, uint64_t offset, int64_t* out, void* stream) {
  NB_ENTER(h);
  if (N == 0) return NB_OK;
  NB_REQUIRE(h, N > 0 && out && H > 0 && W > 0 && nr > 0 && nc > 0 && r0 >= 0 && c0 >= 0 && r0 + nr <= H && c0 + nc <= W,
             "nb_select_pixels: bad region");
  const unsigned long long D = (unsigned long long)nr * nc;
  NB_REQUIRE(h, (unsigned long long)N <= D && D < (1ull << 31), "nb_select_pixels: N must be <= region size < 2^31");
  int bits = 2;
  while ((1ull << bits) < D) ++bits;
  if (bits & 1) ++bits;
  select_pixels_kernel<<<nb_cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(
      (long long)N, W, r0, c0, nc, (uint32_t)D, bits / 2, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0xA511E9B3u,
      (unsigned long long)offset, (long long*)out);
  NB_LAUNCHED(h);
  return NB_OK;
}
