// K4, NB_FP32 precision: the 8x256 skip-connection MLP forward and backward on CUDA-core FFMA.
//
// Replaces model/NeRF.py:33-52 (NeRFModule.forward) and its autograd (train.py:69).  This is the
// parity path (max-abs error <= 1e-4 against the reference's fp32 sgemm); the throughput path is
// the tcgen05 kernel in nb_mlp_tc.cu.  One register-tiled SGEMM (128x128x16 tile, 8x8 per thread,
// split microtile so shared-memory reads are conflict-free) serves the three layouts:
//   NT  forward   C[P,out]  = act(A[P,in] . W[out,in]^T + b)        (skip/view concat = 2 accumulating calls)
//   NN  dgrad     dX[P,in]  = (dY[P,out] . W[out,in]) * (X > 0)
//   TN  wgrad     dW[out,in] += dY[P,out]^T . X[P,in]                (split-K over P, fp32 atomics)
#include "nb_common.cuh"
#include "nb_mlp.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 4;

enum : unsigned { F_BIAS = 1, F_RELU = 2, F_ACCUM = 4, F_MASK = 8, F_ATOMIC = 16 };

// A(m,k) = A_MC ? A[k*lda+m] : A[m*lda+k];  B(k,n) = B_NC ? B[k*ldb+n] : B[n*ldb+k]
template <bool A_MC, bool B_NC>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, long long lda, const float* __restrict__ B,
             long long ldb, float* __restrict__ C, long long ldc, const float* __restrict__ bias,
             const float* __restrict__ mask, long long ld_mask, unsigned flags, int k_chunk) {
  __shared__ __align__(16) float As[BK][LDS_];
  __shared__ __align__(16) float Bs[BK][LDS_];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- global -> shared (guarded scalar loads; mapping chosen for coalescing per layout)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m, k;
      if (A_MC) { m = tid & 127; k = (tid >> 7) + 2 * i; }
      else      { k = tid & 15;  m = (tid >> 4) + 16 * i; }
      const long long gm = m0 + m;
      const int gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = A_MC ? A[(long long)gk * lda + gm] : A[gm * lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int n, k;
      if (B_NC) { n = tid & 127; k = (tid >> 7) + 2 * i; }
      else      { k = tid & 15;  n = (tid >> 4) + 16 * i; }
      const int gn = n0 + n;
      const int gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = B_NC ? B[(long long)gk * ldb + gn] : B[(long long)gn * ldb + gk];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      float* c = C + gm * ldc + gn;
      if (flags & F_ATOMIC) { atomicAdd(c, v); continue; }
      if (flags & F_ACCUM) v += *c;
      if (flags & F_BIAS) v += bias[gn];
      if (flags & F_RELU) v = fmaxf(v, 0.f);
      if (flags & F_MASK) v = (mask[gm * ld_mask + gn] > 0.f) ? v : 0.f;
      *c = v;
    }
  }
}

struct Ctx {
  nb_handle_t h;
  cudaStream_t st;
};

// forward: C = act(A.W^T (+C) (+b))
int gemm_nt(Ctx& c, long long M, int N, int K, const float* A, long long lda, const float* W, long long ldw, float* C,
            long long ldc, const float* bias, unsigned flags) {
  // point tiles ride on grid.y (<= 65535 blocks): larger calls are issued as row slabs
  const long long slab = 65535LL * BM;
  for (long long m0 = 0; m0 < M; m0 += slab) {
    const long long rows = (M - m0 < slab) ? (M - m0) : slab;
    dim3 grid(nb_cdiv(N, BN), nb_cdiv(rows, BM), 1);
    sgemm_kernel<false, false><<<grid, 256, 0, c.st>>>((int)rows, N, K, A + m0 * lda, lda, W, ldw, C + m0 * ldc, ldc, bias, nullptr, 0, flags, K);
    NB_LAUNCHED(c.h);
  }
  return NB_OK;
}
// dgrad: C = (A.W) [masked by mask>0]
int gemm_nn(Ctx& c, long long M, int N, int K, const float* A, long long lda, const float* W, long long ldw, float* C,
            long long ldc, const float* mask, long long ld_mask, unsigned flags) {
  const long long slab = 65535LL * BM;
  for (long long m0 = 0; m0 < M; m0 += slab) {
    const long long rows = (M - m0 < slab) ? (M - m0) : slab;
    dim3 grid(nb_cdiv(N, BN), nb_cdiv(rows, BM), 1);
    sgemm_kernel<false, true><<<grid, 256, 0, c.st>>>((int)rows, N, K, A + m0 * lda, lda, W, ldw, C + m0 * ldc, ldc, nullptr,
                                                       mask ? mask + m0 * ld_mask : nullptr, ld_mask, flags | (mask ? F_MASK : 0), K);
    NB_LAUNCHED(c.h);
  }
  return NB_OK;
}
// wgrad: C[M=out, N=in] += dY[P,out]^T . X[P,in], split-K with atomics (C must be initialised)
int gemm_tn(Ctx& c, int M, int N, long long P, const float* dY, long long ldy, const float* X, long long ldx, float* C,
            long long ldc) {
  const int tiles = nb_cdiv(M, BM) * nb_cdiv(N, BN);
  int splits = (int)((P + 2047) / 2048);
  const int want = (4 * c.h->sm_count + tiles - 1) / tiles;
  if (splits > want) splits = want;
  if (splits < 1) splits = 1;
  int k_chunk = (int)((P + splits - 1) / splits);
  k_chunk = (k_chunk + BK - 1) / BK * BK;
  splits = (int)((P + k_chunk - 1) / k_chunk);
  dim3 grid(nb_cdiv(N, BN), nb_cdiv(M, BM), splits);
  sgemm_kernel<true, true><<<grid, 256, 0, c.st>>>(M, N, (int)P, dY, ldy, X, ldx, C, ldc, nullptr, nullptr, 0, F_ATOMIC,
                                                    k_chunk);
  NB_LAUNCHED(c.h);
  return NB_OK;
}

// bias gradient: out[j] += sum_p dY[p*ld + j]
__global__ void __launch_bounds__(256)
colsum_kernel(long long P, int ncol, const float* __restrict__ dY, long long ld, float* __restrict__ out, int rows_per_block) {
  const long long p0 = (long long)blockIdx.x * rows_per_block;
  const long long p1 = min(P, p0 + rows_per_block);
  for (int j = threadIdx.x; j < ncol; j += blockDim.x) {
    float s = 0.f;
    for (long long p = p0; p < p1; ++p) s += dY[p * ld + j];
    atomicAdd(&out[j], s);
  }
}
int colsum(Ctx& c, long long P, int ncol, const float* dY, long long ld, float* out) {
  int rpb = 512;
  colsum_kernel<<<nb_cdiv(P, rpb), ncol >= 128 ? 256 : 32, 0, c.st>>>(P, ncol, dY, ld, out, rpb);
  NB_LAUNCHED(c.h);
  return NB_OK;
}

// dh[p,j] = (h[p,j] > 0) ? dh[p,j] + dsig[p*4+3] * wsig[j] : 0        (density head + relu of the last trunk layer)
__global__ void __launch_bounds__(256)
rank1_mask_kernel(long long total, int W, float* __restrict__ dh, const float* __restrict__ h, const float* __restrict__ d_raw,
                  const float* __restrict__ wsig) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long p = i / W;
    const int j = (int)(i - p * W);
    dh[i] = (h[i] > 0.f) ? dh[i] + d_raw[p * 4 + 3] * wsig[j] : 0.f;
  }
}

#define TRY(x) do { int rc__ = (x); if (rc__ != NB_OK) return rc__; } while (0)

}  // namespace

// ---------------------------------------------------------------------------------------------
// parameter / activation layouts (shared with the tensor-core path through nb_mlp.h)
// ---------------------------------------------------------------------------------------------
NbParamLayout nb_param_layout(const nb_mlp_desc& d) {
  NbParamLayout L;
  size_t off = 0;
  for (int i = 0; i < d.D; ++i) {
    const int in = (i == 0) ? d.in_x : (d.skip >= 0 && i == d.skip + 1 ? d.W + d.in_x : d.W);
    L.in_dim[i] = in;
    L.w[i] = off; off += (size_t)d.W * in;
    L.b[i] = off; off += d.W;
  }
  L.wd = off; off += (size_t)(d.W / 2) * (d.W + d.in_d);
  L.bd = off; off += d.W / 2;
  L.wf = off; off += (size_t)d.W * d.W;
  L.bf = off; off += d.W;
  L.ws = off; off += d.W;
  L.bs = off; off += 1;
  L.wc = off; off += (size_t)3 * (d.W / 2);
  L.bc = off; off += 3;
  L.total = off;
  return L;
}

int nb_desc_check(nb_handle_t h, const nb_mlp_desc* d) {
  NB_REQUIRE(h, d, "mlp: null desc");
  NB_REQUIRE(h, d->D >= 2 && d->D <= NB_MAX_D && d->W >= 16 && d->W % 16 == 0 && d->W <= 1024 && d->in_x > 0 && d->in_d > 0 &&
                    d->skip < d->D - 1 && d->skip >= -1,
             "mlp: unsupported topology D=%d W=%d in_x=%d in_d=%d skip=%d", d->D, d->W, d->in_x, d->in_d, d->skip);
  return NB_OK;
}

namespace {

// fp32 activation stash (floats per point): emb[in_x+in_d] | h[0..D-1][W] | feat[W] | g[W/2]
struct ActLayout {
  size_t emb, h[NB_MAX_D], feat, g, total_floats;   // offsets in floats (already multiplied by P)
  int ld_emb;
};
ActLayout act_layout_train(const nb_mlp_desc& d, long long P) {
  ActLayout a;
  size_t off = 0;
  a.ld_emb = d.in_x + d.in_d;
  a.emb = off; off += (size_t)P * a.ld_emb;
  off = (off + 3) & ~(size_t)3;
  for (int i = 0; i < d.D; ++i) { a.h[i] = off; off += (size_t)P * d.W; }
  a.feat = off; off += (size_t)P * d.W;
  a.g = off; off += (size_t)P * (d.W / 2);
  a.total_floats = off;
  return a;
}
// inference: h ping-pongs between two [P,W] slots; feat reuses slot 0, g slot 1
ActLayout act_layout_infer(const nb_mlp_desc& d, long long P, bool with_emb) {
  ActLayout a;
  size_t off = 0;
  a.ld_emb = d.in_x + d.in_d;
  a.emb = 0;
  if (with_emb) { off += (size_t)P * a.ld_emb; off = (off + 3) & ~(size_t)3; }
  const size_t s0 = off, s1 = off + (size_t)P * d.W;
  for (int i = 0; i < d.D; ++i) a.h[i] = (i & 1) ? s1 : s0;
  const bool last_in_s1 = ((d.D - 1) & 1) != 0;
  a.feat = last_in_s1 ? s0 : s1;
  a.g = last_in_s1 ? s1 : s0;
  a.total_floats = off + 2 * (size_t)P * d.W;
  return a;
}

int forward_core(Ctx& c, const nb_mlp_desc& d, const float* prm, long long P, const float* emb, long long ld_emb,
                 float* base, const ActLayout& a, float* raw_out) {
  const NbParamLayout L = nb_param_layout(d);
  const int W = d.W;
  const float* in = emb;
  long long ld_in = ld_emb;
  for (int i = 0; i < d.D; ++i) {
    float* out = base + a.h[i];
    if (d.skip >= 0 && i == d.skip + 1) {
      // h = [x, h_prev] (NeRF.py:40-41): two accumulating GEMMs over the column blocks of W_i
      TRY(gemm_nt(c, P, W, d.in_x, emb, ld_emb, prm + L.w[i], L.in_dim[i], out, W, nullptr, 0));
      TRY(gemm_nt(c, P, W, W, in, ld_in, prm + L.w[i] + d.in_x, L.in_dim[i], out, W, prm + L.b[i], F_ACCUM | F_BIAS | F_RELU));
    } else {
      TRY(gemm_nt(c, P, W, L.in_dim[i], in, ld_in, prm + L.w[i], L.in_dim[i], out, W, prm + L.b[i], F_BIAS | F_RELU));
    }
    in = out;
    ld_in = W;
  }
  // density (no activation, NeRF.py:43) -> raw[:,3]; feature (no activation, :44)
  TRY(gemm_nt(c, P, 1, W, in, W, prm + L.ws, W, raw_out + 3, 4, prm + L.bs, F_BIAS));
  float* feat = base + a.feat;
  TRY(gemm_nt(c, P, W, W, in, W, prm + L.wf, W, feat, W, prm + L.bf, F_BIAS));
  // g = relu(W_d [feat, emb_d] + b_d) (NeRF.py:46-48)
  float* g = base + a.g;
  const int ldwd = W + d.in_d;
  TRY(gemm_nt(c, P, W / 2, W, feat, W, prm + L.wd, ldwd, g, W / 2, nullptr, 0));
  TRY(gemm_nt(c, P, W / 2, d.in_d, emb + d.in_x, ld_emb, prm + L.wd + W, ldwd, g, W / 2, prm + L.bd, F_ACCUM | F_BIAS | F_RELU));
  // rgb (NeRF.py:50) -> raw[:,0:3]
  TRY(gemm_nt(c, P, 3, W / 2, g, W / 2, prm + L.wc, W / 2, raw_out, 4, prm + L.bc, F_BIAS));
  return NB_OK;
}

}  // namespace

size_t nb_fp32_act_bytes(const nb_mlp_desc& d, long long P) { return act_layout_train(d, P).total_floats * sizeof(float); }

size_t nb_fp32_ws_bytes(const nb_mlp_desc& d, long long P, int backward) {
  if (backward) return ((size_t)P * d.W * 2 + (size_t)P * (d.W / 2)) * sizeof(float);
  return act_layout_infer(d, P, true).total_floats * sizeof(float);
}

int nb_fp32_forward(nb_handle_t h, const nb_mlp_desc* d, const float* params, int64_t P, const float* x, int64_t ld_x,
                    const float* rays, const float* z, int32_t S, float* raw_out, void* act_save, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  Ctx c{h, st};
  const bool from_rays = rays != nullptr;
  float* base;
  ActLayout a;
  if (act_save) {
    a = act_layout_train(*d, P);
    base = (float*)act_save;
  } else {
    a = act_layout_infer(*d, P, from_rays);
    if (ws_bytes < a.total_floats * sizeof(float) || !ws) {
      NB_SET_ERR(h, "mlp fp32 forward: workspace %zu < %zu bytes", ws_bytes, a.total_floats * sizeof(float));
      return NB_ERR_WORKSPACE;
    }
    base = (float*)ws;
  }
  const float* emb = x;
  long long ld_emb = ld_x;
  if (from_rays) {
    TRY(nb_embed_points(h, P / S, S, d->L_x, d->L_d, rays, z, base + a.emb, a.ld_emb, st));
    emb = base + a.emb;
    ld_emb = a.ld_emb;
  } else if (act_save) {
    // keep a private copy of the input for backward
    NB_CUDA(h, cudaMemcpy2DAsync(base + a.emb, a.ld_emb * sizeof(float), x, ld_x * sizeof(float),
                                 a.ld_emb * sizeof(float), (size_t)P, cudaMemcpyDeviceToDevice, st));
    emb = base + a.emb;
    ld_emb = a.ld_emb;
  }
  return forward_core(c, *d, params, P, emb, ld_emb, base, a, raw_out);
}

int nb_fp32_backward(nb_handle_t h, const nb_mlp_desc* dp, const float* prm, int64_t P, const void* act_save,
                     const float* d_raw, float* grad, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  Ctx c{h, st};
  const nb_mlp_desc& d = *dp;
  const NbParamLayout L = nb_param_layout(d);
  const ActLayout a = act_layout_train(d, P);
  const size_t need = nb_fp32_ws_bytes(d, P, 1);
  if (!ws || ws_bytes < need) {
    NB_SET_ERR(h, "mlp fp32 backward: workspace %zu < %zu bytes", ws_bytes, need);
    return NB_ERR_WORKSPACE;
  }
  const int W = d.W, W2 = d.W / 2;
  const float* base = (const float*)act_save;
  const float* emb = base + a.emb;
  const long long lde = a.ld_emb;
  float* bufA = (float*)ws;
  float* bufB = bufA + (size_t)P * W;
  float* bufG = bufB + (size_t)P * W;
  if (!accumulate) NB_CUDA(h, cudaMemsetAsync(grad, 0, L.total * sizeof(float), st));

  const float* g = base + a.g;
  const float* feat = base + a.feat;
  const float* hl = base + a.h[d.D - 1];
  // rgb head: dWc = d_rgb^T g ; dbc ; dg = (d_rgb Wc) * (g>0)
  TRY(gemm_tn(c, 3, W2, P, d_raw, 4, g, W2, grad + L.wc, W2));
  TRY(colsum(c, P, 3, d_raw, 4, grad + L.bc));
  TRY(gemm_nn(c, P, W2, 3, d_raw, 4, prm + L.wc, W2, bufG, W2, g, W2, 0));
  // view layer: dWd = dg^T [feat, emb_d] ; dbd ; dfeat = dg Wd[:, :W]
  const int ldwd = W + d.in_d;
  TRY(gemm_tn(c, W2, W, P, bufG, W2, feat, W, grad + L.wd, ldwd));
  TRY(gemm_tn(c, W2, d.in_d, P, bufG, W2, emb + d.in_x, lde, grad + L.wd + W, ldwd));
  TRY(colsum(c, P, W2, bufG, W2, grad + L.bd));
  TRY(gemm_nn(c, P, W, W2, bufG, W2, prm + L.wd, ldwd, bufA, W, nullptr, 0, 0));   // dfeat
  // feature + density heads on the last trunk activation
  TRY(gemm_tn(c, W, W, P, bufA, W, hl, W, grad + L.wf, W));
  TRY(colsum(c, P, W, bufA, W, grad + L.bf));
  TRY(gemm_tn(c, 1, W, P, d_raw + 3, 4, hl, W, grad + L.ws, W));
  TRY(colsum(c, P, 1, d_raw + 3, 4, grad + L.bs));
  TRY(gemm_nn(c, P, W, W, bufA, W, prm + L.wf, W, bufB, W, nullptr, 0, 0));
  {
    const long long total = (long long)P * W;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)h->sm_count * 32;
    if (blocks > cap) blocks = cap;
    rank1_mask_kernel<<<(int)blocks, 256, 0, st>>>(total, W, bufB, hl, d_raw, prm + L.ws);
    NB_LAUNCHED(h);
  }
  float* dh = bufB;      // gradient wrt the pre-activation of trunk layer i
  float* other = bufA;
  for (int i = d.D - 1; i >= 0; --i) {
    const bool cat = (d.skip >= 0 && i == d.skip + 1);
    const int ldw = L.in_dim[i];
    if (i == 0) {
      TRY(gemm_tn(c, W, d.in_x, P, dh, W, emb, lde, grad + L.w[0], ldw));
    } else if (cat) {
      TRY(gemm_tn(c, W, d.in_x, P, dh, W, emb, lde, grad + L.w[i], ldw));
      TRY(gemm_tn(c, W, W, P, dh, W, base + a.h[i - 1], W, grad + L.w[i] + d.in_x, ldw));
    } else {
      TRY(gemm_tn(c, W, W, P, dh, W, base + a.h[i - 1], W, grad + L.w[i], ldw));
    }
    TRY(colsum(c, P, W, dh, W, grad + L.b[i]));
    if (i > 0) {
      const float* hprev = base + a.h[i - 1];
      TRY(gemm_nn(c, P, W, W, dh, W, prm + L.w[i] + (cat ? d.in_x : 0), ldw, other, W, hprev, W, 0));
      float* t = dh; dh = other; other = t;
    }
  }
  return NB_OK;
}
