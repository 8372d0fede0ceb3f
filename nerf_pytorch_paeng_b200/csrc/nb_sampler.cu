// K2: stratified coarse sampling and hierarchical inverse-CDF fine sampling.
//
// Replaces nerf_process.py:43-60 (coarse z with unconditional jitter) and nerf_process.py:62-67 +
// 144-182 (mids, pdf, cdf, searchsorted, gather, lerp, cat, sort).  HBM-bound: coarse 536 B/ray,
// fine 1792 B/ray.  The fine kernel is warp-cooperative: one warp owns one ray, the
// ray's z / weights / cdf live in shared memory, each lane inverts S_f/32 samples by binary search;
// the samples are sorted in registers (warp bitonic network) and rank-merged with the sorted coarse depths
// (generic sizes / unsorted z_c: shared-memory bitonic network over all S_c+S_f values).
// All contractual fp32 operations are individually rounded (no FMA contraction).
#include "nb_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// coarse: z[n,s] = lower[s] + span[s] * t_rand[n,s]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stratified_kernel(long long total4, int S, const float* __restrict__ lower, const float* __restrict__ span,
                  const float4* __restrict__ t_rand, uint2 key, unsigned long long offset,
                  const unsigned long long* __restrict__ ctr, float4* __restrict__ z_out) {
  extern __shared__ float s_ls[];  // lower[S], span[S]
  if (ctr) offset += *ctr;         // device-resident Philox counter (CUDA-graph replays advance it without host work)
  for (int i = threadIdx.x; i < S; i += blockDim.x) { s_ls[i] = lower[i]; s_ls[S + i] = span[i]; }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    float4 r;
    if (t_rand) {
      r = __ldg(&t_rand[i]);
    } else {
      unsigned long long c = offset + (unsigned long long)i;
      uint4 x = nb_philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
      r = make_float4(nb_u01(x.x), nb_u01(x.y), nb_u01(x.z), nb_u01(x.w));
    }
    int s = (int)((i * 4) % S);   // S % 4 == 0, so the 4 samples stay inside one ray
    float4 z;
    z.x = __fadd_rn(s_ls[s + 0], __fmul_rn(s_ls[S + s + 0], r.x));
    z.y = __fadd_rn(s_ls[s + 1], __fmul_rn(s_ls[S + s + 1], r.y));
    z.z = __fadd_rn(s_ls[s + 2], __fmul_rn(s_ls[S + s + 2], r.z));
    z.w = __fadd_rn(s_ls[s + 3], __fmul_rn(s_ls[S + s + 3], r.w));
    z_out[i] = z;
  }
}

// ------------------------------------------------------------------------------------------
// fine: one warp per ray
// ------------------------------------------------------------------------------------------
constexpr int kWarpsPerBlock = 4;

// per-warp smem layout (floats): z[S_c] | cdf[S_c] (S_c-1 knots; holds w during the build) | bins[S_c] | sort[P2]
// KF > 0: fast path for S_f == 32*KF <= 2*S_c: every lane owns KF consecutive samples, sorts them in registers (bitonic network
// over the warp: shuffles for strides >= KF), and the sorted samples are MERGED with the (already sorted) coarse depths by
// rank -- position = own index + number of smaller elements in the other list -- instead of sorting all S_c+S_f values.
// The result is the same sorted sequence torch.sort produces (values only; ties are indistinguishable).
template <int KF>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_kernel(long long N, int S_c, int S_f, int P2, const float* __restrict__ z_c, const float* __restrict__ w_c,
                  const float* __restrict__ u_in, int u_mode, uint2 key, unsigned long long offset,
                  const float* __restrict__ cdf_in, const float* __restrict__ bins_in, float* __restrict__ z_fine,
                  float* __restrict__ z_samples, long long* __restrict__ inds_out, float* __restrict__ cdf_out, int scan_ntx, int sum_bw,
                  const unsigned long long* __restrict__ ctr) {
  extern __shared__ float smem[];
  if (ctr) offset += *ctr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 3 * S_c + P2;
  float* sz = smem + warp * per_warp;
  float* scdf = sz + S_c;          // S_c-1 knots (index 0 is the leading zero)
  float* sbins = scdf + S_c;       // S_c-1 bin positions
  float* ssort = sbins + S_c;
  const int n_knots = S_c - 1;     // len(bins) = len(cdf)
  const int n_w = S_c - 2;         // weights[..., 1:-1]
  const int S = S_c + S_f;

  for (long long ray = (long long)blockIdx.x * kWarpsPerBlock + warp; ray < N; ray += (long long)gridDim.x * kWarpsPerBlock) {
    // ---- stage z and w (coalesced) ----
    if (z_c) for (int i = lane; i < S_c; i += 32) sz[i] = z_c[ray * S_c + i];
    __syncwarp();
    // bins = mids of z (nerf_process.py:63), or given directly (stand-alone sample_pdf entry)
    for (int i = lane; i < n_knots; i += 32)
      sbins[i] = bins_in ? bins_in[ray * n_knots + i] : __fmul_rn(0.5f, __fadd_rn(sz[i + 1], sz[i]));
    if (cdf_in) {
      for (int i = lane; i < n_knots; i += 32) scdf[i] = cdf_in[ray * n_knots + i];
      __syncwarp();
    } else {
      // w = weights[1:-1] + 1e-5 ; pdf = w / sum(w) ; cdf = [0, cumsum(pdf)]   (nerf_process.py:150-155)
      if (scan_ntx > 0) {
        // ---- the summation ORDER of torch.sum / torch.cumsum on CUDA (ATen Reduce.cuh / ScanUtils.cuh), fp32 throughout, so that the
        // cdf -- and with it every bin index -- is bit-identical to the reference running on the same device.
        for (int i = lane; i < n_w; i += 32) scdf[1 + i] = __fadd_rn(w_c[ray * S_c + 1 + i], 1e-5f);
        __syncwarp();
        // torch.sum over the last dim (n_w < 128: no input vectorisation): block_width bw = min(2^floor(log2 n_w), 32) lanes, lane x owns
        // elements x, x+bw, x+2bw, x+3bw in four accumulators combined as ((v0+v1)+v2)+v3, then a shuffle-down tree bw/2 .. 1
        float p = 0.f;
        if (lane < sum_bw) {
          const float v0 = scdf[1 + lane];
          const float v1 = lane + sum_bw < n_w ? scdf[1 + lane + sum_bw] : 0.f;
          const float v2 = lane + 2 * sum_bw < n_w ? scdf[1 + lane + 2 * sum_bw] : 0.f;
          const float v3 = lane + 3 * sum_bw < n_w ? scdf[1 + lane + 3 * sum_bw] : 0.f;
          p = __fadd_rn(__fadd_rn(__fadd_rn(v0, v1), v2), v3);
        }
        for (int o = sum_bw >> 1; o > 0; o >>= 1) p = __fadd_rn(p, __shfl_down_sync(0xffffffffu, p, o));
        const float total = __shfl_sync(0xffffffffu, p, 0);
        __syncwarp();
        for (int i = lane; i < n_w; i += 32) scdf[1 + i] = __fdiv_rn(scdf[1 + i], total);
        __syncwarp();
        // torch.cumsum: Sklansky scan over blocks of 2*scan_ntx elements, the running total added to the first element of the next block
        float* buf = scdf + 1;
        const int B = 2 * scan_ntx;
        float carry = 0.f;
        for (int c0 = 0; c0 < n_w; c0 += B) {
          const int len = min(B, n_w - c0);
          if (lane == 0 && c0 > 0) buf[c0] = __fadd_rn(buf[c0], carry);
          __syncwarp();
          for (int m = 0; (1 << m) <= scan_ntx; ++m) {
            const int sft = 1 << m;
            for (int t = lane; t < scan_ntx && t < len; t += 32) {
              const int a = ((t >> m) << (m + 1)) | sft;
              const int ti = a + (t & (sft - 1));
              if (ti < len) buf[c0 + ti] = __fadd_rn(buf[c0 + ti], buf[c0 + a - 1]);
            }
            __syncwarp();
          }
          if (c0 + B <= n_w) carry = buf[c0 + B - 1];
          __syncwarp();
        }
        if (lane == 0) scdf[0] = 0.0f;
        __syncwarp();
      } else {
      // sum and cumsum accumulate in fp64 (order-independent to fp32 precision: ATen's CPU cumsum; DESIGN.md "summation order")
      double part = 0.0;
      for (int i = lane; i < n_w; i += 32) {
        float w = __fadd_rn(w_c[ray * S_c + 1 + i], 1e-5f);
        scdf[1 + i] = w;
        part += (double)w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const float total = (float)part;
      __syncwarp();
      // inclusive scan over n_w (<=  a few hundred) values: chunks of 32 with a running carry
      double carry = 0.0;
      for (int base = 0; base < n_w; base += 32) {
        int i = base + lane;
        double v = (i < n_w) ? (double)__fdiv_rn(scdf[1 + i], total) : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          double t = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += t;
        }
        v += carry;
        carry = __shfl_sync(0xffffffffu, v, 31);
        __syncwarp();
        if (i < n_w) scdf[1 + i] = (float)v;
      }
      if (lane == 0) scdf[0] = 0.0f;
      __syncwarp();
      }
    }
    if (cdf_out) for (int i = lane; i < n_knots; i += 32) cdf_out[ray * n_knots + i] = scdf[i];

    // ---- invert the cdf for S_f samples ----
    constexpr int KFs = KF > 0 ? KF : 1;
    float own[KFs];
    auto draw_u = [&](int j) -> float {
      if (u_mode == 0) return u_in[j];
      if (u_mode == 1) return u_in[ray * S_f + j];
      const unsigned long long e = (unsigned long long)(ray * S_f + j);
      const unsigned long long c = offset + (e >> 2);
      const uint4 x = nb_philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 1u, 0u), key);
      const int q = (int)(e & 3);
      return nb_u01(q == 0 ? x.x : q == 1 ? x.y : q == 2 ? x.z : x.w);
    };
    auto invert = [&](int j, float u) -> float {
      // searchsorted(cdf, u, right=True): first index with cdf[idx] > u, in [0, n_knots]
      int lo = 0, hi = n_knots;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (scdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      const int ind = lo;
      const int below = max(0, ind - 1), above = min(n_knots - 1, ind);
      const float cb = scdf[below], ca = scdf[above];
      const float bb = sbins[below], ba = sbins[above];
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
      const float zs = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      if (z_samples) z_samples[ray * S_f + j] = zs;
      if (inds_out) inds_out[ray * S_f + j] = ind;
      return zs;
    };
    if (KF > 0) {                 // KF consecutive samples per lane (S_f == 32*KF, KF % 4 == 0)
#pragma unroll
      for (int r4 = 0; r4 < KFs; r4 += 4) {
        const int j0 = lane * KFs + r4;
        float u4[4];
        if (u_mode == 2) {        // the four draws of one Philox call belong to this lane
          const unsigned long long e = (unsigned long long)(ray * S_f + j0);
          const unsigned long long c = offset + (e >> 2);
          const uint4 x = nb_philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 1u, 0u), key);
          u4[0] = nb_u01(x.x); u4[1] = nb_u01(x.y); u4[2] = nb_u01(x.z); u4[3] = nb_u01(x.w);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) u4[q] = draw_u(j0 + q);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) if (r4 + q < KFs) own[r4 + q] = invert(j0 + q, u4[q]);
      }
    } else {
      for (int j = lane; j < S_f; j += 32) {
        const float zs = invert(j, draw_u(j));
        if (z_fine) ssort[S_c + j] = zs;
      }
    }
    if (!z_fine) { __syncwarp(); continue; }   // warp-uniform: samples only
    if (KF > 0) {
      // is z_c sorted (it is for stratified depths)?  Otherwise fall through to the full sort below.
      bool ok = true;
      for (int i = lane; i + 1 < S_c; i += 32) ok = ok && (sz[i] <= sz[i + 1]);
      ok = __all_sync(0xffffffffu, ok);
      if (ok) {
        // bitonic sort of 32*KF values, element e = lane*KF + r
#pragma unroll
        for (int k = 2; k <= 32 * KFs; k <<= 1) {
#pragma unroll
          for (int j = k >> 1; j > 0; j >>= 1) {
            if (j < KFs) {
#pragma unroll
              for (int r = 0; r < KFs; ++r) {
                const int q = r ^ j;
                if (q > r) {
                  const bool up = (((lane * KFs + r) & k) == 0);
                  const float a = own[r], b = own[q];
                  if ((a > b) == up) { own[r] = b; own[q] = a; }
                }
              }
            } else {
              const int lj = j / KFs;
#pragma unroll
              for (int r = 0; r < KFs; ++r) {
                const float other = __shfl_xor_sync(0xffffffffu, own[r], lj);
                const bool up = (((lane * KFs + r) & k) == 0);
                const bool lower = ((lane & lj) == 0);
                own[r] = (lower == up) ? fminf(own[r], other) : fmaxf(own[r], other);
              }
            }
          }
        }
        __syncwarp();                            // every lane is done reading cdf / bins: their storage now holds the samples
        float* sS = scdf;                        // S_f <= 2*S_c floats (cdf + bins regions are contiguous)
#pragma unroll
        for (int r = 0; r < KFs; ++r) sS[lane * KFs + r] = own[r];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < KFs; ++r) {           // samples: after every coarse depth <= it (stable: coarse first on ties)
          const float v = own[r];
          int lo = 0, hi = S_c;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (sz[mid] <= v) lo = mid + 1; else hi = mid; }
          ssort[lane * KFs + r + lo] = v;
        }
        for (int i = lane; i < S_c; i += 32) {   // coarse depths: after every sample < it
          const float a = sz[i];
          int lo = 0, hi = S_f;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (sS[mid] < a) lo = mid + 1; else hi = mid; }
          ssort[i + lo] = a;
        }
        __syncwarp();
        for (int i = lane; i < S; i += 32) z_fine[ray * S + i] = ssort[i];
        __syncwarp();
        continue;
      }
#pragma unroll
      for (int r = 0; r < KFs; ++r) ssort[S_c + lane * KFs + r] = own[r];
    }
    for (int i = lane; i < S_c; i += 32) ssort[i] = sz[i];
    for (int i = S + lane; i < P2; i += 32) ssort[i] = __int_as_float(0x7f800000);  // +inf padding
    __syncwarp();

    // ---- bitonic sort of P2 values (torch.sort(cat([z, z_samples])), values only) ----
    for (int k = 2; k <= P2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (P2 >> 1); t += 32) {
          int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j cleared
          int p = i | j;
          bool up = ((i & k) == 0);
          float a = ssort[i], b = ssort[p];
          if ((a > b) == up) { ssort[i] = b; ssort[p] = a; }
        }
        __syncwarp();
      }
    }
    for (int i = lane; i < S; i += 32) z_fine[ray * S + i] = ssort[i];
    __syncwarp();
  }
}

}  // namespace

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long delta) { *ctr += delta; }

extern "C" int nb_counter_add(nb_handle_t h, uint64_t* ctr, uint64_t delta, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, ctr, "nb_counter_add: NULL counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)ctr, (unsigned long long)delta);
  NB_LAUNCHED(h);
  return NB_OK;
}

extern "C" int nb_stratified(nb_handle_t h, int64_t N, int32_t S_c, const float* lower, const float* span,
                             const float* t_rand, uint64_t seed, uint64_t offset, const uint64_t* ctr, float* z_out, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, N >= 0 && S_c > 0 && S_c % 4 == 0 && S_c <= 2048 && lower && span && z_out,
             "nb_stratified: need S_c %% 4 == 0, S_c <= 2048 and non-null buffers");
  if (N == 0) return NB_OK;
  const long long total4 = (long long)N * S_c / 4;
  int blocks = (int)((total4 + 255) / 256);
  const int cap = h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  stratified_kernel<<<blocks, 256, 2 * S_c * sizeof(float), (cudaStream_t)stream>>>(
      total4, S_c, lower, span, (const float4*)t_rand, key, (unsigned long long)offset, (const unsigned long long*)ctr, (float4*)z_out);
  NB_LAUNCHED(h);
  return NB_OK;
}

// Threads per row of ATen's innermost-dim scan kernel for a [num_rows, row_size] input (ScanUtils.cuh,
// get_log_num_threads_x_inner_scan<uint32_t>, including its unsigned wrap-around for num_rows >> row_size).
static int aten_scan_threads_x(uint64_t num_rows, uint32_t row_size) {
  uint32_t lx = 0, ly = 0;
  while (((uint32_t)1 << lx) < row_size) ++lx;
  while (ly < 63 && ((uint64_t)1 << ly) < num_rows) ++ly;
  const uint32_t diff = lx - ly;
  lx = ((uint32_t)9 + diff) / (uint32_t)2;
  lx = lx < 4 ? 4 : (lx > 9 ? 9 : lx);
  return 1 << lx;
}

extern "C" int nb_sample_pdf(nb_handle_t h, int64_t N, int32_t S_c, int32_t S_f, const float* z_c,
                             const float* weights_c, const float* u, int32_t u_mode, uint64_t seed, uint64_t offset,
                             const float* cdf_in, const float* bins_in, float* z_fine, float* z_samples, int64_t* inds,
                             float* cdf_out, int64_t cdf_rows, const uint64_t* ctr, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, N >= 0 && S_c >= 3 && S_f > 0 && S_c + S_f <= 4096, "nb_sample_pdf: bad sizes");
  NB_REQUIRE(h, z_c || (bins_in && !z_fine), "nb_sample_pdf: z_c may be NULL only with bins_in and without z_fine");
  NB_REQUIRE(h, z_fine || z_samples || inds || cdf_out, "nb_sample_pdf: no output requested");
  NB_REQUIRE(h, weights_c || cdf_in, "nb_sample_pdf: need weights or cdf_in");
  NB_REQUIRE(h, u_mode >= 0 && u_mode <= 2 && (u_mode == 2 || u), "nb_sample_pdf: bad u / u_mode");
  if (N == 0) return NB_OK;
  // summation order of the pdf normalisation and the cdf: torch's CUDA order when it is restated here (row sums without input
  // vectorisation: S_c-2 < 128), else fp64 accumulation
  int scan_ntx = 0, sum_bw = 0;
  const int n_w = S_c - 2;
  if (cdf_rows >= 0 && n_w >= 1 && n_w < 128) {
    scan_ntx = aten_scan_threads_x((uint64_t)(cdf_rows > 0 ? cdf_rows : N), (uint32_t)n_w);
    sum_bw = 1;
    while (sum_bw * 2 <= n_w && sum_bw < 32) sum_bw <<= 1;
  }
  int P2 = 1;
  while (P2 < S_c + S_f) P2 <<= 1;
  const size_t smem = (size_t)kWarpsPerBlock * (3 * S_c + P2) * sizeof(float);
  NB_REQUIRE(h, smem <= 200 * 1024, "nb_sample_pdf: S_c/S_f too large for shared memory");
  typedef void (*pdf_kernel_t)(long long, int, int, int, const float*, const float*, const float*, int, uint2, unsigned long long,
                               const float*, const float*, float*, float*, long long*, float*, int, int, const unsigned long long*);
  pdf_kernel_t kern = sample_pdf_kernel<0>;
  if (z_fine && S_f <= 2 * S_c) {
    if (S_f == 128) kern = sample_pdf_kernel<4>;
    else if (S_f == 256) kern = sample_pdf_kernel<8>;
    else if (S_f == 512) kern = sample_pdf_kernel<16>;
  }
  if (smem > 48 * 1024)
    NB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (N + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  kern<<<(int)blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      (long long)N, S_c, S_f, P2, z_c, weights_c, u, u_mode, key, (unsigned long long)offset, cdf_in, bins_in, z_fine,
      z_samples, (long long*)inds, cdf_out, scan_ntx, sum_bw, (const unsigned long long*)ctr);
  NB_LAUNCHED(h);
  return NB_OK;
}
