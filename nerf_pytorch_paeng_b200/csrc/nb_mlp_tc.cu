// K3+K4, NB_BF16 precision: the whole 8x256 skip MLP of one 128-point tile as a chain of tcgen05
// GEMMs whose activations never leave the SM.
//
// Replaces nerf_process.py:69-84 (points + positional encoding) and model/NeRF.py:33-52.
//
// One persistent CTA per SM, 320 threads, two 128-point tiles ("slots") in flight:
//   warp 0      weight producer: streams the pre-swizzled bf16 weight blobs (32 KB = 256 out-features x 64
//               in-features) from L2 into a 2-stage shared-memory ring with cp.async.bulk (TMA unit) + mbarrier.
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=256|128, K=16) with A = the slot's
//               activation tile in shared memory (K-major, 128B swizzle), B = the weight stage, D = the
//               slot's 256 fp32 TMEM columns; tcgen05.commit releases the weight stage / signals the epilogue.
//   warps 2-5   epilogue of slot 0, warps 6-9 epilogue of slot 1 (thread = one point): build the
//               positional encoding of the point straight into the A-operand tile (K3 fused into the first
//               GEMM's operand), then per layer tcgen05.ld the accumulators, add bias, ReLU, round to bf16
//               and write the next layer's A tile in place; sigma and rgb heads are CUDA-core dot products
//               on the fp32 accumulators; raw[N,S,4] is the only HBM write in inference.
// While the epilogue of one slot runs, the tensor core works on the other slot (ping-pong).
// Layer chain ("steps") per tile, K-blocks of 64:  0: PE63->256 | 1-4: 256->256 | 5: [PE63,256]->256 |
// 6,7: 256->256 (+sigma head after 7) | 8: feature 256->256 (no ReLU) | 9: [feat256,PEd27]->128 (+rgb head).
//
// In training the bf16 A tiles (layer inputs) are additionally bulk-stored to HBM as 16 KB swizzled blobs
// which the backward kernels (nb_mlp_tc_bwd.cu) consume directly as UMMA operands.
#include <stdlib.h>
#include "nb_mlp.h"
#include "nb_tc_common.cuh"
#include "nb_mlp_tc.h"

using namespace tc;

namespace {

// ------------------------------------------------------------------------------------------
// forward chain tables (D=8, W=256, skip=4, in_x=63, in_d=27)
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fwd_nkb(int s) { return (s == 0) ? 1 : ((s == 5 || s == 9) ? 5 : 4); }
__host__ __device__ constexpr int fwd_n(int s) { return s == 9 ? 128 : 256; }
__host__ __device__ constexpr uint32_t fwd_blob_bytes(int s) { return (uint32_t)fwd_n(s) * 128u; }
__host__ __device__ constexpr uint32_t fwd_w_off(int s) {   // byte offset of step s's first blob
  uint32_t o = 0;
  for (int i = 0; i < s; ++i) o += (uint32_t)fwd_nkb(i) * fwd_blob_bytes(i);
  return o;
}
// A source of K-block kb of step s: -1 = aux tile, else act K-block index
__host__ __device__ constexpr int fwd_a_src(int s, int kb) {
  if (s == 0) return -1;
  if (s == 5) return kb == 0 ? -1 : kb - 1;
  if (s == 9) return kb == 4 ? -1 : kb;
  return kb;
}

// shared-memory carve-up (bytes, after 1024-byte alignment)
constexpr uint32_t kActBytes = 4 * kBlobBytes;               // 64 KB per slot
constexpr uint32_t kOffAct = 0;                              // [2][4][16 KB]
constexpr uint32_t kOffAux = 2 * kActBytes;                  // [2][16 KB]
constexpr uint32_t kOffW = kOffAux + 2 * kBlobBytes;         // [2][32 KB]
constexpr uint32_t kOffBar = kOffW + 2 * 32768;              // barriers (256 B)
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;        // + alignment slack

constexpr int kThreads = 320;
constexpr int kBarEpi0 = 1;

__constant__ TcSmall c_fw;   // small fp32 parameters of the network being run (see nb_mlp_tc.h)   // named barrier ids of the two epilogue groups

struct FwdParams {
  const float* rays;      // [N,6]
  const float* z;         // [N,S]
  const float* x_emb;     // optional materialised embedding [P, ld_x] (forward_emb entry) or nullptr
  long long ld_x;
  long long P;            // points
  int S;
  const uint8_t* wpk;     // packed forward blobs
  const float* prm;       // flat fp32 params (biases, sigma / rgb heads)
  NbParamLayout L;
  float* raw;             // [P,4]
  uint8_t* stash;         // training: activation blobs, else nullptr
  TcStash st;
  float* dbg;             // optional [P,256] accumulator dump of step dbg_step
  int dbg_step;
  long long* prof;        // optional per-CTA cycle counters [grid][8] (NB_TC_PROF diagnostic)
  int share_w;            // 1: each weight K-block is loaded once and used by both slots (k-interleaved), 0: ping-pong
  int abl;                // ablation bits for profiling experiments (NB_TC_ABLATE env): 1 no masks, 2 no stash stores
};

// positional-encoding features of one 3-vector, written as bf16 into a swizzled 128-byte row.
// Arguments are reduced in "turns": sin(2^k x) = sin(2 pi frac(2^k x / 2 pi)), exact power-of-two scaling,
// so the fast sin.approx/cos.approx see |arg| <= pi (abs error ~1e-6, far below a bf16 ulp).
template <int L>
__device__ __forceinline__ void pe_row_to_smem(uint32_t row_addr, uint32_t r, float x, float y, float z, bool one_pad) {
  constexpr int NF = 3 + 6 * L;
  constexpr int NCH = (NF + 8) / 8;          // chunks that hold features (+ the optional 1.0 pad column)
  float e[NCH * 8];
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) e[i] = 0.f;
  e[0] = x; e[1] = y; e[2] = z;
  const float inv2pi = 0.15915494309189535f;
  const float tx = x * inv2pi, ty = y * inv2pi, tz = z * inv2pi;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const float sc = (float)(1 << k);
    float t[3] = {tx * sc, ty * sc, tz * sc};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float fr = t[d] - rintf(t[d]);
      const float a = fr * 6.283185307179586f;
      e[3 + 6 * k + d] = __sinf(a);
      e[3 + 6 * k + 3 + d] = __cosf(a);
    }
  }
  if (one_pad) e[NF] = 1.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (c < NCH) {
      w0 = pack_bf16(e[c * 8 + 0], e[c * 8 + 1]); w1 = pack_bf16(e[c * 8 + 2], e[c * 8 + 3]);
      w2 = pack_bf16(e[c * 8 + 4], e[c * 8 + 5]); w3 = pack_bf16(e[c * 8 + 6], e[c * 8 + 7]);
    }
    st_shared_v4(row_addr + (((uint32_t)c ^ (r & 7u)) << 4), w0, w1, w2, w3);
  }
}

// copy 64 bf16 features (cols [c0, c0+ncol) of a materialised fp32 embedding row) into a swizzled row
__device__ __forceinline__ void emb_row_to_smem(uint32_t row_addr, uint32_t r, const float* src, int ncol, bool one_pad) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int f = c * 8 + j;
      v[j] = f < ncol ? src[f] : ((one_pad && f == ncol) ? 1.0f : 0.0f);
    }
    st_shared_v4(row_addr + (((uint32_t)c ^ (r & 7u)) << 4), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// One layer's epilogue over `nchunks` groups of 32 accumulator columns (thread = one point / TMEM lane).
// KIND 0: bias+ReLU -> bf16 A tile | 1: same + sigma head (step 7) | 2: view layer + rgb head (step 9; A tile only
// written in training, for the stash) | 3: feature layer, no activation (step 8).
template <bool TRAIN, bool DBG, int KIND>
__device__ __forceinline__ void epi_chunks(const FwdParams& p, int s, int nchunks, uint32_t t_addr, uint32_t act_base, uint32_t r,
                                           long long pt, bool valid, uint32_t* mdst, float& sigma, float (&rgb)[3]) {
  const float* bias = c_fw.bias[s];
#pragma unroll 1
  for (int c32 = 0; c32 < nchunks; ++c32) {
    float v[32];
    tmem_ld32(t_addr + (uint32_t)c32 * 32u, v);
    tmem_ld_wait();
    if (DBG) {
      if (p.dbg != nullptr && s == p.dbg_step && valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) p.dbg[pt * 256 + c32 * 32 + j] = v[j];
      }
    }
    const float* bc = bias + c32 * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += bc[j];
    if (TRAIN && KIND != 3) {   // ReLU mask of this layer's output for the backward chain (bit j = column c32*32+j > 0)
      uint32_t m = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) m |= (v[j] > 0.f) ? (1u << j) : 0u;
      mdst[c32] = m;
    }
    if (KIND == 1) {            // sigma head on the fp32 post-ReLU trunk output (NeRF.py:43)
      const float* w = c_fw.ws + c32 * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) sigma = fmaf(fmaxf(v[j], 0.f), w[j], sigma);
    }
    if (KIND == 2) {            // rgb head on the fp32 post-ReLU view features (NeRF.py:50)
      const float* w = c_fw.wc + c32 * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float hj = fmaxf(v[j], 0.f);
        rgb[0] = fmaf(hj, w[j], rgb[0]); rgb[1] = fmaf(hj, w[128 + j], rgb[1]); rgb[2] = fmaf(hj, w[256 + j], rgb[2]);
      }
    }
    if (KIND != 2 || TRAIN) {
      // next layer's A operand (bf16, swizzled K-major): columns c32*32.. -> K-block c32/2, chunks (c32&1)*4..+3
      const uint32_t row_addr = act_base + (uint32_t)(c32 >> 1) * kBlobBytes + r * 128u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = (uint32_t)((c32 & 1) * 4 + j);
        uint32_t w0, w1, w2, w3;
        if (KIND == 3) {        // feature layer: no activation (NeRF.py:44)
          w0 = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]); w1 = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
          w2 = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]); w3 = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
        } else {
          w0 = pack_bf16_relu(v[j * 8 + 0], v[j * 8 + 1]); w1 = pack_bf16_relu(v[j * 8 + 2], v[j * 8 + 3]);
          w2 = pack_bf16_relu(v[j * 8 + 4], v[j * 8 + 5]); w3 = pack_bf16_relu(v[j * 8 + 6], v[j * 8 + 7]);
        }
        st_shared_v4(row_addr + ((c ^ (r & 7u)) << 4), w0, w1, w2, w3);
      }
    }
  }
}

template <bool TRAIN, bool DBG>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fwd_chain_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_act = sbase + kOffAct, s_aux = sbase + kOffAux, s_w = sbase + kOffW, s_bar = sbase + kOffBar;
  // barriers (8 bytes each): w_full[2] w_empty[2] a_ready[2] acc_ready[2] ; tmem ptr at +64
  const uint32_t b_wfull = s_bar, b_wempty = s_bar + 16, b_aready = s_bar + 32, b_accready = s_bar + 48, s_tmem = s_bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const long long n_tiles = (p.P + 127) / 128;
  // slot s of CTA b owns tiles (2*b + s) + i * 2 * gridDim.x
  const long long tile_stride = 2LL * gridDim.x;
  auto tile_of = [&](int slot, long long it) { return 2LL * blockIdx.x + slot + it * tile_stride; };
  const long long max_it = (n_tiles + tile_stride - 1) / tile_stride;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_wfull + 8 * i, 1);
      mbar_init(b_wempty + 8 * i, 1);
      mbar_init(b_aready + 8 * i, 128);
      mbar_init(b_accready + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

  if (warp == 0) {
    // ============================== weight producer ==============================
    // every weight K-block is loaded ONCE per iteration and consumed by both slots back to back (halves the
    // L2->SM weight stream per tile); a stage is filled by 4 concurrent bulk copies
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      long long prof_acc[1] = {0};
      for (long long it = 0; it < max_it; ++it) {
        if (tile_of(0, it) >= n_tiles) break;
#pragma unroll 1
        for (int s = 0; s < kFwdSteps; ++s) {
          const uint32_t bytes = fwd_blob_bytes(s);
          const uint8_t* src = p.wpk + fwd_w_off(s);
          const int reps = (p.share_w || tile_of(1, it) >= n_tiles) ? 1 : 2;
          for (int rep = 0; rep < reps; ++rep)
            for (int kb = 0; kb < fwd_nkb(s); ++kb) {
              { const long long t0 = clock64(); mbar_wait(b_wempty + 8 * stage, phase ^ 1); prof_acc[0] += clock64() - t0; }
              mbar_expect_tx(b_wfull + 8 * stage, bytes);
              const uint32_t q4 = bytes >> 2;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                bulk_g2s(s_w + stage * 32768u + i * q4, src + (size_t)kb * bytes + i * q4, q4, b_wfull + 8 * stage);
              stage ^= 1; if (stage == 0) phase ^= 1;
            }
        }
      }
      (void)prof_acc;
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    uint32_t stage = 0, phase = 0, par_a[2] = {0, 0};
    long long pa = 0, pw = 0;
    const long long tstart = clock64();
    for (long long it = 0; it < max_it; ++it) {
      const bool v0 = tile_of(0, it) < n_tiles, v1 = tile_of(1, it) < n_tiles;
      if (!v0) break;
#pragma unroll 1
      for (int s = 0; s < kFwdSteps; ++s) {
        const uint32_t idesc = umma_idesc(128, fwd_n(s), 0, 0);
        const int nkb = fwd_nkb(s);
        if (!p.share_w) {
          // ping-pong: slot 0's whole step, then slot 1's (its epilogue overlaps the other slot's MMAs)
          for (int slot = 0; slot < 2; ++slot) {
            if (slot == 1 && !v1) break;
            { const long long t0 = clock64(); mbar_wait(b_aready + 8 * slot, par_a[slot]); pa += clock64() - t0; }
            par_a[slot] ^= 1;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
            for (int kb = 0; kb < nkb; ++kb) {
              { const long long t0 = clock64(); mbar_wait(b_wfull + 8 * stage, phase); pw += clock64() - t0; }
              tc_fence_after();
              if (lane == 0) {
                const int src = fwd_a_src(s, kb);
                const uint32_t a_addr = (src < 0) ? (s_aux + slot * kBlobBytes) : (s_act + slot * kActBytes + (uint32_t)src * kBlobBytes);
                const uint32_t b_addr = s_w + stage * 32768u;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  umma_ss(d_tmem, umma_desc(a_addr + k4 * 32u, 16, 1024), umma_desc(b_addr + k4 * 32u, 16, 1024), idesc,
                          (kb | k4) ? 1u : 0u);
                umma_commit(b_wempty + 8 * stage);
                if (kb == nkb - 1) umma_commit(b_accready + 8 * slot);
              }
              __syncwarp();
              stage ^= 1; if (stage == 0) phase ^= 1;
            }
          }
          continue;
        }
        mbar_wait(b_aready + 0, par_a[0]); par_a[0] ^= 1;
        if (v1) { mbar_wait(b_aready + 8, par_a[1]); par_a[1] ^= 1; }
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(b_wfull + 8 * stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const int src = fwd_a_src(s, kb);
            const uint32_t b_addr = s_w + stage * 32768u;
#pragma unroll
            for (int slot = 0; slot < 2; ++slot) {
              if (slot == 1 && !v1) break;
              const uint32_t a_addr = (src < 0) ? (s_aux + slot * kBlobBytes) : (s_act + slot * kActBytes + (uint32_t)src * kBlobBytes);
              const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_ss(d_tmem, umma_desc(a_addr + k4 * 32u, 16, 1024), umma_desc(b_addr + k4 * 32u, 16, 1024), idesc,
                        (kb | k4) ? 1u : 0u);
              if (kb == nkb - 1) umma_commit(b_accready + 8 * slot);   // this slot's accumulator is complete
            }
            umma_commit(b_wempty + 8 * stage);                         // stage is free once both slots' MMAs retire
          }
          __syncwarp();
          stage ^= 1; if (stage == 0) phase ^= 1;
        }
      }
    }
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 8 + 1] = pa; p.prof[blockIdx.x * 8 + 2] = pw; p.prof[blockIdx.x * 8 + 3] = clock64() - tstart; }
  } else {
    // ============================== epilogue groups ==============================
    const int slot = (warp - 2) >> 2;
    const uint32_t q = (uint32_t)warp & 3u;                 // TMEM lane quarter this warp may access
    const uint32_t r = q * 32u + (uint32_t)lane;            // row of the tile = point
    const uint32_t act_base = s_act + slot * kActBytes, aux_base = s_aux + slot * kBlobBytes;
    const uint32_t t_addr = tmem_base + ((q * 32u) << 16) + (uint32_t)slot * 256u;
    const int grp_tid = threadIdx.x - (64 + slot * 128);    // 0..127 inside the epilogue group
    const int bar_id = kBarEpi0 + slot;
    uint32_t par_acc = 0;
    bool store_pending = false;                              // a bulk store issued by grp_tid 0 still reads smem
    long long pe_wait = 0, pe_body = 0, pe_pro = 0;
    for (long long it = 0; it < max_it; ++it) {
      const long long tile = tile_of(slot, it);
      if (tile >= n_tiles) break;
      const long long pt = tile * 128 + r;
      const bool valid = pt < p.P;
      const long long pc = valid ? pt : p.P - 1;            // clamp: padded rows compute finite garbage
      // ---- layer-0 operand: positional encoding of the point (K3 fused) ----
      float dirx = 0.f, diry = 0.f, dirz = 0.f;
      if (p.x_emb == nullptr) {
        const long long ray = pc / p.S;
        const float* rr = p.rays + ray * 6;
        const float zz = p.z[pc];
        const float ox = rr[0], oy = rr[1], oz = rr[2], dx = rr[3], dy = rr[4], dz = rr[5];
        const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
        dirx = dx * inv; diry = dy * inv; dirz = dz * inv;
        if (TRAIN && store_pending) { if (grp_tid == 0) bulk_wait_read0(); named_bar_sync(bar_id, 128); store_pending = false; }
        pe_row_to_smem<10>(aux_base + r * 128u, r, fmaf(dx, zz, ox), fmaf(dy, zz, oy), fmaf(dz, zz, oz), false);
      } else {
        if (TRAIN && store_pending) { if (grp_tid == 0) bulk_wait_read0(); named_bar_sync(bar_id, 128); store_pending = false; }
        emb_row_to_smem(aux_base + r * 128u, r, p.x_emb + pc * p.ld_x, 63, false);
      }
      fence_proxy_async_smem();
      if (TRAIN) {
        named_bar_sync(bar_id, 128);
        if (grp_tid == 0) { bulk_s2g(p.stash + p.st.off_embx + (size_t)tile * kBlobBytes, aux_base, kBlobBytes); bulk_commit(); }
        store_pending = true;
      }
      mbar_arrive(b_aready + 8 * slot);

      float sigma = 0.f;
      float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int s = 0; s < kFwdSteps; ++s) {
        long long t_e0 = clock64();
        mbar_wait(b_accready + 8 * slot, par_acc); par_acc ^= 1;
        tc_fence_after();
        { const long long t1 = clock64(); pe_wait += t1 - t_e0; t_e0 = t1; }
        if (TRAIN && store_pending) { if (grp_tid == 0) bulk_wait_read0(); named_bar_sync(bar_id, 128); store_pending = false; }
        // The chunk loop is deliberately NOT unrolled and is specialised per step kind: one 32-column body is
        // ~200 instructions (3 KB) and stays resident in the instruction cache across chunks, steps and tiles.  (A fully
        // unrolled epilogue streamed ~30 KB of code per step through the I-cache and ran 5x slower: stall_no_inst.)
        const int kind = (s == 7) ? 1 : (s == 9 ? 2 : (s == 8 ? 3 : 0));
        uint32_t* mdst = nullptr;
        if (TRAIN && s != 8)
          mdst = reinterpret_cast<uint32_t*>(p.stash + p.st.off_mask) + (((size_t)tile * 9 + (s < 8 ? s : 8)) * 128 + r) * 8;
        if (kind == 0) epi_chunks<TRAIN, DBG, 0>(p, s, 8, t_addr, act_base, r, pt, valid, mdst, sigma, rgb);
        else if (kind == 1) epi_chunks<TRAIN, DBG, 1>(p, s, 8, t_addr, act_base, r, pt, valid, mdst, sigma, rgb);
        else if (kind == 2) epi_chunks<TRAIN, DBG, 2>(p, s, 4, t_addr, act_base, r, pt, valid, mdst, sigma, rgb);
        else epi_chunks<TRAIN, DBG, 3>(p, s, 8, t_addr, act_base, r, pt, valid, mdst, sigma, rgb);
        if (s == 5) {
          // the step-9 operand needs PE(viewdir) in aux; aux (PE of the point) was last read by MMA step 5, now retired
          if (p.x_emb == nullptr) pe_row_to_smem<4>(aux_base + r * 128u, r, dirx, diry, dirz, false);
          else emb_row_to_smem(aux_base + r * 128u, r, p.x_emb + pc * p.ld_x + 63, 27, false);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        if (TRAIN) {
          named_bar_sync(bar_id, 128);
          if (grp_tid == 0 && !(p.abl & 2)) {
            const size_t off = (s < 8 ? p.st.off_h[s] : (s == 8 ? p.st.off_feat : p.st.off_g));
            const uint32_t nb = (s == 9) ? 2u : 4u;
            bulk_s2g(p.stash + off + (size_t)tile * nb * kBlobBytes, act_base, nb * kBlobBytes);
            if (s == 5) bulk_s2g(p.stash + p.st.off_embd + (size_t)tile * kBlobBytes, aux_base, kBlobBytes);
            bulk_commit();
          }
          store_pending = true;
        }
        pe_body += clock64() - t_e0;
        if (s < 9) {
          mbar_arrive(b_aready + 8 * slot);
        } else if (valid) {
          const float4 bc = make_float4(c_fw.bc[0], c_fw.bc[1], c_fw.bc[2], c_fw.bc[3]);
          reinterpret_cast<float4*>(p.raw)[pt] = make_float4(rgb[0] + bc.x, rgb[1] + bc.y, rgb[2] + bc.z, sigma + bc.w);
        }
      }
    }
    if (TRAIN && store_pending && grp_tid == 0) bulk_wait_all0();
    if (p.prof && grp_tid == 0) { p.prof[blockIdx.x * 8 + 4 + slot * 2] = pe_wait; p.prof[blockIdx.x * 8 + 5 + slot * 2] = pe_body; }
    (void)pe_pro;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// weight packing: fp32 nn.Linear weights -> bf16 blobs that are exact shared-memory images
// ------------------------------------------------------------------------------------------
struct BlobDesc {
  uint32_t dst_off;     // byte offset in the packed buffer
  uint32_t src_off;     // float offset of W in the flat params
  int ld;               // row stride (in-features) of W
  int transposed;       // 0: blob[n][k] = W[n0+n][k0+k]   1: blob[n][k] = W[k0+k][n0+n]
  int n0, k0;
  int n_rows;           // blob rows (128 or 256)
  int n_lim, k_lim;     // valid extents: n0+n < n_lim, k0+k < k_lim (else 0)
};
constexpr int kMaxBlobs = 80;
struct PackParams { BlobDesc b[kMaxBlobs]; int n; };

__global__ void __launch_bounds__(256)
pack_kernel(const PackParams pp, const float* __restrict__ prm, uint8_t* __restrict__ out, NbParamLayout L, uint32_t small_off) {
  if ((int)blockIdx.y == pp.n) {       // last row of blocks: gather the small fp32 parameters (TcSmall)
    TcSmall* sm = reinterpret_cast<TcSmall*>(out + small_off);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 10 * 256; i += gridDim.x * blockDim.x) {
      const int s = i >> 8, c = i & 255;
      float v = 0.f;
      if (s < 8) v = prm[L.b[s] + c];
      else if (s == 8) v = prm[L.bf + c];
      else if (c < 128) v = prm[L.bd + c];
      sm->bias[s][c] = v;
      if (s == 0) sm->ws[c] = prm[L.ws + c];
      if (i < 384) sm->wc[i] = prm[L.wc + i];
      if (i < 3) sm->bc[i] = prm[L.bc + i];
      if (i == 3) sm->bc[3] = prm[L.bs];
    }
    return;
  }
  const BlobDesc d = pp.b[blockIdx.y];
  const int total = d.n_rows * 8;    // 16-byte chunks
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i >> 3, c = i & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c * 8 + j;
      const int gn = d.n0 + n, gk = d.k0 + k;
      float x = 0.f;
      if (gn < d.n_lim && gk < d.k_lim) x = d.transposed ? prm[d.src_off + (size_t)gk * d.ld + gn] : prm[d.src_off + (size_t)gn * d.ld + gk];
      v[j] = x;
    }
    uint4 w = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(out + d.dst_off + sw128_chunk((uint32_t)n, (uint32_t)c)) = w;
  }
}

void add_blob(PackParams& pp, uint32_t& off, size_t src, int ld, int tr, int n0, int k0, int rows, int n_lim, int k_lim) {
  BlobDesc& b = pp.b[pp.n++];
  b.dst_off = off; b.src_off = (uint32_t)src; b.ld = ld; b.transposed = tr; b.n0 = n0; b.k0 = k0; b.n_rows = rows;
  b.n_lim = n_lim; b.k_lim = k_lim;
  off += (uint32_t)rows * 128u;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
bool nb_tc_supported(const nb_mlp_desc& d) {
  return d.D == 8 && d.W == 256 && d.skip == 4 && d.in_x == 63 && d.in_d == 27 && d.L_x == 10 && d.L_d == 4;
}

size_t nb_tc_fwd_packed_bytes() { return fwd_w_off(kFwdSteps); }
size_t nb_tc_small_offset() { return nb_tc_fwd_packed_bytes() + nb_tc_bwd_packed_bytes(); }
size_t nb_tc_packed_bytes(const nb_mlp_desc&) { return nb_tc_small_offset() + sizeof(TcSmall); }

TcStash nb_tc_stash_layout(long long P) {
  TcStash s;
  const size_t T = (size_t)((P + 127) / 128);
  size_t off = 0;
  s.off_embx = off; off += T * kBlobBytes;
  s.off_embd = off; off += T * kBlobBytes;
  for (int i = 0; i < 8; ++i) { s.off_h[i] = off; off += T * 4 * kBlobBytes; }
  s.off_feat = off; off += T * 4 * kBlobBytes;
  s.off_g = off; off += T * 2 * kBlobBytes;
  s.off_mask = off; off += T * kMaskTileBytes;
  s.total = off;
  s.tiles = (long long)T;
  return s;
}
size_t nb_tc_act_bytes(const nb_mlp_desc&, long long P) { return nb_tc_stash_layout(P).total; }

size_t nb_tc_ws_bytes(const nb_mlp_desc& d, long long P, int backward) { return backward ? nb_tc_bwd_ws_bytes(d, P) : 256; }

int nb_tc_pack(nb_handle_t h, const nb_mlp_desc* d, const float* params, void* packed, cudaStream_t st) {
  const NbParamLayout L = nb_param_layout(*d);
  PackParams pp;
  pp.n = 0;
  uint32_t off = 0;
  // ---- forward blobs, in consumption order
  add_blob(pp, off, L.w[0], 63, 0, 0, 0, 256, 256, 63);                                  // step 0
  for (int l = 1; l <= 4; ++l) for (int kb = 0; kb < 4; ++kb) add_blob(pp, off, L.w[l], 256, 0, 0, 64 * kb, 256, 256, 256);
  add_blob(pp, off, L.w[5], 319, 0, 0, 0, 256, 256, 63);                                 // step 5: PE columns
  for (int kb = 0; kb < 4; ++kb) add_blob(pp, off, L.w[5] + 63, 319, 0, 0, 64 * kb, 256, 256, 256);
  for (int l = 6; l <= 7; ++l) for (int kb = 0; kb < 4; ++kb) add_blob(pp, off, L.w[l], 256, 0, 0, 64 * kb, 256, 256, 256);
  for (int kb = 0; kb < 4; ++kb) add_blob(pp, off, L.wf, 256, 0, 0, 64 * kb, 256, 256, 256);
  for (int kb = 0; kb < 4; ++kb) add_blob(pp, off, L.wd, 283, 0, 0, 64 * kb, 128, 128, 256);   // step 9: feature columns
  add_blob(pp, off, L.wd + 256, 283, 0, 0, 0, 128, 128, 27);                                   //         view-dir PE columns
  if (off != nb_tc_fwd_packed_bytes()) { NB_SET_ERR(h, "nb_tc_pack: internal layout mismatch"); return NB_ERR_INVALID; }
  // ---- backward (dgrad) blobs: B = W^T
  nb_tc_bwd_add_blobs(L, [&](size_t src, int ld, int tr, int n0, int k0, int rows, int n_lim, int k_lim) {
    add_blob(pp, off, src, ld, tr, n0, k0, rows, n_lim, k_lim);
  });
  if (pp.n > kMaxBlobs) { NB_SET_ERR(h, "nb_tc_pack: too many blobs"); return NB_ERR_INVALID; }
  dim3 grid(4, pp.n + 1);
  pack_kernel<<<grid, 256, 0, st>>>(pp, params, (uint8_t*)packed, L, (uint32_t)nb_tc_small_offset());
  NB_LAUNCHED(h);
  return NB_OK;
}

static int launch_fwd(nb_handle_t h, FwdParams& fp, bool train, cudaStream_t st) {
  static bool attr_done[3] = {false, false, false};
  const bool dbg = fp.dbg != nullptr;
  auto kern = dbg ? mlp_fwd_chain_kernel<false, true> : (train ? mlp_fwd_chain_kernel<true, false> : mlp_fwd_chain_kernel<false, false>);
  const int ki = dbg ? 2 : (train ? 1 : 0);
  if (!attr_done[ki]) {
    NB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_done[ki] = true;
  }
  NB_CUDA(h, cudaMemcpyToSymbolAsync(c_fw, fp.wpk + nb_tc_small_offset(), sizeof(TcSmall), 0, cudaMemcpyDeviceToDevice, st));
  const long long n_tiles = (fp.P + 127) / 128;
  long long grid = (n_tiles + 1) / 2;
  if (grid > h->sm_count) grid = h->sm_count;
  static long long* prof_dev = nullptr;
  const bool prof = getenv("NB_TC_PROF") != nullptr;
  if (prof) {
    if (!prof_dev) cudaMalloc(&prof_dev, 256 * 8 * sizeof(long long));
    cudaMemsetAsync(prof_dev, 0, 256 * 8 * sizeof(long long), st);
    fp.prof = prof_dev;
  }
  kern<<<(int)grid, kThreads, kSmemBytes, st>>>(fp);
  NB_LAUNCHED(h);
  if (prof) {   // diagnostic only: synchronous read-back of the per-CTA cycle counters
    static long long host[256 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost);
    double a[8] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 8; ++k) a[k] += (double)host[b * 8 + k] / grid;
    fprintf(stderr, "nb_tc prof (avg cycles/CTA, P=%lld train=%d): epi1_tmem_ld=%.0f mma_wait_aready=%.0f mma_wait_wfull=%.0f "
            "mma_total=%.0f epi0_wait=%.0f epi0_body=%.0f epi1_wait=%.0f epi1_body=%.0f\n", fp.P, (int)train, a[0], a[1], a[2], a[3],
            a[4], a[5], a[6], a[7]);
  }
  return NB_OK;
}

int nb_tc_forward(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P, const float* x,
                  int64_t ld_x, const float* rays, const float* z, int32_t S, float* raw_out, void* act_save, void* ws,
                  size_t ws_bytes, cudaStream_t st) {
  (void)ws; (void)ws_bytes;
  FwdParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.rays = rays; fp.z = z; fp.x_emb = x; fp.ld_x = ld_x; fp.P = P; fp.S = S > 0 ? S : 1;
  fp.wpk = (const uint8_t*)packed; fp.prm = params; fp.L = nb_param_layout(*d); fp.raw = raw_out;
  fp.stash = (uint8_t*)act_save; fp.st = nb_tc_stash_layout(P);
  fp.dbg = nullptr; fp.dbg_step = -1;
  { const char* e = getenv("NB_TC_ABLATE"); fp.abl = e ? atoi(e) : 0; }
  { const char* e = getenv("NB_TC_SHARE_W"); fp.share_w = e ? atoi(e) : 0; }
  return launch_fwd(h, fp, act_save != nullptr, st);
}

// diagnostic: run the forward chain and dump the raw fp32 accumulators (before bias/activation) of `step`
extern "C" int nb_mlp_tc_probe(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t N, int32_t S,
                               const float* rays, const float* z, int32_t step, float* acc_out, float* raw_out, void* stream) {
  NB_ENTER(h);
  NB_REQUIRE(h, d && nb_tc_supported(*d) && params && packed && rays && z && acc_out && raw_out && step >= 0 && step < kFwdSteps,
             "nb_mlp_tc_probe: bad arguments");
  FwdParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.rays = rays; fp.z = z; fp.P = N * S; fp.S = S; fp.wpk = (const uint8_t*)packed; fp.prm = params;
  fp.L = nb_param_layout(*d); fp.raw = raw_out; fp.stash = nullptr; fp.st = nb_tc_stash_layout(fp.P);
  fp.dbg = acc_out; fp.dbg_step = step;
  return launch_fwd(h, fp, false, (cudaStream_t)stream);
}
