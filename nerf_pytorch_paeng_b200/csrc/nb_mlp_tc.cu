// K3+K4, NB_BF16 precision: the whole 8x256 skip MLP of one 128-point tile as a chain of tcgen05
// GEMMs whose activations never leave the SM.
//
// Replaces nerf_process.py:69-84 (points + positional encoding) and model/NeRF.py:33-52.
//
// One persistent CTA per SM (launched as clusters of two that share the weight stream), 576 threads, two 128-point tiles
// ("slots") in flight:
//   warp 0      weight producer: streams the pre-swizzled bf16 weight blobs (32 KB = 256 out-features x 64 in-features, stored as two
//               SWIZZLE_64B images of 32 in-features each, wblob_chunk() in nb_tc_common.cuh) from L2 into a 2-stage shared-memory
//               ring with cp.async.bulk (TMA unit) + mbarrier; in the default cluster mode each CTA fetches half of every stage
//               and multicasts it into both CTAs' rings.  (Ring stages of ONE sub-blob, 4 x 16 KB, were measured: the single
//               issuing thread then pays its wait / fence / commit sequence per two MMAs instead of four and becomes the bound --
//               inference 0.80 -> 1.09 ms.)
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=256|128, K=16) with A = the slot's activation tile in shared
//               memory (K-major, 128B swizzle), B = the weight stage, D = the slot's 256 fp32 TMEM columns; tcgen05.commit
//               releases the weight stage / signals the epilogue.
//   warps 2-9   epilogue of slot 0, warps 10-17 epilogue of slot 1: two warps per TMEM lane quarter, one per half of the
//               accumulator columns (thread = one point): build the positional encoding of the point straight into the
//               A-operand tile (K3 fused into the first GEMM's operand), then per layer tcgen05.ld the accumulators, add bias,
//               ReLU, round to bf16 and write the next layer's A tile in place; sigma and rgb heads are CUDA-core dot products
//               on the fp32 accumulators; raw[N,S,4] is the only HBM write in inference.
// While the epilogue of one slot runs, the tensor core works on the other slot (ping-pong).
// Layer chain ("steps") per tile, K-blocks of 64:  0: PE63->256 | 1-4: 256->256 | 5: [PE63,256]->256 |
// 6,7: 256->256 (+sigma head after 7) | 8: view layer [h7 (256) through W', PEd27]->128 (+rgb head).
// The reference's feature layer has NO activation (NeRF.py:44: feat = linear_feat(h)) and feeds only the view layer
// (NeRF.py:47-49: relu(linear_d(cat[feat, d]))), so it is folded into the weights once per weight update (nb_tc_pack):
//   W' = Wd[:, :256] . Wf  (128x256),  b' = Wd[:, :256] . b_feat + b_d,   g = relu(W' h7 + Wd[:, 256:] PE(d) + b')
// which removes one 256x256 GEMM step per tile (11% of the network's MACs) from inference AND training; the backward
// recovers the gradients of Wf, b_feat and Wd[:, :256] from G = dg^T h7 (nb_mlp_tc_bwd.cu).  The fp32 parity path keeps the
// reference's explicit two-layer sequence.
//
// In training the bf16 layer inputs are additionally written to HBM STRAIGHT FROM THE EPILOGUE'S REGISTERS (the same packed words
// that go to the shared-memory A tile), in the chunk-major blob layout of stash_off() (nb_tc_common.cuh): fully coalesced 16-byte
// stores, no shared-memory read-back, no extra barrier.  The backward kernels (nb_mlp_tc_bwd.cu) consume those blobs directly as
// MN-major UMMA operands.
#include <stdlib.h>
#include "nb_mlp.h"
#include "nb_tc_common.cuh"
#include "nb_mlp_tc.h"

using namespace tc;

namespace {

// ------------------------------------------------------------------------------------------
// forward chain tables (D=8, W=256, skip=4, in_x=63, in_d=27)
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fwd_nkb(int s) { return (s == 0) ? 1 : ((s == 5 || s == 8) ? 5 : 4); }
__host__ __device__ constexpr int fwd_n(int s) { return s == 8 ? 128 : 256; }
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
  if (s == 8) return kb == 4 ? -1 : kb;
  return kb;
}

// shared-memory carve-up (bytes, after 1024-byte alignment)
constexpr uint32_t kActBytes = 4 * kBlobBytes;               // 64 KB per slot
constexpr uint32_t kOffAct = 0;                              // [2][4][16 KB]
constexpr uint32_t kOffAux = 2 * kActBytes;                  // [2][16 KB]
constexpr uint32_t kOffW = kOffAux + 2 * kBlobBytes;         // [2][32 KB]
constexpr uint32_t kOffBar = kOffW + 2 * 32768;              // barriers (256 B)
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;        // + alignment slack

constexpr int kThreads = 576;   // warp 0 producer, warp 1 MMA, 8 epilogue warps per slot (2 per TMEM lane quarter: column halves)
constexpr int kEpiThreads = 256;
constexpr int kBarEpi0 = 1;

__constant__ TcSmall c_fw[kConstBanks];   // small fp32 parameters of the network being run (see nb_mlp_tc.h), one bank per stream in flight (nb_cbank.h)
NbConstBankTable g_fw_banks;   // named barrier ids of the two epilogue groups

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
  int abl;                // ablation bits for profiling experiments (NB_TC_ABLATE env): 1 no masks, 2 no stash stores
  int bank;               // which copy of c_fw holds this network's constants (nb_cbank.h)
};

// positional-encoding features of one 3-vector, written as bf16 into a swizzled 128-byte row.
// Arguments are reduced in "turns": sin(2^k x) = sin(2 pi frac(2^k x / 2 pi)), exact power-of-two scaling,
// so the fast sin.approx/cos.approx see |arg| <= pi (abs error ~1e-6, far below a bf16 ulp).
template <int L>
__device__ __forceinline__ void pe_row_to_smem(uint32_t row_addr, uint32_t r, float x, float y, float z, bool one_pad, int c_lo, int c_hi,
                                               uint8_t* gblob = nullptr) {
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
    if (c >= c_lo && c < c_hi) {                                                                           // this thread's column half
      st_shared_v4(row_addr + (((uint32_t)c ^ (r & 7u)) << 4), w0, w1, w2, w3);
      if (gblob != nullptr) st_global_na_v4(gblob + stash_off(r, (uint32_t)c), w0, w1, w2, w3);          // training stash (wgrad's X operand)
    }
  }
}

// copy 64 bf16 features (cols [c0, c0+ncol) of a materialised fp32 embedding row) into a swizzled row
__device__ __forceinline__ void emb_row_to_smem(uint32_t row_addr, uint32_t r, const float* src, int ncol, bool one_pad, int c_lo, int c_hi,
                                                uint8_t* gblob = nullptr) {
#pragma unroll 1
  for (int c = c_lo; c < c_hi; ++c) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int f = c * 8 + j;
      v[j] = f < ncol ? src[f] : ((one_pad && f == ncol) ? 1.0f : 0.0f);
    }
    const uint32_t w0 = pack_bf16(v[0], v[1]), w1 = pack_bf16(v[2], v[3]), w2 = pack_bf16(v[4], v[5]), w3 = pack_bf16(v[6], v[7]);
    st_shared_v4(row_addr + (((uint32_t)c ^ (r & 7u)) << 4), w0, w1, w2, w3);
    if (gblob != nullptr) st_global_na_v4(gblob + stash_off(r, (uint32_t)c), w0, w1, w2, w3);
  }
}

// One layer's epilogue over `nchunks` groups of 32 accumulator columns (thread = one point / TMEM lane).
// KIND 0: bias+ReLU -> bf16 A tile | 1: same + sigma head (step 7) | 2: view layer + rgb head (step 8; stored only in
// training, for the stash).
template <bool TRAIN, bool DBG, int KIND>
__device__ __forceinline__ void epi_chunks(const FwdParams& p, int s, int c_begin, int c_end, uint32_t t_addr, uint32_t act_base, uint32_t r,
                                           long long pt, bool valid, uint32_t* mdst, uint8_t* gdst, float& sigma, float (&rgb)[3]) {
  const TcSmall& C = c_fw[p.bank];
  const float* bias = C.bias[s];
  uint4 mw = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  for (int c32 = c_begin; c32 < c_end; ++c32) {
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
    for (int j = 0; j < 32; j += 2) add_f32x2(v[j], v[j + 1], bc[j], bc[j + 1]);      // packed FADD2: half the issue slots
    if (TRAIN) {
      // ReLU mask of this layer's output for the backward chain, one funnel shift per element: bit (31-j) of the
      // word = SIGN bit of column c32*32+j, i.e. set = inactive.  (An exact +0.0 pre-activation counts as active,
      // where torch's relu' is 0: a measure-zero difference.)  The (up to) four words of this thread's column half are
      // kept in registers and leave as ONE 16-byte store after the loop (a warp then writes 512 contiguous bytes).
      uint32_t m = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = __funnelshift_l(__float_as_uint(v[j]), m, 1);
      const int k = c32 - c_begin;
      if (k == 0) mw.x = m; else if (k == 1) mw.y = m; else if (k == 2) mw.z = m; else mw.w = m;
    }
    if (KIND == 1) {            // sigma head on the fp32 post-ReLU trunk output (NeRF.py:43)
      const float* w = C.ws + c32 * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) sigma = fmaf(fmaxf(v[j], 0.f), w[j], sigma);
    }
    if (KIND == 2) {            // rgb head on the fp32 post-ReLU view features (NeRF.py:50)
      const float* w = C.wc + c32 * 32;
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
        const uint32_t w0 = pack_bf16_relu(v[j * 8 + 0], v[j * 8 + 1]), w1 = pack_bf16_relu(v[j * 8 + 2], v[j * 8 + 3]);
        const uint32_t w2 = pack_bf16_relu(v[j * 8 + 4], v[j * 8 + 5]), w3 = pack_bf16_relu(v[j * 8 + 6], v[j * 8 + 7]);
        if (KIND != 2 && !(p.abl & 4)) st_shared_v4(row_addr + ((c ^ (r & 7u)) << 4), w0, w1, w2, w3);     // (the view layer feeds no further GEMM)
        // training: the same 16 bytes go straight from the registers to the stash blob in the chunk-major layout (stash_off): the
        // 32 lanes of the warp (consecutive points) write 512 contiguous bytes.  No shared-memory read-back, no barrier, and the
        // TMA unit / shared-memory ports stay with the weight stream and the MMA operands.
        if (TRAIN && gdst != nullptr)
          st_global_na_v4(gdst + (size_t)(c32 >> 1) * kBlobBytes + stash_off(r, c), w0, w1, w2, w3);
      }
    }
  }
  if (TRAIN && mdst != nullptr)
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(mdst), "r"(mw.x), "r"(mw.y), "r"(mw.z), "r"(mw.w) : "memory");
}

// MC = true: clusters of two independent CTAs (each issues its own cta_group::1 MMAs) that SHARE the weight stream: each CTA fetches
// half of every 32 KB weight stage and multicasts it into both CTAs' rings, halving the L2->SM weight traffic; a stage is refilled
// once BOTH CTAs have retired its MMAs (commit multicast, count 2).  (A cta_group::2 variant -- M = 256 MMAs over the pair, each CTA
// holding half of B -- was built in round 1, measured slower than this ring and retired.)
template <bool TRAIN, bool DBG, bool MC>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fwd_chain_kernel(const FwdParams p) {
  constexpr bool PAIR = MC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_act = sbase + kOffAct, s_aux = sbase + kOffAux, s_w = sbase + kOffW, s_bar = sbase + kOffBar;
  // barriers (8 bytes each): w_full[4] w_empty[4] p_full[4] a_ready[2] acc_ready[2] ; tmem ptr
  const uint32_t b_wfull = s_bar, b_wempty = s_bar + 32, b_pfull = s_bar + 64, b_aready = s_bar + 96, b_accready = s_bar + 112,
                 s_tmem = s_bar + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t NSTAGE = 2u;               // weight ring: 2 stages of one K block (two 32-k sub-blobs, 32 KB at N = 256)
  constexpr uint32_t STAGE_BYTES = 32768u;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;

  const long long n_tiles = (p.P + 127) / 128;
  // work units: MC: "pair tiles" q (tiles 2q, 2q+1 for ranks 0,1); else single tiles
  const long long n_units = PAIR ? (n_tiles + 1) / 2 : n_tiles;
  const long long ncl = PAIR ? gridDim.x / 2 : gridDim.x, cid = PAIR ? blockIdx.x / 2 : blockIdx.x;
  auto unit_of = [&](int slot, long long it) { return (it * ncl + cid) * 2 + slot; };
  const long long max_it = (n_units + 2 * ncl - 1) / (2 * ncl);

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < 4; ++i) {
      mbar_init(b_wfull + 8 * i, 1);
      mbar_init(b_wempty + 8 * i, MC ? 2 : 1);
      mbar_init(b_pfull + 8 * i, 1);
    }
    for (uint32_t i = 0; i < 2; ++i) {
      mbar_init(b_aready + 8 * i, kEpiThreads);
      mbar_init(b_accready + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

  if (warp == 0) {
    // ============================== weight producer (every CTA) ==============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long it = 0; it < max_it; ++it) {
#pragma unroll 1
        for (int s = 0; s < kFwdSteps; ++s) {
          const uint32_t bytes = fwd_blob_bytes(s);
          const uint8_t* src = p.wpk + fwd_w_off(s);
          for (int slot = 0; slot < 2; ++slot) {
            if (unit_of(slot, it) >= n_units) continue;
            for (int kb = 0; kb < fwd_nkb(s); ++kb) {
              mbar_wait(b_wempty + 8 * stage, phase ^ 1);
              if (p.abl & 8) { mbar_arrive(b_wfull + 8 * stage); if (++stage == NSTAGE) { stage = 0; phase ^= 1; } continue; }
              mbar_expect_tx(b_wfull + 8 * stage, bytes);
              const uint32_t q4 = bytes >> 2;
              if (MC) {        // my half of the stage, delivered to both CTAs (the other half arrives from the peer)
                const uint32_t hb = bytes >> 1, q2 = hb >> 1;
#pragma unroll
                for (int i = 0; i < 2; ++i)
                  bulk_g2s_mcast(s_w + stage * STAGE_BYTES + rank * hb + i * q2, src + (size_t)kb * bytes + rank * hb + i * q2, q2,
                                 b_wfull + 8 * stage, (uint16_t)3);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  bulk_g2s(s_w + stage * STAGE_BYTES + i * q4, src + (size_t)kb * fwd_blob_bytes(s) + i * q4, q4, b_wfull + 8 * stage);
              }
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, par_a[2] = {0, 0};
    {
      // ============================== MMA issuer ==============================
      long long pa = 0, pw = 0, pp = 0;
      const long long tstart = clock64();
      for (long long it = 0; it < max_it; ++it) {
#pragma unroll 1
        for (int s = 0; s < kFwdSteps; ++s) {
          const uint32_t idesc = umma_idesc(128, fwd_n(s), 0, 0);
          const int nkb = fwd_nkb(s);
          for (int slot = 0; slot < 2; ++slot) {
            if (unit_of(slot, it) >= n_units) continue;
            { const long long t0 = clock64();
              mbar_wait(b_aready + 8 * slot, par_a[slot]);
              pa += clock64() - t0; }
            par_a[slot] ^= 1;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
            for (int kb = 0; kb < nkb; ++kb) {
              { const long long t0 = clock64(); mbar_wait(b_wfull + 8 * stage, phase); const long long t1 = clock64(); pw += t1 - t0;
                (void)t1; }
              tc_fence_after();
              if (lane == 0) {
                const int src = fwd_a_src(s, kb);
                const uint32_t a_addr = (src < 0) ? (s_aux + slot * kBlobBytes) : (s_act + slot * kActBytes + (uint32_t)src * kBlobBytes);
                const uint32_t b_addr = s_w + stage * STAGE_BYTES;
                const uint32_t sub = fwd_blob_bytes(s) >> 1;      // the K block = two 32-k sub-blobs (SWIZZLE_64B images) laid end to end
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const uint64_t ad = umma_desc(a_addr + k4 * 32u, 16, 1024);
                  const uint64_t bd = umma_desc_sw64(b_addr + (uint32_t)(k4 >> 1) * sub + (uint32_t)(k4 & 1) * 32u, 16, 512);
                  umma_ss(d_tmem, ad, bd, idesc, (kb | k4) ? 1u : 0u);
                }
                if (MC) umma_commit_mcast(b_wempty + 8 * stage);        // both CTAs' producers learn that I am done with it
                else umma_commit(b_wempty + 8 * stage);                 // stage is free once these MMAs retire
                if (kb == nkb - 1) umma_commit(b_accready + 8 * slot);  // accumulator complete
              }
              __syncwarp();
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      if (p.prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 8;
        o[0] = pa; o[1] = pw; o[2] = pp; o[3] = clock64() - tstart;
      }
    }
  } else {
    // ============================== epilogue groups ==============================
    const int slot = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;                 // which half of the accumulator columns this warp drains
    const uint32_t q = (uint32_t)warp & 3u;                 // TMEM lane quarter this warp may access
    const uint32_t r = q * 32u + (uint32_t)lane;            // row of the tile = point (two threads per row: column halves)
    const uint32_t act_base = s_act + slot * kActBytes, aux_base = s_aux + slot * kBlobBytes;
    const uint32_t t_addr = tmem_base + ((q * 32u) << 16) + (uint32_t)slot * 256u;
    const int grp_tid = threadIdx.x - (64 + slot * kEpiThreads);    // 0..255 inside the epilogue group
    const int bar_id = kBarEpi0 + slot;
    uint32_t par_acc = 0;
    long long pe_wait = 0, pe_body = 0, pe_pro = 0;

    for (long long it = 0; it < max_it; ++it) {
      const long long unit = unit_of(slot, it);
      if (unit >= n_units) break;
      const long long tile = PAIR ? unit * 2 + rank : unit;
      const bool tile_ok = tile < n_tiles;                  // MC: the last pair may have a ghost second tile
      const long long pt = tile * 128 + r;
      const bool valid = pt < p.P;
      const long long pc = valid ? pt : p.P - 1;            // clamp: padded rows compute finite garbage
      const long long tile_st = (p.abl & 64) ? (tile & 63) : tile;   // experiment: stash writes stay L2-resident
      // ---- layer-0 operand: positional encoding of the point (K3 fused) ----
      float dirx = 0.f, diry = 0.f, dirz = 0.f;
      const long long t_p0 = clock64();
      uint8_t* g_embx = (TRAIN && tile_ok && !(p.abl & 16)) ? p.stash + p.st.off_embx + (size_t)tile_st * kBlobBytes : nullptr;
      uint8_t* g_embd = (TRAIN && tile_ok && !(p.abl & 16)) ? p.stash + p.st.off_embd + (size_t)tile_st * kBlobBytes : nullptr;
      if (p.x_emb == nullptr) {
        const long long ray = pc / p.S;
        const float* rr = p.rays + ray * 6;
        const float zz = p.z[pc];
        const float ox = rr[0], oy = rr[1], oz = rr[2], dx = rr[3], dy = rr[4], dz = rr[5];
        const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
        dirx = dx * inv; diry = dy * inv; dirz = dz * inv;
        pe_row_to_smem<10>(aux_base + r * 128u, r, fmaf(dx, zz, ox), fmaf(dy, zz, oy), fmaf(dz, zz, oz), TRAIN, half * 4, half * 4 + 4, g_embx);
      } else {
        emb_row_to_smem(aux_base + r * 128u, r, p.x_emb + pc * p.ld_x, 63, TRAIN, half * 4, half * 4 + 4, g_embx);
      }
      fence_proxy_async_smem();
      mbar_arrive(b_aready + 8 * slot);
      pe_pro += clock64() - t_p0;

      float sigma = 0.f;
      float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int s = 0; s < kFwdSteps; ++s) {
        long long t_e0 = clock64();
        mbar_wait(b_accready + 8 * slot, par_acc); par_acc ^= 1;
        tc_fence_after();
        { const long long t1 = clock64(); pe_wait += t1 - t_e0; t_e0 = t1; }
        // The chunk loop is deliberately NOT unrolled and is specialised per step kind: one 32-column body is
        // ~200 instructions (3 KB) and stays resident in the instruction cache across chunks, steps and tiles.  (A fully
        // unrolled epilogue streamed ~30 KB of code per step through the I-cache and ran 5x slower: stall_no_inst.)
        const int kind = (s == 7) ? 1 : (s == 8 ? 2 : 0);
        uint32_t* mdst = nullptr;
        if (TRAIN)
          mdst = reinterpret_cast<uint32_t*>(p.stash + p.st.off_mask) +
                 ((((size_t)(tile_ok ? tile_st : 0) * 9 + (s < 8 ? s : 8)) * 2 + half) * 128 + r) * 4;     // [tile][9][half][row][4 words]
        const bool wmask = TRAIN && tile_ok;
        const int c0 = half * 4, c1 = half * 4 + 4;            // this warp's 128 of the 256 columns (64 of 128 in the view layer)
        uint8_t* gdst = nullptr;
        if (TRAIN && tile_ok && !(p.abl & 16)) {
          const size_t off = (s < 8 ? p.st.off_h[s] : p.st.off_g);
          gdst = p.stash + off + (size_t)tile_st * (s == 8 ? 2u : 4u) * kBlobBytes;
        }
        if (kind == 0) epi_chunks<TRAIN, DBG, 0>(p, s, c0, c1, t_addr, act_base, r, pt, valid, wmask ? mdst : nullptr, gdst, sigma, rgb);
        else if (kind == 1) epi_chunks<TRAIN, DBG, 1>(p, s, c0, c1, t_addr, act_base, r, pt, valid, wmask ? mdst : nullptr, gdst, sigma, rgb);
        else epi_chunks<TRAIN, DBG, 2>(p, s, half * 2, half * 2 + 2, t_addr, act_base, r, pt, valid, wmask ? mdst : nullptr, gdst, sigma, rgb);
        // heads: the two column halves of a row exchange their partial rgb / sigma sums through a scratch row in the
        // (now idle) fourth K-block of the activation tile
        const uint32_t scratch = act_base + 3u * kBlobBytes + r * 16u;
        if (s == kFwdSteps - 1 && half == 1)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(scratch), "f"(rgb[0]), "f"(rgb[1]), "f"(rgb[2]), "f"(sigma) : "memory");
        if (s == 5) {
          // the view-layer operand needs PE(viewdir) in aux; aux (PE of the point) was last read by MMA step 5, now retired
          if (p.x_emb == nullptr) pe_row_to_smem<4>(aux_base + r * 128u, r, dirx, diry, dirz, TRAIN, half * 4, half * 4 + 4, g_embd);
          else emb_row_to_smem(aux_base + r * 128u, r, p.x_emb + pc * p.ld_x + 63, 27, TRAIN, half * 4, half * 4 + 4, g_embd);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        pe_body += clock64() - t_e0;
        if (s < kFwdSteps - 1) {      // release the MMA warp first: it needs only every thread's own (fenced) stores, not the group barrier
          mbar_arrive(b_aready + 8 * slot);
        }
        if (s == kFwdSteps - 1) {
          named_bar_sync(bar_id, kEpiThreads);                  // scratch rows of the other column half are visible
          if (half == 0 && valid) {
            float4 o;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(scratch));
            reinterpret_cast<float4*>(p.raw)[pt] = make_float4(rgb[0] + o.x + c_fw[p.bank].bc[0], rgb[1] + o.y + c_fw[p.bank].bc[1],
                                                               rgb[2] + o.z + c_fw[p.bank].bc[2], sigma + o.w + c_fw[p.bank].bc[3]);
          }
        }
      }
    }
    if (p.prof && grp_tid == 0 && slot == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 8;
      o[4] = pe_wait; o[5] = pe_body; o[6] = pe_pro;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // nobody exits while the partner may still signal its barriers / write its smem
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// weight packing: fp32 nn.Linear weights -> bf16 blobs that are exact shared-memory images
// ------------------------------------------------------------------------------------------
struct BlobDesc {
  uint32_t dst_off;     // byte offset in the packed buffer
  uint32_t src_off;     // float offset of W in the flat params (folded != 0: in the folded matrix W' of the packed buffer)
  int folded;
  int ld;               // row stride (in-features) of W
  int transposed;       // 0: blob[n][k] = W[n0+n][k0+k]   1: blob[n][k] = W[k0+k][n0+n]
  int n0, k0;
  int n_rows;           // blob rows (128 or 256)
  int n_lim, k_lim;     // valid extents: n0+n < n_lim, k0+k < k_lim (else 0)
};
constexpr int kMaxBlobs = 72;
struct PackParams { BlobDesc b[kMaxBlobs]; int n; };

// W'[n][k] = sum_j Wd[n][j] * Wf[j][k] (n < 128, k < 256; Wd row stride 283), b'[n] = sum_j Wd[n][j] * bf[j] + bd[n]: the
// activation-free feature layer folded into the view layer, in fp32, once per weight update (8.4 MFLOP).
__global__ void __launch_bounds__(256)
fold_feat_kernel(const float* __restrict__ prm, NbParamLayout L, float* __restrict__ fold) {
  const int n = blockIdx.x;            // 128 blocks: one output row each
  const int k = threadIdx.x;           // 256 threads: one output column each
  __shared__ float wd[256];
  wd[k] = prm[L.wd + (size_t)n * 283 + k];
  __syncthreads();
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains: the loop is latency-bound otherwise
#pragma unroll 16
  for (int j = 0; j < 256; j += 4) {
    a0 = fmaf(wd[j], prm[L.wf + (size_t)j * 256 + k], a0);
    a1 = fmaf(wd[j + 1], prm[L.wf + (size_t)(j + 1) * 256 + k], a1);
    a2 = fmaf(wd[j + 2], prm[L.wf + (size_t)(j + 2) * 256 + k], a2);
    a3 = fmaf(wd[j + 3], prm[L.wf + (size_t)(j + 3) * 256 + k], a3);
  }
  fold[(size_t)n * 256 + k] = (a0 + a1) + (a2 + a3);
  if (k < 32) {                                       // b'[n]: one warp, shuffle-reduced
    float b = 0.f;
    for (int j = k; j < 256; j += 32) b = fmaf(wd[j], prm[L.bf + j], b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (k == 0) fold[128 * 256 + n] = b + prm[L.bd + n];
  }
}

__global__ void __launch_bounds__(256)
pack_kernel(const PackParams pp, const float* __restrict__ prm, uint8_t* __restrict__ out, NbParamLayout L, uint32_t small_off,
            const float* __restrict__ fold) {
  if ((int)blockIdx.y == pp.n) {       // last row of blocks: gather the small fp32 parameters (TcSmall)
    TcSmall* sm = reinterpret_cast<TcSmall*>(out + small_off);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kFwdSteps * 256; i += gridDim.x * blockDim.x) {
      const int s = i >> 8, c = i & 255;
      float v = 0.f;
      if (s < 8) v = prm[L.b[s] + c];
      else if (c < 128) v = fold[128 * 256 + c];       // b' of the folded view layer
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
      const float* src = d.folded ? fold : prm;
      if (gn < d.n_lim && gk < d.k_lim) x = d.transposed ? src[d.src_off + (size_t)gk * d.ld + gn] : src[d.src_off + (size_t)gn * d.ld + gk];
      v[j] = x;
    }
    uint4 w = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(out + d.dst_off + wblob_chunk((uint32_t)d.n_rows, (uint32_t)n, (uint32_t)c)) = w;
  }
}

void add_blob(PackParams& pp, uint32_t& off, size_t src, int ld, int tr, int n0, int k0, int rows, int n_lim, int k_lim, int folded = 0) {
  BlobDesc& b = pp.b[pp.n++];
  b.dst_off = off; b.src_off = (uint32_t)src; b.folded = folded; b.ld = ld; b.transposed = tr; b.n0 = n0; b.k0 = k0; b.n_rows = rows;
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
size_t nb_tc_fold_offset() { return (nb_tc_small_offset() + sizeof(TcSmall) + 255) & ~(size_t)255; }
size_t nb_tc_packed_bytes(const nb_mlp_desc&) { return nb_tc_fold_offset() + kFoldFloats * sizeof(float); }

TcStash nb_tc_stash_layout(long long P) {
  TcStash s;
  const size_t T = (size_t)((P + 127) / 128);
  size_t off = 0;
  s.off_embx = off; off += T * kBlobBytes;
  s.off_embd = off; off += T * kBlobBytes;
  for (int i = 0; i < 8; ++i) { s.off_h[i] = off; off += T * 4 * kBlobBytes; }
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
  for (int kb = 0; kb < 4; ++kb) add_blob(pp, off, 0, 256, 0, 0, 64 * kb, 128, 128, 256, 1);   // step 8: h7 columns through the folded W'
  add_blob(pp, off, L.wd + 256, 283, 0, 0, 0, 128, 128, 27);                                   //         view-dir PE columns
  if (off != nb_tc_fwd_packed_bytes()) { NB_SET_ERR(h, "nb_tc_pack: internal layout mismatch"); return NB_ERR_INVALID; }
  // ---- backward (dgrad) blobs: B = W^T
  nb_tc_bwd_add_blobs(L, [&](size_t src, int ld, int tr, int n0, int k0, int rows, int n_lim, int k_lim, int folded) {
    add_blob(pp, off, src, ld, tr, n0, k0, rows, n_lim, k_lim, folded);
  });
  if (pp.n > kMaxBlobs) { NB_SET_ERR(h, "nb_tc_pack: too many blobs"); return NB_ERR_INVALID; }
  float* fold = reinterpret_cast<float*>((uint8_t*)packed + nb_tc_fold_offset());
  fold_feat_kernel<<<128, 256, 0, st>>>(params, L, fold);
  NB_LAUNCHED(h);
  dim3 grid(4, pp.n + 1);
  pack_kernel<<<grid, 256, 0, st>>>(pp, params, (uint8_t*)packed, L, (uint32_t)nb_tc_small_offset(), fold);
  NB_LAUNCHED(h);
  return NB_OK;
}

// Tensor map over the packed forward blobs for the pair mode, encoded through the driver entry point (no libcuda link).
static int launch_fwd(nb_handle_t h, FwdParams& fp, bool train, cudaStream_t st) {
  const bool dbg = fp.dbg != nullptr;
  // cluster mode: 0 = independent CTAs, 2 = clusters of two CTAs sharing the weight stream by multicast (default)
  static int mode_env = -1;
  if (mode_env < 0) { const char* e = getenv("NB_TC_CLUSTER"); mode_env = e ? atoi(e) : 2; }
  const int mode = mode_env;
  if (mode != 0 && mode != 2) { NB_SET_ERR(h, "NB_TC_CLUSTER must be 0 or 2"); return NB_ERR_INVALID; }
  const bool cta2 = mode != 0;      // launched as clusters of 2
  typedef void (*kern_t)(const FwdParams);
  kern_t kern;
  if (mode == 2) kern = dbg ? mlp_fwd_chain_kernel<false, true, true> : (train ? mlp_fwd_chain_kernel<true, false, true> : mlp_fwd_chain_kernel<false, false, true>);
  else kern = dbg ? mlp_fwd_chain_kernel<false, true, false> : (train ? mlp_fwd_chain_kernel<true, false, false> : mlp_fwd_chain_kernel<false, false, false>);
  const int ki = (dbg ? 2 : (train ? 1 : 0)) + 3 * mode;
  if (!h->fwd_attr_done[ki]) {
    NB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    h->fwd_attr_done[ki] = true;
  }
  NB_CUDA(h, nb_const_bank_acquire(g_fw_banks, h->device, st, &fp.bank));
  NB_CUDA(h, cudaMemcpyToSymbolAsync(c_fw, fp.wpk + nb_tc_small_offset(), sizeof(TcSmall), (size_t)fp.bank * sizeof(TcSmall),
                                     cudaMemcpyDeviceToDevice, st));
  const long long n_tiles = (fp.P + 127) / 128;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  if (cta2) {
    const long long n_pairs = (n_tiles + 1) / 2;
    long long ncl = (n_pairs + 1) / 2;                       // two pair-tiles (slots) per cluster
    const long long max_cl = h->sm_count / 2;
    if (ncl > max_cl) ncl = max_cl;
    if (ncl < 1) ncl = 1;
    cfg.gridDim = dim3((unsigned)(2 * ncl));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  } else {
    long long grid = (n_tiles + 1) / 2;
    if (grid > h->sm_count) grid = h->sm_count;
    cfg.gridDim = dim3((unsigned)grid);
  }
  static long long* prof_dev = nullptr;
  const bool prof = getenv("NB_TC_PROF") != nullptr;
  if (prof) {
    if (!prof_dev) cudaMalloc(&prof_dev, 256 * 8 * sizeof(long long));
    cudaMemsetAsync(prof_dev, 0, 256 * 8 * sizeof(long long), st);
    fp.prof = prof_dev;
  }
  NB_CUDA(h, cudaLaunchKernelEx(&cfg, kern, fp));
  NB_LAUNCHED(h);
  if (prof) {   // diagnostic only: synchronous read-back of the MMA warp's cycle counters
    static long long host[256 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost);
    double a[8] = {0}; int n = 0;
    long long tmin = 1ll << 62, tmax = 0;
    for (unsigned b = 0; b < cfg.gridDim.x; ++b) if (host[b * 8 + 3] > 0) {
      ++n; for (int k = 0; k < 8; ++k) a[k] += (double)host[b * 8 + k];
      if (host[b * 8 + 3] < tmin) tmin = host[b * 8 + 3];
      if (host[b * 8 + 3] > tmax) tmax = host[b * 8 + 3];
    }
    fprintf(stderr, "nb_tc prof: MMA loop cycles per CTA min %lld max %lld (static round-robin tiles: the kernel ends with the slowest)\n", tmin, tmax);
    fprintf(stderr, "nb_tc prof P=%lld train=%d clusters=%d: mma_wait_aready=%.0f mma_wait_wfull=%.0f mma_wait_peerfull=%.0f mma_total=%.0f "
            "epi_wait_acc=%.0f epi_body=%.0f epi_prologue=%.0f\n",
            fp.P, (int)train, (int)cta2, a[0] / n, a[1] / n, a[2] / n, a[3] / n, a[4] / n, a[5] / n, a[6] / n);
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
  { static int ab = -1; if (ab < 0) { const char* e = getenv("NB_TC_ABLATE"); ab = e ? atoi(e) : 0; } fp.abl = ab; }   // timing experiments only
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
