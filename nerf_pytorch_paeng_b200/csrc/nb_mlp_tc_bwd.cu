// K4 backward, NB_BF16 precision: autograd of model/NeRF.py:33-52 wrt the parameters (train.py:69).
//
// Two kernels per network:
//  (1) dgrad chain  -- same structure as the forward chain (nb_mlp_tc.cu): per 128-point tile the gradient
//      wrt each layer's pre-activation is produced by a chain of tcgen05 GEMMs  dY_l = (dY_{l+1} . W_{l+1}) * relu'(h_l)
//      with dY living in shared memory / TMEM; the ReLU masks come from the forward's activation stash, and
//      every dY tile is written to HBM straight from the epilogue's registers (chunk-major blobs, stash_off()) for (2).  Steps per tile:
//        prologue  dg = (d_rgb . Wc) * (g > 0)                     (CUDA cores, K=3)
//        0: dh7 = (dg . W' + dsigma (x) Wsigma) * (h7 > 0)          with the folded W' = Wd[:, :256] . Wf (nb_mlp_tc.cu)
//        1..7: dh_{l-1} = (dh_l . W_l[:, skip cols]) * (h_{l-1} > 0)   for l = 7..1
//  (2) wgrad -- dW_l = dY_l^T . X_l reduced over all points.  The stashed blobs ([128 points x 64 features] in the chunk-major
//      layout of stash_off(): [point/64][feature/8][point%64][8 features]) are exactly SWIZZLE_NONE "MN-major" UMMA operands, so
//      both A = dY_l and B = X_l are bulk-loaded (8 KB half blobs) and fed to tcgen05.mma without any transposition; the 256x256
//      fp32 accumulator of one weight matrix fills the 512 TMEM columns.  The kernel is HBM-bound (78 operand blobs per tile, no
//      reuse), so it is built around the byte stream:
//        * 11 (layer, input-block) jobs; units of 64 points are claimed in chunks from per-job counters, every CTA starting on a
//          home job chosen by byte share and moving on to the job with most bytes left (see the comment above mlp_wgrad_kernel);
//        * a 216 KB operand ring cut into 3..6 stages by the job's unit size, so every job keeps ~200 KB in flight;
//        * the folded feature + view layers are ONE job G = dg^T [h7 | PE(d)] (second accumulator region for the PE block;
//          fold_grads_kernel maps G to dWf, dbf and dWd[:, :256]) and the density head rides on that job's h7 operand;
//        * bias gradients: from the tensor cores where the B operand has a constant 1.0 column (the pad column of the PE blobs),
//          else column sums of the staged dY tiles by the eight flush warps;
//        * results are added to the flat fp32 gradient with red.global.add (v4 where the rows are 16-byte aligned).
//      Measured on the fine pass of a 4096-ray step (7.85 GB): 1.37 ms with a static equal-byte split and a 3-stage ring for every
//      job, 1.08-1.18 ms now = the 7.2 TB/s the same ring streams with no tensor work at all (scripts/wgrad_probe.cu).
#include <stdlib.h>
#include "nb_mlp_tc.h"
#include "nb_tc_common.cuh"

using namespace tc;

namespace {

constexpr int kBwdSteps = 8;
__host__ __device__ constexpr int bwd_nkb(int b) { return b == 0 ? 2 : 4; }
__host__ __device__ constexpr uint32_t bwd_w_off(int b) {
  uint32_t o = 0;
  for (int i = 0; i < b; ++i) o += (uint32_t)bwd_nkb(i) * 32768u;
  return o;
}

// workspace: dY blobs per tile
struct BwdWs {
  size_t off_draw;     // [T][2]  d_raw (cols 0..3) as a blob, stored twice: A operand (M=128) of the head wgrad jobs
  size_t off_dg;       // [T][2]
  size_t off_dh[8];    // [T][4]  dh0..dh7
  size_t total;
};
BwdWs bwd_ws_layout(long long P) {
  BwdWs w;
  const size_t T = (size_t)((P + 127) / 128);
  size_t off = 0;
  w.off_draw = off; off += T * 2 * kBlobBytes;
  w.off_dg = off; off += T * 2 * kBlobBytes;
  for (int i = 0; i < 8; ++i) { w.off_dh[i] = off; off += T * 4 * kBlobBytes; }
  w.total = off;
  return w;
}

// ------------------------------------------------------------------------------------------
// (1) dgrad chain
// ------------------------------------------------------------------------------------------
constexpr uint32_t kActBytes = 4 * kBlobBytes;
constexpr uint32_t kOffAct = 0;
constexpr uint32_t kDgStages = 3;                 // W^T ring: 3 stages of one K block (two 32-k sub-blobs, 32 KB); the chain needs no aux tile, so the
constexpr uint32_t kDgStageBytes = 32768;         // ring gets that space (a third stage: the forward's two leave its MMA warp waiting 14-21 % of the time)
constexpr uint32_t kOffW = 2 * kActBytes;
constexpr uint32_t kOffBar = kOffW + kDgStages * kDgStageBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;
constexpr int kThreads = 576;      // warp 0 producer, warp 1 MMA, 8 epilogue warps per slot (two per TMEM lane quarter: column halves)
constexpr int kEpi = 256;

__constant__ TcSmall c_bw[kConstBanks];   // small fp32 parameters (sigma / rgb head weights) of the network being differentiated, one bank per stream in flight
NbConstBankTable g_bw_banks;

struct DgradParams {
  long long P;
  const uint8_t* wpk;       // packed dgrad blobs (W^T)
  const float* prm;
  NbParamLayout L;
  const float* d_raw;       // [P,4]
  const uint8_t* stash;     // forward activations
  TcStash st;
  uint8_t* ws;              // dY blobs out
  BwdWs w;
  int abl;                  // NB_TC_ABLATE experiments: 64 = dY tiles written to an L2-resident window
  int bank;                 // which copy of c_bw holds this network's constants (nb_cbank.h)
};

// MC: clusters of two CTAs sharing the W^T stream by multicast (see mlp_fwd_chain_kernel)
template <bool MC>
__global__ void __launch_bounds__(kThreads, 1)
mlp_dgrad_chain_kernel(const DgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_act = sbase + kOffAct, s_w = sbase + kOffW, s_bar = sbase + kOffBar;
  const uint32_t b_wfull = s_bar, b_wempty = s_bar + 64, b_aready = s_bar + 128, b_accready = s_bar + 144, s_tmem = s_bar + 160;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (p.P + 127) / 128;
  const uint32_t rank = MC ? cluster_ctarank() : 0u;
  // MC: the two CTAs of a cluster walk "pair tiles" in lockstep (tiles 2q + rank); the last pair may hold a ghost tile
  const long long n_units = MC ? (n_tiles + 1) / 2 : n_tiles;
  const long long ncl = MC ? gridDim.x / 2 : gridDim.x, cid = MC ? blockIdx.x / 2 : blockIdx.x;
  auto unit_of = [&](int slot, long long it) { return (it * ncl + cid) * 2 + slot; };
  const long long max_it = (n_units + 2 * ncl - 1) / (2 * ncl);

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < kDgStages; ++i) {
      mbar_init(b_wfull + 8 * i, 1);
      mbar_init(b_wempty + 8 * i, MC ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_aready + 8 * i, kEpi);
      mbar_init(b_accready + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

  if (warp == 0) {
    // ---- W^T producer: 3 x 32 KB ring, ping-pong order (slot 0's whole step, then slot 1's)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long it = 0; it < max_it; ++it) {
#pragma unroll 1
        for (int b = 0; b < kBwdSteps; ++b) {
          const uint8_t* src = p.wpk + bwd_w_off(b);
          for (int slot = 0; slot < 2; ++slot) {
            if (unit_of(slot, it) >= n_units) continue;
            for (int kb = 0; kb < bwd_nkb(b); ++kb) {
              mbar_wait(b_wempty + 8 * stage, phase ^ 1);
              mbar_expect_tx(b_wfull + 8 * stage, kDgStageBytes);
              const uint8_t* g = src + (size_t)kb * kDgStageBytes;
              if (MC) {     // my half of the stage, multicast into both CTAs' rings
#pragma unroll
                for (int i = 0; i < 2; ++i)
                  bulk_g2s_mcast(s_w + stage * kDgStageBytes + rank * 16384u + i * 8192u, g + rank * 16384u + i * 8192u, 8192u,
                                 b_wfull + 8 * stage, (uint16_t)3);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) bulk_g2s(s_w + stage * kDgStageBytes + i * 8192u, g + i * 8192u, 8192u, b_wfull + 8 * stage);
              }
              if (++stage == kDgStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, par_a[2] = {0, 0};
    const uint32_t idesc = umma_idesc(128, 256, 0, 0);
    for (long long it = 0; it < max_it; ++it) {
#pragma unroll 1
      for (int b = 0; b < kBwdSteps; ++b) {
        const int nkb = bwd_nkb(b);
        for (int slot = 0; slot < 2; ++slot) {
          if (unit_of(slot, it) >= n_units) continue;
          mbar_wait(b_aready + 8 * slot, par_a[slot]); par_a[slot] ^= 1;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(b_wfull + 8 * stage, phase);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_addr = s_act + slot * kActBytes + (uint32_t)kb * kBlobBytes;
              const uint32_t b_addr = s_w + stage * kDgStageBytes;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)      // the K block = two 32-k sub-blobs (SWIZZLE_64B images) laid end to end
                umma_ss(d_tmem, umma_desc(a_addr + k4 * 32u, 16, 1024),
                        umma_desc_sw64(b_addr + (uint32_t)(k4 >> 1) * 16384u + (uint32_t)(k4 & 1) * 32u, 16, 512), idesc, (kb | k4) ? 1u : 0u);
              if (MC) umma_commit_mcast(b_wempty + 8 * stage); else umma_commit(b_wempty + 8 * stage);
              if (kb == nkb - 1) umma_commit(b_accready + 8 * slot);
            }
            __syncwarp();
            if (++stage == kDgStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    const int slot = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;              // column half of the tile this warp handles
    const uint32_t q = (uint32_t)warp & 3u;
    const uint32_t r = q * 32u + (uint32_t)lane;
    const uint32_t act_base = s_act + slot * kActBytes;
    const uint32_t t_addr = tmem_base + ((q * 32u) << 16) + (uint32_t)slot * 256u;
    uint32_t par_acc = 0;
    for (long long it = 0; it < max_it; ++it) {
      if (unit_of(slot, it) >= n_units) break;
      const long long tile_raw = MC ? unit_of(slot, it) * 2 + rank : unit_of(slot, it);
      const bool tile_ok = tile_raw < n_tiles;               // MC: the last pair may hold a ghost tile (computed, never stored)
      const long long tile = tile_ok ? tile_raw : 0;
      const long long pt = tile_raw * 128 + r;
      const bool valid = pt < p.P;
      const long long tile_ws = (p.abl & 64) ? (tile & 63) : tile;
      // ---- prologue: dg = (d_rgb . Wc) * (g > 0) -> act K-blocks 0,1 ; d_raw blob -> aux ----
      float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) dr = __ldg(reinterpret_cast<const float4*>(p.d_raw) + pt);   // padded rows carry zero gradient
      // masks: [tile][9][2 column halves][128 rows][4 words] (nb_mlp_tc.h); mrow -> this row's words of layer 0, half 0
      const uint32_t* mrow = reinterpret_cast<const uint32_t*>(p.stash + p.st.off_mask) + ((size_t)tile * 9 * 2 * 128 + r) * 4;
      const uint4 gm0 = __ldg(reinterpret_cast<const uint4*>(mrow + (size_t)(8 * 2) * 128 * 4));       // mask of g: columns 0..63 in words x,y
      const uint4 gm1 = __ldg(reinterpret_cast<const uint4*>(mrow + (size_t)(8 * 2 + 1) * 128 * 4));   //            columns 64..127
      const uint32_t gmw[4] = {gm0.x, gm0.y, gm1.x, gm1.y};
      uint8_t* g_dg = tile_ok ? p.ws + p.w.off_dg + (size_t)tile_ws * 2 * kBlobBytes : nullptr;
      uint8_t* g_draw = tile_ok ? p.ws + p.w.off_draw + (size_t)tile_ws * 2 * kBlobBytes : nullptr;
#pragma unroll 1
      for (int c = half * 8; c < half * 8 + 8; ++c) {      // dg columns of this half (one K-block)
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = c * 8 + j;
          const float val = fmaf(dr.x, c_bw[p.bank].wc[col], fmaf(dr.y, c_bw[p.bank].wc[128 + col], dr.z * c_bw[p.bank].wc[256 + col]));
          const uint32_t gmsel = (c & 8) ? ((c & 4) ? gmw[3] : gmw[2]) : ((c & 4) ? gmw[1] : gmw[0]);
          v[j] = ((gmsel >> (31 - (col & 31))) & 1u) ? 0.f : val;
        }
        const uint32_t w0 = pack_bf16(v[0], v[1]), w1 = pack_bf16(v[2], v[3]), w2 = pack_bf16(v[4], v[5]), w3 = pack_bf16(v[6], v[7]);
        st_shared_v4(act_base + (uint32_t)(c >> 3) * kBlobBytes + sw128_chunk(r, (uint32_t)(c & 7)), w0, w1, w2, w3);   // A operand of step 0
        if (g_dg) st_global_na_v4(g_dg + (size_t)(c >> 3) * kBlobBytes + stash_off(r, (uint32_t)(c & 7)), w0, w1, w2, w3);   // wgrad operand
      }
      // d_raw as a (mostly zero) 128-feature operand for the rgb-head wgrad job: features 0..3 of the first blob, written by half 0;
      // half 1 zero-fills the second blob (two blobs so that M = 128 is addressable)
      if (g_draw) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w0 = 0, w1 = 0;
          if (half == 0 && c == 0) { w0 = pack_bf16(dr.x, dr.y); w1 = pack_bf16(dr.z, dr.w); }
          st_global_na_v4(g_draw + (size_t)half * kBlobBytes + stash_off(r, (uint32_t)c), w0, w1, 0u, 0u);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(b_aready + 8 * slot);

#pragma unroll 1
      for (int b = 0; b < kBwdSteps; ++b) {
        // ReLU mask of h_{7-b}, the layer whose pre-activation gradient this step produces: 8 words per row, fetched while the MMA runs
        uint32_t mq[8];      // set bit = inactive unit
        {
          const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(mrow + (size_t)((7 - b) * 2) * 128 * 4));
          const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(mrow + (size_t)((7 - b) * 2 + 1) * 128 * 4));
          mq[0] = m0.x; mq[1] = m0.y; mq[2] = m0.z; mq[3] = m0.w; mq[4] = m1.x; mq[5] = m1.y; mq[6] = m1.z; mq[7] = m1.w;
        }
        mbar_wait(b_accready + 8 * slot, par_acc); par_acc ^= 1;
        tc_fence_after();
        uint8_t* gdst = tile_ok ? p.ws + p.w.off_dh[7 - b] + (size_t)tile_ws * 4 * kBlobBytes : nullptr;
        // rolled on purpose: one 32-column body stays resident in the instruction cache (see nb_mlp_tc.cu)
#pragma unroll 1
        for (int c32 = half * 4; c32 < half * 4 + 4; ++c32) {
          float v[32];
          tmem_ld32(t_addr + (uint32_t)c32 * 32u, v);
          tmem_ld_wait();
          if (b == 0) {   // density head: d h7 += d sigma * W_sigma   (NeRF.py:43)
            const float* w = c_bw[p.bank].ws + c32 * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(dr.w, w[j], v[j]);
          }
          {
            const uint32_t m01 = (c32 & 1) ? mq[1] : mq[0], m23 = (c32 & 1) ? mq[3] : mq[2];
            const uint32_t m45 = (c32 & 1) ? mq[5] : mq[4], m67 = (c32 & 1) ? mq[7] : mq[6];
            const uint32_t m03 = (c32 & 2) ? m23 : m01, m47 = (c32 & 2) ? m67 : m45;
            const uint32_t m = (c32 & 4) ? m47 : m03;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((m >> (31 - j)) & 1u) v[j] = 0.f;
          }
          const uint32_t row_addr = act_base + (uint32_t)(c32 >> 1) * kBlobBytes + r * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t c = (uint32_t)((c32 & 1) * 4 + j);
            const uint32_t w0 = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]), w1 = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
            const uint32_t w2 = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), w3 = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
            if (b < kBwdSteps - 1) st_shared_v4(row_addr + ((c ^ (r & 7u)) << 4), w0, w1, w2, w3);       // A operand of the next step
            // dY tile -> workspace for wgrad, straight from the registers (chunk-major blob: 512 contiguous bytes per warp)
            if (gdst) st_global_na_v4(gdst + (size_t)(c32 >> 1) * kBlobBytes + stash_off(r, c), w0, w1, w2, w3);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        if (b < kBwdSteps - 1) mbar_arrive(b_aready + 8 * slot);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// (2) wgrad
// ------------------------------------------------------------------------------------------
constexpr int kMaxJobs = 16;
struct WgradJob {
  const uint8_t* a;      // dY blobs, [T][a_blobs] (uses the first m_blk of them, offset a_first)
  const uint8_t* b;      // X blobs,  [T][b_blobs]
  int a_blobs, a_first, m_blk;    // M = 64*m_blk (2 or 4)
  int b_blobs, b_first, n_blk;    // N = 64*n_blk (1, 2 or 4)
  float* out;            // dW + column offset, row-major [rows, ld]
  int ld, m_first, m_valid, n_valid;   // rows [m_first, m_valid) x cols [0, n_valid) are written; row m -> out + (m - m_first)*ld
  float* bias_out;       // dY column sums of columns [m_first, m_valid) or nullptr
  int bias_col, bias_reg; // bias_col >= 0: column of accumulator region bias_reg (1 = B, 2 = B2) whose B feature is the constant 1.0, i.e. the
                         // tensor cores already produce the column sums there; < 0: summed from the staged operand on the CUDA cores
  // optional second B operand sharing the same A (view layer: [features | PE(viewdir)]): N2 = 64*n2_blk, accumulated in
  // TMEM columns 256.. (needs m_blk == 2), written to out2
  const uint8_t* b2;
  int b2_blobs, b2_first, n2_blk, n2_valid;
  float* out2;
  int ld2;               // row stride of out2
  // optional rank-1 rider on the B operand (density head on the last trunk activation): sig_out[k] += sum_p d_raw[p][3] * B[p][k]
  // and sig_bias += sum_p d_raw[p][3], on the CUDA cores of the bias warps, so that B is not streamed a second time
  const float* sig_draw;
  float* sig_out;
  float* sig_bias;
  int weight;            // operand blobs streamed per unit (64 points)
  int n_stages;          // ring stages for this job: min(kWgMaxStages, kWgRingBytes / stage bytes)
  int share;             // CTAs expected on this job (grid x its share of the bytes, >= 1): sizes the shrinking claims at the job's end
  long long work_begin;  // sum of weight * n_units over the preceding jobs
};
struct WgradParams {
  WgradJob job[kMaxJobs];
  int n_jobs;
  long long n_tiles, n_points, total_work;
  int abl;
  int chunk_units;             // most units (64 points) per dynamically claimed chunk
  unsigned int* counters;      // [kMaxJobs] next unclaimed unit of every job (zeroed before the launch)
  unsigned long long* prof;    // optional [grid][4] ns time stamps (NB_TC_PROF diagnostic)
};

constexpr int kWgThreads = 320;                         // warp0 producer, warp1 MMA, warps 2-9 column sums + accumulator flush
// Operand ring: 216 KB cut into as many stages as fit the job's unit (4..8 half blobs of 8 KB = 64 points of every operand): 3 stages of
// 64 KB for the trunk layers, 3 x 56 KB for the folded view layer, 5 x 40 KB for the two PE(x) jobs, 6 x 32 KB for the rgb head.  What
// a CTA gets of the HBM stream follows the bytes it keeps in flight (measured: with 3 stages for every job the CTAs of the 32 / 40 KB
// jobs took 1.29 ms for the same byte count the 64 KB jobs streamed in 0.96 ms, and the kernel ends with its slowest CTA), so every
// job keeps 170-200 KB in flight whatever its unit size.
constexpr int kWgMaxStages = 6;
constexpr uint32_t kWgRingBytes = 216 * 1024;
constexpr uint32_t kWgOffBar = kWgRingBytes;            // full[6] | empty[6] (+64) | done (+128) | tmem slot (+136) | free (+144)
constexpr uint32_t kWgOffMeta = kWgOffBar + 192;        // [kWgMaxStages] x 8 B: what the stage holds: {job | kWgFirst / kWgLast / kWgEnd, unit index} (written by the producer)
constexpr uint32_t kWgOffSig = kWgOffBar + 256;         // [3][64 rows x 16 B] d_raw rows of the stage's points (density-head rider, <= 3 stages)
constexpr uint32_t kWgSigStages = 3;
constexpr uint32_t kWgSmemBytes = kWgOffSig + kWgSigStages * 1024 + 1024;
static_assert(kWgSmemBytes <= 232448, "wgrad ring exceeds the 227 KB of shared memory a CTA can have");
// stage descriptor word 0: job | flags
constexpr uint32_t kWgFirst = 1u << 8, kWgLast = 1u << 9, kWgEnd = 1u << 10;

// Work distribution.  A job's units are claimed in chunks of at most `chunk_units` (fewer towards the job's end) from a per-job atomic counter.  Every CTA starts on
// its HOME job -- the one that holds the start of its share of the byte-weighted line of work, so jobs get CTAs in proportion to
// their traffic -- and keeps claiming there; when the job is exhausted it flushes its accumulator and moves to the job with the
// most unclaimed bytes left.  (A static equal-byte split finished between 0.92 and 1.15 ms per CTA on the fine pass: SMs differ
// in what they get of the HBM stream, and the kernel ends with its slowest CTA.)  Only the producer thread talks to the counters;
// it publishes what every stage holds (job, unit, first / last stage of a segment, end of work) in shared memory before arming
// the stage's barrier, and the MMA / column-sum warps follow those descriptors.
__device__ __forceinline__ int wg_home_job(const WgradParams& p, long long lo) {
  int h = 0;
  for (int j = 0; j < p.n_jobs; ++j) if (p.job[j].work_begin <= lo) h = j;
  return h;
}

__global__ void __launch_bounds__(kWgThreads, 1)
mlp_wgrad_kernel(const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_bar = sbase + kWgOffBar, s_meta = sbase + kWgOffMeta;
  const uint32_t b_full = s_bar, b_empty = s_bar + 64, b_done = s_bar + 128, s_tmem = s_bar + 136, b_free = s_bar + 144;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_units = p.n_tiles * 2;          // units of work: half tiles (64 points); unit u -> tile u>>1, half u&1

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgMaxStages; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1 + 8); }
    mbar_init(b_done, 1);
    mbar_init(b_free, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));
  auto stamp = [&](int k) {
    if (p.prof) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.prof[(size_t)blockIdx.x * 4 + k] = t; }
  };
  if (threadIdx.x == 0) stamp(0);

  if (warp == 0) {
    if (lane == 0) {
      // Every barrier keeps its own phase parity (bit i of the mask = parity to wait for at the next use of stage i), because the
      // number of stages in use changes from job to job.  Before the first load of a CTA's next job the ring is drained: the new
      // job's stage slots have another size and would overlap slots the MMAs may still be reading.
      uint32_t pmask = (1u << kWgMaxStages) - 1u;
      const long long C = p.chunk_units;
      // Claim [ua, ub) of job jj: up to C units, fewer towards the end of the job (half of an even share of what `seen` -- the last
      // counter value this thread saw -- leaves), so that the CTAs of a job run out of work together instead of one chunk apart.
      auto claim = [&](int jj, long long seen, long long& ua, long long& ub) -> bool {
        const long long share = p.job[jj].share;
        long long k = (n_units - seen + 2 * share - 1) / (2 * share);
        k = k > C ? C : (k < 1 ? 1 : k);
        ua = (long long)atomicAdd(p.counters + jj, (unsigned int)k);
        ub = ua + k < n_units ? ua + k : n_units;
        return ua < n_units;
      };
      int j = wg_home_job(p, p.total_work * (long long)blockIdx.x / (long long)gridDim.x);
      bool any = false;
      long long units_done = 0;
      while (true) {
        // ---- claim the first chunk of a segment: the home job first, afterwards whichever job has the most bytes unclaimed
        long long ua, ub;
        bool got = claim(j, (long long)*reinterpret_cast<volatile unsigned int*>(p.counters + j), ua, ub);
        while (!got) {
          long long best = 0, best_taken = 0; int bj = -1;
          for (int k = 0; k < p.n_jobs; ++k) {
            const long long taken = (long long)*reinterpret_cast<volatile unsigned int*>(p.counters + k);
            const long long left = (n_units - taken) * p.job[k].weight;
            if (left > best) { best = left; bj = k; best_taken = taken; }
          }
          if (bj < 0) break;
          j = bj;
          got = claim(j, best_taken, ua, ub);
        }
        if (!got) break;
        const WgradJob& J = p.job[j];
        const uint32_t stage_bytes = (uint32_t)(J.m_blk + J.n_blk + J.n2_blk) * 8192u;
        const uint32_t ns = (uint32_t)J.n_stages;
        if (any)
          for (int i = 0; i < kWgMaxStages; ++i) mbar_wait(b_empty + 8 * i, (pmask >> i) & 1u);
        any = true;
        uint32_t stage = 0;
        bool first = true;
        while (got) {
          long long na, nb;
          const bool more = claim(j, ub, na, nb);      // next chunk, claimed early: its latency hides behind this chunk
          for (long long u = ua; u < ub; ++u) {
            const long long tile = u >> 1;
            const long long tile_a = (p.abl & 64) ? (tile & 63) : tile;      // experiments: operands from an L2-resident window
            const long long tile_b = (p.abl & 128) ? (tile & 63) : tile;
            const uint32_t half = (uint32_t)(u & 1) * 8192u;
            mbar_wait(b_empty + 8 * stage, (pmask >> stage) & 1u);
            pmask ^= 1u << stage;
            const uint32_t flags = (uint32_t)j | (first ? kWgFirst : 0u) | ((u == ub - 1 && !more) ? kWgLast : 0u);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_meta + 8u * stage), "r"(flags), "r"((uint32_t)u) : "memory");
            first = false;
            long long rows = J.sig_draw ? p.n_points - u * 64 : 0;           // density-head rider: the unit's d_raw rows ride along
            rows = rows > 64 ? 64 : (rows < 0 ? 0 : rows);
            mbar_expect_tx(b_full + 8 * stage, stage_bytes + (uint32_t)rows * 16u);
            const uint32_t dst = sbase + stage * stage_bytes;
            for (int k = 0; k < J.m_blk; ++k)
              bulk_g2s(dst + (uint32_t)k * 8192u, J.a + ((size_t)tile_a * J.a_blobs + J.a_first + k) * kBlobBytes + half, 8192u, b_full + 8 * stage);
            for (int k = 0; k < J.n_blk; ++k)
              bulk_g2s(dst + (uint32_t)(J.m_blk + k) * 8192u, J.b + ((size_t)tile_b * J.b_blobs + J.b_first + k) * kBlobBytes + half, 8192u,
                       b_full + 8 * stage);
            for (int k = 0; k < J.n2_blk; ++k)
              bulk_g2s(dst + (uint32_t)(J.m_blk + J.n_blk + k) * 8192u, J.b2 + ((size_t)tile_b * J.b2_blobs + J.b2_first + k) * kBlobBytes + half,
                       8192u, b_full + 8 * stage);
            if (rows > 0) bulk_g2s(sbase + kWgOffSig + stage * 1024u, J.sig_draw + u * 64 * 4, (uint32_t)rows * 16u, b_full + 8 * stage);
            if (++stage == ns) stage = 0;
            ++units_done;
          }
          got = more; ua = na; ub = nb;
        }
      }
      // ---- end of work: an empty stage that only carries the flag (stage 0 of a drained ring)
      if (any)
        for (int i = 0; i < kWgMaxStages; ++i) mbar_wait(b_empty + 8 * i, (pmask >> i) & 1u);
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_meta), "r"(kWgEnd), "r"(0u) : "memory");
      mbar_arrive(b_full);
      stamp(1);      // last load issued
      if (p.prof) p.prof[(size_t)blockIdx.x * 4 + 2] = (unsigned long long)units_done;
    }
  } else if (warp == 1) {
    uint32_t cmask = 0, seg = 0, stage = 0;
    while (true) {
      mbar_wait(b_full + 8 * stage, (cmask >> stage) & 1u);
      cmask ^= 1u << stage;
      uint32_t flags;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(flags) : "r"(s_meta + 8u * stage));
      if (flags & kWgEnd) break;
      const WgradJob& J = p.job[flags & 0xFFu];
      const uint32_t stage_bytes = (uint32_t)(J.m_blk + J.n_blk + J.n2_blk) * 8192u, ns = (uint32_t)J.n_stages;
      if ((flags & kWgFirst) && seg > 0) mbar_wait(b_free, (seg - 1) & 1);      // the previous segment's accumulator has been drained
      tc_fence_after();
      if (lane == 0) {
        const uint32_t idesc = umma_idesc(128, 64 * J.n_blk, 1, 1);      // both operands MN-major
        const uint32_t idesc2 = umma_idesc(128, J.n2_blk > 0 ? 64 * J.n2_blk : 64, 1, 1);
        const int m_halves = J.m_blk >> 1;
        const uint32_t a_addr = sbase + stage * stage_bytes;
        const uint32_t b_addr = a_addr + (uint32_t)J.m_blk * 8192u;
        const uint32_t acc = (flags & kWgFirst) ? 0u : 1u;
        if (!(p.abl & 512)) {      // 512: timing experiment without the MMAs (stages are released at once)
          // operands: half blobs [feature/8][64 points][8 features] laid end to end => atoms of 8 features every 1024 B (SBO),
          // 8-point groups every 128 B (LBO), a K16 slice every 256 B; 128 output rows (one m_half) = 16 atoms = 16 KB further
          for (int mh = 0; mh < m_halves; ++mh) {
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)      // 64 points per stage = 4 x K16
              umma_ss(tmem_base + (uint32_t)mh * 256u, umma_desc_mn_noswz(a_addr + (uint32_t)mh * 16384u + k16 * 256u, 128, 1024),
                      umma_desc_mn_noswz(b_addr + k16 * 256u, 128, 1024), idesc, (acc | (uint32_t)k16) ? 1u : 0u);
          }
          if (J.n2_blk > 0) {
            const uint32_t b2_addr = b_addr + (uint32_t)J.n_blk * 8192u;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              umma_ss(tmem_base + 256u, umma_desc_mn_noswz(a_addr + k16 * 256u, 128, 1024), umma_desc_mn_noswz(b2_addr + k16 * 256u, 128, 1024),
                      idesc2, (acc | (uint32_t)k16) ? 1u : 0u);
          }
        }
        umma_commit(b_empty + 8 * stage);
        if (flags & kWgLast) umma_commit(b_done);
      }
      __syncwarp();
      if (flags & kWgLast) { ++seg; stage = 0; } else if (++stage == ns) stage = 0;
    }
  } else {
    // ---- column sums (bias gradients, density-head rider) from the staged operands, then the TMEM -> global flush: EIGHT warps ----
    // (two per scheduler: with one warp per scheduler the dependent unpack / add chains ran at ~0.25 IPC and the jobs with the most
    // sums per byte -- the PE(x) jobs and the folded view layer -- released their stages late: their CTAs finished 7-12 % after the
    // others.)  Staged operand = consecutive 1 KB atoms [64 points][8 features].  Eight consecutive lanes read 8 consecutive points
    // of ONE atom (128 contiguous bytes: conflict-free), so a lane accumulates partial column sums over the points j8, j8+8, .. of
    // its atom; the eight partial sums are combined by shuffles once per job segment.  Jobs whose B operand carries a constant 1.0
    // column (the pad column of PE(x) / PE(viewdir), written by the training forward) get their dY column sums from the TENSOR
    // CORES instead: that accumulator column IS sum_p dY[p][m] (bias_col >= 0), and the flush adds it to the bias gradient.
    const int t = threadIdx.x - 64;           // 0..255
    const int j8 = t & 7;                     // point j8 + 8*i of the stage
    const int q32 = t >> 3;                   // 32 groups of 8 lanes: group g owns atom g
    const uint32_t q = (uint32_t)warp & 3u;   // TMEM lane quarter this warp may read
    const uint32_t csel = ((uint32_t)warp - 2u) >> 2;     // the two warps of a quarter take alternate 32-column chunks of the flush
    uint32_t cmask = 0, seg = 0, stage = 0;
    float bs[8], sg[8], sgb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { bs[k] = 0.f; sg[k] = 0.f; }
    while (true) {
      mbar_wait(b_full + 8 * stage, (cmask >> stage) & 1u);
      cmask ^= 1u << stage;
      uint32_t flags, u32;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(flags), "=r"(u32) : "r"(s_meta + 8u * stage));
      if (flags & kWgEnd) break;
      const WgradJob& J = p.job[flags & 0xFFu];
      const uint32_t stage_bytes = (uint32_t)(J.m_blk + J.n_blk + J.n2_blk) * 8192u, ns = (uint32_t)J.n_stages;
      const int a_atoms = 8 * J.m_blk;                      // dY operand: 16 or 32 atoms
      const bool do_bias = J.bias_out != nullptr && J.bias_col < 0 && !(p.abl & 256);      // 256: timing experiment without the sums
      const bool do_sig = J.sig_draw != nullptr && !(p.abl & 256);       // rider on the B operand (N = 256 = 32 atoms)
      const uint32_t st_base = sbase + stage * stage_bytes;
      if (do_bias && q32 < a_atoms) {
        const uint32_t base = st_base + (uint32_t)q32 * 1024u + (uint32_t)j8 * 16u;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(base + i * 128u));
          bs[0] += __uint_as_float(w0 << 16); bs[1] += __uint_as_float(w0 & 0xFFFF0000u);
          bs[2] += __uint_as_float(w1 << 16); bs[3] += __uint_as_float(w1 & 0xFFFF0000u);
          bs[4] += __uint_as_float(w2 << 16); bs[5] += __uint_as_float(w2 & 0xFFFF0000u);
          bs[6] += __uint_as_float(w3 << 16); bs[7] += __uint_as_float(w3 & 0xFFFF0000u);
        }
      }
      if (do_sig) {
        // d_sigma of point r of the stage = float 3 of staged d_raw row r; rows past the last point were not copied and count as zero
        const uint32_t base = st_base + (uint32_t)(a_atoms + q32) * 1024u + (uint32_t)j8 * 16u;
        const uint32_t sig = sbase + kWgOffSig + stage * 1024u;
        const long long left = p.n_points - (long long)u32 * 64;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
          uint32_t w0, w1, w2, w3;
          float ds;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(base + i * 128u));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ds) : "r"(sig + ((uint32_t)j8 + 8u * i) * 16u + 12u));
          if ((long long)(j8 + 8 * (int)i) >= left) ds = 0.f;
          sg[0] = fmaf(ds, __uint_as_float(w0 << 16), sg[0]); sg[1] = fmaf(ds, __uint_as_float(w0 & 0xFFFF0000u), sg[1]);
          sg[2] = fmaf(ds, __uint_as_float(w1 << 16), sg[2]); sg[3] = fmaf(ds, __uint_as_float(w1 & 0xFFFF0000u), sg[3]);
          sg[4] = fmaf(ds, __uint_as_float(w2 << 16), sg[4]); sg[5] = fmaf(ds, __uint_as_float(w2 & 0xFFFF0000u), sg[5]);
          sg[6] = fmaf(ds, __uint_as_float(w3 << 16), sg[6]); sg[7] = fmaf(ds, __uint_as_float(w3 & 0xFFFF0000u), sg[7]);
          if (q32 == 0) sgb += ds;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(b_empty + 8 * stage);
      if (!(flags & kWgLast)) { if (++stage == ns) stage = 0; continue; }
      stage = 0;
      // ---- last stage of a segment: combine the eight point-interleaved partial sums of the atom (lanes j8 = 0..7); lane j8 adds column j8
      if (do_bias || do_sig) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            bs[k] += __shfl_xor_sync(0xffffffffu, bs[k], o);
            sg[k] += __shfl_xor_sync(0xffffffffu, sg[k], o);
          }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) sgb += __shfl_xor_sync(0xffffffffu, sgb, o);
      }
      if (do_bias) {
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k == j8) mine = bs[k];
        const int col = q32 * 8 + j8;
        if (q32 < a_atoms && col >= J.m_first && col < J.m_valid) atomicAdd(J.bias_out + col - J.m_first, mine);
      }
      if (do_sig) {
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k == j8) mine = sg[k];
        atomicAdd(J.sig_out + q32 * 8 + j8, mine);
        if (t == 0) atomicAdd(J.sig_bias, sgb);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { bs[k] = 0.f; sg[k] = 0.f; }
      sgb = 0.f;
      // accumulators -> flat gradient
      mbar_wait(b_done, seg & 1);
      tc_fence_after();
      const int m_halves = J.m_blk >> 1;
      const bool vec4 = (J.ld & 3) == 0 && ((uintptr_t)J.out & 15) == 0 && (J.n_valid & 3) == 0;
      const bool ones1 = J.bias_out != nullptr && J.bias_col >= 0 && J.bias_reg == 1;      // dY column sums = accumulator column bias_col
      const bool ones2 = J.bias_out != nullptr && J.bias_col >= 0 && J.bias_reg == 2;
      for (int mh = 0; mh < m_halves; ++mh) {
        const int m = mh * 128 + (int)(q * 32u) + lane;          // output row (out-feature)
        for (int c32 = (int)csel; c32 < 2 * J.n_blk; c32 += 2) {
          float v[32];
          tmem_ld32(tmem_base + ((q * 32u) << 16) + (uint32_t)mh * 256u + (uint32_t)c32 * 32u, v);
          tmem_ld_wait();
          if (m >= J.m_first && m < J.m_valid) {
            float* dst = J.out + (size_t)(m - J.m_first) * J.ld + c32 * 32;
            if (vec4) {
#pragma unroll
              for (int jj = 0; jj < 32; jj += 4)
                if (c32 * 32 + jj < J.n_valid)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + jj), "f"(v[jj]), "f"(v[jj + 1]), "f"(v[jj + 2]),
                               "f"(v[jj + 3]) : "memory");
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (c32 * 32 + jj < J.n_valid) atomicAdd(dst + jj, v[jj]);
            }
            if (ones1 && (J.bias_col >> 5) == c32) {
              float b = 0.f;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) if (jj == (J.bias_col & 31)) b = v[jj];
              atomicAdd(J.bias_out + m - J.m_first, b);
            }
          }
        }
      }
      for (int c32 = (int)csel; c32 < 2 * J.n2_blk; c32 += 2) {            // second B operand (m_blk == 2): TMEM columns 256..
        const int m = (int)(q * 32u) + lane;
        float v[32];
        tmem_ld32(tmem_base + ((q * 32u) << 16) + 256u + (uint32_t)c32 * 32u, v);
        tmem_ld_wait();
        if (m >= J.m_first && m < J.m_valid) {
          float* dst = J.out2 + (size_t)(m - J.m_first) * J.ld2 + c32 * 32;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (c32 * 32 + jj < J.n2_valid) atomicAdd(dst + jj, v[jj]);
          if (ones2 && (J.bias_col >> 5) == c32) {
            float b = 0.f;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) if (jj == (J.bias_col & 31)) b = v[jj];
            atomicAdd(J.bias_out + m - J.m_first, b);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_free);       // the MMA warp may overwrite the accumulator for this CTA's next segment
      ++seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(3);   // accumulators flushed
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// (3) gradients of the folded layers from G = dg^T h7 [128][256] and s = column sums of dg [128]  (fp32, 25 MFLOP)
//     feat = Wf h7 + bf,  pre_g = Wd_a feat + Wd_b PE(d) + bd   (Wd_a = Wd[:, :256])
//       dWf   = Wd_a^T G              dWd_a = G Wf^T + s (x) bf            dbf = Wd_a^T s            dbd = s
// ------------------------------------------------------------------------------------------
// grid = 128 (dWf, db_feat: two rows j per block) + 512 (dWd_a: one row n and a quarter of k per block) + 1 (db_d); every sum keeps four independent
// accumulators, and the Wf tiles of the second part go through shared memory so that both the global reads (rows of 128 B) and the
// per-thread reads (stride 33 words) are conflict-free (the first version read Wf rows with a 1 KB stride per lane: 74 us under ncu).
__global__ void __launch_bounds__(256)
fold_grads_kernel(const float* __restrict__ prm, NbParamLayout L, const float* __restrict__ fold_g, float* __restrict__ grad) {
  const float* G = fold_g;
  const float* sdg = fold_g + 128 * 256;
  const int t = threadIdx.x;
  __shared__ float sh[256];
  __shared__ float tile[256 * 33];
  if (blockIdx.x < 128) {                      // dWf[j][:] += sum_n Wd_a[n][j] G[n][:]  ;  dbf[j] += sum_n Wd_a[n][j] s[n]   (j = 2b, 2b+1)
    const int j0 = 2 * (int)blockIdx.x;
    sh[t] = prm[L.wd + (size_t)(t & 127) * 283 + j0 + (t >> 7)];      // sh[0..127] = Wd[:, j0], sh[128..255] = Wd[:, j0+1]
    __syncthreads();
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 16
    for (int n = 0; n < 128; n += 2) {
      const float g0 = G[(size_t)n * 256 + t], g1 = G[(size_t)(n + 1) * 256 + t];
      a0 = fmaf(sh[n], g0, a0); a1 = fmaf(sh[n + 1], g1, a1);
      b0 = fmaf(sh[128 + n], g0, b0); b1 = fmaf(sh[128 + n + 1], g1, b1);
    }
    grad[L.wf + (size_t)j0 * 256 + t] += a0 + a1;
    grad[L.wf + (size_t)(j0 + 1) * 256 + t] += b0 + b1;
    if (t < 64) {                              // warp 0 / warp 1: db_feat[j0] / db_feat[j0+1]
      const int w = t >> 5, l = t & 31;
      float b = 0.f;
      for (int n = l; n < 128; n += 32) b = fmaf(sh[128 * w + n], sdg[n], b);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
      if (l == 0) grad[L.bf + j0 + w] += b;
    }
  } else if (blockIdx.x < 128 + 512) {         // dWd_a[n][j] += sum_k G[n][k] Wf[j][k] + s[n] bf[j]   (thread = j; k in four slices)
    const int n = ((int)blockIdx.x - 128) >> 2, ks = ((int)blockIdx.x - 128) & 3;
    const int warp = t >> 5, lane = t & 31;
    sh[t] = G[(size_t)n * 256 + t];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k0 = 64 * ks; k0 < 64 * ks + 64; k0 += 32) {
      __syncthreads();                         // previous tile consumed (and sh[] visible on the first pass)
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const int j = warp * 32 + i;
        tile[j * 33 + lane] = prm[L.wf + (size_t)j * 256 + k0 + lane];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        a0 = fmaf(sh[k0 + kk], tile[t * 33 + kk], a0);
        a1 = fmaf(sh[k0 + kk + 1], tile[t * 33 + kk + 1], a1);
        a2 = fmaf(sh[k0 + kk + 2], tile[t * 33 + kk + 2], a2);
        a3 = fmaf(sh[k0 + kk + 3], tile[t * 33 + kk + 3], a3);
      }
    }
    atomicAdd(grad + L.wd + (size_t)n * 283 + t, (a0 + a1) + (a2 + a3) + (ks == 0 ? sdg[n] * prm[L.bf + t] : 0.f));
  } else {                                     // dbd += s
    if (t < 128) grad[L.bd + t] += sdg[t];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
size_t nb_tc_bwd_packed_bytes() { return bwd_w_off(kBwdSteps); }
size_t nb_tc_bwd_ws_bytes(const nb_mlp_desc&, long long P) { return bwd_ws_layout(P).total + kFoldFloats * sizeof(float) + 256; }   // 256 >= chunk counters

void nb_tc_bwd_add_blobs(const NbParamLayout& L, const std::function<void(size_t, int, int, int, int, int, int, int, int)>& add) {
  // blob[n][k] = W[k0+k][n0+n]  (B = W^T, rows = in-features, K = out-features)
  for (int kb = 0; kb < 2; ++kb) add(0, 256, 1, 0, 64 * kb, 256, 256, 128, 1);              // step 0: the folded W' = Wd[:, :256] . Wf  (128 x 256)
  const int layers[7] = {7, 6, 5, 4, 3, 2, 1};                                              // steps 1..7
  for (int i = 0; i < 7; ++i) {
    const int l = layers[i];
    const size_t src = L.w[l] + (l == 5 ? 63 : 0);                                           // skip layer: the h columns of W5
    for (int kb = 0; kb < 4; ++kb) add(src, L.in_dim[l], 1, 0, 64 * kb, 256, 256, 256, 0);
  }
}

int nb_tc_backward(nb_handle_t h, const nb_mlp_desc* d, const float* params, const void* packed, int64_t P,
                   const void* act_save, const float* d_raw, float* grad, int accumulate, void* ws, size_t ws_bytes,
                   cudaStream_t st, int stages) {
  const NbParamLayout L = nb_param_layout(*d);
  const BwdWs W = bwd_ws_layout(P);
  if (!ws || ws_bytes < W.total + kFoldFloats * sizeof(float) + 256) {
    NB_SET_ERR(h, "mlp bf16 backward: workspace %zu < %zu bytes", ws_bytes, W.total + kFoldFloats * sizeof(float) + 256);
    return NB_ERR_WORKSPACE;
  }
  NB_REQUIRE(h, ((uintptr_t)d_raw & 15) == 0 && ((uintptr_t)ws & 15) == 0 && ((uintptr_t)act_save & 15) == 0,
             "mlp bf16 backward: d_raw / act_save / ws must be 16-byte aligned");
  if (!accumulate && (stages & 1)) NB_CUDA(h, cudaMemsetAsync(grad, 0, L.total * sizeof(float), st));
  const TcStash S = nb_tc_stash_layout(P);
  const long long n_tiles = S.tiles;
  if (!h->bwd_attr_done) {
    NB_CUDA(h, cudaFuncSetAttribute(mlp_dgrad_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(h, cudaFuncSetAttribute(mlp_dgrad_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(h, cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemBytes));
    h->bwd_attr_done = true;
  }
  // ---- (1) dgrad chain
  if (stages & 1) {
    DgradParams dp;
    memset(&dp, 0, sizeof(dp));
    dp.P = P; dp.wpk = (const uint8_t*)packed + nb_tc_fwd_packed_bytes(); dp.prm = params; dp.L = L; dp.d_raw = d_raw;
    dp.stash = (const uint8_t*)act_save; dp.st = S; dp.ws = (uint8_t*)ws; dp.w = W;
    { const char* e = getenv("NB_TC_ABLATE"); dp.abl = e ? atoi(e) : 0; }   // timing experiments only
    NB_CUDA(h, nb_const_bank_acquire(g_bw_banks, h->device, st, &dp.bank));
    NB_CUDA(h, cudaMemcpyToSymbolAsync(c_bw, (const uint8_t*)packed + nb_tc_small_offset(), sizeof(TcSmall),
                                       (size_t)dp.bank * sizeof(TcSmall), cudaMemcpyDeviceToDevice, st));
    static int mode_env = -1;
    if (mode_env < 0) { const char* e = getenv("NB_TC_CLUSTER"); mode_env = e ? atoi(e) : 2; }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    if (mode_env == 2) {      // clusters of 2 CTAs sharing the weight stream by multicast
      const long long n_pairs = (n_tiles + 1) / 2;
      long long ncl = (n_pairs + 1) / 2;
      if (ncl > h->sm_count / 2) ncl = h->sm_count / 2;
      if (ncl < 1) ncl = 1;
      cfg.gridDim = dim3((unsigned)(2 * ncl));
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      NB_CUDA(h, cudaLaunchKernelEx(&cfg, mlp_dgrad_chain_kernel<true>, dp));
    } else {
      long long grid = (n_tiles + 1) / 2;
      if (grid > h->sm_count) grid = h->sm_count;
      cfg.gridDim = dim3((unsigned)grid);
      NB_CUDA(h, cudaLaunchKernelEx(&cfg, mlp_dgrad_chain_kernel<false>, dp));
    }
    NB_LAUNCHED(h);
  }
  // ---- (2) wgrad jobs
  if (stages & 2) {
    WgradParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.n_tiles = n_tiles;
    { const char* e = getenv("NB_TC_ABLATE"); wp.abl = e ? atoi(e) : 0; }   // timing experiments only
    const uint8_t* stash = (const uint8_t*)act_save;
    const uint8_t* w8 = (const uint8_t*)ws;
    int weight[kMaxJobs];
    int nj = 0;
    auto add = [&](const uint8_t* a, int a_blobs, int a_first, int m_blk, const uint8_t* b, int b_blobs, int b_first, int n_blk,
                   float* out, int ld, int m_first, int m_valid, int n_valid, float* bias) {
      WgradJob& j = wp.job[nj];
      j.a = a; j.a_blobs = a_blobs; j.a_first = a_first; j.m_blk = m_blk; j.b = b; j.b_blobs = b_blobs; j.b_first = b_first;
      j.n_blk = n_blk; j.out = out; j.ld = ld; j.m_first = m_first; j.m_valid = m_valid; j.n_valid = n_valid; j.bias_out = bias;
      j.bias_col = -1; j.bias_reg = 0;
      weight[nj] = m_blk + n_blk;       // HBM bytes per point ~ blobs loaded
      ++nj;
    };
    const uint8_t* dh[8];
    for (int i = 0; i < 8; ++i) dh[i] = w8 + W.off_dh[i];
    // trunk layers
    // PE(x) blobs carry 1.0 in their pad column 63 (PE(viewdir): column 27), so db = dY^T 1 comes out of the same MMAs
    add(dh[0], 4, 0, 4, stash + S.off_embx, 1, 0, 1, grad + L.w[0], 63, 0, 256, 63, grad + L.b[0]);
    wp.job[nj - 1].bias_col = 63; wp.job[nj - 1].bias_reg = 1;
    for (int l = 1; l < 8; ++l) {
      if (l == 5) {
        add(dh[5], 4, 0, 4, stash + S.off_embx, 1, 0, 1, grad + L.w[5], 319, 0, 256, 63, grad + L.b[5]);
        wp.job[nj - 1].bias_col = 63; wp.job[nj - 1].bias_reg = 1;
        add(dh[5], 4, 0, 4, stash + S.off_h[4], 4, 0, 4, grad + L.w[5] + 63, 319, 0, 256, 256, nullptr);
      } else {
        add(dh[l], 4, 0, 4, stash + S.off_h[l - 1], 4, 0, 4, grad + L.w[l], 256, 0, 256, 256, grad + L.b[l]);
      }
    }
    // folded feature + view layers: ONE job G = dg^T [h7 | PE(viewdir)].  Its h7 block (128 x 256, into the fold scratch) carries the
    // gradients of Wf, b_feat and Wd[:, :256] (fold_grads_kernel below); its PE block is dWd[:, 256:283] directly.  The density head
    // (dW_sigma = d_sigma^T h7, db_sigma) rides on the h7 operand.  dg is read once, h7 once: 7 blobs where the unfolded layers took 15.
    float* fold_g = reinterpret_cast<float*>((uint8_t*)ws + W.total);              // [128][256] G  +  [128] column sums of dg
    NB_CUDA(h, cudaMemsetAsync(fold_g, 0, kFoldFloats * sizeof(float) + kMaxJobs * sizeof(unsigned int), st));   // + the chunk counters
    add(w8 + W.off_dg, 2, 0, 2, stash + S.off_h[7], 4, 0, 4, fold_g, 256, 0, 128, 256, fold_g + 128 * 256);
    { WgradJob& j = wp.job[nj - 1]; j.b2 = stash + S.off_embd; j.b2_blobs = 1; j.b2_first = 0; j.n2_blk = 1; j.n2_valid = 27;
      j.out2 = grad + L.wd + 256; j.ld2 = 283; weight[nj - 1] += 1; j.bias_col = 27; j.bias_reg = 2;
      j.sig_draw = d_raw; j.sig_out = grad + L.ws; j.sig_bias = grad + L.bs; }
    // rgb head: A = d_raw blob (cols 0..2 = d_rgb), stored twice so that M = 128 is addressable
    add(w8 + W.off_draw, 2, 0, 2, stash + S.off_g, 2, 0, 2, grad + L.wc, 128, 0, 3, 128, grad + L.bc);          // dWc, dbc
    wp.n_jobs = nj;
    wp.n_points = P;
    // the byte-weighted line of work (work_begin / total_work) only picks every CTA's HOME job; units are claimed dynamically
    long long work = 0;
    for (int j = 0; j < nj; ++j) {
      wp.job[j].weight = weight[j]; wp.job[j].work_begin = work; work += (long long)weight[j] * n_tiles * 2;
      int fit = (int)(kWgRingBytes / ((uint32_t)weight[j] * 8192u));
      if (wp.job[j].sig_draw && fit > (int)kWgSigStages) fit = (int)kWgSigStages;
      wp.job[j].n_stages = fit > kWgMaxStages ? kWgMaxStages : fit;
    }
    wp.total_work = work;
    // chunks of 16 units (1024 points) when every CTA gets >= 32 of them per job visit, else 8 / 4 / 1
    {
      const long long per_cta = n_tiles * 2 * nj / (h->sm_count > 0 ? h->sm_count : 1);
      wp.chunk_units = per_cta >= 512 ? 16 : (per_cta >= 128 ? 8 : (per_cta >= 16 ? 4 : 1));
    }
    wp.counters = reinterpret_cast<unsigned int*>(fold_g + kFoldFloats);          // zeroed with the fold scratch above
    {
      int grid_n = h->sm_count;
      if ((long long)grid_n > n_tiles * 2) grid_n = (int)(n_tiles * 2);
      for (int j = 0; j < nj; ++j) {
        const long long sh = (long long)grid_n * weight[j] * n_tiles * 2 / (work > 0 ? work : 1);
        wp.job[j].share = sh < 1 ? 1 : (int)sh;
      }
    }
    int begin = h->sm_count;
    if ((long long)begin > n_tiles * 2) begin = (int)(n_tiles * 2);
    static unsigned long long* prof_dev = nullptr;
    const bool prof = getenv("NB_TC_PROF") != nullptr;
    if (prof) {
      if (!prof_dev) cudaMalloc(&prof_dev, 256 * 4 * sizeof(unsigned long long));
      cudaMemsetAsync(prof_dev, 0, 256 * 4 * sizeof(unsigned long long), st);
      wp.prof = prof_dev;
    }
    mlp_wgrad_kernel<<<begin, kWgThreads, kWgSmemBytes, st>>>(wp);
    NB_LAUNCHED(h);
    if (prof) {   // diagnostic only: synchronous read-back of the per-CTA time stamps (ns since the earliest CTA start)
      static unsigned long long host[256 * 4];
      cudaStreamSynchronize(st);
      cudaMemcpy(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull;
      for (int b = 0; b < begin; ++b) if (host[b * 4] < t0) t0 = host[b * 4];
      fprintf(stderr, "nb_tc wgrad prof P=%lld grid=%d chunk=%d units (cta: home job, units done | start, end of loads, flushed [us])\n", (long long)P, begin,
              wp.chunk_units);
      for (int b = 0; b < begin; ++b) {
        const long long lo = work * b / begin;
        int j0 = 0;
        for (int j = 0; j < nj; ++j) if (wp.job[j].work_begin <= lo) j0 = j;
        fprintf(stderr, "  cta %3d: jobs %2d..%2d | %7.1f %7.1f %7.1f %7.1f\n", b, j0, j0, (host[b * 4] - t0) * 1e-3, (host[b * 4 + 1] - t0) * 1e-3,
                (double)host[b * 4 + 2], (host[b * 4 + 3] - t0) * 1e-3);
      }
    }
    fold_grads_kernel<<<128 + 512 + 1, 256, 0, st>>>(params, L, fold_g, grad);
    NB_LAUNCHED(h);
  }
  return NB_OK;
}
