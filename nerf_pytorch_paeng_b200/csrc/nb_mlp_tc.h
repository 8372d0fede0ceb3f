// Shared between the forward (nb_mlp_tc.cu) and backward (nb_mlp_tc_bwd.cu) tensor-core MLP kernels.
#pragma once
#include <functional>
#include "nb_mlp.h"
#include "nb_cbank.h"

constexpr uint32_t kBlobBytes = 16384;   // one K-block of a 128-point tile: 128 points x 64 bf16 (shared memory: 128B-swizzled K-major rows;
                                         // stash / dY blobs in HBM: chunk-major, stash_off() in nb_tc_common.cuh)
constexpr int kFwdSteps = 9;        // the activation-free feature layer is folded into the view layer (W' = Wd[:, :W] . Wf, see nb_mlp_tc.cu)
constexpr uint32_t kMaskTileBytes = 9 * 128 * 32;   // per tile

// Activation stash written by the training forward: per tensor, per tile, consecutive 16 KB blobs in the chunk-major layout
// [point/64][feature/8][point%64][8 features] (written straight from the epilogue's registers with coalesced 16-byte stores;
// each 8 KB half is a SWIZZLE_NONE MN-major UMMA operand, so the weight-gradient kernel bulk-loads them as they are).
struct TcStash {
  size_t off_embx, off_embd;   // PE(x) 63(+1.0 pad) cols, PE(d) 27(+1.0 pad) cols : 1 blob / tile
  size_t off_h[8];             // post-ReLU trunk outputs h0..h7                   : 4 blobs / tile
  size_t off_g;                // post-ReLU view layer output (128 wide)           : 2 blobs / tile
  size_t off_mask;             // ReLU masks: [tile][9 = h0..h7, g][2 column halves][128 rows][4 x u32]; bit (31-j) of word w of half h = sign bit of
                               // column 128h+32w+j (g, 128 wide: 64h+32w+j, words 2,3 unused), set = inactive
  size_t total;
  long long tiles;
};
TcStash nb_tc_stash_layout(long long P);

// Small fp32 parameters every epilogue thread needs for every column (biases, sigma / rgb head weights).  With
// 227 KB of shared memory per CTA the L1 is empty, so reading them through the global path costs an L2 round
// trip per chunk; they are gathered once per weight update (nb_mlp_pack) behind the packed blobs and copied to
// __constant__ memory before each launch, where warp-uniform reads are broadcast from the constant cache.
struct TcSmall {
  float bias[9][256];    // forward chain steps 0..8: b0..b7, b' = Wd[:, :W] . b_feat + b_d (128 used)
  float ws[256];         // linear_density.weight
  float wc[384];         // linear_color.weight [3][128]
  float bc[4];           // linear_color.bias[0..2], linear_density.bias
};
size_t nb_tc_small_offset();   // byte offset of the TcSmall block inside the packed buffer
size_t nb_tc_fold_offset();    // byte offset of the folded fp32 matrix W' [W/2][W] (+ b' [W/2]) inside the packed buffer
constexpr size_t kFoldFloats = 128 * 256 + 128;

size_t nb_tc_fwd_packed_bytes();
size_t nb_tc_bwd_packed_bytes();
size_t nb_tc_bwd_ws_bytes(const nb_mlp_desc& d, long long P);
// appends the dgrad (W^T) blobs in consumption order: add(src_float_off, ld, transposed, n0, k0, rows, n_lim, k_lim, folded)
void nb_tc_bwd_add_blobs(const NbParamLayout& L, const std::function<void(size_t, int, int, int, int, int, int, int, int)>& add);
