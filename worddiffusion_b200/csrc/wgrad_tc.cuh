// Weight-gradient contraction on the 5th-gen tensor cores (training step, reference train.py:281-294 `loss.backward()`):
//
//     dW[n, tap, c] += sum over tokens m of  dY[m, n] * X[shift_tap(m), c]
//
// i.e. dW = dY^T . im2col(X) for the 3x3 convolutions (unet.py:595,621,488,540,1251,1457) and dW = dY^T . X for the 1x1
// convolutions / nn.Linear layers (unet.py:364,375,632,175-183,125,145,611,1202-1204).  The reduction runs over TOKENS, the
// dimension along which both operands are strided in memory, so both shared-memory operands are "MN-major": the very same
// [64 channel x 64 token] SWIZZLE_128B TMA boxes the forward kernel loads (4-D shifted boxes for a filter tap, zero fill =
// conv padding) are handed to tcgen05.mma with the transpose (a_major = b_major = MN) bits set -- no transposed copy of an
// activation is ever made.
//
// Work item (one CTA) = (token split, filter tap, 128-channel slice of X, BN-column group of dY):
//   UMMA M = 128 (X channels -> TMEM lanes), N = BN (dY channels -> TMEM columns; 320 = one N=192 + one N=128 MMA), K = 16 tokens.
//   fp32 accumulators stay in TMEM for the whole token range of the split; the epilogue adds them to the fp32 gradient tensor
//   (reference state_dict layout, arbitrary strides) with red.global.add.f32.
#pragma once
#include "common.cuh"

namespace wd {

constexpr int WG_BLOCK_C = 128;   // X channels per work item (UMMA M)
constexpr int WG_BLOCK_TOK = 64;  // tokens per pipeline stage (4 UMMA K steps)
constexpr int WG_MAX_GROUPS = 8;

struct WgradArgs {
  int M;       // tokens (rows of dY)
  int Cin;     // channels of X (>= 128, multiple of 64)
  int taps;    // 1 or 9
  int conv;    // 1: X is addressed through the 4-D (c, w, h, n) map
  int HWout, Wout, stride;  // conv geometry of the OUTPUT grid (tokens of dY); stride of the forward conv
  int n_groups;             // dY column groups of BN
  int splits;               // token splits (work items along the reduction)
  int n_valid;              // valid dY columns per group (BN, or 4 for the output conv)
  float* dst[WG_MAX_GROUPS];  // gradient tensor of each group: element (n, c, tap) at n*sN + c*sC + tap*sT
  long long sN, sC, sT;
};

struct WgradLaunch {
  CUtensorMap mapX;   // 2-D [M, Cin] or 4-D (c,w,h,n); box = 64 channels x 64 tokens, SWIZZLE_128B
  CUtensorMap mapDY;  // 2-D [M, n_groups*BN]; box = 64 x 64
  WgradArgs args;
  int bn;  // 320 or 64
};

int wgrad_pick_splits(int M, int items_base);
cudaError_t wgrad_tc_launch(const WgradLaunch& L, cudaStream_t stream);

}  // namespace wd
