// One kernel for the whole transformer block of unet.UNetModel's SpatialTransformer (reference unet.py:337-345, 381-412):
//
//     x  = proj_in(g)                                  g = GroupNorm(x_in), 1x1 conv
//     x += attn1.to_out(softmax(to_q(LN2 x) K1^T) V1)   cross-attention over the <= 16 character tokens
//     x += attn2.to_out(softmax(to_q(LN2 x) K2^T) V2)   (unet.py applies norm2 twice; norm1 is a dead parameter)
//     x += ff.net.2(GEGLU(ff.net.0.proj(LN3 x)))
//     out = proj_out(x) + x_in
//
// Every op is row-local, so a CTA carries a 128-token tile through all of it: the fp32 residual stream never leaves TMEM
// (each residual branch is a tcgen05.mma that ACCUMULATES into the same 320 columns), the 16-bit operand copies live in one
// 80 KB shared-memory buffer, and only weights stream in (TMA ring) -- one read of g / x_in and one write of `out` per token
// instead of nine HBM round trips through seven GEMM launches.
//
// The two cross-attentions are folded algebraically: with K, V fixed per trajectory,
//     scores[row, (h, j)] = LN(x)[row, :] . (Wq_h^T K_h[j, :])            -> "S" GEMM, N = 64 (4 heads x 16 key slots), K = 320
//     x += P[row, (h, j)] . (V_h[j, :] Wout_h^T)                           -> "PN" GEMM, N = 320, K = 64
// replace to_q (320 x 320), QK^T, PV and to_out (320 x 320): 8x fewer FLOPs for the sub-block, and q never exists.  The
// per-sample operands (M = Wq^T K^T and N = V Wout^T) come from ONE GEMM per trajectory: ctx [B L, 320] x W_fold^T, where
// W_fold holds Wk_h^T Wq_h and Wv_h^T Wout_h^T, built once per weight load (tblock_fold_weights_launch).
#pragma once
#include "common.cuh"

namespace wd {

constexpr int TB_M = 128;          // tokens per tile
constexpr int TB_C = 320;          // channels (4 heads x 80)
constexpr int TB_HEADS = 4;
constexpr int TB_DH = 80;
constexpr int TB_KEYS = 16;        // key slots per head in the S / PN GEMMs (context length <= 16)
constexpr int TB_HID = 1280;       // GEGLU hidden width
constexpr int TB_CHUNK = 64;       // hidden columns per feed-forward chunk
constexpr int TB_FOLD_N = 2 * TB_HEADS * TB_C;  // rows of W_fold per attention: [M part: (h, k)] ++ [N part: (h, n)] = 2560

struct TBlockArgs {
  int M;    // rows = batch * HW, a multiple of 128
  int HW;   // tokens per sample: a multiple of 128 (a tile lies inside one sample), or 64 (two samples per tile)
  int L;    // context keys, 1 .. 16
  const float* cb;     // [4][320] cumulative biases of the residual stream after proj_in / attn1 / attn2 / ff
  const float* b_ff;   // [2560] GEGLU bias with LayerNorm-3 folded in, chunk-interleaved (64 values ++ 64 gates per chunk)
  const float* cvec1;  // additive score constants of attn1 (beta of the folded LayerNorm, already x scale log2 e): row b L + j holds
  const float* cvec2;  //   the 4 heads at cvec[(b L + j) * cvec_ld + h]
  int cvec_ld;
  const float* b_po;   // [320] proj_out bias
  const __half* x_in;  // [M, x_in_ld] fp16: the SpatialTransformer input (residual of proj_out)
  int x_in_ld;
  float* gn_partial;   // GroupNorm partials of `out`: [sample][32 groups][HW / 32][2], or null
  // GroupNorm of the block's input inside the kernel (SpatialTransformer.norm, unet.py:388: GroupNorm32 without SiLU, eps 1e-6): when
  // gn_in_partial != null, mapG maps x_in itself (fp16) and the epilogue warps normalise the tile in shared memory before proj_in
  const float* gn_in_partial;  // partial statistics of x_in: [sample][32 groups][gn_in_slots][2]
  int gn_in_slots;
  const float* gn_gamma;       // [320]
  const float* gn_beta;
  float gn_eps;
  float ln_eps;
  int pair;            // 1: CTA-pair kernel (256-token tiles, HW % 256 == 0; weight maps encoded with half-unit boxes)
  int mid;             // 1: "middle" form for SpatialTransformers whose in/out channels differ from 320 (proj_in / proj_out stay
                       // separate GEMMs): mapG = the fp16 residual stream x [M, 320], mapWpi = a 320 x 320 fp16 identity, cb[0] = 0,
                       // stage = 4 (the tile leaves as the raw fp16 stream after the feed-forward)
  int trace;           // set by tblock_launch from env WD_TBLOCK_TRACE: CTA 0 records clock64 phase stamps (tools/tblock_trace.py)
  int stage;           // 0: full block.  Debug (operator test): 1..4 -> `out` receives the normalised operand copy after
                       // proj_in / attn1 / attn2 (LayerNorm without gamma / beta) or the raw residual stream after ff (4)
};

struct TBlockLaunch {
  CUtensorMap mapG;     // g     bf16 [M, 320]            box {64, 128}
  // weight boxes below are for the single-CTA kernel; the pair kernel (args.pair) takes half of each: {64, 80} / W1 {64, 64}
  CUtensorMap mapWpi;   // proj_in  bf16 [320, 320]       box {64, 160}
  CUtensorMap mapF[4];  // per-sample fold operands (fp16): attn1 M, attn1 N, attn2 M, attn2 N; 3-D (1280, L, batch), box {64, 16, 1}
  CUtensorMap mapW1;    // ff.net.0.proj folded, fp16 [2560, 320], box {64, 128}
  CUtensorMap mapW2;    // ff.net.2 bf16 [320, 1280]      box {64, 160}
  CUtensorMap mapWpo;   // proj_out fp16 [320, 320]       box {64, 160}
  CUtensorMap mapOut;   // out   fp16 [M, 320]            box {64, 128}
  TBlockArgs args;
};

bool tblock_enabled();  // env WD_TBLOCK (default on)
bool tblock_gn_fused();        // env WD_TBLOCK_GN (default on): SpatialTransformer.norm inside the kernel
bool tblock_use_pair(int HW);  // env WD_TBLOCK_PAIR (default on) and HW % 256 == 0
cudaError_t tblock_launch(const TBlockLaunch& L, int num_sms, cudaStream_t stream);

// Weight-load time: W_fold rows of one attention (bf16 [2560, 320]) and the score-constant vectors u (fp32 [4][320]).
//   wq [320, 320] (to_q.weight), wk / wv [320, ctx_dim = 320] (to_k / to_v), wo [320, 320] (to_out.0.weight); gamma / beta: norm2
cudaError_t tblock_fold_weights_launch(const float* wq, const float* wk, const float* wv, const float* wo, const float* gamma,
                                       const float* beta, __nv_bfloat16* w_fold, float* u, cudaStream_t s);
// Per trajectory: cvec[row][hh] = sum_c ctx[row][c] u[hh][c] for all `heads` = (attentions x 4) pooled u vectors
cudaError_t tblock_cvec_launch(const __nv_bfloat16* ctx, const float* u, float* cvec, int rows, int heads, cudaStream_t s);

}  // namespace wd
