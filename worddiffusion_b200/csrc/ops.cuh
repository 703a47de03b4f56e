// Non-GEMM kernels of the WordDiffusion hot path (HBM / latency bound work).  See ops.cu.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"  // StepMode

namespace wd {

// ---------------- GroupNorm (+SiLU), NHWC bf16 -> NHWC bf16 (reference unet.py:429-431,161-162) ----------------
// Per-tensor partial statistics: partial[sample][C/pcpg groups][slots][2] fp32 ({sum, sum of squares}).
struct GroupNormStatsArgs {
  const __nv_bfloat16* x;  // [B, HW, ld]
  int ld;
  float* partial;
  int HW, C;
  int pcpg;    // channels per partial group
  int pslots;  // pixel chunks per sample (groupnorm_stats_slots(HW))
  int x_f16;   // 1: x is fp16 (residual-stream tensor), 0: bf16
};
int groupnorm_stats_slots(int HW);
cudaError_t groupnorm_stats_launch(const GroupNormStatsArgs& a, int B, cudaStream_t s);

struct GroupNormArgs {
  const __nv_bfloat16* x[2];  // per channel slab: source tensor
  int x_ld[2];                // pixel stride (elements) of each source
  const float* partial[2];    // partial statistics of each source tensor
  int pslots[2];              // slots per (sample, partial group) of each source
  __nv_bfloat16* out;         // [B, HW, out_ld]; slab s writes channels [s*Cs, (s+1)*Cs)
  int out_ld;
  const float* gamma;  // [nslab*Cs]
  const float* beta;
  int HW;
  int Cs;    // channels per slab (= channels of each source; multiple of 8, whole groups)
  int cpg;   // channels per GroupNorm group
  int pcpg;  // channels per partial-statistics group (cpg % pcpg == 0)
  float eps;
  int silu;
  int nchunk;  // pixel chunks per sample of the apply grid
  int x_f16[2];  // per source: 1 = fp16 input (residual-stream tensors), 0 = bf16.  The output is always bf16.
  int reverse;   // bulk kernel: walk the items from the end of the tensor (set by groupnorm_launch)
};
// normalise (+SiLU) from the partial statistics
int groupnorm_apply_chunks(int HW);  // pixel chunks per sample of the apply grid (GroupNormArgs::nchunk)
cudaError_t groupnorm_launch(const GroupNormArgs& a, int B, int nslab, cudaStream_t s);

cudaError_t noise_images_launch(const float* x, const long long* t, const float* alpha_hat, int T, const float* eps_in,
                                unsigned long long seed, unsigned long long elem_offset, uint32_t stream_id, float* x_t, float* eps_out,
                                size_t n, int per, int* bad, cudaStream_t s);
size_t mse_grad_workspace_bytes(size_t n);
cudaError_t mse_grad_launch(const float* pred, const float* target, float* d, float* loss, void* workspace, size_t n, cudaStream_t s);
cudaError_t lerp_launch(const float* a, const float* b, float w, float* out, size_t n, cudaStream_t s);
cudaError_t label_mix_launch(float* table, int D, int row, int s1, int s2, float mix, cudaStream_t s);

// ---------------- LayerNorm over the channel dim, bf16 -> bf16 (unet.py:314-316) ----------------
cudaError_t layernorm_launch(const __nv_bfloat16* x, __nv_bfloat16* out, const float* gamma, const float* beta, int M,
                             int C, float eps, int x_f16, cudaStream_t s);

// ---------------- attention with a short key/value sequence (char context, L <= 16), unet.py:185-279 ----------------
struct AttnSmallArgs {
  const __nv_bfloat16* q;  // [B, Sq, q_ld]
  int q_ld;
  const __nv_bfloat16* k;  // [B, L, kv_ld]
  const __nv_bfloat16* v;
  int kv_ld;
  __nv_bfloat16* out;  // [B, Sq, out_ld]
  int out_ld;
  float* probs;  // optional [B, heads, Sq, L] fp32 (attention maps), may be null
  int Sq, L, heads;
  float scale;
};
cudaError_t attn_small_launch(const AttnSmallArgs& a, int B, cudaStream_t s);

// ---------------- flash-style attention, general Skv (self-attention 256/64 tokens, PHOSC context 779) ----------------
struct AttnFlashArgs {
  const __nv_bfloat16* q;  // [B, Sq, q_ld]
  int q_ld;
  const __nv_bfloat16* k;  // [B, Skv, kv_ld]
  const __nv_bfloat16* v;
  int kv_ld;
  __nv_bfloat16* out;
  int out_ld;
  int Sq, Skv, heads;
  float scale;
};
cudaError_t attn_flash_launch(const AttnFlashArgs& a, int B, cudaStream_t s);
// tcgen05 attention (attn_tc.cu), Skv > 16: returns false when it does not take the launch (the mma.sync kernel runs instead)
bool attn_tc_try_launch(const AttnFlashArgs& a, int B, cudaStream_t s, cudaError_t* err);

// ---------------- PHOSC tokenizer: words [B, max_len] zero-padded bytes -> int32 [B, 769]; *bad_flag |= 1 on a non-letter ----------------
cudaError_t phosc_tokenize_launch(const unsigned char* words, int B, int max_len, int* out, int* bad_flag, cudaStream_t s);

// emb_act = SiLU(table[t] + label_emb[y]) -> bf16 [B, dim]; table fp32 [timesteps, dim] (timestep_embed_launch with t_scalar < 0
// embeds t = row index, which is how the table's sinusoid rows are made)
cudaError_t emb_from_table_launch(const float* table, long long t, const StepParams* sp, const float* label_emb, const long long* y,
                                  __nv_bfloat16* out, int B, int dim, cudaStream_t s);
cudaError_t set_step_params_launch(StepParams* dst, const StepParams& v, cudaStream_t s);

// ---------------- sampler update with a given (possibly stale) predicted noise; same arithmetic as the fused epilogue ----------------
cudaError_t sampler_update_launch(float* x, const float* eps, const float* noise, int use_philox, unsigned long long seed,
                                  unsigned long long elem_offset, int step_index, float4 coef, int mode, size_t n, cudaStream_t s);

// ---------------- sinusoidal timestep embedding (unet.py:96-116) ----------------
// t_dev: per-sample int64 timesteps, or null -> every row uses t_scalar.  out bf16 [B, dim]
cudaError_t timestep_embed_launch(const long long* t_dev, long long t_scalar, __nv_bfloat16* out, int B, int dim,
                                  cudaStream_t s);

// ---------------- conv_in (unet.py:1251): im2col of the fp32 NCHW latent into the bf16 GEMM operand [B*H*W, 128] ----------------
// columns j = c*9+ky*3+kx: x_hi, 36 + j: x_lo (x - x_hi), 72 + j: x_hi again (pairs with w_lo), 108..127: zero
cudaError_t conv_in_im2col_launch(const float* x, __nv_bfloat16* out, int B, int H, int W, cudaStream_t s);

// ---------------- output head in one kernel: GroupNorm32 + SiLU + conv3x3 C -> 4 + sampler update ----------------
// (unet.py:1454-1458,1815 then train.py:229-236).  One CTA per sample: the raw fp16 tensor is read ONCE, normalised into a
// shared-memory image (bf16, one padded row per pixel + a zero row that plays the conv's padding), the 3 x 3 x C x 8 contraction
// (4 output channels as bf16 hi + lo weight rows) runs on mma.sync from that image, and the epilogue applies the DDPM / DDIM
// update exactly like the tcgen05 output conv's sampler epilogue.  Replaces a GroupNorm launch (write + re-read of the
// normalised tensor) and an N = 16 tcgen05 launch that fetched every activation nine times through L2.
struct OutHeadArgs {
  const __half* h;        // [B, HW, C] fp16 (residual stream)
  const float* partial;   // GroupNorm partial statistics of h: [B][32][pslots][2]
  int pslots;
  const float* gamma;
  const float* beta;
  float gn_eps;
  const __nv_bfloat16* w;  // [>= 8][9 C]: rows 0..3 = hi part of out.2.weight (k = tap C + c), rows 4..7 = lo part
  const float* bias;       // [4]
  int B, H, W, C;
  // sampler (same meaning as GemmArgs)
  float* eps_out;
  float* x;
  const float* noise;
  int use_philox;
  unsigned long long seed;
  unsigned long long sample_offset;
  int step_index;
  float4 coef;
  int mode;
  const StepParams* sp;
};
bool out_head_supported(int H, int W, int C);  // shared-memory image fits, whole 16-pixel tiles, whole groups of C / 32 channels
cudaError_t out_head_launch(const OutHeadArgs& a, cudaStream_t s);

// ---------------- nearest 2x upsample NHWC bf16 (unet.py:497) ----------------
cudaError_t upsample2x_launch(const __nv_bfloat16* x, __nv_bfloat16* out, int B, int H, int W, int C, cudaStream_t s);

// ---------------- sub-pixel weights of "nearest 2x upsample, then conv3x3" (GemmArgs::up_phase) ----------------
// w [Cout, Cin, 3, 3] fp32 -> dst [4 phases][Cout][4 taps][Cin] (bf16, or fp16 bit patterns when as_f16): phase (a, b), tap
// (ty, tx) holds the sum of the 3x3 taps (ky, kx) that read input pixel (ty - 1 + a, tx - 1 + b): rows a = 0: {0}, {1, 2};
// a = 1: {0, 1}, {2} (same for the columns), summed in fp32 and rounded once.
cudaError_t upconv_phase_fold_launch(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int as_f16, cudaStream_t s);

// ---------------- weight repacking (fp32 state_dict tensors -> packed bf16 / fp32 layouts) ----------------
// conv3x3 weight [Cout, Cin, 3, 3] -> dst[n, k_off + tap*Cin + c]   (row stride ldk); as_f16: store fp16 bit patterns
// (weight columns that multiply an fp16 A source, see GemmArgs::a_f16)
cudaError_t repack_conv3x3_launch(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int ldk, int k_off, int as_f16,
                                  cudaStream_t s);
// linear / 1x1 weight [N, K] -> dst[perm(n) + n_off, k_off + k]; geglu_bn > 0 applies the value/gate tile permutation
cudaError_t repack_linear_launch(const float* w, __nv_bfloat16* dst, int N, int K, int ldk, int k_off, int n_off,
                                 int geglu_bn, int as_f16, cudaStream_t s);
// vector [N] -> dst[perm(n) + n_off] (accumulate: dst += src)
cudaError_t repack_vec_launch(const float* v, float* dst, int N, int n_off, int geglu_bn, int accumulate,
                              cudaStream_t s);
// LayerNorm (gamma, beta over K) folded into nn.Linear(K, N): dst = fp16(gamma (.) W) at rows perm(n) + n_off,
// s_out[n'] = row sums of the rounded weights, b_out[n'] = W beta + bias
cudaError_t fold_ln_linear_launch(const float* w, const float* gamma, const float* beta, const float* bias, __nv_bfloat16* dst,
                                  float* s_out, float* b_out, int N, int K, int ldk, int n_off, int geglu_bn, cudaStream_t s);
// conv_in weight [Cout,4,3,3] -> bf16 [Cout, 128]: w_hi | w_hi | w_lo | 0
cudaError_t repack_conv_in_launch(const float* w, __nv_bfloat16* dst, int Cout, int Cin, cudaStream_t s);

// ---------------- fp32 context encoder (unet.py:815-882) ----------------
// tokens [B, L] (int64 or int32) -> emb[B, L, D] = E[token] (+ pe[l] when add_pe)
cudaError_t embed_tokens_launch(const void* tokens, int tokens_are_i64, const float* E, int vocab, const float* pe,
                                int add_pe, float* out, int B, int L, int D, cudaStream_t s);
// out[m, n] = sum_k x[m,k] W[n,k] + b[n]   (fp32 SIMT)
cudaError_t linear_f32_launch(const float* x, const float* W, const float* b, float* out, int M, int N, int K,
                              cudaStream_t s);
// unscaled single-head softmax(Q K^T) V, fp32; writes bf16 ctx rows [b, row_off + l, :] of a [B, Ltot, D] tensor
// position-free segments (no positional encoding): histogram form, see ops.cu
cudaError_t word_attn_hist_launch(const void* tokens, int tokens_are_i64, const float* G, const float* TV, int vocab,
                                  __nv_bfloat16* ctx_out, int B, int L, int D, int Ltot, int row_off, cudaStream_t s);
cudaError_t word_attn_gram_launch(const float* TQ, const float* TK, float* G, int vocab, int D, cudaStream_t s);
cudaError_t word_attn_launch(const float* q, const float* k, const float* v, __nv_bfloat16* ctx_out, float* ctx_out_f32,
                             int B, int L, int D, int Ltot, int row_off, cudaStream_t s);

}  // namespace wd
