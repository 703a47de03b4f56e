// Backward / training-only kernels of the WordDiffusion noise-prediction step (reference train.py:281-294:
// `loss.backward()`, `optimizer.step()`, `ema.step_ema`).  HBM-/latency-bound work; the dense contractions of the backward
// pass run on the tensor cores (gemm_tc.cu with transposed weights for the data gradients, wgrad_tc.cu for the weight
// gradients).  All activations and activation gradients are token-major bf16, parameter gradients are fp32 in the reference
// state_dict layout and are ACCUMULATED (+=): the caller zeroes them (optimizer.zero_grad()).
#pragma once
#include "common.cuh"

namespace wd {

typedef __nv_bfloat16 bf16_t;

// ---------------- GroupNorm(+SiLU) backward (unet.py:429-431,592-596,161-162) ----------------
// y = act(xhat * gamma + beta) over the channel concatenation of <= 2 source tensors (slabs of Cs channels).
struct GroupNormBwdArgs {
  const bf16_t* x[2];       // source tensors [B, HW, x_ld]
  int x_ld[2];
  const float* partial[2];  // forward partial statistics of each source (GroupNormArgs::partial)
  int pslots[2];
  const bf16_t* dy;         // gradient of the output [B, HW, nslab*Cs]
  int dy_ld;
  const float* gamma;       // [nslab*Cs]
  const float* beta;
  float* ws;                // workspace [B][4][nslab*Cs][2] fp32: per-sample (x <= 4 row chunks) per-channel {sum dz, sum dz*xhat}
  bf16_t* dx[2];            // gradient w.r.t. each source [B, HW, dx_ld]
  int dx_ld[2];
  const bf16_t* add[2];     // optional extra term added to dx (the branch that by-passes the norm), [B, HW, add_ld]
  int add_ld[2];
  int accumulate[2];        // 1: dx += (it already holds a gradient)
  float* dgamma;            // [nslab*Cs] fp32, +=
  float* dbeta;
  int HW, Cs, cpg, pcpg;
  float eps;
  int silu;
};
cudaError_t groupnorm_bwd_launch(const GroupNormBwdArgs& a, int B, int nslab, cudaStream_t s);

// ---------------- LayerNorm backward (unet.py:314-316) ----------------
// dx = LN'(x) dy (+ add) ; dgamma += sum dy*xhat ; dbeta += sum dy
cudaError_t layernorm_bwd_launch(const bf16_t* x, const bf16_t* dy, const float* gamma, const bf16_t* add, bf16_t* dx,
                                 float* dgamma, float* dbeta, int M, int C, float eps, cudaStream_t s);

// ---------------- GEGLU (unet.py:127-129): out[m, j] = p[m, j] * gelu(p[m, H + j]), exact erf GELU ----------------
cudaError_t geglu_fwd_launch(const bf16_t* p, bf16_t* out, int M, int H, cudaStream_t s);
cudaError_t geglu_bwd_launch(const bf16_t* p, const bf16_t* dout, bf16_t* dp, int M, int H, cudaStream_t s);

// ---------------- SiLU (time-embedding MLP, unet.py:1201-1205,609-610) ----------------
cudaError_t silu_fwd_launch(const bf16_t* x, bf16_t* y, size_t n, cudaStream_t s);
cudaError_t silu_bwd_launch(const bf16_t* x, const bf16_t* dy, bf16_t* dx, size_t n, cudaStream_t s);

// ---------------- short-context cross-attention backward (L <= 16, d_head = 80; unet.py:185-279) ----------------
// q [B,Sq,q_ld], k/v [B,L,kv_ld], dout [B,Sq,do_ld] -> dq [B,Sq,dq_ld], dk/dv [B,L,dkv_ld] (every element written)
struct AttnSmallBwdArgs {
  const bf16_t* q; int q_ld;
  const bf16_t* k; const bf16_t* v; int kv_ld;
  const bf16_t* dout; int do_ld;
  bf16_t* dq; int dq_ld;
  bf16_t* dk; bf16_t* dv; int dkv_ld;
  int Sq, L, heads;
  float scale;
};
cudaError_t attn_small_bwd_launch(const AttnSmallBwdArgs& a, int B, cudaStream_t s);

// ---------------- column sums (bias gradients; per-sample sums = gradient of the timestep-embedding row bias) ----------------
// dy [groups*rows_per_group, ld] (N columns used).  total[N] += sum over all rows (fp32);
// per_group (optional, bf16 [groups, pg_ld]) = sum over the rows of each group.
cudaError_t colsum_launch(const bf16_t* dy, int ld, int N, int groups, int rows_per_group, float* total, bf16_t* per_group,
                          int pg_ld, cudaStream_t s);

// ---------------- resampling ----------------
// gradient of nearest 2x upsampling: dx[b,y,x,:] (+)= sum of the 2x2 block of dup  (unet.py:497)
cudaError_t upsample2x_bwd_launch(const bf16_t* dup, bf16_t* dx, int B, int H, int W, int C, int accumulate, cudaStream_t s);
// zero-insertion: out[b,2y,2x,:] = x[b,y,x,:], 0 elsewhere (data gradient of the stride-2 conv, unet.py:540)
cudaError_t dilate2x_launch(const bf16_t* x, bf16_t* out, int B, int H, int W, int C, cudaStream_t s);

// ---------------- layout / dtype glue ----------------
// fp32 NCHW [B,4,H,W] -> bf16 token-major [B*H*W, 64] columns 0..3 (4..7 zeroed, 8..63 untouched)
cudaError_t nchw4_to_tok64_launch(const float* g, bf16_t* out, int B, int HW, cudaStream_t s);
cudaError_t f32_to_bf16_launch(const float* x, bf16_t* out, size_t n, cudaStream_t s);
// y (+)= a + b elementwise (b may be null)
cudaError_t add_bf16_launch(const bf16_t* a, const bf16_t* b, bf16_t* y, size_t n, int accumulate, cudaStream_t s);
// table[idx[r], :] += rows[r, :]  (nn.Embedding backward; idx int64 or int32)
cudaError_t scatter_add_rows_launch(const bf16_t* rows, int ld, const void* idx, int idx_i64, float* table, int nrows, int D,
                                    int table_rows, cudaStream_t s);
// conv_in weight gradient: dW[n, j] += g[n, j] + g[n, 36 + j]  (g = wgrad against the hi|lo im2col operand, [N,128] fp32)
cudaError_t conv_in_wgrad_fold_launch(const float* g, float* dW, int N, cudaStream_t s);

// ---------------- Word_Attention backward (fp32 context encoder, unet.py:815-836) ----------------
// q,k,v fp32 [B,L,D]; dctx bf16 [B, Ltot, D] rows row_off..row_off+L-1; d_qkv bf16 [B*L, 3D] = dq | dk | dv
cudaError_t word_attn_bwd_launch(const float* q, const float* k, const float* v, const bf16_t* dctx, bf16_t* d_qkv, int B,
                                 int L, int D, int Ltot, int row_off, cudaStream_t s);

// ---------------- transposed weight packs for the data-gradient GEMMs ----------------
// nn.Linear / 1x1 conv weight fp32 [N, K] -> bf16 dst[k_row_off + k][n_off + n]   (row stride ldn)
cudaError_t repack_linear_T_launch(const float* w, bf16_t* dst, int N, int K, int ldn, int n_off, int k_row_off, cudaStream_t s);
// conv3x3 weight fp32 [Cout, Cin, 3, 3] -> bf16 dst[c][(8 - tap) * cout_pad + n]   (row stride 9 * cout_pad)
// Every weight pack of the trainer in ONE launch (wd_trainer_sync_weights ran ~200 small launches per step: 1.0 ms of the 16 ms
// step at batch 224, a sixth of the 28-latent step): `jobs` is a device table sorted by `start` (first element index of the job).
// kind: 0 linear [N,K] -> dst[(n + off1) ld + off0 + k] | 1 linear transposed -> dst[(off1 + k) ld + off0 + n] |
//       2 conv3x3 [Cout=N, Cin=K, 3, 3] -> dst[n ld + off0 + tap K + c] (off1: 0 bf16, 1 fp16 bits, 2 bf16 hi / lo rows n, n + 4) |
//       3 conv3x3 transposed, flipped taps -> dst[c 9 ld + (8 - tap) ld + n] | 4 conv_in [Cout=N, 4, 3, 3] -> bf16 [N, 128] hi|hi|lo|0
struct PackDesc {
  const float* src;
  void* dst;
  long long start;
  int kind, N, K, ld, off0, off1;
};
cudaError_t repack_multi_launch(const PackDesc* jobs_dev, int njobs, long long total, cudaStream_t s);
// the transposed jobs (kind 1, 3) as 32 x 32 shared-memory tiles; `start` = first tile index of the job
cudaError_t repack_multi_T_launch(const PackDesc* jobs_dev, int njobs, long long total_tiles, cudaStream_t s);
cudaError_t repack_conv3x3_T_launch(const float* w, bf16_t* dst, int Cout, int Cin, int cout_pad, cudaStream_t s);

// ---------------- AdamW + EMA (train.py:405 `optim.AdamW(lr=1e-4)`, train.py:140-170 `EMA(0.995)`) ----------------
// torch.optim.AdamW semantics (decoupled weight decay, bias correction, eps outside the sqrt of the corrected v).
// ema_mode 0: none, 1: ema = p (warm-up copy, EMA.reset_parameters), 2: ema = beta*ema + (1-beta)*p
cudaError_t adamw_ema_launch(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, float ema_beta, int ema_mode,
                             float grad_scale, cudaStream_t s);

}  // namespace wd
