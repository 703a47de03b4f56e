/*
 * wd_b200 -- C ABI of the B200-native WordDiffusion denoising hot path.
 *
 * The reference (aniketntnu/WordDiffusion) has no FFI / plugin layer: the seam of its hot path is the
 * Python nn.Module handed to the sampling / training loops (reference train.py:403,424;
 * regenerateFromtrain2.py:1172,1291).  This header is the boundary a binding for that seam talks to:
 * plain pointers and sizes, no torch types, every function returns 0 on success and a negative code on
 * failure (wd_last_error() gives the message), nothing throws.  All pointers are DEVICE pointers unless
 * a parameter says "host".  `stream` is a cudaStream_t passed as void*.  Calls on one engine must be
 * serialised by the caller; different engines may run concurrently on different streams.
 *
 * Each entry point cites the reference interface it replaces (file:line under /root/reference).
 */
#ifndef WD_B200_H
#define WD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WD_OK 0
#define WD_IGNORED 1 /* load_param: a parameter the reference forward never reads (attnc.*, *.to_kv, res.*, ...) */
#define WD_ERR_INVALID (-1)
#define WD_ERR_UNSUPPORTED (-2)
#define WD_ERR_CUDA (-3)
#define WD_ERR_STATE (-4)

#define WD_VARIANT_UNET 0  /* unet.py UNetModel: both attentions of a block are cross-attention (unet.py:337-345) */
#define WD_VARIANT_PHOSC 1 /* unetPhosc.py / unetPhosc2.py UNetModelPhosc: self-attn + cross-attn (unetPhosc.py:241-246) */

#define WD_STEP_EPS_ONLY 0
#define WD_STEP_DDPM 1
#define WD_STEP_DDIM 2

typedef struct wd_engine wd_engine;

/* Mirrors the constructor arguments of UNetModel / UNetModelPhosc (unet.py:1126-1156, unetPhosc.py:781-811). */
typedef struct wd_config {
  int variant;
  int in_channels;
  int model_channels;
  int out_channels;
  int num_res_blocks;
  int n_channel_mult;
  int channel_mult[8];
  int n_attention_resolutions;
  int attention_resolutions[8];
  int num_heads;         /* -1 if num_head_channels is used */
  int num_head_channels; /* -1 if num_heads is used */
  int transformer_depth;
  int context_dim;
  int vocab_size;
  int num_classes; /* 0: no label embedding */
  int max_seq_len;
  int latent_h; /* 8  */
  int latent_w; /* 32 */
  int add_label_emb; /* unet.py:1578-1581: label_emb is skipped when args.imgConditioned == 1 */
  int phosc_len;     /* number of PHOSC tokens concatenated to the context (769), 0 = none (unetPhosc.py:1120-1130) */
} wd_config;

const char* wd_last_error(void);
int wd_version(void);

/* ---- engine life cycle -------------------------------------------------------------------------------- */
int wd_engine_create(const wd_config* cfg, wd_engine** out);
void wd_engine_destroy(wd_engine* e);

/* state_dict entry -> packed device weights.  `name` is the reference state_dict key
 * (e.g. "input_blocks.1.0.in_layers.2.weight"); `src` is the fp32 device tensor, contiguous. */
int wd_engine_load_param(wd_engine* e, const char* name, const float* src, const int64_t* shape, int ndim, void* stream);
/* CharacterEncoder.positional_encoding (unet.py:876-882) is not in the state_dict: [max_seq_len, context_dim] fp32 */
int wd_engine_set_pos_encoding(wd_engine* e, const float* pe, void* stream);
/* call once after all load_param calls (sums fused biases); returns the number of parameters still missing */
int wd_engine_finalize_params(wd_engine* e, void* stream);

/* allocate activations for up to `batch` latents and build the launch plan */
int wd_engine_reserve(wd_engine* e, int batch);
size_t wd_engine_workspace_bytes(const wd_engine* e);
size_t wd_engine_weight_bytes(const wd_engine* e);
/* kernels launched by the most recent wd_unet_eval / wd_sampler_step / wd_encode_context call */
int wd_engine_last_launch_count(const wd_engine* e);

/* ---- per-op device timing (bench.py's roofline leg) -------------------------------------------------------
 * While enabled, every wd_unet_eval / wd_sampler_step records CUDA events on `stream` around each kernel launch of the
 * step (at most 256 steps are kept).  wd_engine_profile_read synchronises the device and returns, per launch of the
 * step plan: its kernel class (0 timestep-embed, 1 tcgen05 GEMM/conv, 2 GroupNorm, 3 LayerNorm, 4 short-context
 * attention, 5 flash attention, 6 conv_in, 7 GroupNorm statistics, 8 upsample; the output conv + sampler update is a
 * class-1 launch), its algorithmic FLOPs and bytes,
 * and the summed milliseconds over the recorded steps.  Returns the number of launches per step (or < 0). */
int wd_engine_set_profiling(wd_engine* e, int enable);
int wd_engine_profile_read(wd_engine* e, int cap, int* kinds, double* flops, double* bytes, float* ms_sum, int* n_steps);

/* ---- the hot path ------------------------------------------------------------------------------------ */
/* Time-invariant conditioning: CharacterEncoder (+PHOSC tokens) and the K/V projections of every
 * cross-attention (unet.py:1626-1636,839-882,180-183 ; unetPhosc.py:1117-1130).
 * ctx_tokens: int64 [B, L]; phosc: int32 [B, phosc_len] or NULL. */
int wd_encode_context(wd_engine* e, int batch, const int64_t* ctx_tokens, int L, const int32_t* phosc, void* stream);

/* One UNet evaluation: eps = UNetModel(x, timesteps, context, y) (unet.py:1499-1836 / unetPhosc.py:1068-1159),
 * using the context encoded by the last wd_encode_context.  x, eps_out: fp32 NCHW [B,4,H,W].
 * timesteps: int64 [B] device pointer, or NULL to use t_scalar for every row. y: int64 [B] (may be NULL if
 * num_classes == 0 or add_label_emb == 0). */
int wd_unet_eval(wd_engine* e, int batch, const float* x, const int64_t* timesteps, int64_t t_scalar, const int64_t* y,
                 float* eps_out, void* stream);

/* UNet evaluation fused with the sampler update of Diffusion.sampling (train.py:221-236): x is updated in place.
 * mode WD_STEP_DDPM: coef = {1/sqrt(alpha_t), (1-alpha_t)/sqrt(1-alpha_hat_t), sqrt(beta_t), 0}
 * mode WD_STEP_DDIM: coef = {1/sqrt(ah_t), sqrt(1-ah_t), sqrt(ah_prev), sqrt(1-ah_prev)}   (eta = 0)
 * noise: fp32 NCHW [B,4,H,W] or NULL; when NULL and use_philox != 0 the kernel draws N(0,1) from Philox4x32-10 keyed by
 * (seed, step_index, sample_offset + sample, element).  eps_out may be NULL. */
int wd_sampler_step(wd_engine* e, int batch, float* x, int64_t t_scalar, const int64_t* y, int mode, const float* coef4_host,
                    const float* noise, int use_philox, uint64_t seed, uint64_t sample_offset, int step_index,
                    float* eps_out, void* stream);

/* PHOSC labels on the device (reference ResPhoSCNetZSL/modules/utils/phos_generator.py:59-78 + phoc_generator.py:17-90, 'eng'
 * alphabet; the vector the reference's dataset / generator builds on the host and passes as `phoscLabels`, unetPhosc.py:1068).
 * words: device bytes [batch, max_len], zero padded, spaces / underscores already removed; out: int32 [batch, 769] (165 PHOS
 * counts ++ 604 PHOC bits); *bad_flag (device int, zeroed by the caller) becomes non-zero if a word holds a character outside
 * a-zA-Z, for which the reference raises KeyError. */
int wd_phosc_tokenize(const unsigned char* words, int batch, int max_len, int32_t* out, int32_t* bad_flag, void* stream);

/* The sampler update alone, with a predicted noise the caller kept: the reference's production generator evaluates the UNet only at
 * some steps and re-uses the last eps in between (regenerateFromtrain2.py:536,615-618).  Same fp32 arithmetic and op order as the
 * fused epilogue of wd_sampler_step.  x, eps, noise: fp32 [batch, elems_per_latent] device tensors; coef4_host / mode / noise /
 * Philox arguments as in wd_sampler_step. */
int wd_sampler_update(float* x, const float* eps, int batch, int elems_per_latent, int mode, const float* coef4_host,
                      const float* noise, int use_philox, uint64_t seed, uint64_t sample_offset, int step_index, void* stream);

/* ---- single operators (used by the parity tests; same kernels as the engine) -------------------------- */
/* GroupNorm32(+SiLU) (unet.py:429-431,592-596): x,out bf16 NHWC [B,HW,C] */
int wd_op_groupnorm(const void* x_bf16, void* out_bf16, const float* gamma, const float* beta, int B, int HW, int C,
                    int groups, float eps, int silu, void* stream);
/* nn.LayerNorm(C) (unet.py:314-316): bf16 [M,C] */
int wd_op_layernorm(const void* x_bf16, void* out_bf16, const float* gamma, const float* beta, int M, int C, float eps,
                    void* stream);
/* out[M,N] = act(A[M,K] W[N,K]^T + bias + residual) on tcgen05; A,W,residual,out bf16 (out fp32 if out_f32).
 * geglu: W rows are nn.Linear(K, N) of GEGLU.proj in packed (tile-permuted) order, out is [M, N/2]. */
int wd_op_gemm(const void* a_bf16, const void* w_bf16, const float* bias, const void* residual_bf16, void* out, int M,
               int N, int K, int act_silu, int geglu, int out_f32, void* stream);
/* the residual-stream flavour of the transformer blocks (unet.py:337-345: x = attn(norm(x)) + x): A, W bf16; residual and out
 * fp16 [M,N].  For K <= 320 the residual rides through the operand ring as extra K blocks against an identity tile. */
int wd_op_gemm_f16(const void* a_bf16, const void* w_bf16, const float* bias, const void* residual_f16, void* out_f16, int M, int N,
                   int K, void* stream);
/* 3x3 conv, pad 1, stride 1|2, NHWC bf16, weights pre-packed [Cout, 9*Cin] by wd_op_pack_conv3x3 */
int wd_op_conv3x3(const void* x_bf16, const void* w_packed_bf16, const float* bias, const float* rowbias, int rb_ld,
                  const void* residual_bf16, void* out_bf16, int B, int H, int W, int Cin, int Cout, int stride, void* stream);
/* conv3x3 (stride 1) + bias + per-sample row bias -> GroupNorm32 -> SiLU, the front half of ResBlock._forward (unet.py:657-667
 * then :592-594), with the normalisation applied by the conv kernel's own epilogue (the raw conv output never reaches HBM).
 * Cout == 320, H*W divides 256 and is a multiple of 32, K = 9*Cin/64 blocks long enough for the CTA-pair kernel (else
 * WD_ERR_UNSUPPORTED).  stats_ws: fp32 [B][32][H*W/32][2] scratch.  out: bf16 NHWC. */
int wd_op_conv3x3_gn_silu(const void* x_bf16, const void* w_packed_bf16, const float* bias, const float* rowbias, int rb_ld,
                          const float* gamma, const float* beta, float eps, void* out_bf16, float* stats_ws, int B, int H, int W,
                          int Cin, int Cout, void* stream);
int wd_op_pack_conv3x3(const float* w_oihw, void* dst_bf16, int Cout, int Cin, void* stream);
/* fp32 [N,K] -> bf16 [N,K] (geglu_perm != 0 applies the GEGLU tile permutation used by wd_op_gemm) */
int wd_op_pack_linear(const float* w, void* dst_bf16, int N, int K, int geglu_perm, void* stream);
int wd_op_pack_vec_geglu(const float* v, float* dst, int N, void* stream);
/* softmax(q k^T scale) v with a short key sequence L <= 16; q [B,Sq,C], k,v [B,L,C], out [B,Sq,C] bf16; C = heads*80 */
int wd_op_attention_small(const void* q, const void* k, const void* v, void* out, float* probs, int B, int Sq, int L,
                          int heads, float scale, void* stream);
/* CrossAttention.forward (unet.py:185-207) for a short context with the to_q Linear fused in: out = softmax((a Wq^T + bias) K^T
 * scale) V.  a [B*Sq, C] bf16, wq [C, C] bf16, kv [B, L, 2C] bf16 (K | V), out [B*Sq, C] bf16; C = heads*80 <= 320, Sq % 128 == 0,
 * L <= 16.  The attention runs in the epilogue of the tcgen05 GEMM (q never reaches HBM). */
int wd_op_q_ctx_attention(const void* a_bf16, const void* wq_bf16, const float* bias, const void* kv_bf16, void* out_bf16, int B,
                          int Sq, int L, int heads, float scale, void* stream);
/* flash-style attention, any Skv; q [B,Sq,ldq], k,v [B,Skv,ldkv] (row strides in elements) */
int wd_op_attention(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B, int Sq,
                    int Skv, int heads, float scale, void* stream);
int wd_op_gemm_block_n(void);

/* The whole transformer block of unet.UNetModel's SpatialTransformer as ONE kernel (csrc/tblock.cu; reference unet.py:337-345,
 * 381-412: proj_in, two cross-attentions over the <= 16 context tokens fed by norm2, GEGLU feed-forward fed by norm3, proj_out +
 * residual).  Operator form for the parity tests: tensors[25] = g (bf16 [B*HW,320] = GroupNorm(x_in)), x_in (fp16), ctx (bf16
 * [B*L,320], the encoded context), then the 22 fp32 state_dict tensors of the block in the order listed at the definition.
 * stage 0: out = the block's output (fp16 [B*HW,320]); 1..3: the LayerNorm-normalised residual stream (no affine part) after
 * proj_in / attn1 / attn2; 4: the residual stream after the feed-forward.  gn_partial: GroupNorm partial sums of out or NULL. */
int wd_op_tblock_unet(const void* const* tensors, int n_tensors, int B, int HW, int L, int stage, void* out_f16,
                      float* gn_partial, void* stream);

/* ==== training step (reference train.py:281-294; model = unet.UNetModel, train.py:403) ==================================
 * predicted_noise = model(x_t, ..., timesteps=t, context=text_features, y=s_id); loss.backward(); optimizer.step();
 * ema.step_ema().  The trainer BINDS the caller's fp32 parameter and gradient tensors (reference state_dict layout, e.g.
 * the storage of torch nn.Parameters and of their .grad): it never owns master weights.  bf16 tensor-core packs of the
 * parameters are refreshed by wd_trainer_sync_weights (call it after every optimizer step).  Gradients are ACCUMULATED
 * (+=) into the bound buffers: zero them first (optimizer.zero_grad(), train.py:289). */
typedef struct wd_trainer wd_trainer;
int wd_trainer_create(const wd_config* cfg, wd_trainer** out);
void wd_trainer_destroy(wd_trainer* t);
/* name = reference state_dict key; w, grad: fp32 device tensors of `shape`.  Returns WD_IGNORED for parameters the
 * reference forward never reads (they receive no gradient in the reference either: SURVEY 8a, a17). */
int wd_trainer_bind_param(wd_trainer* t, const char* name, const float* w, float* grad, const int64_t* shape, int ndim);
int wd_trainer_set_pos_encoding(wd_trainer* t, const float* pe, void* stream);
int wd_trainer_sync_weights(wd_trainer* t, void* stream);
/* forward of UNetModel.forward (unet.py:1499-1836) keeping every activation the backward pass needs.
 * x: fp32 NCHW [B,4,H,W]; timesteps, y: int64 [B]; ctx_tokens: int64 [B,L]; eps_out: fp32 NCHW [B,4,H,W]. */
int wd_trainer_forward(wd_trainer* t, int batch, const float* x, const int64_t* timesteps, const int64_t* y,
                       const int64_t* ctx_tokens, int L, float* eps_out, void* stream);
/* backward of the last wd_trainer_forward: d_eps = dLoss/d eps_out, fp32 NCHW [B,4,H,W] (for train.py:287's nn.MSELoss:
 * 2 (eps - noise) / numel).  y / ctx_tokens: the same index tensors as in the forward call (embedding gradients). */
int wd_trainer_backward(wd_trainer* t, const float* d_eps, const int64_t* y, const int64_t* ctx_tokens, void* stream);
int wd_trainer_launch_counts(const wd_trainer* t, int* fwd, int* bwd);
/* Staged backward = the same launch list cut at layer boundaries, for a bucketed gradient all-reduce that overlaps the backward
 * pass (what DistributedDataParallel does for the reference's `--ddp` training, train.py:253-316).
 * wd_trainer_num_grad_stages: number of stages (one per layer, in execution order: output conv first, embeddings last).
 * wd_trainer_grad_stage: the stage after which the gradient of state_dict key `name` is final (it may be reduced then).
 * wd_trainer_backward_stages: runs stages [stage_begin, stage_end) of the last forward on `stream`; a whole backward pass is the
 * calls (0,a), (a,b), ..., (z,n) in order; d_eps is read by the call that starts at stage 0.  Each range replays as its own
 * CUDA graph from its third call on. */
int wd_trainer_num_grad_stages(wd_trainer* t, int* n);
int wd_trainer_grad_stage(wd_trainer* t, const char* name, int* stage);
int wd_trainer_backward_stages(wd_trainer* t, const float* d_eps, int stage_begin, int stage_end, void* stream);
/* test hook: copies a named intermediate of the last forward ("temb", "h1p", "h1", "embp", "emb_act" bf16 [B,*];
 * "emb_out" fp32 [B, 8*320]; "ctx", "kv_all" bf16) into dst (device); `bytes` must match. */
int wd_trainer_read_tensor(const wd_trainer* t, const char* name, void* dst, size_t bytes, void* stream);
size_t wd_trainer_workspace_bytes(const wd_trainer* t);
size_t wd_trainer_weight_bytes(const wd_trainer* t);
/* torch.optim.AdamW step (train.py:405, lr 1e-4, default betas / eps / weight_decay) fused with the EMA update of
 * train.py:140-170 over flat fp32 buffers.  step >= 1 (bias correction); grad_scale multiplies g (1/world_size after a
 * sum all-reduce).  ema_mode 0: no EMA, 1: ema = p (EMA.reset_parameters during the 2000-step warm-up), 2: ema =
 * ema_beta * ema + (1 - ema_beta) * p. */
int wd_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int step, float ema_beta, int ema_mode, float grad_scale, void* stream);

/* ---- single backward operators (parity tests; same kernels as the trainer) ---- */
/* dw[N,K] += dy[M,N]^T x[M,K] on tcgen05 (MN-major operands).  N % 320 == 0 (or N == 64), K >= 128, K % 64 == 0 */
int wd_op_wgrad_linear(const void* x_bf16, const void* dy_bf16, float* dw, int M, int N, int K, void* stream);
/* dw[Cout,Cin,3,3] += conv3x3 weight gradient; x NHWC [B,H,W,Cin], dy NHWC [B,H/stride,W/stride,Cout] */
int wd_op_wgrad_conv3x3(const void* x_bf16, const void* dy_bf16, float* dw, int B, int H, int W, int Cin, int Cout, int stride,
                        void* stream);
/* transposed packs for the data-gradient GEMMs: wd_op_conv3x3(dy, pack_T(w)) = d x (stride 1), wd_op_gemm(dy, pack_T(w)) = d x */
int wd_op_pack_conv3x3_t(const float* w_oihw, void* dst_bf16, int Cout, int Cin, void* stream);
int wd_op_pack_linear_t(const float* w, void* dst_bf16, int N, int K, void* stream);
int wd_op_groupnorm_bwd(const void* x_bf16, const void* dy_bf16, const float* gamma, const float* beta, void* dx_bf16,
                        float* dgamma, float* dbeta, int B, int HW, int C, int groups, float eps, int silu, void* stream);
int wd_op_layernorm_bwd(const void* x_bf16, const void* dy_bf16, const float* gamma, const void* add_bf16, void* dx_bf16,
                        float* dgamma, float* dbeta, int M, int C, float eps, void* stream);
int wd_op_geglu_fwd(const void* p_bf16, void* out_bf16, int M, int H, void* stream);
int wd_op_geglu_bwd(const void* p_bf16, const void* dout_bf16, void* dp_bf16, int M, int H, void* stream);
int wd_op_attention_small_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int B,
                              int Sq, int L, int heads, float scale, void* stream);

/* ==== fp32 mode (BASELINE.json north_star: "per-step predicted noise within ... 1e-4 (fp32 mode)"; configs[1] "fp32 and bf16") ====
 * The same UNet evaluated with fp32 storage and fp32 FFMA arithmetic (csrc/f32_path.cu), as the reference does with
 * use_fp16=False (unet.py:1193,1638): the accuracy mode beside the tcgen05 engine above.  Same call sequence as the engine:
 * load every state_dict entry, set the positional encoding, encode the context once per trajectory, evaluate. */
typedef struct wd_f32 wd_f32;
int wd_f32_create(const wd_config* cfg, wd_f32** out);
void wd_f32_destroy(wd_f32* e);
/* copies the fp32 tensor (3x3 conv weights are repacked [Cout][tap][Cin]); every state_dict key is accepted */
int wd_f32_load_param(wd_f32* e, const char* name, const float* src, const int64_t* shape, int ndim, void* stream);
int wd_f32_set_pos_encoding(wd_f32* e, const float* pe, void* stream);
/* CharacterEncoder (+ PHOSC tokens) in fp32 (unet.py:1626-1636,839-882 ; unetPhosc.py:1117-1130); synchronises `stream` once to
 * report token ids outside the embedding table (the reference raises IndexError) */
int wd_f32_encode_context(wd_f32* e, int batch, const int64_t* ctx_tokens, int L, const int32_t* phosc, void* stream);
/* eps = UNetModel(x, timesteps, context, y) in fp32; arguments as wd_unet_eval */
int wd_f32_unet_eval(wd_f32* e, int batch, const float* x, const int64_t* timesteps, int64_t t_scalar, const int64_t* y,
                     float* eps_out, void* stream);
/* args.attentionMaps == 1 (unet.py:1336-1364,1645-1836; unet.UNetModel only): the same evaluation, keeping the attention
 * probabilities of attn2 summed over the heads (CrossAttention returns attn, unet.py:276-279; UNetModel.forward sums it over dim 1,
 * unet.py:1786) of the LAST SpatialTransformer of the input blocks (which = 0), of the middle block (1) and of the output blocks (2)
 * until the next evaluation.  wd_f32_read_attention_map reports the stored map's H, W, L and, when dst != NULL, writes it
 * nearest-upsampled by `scale` (the reference uses 8, 16, 8: unet.py:1787,1791,1795) as fp32 [B, H*scale, W*scale, L].
 * wd_f32_read_context copies the encoded context [B, L, context_dim] (the fifth element of the reference's return tuple). */
int wd_f32_unet_eval_maps(wd_f32* e, int batch, const float* x, const int64_t* timesteps, int64_t t_scalar, const int64_t* y,
                          float* eps_out, void* stream);
int wd_f32_read_attention_map(wd_f32* e, int which, int scale, float* dst, int* H, int* W, int* L, void* stream);
int wd_f32_read_context(wd_f32* e, float* dst, size_t bytes, void* stream);
/* ==== variants of unet.UNetModel ====
 * wd_set_context / wd_f32_set_context: the context as a dense fp32 tensor [batch, L, context_dim] instead of tokens -- the
 *   reference's args.wrdChrWrStyl == 1 path, context = wrd_proj(wrdChrWrStyl) (unet.py:1590-1591,1617-1618); replaces
 *   wd_encode_context / wd_f32_encode_context for that trajectory.  wd_f32_op_linear: the projection itself in fp32.
 * wd_engine_set_label_mix / wd_f32_set_label_mix: label_emb.weight[row] = (1 - mix) w[s1] + mix w[s2], the style interpolation of
 *   args.interpolation (unet.py:1558-1572); `row` is a scratch class (the module creates its engines with one class more).
 * wd_f32_ctc_head: tdec = auxhead(eps) of args.ocrTraining == 1 (CTCtopC, unet.py:1054-1092,1829) in eval mode,
 *   eps fp32 NCHW [batch, C, H, W] -> out fp32 [256, batch, nclasses]. */
int wd_set_context(wd_engine* e, int batch, const float* ctx_f32, int L, void* stream);
/* Front and back end of the noise-prediction step (train.py:190-194 Diffusion.noise_images, :287 nn.MSELoss):
 * wd_noise_images: x_t = sqrt(alpha_hat[t]) x + sqrt(1 - alpha_hat[t]) eps in one pass over fp32 [batch, elems_per_latent]; eps comes
 *   from eps_in, or (eps_in == NULL) from the sampler's Philox4x32-10 stream keyed by (seed, sample_offset + latent, stream_id);
 *   alpha_hat: device fp32 [T]; t: device int64 [batch].  Both x_t and eps are written.
 * wd_mse_loss_grad: loss = mean((pred - target)^2) (device scalar), d_pred = 2 (pred - target) / n; deterministic reduction.
 *   workspace: wd_mse_workspace_bytes(n) bytes of device memory, zeroed once. */
int wd_noise_images(const float* x, const int64_t* t, const float* alpha_hat, int T, const float* eps_in, uint64_t seed,
                    uint64_t sample_offset, uint32_t stream_id, float* x_t, float* eps_out, int batch, int elems_per_latent, void* stream);
size_t wd_mse_workspace_bytes(size_t n);
int wd_mse_loss_grad(const float* pred, const float* target, float* d_pred, float* loss, void* workspace, size_t n, void* stream);
/* out = torch.lerp(start, end, weight), fp32 (train.py:226-228: the guidance mix of two evaluations) */
int wd_lerp(const float* start, const float* end, float weight, float* out, size_t n, void* stream);
int wd_engine_set_label_mix(wd_engine* e, int row, int s1, int s2, float mix, void* stream);
int wd_f32_set_context(wd_f32* e, int batch, const float* ctx, int L, void* stream);
int wd_f32_set_label_mix(wd_f32* e, int row, int s1, int s2, float mix, void* stream);
int wd_f32_ctc_head(wd_f32* e, int batch, const float* eps, int C, int H, int W, float* out, void* stream);
int wd_f32_op_linear(const float* x, const float* w, const float* bias, float* out, int M, int N, int K, void* stream);

/* ==== VAE decode: reference train.py:239-247, regenerateFromtrain2.py:624-636 --
 *     latents = 1 / 0.18215 * x ; image = vae.decode(latents).sample ; image = (image / 2 + 0.5).clamp(0, 1)
 * with vae = diffusers' AutoencoderKL (train.py:415).  The handle is a wd_f32 that holds the VAE's state_dict (load every
 * `post_quant_conv.*` / `decoder.*` entry with wd_f32_load_param, free with wd_f32_destroy); the layer sequence is recovered from
 * the keys.  wd_vae_decode: latents fp32 NCHW [n, 4, h, w] (multiplied by `scale`, the reference's 1 / 0.18215) ->
 * images fp32 NCHW [n, 3, 8h, 8w]; postprocess != 0 applies the reference's (image / 2 + 0.5).clamp(0, 1).  The batch is walked in
 * chunks of `chunk` latents (<= 0: 32) so the activation arena stays bounded. */
int wd_vae_create(wd_f32** out);
int wd_vae_decode(wd_f32* e, int n, const float* latents, int h, int w, float scale, int postprocess, float* images, int chunk,
                  void* stream);
int wd_f32_last_launch_count(const wd_f32* e);
size_t wd_f32_workspace_bytes(const wd_f32* e);
/* single operators of the fp32 path (parity tests): 3x3 conv pad 1 (stride 1|2, or nearest-2x upsampling first), fp32 NHWC,
 * weights in the state_dict layout [Cout,Cin,3,3]; softmax(q k^T scale) v with q [B,Sq,heads*d], k,v [B,Skv,heads*d] */
int wd_f32_op_conv3x3(const float* x_nhwc, const float* w_oihw, const float* bias, float* out_nhwc, int B, int H, int W, int Cin,
                      int Cout, int stride, int up, void* stream);
/* out[M,N] = A[M,K] W[N,K]^T + bias through the split-TF32 tcgen05 kernel (csrc/f32_gemm_tc.cu: three kind::tf32 MMAs per K step on
 * pre-split operands, chunks of K = 320 summed in fp32 registers; the fp32 mode's Linear / 1x1 / 3x3 contractions run on it unless
 * WD_F32_TC=0).  M % 128 == 0, N % 160 == 0 or N % 128 == 0, K % 32 == 0 */
int wd_f32_op_gemm_tc(const float* a, const float* w, const float* bias, float* out, int M, int N, int K, void* stream);
/* 3x3 pad-1 stride-1 convolution over cat([x1, x2], channel) (x2 NULL: one source) as an implicit GEMM on the split-TF32 kernel */
int wd_f32_op_conv3x3_tc(const float* x1, const float* x2, const float* w_oihw, const float* bias, float* out_nhwc, int B, int H, int W,
                         int C1, int C2, int Cout, void* stream);
int wd_f32_op_attention(const float* q, const float* k, const float* v, float* out, int B, int Sq, int Skv, int heads, int d,
                        float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WD_B200_H */
