// Implicit-GEMM on the 5th-gen tensor cores (tcgen05 + TMEM accumulators + TMA operand staging).
//
// One persistent, warp-specialised kernel serves every dense contraction on the WordDiffusion hot path:
//   * ResBlock / Upsample / Downsample 3x3 convolutions (reference unet.py:595,621,488,540):
//       A rows are gathered straight from the NHWC bf16 activation by 4-D TMA boxes, one box per
//       filter tap with shifted (w,h) start coordinates; out-of-bounds rows are zero-filled by the TMA
//       unit, which is exactly the conv's zero padding.  No im2col buffer exists.
//   * 1x1 convolutions and nn.Linear (unet.py:364,375,632,175-183,125,145,611,1202-1204): plain 2-D TMA.
//   * "K-concatenated" fusions: up to three A sources are walked back to back along K, so the
//       640->320 skip 1x1 conv of the decoder ResBlocks (unet.py:632,671) accumulates into the same TMEM
//       tile as the block's second 3x3 conv.
//   * the output convolution 320 -> 4 (unet.py:1457) with a 16-column tile whose epilogue applies the
//       DDPM / DDIM update of train.py:229-236 and writes x_{t-1} (fp32 NCHW) directly.
// Structure: grid = min(tiles, SMs) CTAs, 1 CTA / SM, tiles round-robin.  warp 0 = TMA producer (smem ring),
// warp 1 = single-thread tcgen05.mma issuer, warps 2..9 = epilogue.  The fp32 accumulator is double-buffered in
// TMEM, so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue: a warp may only read the TMEM lane quarter (warp % 4), so two warps share each quarter and split the
// 160 tile columns 80/80 ("column halves"; the two halves are independent pipelines).  Each warp drains its 32 rows x 80
// columns with five tcgen05.ld.x16 in flight, hands the accumulator back to the MMA warp, applies the fusions in
// registers, and writes bf16 into a shared-memory staging tile made of dense [128 rows][40 columns] sub-tiles (80-byte
// rows: bank-conflict-free for one-row-per-thread 16-byte accesses), which one elected thread per half stores with TMA
// (cp.async.bulk.tensor, bulk groups).  A residual operand is prefetched by TMA into the same staging sub-tiles one tile
// ahead and added in place.  Nothing on the output path is an uncoalesced per-thread global access.
// Epilogue fusions: +bias[N], +row-bias[sample,N] (timestep-embedding add, unet.py:657-666), +residual, SiLU,
// GEGLU (unet.py:127-129), bf16 / fp32 store, and per-(sample, group) GroupNorm partial statistics of the tensor
// being written (consumed by groupnorm_apply_kernel; unet.py:429-431).
#pragma once
#include "common.cuh"

namespace wd {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GEMM_BLOCK_N = 160;  // 320 = 2 x 160; UMMA shape 128 x 160 x 16
constexpr int GEMM_BLOCK_N_OUT = 16;  // output-conv tile (4 real columns)
constexpr int GEMM_EPI_WARPS = 8;   // two per TMEM lane quarter
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp0: TMA producer, warp1: TMEM alloc + MMA issuer, warps2-9: epilogue
constexpr int GEMM_SUB_N = 40;      // staging / TMA-store sub-tile width (columns)
constexpr int GEMM_MAX_SRC = 3;
constexpr int GEMM_ATT_MAXL = 16;  // longest key sequence of the fused context-attention epilogue

enum GemmAct : int { ACT_NONE = 0, ACT_SILU = 1 };
enum GemmEpi : int { EPI_STD = 0, EPI_SAMPLER = 1 };
enum StepMode : int { STEP_EPS_ONLY = 0, STEP_DDPM = 1, STEP_DDIM = 2 };

// Per-step quantities of the sampling loop in DEVICE memory: a captured CUDA graph of the step cannot carry them as kernel
// arguments (they change every step), so the graphed launch sequence reads them through a pointer that a tiny kernel refreshes
// before each replay (engine.cu: wd_sampler_step).
struct StepParams {
  long long t;  // timestep of the whole batch
  float4 coef;
  unsigned long long seed;
  unsigned long long sample_offset;
  int mode;  // StepMode
  int use_philox;
  int step_index;
  int pad;
};

struct GemmArgs {
  int M;  // rows (pixels / tokens)
  int N;  // accumulator columns (weight rows)
  int num_src;
  int taps[GEMM_MAX_SRC];    // 1 or 9
  int chunks[GEMM_MAX_SRC];  // channels / 64
  int stride[GEMM_MAX_SRC];  // conv stride of that source (1 or 2)
  int a_f16[GEMM_MAX_SRC];   // 1: this source (and the weight columns it multiplies) is fp16, not bf16 (residual-stream tensors)
  int conv;                  // 1: 4-D (c,w,h,n) coordinates, 0: 2-D (c, row)
  int Wout;                  // output width  (conv)
  int HWout;                 // output pixels per image (conv)
  int epi;                   // GemmEpi
  // ---- standard epilogue ----
  const float* bias;             // [N] or null (already permuted for GEGLU)
  const float* rowbias;          // [rows, rb_ld] fp32 or null
  const long long* rowbias_idx;  // optional: row = rowbias_idx[m / rows_per_sample]
  int rb_ld;
  int rows_per_sample;
  const __nv_bfloat16* residual;  // [M, res_ld] or null
  int res_ld;
  void* out;  // bf16 (or fp32 when out_f32) [M, out_ld]
  int out_ld;
  int out_f32;
  int out_f16;  // 16-bit output format: 0 bf16 (MMA operands of later GEMMs), 1 fp16 (residual stream / tensors consumed by norms)
  int res_f16;  // format of the residual tensor
  // Residual folded into the accumulator (weight-stationary build, fp16 residual): the residual tensor is streamed through the
  // operand ring as extra K blocks that multiply a 64 x 64 fp16 IDENTITY tile, so `acc += residual` happens on the tensor
  // pipe and the epilogue never waits for a residual tile (the TMA-prefetched residual arrived late: ncu showed a quarter of
  // all warp samples of the K = 320 GEMMs on that wait).  Set by gemm_prepare_res_k(); mapA[1] = residual as an A operand,
  // mapA[2] = the identity tile.
  int res_k;
  int act;
  int geglu;  // 1: tile columns [0,BN/2) are values, [BN/2,BN) gates; writes BN/2 columns per tile
  // LayerNorm folded into the GEMMs around it (unet.py:314-316 + the Linear that follows, e.g. :337 / :175):
  //   producer side: ln_out != null -> every epilogue thread also writes {sum x, sum x^2} of its 80 output columns of its row:
  //                  ln_out[(m * (N/80) + column_block) * 2 + {0,1}]   (fp32, before the 16-bit rounding)
  //   consumer side: ln_stats != null -> A is the un-normalised tensor, the weights are gamma (.) W, and the epilogue applies
  //                  y = rstd_m * (acc - mu_m * ln_s[n]) + bias[n]  with bias = W beta + b and ln_s[n] = sum_k gamma_k W[n,k]
  float* ln_out;
  const float* ln_stats;  // [M][ln_slots][2]
  int ln_slots;           // 80-column blocks per row of the normalised tensor (its channels / 80)
  int ln_dim;             // channels of the normalised tensor
  float ln_eps;
  const float* ln_s;      // [N]
  // Fused short-context cross-attention (unet.py:185-207 with the 10-token character context): the GEMM is the to_q projection
  // (N = heads * 80); each epilogue thread holds the 80 q values of one (row, head) and computes softmax(q K^T scale) V against
  // the sample's precomputed K / V rows, so q never reaches HBM.  att_kv: bf16 [samples, att_L, att_ld], K of head h at column
  // 80 h, V at column att_voff + 80 h.  Needs the weight-stationary build, rows_per_sample % 128 == 0, M % 128 == 0, att_L <= 16.
  const __nv_bfloat16* att_kv;  // null: off
  int att_ld;
  int att_voff;
  int att_L;
  float att_scale;
  // GroupNorm partial statistics of the written tensor: gn_partial[sample][N/gn_cpg groups][rows_per_sample/32][2]
  float* gn_partial;  // null: off.  Needs gn_cpg == 10, rows_per_sample % 32 == 0
  int gn_cpg;
  // GroupNorm + SiLU of the written tensor applied by the PRODUCER (pair kernel only; unet.py:592-596 after :657-666):
  // `out` receives silu(GroupNorm(acc + bias + row-bias)) as bf16 and the raw tensor never reaches HBM.  A 256-row pair tile
  // holds whole samples (256 % rows_per_sample == 0), so the per-(sample, group) statistics are complete once the 16 epilogue
  // warps of the pair have published their partials (gn_partial, exchanged through L2 behind a cluster-scope mbarrier).
  // The values wait as fp16 (what the separate GroupNorm kernel used to read): round 0 in the staging tile, round 1 in registers.
  int gn_apply;
  // Sub-pixel form of "nearest 2x upsample, then conv3x3" (unet.py:497-499): output phase (a, b) = (up_phase - 1) >> 1, & 1 of the
  // 2H x 2W image is a 2 x 2 convolution of the H x W input with the taps that hit the same input pixel summed into one weight
  // (9/4 fewer MACs, no upsampled tensor).  taps[0] == 4: tap t reads input offset (dy, dx) = ((t >> 1) - 1 + a, (t & 1) - 1 + b);
  // the tile is stored through a 4-D output map (c, x, y, n) whose strides step two pixels; the GroupNorm partials of the
  // 2H x 2W output go to slot gn_slot_base + (row-in-phase >> 5) of gn_nslot.  0: off.  Needs 128 % (H W) == 0.
  int up_phase;  // 1..4: one phase per launch; 5: all four phases in ONE launch as N tiles (N = 4 x 320: phase = n0 / 320; single-CTA
                 // kernel only; bias replicated per phase; 5-D output map (c, b, x, a, image-row))
  int gn_slot_base;
  int gn_nslot;
  int tail_split;  // pair kernel: cut the tiles of a last, at most half-full round into 160-column halves (set by gemm_pair_launch)
  const float* gn_gamma;  // [N]
  const float* gn_beta;   // [N]
  float gn_eps;
  // ---- sampler epilogue (EPI_SAMPLER; N tile = 16, columns 0..3 = predicted-noise channels) ----
  float* eps_out;      // fp32 NCHW [B,4,H,W] or null
  float* x;            // fp32 NCHW latent, updated in place when mode != STEP_EPS_ONLY
  const float* noise;  // fp32 NCHW or null
  int use_philox;
  unsigned long long seed;
  unsigned long long sample_offset;  // global index of sample 0 of this shard (GPU-count invariant noise)
  int step_index;
  float4 coef;
  int mode;
  const StepParams* sp;  // non-null (graph replay): coef / mode / Philox arguments come from *sp instead of the fields above
  int dbg;  // experiment switches (env WD_GEMM_DBG, tools/op_bench.py): 1 no TMA store, 2 no epilogue math, 4 no B loads
};

struct GemmLaunch {
  CUtensorMap mapA[GEMM_MAX_SRC];
  CUtensorMap mapB;
  CUtensorMap mapOut;  // bf16 output [M, out columns], box {GEMM_SUB_N, GEMM_BLOCK_M}, no swizzle (unused: out_f32 / sampler)
  CUtensorMap mapRes;  // residual, same box (unused when args.residual == nullptr)
  GemmArgs args;
};

// Host helpers (gemm_tc.cu)
// If the launch qualifies (see GemmArgs::res_k), encodes L.mapA[1] / L.mapA[2] and sets args.res_k.  Call after the sources,
// epilogue fields and args.residual / res_ld are final.  Returns false only when a tensor map cannot be encoded.
bool gemm_prepare_res_k(GemmLaunch& L);
bool tmap_encode_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                         uint32_t box_inner, uint32_t box_rows);
bool tmap_encode_4d_bf16(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N,
                         uint64_t pix_stride_elems, uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n,
                         uint32_t stride_wh);
// 3-D (inner, rows, batch) map with a {box_inner, box_rows, 1} SWIZZLE_128B box (attention operands, per-sample fold operands)
bool tmap_encode_3d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t row_stride_elems,
                         uint64_t batch_stride_elems, uint32_t box_inner, uint32_t box_rows);
// output map of one sub-pixel phase (GemmArgs::up_phase): dims (C, W, H, N) of the PHASE grid over a [N, 2H, 2W, C] tensor whose
// base already points at the phase's first pixel; box {GEMM_SUB_N, W, H, 128 / (W H)}, no swizzle
bool tmap_encode_out_phase_bf16(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N);
// all four phases behind one map: dims (C, b, W, a, N H) over the [N, 2H, 2W, C] tensor (strides ascending), box {GEMM_SUB_N, 1, W, 1, 128 / W}
bool tmap_encode_out_phase5_bf16(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N);
// output / residual tensor map of the staging sub-tiles
bool tmap_encode_out_bf16(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_elems);
cudaError_t gemm_tc_launch(const GemmLaunch& L, cudaStream_t stream);
inline int gemm_tc_block_n() { return GEMM_BLOCK_N; }  // N granularity accepted by the GEMM entry points

// ---- CTA-pair kernel (gemm_pair.cu): 256 x 320 tiles, tcgen05 cta_group::2 ----
constexpr int GEMM_PAIR_BLOCK_N = 320;
bool gemm_pair_enabled();                       // env WD_GEMM_PAIR (default on)
bool gemm_pair_supported(const GemmArgs& a);    // shape / epilogue combination handled by the pair kernel
cudaError_t gemm_pair_launch(const GemmLaunch& L, int num_sms, cudaStream_t stream);
// which kernel gemm_tc_launch() picks decides two host-side layouts:
bool gemm_uses_pair(const GemmArgs& a);         // -> B tensor-map box rows (gemm_b_box_rows) ...
int gemm_b_box_rows(const GemmArgs& a);         // 16 (output conv), 80 (pair kernel), 160
int gemm_geglu_block(int N);                    // ... and the value/gate interleave width of GEGLU weights (320 or 160)

}  // namespace wd
