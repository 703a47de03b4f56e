// Host-side engine of the WordDiffusion hot path: builds the layer inventory from the constructor
// arguments (same loops as reference unet.py:1248-1458 / unetPhosc.py:864-1040), repacks state_dict tensors
// into the layouts the kernels want, owns the activation arena and the per-batch launch plan (TMA
// descriptors are encoded once per plan), and exposes the C ABI of include/wd_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/wd_b200.h"
#include "gemm_tc.cuh"
#include "ops.cuh"
#include "tblock.cuh"

using namespace wd;
typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return fail(WD_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

extern "C" const char* wd_last_error(void) { return g_err.c_str(); }
// shared with train.cu (engine_internal.h)
int wd_set_error(int code, const char* msg) {
  g_err = msg ? msg : "";
  return code;
}
extern "C" int wd_version(void) { return 1; }
static int op_gemm_impl(const void* a_, const void* w, const float* bias, const void* residual, void* out, int M, int N, int K,
                        int act_silu, int geglu, int out_f32, int res_f16, int out_f16, void* stream);
extern "C" int wd_op_gemm_block_n(void) { return gemm_tc_block_n(); }

// ----------------------------------------------------------------------------------------------
// arena
// ----------------------------------------------------------------------------------------------
struct Arena {
  char* base = nullptr;
  size_t used = 0;
  template <class T>
  T* alloc(size_t n) {
    const size_t bytes = (n * sizeof(T) + 1023) & ~size_t(1023);
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
};

// ----------------------------------------------------------------------------------------------
// layer inventory
// ----------------------------------------------------------------------------------------------
struct GemmW {
  bf16* w = nullptr;  // [N, K] bf16, K-major
  float* bias = nullptr;
  int N = 0, K = 0;
  float* ln_s = nullptr;  // LayerNorm folded into this Linear: column sums of the gamma-scaled weights (GemmArgs::ln_s)
  float* raw_w = nullptr;  // LayerNorm-folded Linears keep the fp32 state_dict tensors (the fold runs at finalize time)
  float* raw_b = nullptr;
};
// one nn.Linear whose preceding LayerNorm is folded into it at finalize time (ops.cu: fold_ln_linear_kernel)
struct LnFold {
  const float* raw_w;  // fp32 [N, K] copy of the state_dict weight
  const float* raw_b;  // fp32 [N] copy of the bias, or null
  const float* gamma;
  const float* beta;
  bf16* dst;
  float* s_out;
  float* b_out;
  int N, K, ldk, n_off, geglu_bn;
};
struct NormW {
  float* g = nullptr;
  float* b = nullptr;
  int C = 0;
};
struct ResL {
  int Cin = 0, Cout = 0;
  NormW gn1, gn2;
  GemmW conv1, conv2;  // conv2.K = 9*Cout (+ Cin when the 1x1 skip conv is fused along K)
  bool skip_conv = false;
  float* b_main = nullptr;  // out_layers.3.bias
  float* b_skip = nullptr;  // skip_connection.bias
  int emb_off = 0;          // column offset inside the fused emb_layers GEMM
};
struct TBlockL {
  NormW ln1, ln2, ln3;
  GemmW a1_q;   // unet: attn1.to_q ; phosc: fused [to_q; to_k; to_v] of the self-attention
  GemmW a1_kv;  // unet only: [to_k; to_v] applied to the context
  GemmW a1_out;
  GemmW a2_q, a2_kv, a2_out;
  GemmW ff_proj, ff_out;
  int kv1 = -1, kv2 = -1;  // index of the per-trajectory K/V buffer
  // ---- fused transformer-block kernel (tblock.cu; unet.UNetModel only) ----
  float *raw_k[2] = {nullptr, nullptr}, *raw_v[2] = {nullptr, nullptr}, *raw_o[2] = {nullptr, nullptr};  // fp32 to_k / to_v / to_out.0
  bf16* w_fold = nullptr;   // [2 attentions][2560][320] rows of the pooled fold GEMM (tblock_fold_weights_launch)
  float* u_fold = nullptr;  // [2][4][320]
  int fold_idx = -1;        // position of this block's first attention in the pools
  GemmW ff128;              // ff.net.0.proj with LayerNorm 3 folded in, value / gate rows interleaved per 64-column chunk
  float* cb = nullptr;      // [4][320] cumulative residual-stream biases
  bool mid = false;         // SpatialTransformer channels != 320: proj_in / proj_out stay GEMMs, the kernel runs its "middle" form
};
struct STL {
  int C = 0, heads = 0, dh = 0;
  NormW gn;
  GemmW proj_in, proj_out;
  std::vector<TBlockL> blocks;
};
struct SampL {
  int C = 0;
  GemmW conv;
  // Upsample only: sub-pixel weights [4 phases][C][4 taps][C] (GemmArgs::up_phase) folded from the fp32 copy at finalize
  bf16* w_phase = nullptr;
  float* raw_w = nullptr;
  float* bias4 = nullptr;  // the bias once per phase ([4 C]: all phases in one launch, GemmArgs::up_phase == 5)
  int as_f16 = 0;
};
enum LayerKind { L_CONVIN, L_RES, L_ST, L_DOWN, L_UP };
struct Layer {
  LayerKind kind;
  int idx;
};
typedef std::vector<Layer> Block;

enum SlotKind { S_VEC, S_CONV3, S_LIN, S_CONV_IN, S_F32 };
struct Slot {
  SlotKind kind;
  void* dst;
  int64_t numel;
  int N, K;  // S_LIN: [N,K]; S_CONV3: Cout, Cin
  int ldk, k_off, n_off, geglu_bn;
  bool loaded;
  int as_f16;  // weight columns that multiply an fp16 (residual-stream) A source are stored as fp16
};

// ----------------------------------------------------------------------------------------------
// launch plan
// ----------------------------------------------------------------------------------------------
enum OpKind { OP_TEMB, OP_GEMM, OP_GN, OP_LN, OP_ATTN_SMALL, OP_ATTN_FLASH, OP_CONV_IN, OP_GNSTATS, OP_UPSAMPLE,
              OP_EMBED, OP_LINF32, OP_WORDATTN, OP_EMBTBL, OP_TBLOCK, OP_TB_CVEC, OP_OUTHEAD };
// step ops that exist in two flavours: the time-embedding MLP per step (per-row timesteps: wd_unet_eval) or the lookup in
// the per-trajectory table (one timestep for the whole batch: wd_sampler_step)
enum OpCond { COND_ALWAYS = 0, COND_NO_TABLE = 1, COND_TABLE = 2 };
constexpr int TEMB_TABLE_ROWS = 1024;  // timesteps 0 .. 1023 (the reference schedules use T = 1000 / 600)
enum Patch { P_NONE = 0, P_Y = 1, P_SAMPLER = 2 };
struct Op {
  OpKind kind;
  int patch = P_NONE;
  int cond = COND_ALWAYS;
  double flops = 0;  // algorithmic work of this launch ("algorithmic work per op")
  double bytes = 0;
  GemmLaunch gemm;
  GroupNormArgs gn;
  GroupNormStatsArgs gs;
  int gn_B = 0, gn_nslab = 0;
  struct { const bf16* x; bf16* out; const float* g; const float* b; int M, C; float eps; int x_f16; } ln;
  AttnSmallArgs as;
  AttnFlashArgs af;
  struct { bf16* out; int B, dim; int table; } temb;  // table != 0: row b embeds t = b
  struct { const float* table; bf16* out; int B, dim; int use_label; } etbl;
  struct { bf16* out; int B, H, W; } cin;  // im2col of the latent
  struct { const bf16* x; bf16* out; int B, H, W, C; } up;
  struct { int which; const float* E; int vocab; const float* pe; int add_pe; float* out; int B, L, D; } emb;
  struct { const float* x; const float* W; const float* b; float* out; int M, N, K; } lin;
  struct { const float* q; const float* k; const float* v; bf16* ctx; int B, L, D, Ltot, row_off; } wa;
  TBlockLaunch tb;
  struct { const bf16* ctx; const float* u; float* out; int rows, heads; } cv;
  OutHeadArgs oh;  // per-call fields (x, eps_out, noise, sampler scalars) are patched in like P_SAMPLER
};

struct Plan {
  int B = 0, L = 0, Ltot = 0;
  std::vector<Op> ctx_ops;
  std::vector<Op> step_ops;
  size_t bytes = 0;
  bool context_valid = false;
  bf16* ctx_buf = nullptr;  // encoded context [B, Ltot, context_dim] (wd_set_context writes it directly)
  // CUDA graphs of the step launch sequence, each valid for one set of caller pointers: g_step for wd_sampler_step (the
  // per-step scalars live in wd_engine::sp_dev), g_eval for wd_unet_eval with per-row timesteps (the forward() path).
  // `seen` counts consecutive calls with the same pointers: the first runs eagerly, the second captures, later ones replay.
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    const void* key[4] = {nullptr, nullptr, nullptr, nullptr};
    int seen = 0;
    int launches = 0;
    ~GraphSlot() {
      if (exec) cudaGraphExecDestroy(exec);
    }
  };
  GraphSlot g_step, g_eval;
};

struct Act {
  bf16* p;
  int C, H, W;
  float* stats;  // GroupNorm partial statistics of this tensor: [B][32][pslots][2] fp32 (groups of C/32 channels)
  int pslots;    // 0: not computed
  bool f16;      // stored as fp16 (residual-stream tensors and tensors only read by norms), else bf16 (MMA operands)
};

struct wd_engine {
  wd_config cfg;
  int time_dim = 0;
  // weights
  Arena warena;
  char* wbase = nullptr;
  size_t wbytes = 0;
  std::unordered_map<std::string, Slot> slots;
  std::vector<ResL> res;
  std::vector<STL> st;
  std::vector<SampL> samp;
  std::vector<Block> input_blocks, output_blocks;
  Block middle;
  GemmW te0, te2, emb_all;
  float* label_emb = nullptr;
  GemmW conv_in;  // [model_channels, 128] bf16: im2col columns hi(36) | lo(36) | zero pad
  NormW out_gn;
  GemmW conv_out;  // [16, 9*C] bf16, rows >= out_channels are zero
  // context encoder (fp32)
  float *we_E = nullptr, *we_qw = nullptr, *we_qb = nullptr, *we_kw = nullptr, *we_kb = nullptr, *we_vw = nullptr,
        *we_vb = nullptr, *we_pe = nullptr;
  bool pe_set = false;
  int n_kv = 0;
  std::vector<GemmW*> kv_weights;  // per K/V buffer: the fused [to_k; to_v] weight
  StepParams* sp_dev = nullptr;    // device-resident per-step scalars of the graphed sampling step
  cudaStream_t cap_stream = nullptr;  // private stream the step graph is captured on
  bool prof_used_table = false;    // flavour of the time-embedding ops in the profiled steps (wd_engine_profile_read)
  int kv_fused_count = 0;          // kv_fused() calls (counted by the dry layout pass): their weights share one pool, so the
  bf16* kv_pool = nullptr;         //   K/V projections of ALL cross-attentions run as a single GEMM per trajectory
  int kv_pool_next = 0;
  // CharacterEncoder folded into lookup tables (finalize_params): T* = E W^T + b [vocab, D], P* = pe W^T [max_seq_len, D]
  float *we_tq = nullptr, *we_tk = nullptr, *we_tv = nullptr, *we_pq = nullptr, *we_pk = nullptr, *we_pv = nullptr;
  float* temb_table = nullptr;  // time_embed(t) for t < TEMB_TABLE_ROWS, fp32 [rows, 4 mc]: built once per weight load
  float* we_gram = nullptr;  // TQ TK^T [vocab, vocab]: the scores of a position-free Word_Attention segment (ops.cu: word_attn_hist_kernel)
  std::vector<LnFold> ln_folds;
  // fused transformer block (tblock.cu): extra fp32 copies of state_dict tensors, and the pooled per-attention fold weights
  std::unordered_multimap<std::string, float*> raw_extra;
  int fold_count = 0;      // attentions with fold weights (counted by the dry layout pass)
  int fold_next = 0;
  bf16* fold_pool = nullptr;   // [fold_count][2560][320] bf16: ONE GEMM per trajectory produces every per-sample attention operand
  float* u_pool = nullptr;     // [fold_count][4][320]
  bf16* tb_identity = nullptr; // 320 x 320 fp16 identity: "proj_in" of the fused kernel's middle form (x I = the residual stream)
  // activations
  char* abase = nullptr;
  size_t acap = 0;
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans;
  Plan* cur = nullptr;
  int last_launches = 0;
  // per-op device timing (bench.py): events recorded on the launching stream around every op of a step
  bool prof_on = false;
  std::vector<std::vector<cudaEvent_t>> prof_steps;  // each: n_ops + 1 events
  const Plan* prof_plan = nullptr;
};

// ----------------------------------------------------------------------------------------------
// model builder
// ----------------------------------------------------------------------------------------------
namespace {

struct Builder {
  wd_engine* e;
  Arena& A;
  bool dry;
  int bn;

  void slot(const std::string& name, SlotKind kind, void* dst, int64_t numel, int N = 0, int K = 0, int ldk = 0,
            int k_off = 0, int n_off = 0, int geglu_bn = 0, int as_f16 = 0) {
    if (dry) return;
    Slot s{kind, dst, numel, N, K, ldk, k_off, n_off, geglu_bn, false, as_f16};
    e->slots[name] = s;
  }
  NormW norm(const std::string& pfx, int C) {
    NormW n;
    n.C = C;
    n.g = A.alloc<float>(C);
    n.b = A.alloc<float>(C);
    slot(pfx + ".weight", S_VEC, n.g, C);
    slot(pfx + ".bias", S_VEC, n.b, C);
    return n;
  }
  // nn.Linear / 1x1 conv [N, K] (+ optional bias)
  GemmW linear(const std::string& pfx, int N, int K, bool bias, int geglu_bn = 0, int as_f16 = 0) {
    GemmW g;
    g.N = N;
    g.K = K;
    g.w = A.alloc<bf16>(static_cast<size_t>(N) * K);
    slot(pfx + ".weight", S_LIN, g.w, static_cast<int64_t>(N) * K, N, K, K, 0, 0, geglu_bn, as_f16);
    if (bias) {
      g.bias = A.alloc<float>(N);
      slot(pfx + ".bias", S_VEC, g.bias, N, N, 0, 0, 0, 0, geglu_bn);
    }
    return g;
  }
  // an additional fp32 copy of a state_dict tensor (besides the destination of its slot)
  float* raw_copy(const std::string& name, int64_t numel) {
    float* p = A.alloc<float>(static_cast<size_t>(numel));
    if (!dry) e->raw_extra.emplace(name, p);
    return p;
  }
  // nn.Linear(K, Nrows) whose input is LayerNorm `ln` of an fp16 token tensor: rows land at [n_off, n_off + Nrows) of `g`
  // (g.w / g.bias / g.ln_s are allocated by the caller for fused weights, or here when g.w is null)
  void linear_ln(GemmW& g, const std::string& pfx, int Nrows, int K, bool bias, const NormW& ln, int n_off = 0, int Ntotal = 0,
                 int geglu_bn = 0) {
    if (!g.w) {
      g.N = Ntotal ? Ntotal : Nrows;
      g.K = K;
      g.w = A.alloc<bf16>(static_cast<size_t>(g.N) * K);
      g.bias = A.alloc<float>(g.N);
      g.ln_s = A.alloc<float>(g.N);
    }
    float* raw_w = A.alloc<float>(static_cast<size_t>(Nrows) * K);
    float* raw_b = bias ? A.alloc<float>(Nrows) : nullptr;
    slot(pfx + ".weight", S_F32, raw_w, static_cast<int64_t>(Nrows) * K);
    if (bias) slot(pfx + ".bias", S_F32, raw_b, Nrows);
    g.raw_w = raw_w;
    g.raw_b = raw_b;
    if (!dry) e->ln_folds.push_back(LnFold{raw_w, raw_b, ln.g, ln.b, g.w, g.ln_s, g.bias, Nrows, K, K, n_off, geglu_bn});
  }
  GemmW conv3(const std::string& pfx, int Cout, int Cin, int extraK = 0, int as_f16 = 0) {
    GemmW g;
    g.N = Cout;
    g.K = 9 * Cin + extraK;
    g.w = A.alloc<bf16>(static_cast<size_t>(g.N) * g.K);
    g.bias = A.alloc<float>(Cout);
    slot(pfx + ".weight", S_CONV3, g.w, static_cast<int64_t>(Cout) * Cin * 9, Cout, Cin, g.K, 0, 0, 0, as_f16);
    return g;
  }

  int add_res(const std::string& pfx, int Cin, int Cout, int& emb_cols) {
    ResL r;
    r.Cin = Cin;
    r.Cout = Cout;
    r.gn1 = norm(pfx + "in_layers.0", Cin);
    r.conv1 = conv3(pfx + "in_layers.2", Cout, Cin);
    slot(pfx + "in_layers.2.bias", S_VEC, r.conv1.bias, Cout);
    r.emb_off = emb_cols;
    emb_cols += Cout;
    r.gn2 = norm(pfx + "out_layers.0", Cout);
    r.skip_conv = (Cin != Cout);
    r.conv2 = conv3(pfx + "out_layers.3", Cout, Cout, r.skip_conv ? Cin : 0);
    r.b_main = A.alloc<float>(Cout);
    slot(pfx + "out_layers.3.bias", S_VEC, r.b_main, Cout);
    if (r.skip_conv) {
      r.b_skip = A.alloc<float>(Cout);
      // the 1x1 skip conv reads the (fp16) residual-stream tensors directly: its K range of the fused weight is fp16
      slot(pfx + "skip_connection.weight", S_LIN, r.conv2.w, static_cast<int64_t>(Cout) * Cin, Cout, Cin, r.conv2.K,
           9 * Cout, 0, 0, 1);
      slot(pfx + "skip_connection.bias", S_VEC, r.b_skip, Cout);
    }
    e->res.push_back(r);
    return static_cast<int>(e->res.size()) - 1;
  }

  // [to_k ; to_v] fused along N
  GemmW kv_fused(const std::string& pfx, int inner, int ctx_dim) {
    GemmW g;
    g.N = 2 * inner;
    g.K = ctx_dim;
    const size_t elems = static_cast<size_t>(g.N) * g.K;
    if ((elems * sizeof(bf16)) % 1024 != 0) {
      g.w = A.alloc<bf16>(elems);
    } else if (dry) {
      ++e->kv_fused_count;
      g.w = A.alloc<bf16>(elems);
    } else {  // one contiguous pool in call order (same bytes as the dry pass: every slice is a whole number of 1 KB units)
      if (!e->kv_pool) e->kv_pool = A.alloc<bf16>(elems * e->kv_fused_count);
      g.w = e->kv_pool + elems * e->kv_pool_next++;
    }
    slot(pfx + ".to_k.weight", S_LIN, g.w, static_cast<int64_t>(inner) * ctx_dim, inner, ctx_dim, ctx_dim, 0, 0, 0);
    slot(pfx + ".to_v.weight", S_LIN, g.w, static_cast<int64_t>(inner) * ctx_dim, inner, ctx_dim, ctx_dim, 0, inner, 0);
    return g;
  }

  int add_st(const std::string& pfx, int C, int heads, int dh) {
    STL s;
    s.C = C;
    s.heads = heads;
    s.dh = dh;
    const int inner = heads * dh;
    const int ctx_dim = e->cfg.context_dim;
    s.gn = norm(pfx + "norm", C);
    s.proj_in = linear(pfx + "proj_in", inner, C, true);
    for (int d = 0; d < e->cfg.transformer_depth; ++d) {
      const std::string tp = pfx + "transformer_blocks." + std::to_string(d) + ".";
      TBlockL t;
      // the three LayerNorms are folded into the Linear that consumes them (q / qkv projections and the GEGLU projection)
      if (e->cfg.variant == WD_VARIANT_PHOSC) t.ln1 = norm(tp + "norm1", inner);
      t.ln2 = norm(tp + "norm2", inner);
      t.ln3 = norm(tp + "norm3", inner);
      if (e->cfg.variant == WD_VARIANT_PHOSC) {
        // self-attention: q, k, v all from LN1(x) -> one N = 3*inner GEMM
        linear_ln(t.a1_q, tp + "attn1.to_q", inner, inner, false, t.ln1, 0, 3 * inner);
        linear_ln(t.a1_q, tp + "attn1.to_k", inner, inner, false, t.ln1, inner);
        linear_ln(t.a1_q, tp + "attn1.to_v", inner, inner, false, t.ln1, 2 * inner);
      } else {
        // unet.py:337-341 -- attn1 is a cross-attention over the context, fed by norm2 (norm1 is never applied)
        linear_ln(t.a1_q, tp + "attn1.to_q", inner, inner, false, t.ln2);
        t.a1_kv = kv_fused(tp + "attn1", inner, ctx_dim);
      }
      t.a1_out = linear(tp + "attn1.to_out.0", inner, inner, true);
      linear_ln(t.a2_q, tp + "attn2.to_q", inner, inner, false, t.ln2);
      t.a2_kv = kv_fused(tp + "attn2", inner, ctx_dim);
      t.a2_out = linear(tp + "attn2.to_out.0", inner, inner, true);
      linear_ln(t.ff_proj, tp + "ff.net.0.proj", inner * 8, inner, true, t.ln3, 0, 0, gemm_geglu_block(inner * 8));
      t.ff_out = linear(tp + "ff.net.2", inner, inner * 4, true);
      if (e->cfg.variant == WD_VARIANT_UNET && inner == TB_C && heads == TB_HEADS && dh == TB_DH && ctx_dim == TB_C) {
        t.mid = C != TB_C;
        if (t.mid && !e->tb_identity) e->tb_identity = A.alloc<bf16>(static_cast<size_t>(TB_C) * TB_C);
        // operands of the fused transformer-block kernel (tblock.cuh)
        for (int a = 0; a < 2; ++a) {
          const std::string ap = tp + (a ? "attn2" : "attn1");
          t.raw_k[a] = raw_copy(ap + ".to_k.weight", static_cast<int64_t>(inner) * ctx_dim);
          t.raw_v[a] = raw_copy(ap + ".to_v.weight", static_cast<int64_t>(inner) * ctx_dim);
          t.raw_o[a] = raw_copy(ap + ".to_out.0.weight", static_cast<int64_t>(inner) * inner);
        }
        const size_t fold_elems = static_cast<size_t>(2) * TB_FOLD_N * TB_C;
        if (dry) {
          e->fold_count += 2;
          t.w_fold = A.alloc<bf16>(fold_elems);
          t.u_fold = A.alloc<float>(2 * TB_HEADS * TB_C);
        } else {
          if (!e->fold_pool) {
            e->fold_pool = A.alloc<bf16>(static_cast<size_t>(e->fold_count) * TB_FOLD_N * TB_C);
            e->u_pool = A.alloc<float>(static_cast<size_t>(e->fold_count) * TB_HEADS * TB_C);
          }
          t.fold_idx = e->fold_next;
          t.w_fold = e->fold_pool + static_cast<size_t>(e->fold_next) * TB_FOLD_N * TB_C;
          t.u_fold = e->u_pool + static_cast<size_t>(e->fold_next) * TB_HEADS * TB_C;
          e->fold_next += 2;
        }
        // second fold of ff.net.0.proj from the same fp32 tensors, value / gate rows interleaved per 64-column chunk
        t.ff128.N = inner * 8;
        t.ff128.K = inner;
        t.ff128.w = A.alloc<bf16>(static_cast<size_t>(t.ff128.N) * inner);
        t.ff128.bias = A.alloc<float>(t.ff128.N);
        t.ff128.ln_s = A.alloc<float>(t.ff128.N);
        if (!dry)
          e->ln_folds.push_back(LnFold{t.ff_proj.raw_w, t.ff_proj.raw_b, t.ln3.g, t.ln3.b, t.ff128.w, t.ff128.ln_s, t.ff128.bias,
                                       inner * 8, inner, inner, 0, 2 * TB_CHUNK});
        t.cb = A.alloc<float>(4 * TB_C);
      }
      s.blocks.push_back(t);
    }
    s.proj_out = linear(pfx + "proj_out", C, inner, true, 0, 1);  // A operand = x3, a residual-stream (fp16) tensor
    e->st.push_back(s);
    return static_cast<int>(e->st.size()) - 1;
  }

  int add_samp(const std::string& pfx, int C, bool upsample = false) {
    SampL s;
    s.C = C;
    s.conv = conv3(pfx, C, C, 0, 1);  // Down/Upsample convs read residual-stream (fp16) tensors
    slot(pfx + ".bias", S_VEC, s.conv.bias, C);
    if (upsample) {
      s.w_phase = A.alloc<bf16>(static_cast<size_t>(4) * C * 4 * C);
      s.raw_w = raw_copy(pfx + ".weight", static_cast<int64_t>(C) * C * 9);
      s.bias4 = A.alloc<float>(static_cast<size_t>(4) * C);
      s.as_f16 = 1;
    }
    e->samp.push_back(s);
    return static_cast<int>(e->samp.size()) - 1;
  }

  bool in_attn_res(int ds) const {
    for (int i = 0; i < e->cfg.n_attention_resolutions; ++i)
      if (e->cfg.attention_resolutions[i] == ds) return true;
    return false;
  }
  void heads_for(int ch, int& heads, int& dh) const {
    if (e->cfg.num_head_channels == -1) {
      heads = e->cfg.num_heads;
      dh = ch / heads;
    } else {
      heads = ch / e->cfg.num_head_channels;
      dh = e->cfg.num_head_channels;
    }
  }

  void build() {
    const wd_config& c = e->cfg;
    const int mc = c.model_channels;
    const int ted = mc * 4;
    e->time_dim = ted;
    e->res.clear();
    e->ln_folds.clear();
    if (dry) e->fold_count = 0;
    e->fold_next = 0;
    e->fold_pool = nullptr;
    e->u_pool = nullptr;
    e->tb_identity = nullptr;
    e->raw_extra.clear();
    e->st.clear();
    e->samp.clear();
    e->input_blocks.clear();
    e->output_blocks.clear();
    e->middle.clear();

    e->te0 = linear("time_embed.0", ted, mc, true);
    e->te2 = linear("time_embed.2", ted, ted, true);
    e->temb_table = A.alloc<float>(static_cast<size_t>(TEMB_TABLE_ROWS) * ted);
    // CharacterEncoder (fp32)
    const int D = c.context_dim;
    e->we_E = A.alloc<float>(static_cast<size_t>(c.vocab_size) * D);
    slot("word_emb.embedding.weight", S_F32, e->we_E, static_cast<int64_t>(c.vocab_size) * D);
    e->we_qw = A.alloc<float>(static_cast<size_t>(D) * D);
    e->we_kw = A.alloc<float>(static_cast<size_t>(D) * D);
    e->we_vw = A.alloc<float>(static_cast<size_t>(D) * D);
    e->we_qb = A.alloc<float>(D);
    e->we_kb = A.alloc<float>(D);
    e->we_vb = A.alloc<float>(D);
    e->we_pe = A.alloc<float>(static_cast<size_t>(c.max_seq_len) * D);
    e->we_tq = A.alloc<float>(static_cast<size_t>(c.vocab_size) * D);
    e->we_tk = A.alloc<float>(static_cast<size_t>(c.vocab_size) * D);
    e->we_tv = A.alloc<float>(static_cast<size_t>(c.vocab_size) * D);
    e->we_pq = A.alloc<float>(static_cast<size_t>(c.max_seq_len) * D);
    e->we_pk = A.alloc<float>(static_cast<size_t>(c.max_seq_len) * D);
    e->we_pv = A.alloc<float>(static_cast<size_t>(c.max_seq_len) * D);
    e->we_gram = A.alloc<float>(static_cast<size_t>(c.vocab_size) * c.vocab_size);
    slot("word_emb.attention.linear_query.weight", S_F32, e->we_qw, static_cast<int64_t>(D) * D);
    slot("word_emb.attention.linear_query.bias", S_F32, e->we_qb, D);
    slot("word_emb.attention.linear_key.weight", S_F32, e->we_kw, static_cast<int64_t>(D) * D);
    slot("word_emb.attention.linear_key.bias", S_F32, e->we_kb, D);
    slot("word_emb.attention.linear_value.weight", S_F32, e->we_vw, static_cast<int64_t>(D) * D);
    slot("word_emb.attention.linear_value.bias", S_F32, e->we_vb, D);
    if (c.num_classes > 0) {
      e->label_emb = A.alloc<float>(static_cast<size_t>(c.num_classes) * ted);
      slot("label_emb.weight", S_F32, e->label_emb, static_cast<int64_t>(c.num_classes) * ted);
    }
    // input_blocks.0.0 : conv_in
    e->conv_in.N = mc;
    e->conv_in.K = 128;
    e->conv_in.w = A.alloc<bf16>(static_cast<size_t>(mc) * 128);
    e->conv_in.bias = A.alloc<float>(mc);
    slot("input_blocks.0.0.weight", S_CONV_IN, e->conv_in.w, static_cast<int64_t>(mc) * c.in_channels * 9, mc, c.in_channels);
    slot("input_blocks.0.0.bias", S_VEC, e->conv_in.bias, mc);
    e->input_blocks.push_back(Block{Layer{L_CONVIN, 0}});

    int emb_cols = 0;
    std::vector<int> chans{mc};
    int ch = mc, ds = 1;
    // unet.py:1258-1318
    for (int level = 0; level < c.n_channel_mult; ++level) {
      const int mult = c.channel_mult[level];
      for (int i = 0; i < c.num_res_blocks; ++i) {
        const std::string pfx = "input_blocks." + std::to_string(e->input_blocks.size()) + ".";
        Block b;
        b.push_back(Layer{L_RES, add_res(pfx + "0.", ch, mult * mc, emb_cols)});
        ch = mult * mc;
        if (in_attn_res(ds)) {
          int heads, dh;
          heads_for(ch, heads, dh);
          b.push_back(Layer{L_ST, add_st(pfx + "1.", ch, heads, dh)});
        }
        e->input_blocks.push_back(b);
        chans.push_back(ch);
      }
      if (level != c.n_channel_mult - 1) {
        const std::string pfx = "input_blocks." + std::to_string(e->input_blocks.size()) + ".0.op";
        e->input_blocks.push_back(Block{Layer{L_DOWN, add_samp(pfx, ch)}});
        chans.push_back(ch);
        ds *= 2;
      }
    }
    // middle block, unet.py:1366-1394
    {
      int heads, dh;
      heads_for(ch, heads, dh);
      e->middle.push_back(Layer{L_RES, add_res("middle_block.0.", ch, ch, emb_cols)});
      e->middle.push_back(Layer{L_ST, add_st("middle_block.1.", ch, heads, dh)});
      e->middle.push_back(Layer{L_RES, add_res("middle_block.2.", ch, ch, emb_cols)});
    }
    // output blocks, unet.py:1398-1451
    for (int level = c.n_channel_mult - 1; level >= 0; --level) {
      const int mult = c.channel_mult[level];
      for (int i = 0; i < c.num_res_blocks + 1; ++i) {
        const int ich = chans.back();
        chans.pop_back();
        const std::string pfx = "output_blocks." + std::to_string(e->output_blocks.size()) + ".";
        Block b;
        int li = 0;
        b.push_back(Layer{L_RES, add_res(pfx + std::to_string(li++) + ".", ch + ich, mc * mult, emb_cols)});
        ch = mc * mult;
        if (in_attn_res(ds)) {
          int heads, dh;
          heads_for(ch, heads, dh);
          b.push_back(Layer{L_ST, add_st(pfx + std::to_string(li++) + ".", ch, heads, dh)});
        }
        if (level && i == c.num_res_blocks) {
          b.push_back(Layer{L_UP, add_samp(pfx + std::to_string(li++) + ".conv", ch, true)});
          ds /= 2;
        }
        e->output_blocks.push_back(b);
      }
    }
    // out, unet.py:1454-1458
    e->out_gn = norm("out.0", ch);
    e->conv_out.N = GEMM_BLOCK_N_OUT;
    e->conv_out.K = 9 * ch;
    e->conv_out.w = A.alloc<bf16>(static_cast<size_t>(GEMM_BLOCK_N_OUT) * 9 * ch);
    e->conv_out.bias = A.alloc<float>(GEMM_BLOCK_N_OUT);
    slot("out.2.weight", S_CONV3, e->conv_out.w, static_cast<int64_t>(c.out_channels) * ch * 9, c.out_channels, ch, 9 * ch, 0, 0, 0,
         2 /* hi rows 0..3, lo rows 4..7 */);
    slot("out.2.bias", S_VEC, e->conv_out.bias, c.out_channels);

    // fused emb_layers GEMM: [sum Cout, time_dim]
    e->emb_all.N = emb_cols;
    e->emb_all.K = ted;
    e->emb_all.w = A.alloc<bf16>(static_cast<size_t>(emb_cols) * ted);
    e->emb_all.bias = A.alloc<float>(emb_cols);
    if (!dry) {
      int ri = 0;
      auto reg = [&](const std::string& pfx, const ResL& r) {
        slot(pfx + "emb_layers.1.weight", S_LIN, e->emb_all.w, static_cast<int64_t>(r.Cout) * ted, r.Cout, ted, ted, 0,
             r.emb_off, 0);
        slot(pfx + "emb_layers.1.bias", S_VEC, e->emb_all.bias, r.Cout, r.Cout, 0, 0, 0, r.emb_off, 0);
        ++ri;
      };
      for (size_t bi = 0; bi < e->input_blocks.size(); ++bi)
        for (size_t li = 0; li < e->input_blocks[bi].size(); ++li)
          if (e->input_blocks[bi][li].kind == L_RES)
            reg("input_blocks." + std::to_string(bi) + "." + std::to_string(li) + ".", e->res[e->input_blocks[bi][li].idx]);
      for (size_t li = 0; li < e->middle.size(); ++li)
        if (e->middle[li].kind == L_RES) reg("middle_block." + std::to_string(li) + ".", e->res[e->middle[li].idx]);
      for (size_t bi = 0; bi < e->output_blocks.size(); ++bi)
        for (size_t li = 0; li < e->output_blocks[bi].size(); ++li)
          if (e->output_blocks[bi][li].kind == L_RES)
            reg("output_blocks." + std::to_string(bi) + "." + std::to_string(li) + ".", e->res[e->output_blocks[bi][li].idx]);
    }
    // K/V buffers of the cross-attentions (time-invariant)
    e->n_kv = 0;
    e->kv_weights.clear();
    for (auto& s : e->st)
      for (auto& t : s.blocks) {
        if (c.variant == WD_VARIANT_UNET) {
          t.kv1 = e->n_kv++;
          e->kv_weights.push_back(&t.a1_kv);
        }
        t.kv2 = e->n_kv++;
        e->kv_weights.push_back(&t.a2_kv);
      }
  }
};

bool is_dead_param(const std::string& n, int variant) {
  // parameters that exist in the reference state_dict but are never read by its forward (SURVEY 8a, a8')
  if (n.find(".attnc.") != std::string::npos) return true;
  if (n.find(".to_kv.") != std::string::npos) return true;
  if (n.rfind("res.", 0) == 0) return true;
  if (n.rfind("wrd_proj.", 0) == 0) return true;
  // args.ocrTraining == 1 / args.charImages == 1 (unet.py:1468,1217-1223): the OCR head reads the UNet's output (wd_f32_ctc_head),
  // the character-image convolutions feed a value the reference forward discards (unet.py:1625-1627)
  if (n.rfind("auxhead.", 0) == 0 || n.rfind("conv_layer1.", 0) == 0 || n.rfind("conv_layer2.", 0) == 0 || n.rfind("conv_layer3.", 0) == 0)
    return true;
  if (variant == WD_VARIANT_UNET && n.find(".norm1.") != std::string::npos) return true;
  if (variant == WD_VARIANT_PHOSC) {
    // the self-attention's to_k/to_v are used; nothing else is dead
  }
  return false;
}

int validate(const wd_config& c) {
  const int bn = gemm_tc_block_n();
  if (c.variant != WD_VARIANT_UNET && c.variant != WD_VARIANT_PHOSC) return fail(WD_ERR_INVALID, "bad variant");
  if (c.in_channels != 4 || c.out_channels != 4) return fail(WD_ERR_UNSUPPORTED, "in/out channels must be 4");
  if (c.model_channels % 64 || c.model_channels % bn)
    return fail(WD_ERR_UNSUPPORTED, "model_channels must be a multiple of 64 and of the GEMM N tile (%d)", bn);
  if (c.n_channel_mult < 1 || c.n_channel_mult > 8) return fail(WD_ERR_INVALID, "channel_mult length");
  if (c.context_dim % 64) return fail(WD_ERR_UNSUPPORTED, "context_dim must be a multiple of 64");
  if (c.num_heads == -1 && c.num_head_channels == -1) return fail(WD_ERR_INVALID, "num_heads or num_head_channels");
  if (c.transformer_depth < 1) return fail(WD_ERR_INVALID, "transformer_depth");
  if (c.latent_h < 1 || c.latent_w < 1) return fail(WD_ERR_INVALID, "latent size");
  return WD_OK;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// C ABI: life cycle and weights
// ----------------------------------------------------------------------------------------------
extern "C" int wd_engine_create(const wd_config* cfg, wd_engine** out) {
  if (!cfg || !out) return fail(WD_ERR_INVALID, "null argument");
  int rc = validate(*cfg);
  if (rc) return rc;
  int dev = 0, major = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(WD_ERR_UNSUPPORTED, "wd_b200 needs an sm_100a GPU (found compute capability %d.x)", major);
  std::unique_ptr<wd_engine> e(new wd_engine());
  e->cfg = *cfg;
  {
    Arena dryA;
    Builder b{e.get(), dryA, true, gemm_tc_block_n()};
    b.build();
    e->wbytes = dryA.used;
  }
  CUDA_TRY(cudaMalloc(&e->wbase, e->wbytes));
  CUDA_TRY(cudaMemset(e->wbase, 0, e->wbytes));
  e->warena.base = e->wbase;
  e->warena.used = 0;
  Builder b{e.get(), e->warena, false, gemm_tc_block_n()};
  b.build();
  for (auto& s : e->st)
    for (auto& t : s.blocks)
      if (s.dh != 80) return fail(WD_ERR_UNSUPPORTED, "attention kernels are built for d_head = 80 (got %d)", s.dh);
  *out = e.release();
  return WD_OK;
}

extern "C" void wd_engine_destroy(wd_engine* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  if (e->wbase) cudaFree(e->wbase);
  if (e->abase) cudaFree(e->abase);
  if (e->sp_dev) cudaFree(e->sp_dev);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  e->plans.clear();
  delete e;
}

extern "C" size_t wd_engine_weight_bytes(const wd_engine* e) { return e ? e->wbytes : 0; }
extern "C" size_t wd_engine_workspace_bytes(const wd_engine* e) { return e ? e->acap : 0; }
extern "C" int wd_engine_last_launch_count(const wd_engine* e) { return e ? e->last_launches : 0; }

extern "C" int wd_engine_load_param(wd_engine* e, const char* name, const float* src, const int64_t* shape, int ndim,
                                    void* stream) {
  if (!e || !name || !src) return fail(WD_ERR_INVALID, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto it = e->slots.find(name);
  if (it == e->slots.end()) {
    if (is_dead_param(name, e->cfg.variant)) return WD_IGNORED;
    return fail(WD_ERR_INVALID, "unknown parameter '%s'", name);
  }
  Slot& sl = it->second;
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  if (numel != sl.numel) return fail(WD_ERR_INVALID, "parameter '%s': %lld elements, expected %lld", name, (long long)numel,
                                     (long long)sl.numel);
  switch (sl.kind) {
    case S_VEC:
      CUDA_TRY(repack_vec_launch(src, static_cast<float*>(sl.dst), static_cast<int>(numel), sl.n_off, sl.geglu_bn, 0, s));
      break;
    case S_F32:
      CUDA_TRY(cudaMemcpyAsync(sl.dst, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
      break;
    case S_CONV3:
      CUDA_TRY(repack_conv3x3_launch(src, static_cast<bf16*>(sl.dst), sl.N, sl.K, sl.ldk, sl.k_off, sl.as_f16, s));
      break;
    case S_LIN:
      CUDA_TRY(repack_linear_launch(src, static_cast<bf16*>(sl.dst), sl.N, sl.K, sl.ldk, sl.k_off, sl.n_off, sl.geglu_bn,
                                    sl.as_f16, s));
      break;
    case S_CONV_IN:
      CUDA_TRY(repack_conv_in_launch(src, static_cast<bf16*>(sl.dst), sl.N, sl.K, s));
      break;
  }
  sl.loaded = true;
  auto extra = e->raw_extra.equal_range(name);
  for (auto x = extra.first; x != extra.second; ++x)
    CUDA_TRY(cudaMemcpyAsync(x->second, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return WD_OK;
}

extern "C" int wd_engine_set_pos_encoding(wd_engine* e, const float* pe, void* stream) {
  if (!e || !pe) return fail(WD_ERR_INVALID, "null argument");
  CUDA_TRY(cudaMemcpyAsync(e->we_pe, pe, static_cast<size_t>(e->cfg.max_seq_len) * e->cfg.context_dim * sizeof(float),
                           cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  e->pe_set = true;
  return WD_OK;
}

extern "C" int wd_engine_finalize_params(wd_engine* e, void* stream) {
  if (!e) return fail(WD_ERR_INVALID, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int missing = 0;
  std::string first;
  for (auto& kv : e->slots)
    if (!kv.second.loaded) {
      if (!missing) first = kv.first;
      ++missing;
    }
  if (!e->pe_set) {
    ++missing;
    if (first.empty()) first = "<positional encoding>";
  }
  if (missing) {
    fail(WD_ERR_STATE, "%d parameters not loaded (first: %s)", missing, first.c_str());
    return missing;
  }
  {
    // CharacterEncoder (unet.py:839-882): q/k/v = Linear(E[token] + pe[position]) = (E W^T + b)[token] + (pe W^T)[position].  The
    // inputs come from a finite vocabulary x position set, so the three fp32 Linears over [B, L, D] become two small tables
    // each, built here once per weight load; per trajectory only a gather + add remains (the Linears were 180 of the 250 us
    // the context encoding took at batch 256).
    const wd_config& c = e->cfg;
    const int D = c.context_dim;
    const float* Ws[3] = {e->we_qw, e->we_kw, e->we_vw};
    const float* bs[3] = {e->we_qb, e->we_kb, e->we_vb};
    float* Ts[3] = {e->we_tq, e->we_tk, e->we_tv};
    float* Ps[3] = {e->we_pq, e->we_pk, e->we_pv};
    for (int i = 0; i < 3; ++i) {
      CUDA_TRY(linear_f32_launch(e->we_E, Ws[i], bs[i], Ts[i], c.vocab_size, D, D, s));
      CUDA_TRY(linear_f32_launch(e->we_pe, Ws[i], nullptr, Ps[i], c.max_seq_len, D, D, s));
    }
    CUDA_TRY(word_attn_gram_launch(e->we_tq, e->we_tk, e->we_gram, c.vocab_size, D, s));
    // time_embed(t) for every timestep (unet.py:1575-1576: Linear, SiLU, Linear of the sinusoid): weights only, so once per load
    {
      const int mc = c.model_channels, ted = e->time_dim;
      bf16 *temb_t = nullptr, *h1_t = nullptr;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&temb_t), static_cast<size_t>(TEMB_TABLE_ROWS) * mc * sizeof(bf16)));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h1_t), static_cast<size_t>(TEMB_TABLE_ROWS) * ted * sizeof(bf16)));
      int rc = WD_OK;
      if (timestep_embed_launch(nullptr, -1, temb_t, TEMB_TABLE_ROWS, mc, s) != cudaSuccess) rc = fail(WD_ERR_CUDA, "time-embedding table: sinusoid");
      if (rc == WD_OK) rc = op_gemm_impl(temb_t, e->te0.w, e->te0.bias, nullptr, h1_t, TEMB_TABLE_ROWS, ted, mc, 1, 0, 0, 0, 0, s);
      if (rc == WD_OK) rc = op_gemm_impl(h1_t, e->te2.w, e->te2.bias, nullptr, e->temb_table, TEMB_TABLE_ROWS, ted, ted, 0, 0, 1, 0, 0, s);
      cudaStreamSynchronize(s);
      cudaFree(temb_t);
      cudaFree(h1_t);
      if (rc != WD_OK) return rc;
    }
  }
  for (auto& sp : e->samp)
    if (sp.w_phase && sp.raw_w) {
      CUDA_TRY(upconv_phase_fold_launch(sp.raw_w, sp.w_phase, sp.C, sp.C, sp.as_f16, s));
      for (int ph = 0; ph < 4; ++ph)
        CUDA_TRY(cudaMemcpyAsync(sp.bias4 + ph * sp.C, sp.conv.bias, static_cast<size_t>(sp.C) * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
  for (auto& f : e->ln_folds)
    CUDA_TRY(fold_ln_linear_launch(f.raw_w, f.gamma, f.beta, f.raw_b, f.dst, f.s_out, f.b_out, f.N, f.K, f.ldk, f.n_off, f.geglu_bn, s));
  for (auto& st : e->st)
    for (auto& t : st.blocks) {
      if (!t.w_fold) continue;
      // fused transformer block: W_fold / u of both cross-attentions (norm2 feeds both, unet.py:337-341), cumulative biases
      const GemmW* q[2] = {&t.a1_q, &t.a2_q};
      for (int a = 0; a < 2; ++a)
        CUDA_TRY(tblock_fold_weights_launch(q[a]->raw_w, t.raw_k[a], t.raw_v[a], t.raw_o[a], t.ln2.g, t.ln2.b,
                                            t.w_fold + static_cast<size_t>(a) * TB_FOLD_N * TB_C, t.u_fold + a * TB_HEADS * TB_C, s));
      const float* add[4] = {st.proj_in.bias, t.a1_out.bias, t.a2_out.bias, t.ff_out.bias};
      for (int i = 0; i < 4; ++i) {
        if (i > 0) CUDA_TRY(repack_vec_launch(t.cb + (i - 1) * TB_C, t.cb + i * TB_C, TB_C, 0, 0, 0, s));
        if (i == 0 && t.mid) {  // the proj_in GEMM in front of the kernel has already added its bias
          CUDA_TRY(cudaMemsetAsync(t.cb, 0, TB_C * sizeof(float), s));
        } else {
          CUDA_TRY(repack_vec_launch(add[i], t.cb + i * TB_C, TB_C, 0, 0, i > 0 ? 1 : 0, s));
        }
      }
    }
  if (e->tb_identity) {
    std::vector<uint16_t> eye(static_cast<size_t>(TB_C) * TB_C, 0);
    for (int i = 0; i < TB_C; ++i) eye[static_cast<size_t>(i) * TB_C + i] = 0x3C00;  // fp16 1.0
    CUDA_TRY(cudaMemcpyAsync(e->tb_identity, eye.data(), eye.size() * 2, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
  }
  for (auto& r : e->res) {
    CUDA_TRY(repack_vec_launch(r.b_main, r.conv2.bias, r.Cout, 0, 0, 0, s));
    if (r.skip_conv) CUDA_TRY(repack_vec_launch(r.b_skip, r.conv2.bias, r.Cout, 0, 0, 1, s));
  }
  return 0;
}

// ----------------------------------------------------------------------------------------------
// plan builder
// ----------------------------------------------------------------------------------------------
namespace {

struct ASrc {
  const bf16* p;
  int C;   // channels taken from this source
  int ld;  // pixel / row stride in elements
  int taps;
  int stride;
  int H, W;  // input spatial size (conv mode)
  bool f16 = false;  // fp16 source (its weight columns are packed as fp16 too)
};
struct Epi {
  const float* rowbias = nullptr;
  int rb_ld = 0;
  int rows_per_sample = 1;
  const bf16* residual = nullptr;
  int res_ld = 0;
  void* out = nullptr;
  int out_ld = 0;
  int out_f32 = 0;
  int out_f16 = 0;
  int res_f16 = 0;
  int act = 0;
  int geglu = 0;
  float* ln_out = nullptr;          // write LayerNorm row statistics of the output tensor
  const float* ln_stats = nullptr;  // A is an un-normalised tensor with these row statistics (weights carry gamma, see GemmW::ln_s)
  int ln_dim = 0;
  Act* stats_for = nullptr;  // output tensor whose GroupNorm partials the epilogue should write (if it can)
  int up_phase = 0;  // sub-pixel phase of "upsample, then conv3x3" (GemmArgs::up_phase): `out` = the [B, 2H, 2W, C] tensor
  const NormW* gn_apply = nullptr;  // producer-side GroupNorm + SiLU (GemmArgs::gn_apply): `out` receives the normalised tensor
  float gn_eps = 1e-5f;
  int epi = EPI_STD;
  // context attention fused into the to_q projection (GemmArgs::att_*): K / V rows [B, att_L, att_ld], `out` receives softmax(qK^T)V
  const bf16* att_kv = nullptr;
  int att_ld = 0, att_voff = 0, att_L = 0;
  float att_scale = 0.f;
};

// env WD_UP_PHASE: Upsample + conv3x3 as four sub-pixel 2 x 2 convolutions.  0: off (upsample kernel + 3 x 3 conv); 1: one launch
// per phase; 2 (default): all four phases in one launch
static int up_phase_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_UP_PHASE");
    v = e ? atoi(e) : 2;
  }
  return v;
}
static bool up_phase_enabled() { return up_phase_mode() != 0; }

// env WD_OUT_HEAD=1: out GroupNorm + conv_out + sampler update as ONE kernel (ops.cu out_head_kernel).  Default OFF: measured
// neutral at batch 32 / 256 (the two launches it replaces overlap their neighbours through PDL; its own body is bound by the
// legacy HMMA rate) and slower at batch 1 (one CTA per sample: 0.778 vs 0.720 ms per step); and the choice must not depend on
// the batch size, because the two forms sum in different orders and batch invariance is bit-exact (tests/test_gpu_model.py).
static bool out_head_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_OUT_HEAD");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}

// env WD_GN_PRODUCER: GroupNorm of a ResBlock's h applied by conv1's epilogue.  0: never; 1: only when the conv has more
// 256-row tiles than the GPU has CTA pairs, so that all but the last tile's epilogue runs under the next tile's MMAs;
// 2: always (conv1 then runs on the pair kernel whatever its size).  Default 1.
static int gn_producer_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_GN_PRODUCER");
    v = e ? atoi(e) : 1;
  }
  return v;
}

struct PlanBuilder {
  wd_engine* e;
  Plan* plan;
  Arena A;
  bool dry;
  int B;
  std::string err;
  int kv_ld = 0;  // row stride of the K/V buffers when all of them come from one pooled GEMM (0: each buffer is [.., 2C])
  // fused transformer block: output of the pooled fold GEMM, fp16 [B * Ltot, fold_ld] (per attention: 1280 "M" ++ 1280 "N" columns),
  // and the per-row score constants fp32 [B * Ltot, cvec_ld]
  bf16* fold_out = nullptr;
  int fold_ld = 0;
  float* cvec_all = nullptr;
  int cvec_ld = 0;

  Act new_act(int H, int W, int C, bool f16 = false) {
    Act a{A.alloc<bf16>(static_cast<size_t>(B) * H * W * C), C, H, W, nullptr, 0, f16};
    a.stats = A.alloc<float>(static_cast<size_t>(B) * 32 * 8 * 2);
    return a;
  }
  // can the GEMM epilogue emit the GroupNorm partials of an [B, H*W, C] output?  (16 groups of 10 columns per 160-wide tile)
  static bool epilogue_stats_ok(int HW, int C) { return C % 32 == 0 && C / 32 == 10 && HW % 32 == 0 && HW / 32 <= 8; }
  // make sure `a` carries GroupNorm partials: free when the producing GEMM wrote them, else one reduction kernel
  bool ensure_stats(std::vector<Op>& ops, Act& a) {
    if (a.pslots) return true;
    if (a.C % 32 || a.C % 8) { err = "groupnorm: channels must be a multiple of 32"; return false; }
    Op op;
    memset(&op, 0, sizeof(op));
    op.kind = OP_GNSTATS;
    a.pslots = groupnorm_stats_slots(a.H * a.W);
    op.gs = GroupNormStatsArgs{a.p, a.C, a.stats, a.H * a.W, a.C, a.C / 32, a.pslots, a.f16 ? 1 : 0};
    op.gn_B = B;
    op.bytes = 2.0 * B * a.H * a.W * a.C;
    ops.push_back(op);
    return true;
  }

  bool gemm_op(std::vector<Op>& ops, int M, bool conv, int Hout, int Wout, const std::vector<ASrc>& srcs, const GemmW& w,
               const Epi& ep, int patch = P_NONE) {
    Op op;
    memset(&op, 0, sizeof(op));
    op.kind = OP_GEMM;
    op.patch = patch;
    GemmArgs& a = op.gemm.args;
    a.M = M;
    a.N = w.N;
    a.num_src = static_cast<int>(srcs.size());
    a.conv = conv ? 1 : 0;
    a.Wout = conv ? Wout : 1;
    a.HWout = conv ? Hout * Wout : 1;
    a.bias = w.bias;
    a.rowbias = ep.rowbias;
    a.rowbias_idx = nullptr;
    a.rb_ld = ep.rb_ld;
    a.rows_per_sample = ep.rows_per_sample;
    a.residual = ep.residual;
    a.res_ld = ep.res_ld;
    a.out = ep.out;
    a.out_ld = ep.out_ld;
    a.out_f32 = ep.out_f32;
    a.out_f16 = ep.out_f16;
    a.res_f16 = ep.res_f16;
    a.ln_out = ep.ln_out;
    a.ln_stats = ep.ln_stats;
    a.ln_slots = ep.ln_dim / 80;
    a.ln_dim = ep.ln_dim;
    a.ln_eps = 1e-5f;  // nn.LayerNorm default (unet.py:314-316)
    a.ln_s = w.ln_s;
    if (ep.ln_stats && (!w.ln_s || ep.ln_dim % 80)) { err = "gemm: LayerNorm folding needs folded weights and channels % 80 == 0"; return false; }
    a.act = ep.act;
    a.geglu = ep.geglu;
    a.epi = ep.epi;
    a.att_kv = ep.att_kv;
    a.att_ld = ep.att_ld;
    a.att_voff = ep.att_voff;
    a.att_L = ep.att_L;
    a.att_scale = ep.att_scale;
    const int n_stats = ep.up_phase == 5 ? w.N / 4 : w.N;  // (all four phases as N tiles: the output tensor has N / 4 channels)
    if (ep.stats_for && epilogue_stats_ok(ep.rows_per_sample, n_stats) && ep.stats_for->C == n_stats) {
      a.gn_partial = ep.stats_for->stats;
      a.gn_cpg = 10;
      ep.stats_for->pslots = ep.rows_per_sample / 32;
      if (ep.up_phase) {  // the phase grid's rows are a quarter of the output sample: 4 phases x (H W / 32) slots
        a.gn_nslot = 4 * (ep.rows_per_sample / 32);
        a.gn_slot_base = ep.up_phase == 5 ? 0 : (ep.up_phase - 1) * (ep.rows_per_sample / 32);
        ep.stats_for->pslots = a.gn_nslot;
      }
    }
    a.up_phase = ep.up_phase;
    if (ep.up_phase && (!conv || srcs.size() != 1 || srcs[0].taps != 4 || GEMM_BLOCK_M % (Hout * Wout) || !a.gn_partial)) {
      err = "gemm: the sub-pixel upsample conv needs one 4-tap source, whole images per tile and epilogue statistics";
      return false;
    }
    if (ep.gn_apply) {
      if (!a.gn_partial || ep.gn_apply->C != w.N) { err = "gemm: producer-side GroupNorm needs epilogue statistics"; return false; }
      a.gn_apply = 1;
      a.gn_gamma = ep.gn_apply->g;
      a.gn_beta = ep.gn_apply->b;
      a.gn_eps = ep.gn_eps;
    }
    int ktot = 0;
    if (srcs.empty() || srcs.size() > GEMM_MAX_SRC) { err = "gemm: bad source count"; return false; }
    for (size_t i = 0; i < srcs.size(); ++i) {
      const ASrc& s = srcs[i];
      if (s.C % GEMM_BLOCK_K) { err = "gemm: source channels must be a multiple of 64"; return false; }
      a.taps[i] = s.taps;
      a.chunks[i] = s.C / GEMM_BLOCK_K;
      a.stride[i] = s.stride;
      a.a_f16[i] = s.f16 ? 1 : 0;
      ktot += s.taps * s.C;
      if (dry) continue;
      bool ok;
      if (!conv) {
        ok = tmap_encode_2d_bf16(&op.gemm.mapA[i], s.p, s.C, M, s.ld, GEMM_BLOCK_K, GEMM_BLOCK_M);
      } else {
        const int HWout = Hout * Wout;
        uint32_t bw, bh, bnn;
        if (HWout >= GEMM_BLOCK_M) {
          if (HWout % GEMM_BLOCK_M || GEMM_BLOCK_M % Wout) { err = "conv: unsupported spatial size"; return false; }
          bw = Wout * s.stride;
          bh = (GEMM_BLOCK_M / Wout) * s.stride;
          bnn = 1;
        } else {
          if (GEMM_BLOCK_M % HWout) { err = "conv: unsupported spatial size"; return false; }
          bw = s.W;
          bh = s.H;
          bnn = GEMM_BLOCK_M / HWout;
        }
        if (bw > 256 || bh > 256 || bnn > 256) { err = "conv: TMA box too large"; return false; }
        ok = tmap_encode_4d_bf16(&op.gemm.mapA[i], s.p, s.C, s.W, s.H, B, s.ld, GEMM_BLOCK_K, bw, bh, bnn, s.stride);
      }
      if (!ok) { err = "cuTensorMapEncodeTiled failed (A)"; return false; }
    }
    if (ktot != w.K) { err = "gemm: K mismatch between sources and weight"; return false; }
    op.flops = 2.0 * M * static_cast<double>(w.N) * w.K;
    op.bytes = 2.0 * (static_cast<double>(M) * w.K + static_cast<double>(w.N) * w.K + static_cast<double>(M) * (ep.geglu ? w.N / 2 : w.N));
    if (ep.att_kv) {  // + QK^T and PV of the fused context attention, + the K / V rows
      op.flops += 4.0 * M * static_cast<double>(w.N) * ep.att_L;
      op.bytes += 2.0 * 2.0 * (static_cast<double>(M) / ep.rows_per_sample) * ep.att_L * w.N;
    }
    const int bn = (ep.epi == EPI_SAMPLER) ? GEMM_BLOCK_N_OUT : gemm_tc_block_n();
    if (w.N % bn) { err = "gemm: N must be a multiple of the N tile"; return false; }
    const int b_box = gemm_b_box_rows(a);
    if (!dry) {
      for (size_t i = srcs.size(); i < GEMM_MAX_SRC; ++i) op.gemm.mapA[i] = op.gemm.mapA[0];
      if (!tmap_encode_2d_bf16(&op.gemm.mapB, w.w, w.K, w.N, w.K, GEMM_BLOCK_K, b_box)) {
        err = "cuTensorMapEncodeTiled failed (B)";
        return false;
      }
      op.gemm.mapOut = op.gemm.mapB;
      op.gemm.mapRes = op.gemm.mapB;
      if (ep.up_phase == 5) {
        if (ep.out_ld != w.N / 4 || !tmap_encode_out_phase5_bf16(&op.gemm.mapOut, ep.out, w.N / 4, Wout, Hout, B)) {
          err = "cuTensorMapEncodeTiled failed (out phase5)";
          return false;
        }
      } else if (ep.up_phase) {
        // phase (a, b) starts at pixel (a, b) of every [2H, 2W] image and steps two pixels / two rows
        const int pa = (ep.up_phase - 1) >> 1, pb = (ep.up_phase - 1) & 1;
        const bf16* base = static_cast<const bf16*>(ep.out) + (static_cast<size_t>(pa) * 2 * Wout + pb) * w.N;
        if (ep.out_ld != w.N || !tmap_encode_out_phase_bf16(&op.gemm.mapOut, base, w.N, Wout, Hout, B)) {
          err = "cuTensorMapEncodeTiled failed (out phase)";
          return false;
        }
      } else if (ep.epi != EPI_SAMPLER && !ep.out_f32 &&
                 !tmap_encode_out_bf16(&op.gemm.mapOut, ep.out, ep.geglu ? w.N / 2 : w.N, M, ep.out_ld)) {
        err = "cuTensorMapEncodeTiled failed (out)";
        return false;
      }
      if (ep.residual && !tmap_encode_out_bf16(&op.gemm.mapRes, ep.residual, w.N, M, ep.res_ld)) {
        err = "cuTensorMapEncodeTiled failed (residual)";
        return false;
      }
      if (!gemm_prepare_res_k(op.gemm)) { err = "cuTensorMapEncodeTiled failed (residual as K blocks)"; return false; }
    }
    ops.push_back(op);
    if (ep.stats_for && !ensure_stats(ops, *ep.stats_for)) return false;
    return true;
  }

  // GroupNorm over the channel-concatenation of `srcs` -> one [B,HW,sumC] tensor
  bool gn_op(std::vector<Op>& ops, const std::vector<Act>& srcs, const NormW& nw, float eps, int silu, Act& out) {
    int totalC = 0;
    for (auto& s : srcs) totalC += s.C;
    if (totalC != nw.C || totalC % 32) { err = "groupnorm: channel mismatch"; return false; }
    const int cpg = totalC / 32;
    const int H = srcs[0].H, W = srcs[0].W, HW = H * W;
    // one slab per concatenated source: it must hold whole groups and whole 16-byte vectors
    const int Cs = srcs[0].C;
    for (auto& s : srcs)
      if (s.C != Cs) { err = "groupnorm: concat sources must have equal channels"; return false; }
    if (Cs % cpg || Cs % 8 || Cs / 8 * 8 > 1024) { err = "groupnorm: slab does not hold whole groups"; return false; }
    const int per_src = 1;
    const int nslab = static_cast<int>(srcs.size());
    if (nslab > 2) { err = "groupnorm: too many slabs"; return false; }
    out = new_act(H, W, totalC);
    Op op;
    memset(&op, 0, sizeof(op));
    op.kind = OP_GN;
    for (int i = 0; i < nslab; ++i) {
      const Act& s = srcs[i / per_src];
      op.gn.x[i] = s.p + (i % per_src) * Cs;
      op.gn.x_ld[i] = s.C;
      op.gn.x_f16[i] = s.f16 ? 1 : 0;
    }
    op.gn.out = out.p;
    op.gn.out_ld = totalC;
    op.gn.gamma = nw.g;
    op.gn.beta = nw.b;
    op.gn.HW = HW;
    op.gn.Cs = Cs;
    op.gn.cpg = cpg;
    op.gn.eps = eps;
    op.gn.silu = silu;
    op.gn.pcpg = Cs / 32;
    op.gn.nchunk = groupnorm_apply_chunks(HW);
    for (int i = 0; i < nslab; ++i) {
      if (!srcs[i].pslots) { err = "groupnorm: source tensor carries no statistics"; return false; }
      op.gn.partial[i] = srcs[i].stats;
      op.gn.pslots[i] = srcs[i].pslots;
    }
    op.gn_B = B;
    op.gn_nslab = nslab;
    op.bytes = 4.0 * B * HW * totalC;  // bf16 read + bf16 write
    ops.push_back(op);
    return true;
  }

  void ln_op(std::vector<Op>& ops, const Act& x, bf16* out, const NormW& nw, int M) {
    Op op;
    memset(&op, 0, sizeof(op));
    op.kind = OP_LN;
    op.ln = {x.p, out, nw.g, nw.b, M, nw.C, 1e-5f, x.f16 ? 1 : 0};
    op.bytes = 4.0 * M * nw.C;
    ops.push_back(op);
  }

  bool attn_op(std::vector<Op>& ops, const bf16* q, int q_ld, const bf16* k, const bf16* v, int kv_ld, bf16* out, int out_ld,
               int Sq, int Skv, int heads, int dh) {
    Op op;
    memset(&op, 0, sizeof(op));
    const float scale = 1.0f / sqrtf(static_cast<float>(dh));
    // tensor-core flash kernel for every key length (16-key tile for the character context); the SIMT
    // attn_small kernel only remains for the attention-probability output (wd_op_attention_small)
    op.kind = OP_ATTN_FLASH;
    op.af = AttnFlashArgs{q, q_ld, k, v, kv_ld, out, out_ld, Sq, Skv, heads, scale};
    op.flops = 4.0 * B * heads * static_cast<double>(Sq) * Skv * dh;  // QK^T + PV
    op.bytes = 2.0 * B * heads * dh * (2.0 * Sq + 2.0 * Skv);
    ops.push_back(op);
    return true;
  }

  // ---- ResBlock (unet.py:646-671) ----
  bool res_block(std::vector<Op>& ops, const ResL& r, const std::vector<Act>& in, const float* emb_out, int emb_ld, Act& out) {
    const int H = in[0].H, W = in[0].W, HW = H * W;
    Act a1;
    if (!gn_op(ops, in, r.gn1, 1e-5f, 1, a1)) return false;
    Act a2;
    // conv1 -> GroupNorm -> SiLU (unet.py:657-667 then :592-594): when conv1 runs on the pair kernel and its 256-row tiles hold
    // whole samples, the epilogue normalises the tile itself and h never reaches HBM (one launch and one HBM round trip fewer)
    // (R4d: 2.405 -> 2.353 ms per batch-256 step when only the multi-tile launches are fused; single-wave launches expose the
    //  longer epilogue -- its straight-line code runs once, from a cold instruction cache -- and gain nothing)
    bool fuse_gn2 = gn_producer_mode() > 0 && gemm_pair_enabled() && epilogue_stats_ok(HW, r.Cout) && 256 % HW == 0 &&
                    r.Cout == GEMM_PAIR_BLOCK_N && r.gn2.C == r.Cout && a1.C % GEMM_BLOCK_K == 0 && 9 * a1.C / GEMM_BLOCK_K >= 40;
    if (fuse_gn2 && gn_producer_mode() == 1) {
      int sms = 148, dev = 0;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      fuse_gn2 = (static_cast<long long>(B) * HW + 255) / 256 > sms / 2;
    }
    if (fuse_gn2) {
      a2 = new_act(H, W, r.Cout);
      Act hs{nullptr, r.Cout, H, W, A.alloc<float>(static_cast<size_t>(B) * 32 * 8 * 2), 0, true};  // statistics of h only
      Epi ep;
      ep.rowbias = emb_out + r.emb_off;
      ep.rb_ld = emb_ld;
      ep.rows_per_sample = HW;
      ep.out = a2.p;
      ep.out_ld = r.Cout;
      ep.stats_for = &hs;
      ep.gn_apply = &r.gn2;
      ep.gn_eps = 1e-5f;
      if (!gemm_op(ops, B * HW, true, H, W, {ASrc{a1.p, a1.C, a1.C, 9, 1, H, W}}, r.conv1, ep)) return false;
    } else {
      Act h2 = new_act(H, W, r.Cout, true);  // only read by GroupNorm: fp16
      Epi ep;
      ep.rowbias = emb_out + r.emb_off;
      ep.rb_ld = emb_ld;
      ep.rows_per_sample = HW;
      ep.out = h2.p;
      ep.out_ld = r.Cout;
      ep.out_f16 = 1;
      ep.stats_for = &h2;
      if (!gemm_op(ops, B * HW, true, H, W, {ASrc{a1.p, a1.C, a1.C, 9, 1, H, W}}, r.conv1, ep)) return false;
      if (!gn_op(ops, {h2}, r.gn2, 1e-5f, 1, a2)) return false;
    }
    out = new_act(H, W, r.Cout, true);  // residual stream: fp16
    {
      Epi ep;
      ep.rows_per_sample = HW;
      ep.out = out.p;
      ep.out_ld = r.Cout;
      ep.out_f16 = 1;
      ep.stats_for = &out;
      std::vector<ASrc> srcs{ASrc{a2.p, a2.C, a2.C, 9, 1, H, W}};
      if (r.skip_conv) {
        for (auto& s : in) {
          if (!s.f16) { err = "resblock: the fused skip conv expects fp16 residual-stream sources"; return false; }
          srcs.push_back(ASrc{s.p, s.C, s.C, 1, 1, H, W, true});
        }
      } else {
        if (in.size() != 1 || in[0].C != r.Cout) { err = "resblock: identity skip needs a single source"; return false; }
        ep.residual = in[0].p;
        ep.res_ld = in[0].C;
        ep.res_f16 = in[0].f16 ? 1 : 0;
      }
      if (!gemm_op(ops, B * HW, true, H, W, srcs, r.conv2, ep)) return false;
    }
    return true;
  }

  // The character-context attention (<= 12 keys, 80-channel heads) runs inside the epilogue of its to_q projection when every
  // 128-row tile lies inside one sample (env WD_FUSE_CTX_ATTN=0 keeps the separate attention kernel for A/B measurements).
  static bool fuse_ctx_attn(int HW, int Ltot, int dh) {
    static int en = -1;
    if (en < 0) {
      const char* v = getenv("WD_FUSE_CTX_ATTN");
      en = v ? (atoi(v) != 0) : 1;
    }
    return en && dh == 80 && Ltot >= 1 && Ltot <= GEMM_ATT_MAXL && HW % GEMM_BLOCK_M == 0;
  }
  static void set_ctx_attn(Epi& ep, const bf16* kvbuf, int C, int Ltot, int HW, int dh, int kv_ld) {
    ep.att_kv = kvbuf;
    ep.att_ld = kv_ld ? kv_ld : 2 * C;
    ep.att_voff = C;
    ep.att_L = Ltot;
    ep.att_scale = 1.0f / sqrtf(static_cast<float>(dh));
    ep.rows_per_sample = HW;
  }

  // ---- the whole transformer block as ONE kernel (tblock.cu): unet.UNetModel, 128-token tiles inside one sample ----
  // mid form (t.mid): `g` is the fp16 residual stream after the proj_in GEMM, `out` the raw fp16 stream after the feed-forward
  bool st_block_fused(std::vector<Op>& ops, const STL& s, const TBlockL& t, const Act& x_in, const Act& g, Act& out,
                      const NormW* gn_in = nullptr) {
    const int H = x_in.H, W = x_in.W, HW = H * W, M = B * HW, C = TB_C;
    const int Ltot = plan->Ltot;
    const bool mid = t.mid;
    out = new_act(H, W, C, true);
    Op op;
    memset(&op, 0, sizeof(op));
    op.kind = OP_TBLOCK;
    TBlockArgs& a = op.tb.args;
    a.M = M;
    a.HW = HW;
    a.L = Ltot;
    a.cb = t.cb;
    a.b_ff = t.ff128.bias;
    a.cvec1 = cvec_all + static_cast<size_t>(t.fold_idx) * TB_HEADS;
    a.cvec2 = cvec_all + static_cast<size_t>(t.fold_idx + 1) * TB_HEADS;
    a.cvec_ld = cvec_ld;
    a.b_po = s.proj_out.bias;
    a.x_in = reinterpret_cast<const __half*>(x_in.p);
    a.x_in_ld = x_in.C;
    a.ln_eps = 1e-5f;  // nn.LayerNorm default (unet.py:314-316)
    a.stage = mid ? 4 : 0;
    a.mid = mid ? 1 : 0;
    a.pair = tblock_use_pair(HW) ? 1 : 0;
    if (gn_in) {
      if (!x_in.pslots || mid) { err = "fused transformer block: the input carries no GroupNorm statistics"; return false; }
      a.gn_in_partial = x_in.stats;
      a.gn_in_slots = x_in.pslots;
      a.gn_gamma = gn_in->g;
      a.gn_beta = gn_in->b;
      a.gn_eps = 1e-6f;  // SpatialTransformer.norm: Normalize(in_channels), unet.py:161-162
    }
    const int wbox = a.pair ? 80 : 160, w1box = a.pair ? TB_CHUNK : 2 * TB_CHUNK;
    if (!mid && epilogue_stats_ok(HW, C)) {
      a.gn_partial = out.stats;
      out.pslots = HW / 32;
    }
    // algorithmic work = the reference's ops (unet.py:337-345, 381-412): proj_in, 2 x (to_q, QK^T, PV, to_out), GEGLU ff, proj_out
    // (the middle form leaves proj_in / proj_out to their own GEMM ops)
    const double lin = 2.0 * M * C * C;
    op.flops = (mid ? 0.0 : 2.0 * lin) + 2.0 * (2.0 * lin + 4.0 * M * C * Ltot) + 2.0 * M * C * (8.0 * C) + 2.0 * M * (4.0 * C) * C;
    op.bytes = 2.0 * M * C * 3 + 2.0 * (3.0 * C * C + 12.0 * C * C) + 2.0 * 4 * (static_cast<double>(M) / HW) * Ltot * TB_HEADS * C;
    if (!dry) {
      bool ok = tmap_encode_2d_bf16(&op.tb.mapG, g.p, C, M, g.C, 64, TB_M) &&
                tmap_encode_2d_bf16(&op.tb.mapWpi, mid ? e->tb_identity : s.proj_in.w, C, C, C, 64, wbox) &&
                tmap_encode_2d_bf16(&op.tb.mapW1, t.ff128.w, C, 8 * C, C, 64, w1box) &&
                tmap_encode_2d_bf16(&op.tb.mapW2, t.ff_out.w, 4 * C, C, 4 * C, 64, wbox) &&
                tmap_encode_2d_bf16(&op.tb.mapWpo, s.proj_out.w, C, C, C, 64, wbox) &&
                tmap_encode_2d_bf16(&op.tb.mapOut, out.p, C, M, C, 64, TB_M);
      for (int i = 0; i < 4 && ok; ++i) {
        const bf16* base = fold_out + static_cast<size_t>(t.fold_idx + i / 2) * TB_FOLD_N + (i & 1) * (TB_HEADS * TB_C);
        ok = tmap_encode_3d_bf16(&op.tb.mapF[i], base, TB_HEADS * TB_C, Ltot, B, fold_ld, static_cast<uint64_t>(Ltot) * fold_ld, 64,
                                 TB_KEYS);
      }
      if (!ok) { err = "cuTensorMapEncodeTiled failed (fused transformer block)"; return false; }
    }
    ops.push_back(op);
    if (!mid && !out.pslots && !ensure_stats(ops, out)) return false;
    return true;
  }

  // ---- SpatialTransformer (unet.py:381-412, 337-345 ; unetPhosc.py:282-300, 241-246) ----
  bool st_block(std::vector<Op>& ops, const STL& s, const Act& x_in, const std::vector<bf16*>& kv, Act& out) {
    const int H = x_in.H, W = x_in.W, HW = H * W, M = B * HW, C = s.heads * s.dh;
    const int Ltot = plan->Ltot;
    const bool tb_ok = fold_out && s.blocks.size() == 1 && s.blocks[0].w_fold && x_in.f16 && s.heads * s.dh == TB_C;
    if (tb_ok && !s.blocks[0].mid && (HW % TB_M == 0 || HW == TB_M / 2) && x_in.C == TB_C && s.C == TB_C && tblock_gn_fused()) {
      // the SpatialTransformer's GroupNorm runs inside the fused kernel (x_in's partial statistics come from its producer)
      Act xs = x_in;
      if (!ensure_stats(ops, xs)) return false;
      return st_block_fused(ops, s, s.blocks[0], xs, xs, out, &s.gn);
    }
    Act g;
    if (!gn_op(ops, {x_in}, s.gn, 1e-6f, 0, g)) return false;
    if (tb_ok && !s.blocks[0].mid && (HW % TB_M == 0 || HW == TB_M / 2) && x_in.C == TB_C && s.C == TB_C)
      return st_block_fused(ops, s, s.blocks[0], x_in, g, out);
    if (tb_ok && s.blocks[0].mid && (HW % TB_M == 0 || HW == TB_M / 2)) {
      // channels != 320 (the 4 x 16 level: 640): proj_in and proj_out as GEMMs around the kernel's middle form
      Act x = new_act(H, W, TB_C, true), x3;
      {
        Epi ep;
        ep.out = x.p;
        ep.out_ld = TB_C;
        ep.out_f16 = 1;
        if (!gemm_op(ops, M, false, 0, 0, {ASrc{g.p, g.C, g.C, 1, 1, H, W}}, s.proj_in, ep)) return false;
      }
      if (!st_block_fused(ops, s, s.blocks[0], x_in, x, x3)) return false;
      out = new_act(H, W, s.C, true);
      Epi ep;
      ep.out = out.p;
      ep.out_ld = s.C;
      ep.out_f16 = 1;
      ep.residual = x_in.p;
      ep.res_ld = x_in.C;
      ep.res_f16 = x_in.f16 ? 1 : 0;
      ep.rows_per_sample = HW;
      ep.stats_for = &out;
      return gemm_op(ops, M, false, 0, 0, {ASrc{x3.p, TB_C, TB_C, 1, 1, H, W, true}}, s.proj_out, ep);
    }
    if (C % 80) { err = "spatial transformer: inner channels must be a multiple of 80 (LayerNorm folding)"; return false; }
    // LayerNorm is folded into the GEMMs around it: the producer of each token tensor writes per-row {sum, sum of squares}
    // (4 column blocks of 80), the consumer GEMM reads the raw fp16 tensor with gamma-scaled weights and normalises in its epilogue
    auto new_rowstats = [&]() { return A.alloc<float>(static_cast<size_t>(M) * (C / 80) * 2); };
    Act x = new_act(H, W, C, true);  // token residual stream: fp16
    float* rs_x = new_rowstats();
    {
      Epi ep;
      ep.out = x.p;
      ep.out_ld = C;
      ep.out_f16 = 1;
      ep.ln_out = rs_x;
      if (!gemm_op(ops, M, false, 0, 0, {ASrc{g.p, g.C, g.C, 1, 1, H, W}}, s.proj_in, ep)) return false;
    }
    Act o = new_act(H, W, C);
    for (size_t bi = 0; bi < s.blocks.size(); ++bi) {
      const TBlockL& t = s.blocks[bi];
      // --- attn1 ---
      Act x1 = new_act(H, W, C, true);
      float* rs_x1 = new_rowstats();
      if (e->cfg.variant == WD_VARIANT_PHOSC) {
        bf16* qkv = A.alloc<bf16>(static_cast<size_t>(M) * 3 * C);
        Epi ep;
        ep.out = qkv;
        ep.out_ld = 3 * C;
        ep.ln_stats = rs_x;  // LN1 (unetPhosc.py:241)
        ep.ln_dim = C;
        if (!gemm_op(ops, M, false, 0, 0, {ASrc{x.p, C, C, 1, 1, H, W, true}}, t.a1_q, ep)) return false;
        attn_op(ops, qkv, 3 * C, qkv + C, qkv + 2 * C, 3 * C, o.p, C, HW, HW, s.heads, s.dh);
      } else {
        Epi ep;
        ep.ln_stats = rs_x;  // unet.py:337 applies norm2 before attn1
        ep.ln_dim = C;
        if (fuse_ctx_attn(HW, Ltot, s.dh)) {
          ep.out = o.p;
          ep.out_ld = C;
          set_ctx_attn(ep, kv[t.kv1], C, Ltot, HW, s.dh, kv_ld);
          if (!gemm_op(ops, M, false, 0, 0, {ASrc{x.p, C, C, 1, 1, H, W, true}}, t.a1_q, ep)) return false;
        } else {
          bf16* q = A.alloc<bf16>(static_cast<size_t>(M) * C);
          ep.out = q;
          ep.out_ld = C;
          if (!gemm_op(ops, M, false, 0, 0, {ASrc{x.p, C, C, 1, 1, H, W, true}}, t.a1_q, ep)) return false;
          attn_op(ops, q, C, kv[t.kv1], kv[t.kv1] + C, kv_ld ? kv_ld : 2 * C, o.p, C, HW, Ltot, s.heads, s.dh);
        }
      }
      {
        Epi ep;
        ep.out = x1.p;
        ep.out_ld = C;
        ep.out_f16 = 1;
        ep.residual = x.p;
        ep.res_ld = C;
        ep.res_f16 = 1;
        ep.ln_out = rs_x1;
        if (!gemm_op(ops, M, false, 0, 0, {ASrc{o.p, C, C, 1, 1, H, W}}, t.a1_out, ep)) return false;
      }
      // --- attn2 (cross) ---
      Act x2 = new_act(H, W, C, true);
      float* rs_x2 = new_rowstats();
      {
        Epi ep;
        ep.ln_stats = rs_x1;  // LN2
        ep.ln_dim = C;
        if (fuse_ctx_attn(HW, Ltot, s.dh)) {
          ep.out = o.p;
          ep.out_ld = C;
          set_ctx_attn(ep, kv[t.kv2], C, Ltot, HW, s.dh, kv_ld);
          if (!gemm_op(ops, M, false, 0, 0, {ASrc{x1.p, C, C, 1, 1, H, W, true}}, t.a2_q, ep)) return false;
        } else {
          bf16* q = A.alloc<bf16>(static_cast<size_t>(M) * C);
          ep.out = q;
          ep.out_ld = C;
          if (!gemm_op(ops, M, false, 0, 0, {ASrc{x1.p, C, C, 1, 1, H, W, true}}, t.a2_q, ep)) return false;
          attn_op(ops, q, C, kv[t.kv2], kv[t.kv2] + C, kv_ld ? kv_ld : 2 * C, o.p, C, HW, Ltot, s.heads, s.dh);
        }
        Epi ep2;
        ep2.out = x2.p;
        ep2.out_ld = C;
        ep2.out_f16 = 1;
        ep2.residual = x1.p;
        ep2.res_ld = C;
        ep2.res_f16 = 1;
        ep2.ln_out = rs_x2;
        if (!gemm_op(ops, M, false, 0, 0, {ASrc{o.p, C, C, 1, 1, H, W}}, t.a2_out, ep2)) return false;
      }
      // --- GEGLU feed-forward ---
      Act x3 = new_act(H, W, C, true);
      float* rs_x3 = (bi + 1 < s.blocks.size()) ? new_rowstats() : nullptr;  // only a further transformer block normalises x3
      {
        bf16* gg = A.alloc<bf16>(static_cast<size_t>(M) * 4 * C);
        Epi ep;
        ep.out = gg;
        ep.out_ld = 4 * C;
        ep.geglu = 1;
        ep.ln_stats = rs_x2;  // LN3
        ep.ln_dim = C;
        if (!gemm_op(ops, M, false, 0, 0, {ASrc{x2.p, C, C, 1, 1, H, W, true}}, t.ff_proj, ep)) return false;
        Epi ep2;
        ep2.out = x3.p;
        ep2.out_ld = C;
        ep2.out_f16 = 1;
        ep2.residual = x2.p;
        ep2.res_ld = C;
        ep2.res_f16 = 1;
        ep2.ln_out = rs_x3;
        if (!gemm_op(ops, M, false, 0, 0, {ASrc{gg, 4 * C, 4 * C, 1, 1, H, W}}, t.ff_out, ep2)) return false;
      }
      x = x3;
      rs_x = rs_x3;
    }
    out = new_act(H, W, s.C, true);
    Epi ep;
    ep.out = out.p;
    ep.out_ld = s.C;
    ep.out_f16 = 1;
    ep.residual = x_in.p;
    ep.res_ld = x_in.C;
    ep.res_f16 = x_in.f16 ? 1 : 0;
    ep.rows_per_sample = HW;
    ep.stats_for = &out;
    if (!x.f16) { err = "spatial transformer: proj_out expects the fp16 token stream"; return false; }
    return gemm_op(ops, M, false, 0, 0, {ASrc{x.p, C, C, 1, 1, H, W, true}}, s.proj_out, ep);
  }

  bool build() {
    const wd_config& c = e->cfg;
    const int mc = c.model_channels, ted = e->time_dim, D = c.context_dim;
    const int L = plan->L, Ltot = plan->Ltot;
    auto& cops = plan->ctx_ops;
    auto& sops = plan->step_ops;

    // ================= context (time-invariant) =================
    bf16* ctx = A.alloc<bf16>(static_cast<size_t>(B) * Ltot * D);
    plan->ctx_buf = ctx;
    {
      const int nseg = c.phosc_len > 0 ? 2 : 1;
      for (int seg = 0; seg < nseg; ++seg) {
        const int Ls = seg == 0 ? L : c.phosc_len;
        const int row_off = seg == 0 ? 0 : L;
        // unet.py:872 always adds the PE; unetPhosc.py:726-729 only when the sequence fits max_seq_len
        const int add_pe = (c.variant == WD_VARIANT_UNET) ? 1 : (Ls <= c.max_seq_len ? 1 : 0);
        if (add_pe && Ls > c.max_seq_len) { err = "context longer than max_seq_len"; return false; }
        if (!add_pe) {
          // no positional term: the attention output depends only on each token's value and the sample's token histogram
          Op wh;
          memset(&wh, 0, sizeof(wh));
          wh.kind = OP_WORDATTN;
          wh.wa = {nullptr, nullptr, nullptr, ctx, B, Ls, D, Ltot, row_off};
          wh.emb.which = seg;  // which token tensor (0: characters, 1: PHOSC)
          wh.flops = 1;        // marks the histogram form
          cops.push_back(wh);
          continue;
        }
        float* q = A.alloc<float>(static_cast<size_t>(B) * Ls * D);
        float* k = A.alloc<float>(static_cast<size_t>(B) * Ls * D);
        float* v = A.alloc<float>(static_cast<size_t>(B) * Ls * D);
        // q / k / v of Word_Attention straight from the folded tables (wd_engine_finalize_params): gather + add
        const float* Ts[3] = {e->we_tq, e->we_tk, e->we_tv};
        const float* Ps[3] = {e->we_pq, e->we_pk, e->we_pv};
        float* outs[3] = {q, k, v};
        for (int i = 0; i < 3; ++i) {
          Op op;
          memset(&op, 0, sizeof(op));
          op.kind = OP_EMBED;
          op.emb = {seg, Ts[i], c.vocab_size, Ps[i], add_pe, outs[i], B, Ls, D};
          cops.push_back(op);
        }
        Op wa;
        memset(&wa, 0, sizeof(wa));
        wa.kind = OP_WORDATTN;
        wa.wa = {q, k, v, ctx, B, Ls, D, Ltot, row_off};
        cops.push_back(wa);
      }
    }
    std::vector<bf16*> kv(e->n_kv, nullptr);
    kv_ld = 0;
    bool pooled = e->n_kv > 0;
    for (int i = 0; i < e->n_kv; ++i) {
      const GemmW& w = *e->kv_weights[i];
      const GemmW& w0 = *e->kv_weights[0];
      pooled = pooled && !w.bias && w.N == w0.N && w.K == w0.K && w.w == w0.w + static_cast<size_t>(i) * w0.N * w0.K;
    }
    if (pooled) {
      // the [to_k; to_v] weights of every cross-attention are contiguous: ONE GEMM [B*Ltot, D] x [n_kv*2C, D]^T per trajectory
      const GemmW& w0 = *e->kv_weights[0];
      GemmW wall;
      wall.w = w0.w;
      wall.N = w0.N * e->n_kv;
      wall.K = w0.K;
      kv_ld = wall.N;
      bf16* all = A.alloc<bf16>(static_cast<size_t>(B) * Ltot * wall.N);
      for (int i = 0; i < e->n_kv; ++i) kv[i] = all + static_cast<size_t>(i) * w0.N;
      Epi ep;
      ep.out = all;
      ep.out_ld = wall.N;
      if (!gemm_op(cops, B * Ltot, false, 0, 0, {ASrc{ctx, D, D, 1, 1, 1, 1}}, wall, ep)) return false;
    } else {
      for (int i = 0; i < e->n_kv; ++i) {
        const GemmW& w = *e->kv_weights[i];
        kv[i] = A.alloc<bf16>(static_cast<size_t>(B) * Ltot * w.N);
        Epi ep;
        ep.out = kv[i];
        ep.out_ld = w.N;
        if (!gemm_op(cops, B * Ltot, false, 0, 0, {ASrc{ctx, D, D, 1, 1, 1, 1}}, w, ep)) return false;
      }
    }

    // fused transformer blocks: every per-sample attention operand (M = Wq^T K^T and N = V Wout^T of each cross-attention) from
    // ONE GEMM ctx [B L, D] x W_fold^T, plus the score constants; unet.UNetModel with a context of at most 16 tokens
    fold_out = nullptr;
    if (tblock_enabled() && e->fold_pool && e->fold_count > 0 && Ltot <= TB_KEYS && D == TB_C) {
      GemmW wall;
      wall.w = e->fold_pool;
      wall.N = e->fold_count * TB_FOLD_N;
      wall.K = TB_C;
      fold_ld = wall.N;
      fold_out = A.alloc<bf16>(static_cast<size_t>(B) * Ltot * wall.N);
      Epi ep;
      ep.out = fold_out;
      ep.out_ld = wall.N;
      ep.out_f16 = 1;
      if (!gemm_op(cops, B * Ltot, false, 0, 0, {ASrc{ctx, D, D, 1, 1, 1, 1}}, wall, ep)) return false;
      cvec_ld = e->fold_count * TB_HEADS;
      cvec_all = A.alloc<float>(static_cast<size_t>(B) * Ltot * cvec_ld);
      Op cv;
      memset(&cv, 0, sizeof(cv));
      cv.kind = OP_TB_CVEC;
      cv.cv = {ctx, e->u_pool, cvec_all, B * Ltot, cvec_ld};
      cops.push_back(cv);
    }

    // ================= per-step =================
    bf16* temb = A.alloc<bf16>(static_cast<size_t>(B) * mc);
    bf16* h1 = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    bf16* emb_act = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    float* emb_out = A.alloc<float>(static_cast<size_t>(B) * e->emb_all.N);
    {
      float* table = e->temb_table;  // time_embed(t) for every timestep, built at weight load (wd_engine_finalize_params)
      {
        Op op;
        memset(&op, 0, sizeof(op));
        op.kind = OP_EMBTBL;
        op.cond = COND_TABLE;
        op.etbl = {table, emb_act, B, ted, (c.num_classes > 0 && c.add_label_emb) ? 1 : 0};
        op.bytes = 6.0 * B * ted;
        sops.push_back(op);
      }
      const size_t first_general = sops.size();
      Op op;
      memset(&op, 0, sizeof(op));
      op.kind = OP_TEMB;
      op.temb = {temb, B, mc, 0};
      sops.push_back(op);
      Epi ep;
      ep.out = h1;
      ep.out_ld = ted;
      ep.act = ACT_SILU;
      if (!gemm_op(sops, B, false, 0, 0, {ASrc{temb, mc, mc, 1, 1, 1, 1}}, e->te0, ep)) return false;
      Epi ep2;
      ep2.out = emb_act;
      ep2.out_ld = ted;
      ep2.act = ACT_SILU;  // every consumer applies SiLU first (emb_layers.0, unet.py:609-610)
      int patch = P_NONE;
      if (c.num_classes > 0 && c.add_label_emb) {
        ep2.rowbias = e->label_emb;
        ep2.rb_ld = ted;
        ep2.rows_per_sample = 1;
        patch = P_Y;
      }
      if (!gemm_op(sops, B, false, 0, 0, {ASrc{h1, ted, ted, 1, 1, 1, 1}}, e->te2, ep2, patch)) return false;
      for (size_t i = first_general; i < sops.size(); ++i) sops[i].cond = COND_NO_TABLE;
      Epi ep3;
      ep3.out = emb_out;
      ep3.out_ld = e->emb_all.N;
      ep3.out_f32 = 1;
      if (!gemm_op(sops, B, false, 0, 0, {ASrc{emb_act, ted, ted, 1, 1, 1, 1}}, e->emb_all, ep3)) return false;
    }
    const int emb_ld = e->emb_all.N;

    std::vector<Act> hs;
    Act h{nullptr, 0, 0, 0};
    auto run_block = [&](const Block& blk, std::vector<Act> in) -> bool {
      for (const Layer& l : blk) {
        Act out;
        switch (l.kind) {
          case L_CONVIN: {
            // conv_in = im2col (hi/lo bf16 split of the fp32 latent) + tcgen05 GEMM with K = 128
            out = new_act(c.latent_h, c.latent_w, mc, true);
            const int HW = c.latent_h * c.latent_w;
            bf16* col = A.alloc<bf16>(static_cast<size_t>(B) * HW * 128);
            Op op;
            memset(&op, 0, sizeof(op));
            op.kind = OP_CONV_IN;
            op.cin = {col, B, c.latent_h, c.latent_w};
            op.bytes = static_cast<double>(B) * HW * (4 * 4 + 2 * 128);
            sops.push_back(op);
            Epi ep;
            ep.out = out.p;
            ep.out_ld = mc;
            ep.out_f16 = 1;
            ep.rows_per_sample = HW;
            ep.stats_for = &out;
            if (!gemm_op(sops, B * HW, false, 0, 0, {ASrc{col, 128, 128, 1, 1, 1, 1}}, e->conv_in, ep)) return false;
            for (auto it = sops.rbegin(); it != sops.rend(); ++it)
              if (it->kind == OP_GEMM) {
                it->flops = 2.0 * B * HW * mc * 36;  // algorithmic work of the 3x3x4 convolution, not of the padded K
                break;
              }
            break;
          }
          case L_RES:
            if (!res_block(sops, e->res[l.idx], in, emb_out, emb_ld, out)) return false;
            break;
          case L_ST:
            if (!st_block(sops, e->st[l.idx], in[0], kv, out)) return false;
            break;
          case L_DOWN: {
            const Act& x = in[0];
            if (x.H % 2 || x.W % 2) { err = "downsample needs even spatial size"; return false; }
            if (!x.f16) { err = "downsample expects an fp16 residual-stream input"; return false; }
            out = new_act(x.H / 2, x.W / 2, x.C, true);
            Epi ep;
            ep.out = out.p;
            ep.out_ld = x.C;
            ep.out_f16 = 1;
            ep.rows_per_sample = out.H * out.W;
            ep.stats_for = &out;
            if (!gemm_op(sops, B * out.H * out.W, true, out.H, out.W, {ASrc{x.p, x.C, x.C, 9, 2, x.H, x.W, true}},
                         e->samp[l.idx].conv, ep))
              return false;
            break;
          }
          case L_UP: {
            const Act& x = in[0];
            if (!x.f16) { err = "upsample expects an fp16 residual-stream input"; return false; }
            const SampL& sp = e->samp[l.idx];
            if (up_phase_enabled() && sp.w_phase && GEMM_BLOCK_M % (x.H * x.W) == 0 && epilogue_stats_ok(x.H * x.W, x.C) &&
                (4 * x.H * x.W) / 32 <= 8) {
              // sub-pixel form (GemmArgs::up_phase): four 2 x 2 convolutions of the H x W input, one per output phase, 9/4 fewer
              // MACs than the 3 x 3 conv on the upsampled tensor, which is never materialised
              out = new_act(x.H * 2, x.W * 2, x.C, true);
              if (up_phase_mode() >= 2 && GEMM_BLOCK_M % x.W == 0 && x.C == 320) {
                // all four phases in ONE launch: the phases are N tiles of a [4 C, 4 C] weight (1 024 tiles at batch 256 instead of
                // four launches of 256 tiles each)
                GemmW wp = sp.conv;
                wp.w = sp.w_phase;
                wp.N = 4 * x.C;
                wp.K = 4 * x.C;
                wp.bias = sp.bias4;
                Epi ep;
                ep.out = out.p;
                ep.out_ld = x.C;
                ep.out_f16 = 1;
                ep.rows_per_sample = x.H * x.W;
                ep.stats_for = &out;
                ep.up_phase = 5;
                if (!gemm_op(sops, B * x.H * x.W, true, x.H, x.W, {ASrc{x.p, x.C, x.C, 4, 1, x.H, x.W, true}}, wp, ep)) return false;
                sops.back().flops = 2.0 * B * 4 * x.H * x.W * x.C * 9.0 * x.C;  // the reference's conv on the upsampled tensor
                break;
              }
              for (int ph = 0; ph < 4; ++ph) {
                GemmW wp = sp.conv;
                wp.w = sp.w_phase + static_cast<size_t>(ph) * x.C * 4 * x.C;
                wp.K = 4 * x.C;
                Epi ep;
                ep.out = out.p;
                ep.out_ld = x.C;
                ep.out_f16 = 1;
                ep.rows_per_sample = x.H * x.W;
                ep.stats_for = &out;
                ep.up_phase = ph + 1;
                if (!gemm_op(sops, B * x.H * x.W, true, x.H, x.W, {ASrc{x.p, x.C, x.C, 4, 1, x.H, x.W, true}}, wp, ep)) return false;
                // algorithmic work = the reference's 3 x 3 conv on the upsampled tensor (a quarter of it per phase); 4/9 of it is executed
                if (!dry) sops.back().flops = 2.0 * B * x.H * x.W * x.C * 9.0 * x.C;
              }
              break;
            }
            Act up = new_act(x.H * 2, x.W * 2, x.C, true);  // 16-bit copy, format-agnostic
            Op op;
            memset(&op, 0, sizeof(op));
            op.kind = OP_UPSAMPLE;
            op.up = {x.p, up.p, B, x.H, x.W, x.C};
            op.bytes = 2.0 * B * x.H * x.W * x.C * 5;
            sops.push_back(op);
            out = new_act(up.H, up.W, x.C, true);
            Epi ep;
            ep.out = out.p;
            ep.out_ld = x.C;
            ep.out_f16 = 1;
            ep.rows_per_sample = up.H * up.W;
            ep.stats_for = &out;
            if (!gemm_op(sops, B * up.H * up.W, true, up.H, up.W, {ASrc{up.p, up.C, up.C, 9, 1, up.H, up.W, true}},
                         e->samp[l.idx].conv, ep))
              return false;
            break;
          }
        }
        in = {out};
        h = out;
      }
      return true;
    };

    for (auto& blk : e->input_blocks) {
      if (!run_block(blk, h.p ? std::vector<Act>{h} : std::vector<Act>{})) return false;
      hs.push_back(h);
    }
    if (!run_block(e->middle, {h})) return false;
    for (auto& blk : e->output_blocks) {
      Act skip = hs.back();
      hs.pop_back();
      if (skip.H != h.H || skip.W != h.W) { err = "skip connection spatial mismatch"; return false; }
      if (!run_block(blk, {h, skip})) return false;
    }
    // out: GN + SiLU + conv_out + sampler update as ONE kernel with the sample's image in shared memory (ops.cuh: OutHeadArgs)
    if (out_head_enabled() && h.f16 && h.pslots > 0 && h.C % 32 == 0 && c.out_channels == 4 && out_head_supported(h.H, h.W, h.C)) {
      Op op;
      memset(&op, 0, sizeof(op));
      op.kind = OP_OUTHEAD;
      op.oh.h = reinterpret_cast<const __half*>(h.p);
      op.oh.partial = h.stats;
      op.oh.pslots = h.pslots;
      op.oh.gamma = e->out_gn.g;
      op.oh.beta = e->out_gn.b;
      op.oh.gn_eps = 1e-5f;
      op.oh.w = e->conv_out.w;
      op.oh.bias = e->conv_out.bias;
      op.oh.B = B;
      op.oh.H = h.H;
      op.oh.W = h.W;
      op.oh.C = h.C;
      op.bytes = static_cast<double>(B) * h.H * h.W * (2.0 * h.C + 4.0 * c.out_channels * 4);  // h read; x in/out, noise, eps
      op.flops = 2.0 * B * h.H * h.W * h.C * 9 * c.out_channels;
      sops.push_back(op);
      plan->bytes = A.used;
      return true;
    }
    // fallback: GN + SiLU kernel, then conv_out on the tensor cores fused with the sampler update
    Act a;
    if (!gn_op(sops, {h}, e->out_gn, 1e-5f, 1, a)) return false;
    {
      // output conv 320 -> 4 on the tensor cores (16-column tile), epilogue = sampler update; x / eps / noise /
      // coefficients are patched in per call (P_SAMPLER)
      Epi ep;
      ep.epi = EPI_SAMPLER;
      ep.rows_per_sample = a.H * a.W;
      if (!gemm_op(sops, B * a.H * a.W, true, a.H, a.W, {ASrc{a.p, a.C, a.C, 9, 1, a.H, a.W}}, e->conv_out, ep, P_SAMPLER))
        return false;
      sops.back().bytes = static_cast<double>(B) * a.H * a.W * (2.0 * a.C + 4.0 * c.out_channels * 4);  // h read; x in/out, noise, eps
      sops.back().flops = 2.0 * B * a.H * a.W * a.C * 9 * c.out_channels;
    }
    plan->bytes = A.used;
    return true;
  }
};

int ensure_plan(wd_engine* e, int B, int L, Plan** out) {
  if (B < 1) return fail(WD_ERR_INVALID, "batch must be >= 1");
  if (L < 1) return fail(WD_ERR_INVALID, "context length must be >= 1");
  auto key = std::make_pair(B, L);
  auto it = e->plans.find(key);
  if (it != e->plans.end()) {
    *out = it->second.get();
    return WD_OK;
  }
  std::unique_ptr<Plan> p(new Plan());
  p->B = B;
  p->L = L;
  p->Ltot = L + (e->cfg.phosc_len > 0 ? e->cfg.phosc_len : 0);
  {
    Plan tmp = *p;
    PlanBuilder pb{e, &tmp, Arena(), true, B, ""};
    if (!pb.build()) return fail(WD_ERR_UNSUPPORTED, "plan: %s", pb.err.c_str());
    p->bytes = tmp.bytes;
  }
  if (p->bytes > e->acap) {
    CUDA_TRY(cudaDeviceSynchronize());
    if (e->abase) CUDA_TRY(cudaFree(e->abase));
    e->abase = nullptr;
    e->acap = 0;
    e->plans.clear();
    e->cur = nullptr;
    CUDA_TRY(cudaMalloc(&e->abase, p->bytes));
    e->acap = p->bytes;
  }
  Arena A;
  A.base = e->abase;
  PlanBuilder pb{e, p.get(), A, false, B, ""};
  if (!pb.build()) return fail(WD_ERR_UNSUPPORTED, "plan: %s", pb.err.c_str());
  *out = p.get();
  e->plans[key] = std::move(p);
  return WD_OK;
}

struct RunCtx {
  const StepParams* sp = nullptr;  // graph capture / replay: per-step scalars are read from device memory
  const float* x = nullptr;
  const long long* t_dev = nullptr;
  long long t_scalar = 0;
  const long long* y = nullptr;
  const long long* ctx_tokens = nullptr;
  const int* phosc = nullptr;
  // conv_out
  float* eps_out = nullptr;
  float* x_rw = nullptr;
  const float* noise = nullptr;
  int use_philox = 0;
  unsigned long long seed = 0, sample_offset = 0;
  int step_index = 0;
  float4 coef = make_float4(0, 0, 0, 0);
  int mode = STEP_EPS_ONLY;
};

static bool step_graph_enabled() {  // env WD_STEP_GRAPH (default on)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_STEP_GRAPH");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
static bool temb_table_enabled() {  // env WD_TEMB_TABLE (default on)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_TEMB_TABLE");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

static int engine_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int run_ops(wd_engine* e, const std::vector<Op>& ops, const RunCtx& rc, cudaStream_t s) {
  int n = 0, kernels = 0;
  std::vector<cudaEvent_t>* ev = nullptr;
  if (e->prof_on && &ops == &e->cur->step_ops && e->prof_steps.size() < 256) {
    if (e->prof_plan != e->cur) {
      for (auto& v : e->prof_steps)
        for (auto x : v) cudaEventDestroy(x);
      e->prof_steps.clear();
      e->prof_plan = e->cur;
    }
    e->prof_steps.emplace_back(ops.size() + 1);
    ev = &e->prof_steps.back();
    for (auto& x : *ev)
      if (cudaEventCreate(&x) != cudaSuccess) return fail(WD_ERR_CUDA, "cudaEventCreate");
  }
  const bool use_table = rc.t_dev == nullptr && rc.t_scalar >= 0 && rc.t_scalar < TEMB_TABLE_ROWS && temb_table_enabled();
  if (ev) e->prof_used_table = use_table;
  for (const Op& op : ops) {
    cudaError_t err = cudaSuccess;
    if (ev) cudaEventRecord((*ev)[n], s);
    if ((op.cond == COND_TABLE && !use_table) || (op.cond == COND_NO_TABLE && use_table)) {
      ++n;
      continue;
    }
    switch (op.kind) {
      case OP_TEMB:
        if (op.temb.table) err = timestep_embed_launch(nullptr, -1, op.temb.out, op.temb.B, op.temb.dim, s);
        else err = timestep_embed_launch(rc.t_dev, rc.t_scalar, op.temb.out, op.temb.B, op.temb.dim, s);
        break;
      case OP_EMBTBL:
        if (op.etbl.use_label && !rc.y) return fail(WD_ERR_INVALID, "y (writer ids) is required by this model");
        err = emb_from_table_launch(op.etbl.table, rc.t_scalar, rc.sp, op.etbl.use_label ? e->label_emb : nullptr, rc.y, op.etbl.out,
                                    op.etbl.B, op.etbl.dim, s);
        break;
      case OP_GEMM:
        if (op.patch == P_Y) {
          if (!rc.y) return fail(WD_ERR_INVALID, "y (writer ids) is required by this model");
          GemmLaunch L = op.gemm;
          L.args.rowbias_idx = rc.y;
          err = gemm_tc_launch(L, s);
        } else if (op.patch == P_SAMPLER) {
          GemmLaunch L = op.gemm;
          L.args.eps_out = rc.eps_out;
          L.args.x = rc.x_rw;
          L.args.noise = rc.noise;
          L.args.use_philox = rc.use_philox;
          L.args.seed = rc.seed;
          L.args.sample_offset = rc.sample_offset;
          L.args.step_index = rc.step_index;
          L.args.coef = rc.coef;
          L.args.mode = rc.mode;
          L.args.sp = rc.sp;
          err = gemm_tc_launch(L, s);
        } else {
          err = gemm_tc_launch(op.gemm, s);
        }
        break;
      case OP_GN:
        err = groupnorm_launch(op.gn, op.gn_B, op.gn_nslab, s);
        break;
      case OP_LN:
        err = layernorm_launch(op.ln.x, op.ln.out, op.ln.g, op.ln.b, op.ln.M, op.ln.C, op.ln.eps, op.ln.x_f16, s);
        break;
      case OP_ATTN_SMALL:
        err = attn_small_launch(op.as, e->cur->B, s);
        break;
      case OP_ATTN_FLASH:
        err = attn_flash_launch(op.af, e->cur->B, s);
        break;
      case OP_CONV_IN:
        err = conv_in_im2col_launch(rc.x, op.cin.out, op.cin.B, op.cin.H, op.cin.W, s);
        break;
      case OP_GNSTATS:
        err = groupnorm_stats_launch(op.gs, op.gn_B, s);
        break;
      case OP_UPSAMPLE:
        err = upsample2x_launch(op.up.x, op.up.out, op.up.B, op.up.H, op.up.W, op.up.C, s);
        break;
      case OP_EMBED:
        if (op.emb.which == 0)
          err = embed_tokens_launch(rc.ctx_tokens, 1, op.emb.E, op.emb.vocab, op.emb.pe, op.emb.add_pe, op.emb.out,
                                    op.emb.B, op.emb.L, op.emb.D, s);
        else {
          if (!rc.phosc) return fail(WD_ERR_INVALID, "phoscLabels are required by this model");
          err = embed_tokens_launch(rc.phosc, 0, op.emb.E, op.emb.vocab, op.emb.pe, op.emb.add_pe, op.emb.out, op.emb.B,
                                    op.emb.L, op.emb.D, s);
        }
        break;
      case OP_TBLOCK:
        err = tblock_launch(op.tb, engine_num_sms(), s);
        break;
      case OP_TB_CVEC:
        err = tblock_cvec_launch(op.cv.ctx, op.cv.u, op.cv.out, op.cv.rows, op.cv.heads, s);
        break;
      case OP_OUTHEAD: {
        OutHeadArgs oa = op.oh;
        oa.eps_out = rc.eps_out;
        oa.x = rc.x_rw;
        oa.noise = rc.noise;
        oa.use_philox = rc.use_philox;
        oa.seed = rc.seed;
        oa.sample_offset = rc.sample_offset;
        oa.step_index = rc.step_index;
        oa.coef = rc.coef;
        oa.mode = rc.mode;
        oa.sp = rc.sp;
        err = out_head_launch(oa, s);
        break;
      }
      case OP_LINF32:
        err = linear_f32_launch(op.lin.x, op.lin.W, op.lin.b, op.lin.out, op.lin.M, op.lin.N, op.lin.K, s);
        break;
      case OP_WORDATTN:
        if (op.wa.q == nullptr) {  // histogram form (position-free segment)
          const void* toks = op.emb.which == 0 ? static_cast<const void*>(rc.ctx_tokens) : static_cast<const void*>(rc.phosc);
          if (!toks) return fail(WD_ERR_INVALID, "phoscLabels are required by this model");
          err = word_attn_hist_launch(toks, op.emb.which == 0 ? 1 : 0, e->we_gram, e->we_tv, e->cfg.vocab_size, op.wa.ctx, op.wa.B, op.wa.L,
                                      op.wa.D, op.wa.Ltot, op.wa.row_off, s);
          break;
        }
        err = word_attn_launch(op.wa.q, op.wa.k, op.wa.v, op.wa.ctx, nullptr, op.wa.B, op.wa.L, op.wa.D, op.wa.Ltot,
                               op.wa.row_off, s);
        break;
    }
    if (err != cudaSuccess) return fail(WD_ERR_CUDA, "launch of op kind %d failed: %s", (int)op.kind, cudaGetErrorString(err));
    ++n;
    kernels += 1;
  }
  if (ev) cudaEventRecord((*ev)[n], s);
  e->last_launches = kernels;
  return WD_OK;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// C ABI: hot path
// ----------------------------------------------------------------------------------------------
extern "C" int wd_engine_reserve(wd_engine* e, int batch) {
  if (!e) return fail(WD_ERR_INVALID, "null engine");
  Plan* p = nullptr;
  return ensure_plan(e, batch, e->cfg.max_seq_len, &p);
}

extern "C" int wd_encode_context(wd_engine* e, int batch, const int64_t* ctx_tokens, int L, const int32_t* phosc,
                                 void* stream) {
  if (!e || !ctx_tokens) return fail(WD_ERR_INVALID, "null argument");
  Plan* p = nullptr;
  int rc = ensure_plan(e, batch, L, &p);
  if (rc) return rc;
  e->cur = p;
  RunCtx r;
  r.ctx_tokens = reinterpret_cast<const long long*>(ctx_tokens);
  r.phosc = phosc;
  rc = run_ops(e, p->ctx_ops, r, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  p->context_valid = true;
  return WD_OK;
}

// The context given directly as a dense fp32 tensor [batch, L, context_dim] instead of character tokens: the reference's
// args.wrdChrWrStyl == 1 path replaces word_emb(context) by wrd_proj(wrdChrWrStyl) (unet.py:1590-1591,1617-1618).  Runs the
// time-invariant part that follows the context encoder (K/V projections, fused-block operands).
extern "C" int wd_set_context(wd_engine* e, int batch, const float* ctx_f32, int L, void* stream) {
  if (!e || !ctx_f32) return fail(WD_ERR_INVALID, "null argument");
  if (e->cfg.phosc_len > 0) return fail(WD_ERR_UNSUPPORTED, "wd_set_context: the PHOSC variants build their context from tokens");
  Plan* p = nullptr;
  int rc = ensure_plan(e, batch, L, &p);
  if (rc) return rc;
  e->cur = p;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUDA_TRY(repack_linear_launch(ctx_f32, p->ctx_buf, batch * p->Ltot, e->cfg.context_dim, e->cfg.context_dim, 0, 0, 0, 0, s));
  std::vector<Op> rest;
  for (const Op& op : p->ctx_ops)
    if (op.kind != OP_EMBED && op.kind != OP_WORDATTN) rest.push_back(op);
  RunCtx r;
  rc = run_ops(e, rest, r, s);
  if (rc) return rc;
  p->context_valid = true;
  return WD_OK;
}

// label_emb.weight[row] = (1 - mix) * label_emb.weight[s1] + mix * label_emb.weight[s2]: the style interpolation of
// args.interpolation (unet.py:1558-1572) with the reference's rounding (two products, one sum).  `row` is a scratch row of the
// table (the drop-in module creates the engine with one more class than the model has).
extern "C" int wd_engine_set_label_mix(wd_engine* e, int row, int s1, int s2, float mix, void* stream) {
  if (!e || !e->label_emb) return fail(WD_ERR_STATE, "wd_engine_set_label_mix: the model has no label embedding");
  const int n = e->cfg.num_classes;
  if (row < 0 || row >= n || s1 < 0 || s1 >= n || s2 < 0 || s2 >= n) return fail(WD_ERR_INVALID, "wd_engine_set_label_mix: class out of range");
  CUDA_TRY(label_mix_launch(e->label_emb, e->time_dim, row, s1, s2, mix, static_cast<cudaStream_t>(stream)));
  if (e->cur) {  // captured graphs read the table by pointer: nothing to refresh
  }
  return WD_OK;
}

// Training-step front and back ends (train.py:190-194,287): wd_noise_images writes x_t and eps in one pass (eps from
// `eps_in`, or Philox keyed by (seed, sample_offset, stream_id) when eps_in is NULL); wd_mse_loss_grad writes the MSE loss (device
// scalar) and d loss / d pred.  `workspace`: wd_mse_workspace_bytes(n) bytes, zeroed once by the caller.
extern "C" int wd_noise_images(const float* x, const int64_t* t, const float* alpha_hat, int T, const float* eps_in, uint64_t seed,
                               uint64_t sample_offset, uint32_t stream_id, float* x_t, float* eps_out, int batch, int elems_per_latent,
                               void* stream) {
  if (!x || !t || !alpha_hat || !x_t || !eps_out || batch < 0 || elems_per_latent < 1) return fail(WD_ERR_INVALID, "wd_noise_images: invalid argument");
  CUDA_TRY(noise_images_launch(x, reinterpret_cast<const long long*>(t), alpha_hat, T, eps_in, seed,
                               sample_offset * static_cast<uint64_t>(elems_per_latent), stream_id, x_t, eps_out,
                               static_cast<size_t>(batch) * elems_per_latent, elems_per_latent, nullptr, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}
extern "C" size_t wd_mse_workspace_bytes(size_t n) { return mse_grad_workspace_bytes(n); }
extern "C" int wd_mse_loss_grad(const float* pred, const float* target, float* d_pred, float* loss, void* workspace, size_t n,
                                void* stream) {
  if (!pred || !target || !d_pred || !loss || !workspace || !n) return fail(WD_ERR_INVALID, "wd_mse_loss_grad: invalid argument");
  CUDA_TRY(mse_grad_launch(pred, target, d_pred, loss, workspace, n, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

// out = torch.lerp(start, end, weight) on fp32 device tensors (train.py:228: predicted_noise = lerp(uncond, cond, cfg_scale))
extern "C" int wd_lerp(const float* start, const float* end, float weight, float* out, size_t n, void* stream) {
  if (!start || !end || !out) return fail(WD_ERR_INVALID, "wd_lerp: null argument");
  if (n) CUDA_TRY(lerp_launch(start, end, weight, out, n, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

static int check_ready(wd_engine* e, int batch) {
  if (!e) return fail(WD_ERR_INVALID, "null engine");
  if (!e->cur || !e->cur->context_valid) return fail(WD_ERR_STATE, "wd_encode_context must run before the UNet");
  if (e->cur->B != batch) return fail(WD_ERR_STATE, "batch %d does not match the encoded context (%d)", batch, e->cur->B);
  return WD_OK;
}

// Step launch sequence through a CUDA graph: the first call with a set of caller pointers runs eagerly (and warms every kernel
// up), the second captures the sequence on a private stream (the caller's may be the legacy default stream, which cannot be
// captured; nothing runs during capture) and instantiates it, later ones replay it on the caller's stream.  `sp`: per-step
// scalars, refreshed in device memory by a one-thread kernel (by-value argument: no host staging buffer to keep alive) before
// every replay.  Any failure to capture leaves the slot in eager mode.
static int run_graphed(wd_engine* e, Plan* p, Plan::GraphSlot& g, RunCtx& r, const void* const key[4], const StepParams* sp,
                       cudaStream_t s) {
  if (memcmp(key, g.key, sizeof(g.key)) != 0) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
    memcpy(g.key, key, sizeof(g.key));
    g.seen = 0;
  }
  if (++g.seen == 1 || g.seen < 0) return run_ops(e, p->step_ops, r, s);
  if (sp) {
    if (!e->sp_dev) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&e->sp_dev), sizeof(StepParams)));
    CUDA_TRY(set_step_params_launch(e->sp_dev, *sp, s));
  }
  if (!g.exec) {
    if (sp) r.sp = e->sp_dev;
    cudaGraph_t graph = nullptr;
    cudaError_t ie = cudaErrorUnknown;
    if (!e->cap_stream && cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking) != cudaSuccess) e->cap_stream = nullptr;
    if (e->cap_stream && cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int rc2 = run_ops(e, p->step_ops, r, e->cap_stream);
      const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
      if (rc2 == WD_OK && ce == cudaSuccess && graph) {
        g.launches = e->last_launches;
        ie = cudaGraphInstantiate(&g.exec, graph, 0);
      }
    }
    if (graph) cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {  // capture is not available here: stay eager for these pointers
      g.exec = nullptr;
      g.seen = -1000000000;
      cudaGetLastError();
      r.sp = nullptr;
      return run_ops(e, p->step_ops, r, s);
    }
  }
  CUDA_TRY(cudaGraphLaunch(g.exec, s));
  e->last_launches = g.launches;
  return WD_OK;
}

extern "C" int wd_unet_eval(wd_engine* e, int batch, const float* x, const int64_t* timesteps, int64_t t_scalar,
                            const int64_t* y, float* eps_out, void* stream) {
  int rc = check_ready(e, batch);
  if (rc) return rc;
  if (!x || !eps_out) return fail(WD_ERR_INVALID, "null argument");
  RunCtx r;
  r.x = x;
  r.t_dev = reinterpret_cast<const long long*>(timesteps);
  r.t_scalar = t_scalar;
  r.y = reinterpret_cast<const long long*>(y);
  r.eps_out = eps_out;
  r.mode = STEP_EPS_ONLY;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // per-row timesteps (the forward() path): every argument of the sequence is a pointer, so the graph needs no StepParams
  if (timesteps && step_graph_enabled() && !e->prof_on) {
    const void* key[4] = {x, eps_out, timesteps, y};
    return run_graphed(e, e->cur, e->cur->g_eval, r, key, nullptr, s);
  }
  return run_ops(e, e->cur->step_ops, r, s);
}

extern "C" int wd_sampler_step(wd_engine* e, int batch, float* x, int64_t t_scalar, const int64_t* y, int mode,
                               const float* coef4_host, const float* noise, int use_philox, uint64_t seed,
                               uint64_t sample_offset, int step_index, float* eps_out, void* stream) {
  int rc = check_ready(e, batch);
  if (rc) return rc;
  if (!x || !coef4_host) return fail(WD_ERR_INVALID, "null argument");
  if (mode != WD_STEP_DDPM && mode != WD_STEP_DDIM && mode != WD_STEP_EPS_ONLY) return fail(WD_ERR_INVALID, "bad mode");
  RunCtx r;
  r.x = x;
  r.x_rw = x;
  r.t_scalar = t_scalar;
  r.y = reinterpret_cast<const long long*>(y);
  r.eps_out = eps_out;
  r.noise = noise;
  r.use_philox = use_philox;
  r.seed = seed;
  r.sample_offset = sample_offset;
  r.step_index = step_index;
  r.coef = make_float4(coef4_host[0], coef4_host[1], coef4_host[2], coef4_host[3]);
  r.mode = mode;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // ---- CUDA-graph replay of the launch sequence (79 kernels); the scalars that change per step travel through sp_dev.  Small
  // batches are bound by the host's launch rate otherwise.  Needs the table flavour of the time embedding.
  Plan* p = e->cur;
  const void* key[4] = {x, eps_out, noise, y};
  const bool graph_ok = step_graph_enabled() && !e->prof_on && t_scalar >= 0 && t_scalar < TEMB_TABLE_ROWS && temb_table_enabled();
  if (!graph_ok) return run_ops(e, p->step_ops, r, s);
  StepParams sp;
  sp.t = t_scalar;
  sp.coef = r.coef;
  sp.seed = seed;
  sp.sample_offset = sample_offset;
  sp.mode = mode;
  sp.use_philox = use_philox;
  sp.step_index = step_index;
  sp.pad = 0;
  return run_graphed(e, p, p->g_step, r, key, &sp, s);
}

// ----------------------------------------------------------------------------------------------
// C ABI: per-op device timing
// ----------------------------------------------------------------------------------------------
extern "C" int wd_engine_set_profiling(wd_engine* e, int enable) {
  if (!e) return fail(WD_ERR_INVALID, "null engine");
  e->prof_on = enable != 0;
  if (!enable) {
    for (auto& v : e->prof_steps)
      for (auto x : v) cudaEventDestroy(x);
    e->prof_steps.clear();
    e->prof_plan = nullptr;
  }
  return WD_OK;
}

extern "C" int wd_engine_profile_read(wd_engine* e, int cap, int* kinds, double* flops, double* bytes, float* ms_sum,
                                      int* n_steps) {
  if (!e || !e->prof_plan) return fail(WD_ERR_STATE, "no profiled step recorded");
  const std::vector<Op>& ops = e->prof_plan->step_ops;
  const int n = static_cast<int>(ops.size());
  if (cap < n) return fail(WD_ERR_INVALID, "profile_read: capacity %d < %d ops", cap, n);
  CUDA_TRY(cudaDeviceSynchronize());
  for (int i = 0; i < n; ++i) {
    // the table lookup reports as the timestep-embedding class; ops of the flavour the profiled steps skipped carry no work
    const bool skipped = (ops[i].cond == COND_TABLE && !e->prof_used_table) || (ops[i].cond == COND_NO_TABLE && e->prof_used_table);
    // indices into engine.py's OP_KINDS: 0 .. 8 = OP_TEMB .. OP_UPSAMPLE, 9 = the fused transformer block
    kinds[i] = ops[i].kind == OP_EMBTBL ? static_cast<int>(OP_TEMB) : (ops[i].kind == OP_TBLOCK ? 9 : (ops[i].kind == OP_OUTHEAD ? 10 : static_cast<int>(ops[i].kind)));
    flops[i] = skipped ? 0.0 : ops[i].flops;
    bytes[i] = skipped ? 0.0 : ops[i].bytes;
    ms_sum[i] = 0.f;
  }
  for (auto& v : e->prof_steps)
    for (int i = 0; i < n; ++i) {
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
      ms_sum[i] += ms;
    }
  if (n_steps) *n_steps = static_cast<int>(e->prof_steps.size());
  return n;
}

// ----------------------------------------------------------------------------------------------
// C ABI: single operators (parity tests)
// ----------------------------------------------------------------------------------------------
extern "C" int wd_phosc_tokenize(const unsigned char* words, int batch, int max_len, int32_t* out, int32_t* bad_flag, void* stream) {
  if (!words || !out || !bad_flag) return fail(WD_ERR_INVALID, "phosc_tokenize: null argument");
  CUDA_TRY(phosc_tokenize_launch(words, batch, max_len, out, bad_flag, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_sampler_update(float* x, const float* eps, int batch, int elems_per_latent, int mode, const float* coef4_host,
                                 const float* noise, int use_philox, uint64_t seed, uint64_t sample_offset, int step_index,
                                 void* stream) {
  if (!x || !eps || !coef4_host || batch < 1 || elems_per_latent < 1) return fail(WD_ERR_INVALID, "sampler_update: null / empty argument");
  if (mode != WD_STEP_DDPM && mode != WD_STEP_DDIM) return fail(WD_ERR_INVALID, "sampler_update: mode must be WD_STEP_DDPM or WD_STEP_DDIM");
  const float4 coef = make_float4(coef4_host[0], coef4_host[1], coef4_host[2], coef4_host[3]);
  CUDA_TRY(sampler_update_launch(x, eps, noise, use_philox, seed, sample_offset * static_cast<uint64_t>(elems_per_latent), step_index, coef,
                                 mode == WD_STEP_DDPM ? STEP_DDPM : STEP_DDIM, static_cast<size_t>(batch) * elems_per_latent,
                                 static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_groupnorm(const void* x, void* out, const float* gamma, const float* beta, int B, int HW, int C,
                               int groups, float eps, int silu, void* stream) {
  if (C % groups) return fail(WD_ERR_INVALID, "C %% groups");
  if (C % 8 || C > 1024) return fail(WD_ERR_UNSUPPORTED, "groupnorm: unsupported channel count");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  const int slots = groupnorm_stats_slots(HW);
  float* partial = nullptr;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&partial), static_cast<size_t>(B) * groups * slots * 2 * sizeof(float), s));
  GroupNormStatsArgs st{static_cast<const bf16*>(x), C, partial, HW, C, cpg, slots, 0};
  CUDA_TRY(groupnorm_stats_launch(st, B, s));
  GroupNormArgs a;
  memset(&a, 0, sizeof(a));
  a.x[0] = static_cast<const bf16*>(x);
  a.x_ld[0] = C;
  a.partial[0] = partial;
  a.pslots[0] = slots;
  a.out = static_cast<bf16*>(out);
  a.out_ld = C;
  a.gamma = gamma;
  a.beta = beta;
  a.HW = HW;
  a.Cs = C;
  a.cpg = cpg;
  a.pcpg = cpg;
  a.eps = eps;
  a.silu = silu;
  a.nchunk = groupnorm_apply_chunks(HW);
  CUDA_TRY(groupnorm_launch(a, B, 1, s));
  CUDA_TRY(cudaFreeAsync(partial, s));
  return WD_OK;
}

extern "C" int wd_op_layernorm(const void* x, void* out, const float* gamma, const float* beta, int M, int C, float eps,
                               void* stream) {
  CUDA_TRY(layernorm_launch(static_cast<const bf16*>(x), static_cast<bf16*>(out), gamma, beta, M, C, eps, 0,
                            static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

static int op_gemm_impl(const void* a_, const void* w, const float* bias, const void* residual, void* out, int M, int N, int K,
                        int act_silu, int geglu, int out_f32, int res_f16, int out_f16, void* stream);
extern "C" int wd_op_gemm(const void* a_, const void* w, const float* bias, const void* residual, void* out, int M, int N,
                          int K, int act_silu, int geglu, int out_f32, void* stream) {
  return op_gemm_impl(a_, w, bias, residual, out, M, N, K, act_silu, geglu, out_f32, 0, 0, stream);
}
extern "C" int wd_op_gemm_f16(const void* a_, const void* w, const float* bias, const void* residual_f16, void* out_f16, int M,
                              int N, int K, void* stream) {
  return op_gemm_impl(a_, w, bias, residual_f16, out_f16, M, N, K, 0, 0, 0, 1, 1, stream);
}
static int op_gemm_impl(const void* a_, const void* w, const float* bias, const void* residual, void* out, int M, int N, int K,
                        int act_silu, int geglu, int out_f32, int res_f16, int out_f16, void* stream) {
  if (K % GEMM_BLOCK_K || N % gemm_tc_block_n()) return fail(WD_ERR_UNSUPPORTED, "gemm: K %% 64 or N %% %d", gemm_tc_block_n());
  GemmLaunch L;
  memset(&L, 0, sizeof(L));
  GemmArgs& a = L.args;
  a.M = M;
  a.N = N;
  a.num_src = 1;
  a.taps[0] = 1;
  a.chunks[0] = K / GEMM_BLOCK_K;
  a.stride[0] = 1;
  a.conv = 0;
  a.Wout = 1;
  a.HWout = 1;
  a.bias = bias;
  a.rows_per_sample = 1;
  a.residual = static_cast<const bf16*>(residual);
  const int out_cols = geglu ? N / 2 : N;
  a.res_ld = out_cols;
  a.out = out;
  a.out_ld = out_cols;
  a.out_f32 = out_f32;
  a.res_f16 = res_f16;
  a.out_f16 = out_f16;
  a.act = act_silu ? ACT_SILU : ACT_NONE;
  a.geglu = geglu;
  if (!tmap_encode_2d_bf16(&L.mapA[0], a_, K, M, K, GEMM_BLOCK_K, GEMM_BLOCK_M)) return fail(WD_ERR_CUDA, "tensor map A");
  L.mapA[1] = L.mapA[2] = L.mapA[0];
  if (!tmap_encode_2d_bf16(&L.mapB, w, K, N, K, GEMM_BLOCK_K, gemm_b_box_rows(a))) return fail(WD_ERR_CUDA, "tensor map B");
  L.mapOut = L.mapRes = L.mapB;
  if (!out_f32 && !tmap_encode_out_bf16(&L.mapOut, out, out_cols, M, out_cols)) return fail(WD_ERR_CUDA, "tensor map out");
  if (residual && !tmap_encode_out_bf16(&L.mapRes, residual, out_cols, M, out_cols)) return fail(WD_ERR_CUDA, "tensor map residual");
  if (!gemm_prepare_res_k(L)) return fail(WD_ERR_CUDA, "tensor map residual (K blocks)");
  CUDA_TRY(gemm_tc_launch(L, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_q_ctx_attention(const void* a_, const void* wq, const float* bias, const void* kv, void* out, int B, int Sq,
                                     int L, int heads, float scale, void* stream) {
  const int C = heads * 80, M = B * Sq;
  if (C % gemm_tc_block_n() || C % GEMM_BLOCK_K || C > 5 * GEMM_BLOCK_K)
    return fail(WD_ERR_UNSUPPORTED, "q_ctx_attention: heads * 80 must be a multiple of %d and <= 320", gemm_tc_block_n());
  if (Sq % GEMM_BLOCK_M || L < 1 || L > GEMM_ATT_MAXL)
    return fail(WD_ERR_UNSUPPORTED, "q_ctx_attention: Sq %% 128 != 0 or L > %d", GEMM_ATT_MAXL);
  GemmLaunch Lh;
  memset(&Lh, 0, sizeof(Lh));
  GemmArgs& a = Lh.args;
  a.M = M;
  a.N = C;
  a.num_src = 1;
  a.taps[0] = 1;
  a.chunks[0] = C / GEMM_BLOCK_K;
  a.stride[0] = 1;
  a.Wout = 1;
  a.HWout = 1;
  a.bias = bias;
  a.rows_per_sample = Sq;
  a.out = out;
  a.out_ld = C;
  a.att_kv = static_cast<const bf16*>(kv);
  a.att_ld = 2 * C;
  a.att_voff = C;
  a.att_L = L;
  a.att_scale = scale;
  if (!tmap_encode_2d_bf16(&Lh.mapA[0], a_, C, M, C, GEMM_BLOCK_K, GEMM_BLOCK_M)) return fail(WD_ERR_CUDA, "tensor map A");
  Lh.mapA[1] = Lh.mapA[2] = Lh.mapA[0];
  if (!tmap_encode_2d_bf16(&Lh.mapB, wq, C, C, C, GEMM_BLOCK_K, gemm_b_box_rows(a))) return fail(WD_ERR_CUDA, "tensor map B");
  Lh.mapOut = Lh.mapRes = Lh.mapB;
  if (!tmap_encode_out_bf16(&Lh.mapOut, out, C, M, C)) return fail(WD_ERR_CUDA, "tensor map out");
  CUDA_TRY(gemm_tc_launch(Lh, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_conv3x3(const void* x, const void* w_packed, const float* bias, const float* rowbias, int rb_ld,
                             const void* residual, void* out, int B, int H, int W, int Cin, int Cout, int stride,
                             void* stream) {
  if (Cin % GEMM_BLOCK_K || Cout % gemm_tc_block_n()) return fail(WD_ERR_UNSUPPORTED, "conv3x3: channel counts");
  if (stride != 1 && stride != 2) return fail(WD_ERR_INVALID, "stride");
  const int Ho = H / stride, Wo = W / stride, HWo = Ho * Wo;
  GemmLaunch L;
  memset(&L, 0, sizeof(L));
  GemmArgs& a = L.args;
  a.M = B * HWo;
  a.N = Cout;
  a.num_src = 1;
  a.taps[0] = 9;
  a.chunks[0] = Cin / GEMM_BLOCK_K;
  a.stride[0] = stride;
  a.conv = 1;
  a.Wout = Wo;
  a.HWout = HWo;
  a.bias = bias;
  a.rowbias = rowbias;
  a.rb_ld = rb_ld;
  a.rows_per_sample = HWo;
  a.residual = static_cast<const bf16*>(residual);
  a.res_ld = Cout;
  a.out = out;
  a.out_ld = Cout;
  uint32_t bw, bh, bnn;
  if (HWo >= GEMM_BLOCK_M) {
    if (HWo % GEMM_BLOCK_M || GEMM_BLOCK_M % Wo) return fail(WD_ERR_UNSUPPORTED, "conv3x3: spatial size");
    bw = Wo * stride;
    bh = (GEMM_BLOCK_M / Wo) * stride;
    bnn = 1;
  } else {
    if (GEMM_BLOCK_M % HWo) return fail(WD_ERR_UNSUPPORTED, "conv3x3: spatial size");
    bw = W;
    bh = H;
    bnn = GEMM_BLOCK_M / HWo;
  }
  if (!tmap_encode_4d_bf16(&L.mapA[0], x, Cin, W, H, B, Cin, GEMM_BLOCK_K, bw, bh, bnn, stride))
    return fail(WD_ERR_CUDA, "tensor map A");
  L.mapA[1] = L.mapA[2] = L.mapA[0];
  if (!tmap_encode_2d_bf16(&L.mapB, w_packed, 9 * Cin, Cout, 9 * Cin, GEMM_BLOCK_K, gemm_b_box_rows(a)))
    return fail(WD_ERR_CUDA, "tensor map B");
  L.mapOut = L.mapRes = L.mapB;
  if (!tmap_encode_out_bf16(&L.mapOut, out, Cout, a.M, Cout)) return fail(WD_ERR_CUDA, "tensor map out");
  if (residual && !tmap_encode_out_bf16(&L.mapRes, residual, Cout, a.M, Cout)) return fail(WD_ERR_CUDA, "tensor map residual");
  CUDA_TRY(gemm_tc_launch(L, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

// conv3x3 (stride 1) + bias + per-sample row bias -> GroupNorm(32 groups) -> SiLU with the normalisation applied by the conv's own
// epilogue (GemmArgs::gn_apply, pair kernel): the front half of ResBlock._forward (unet.py:657-667 then :592-594)
extern "C" int wd_op_conv3x3_gn_silu(const void* x, const void* w_packed, const float* bias, const float* rowbias, int rb_ld,
                                     const float* gamma, const float* beta, float eps, void* out, float* stats_ws, int B, int H, int W,
                                     int Cin, int Cout, void* stream) {
  if (!x || !w_packed || !gamma || !beta || !out || !stats_ws) return fail(WD_ERR_INVALID, "conv3x3_gn_silu: null argument");
  if (Cin % GEMM_BLOCK_K || Cout != GEMM_PAIR_BLOCK_N) return fail(WD_ERR_UNSUPPORTED, "conv3x3_gn_silu: channel counts");
  const int HW = H * W;
  if (HW % 32 || 256 % HW) return fail(WD_ERR_UNSUPPORTED, "conv3x3_gn_silu: H*W must divide 256 and be a multiple of 32");
  GemmLaunch L;
  memset(&L, 0, sizeof(L));
  GemmArgs& a = L.args;
  a.M = B * HW;
  a.N = Cout;
  a.num_src = 1;
  a.taps[0] = 9;
  a.chunks[0] = Cin / GEMM_BLOCK_K;
  a.stride[0] = 1;
  a.conv = 1;
  a.Wout = W;
  a.HWout = HW;
  a.bias = bias;
  a.rowbias = rowbias;
  a.rb_ld = rb_ld;
  a.rows_per_sample = HW;
  a.out = out;
  a.out_ld = Cout;
  a.gn_partial = stats_ws;  // [B][32][HW / 32][2] fp32
  a.gn_cpg = 10;
  a.gn_apply = 1;
  a.gn_gamma = gamma;
  a.gn_beta = beta;
  a.gn_eps = eps;
  if (!gemm_uses_pair(a)) return fail(WD_ERR_UNSUPPORTED, "conv3x3_gn_silu: this shape does not run on the pair kernel");
  uint32_t bw, bh, bnn;
  if (HW >= GEMM_BLOCK_M) {
    if (HW % GEMM_BLOCK_M || GEMM_BLOCK_M % W) return fail(WD_ERR_UNSUPPORTED, "conv3x3_gn_silu: spatial size");
    bw = W;
    bh = GEMM_BLOCK_M / W;
    bnn = 1;
  } else {
    if (GEMM_BLOCK_M % HW) return fail(WD_ERR_UNSUPPORTED, "conv3x3_gn_silu: spatial size");
    bw = W;
    bh = H;
    bnn = GEMM_BLOCK_M / HW;
  }
  if (!tmap_encode_4d_bf16(&L.mapA[0], x, Cin, W, H, B, Cin, GEMM_BLOCK_K, bw, bh, bnn, 1)) return fail(WD_ERR_CUDA, "tensor map A");
  L.mapA[1] = L.mapA[2] = L.mapA[0];
  if (!tmap_encode_2d_bf16(&L.mapB, w_packed, 9 * Cin, Cout, 9 * Cin, GEMM_BLOCK_K, gemm_b_box_rows(a)))
    return fail(WD_ERR_CUDA, "tensor map B");
  L.mapOut = L.mapRes = L.mapB;
  if (!tmap_encode_out_bf16(&L.mapOut, out, Cout, a.M, Cout)) return fail(WD_ERR_CUDA, "tensor map out");
  CUDA_TRY(gemm_tc_launch(L, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_pack_conv3x3(const float* w, void* dst, int Cout, int Cin, void* stream) {
  CUDA_TRY(repack_conv3x3_launch(w, static_cast<bf16*>(dst), Cout, Cin, 9 * Cin, 0, 0, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}
extern "C" int wd_op_pack_linear(const float* w, void* dst, int N, int K, int geglu_perm, void* stream) {
  CUDA_TRY(repack_linear_launch(w, static_cast<bf16*>(dst), N, K, K, 0, 0, geglu_perm ? gemm_geglu_block(N) : 0, 0,
                                static_cast<cudaStream_t>(stream)));
  return WD_OK;
}
extern "C" int wd_op_pack_vec_geglu(const float* v, float* dst, int N, void* stream) {
  CUDA_TRY(repack_vec_launch(v, dst, N, 0, gemm_geglu_block(N), 0, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_attention_small(const void* q, const void* k, const void* v, void* out, float* probs, int B, int Sq,
                                     int L, int heads, float scale, void* stream) {
  const int C = heads * 80;
  AttnSmallArgs a{static_cast<const bf16*>(q), C, static_cast<const bf16*>(k), static_cast<const bf16*>(v), C,
                  static_cast<bf16*>(out), C, probs, Sq, L, heads, scale};
  CUDA_TRY(attn_small_launch(a, B, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_attention(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int ldo, int B,
                               int Sq, int Skv, int heads, float scale, void* stream) {
  AttnFlashArgs a{static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), static_cast<const bf16*>(v), ldkv,
                  static_cast<bf16*>(out), ldo, Sq, Skv, heads, scale};
  CUDA_TRY(attn_flash_launch(a, B, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

// Fused transformer block as a single operator (tests/test_gpu_tblock.py).  tensors[] (device pointers, fp32 unless noted):
//  0 g bf16 [M,320]   1 x_in fp16 [M,320]   2 ctx bf16 [B*L,320]
//  3 proj_in.w [320,320]  4 proj_in.b  5 norm2.w  6 norm2.b  7 norm3.w  8 norm3.b
//  9 attn1.to_q.w  10 attn1.to_k.w  11 attn1.to_v.w  12 attn1.to_out.0.w  13 attn1.to_out.0.b
// 14 attn2.to_q.w  15 attn2.to_k.w  16 attn2.to_v.w  17 attn2.to_out.0.w  18 attn2.to_out.0.b
// 19 ff.net.0.proj.w [2560,320]  20 ff.net.0.proj.b [2560]  21 ff.net.2.w [320,1280]  22 ff.net.2.b  23 proj_out.w  24 proj_out.b
// stage: TBlockArgs::stage, or 5 = the middle form (x_in is the residual stream itself: no proj_in / proj_out, out = the raw
// stream after the feed-forward).  HW: a multiple of 128, or 64 (two samples per tile; B may be odd).
// out: fp16 [M,320]; gn_partial: fp32 [B][32][HW/32][2] or NULL.  Synchronises the stream.
extern "C" int wd_op_tblock_unet(const void* const* tensors, int n_tensors, int B, int HW, int L, int stage, void* out_f16,
                                 float* gn_partial, void* stream) {
  // 27 tensors: [25] / [26] = norm.weight / norm.bias of the SpatialTransformer's GroupNorm -- the kernel normalises x_in itself
  // (tensors[0] is ignored) from partial statistics computed here by groupnorm_stats_kernel
  if (!tensors || (n_tensors != 25 && n_tensors != 27) || !out_f16) return fail(WD_ERR_INVALID, "op_tblock_unet: expects 25 or 27 tensors");
  const bool gn_in = n_tensors == 27;
  for (int i = 0; i < n_tensors; ++i)
    if (!tensors[i]) return fail(WD_ERR_INVALID, "op_tblock_unet: tensor %d is null", i);
  if (B < 1 || (HW % TB_M && HW != TB_M / 2) || L < 1 || L > TB_KEYS || stage < 0 || stage > 5)
    return fail(WD_ERR_INVALID, "op_tblock_unet: bad shape");
  const bool mid = stage == 5;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto F = [&](int i) { return static_cast<const float*>(tensors[i]); };
  const int C = TB_C, M = B * HW;
  char* ws = nullptr;
  Arena A;
  const size_t need = 64ull << 20;
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ws), need + static_cast<size_t>(B) * L * (2 * TB_FOLD_N * 2 + 64)));
  A.base = ws;
  int rc = WD_OK;
  do {
    bf16* w_pi = A.alloc<bf16>(C * C);
    bf16* w_po = A.alloc<bf16>(C * C);
    bf16* w_ff2 = A.alloc<bf16>(C * 4 * C);
    bf16* w_ff1 = A.alloc<bf16>(8 * C * C);
    float* b_ff1 = A.alloc<float>(8 * C);
    float* s_ff1 = A.alloc<float>(8 * C);
    bf16* w_fold = A.alloc<bf16>(static_cast<size_t>(2) * TB_FOLD_N * C);
    float* u = A.alloc<float>(2 * TB_HEADS * C);
    float* cb = A.alloc<float>(4 * C);
    bf16* fold_out = A.alloc<bf16>(static_cast<size_t>(B) * L * 2 * TB_FOLD_N);
    float* cvec = A.alloc<float>(static_cast<size_t>(B) * L * 2 * TB_HEADS);
    bf16* eye_d = A.alloc<bf16>(static_cast<size_t>(C) * C);
    const int gn_slots = groupnorm_stats_slots(HW);
    float* gn_part = A.alloc<float>(static_cast<size_t>(B) * 32 * gn_slots * 2);
#define TB_TRY(x) if ((x) != cudaSuccess) { rc = fail(WD_ERR_CUDA, "op_tblock_unet: %s", cudaGetErrorString(cudaGetLastError())); break; }
    TB_TRY(repack_linear_launch(F(3), w_pi, C, C, C, 0, 0, 0, 0, s));
    TB_TRY(repack_linear_launch(F(23), w_po, C, C, C, 0, 0, 0, 1, s));
    TB_TRY(repack_linear_launch(F(21), w_ff2, C, 4 * C, 4 * C, 0, 0, 0, 0, s));
    TB_TRY(fold_ln_linear_launch(F(19), F(7), F(8), F(20), w_ff1, s_ff1, b_ff1, 8 * C, C, C, 0, 2 * TB_CHUNK, s));
    for (int a = 0; a < 2; ++a)
      TB_TRY(tblock_fold_weights_launch(F(9 + 5 * a), F(10 + 5 * a), F(11 + 5 * a), F(12 + 5 * a), F(5), F(6),
                                        w_fold + static_cast<size_t>(a) * TB_FOLD_N * C, u + a * TB_HEADS * C, s));
    if (rc != WD_OK) break;
    const float* add[4] = {F(4), F(13), F(18), F(22)};
    for (int i = 0; i < 4; ++i) {
      if (i > 0) TB_TRY(repack_vec_launch(cb + (i - 1) * C, cb + i * C, C, 0, 0, 0, s));
      if (i == 0 && mid) {
        TB_TRY(cudaMemsetAsync(cb, 0, C * sizeof(float), s));
      } else {
        TB_TRY(repack_vec_launch(add[i], cb + i * C, C, 0, 0, i > 0 ? 1 : 0, s));
      }
    }
    if (rc != WD_OK) break;
    if (mid) {
      std::vector<uint16_t> eye(static_cast<size_t>(C) * C, 0);
      for (int i = 0; i < C; ++i) eye[static_cast<size_t>(i) * C + i] = 0x3C00;
      TB_TRY(cudaMemcpyAsync(eye_d, eye.data(), eye.size() * 2, cudaMemcpyHostToDevice, s));
      TB_TRY(cudaStreamSynchronize(s));
    }
    rc = op_gemm_impl(tensors[2], w_fold, nullptr, nullptr, fold_out, B * L, 2 * TB_FOLD_N, C, 0, 0, 0, 0, 1, s);
    if (rc != WD_OK) break;
    TB_TRY(tblock_cvec_launch(static_cast<const bf16*>(tensors[2]), u, cvec, B * L, 2 * TB_HEADS, s));
    TBlockLaunch T;
    memset(&T, 0, sizeof(T));
    TBlockArgs& a = T.args;
    a.M = M; a.HW = HW; a.L = L; a.cb = cb; a.b_ff = b_ff1; a.cvec1 = cvec; a.cvec2 = cvec + TB_HEADS; a.cvec_ld = 2 * TB_HEADS;
    a.b_po = F(24); a.x_in = static_cast<const __half*>(tensors[1]); a.x_in_ld = C; a.gn_partial = gn_partial; a.ln_eps = 1e-5f;
    a.stage = mid ? 4 : stage;
    a.mid = mid ? 1 : 0;
    a.pair = tblock_use_pair(HW) ? 1 : 0;
    if (gn_in) {
      if (mid) { rc = fail(WD_ERR_INVALID, "op_tblock_unet: the middle form has no input GroupNorm"); break; }
      GroupNormStatsArgs st{static_cast<const bf16*>(tensors[1]), C, gn_part, HW, C, C / 32, gn_slots, 1};
      TB_TRY(groupnorm_stats_launch(st, B, s));
      a.gn_in_partial = gn_part;
      a.gn_in_slots = gn_slots;
      a.gn_gamma = F(25);
      a.gn_beta = F(26);
      a.gn_eps = 1e-6f;
    }
    const int wbox = a.pair ? 80 : 160, w1box = a.pair ? TB_CHUNK : 2 * TB_CHUNK;
    const int fold_ld = 2 * TB_FOLD_N;
    bool ok = tmap_encode_2d_bf16(&T.mapG, (mid || gn_in) ? tensors[1] : tensors[0], C, M, C, 64, TB_M) &&
              tmap_encode_2d_bf16(&T.mapWpi, mid ? eye_d : w_pi, C, C, C, 64, wbox) &&
              tmap_encode_2d_bf16(&T.mapW1, w_ff1, C, 8 * C, C, 64, w1box) &&
              tmap_encode_2d_bf16(&T.mapW2, w_ff2, 4 * C, C, 4 * C, 64, wbox) && tmap_encode_2d_bf16(&T.mapWpo, w_po, C, C, C, 64, wbox) &&
              tmap_encode_2d_bf16(&T.mapOut, out_f16, C, M, C, 64, TB_M);
    for (int i = 0; i < 4 && ok; ++i)
      ok = tmap_encode_3d_bf16(&T.mapF[i], fold_out + static_cast<size_t>(i / 2) * TB_FOLD_N + (i & 1) * (TB_HEADS * C), TB_HEADS * C, L, B,
                               fold_ld, static_cast<uint64_t>(L) * fold_ld, 64, TB_KEYS);
    if (!ok) { rc = fail(WD_ERR_CUDA, "op_tblock_unet: cuTensorMapEncodeTiled failed"); break; }
    TB_TRY(tblock_launch(T, engine_num_sms(), s));
#undef TB_TRY
  } while (0);
  const cudaError_t se = cudaStreamSynchronize(s);
  cudaFree(ws);
  if (rc == WD_OK && se != cudaSuccess) rc = fail(WD_ERR_CUDA, "op_tblock_unet: %s", cudaGetErrorString(se));
  return rc;
}
