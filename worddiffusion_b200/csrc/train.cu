// Training engine of the WordDiffusion hot path: the noise-prediction step of reference train.py:281-294
//   predicted_noise = model(x_t, ..., timesteps=t, context=text_features, y=s_id); loss = mse(noise, predicted_noise);
//   loss.backward(); optimizer.step(); ema.step_ema(...)
// for unet.UNetModel (the model train.py:403 builds).  `wd_trainer` owns bf16 packs of the caller's fp32 parameters (forward
// layout [N, K] and transposed [K, N] for the data-gradient GEMMs), an activation arena in which every forward tensor stays
// alive until the backward pass has consumed it, and a per-batch plan: a flat list of forward launches and a flat list of
// backward launches (built once by walking the layer inventory forwards, then its tape backwards).
//   forward  : same kernels as the inference engine (gemm_tc.cu, ops.cu, attn_flash.cu), all tensors bf16, LayerNorm and
//              GEGLU as separate kernels so that their inputs are available to the backward pass
//   backward : data gradients  = gemm_tc.cu with transposed weight packs (3x3 conv: flipped taps),
//              weight gradients = wgrad_tc.cu (tcgen05, MN-major operands), accumulated into the caller's fp32 .grad tensors
//              norms / activations / attention / resampling = ops_bwd.cu
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/wd_b200.h"
#include "engine_internal.h"
#include "gemm_tc.cuh"
#include "ops.cuh"
#include "ops_bwd.cuh"
#include "wgrad_tc.cuh"

using namespace wd;
typedef __nv_bfloat16 bf16;

static int tfail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return wd_set_error(code, buf);
}
#define T_CUDA_TRY(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return tfail(WD_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

namespace {

struct TArena {
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

struct TParam {
  const float* w = nullptr;
  float* g = nullptr;
  int64_t numel = 0;
};

enum PackKind { PK_LIN, PK_LIN_T, PK_CONV3, PK_CONV3_T, PK_CONV_IN, PK_VEC };
struct PackJob {
  PackKind kind;
  std::string src;
  void* dst;
  int N, K;          // PK_LIN*: [N, K]; PK_CONV3*: Cout, Cin; PK_VEC: N elements
  int ld, off0, off1;  // PK_LIN: ldk, k_off, n_off | PK_LIN_T: ldn, n_off, k_row_off | PK_CONV3: ldk, k_off, as_f16 | PK_CONV3_T: cout_pad | PK_VEC: -, n_off, accumulate
};

// ---------------- layer inventory (state_dict keys + packs) ----------------
struct TNorm { std::string g, b; int C = 0; };
struct TLin {
  std::string w, b;  // b empty: no bias
  int N = 0, K = 0;
  bf16* pw = nullptr;   // [N, K]
  bf16* pwt = nullptr;  // [K, N]
};
struct TConv {
  std::string w, b;
  int Cout = 0, Cin = 0;
  bf16* pw = nullptr;   // [Cout (padded), Ktot]
  int Ktot = 0;
  bf16* pwt = nullptr;  // [Cin, 9 * cout_pad]
  int cout_pad = 0;
};
struct TRes {
  int Cin = 0, Cout = 0;
  TNorm gn1, gn2;
  TConv conv1, conv2;
  bool skip_conv = false;
  TLin skip;  // pw unused (fused along K of conv2.pw)
  float* bias2 = nullptr;  // out_layers.3.bias (+ skip_connection.bias)
  std::string emb_w, emb_b;
  int emb_off = 0;
};
struct TBlk {
  TNorm ln2, ln3;
  TLin q1, o1, q2, o2, ffp, ffo;
  std::string k1, v1, k2, v2;
  int kv1 = 0, kv2 = 0;  // index of the attention's [to_k; to_v] block inside the fused context projection
};
struct TST {
  int C = 0, heads = 0, dh = 0;
  TNorm gn;
  TLin proj_in, proj_out;
  std::vector<TBlk> blocks;
};
struct TSamp { int C = 0; TConv conv; };
enum TLayerKind { TL_CONVIN, TL_RES, TL_ST, TL_DOWN, TL_UP };
struct TLayer { TLayerKind kind; int idx; };
typedef std::vector<TLayer> TBlock;

// ---------------- plan ----------------
struct TRun {
  const float* x = nullptr;
  const long long* t = nullptr;
  const long long* y = nullptr;
  const long long* ctx = nullptr;
  float* eps_out = nullptr;
  const float* d_eps = nullptr;
};
typedef std::function<cudaError_t(const TRun&, cudaStream_t)> TFn;
struct TOp {
  const char* what;
  TFn fn;
};
struct TT {  // activation tensor + its gradient
  bf16* p = nullptr;
  bf16* g = nullptr;
  int C = 0, H = 0, W = 0;
  float* stats = nullptr;
  int pslots = 0;
  bool g_init = false;  // (plan-build time) some backward op already wrote g
};
struct TPlan {
  int B = 0, L = 0;
  std::vector<TOp> fwd, bwd;
  std::deque<TT> tensors;
  size_t bytes = 0;
  std::vector<std::pair<void*, size_t>> zero_on_bwd;  // scratch that must be zero when the backward pass starts
  std::map<std::string, std::pair<const void*, size_t>> named;  // intermediates readable through wd_trainer_read_tensor (tests)
  // per-call inputs / outputs are staged at fixed addresses so that the two launch lists can be replayed as CUDA graphs
  float* in_x = nullptr;
  long long *in_t = nullptr, *in_y = nullptr, *in_ctx = nullptr;
  float* out_eps = nullptr;
  float* in_deps = nullptr;
  cudaGraphExec_t g_fwd = nullptr, g_bwd = nullptr;
  int fwd_calls = 0, bwd_calls = 0;
  // Backward stages (one per layer of the tape, in execution order): stage s covers bwd[stage_end[s-1], stage_end[s]).
  // grad_stage[name] = the last stage that writes the gradient of that parameter (it is final once the stage has run), so
  // the caller can all-reduce a bucket of gradients while the later stages still run (wd_trainer_backward_stages).
  std::vector<int> stage_end;
  std::map<std::string, int> grad_stage;
  struct Seg { cudaGraphExec_t exec = nullptr; int calls = 0; };
  std::map<std::pair<int, int>, Seg> segs;  // launch list [stage_begin, stage_end) as its own CUDA graph
  ~TPlan() {
    if (g_fwd) cudaGraphExecDestroy(g_fwd);
    if (g_bwd) cudaGraphExecDestroy(g_bwd);
    for (auto& kv : segs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  }
};

}  // namespace

struct wd_trainer {
  wd_config cfg;
  int time_dim = 0;
  std::unordered_map<std::string, TParam> params;
  std::unordered_map<std::string, int64_t> expected;  // live parameters: key -> element count
  std::vector<PackJob> jobs;
  void* pack_tab_dev = nullptr;       // device copy of the matrix-pack job table (repack_multi_launch)
  std::vector<char> pack_tab_host;    // what was uploaded last
  char* wbase = nullptr;
  size_t wbytes = 0;
  // inventory
  std::vector<TRes> res;
  std::vector<TST> st;
  std::vector<TSamp> samp;
  std::vector<TBlock> input_blocks, output_blocks;
  TBlock middle;
  TLin te0, te2;
  bf16* emb_all_w = nullptr;   // [emb_cols, ted]
  bf16* emb_all_wt = nullptr;  // [ted, emb_cols]
  float* emb_all_b = nullptr;  // [emb_cols]
  int emb_cols = 0;
  bf16* kv_all_w = nullptr;   // [n_kv * 2 * inner, ctx_dim]
  bf16* kv_all_wt = nullptr;  // [ctx_dim, n_kv * 2 * inner]
  int n_kv = 0, kv_cols = 0;
  std::vector<std::string> kv_names;  // 2 per attention: to_k, to_v
  bf16* wqkv_t = nullptr;  // word_emb attention [D, 3D] (transposed packs of query | key | value)
  bf16* conv_in_w = nullptr;  // [mc, 128]
  TNorm out_gn;
  TConv conv_out;
  float* pe = nullptr;
  bool pe_set = false;
  // activations
  char* abase = nullptr;
  size_t acap = 0;
  std::map<std::pair<int, int>, std::unique_ptr<TPlan>> plans;
  TPlan* cur = nullptr;
  bool fwd_done = false;
  bool packs_valid = false;
  std::map<std::string, int> grad_stage;  // from a dry plan build (the order does not depend on batch / context length)
  int n_stages = 0;
  int bwd_next_stage = 0;  // staged backward: the stage the next wd_trainer_backward_stages call must start at
};

namespace {

// ----------------------------------------------------------------------------------------------
// inventory builder (same loops as reference unet.py:1248-1458)
// ----------------------------------------------------------------------------------------------
struct TBuilder {
  wd_trainer* e;
  TArena& A;
  bool dry;

  void expect(const std::string& name, int64_t numel) {
    if (!dry) e->expected[name] = numel;
  }
  void job(PackKind k, const std::string& src, void* dst, int N, int K, int ld, int off0, int off1) {
    if (!dry) e->jobs.push_back(PackJob{k, src, dst, N, K, ld, off0, off1});
  }
  TNorm norm(const std::string& pfx, int C) {
    expect(pfx + ".weight", C);
    expect(pfx + ".bias", C);
    return TNorm{pfx + ".weight", pfx + ".bias", C};
  }
  TLin linear(const std::string& pfx, int N, int K, bool bias, bool fwd_pack = true) {
    TLin l;
    l.w = pfx + ".weight";
    l.N = N;
    l.K = K;
    expect(l.w, static_cast<int64_t>(N) * K);
    if (bias) {
      l.b = pfx + ".bias";
      expect(l.b, N);
    }
    if (fwd_pack) {
      l.pw = A.alloc<bf16>(static_cast<size_t>(N) * K);
      job(PK_LIN, l.w, l.pw, N, K, K, 0, 0);
    }
    l.pwt = A.alloc<bf16>(static_cast<size_t>(N) * K);
    job(PK_LIN_T, l.w, l.pwt, N, K, N, 0, 0);
    return l;
  }
  TConv conv3(const std::string& pfx, int Cout, int Cin, int extraK = 0) {
    TConv c;
    c.w = pfx + ".weight";
    c.b = pfx + ".bias";
    c.Cout = Cout;
    c.Cin = Cin;
    c.Ktot = 9 * Cin + extraK;
    expect(c.w, static_cast<int64_t>(Cout) * Cin * 9);
    expect(c.b, Cout);
    c.pw = A.alloc<bf16>(static_cast<size_t>(Cout) * c.Ktot);
    job(PK_CONV3, c.w, c.pw, Cout, Cin, c.Ktot, 0, 0);
    c.cout_pad = Cout;
    c.pwt = A.alloc<bf16>(static_cast<size_t>(Cin) * 9 * Cout);
    job(PK_CONV3_T, c.w, c.pwt, Cout, Cin, Cout, 0, 0);
    return c;
  }
  int add_res(const std::string& pfx, int Cin, int Cout, int& emb_cols) {
    TRes r;
    r.Cin = Cin;
    r.Cout = Cout;
    r.gn1 = norm(pfx + "in_layers.0", Cin);
    r.conv1 = conv3(pfx + "in_layers.2", Cout, Cin);
    r.emb_w = pfx + "emb_layers.1.weight";
    r.emb_b = pfx + "emb_layers.1.bias";
    expect(r.emb_w, static_cast<int64_t>(Cout) * e->time_dim);
    expect(r.emb_b, Cout);
    r.emb_off = emb_cols;
    emb_cols += Cout;
    r.gn2 = norm(pfx + "out_layers.0", Cout);
    r.skip_conv = (Cin != Cout);
    r.conv2 = conv3(pfx + "out_layers.3", Cout, Cout, r.skip_conv ? Cin : 0);
    r.bias2 = A.alloc<float>(Cout);
    job(PK_VEC, r.conv2.b, r.bias2, Cout, 0, 0, 0, 0);
    if (r.skip_conv) {
      r.skip = linear(pfx + "skip_connection", Cout, Cin, true, false);
      job(PK_LIN, r.skip.w, r.conv2.pw, Cout, Cin, r.conv2.Ktot, 9 * Cout, 0);
      job(PK_VEC, r.skip.b, r.bias2, Cout, 0, 0, 0, 1);
    }
    e->res.push_back(r);
    return static_cast<int>(e->res.size()) - 1;
  }
  int add_st(const std::string& pfx, int C, int heads, int dh) {
    TST s;
    s.C = C;
    s.heads = heads;
    s.dh = dh;
    const int inner = heads * dh, D = e->cfg.context_dim;
    s.gn = norm(pfx + "norm", C);
    s.proj_in = linear(pfx + "proj_in", inner, C, true);
    for (int d = 0; d < e->cfg.transformer_depth; ++d) {
      const std::string tp = pfx + "transformer_blocks." + std::to_string(d) + ".";
      TBlk t;
      t.ln2 = norm(tp + "norm2", inner);
      t.ln3 = norm(tp + "norm3", inner);
      t.q1 = linear(tp + "attn1.to_q", inner, inner, false);
      t.o1 = linear(tp + "attn1.to_out.0", inner, inner, true);
      t.q2 = linear(tp + "attn2.to_q", inner, inner, false);
      t.o2 = linear(tp + "attn2.to_out.0", inner, inner, true);
      t.ffp = linear(tp + "ff.net.0.proj", inner * 8, inner, true);
      t.ffo = linear(tp + "ff.net.2", inner, inner * 4, true);
      t.k1 = tp + "attn1.to_k.weight";
      t.v1 = tp + "attn1.to_v.weight";
      t.k2 = tp + "attn2.to_k.weight";
      t.v2 = tp + "attn2.to_v.weight";
      for (const std::string* n : {&t.k1, &t.v1, &t.k2, &t.v2}) expect(*n, static_cast<int64_t>(inner) * D);
      t.kv1 = e->n_kv++;
      t.kv2 = e->n_kv++;
      if (!dry) {
        e->kv_names.push_back(t.k1);
        e->kv_names.push_back(t.v1);
        e->kv_names.push_back(t.k2);
        e->kv_names.push_back(t.v2);
      }
      s.blocks.push_back(t);
    }
    s.proj_out = linear(pfx + "proj_out", C, inner, true);
    e->st.push_back(s);
    return static_cast<int>(e->st.size()) - 1;
  }
  int add_samp(const std::string& pfx, int C) {
    TSamp s;
    s.C = C;
    s.conv = conv3(pfx, C, C);
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
    const int mc = c.model_channels, ted = mc * 4, D = c.context_dim;
    e->time_dim = ted;
    e->res.clear();
    e->st.clear();
    e->samp.clear();
    e->input_blocks.clear();
    e->output_blocks.clear();
    e->middle.clear();
    e->jobs.clear();
    e->expected.clear();
    e->kv_names.clear();
    e->n_kv = 0;

    e->te0 = linear("time_embed.0", ted, mc, true);
    e->te2 = linear("time_embed.2", ted, ted, true);
    expect("word_emb.embedding.weight", static_cast<int64_t>(c.vocab_size) * D);
    e->wqkv_t = A.alloc<bf16>(static_cast<size_t>(D) * 3 * D);
    const char* qkv[3] = {"linear_query", "linear_key", "linear_value"};
    for (int i = 0; i < 3; ++i) {
      const std::string p = std::string("word_emb.attention.") + qkv[i];
      expect(p + ".weight", static_cast<int64_t>(D) * D);
      expect(p + ".bias", D);
      job(PK_LIN_T, p + ".weight", e->wqkv_t, D, D, 3 * D, i * D, 0);
    }
    if (c.num_classes > 0 && c.add_label_emb) expect("label_emb.weight", static_cast<int64_t>(c.num_classes) * ted);
    e->pe = A.alloc<float>(static_cast<size_t>(c.max_seq_len) * D);
    expect("input_blocks.0.0.weight", static_cast<int64_t>(mc) * c.in_channels * 9);
    expect("input_blocks.0.0.bias", mc);
    e->conv_in_w = A.alloc<bf16>(static_cast<size_t>(mc) * 128);
    job(PK_CONV_IN, "input_blocks.0.0.weight", e->conv_in_w, mc, c.in_channels, 0, 0, 0);
    e->input_blocks.push_back(TBlock{TLayer{TL_CONVIN, 0}});

    int emb_cols = 0;
    std::vector<int> chans{mc};
    int ch = mc, ds = 1;
    for (int level = 0; level < c.n_channel_mult; ++level) {
      const int mult = c.channel_mult[level];
      for (int i = 0; i < c.num_res_blocks; ++i) {
        const std::string pfx = "input_blocks." + std::to_string(e->input_blocks.size()) + ".";
        TBlock b;
        b.push_back(TLayer{TL_RES, add_res(pfx + "0.", ch, mult * mc, emb_cols)});
        ch = mult * mc;
        if (in_attn_res(ds)) {
          int heads, dh;
          heads_for(ch, heads, dh);
          b.push_back(TLayer{TL_ST, add_st(pfx + "1.", ch, heads, dh)});
        }
        e->input_blocks.push_back(b);
        chans.push_back(ch);
      }
      if (level != c.n_channel_mult - 1) {
        const std::string pfx = "input_blocks." + std::to_string(e->input_blocks.size()) + ".0.op";
        e->input_blocks.push_back(TBlock{TLayer{TL_DOWN, add_samp(pfx, ch)}});
        chans.push_back(ch);
        ds *= 2;
      }
    }
    {
      int heads, dh;
      heads_for(ch, heads, dh);
      e->middle.push_back(TLayer{TL_RES, add_res("middle_block.0.", ch, ch, emb_cols)});
      e->middle.push_back(TLayer{TL_ST, add_st("middle_block.1.", ch, heads, dh)});
      e->middle.push_back(TLayer{TL_RES, add_res("middle_block.2.", ch, ch, emb_cols)});
    }
    for (int level = c.n_channel_mult - 1; level >= 0; --level) {
      const int mult = c.channel_mult[level];
      for (int i = 0; i < c.num_res_blocks + 1; ++i) {
        const int ich = chans.back();
        chans.pop_back();
        const std::string pfx = "output_blocks." + std::to_string(e->output_blocks.size()) + ".";
        TBlock b;
        int li = 0;
        b.push_back(TLayer{TL_RES, add_res(pfx + std::to_string(li++) + ".", ch + ich, mc * mult, emb_cols)});
        ch = mc * mult;
        if (in_attn_res(ds)) {
          int heads, dh;
          heads_for(ch, heads, dh);
          b.push_back(TLayer{TL_ST, add_st(pfx + std::to_string(li++) + ".", ch, heads, dh)});
        }
        if (level && i == c.num_res_blocks) {
          b.push_back(TLayer{TL_UP, add_samp(pfx + std::to_string(li++) + ".conv", ch)});
          ds /= 2;
        }
        e->output_blocks.push_back(b);
      }
    }
    e->out_gn = norm("out.0", ch);
    {
      TConv& o = e->conv_out;
      o.w = "out.2.weight";
      o.b = "out.2.bias";
      o.Cout = c.out_channels;
      o.Cin = ch;
      o.Ktot = 9 * ch;
      expect(o.w, static_cast<int64_t>(c.out_channels) * ch * 9);
      expect(o.b, c.out_channels);
      o.pw = A.alloc<bf16>(static_cast<size_t>(GEMM_BLOCK_N_OUT) * 9 * ch);  // rows 0..3 hi, 4..7 lo (zero elsewhere)
      job(PK_CONV3, o.w, o.pw, c.out_channels, ch, 9 * ch, 0, 2);
      o.cout_pad = 64;
      o.pwt = A.alloc<bf16>(static_cast<size_t>(ch) * 9 * 64);
      job(PK_CONV3_T, o.w, o.pwt, c.out_channels, ch, 64, 0, 0);
    }
    // fused emb_layers projection
    e->emb_cols = emb_cols;
    e->emb_all_w = A.alloc<bf16>(static_cast<size_t>(emb_cols) * ted);
    e->emb_all_wt = A.alloc<bf16>(static_cast<size_t>(emb_cols) * ted);
    e->emb_all_b = A.alloc<float>(emb_cols);
    for (const TRes& r : e->res) {
      job(PK_LIN, r.emb_w, e->emb_all_w, r.Cout, ted, ted, 0, r.emb_off);
      job(PK_LIN_T, r.emb_w, e->emb_all_wt, r.Cout, ted, emb_cols, r.emb_off, 0);
      job(PK_VEC, r.emb_b, e->emb_all_b, r.Cout, 0, 0, r.emb_off, 0);
    }
    // fused context K/V projection of every cross-attention
    const int inner = e->st.empty() ? 0 : e->st[0].heads * e->st[0].dh;
    e->kv_cols = e->n_kv * 2 * inner;
    e->kv_all_w = A.alloc<bf16>(static_cast<size_t>(e->kv_cols) * D);
    e->kv_all_wt = A.alloc<bf16>(static_cast<size_t>(e->kv_cols) * D);
    if (!dry)
      for (size_t i = 0; i < e->kv_names.size(); ++i) {
        job(PK_LIN, e->kv_names[i], e->kv_all_w, inner, D, D, 0, static_cast<int>(i) * inner);
        job(PK_LIN_T, e->kv_names[i], e->kv_all_wt, inner, D, e->kv_cols, static_cast<int>(i) * inner, 0);
      }
  }
};

// ----------------------------------------------------------------------------------------------
// plan builder
// ----------------------------------------------------------------------------------------------
struct GSrc {
  const bf16* p;
  int C, ld, taps, stride, H, W;
};
struct GEpi {
  const float* bias = nullptr;
  const float* rowbias = nullptr;
  int rb_ld = 0;
  int rows_per_sample = 1;
  const bf16* residual = nullptr;
  int res_ld = 0;
  void* out = nullptr;
  int out_ld = 0;
  int out_f32 = 0;
  float* gn_partial = nullptr;
  int epi = EPI_STD;
  int patch_y = 0;
};

struct TPlanBuilder {
  wd_trainer* e;
  TPlan* plan;
  TArena A;
  bool dry;
  int B;
  std::string err;
  std::vector<std::function<bool()>> tape;  // backward emitters, run in reverse order

  const float* W(const std::string& n) { return dry ? nullptr : e->params[n].w; }
  std::vector<std::string> g_touched;  // gradients named by the backward emitter that is running
  float* G(const std::string& n) {
    g_touched.push_back(n);
    return dry ? nullptr : e->params[n].g;
  }

  TT* new_t(int H, int Wd, int C, bool with_grad = true) {
    plan->tensors.emplace_back();
    TT* t = &plan->tensors.back();
    t->p = A.alloc<bf16>(static_cast<size_t>(B) * H * Wd * C);
    if (with_grad) t->g = A.alloc<bf16>(static_cast<size_t>(B) * H * Wd * C);
    t->stats = A.alloc<float>(static_cast<size_t>(B) * 32 * 8 * 2);
    t->C = C;
    t->H = H;
    t->W = Wd;
    return t;
  }
  static bool epilogue_stats_ok(int HW, int C) { return C % 32 == 0 && C / 32 == 10 && HW % 32 == 0 && HW / 32 <= 8; }

  // ---- tcgen05 GEMM / implicit-GEMM conv (forward kernels; also every data gradient) ----
  bool gemm(std::vector<TOp>& ops, const char* what, int M, bool conv, int Hout, int Wout, const std::vector<GSrc>& srcs,
            const bf16* w, int N, int K, const GEpi& ep) {
    GemmLaunch L;
    memset(&L, 0, sizeof(L));
    GemmArgs& a = L.args;
    a.M = M;
    a.N = N;
    a.num_src = static_cast<int>(srcs.size());
    a.conv = conv ? 1 : 0;
    a.Wout = conv ? Wout : 1;
    a.HWout = conv ? Hout * Wout : 1;
    a.bias = ep.bias;
    a.rowbias = ep.rowbias;
    a.rb_ld = ep.rb_ld;
    a.rows_per_sample = ep.rows_per_sample;
    a.residual = ep.residual;
    a.res_ld = ep.res_ld;
    a.out = ep.out;
    a.out_ld = ep.out_ld;
    a.out_f32 = ep.out_f32;
    a.ln_eps = 1e-5f;
    a.epi = ep.epi;
    a.gn_partial = ep.gn_partial;
    a.gn_cpg = ep.gn_partial ? 10 : 0;
    int ktot = 0;
    if (srcs.empty() || srcs.size() > GEMM_MAX_SRC) { err = std::string(what) + ": bad source count"; return false; }
    for (size_t i = 0; i < srcs.size(); ++i) {
      const GSrc& s = srcs[i];
      if (s.C % GEMM_BLOCK_K) { err = std::string(what) + ": source channels must be a multiple of 64"; return false; }
      a.taps[i] = s.taps;
      a.chunks[i] = s.C / GEMM_BLOCK_K;
      a.stride[i] = s.stride;
      ktot += s.taps * s.C;
      if (dry) continue;
      bool ok;
      if (!conv) {
        ok = tmap_encode_2d_bf16(&L.mapA[i], s.p, s.C, M, s.ld, GEMM_BLOCK_K, GEMM_BLOCK_M);
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
        ok = tmap_encode_4d_bf16(&L.mapA[i], s.p, s.C, s.W, s.H, B, s.ld, GEMM_BLOCK_K, bw, bh, bnn, s.stride);
      }
      if (!ok) { err = std::string(what) + ": cuTensorMapEncodeTiled failed (A)"; return false; }
    }
    if (ktot != K) { err = std::string(what) + ": K mismatch"; return false; }
    const int bn = (ep.epi == EPI_SAMPLER) ? GEMM_BLOCK_N_OUT : gemm_tc_block_n();
    if (N % bn) { err = std::string(what) + ": N must be a multiple of the N tile"; return false; }
    if (!dry) {
      for (size_t i = srcs.size(); i < GEMM_MAX_SRC; ++i) L.mapA[i] = L.mapA[0];
      if (!tmap_encode_2d_bf16(&L.mapB, w, K, N, K, GEMM_BLOCK_K, gemm_b_box_rows(a))) { err = "tensor map B"; return false; }
      L.mapOut = L.mapB;
      L.mapRes = L.mapB;
      if (ep.epi != EPI_SAMPLER && !ep.out_f32 && !tmap_encode_out_bf16(&L.mapOut, ep.out, N, M, ep.out_ld)) { err = "tensor map out"; return false; }
      if (ep.residual && !tmap_encode_out_bf16(&L.mapRes, ep.residual, N, M, ep.res_ld)) { err = "tensor map residual"; return false; }
      const int patch_y = ep.patch_y;
      const bool sampler = ep.epi == EPI_SAMPLER;
      ops.push_back(TOp{what, [L, patch_y, sampler](const TRun& r, cudaStream_t s) {
        if (!patch_y && !sampler) return gemm_tc_launch(L, s);
        GemmLaunch L2 = L;
        if (patch_y) L2.args.rowbias_idx = r.y;
        if (sampler) {
          L2.args.eps_out = r.eps_out;
          L2.args.mode = STEP_EPS_ONLY;
        }
        return gemm_tc_launch(L2, s);
      }});
    }
    return true;
  }
  // plain linear forward / data gradient: out[M, N] = A[M, K] w[N, K]^T (+bias) (+residual)
  bool lin(std::vector<TOp>& ops, const char* what, const bf16* a, int a_ld, int M, const bf16* w, int N, int K, bf16* out,
           int out_ld, const float* bias = nullptr, const bf16* residual = nullptr, int res_ld = 0) {
    GEpi ep;
    ep.bias = bias;
    ep.out = out;
    ep.out_ld = out_ld;
    ep.residual = residual;
    ep.res_ld = res_ld;
    return gemm(ops, what, M, false, 0, 0, {GSrc{a, K, a_ld, 1, 1, 1, 1}}, w, N, K, ep);
  }

  // ---- tcgen05 weight gradient ----
  struct WX { const bf16* p; int C, ld; bool conv; int taps, stride, H, W; };  // X operand (H, W: its own spatial size)
  bool wgrad(std::vector<TOp>& ops, const char* what, const WX& x, const bf16* dy, int dy_ld, int dy_cols, int M, int Hout,
             int Wout, const std::vector<float*>& dst, long long sN, long long sC, long long sT, int bn = 320, int n_valid = 0) {
    if (dy_cols % bn) { err = std::string(what) + ": dY columns must be a multiple of the wgrad N tile"; return false; }
    const int groups_total = dy_cols / bn;
    if (!dry && static_cast<int>(dst.size()) != groups_total) { err = std::string(what) + ": one gradient tensor per column group"; return false; }
    if (x.C < WG_BLOCK_C || x.C % 64) { err = std::string(what) + ": X channels"; return false; }
    if (dry) return true;
    for (int g0 = 0; g0 < groups_total; g0 += WG_MAX_GROUPS) {
      const int ng = std::min(WG_MAX_GROUPS, groups_total - g0);
      WgradLaunch L;
      memset(&L, 0, sizeof(L));
      L.bn = bn;
      WgradArgs& a = L.args;
      a.M = M;
      a.Cin = x.C;
      a.taps = x.taps;
      a.conv = x.conv ? 1 : 0;
      a.stride = x.stride;
      a.HWout = x.conv ? Hout * Wout : 1;
      a.Wout = x.conv ? Wout : 1;
      a.n_groups = ng;
      a.n_valid = n_valid ? n_valid : bn;
      for (int g = 0; g < ng; ++g) a.dst[g] = dst[g0 + g];
      a.sN = sN;
      a.sC = sC;
      a.sT = sT;
      const int cin_tiles = (x.C + WG_BLOCK_C - 1) / WG_BLOCK_C;
      a.splits = wgrad_pick_splits(M, ng * cin_tiles * x.taps);
      bool ok;
      if (!x.conv) {
        ok = tmap_encode_2d_bf16(&L.mapX, x.p, x.C, M, x.ld, 64, WG_BLOCK_TOK);
      } else {
        const int HWo = Hout * Wout;
        uint32_t bw, bh, bnn;
        if (HWo >= WG_BLOCK_TOK) {
          if (HWo % WG_BLOCK_TOK || WG_BLOCK_TOK % Wout) { err = "wgrad conv: unsupported spatial size"; return false; }
          bw = Wout * x.stride;
          bh = (WG_BLOCK_TOK / Wout) * x.stride;
          bnn = 1;
        } else {
          if (WG_BLOCK_TOK % HWo) { err = "wgrad conv: unsupported spatial size"; return false; }
          bw = x.W;
          bh = x.H;
          bnn = WG_BLOCK_TOK / HWo;
        }
        ok = tmap_encode_4d_bf16(&L.mapX, x.p, x.C, x.W, x.H, B, x.ld, 64, bw, bh, bnn, x.stride);
      }
      if (!ok) { err = std::string(what) + ": tensor map X"; return false; }
      if (!tmap_encode_2d_bf16(&L.mapDY, dy + static_cast<size_t>(g0) * bn, static_cast<uint64_t>(ng) * bn, M, dy_ld, 64, WG_BLOCK_TOK)) {
        err = std::string(what) + ": tensor map dY";
        return false;
      }
      ops.push_back(TOp{what, [L](const TRun&, cudaStream_t s) { return wgrad_tc_launch(L, s); }});
    }
    return true;
  }
  // nn.Linear weight gradient into the state_dict tensor [N, K] (N a multiple of 320)
  bool wgrad_lin(std::vector<TOp>& ops, const char* what, const bf16* x, int x_ld, int K, const bf16* dy, int dy_ld, int N, int M,
                 float* gw) {
    std::vector<float*> dst;
    for (int g = 0; g < N / 320; ++g) dst.push_back(gw ? gw + static_cast<size_t>(g) * 320 * K : nullptr);
    return wgrad(ops, what, WX{x, K, x_ld, false, 1, 1, 1, 1}, dy, dy_ld, N, M, 0, 0, dst, K, 1, 0);
  }
  void colsum(std::vector<TOp>& ops, const bf16* dy, int ld, int N, int M, float* total, int rows_per_group = 64,
              bf16* per_group = nullptr, int pg_ld = 0) {
    if (dry) return;
    const int groups = (M + rows_per_group - 1) / rows_per_group;
    ops.push_back(TOp{"colsum", [=](const TRun&, cudaStream_t s) {
      // (rows_total is passed through groups * rows_per_group; a ragged tail only occurs for rows_per_group == 256 blocks)
      return colsum_launch_ragged(dy, ld, N, groups, rows_per_group, M, total, per_group, pg_ld, s);
    }});
  }
  static cudaError_t colsum_launch_ragged(const bf16* dy, int ld, int N, int groups, int rows_per_group, int M, float* total,
                                          bf16* per_group, int pg_ld, cudaStream_t s) {
    if (groups * rows_per_group == M) return colsum_launch(dy, ld, N, groups, rows_per_group, total, per_group, pg_ld, s);
    // full groups, then the tail as one smaller group
    const int full = M / rows_per_group;
    cudaError_t e = cudaSuccess;
    if (full > 0) e = colsum_launch(dy, ld, N, full, rows_per_group, total, per_group, pg_ld, s);
    if (e != cudaSuccess) return e;
    const int tail = M - full * rows_per_group;
    return colsum_launch(dy + static_cast<size_t>(full) * rows_per_group * ld, ld, N, 1, tail, total,
                         per_group ? per_group + static_cast<size_t>(full) * pg_ld : nullptr, pg_ld, s);
  }

  // ---- GroupNorm forward (ops.cu kernel) over the concat of `srcs`; returns the output tensor ----
  bool ensure_stats(std::vector<TOp>& ops, TT* a) {
    if (a->pslots) return true;
    a->pslots = groupnorm_stats_slots(a->H * a->W);
    if (dry) return true;
    GroupNormStatsArgs gs{a->p, a->C, a->stats, a->H * a->W, a->C, a->C / 32, a->pslots, 0};
    const int Bc = B;
    ops.push_back(TOp{"groupnorm_stats", [gs, Bc](const TRun&, cudaStream_t s) { return groupnorm_stats_launch(gs, Bc, s); }});
    return true;
  }
  struct GNRec { GroupNormArgs fa; int nslab; };
  bool gn_fwd(std::vector<TOp>& ops, const std::vector<TT*>& srcs, const TNorm& nw, float eps, int silu, TT*& out, GroupNormArgs& fa) {
    int totalC = 0;
    for (TT* s : srcs) totalC += s->C;
    if (totalC != nw.C || totalC % 32) { err = "groupnorm: channel mismatch"; return false; }
    const int cpg = totalC / 32;
    const int H = srcs[0]->H, Wd = srcs[0]->W, HW = H * Wd;
    const int Cs = srcs[0]->C;
    for (TT* s : srcs)
      if (s->C != Cs) { err = "groupnorm: concat sources must have equal channels"; return false; }
    if (Cs % cpg || Cs % 8 || srcs.size() > 2) { err = "groupnorm: unsupported slab layout"; return false; }
    out = new_t(H, Wd, totalC);
    memset(&fa, 0, sizeof(fa));
    for (size_t i = 0; i < srcs.size(); ++i) {
      fa.x[i] = srcs[i]->p;
      fa.x_ld[i] = srcs[i]->C;
      if (!srcs[i]->pslots) { err = "groupnorm: source tensor carries no statistics"; return false; }
      fa.partial[i] = srcs[i]->stats;
      fa.pslots[i] = srcs[i]->pslots;
    }
    fa.out = out->p;
    fa.out_ld = totalC;
    fa.gamma = W(nw.g);
    fa.beta = W(nw.b);
    fa.HW = HW;
    fa.Cs = Cs;
    fa.cpg = cpg;
    fa.pcpg = Cs / 32;
    fa.eps = eps;
    fa.silu = silu;
    fa.nchunk = groupnorm_apply_chunks(HW);
    if (!dry) {
      const GroupNormArgs fa_c = fa;
      const int Bc = B, ns = static_cast<int>(srcs.size());
      ops.push_back(TOp{"groupnorm", [fa_c, Bc, ns](const TRun&, cudaStream_t s) { return groupnorm_launch(fa_c, Bc, ns, s); }});
    }
    return true;
  }
  // backward of gn_fwd: dx[s] (+)= GN'(x) out.g + add[s]
  bool gn_bwd(std::vector<TOp>& ops, const std::vector<TT*>& srcs, const TNorm& nw, const GroupNormArgs& fa, TT* out,
              const bf16* add0, int add0_ld, const bf16* add1, int add1_ld, float* ws) {
    GroupNormBwdArgs a;
    memset(&a, 0, sizeof(a));
    for (size_t i = 0; i < srcs.size(); ++i) {
      a.x[i] = fa.x[i];
      a.x_ld[i] = fa.x_ld[i];
      a.partial[i] = fa.partial[i];
      a.pslots[i] = fa.pslots[i];
      a.dx[i] = srcs[i]->g;
      a.dx_ld[i] = srcs[i]->C;
      a.accumulate[i] = srcs[i]->g_init ? 1 : 0;
      srcs[i]->g_init = true;
    }
    a.add[0] = add0;
    a.add_ld[0] = add0_ld;
    a.add[1] = add1;
    a.add_ld[1] = add1_ld;
    a.dy = out->g;
    a.dy_ld = fa.out_ld;
    a.gamma = fa.gamma;
    a.beta = fa.beta;
    a.ws = ws;
    a.dgamma = G(nw.g);
    a.dbeta = G(nw.b);
    a.HW = fa.HW;
    a.Cs = fa.Cs;
    a.cpg = fa.cpg;
    a.pcpg = fa.pcpg;
    a.eps = fa.eps;
    a.silu = fa.silu;
    if (!dry) {
      const int Bc = B, ns = static_cast<int>(srcs.size());
      ops.push_back(TOp{"groupnorm_bwd", [a, Bc, ns](const TRun&, cudaStream_t s) { return groupnorm_bwd_launch(a, Bc, ns, s); }});
    }
    return true;
  }

  // scratch shared by the backward emitters (the backward pass is one stream: lifetimes never overlap across layers)
  bf16 *scr_a = nullptr, *scr_b = nullptr, *scr_c = nullptr, *scr_wide = nullptr, *scr_mid = nullptr, *scr_cat = nullptr;
  float* gn_ws = nullptr;
  bf16* d_kv_all = nullptr;
  bf16* d_emb_out = nullptr;
  bf16* kv_all = nullptr;

  // ---- ResBlock (unet.py:646-671) ----
  bool res_block(const TRes& r, const std::vector<TT*>& in, const float* emb_out, int emb_ld, TT*& out) {
    auto& F = plan->fwd;
    const int H = in[0]->H, Wd = in[0]->W, HW = H * Wd, M = B * HW;
    TT* a1;
    GroupNormArgs fa1, fa2;
    if (!gn_fwd(F, in, r.gn1, 1e-5f, 1, a1, fa1)) return false;
    TT* h2 = new_t(H, Wd, r.Cout);
    {
      GEpi ep;
      ep.bias = W(r.conv1.b);
      ep.rowbias = emb_out + r.emb_off;
      ep.rb_ld = emb_ld;
      ep.rows_per_sample = HW;
      ep.out = h2->p;
      ep.out_ld = r.Cout;
      if (epilogue_stats_ok(HW, r.Cout)) { ep.gn_partial = h2->stats; h2->pslots = HW / 32; }
      if (!gemm(F, "res.conv1", M, true, H, Wd, {GSrc{a1->p, a1->C, a1->C, 9, 1, H, Wd}}, r.conv1.pw, r.Cout, r.conv1.Ktot, ep)) return false;
      if (!ensure_stats(F, h2)) return false;
    }
    TT* a2;
    if (!gn_fwd(F, {h2}, r.gn2, 1e-5f, 1, a2, fa2)) return false;
    out = new_t(H, Wd, r.Cout);
    {
      GEpi ep;
      ep.bias = r.bias2;
      ep.rows_per_sample = HW;
      ep.out = out->p;
      ep.out_ld = r.Cout;
      if (epilogue_stats_ok(HW, r.Cout)) { ep.gn_partial = out->stats; out->pslots = HW / 32; }
      std::vector<GSrc> srcs{GSrc{a2->p, a2->C, a2->C, 9, 1, H, Wd}};
      if (r.skip_conv) {
        for (TT* s : in) srcs.push_back(GSrc{s->p, s->C, s->C, 1, 1, H, Wd});
      } else {
        if (in.size() != 1 || in[0]->C != r.Cout) { err = "resblock: identity skip needs a single source"; return false; }
        ep.residual = in[0]->p;
        ep.res_ld = in[0]->C;
      }
      if (!gemm(F, "res.conv2", M, true, H, Wd, srcs, r.conv2.pw, r.Cout, r.conv2.Ktot, ep)) return false;
      if (!ensure_stats(F, out)) return false;
    }
    // ---------------- backward ----------------
    const std::vector<TT*> in_c = in;
    tape.push_back([this, r, in_c, a1, h2, a2, out, fa1, fa2, H, Wd, HW, M]() -> bool {
      auto& Bk = plan->bwd;
      const bf16* dout = out->g;
      // conv2 (+ skip conv): weight / bias gradients
      if (!wgrad(Bk, "res.conv2.wgrad", WX{a2->p, a2->C, a2->C, true, 9, 1, H, Wd}, dout, r.Cout, r.Cout, M, H, Wd,
                 {G(r.conv2.w)}, static_cast<long long>(r.Cout) * 9, 9, 1))
        return false;
      colsum(Bk, dout, r.Cout, r.Cout, M, G(r.conv2.b));
      const bf16* add0 = nullptr;
      const bf16* add1 = nullptr;
      int add_ld = 0;
      if (r.skip_conv) {
        int coff = 0;
        for (TT* s : in_c) {
          if (!wgrad(Bk, "res.skip.wgrad", WX{s->p, s->C, s->C, false, 1, 1, 1, 1}, dout, r.Cout, r.Cout, M, 0, 0,
                     {dry ? nullptr : G(r.skip.w) + coff}, r.Cin, 1, 0))
            return false;
          coff += s->C;
        }
        colsum(Bk, dout, r.Cout, r.Cout, M, G(r.skip.b));
        // d(cat input) through the 1x1 skip conv: [M, Cin]
        if (!lin(Bk, "res.skip.dgrad", dout, r.Cout, M, r.skip.pwt, r.Cin, r.Cout, scr_cat, r.Cin)) return false;
        add0 = scr_cat;
        add1 = in_c.size() > 1 ? scr_cat + in_c[0]->C : nullptr;
        add_ld = r.Cin;
      } else {
        add0 = dout;
        add_ld = r.Cout;
      }
      // d a2 = conv2^T(d out)
      {
        GEpi ep;
        ep.out = a2->g;
        ep.out_ld = r.Cout;
        if (!gemm(Bk, "res.conv2.dgrad", M, true, H, Wd, {GSrc{dout, r.Cout, r.Cout, 9, 1, H, Wd}}, r.conv2.pwt, r.Cout, 9 * r.Cout, ep))
          return false;
      }
      if (!gn_bwd(Bk, {h2}, r.gn2, fa2, a2, nullptr, 0, nullptr, 0, gn_ws)) return false;
      // conv1: weight / bias / timestep-embedding gradients
      {
        int coff = 0;
        // A of conv1 is the GroupNorm output a1 over the concatenated channels
        if (!wgrad(Bk, "res.conv1.wgrad", WX{a1->p, a1->C, a1->C, true, 9, 1, H, Wd}, h2->g, r.Cout, r.Cout, M, H, Wd,
                   {G(r.conv1.w)}, static_cast<long long>(r.Cin) * 9, 9, 1))
          return false;
        (void)coff;
      }
      colsum(Bk, h2->g, r.Cout, r.Cout, M, G(r.conv1.b), HW, d_emb_out + r.emb_off, e->emb_cols);
      {
        GEpi ep;
        ep.out = a1->g;
        ep.out_ld = r.Cin;
        if (!gemm(Bk, "res.conv1.dgrad", M, true, H, Wd, {GSrc{h2->g, r.Cout, r.Cout, 9, 1, H, Wd}}, r.conv1.pwt, r.Cin, 9 * r.Cout, ep))
          return false;
      }
      return gn_bwd(Bk, in_c, r.gn1, fa1, a1, add0, add_ld, add1, add_ld, gn_ws);
    });
    return true;
  }

  // ---- SpatialTransformer, unet.py variant (unet.py:381-412,337-345) ----
  bool st_block(const TST& s, TT* x_in, TT*& out) {
    auto& F = plan->fwd;
    const int H = x_in->H, Wd = x_in->W, HW = H * Wd, M = B * HW, C = s.heads * s.dh;
    const int L = plan->L, KVLD = e->kv_cols;
    const float scale = 1.0f / sqrtf(static_cast<float>(s.dh));
    TT* g;
    GroupNormArgs fag;
    if (!gn_fwd(F, {x_in}, s.gn, 1e-6f, 0, g, fag)) return false;
    TT* x = new_t(H, Wd, C);
    if (!lin(F, "st.proj_in", g->p, g->C, M, s.proj_in.pw, C, s.C, x->p, C, W(s.proj_in.b))) return false;
    struct BlkRec { TT *x0, *x1, *x2, *x3; bf16 *n1, *q1, *o1, *n2, *q2, *o2, *n3, *p, *gg; };
    std::vector<BlkRec> recs;
    for (const TBlk& t : s.blocks) {
      BlkRec r;
      r.x0 = x;
      auto ln = [&](const char* what, TT* src, const TNorm& nw, bf16* dst) {
        if (dry) return;
        const bf16* xp = src->p;
        const float* gw = W(nw.g);
        const float* bw = W(nw.b);
        F.push_back(TOp{what, [=](const TRun&, cudaStream_t st) { return layernorm_launch(xp, dst, gw, bw, M, C, 1e-5f, 0, st); }});
      };
      auto attn = [&](const bf16* q, int kvi, bf16* o) {
        if (dry) return;
        AttnFlashArgs af{q, C, kv_all + static_cast<size_t>(kvi) * 2 * C, kv_all + static_cast<size_t>(kvi) * 2 * C + C, KVLD, o, C, HW, L,
                         s.heads, scale};
        const int Bc = B;
        F.push_back(TOp{"st.attn", [af, Bc](const TRun&, cudaStream_t st) { return attn_flash_launch(af, Bc, st); }});
      };
      r.n1 = A.alloc<bf16>(static_cast<size_t>(M) * C);
      r.q1 = A.alloc<bf16>(static_cast<size_t>(M) * C);
      r.o1 = A.alloc<bf16>(static_cast<size_t>(M) * C);
      ln("st.ln2a", r.x0, t.ln2, r.n1);  // unet.py:337 applies norm2 before attn1
      if (!lin(F, "st.q1", r.n1, C, M, t.q1.pw, C, C, r.q1, C)) return false;
      attn(r.q1, t.kv1, r.o1);
      r.x1 = new_t(H, Wd, C);
      if (!lin(F, "st.o1", r.o1, C, M, t.o1.pw, C, C, r.x1->p, C, W(t.o1.b), r.x0->p, C)) return false;
      r.n2 = A.alloc<bf16>(static_cast<size_t>(M) * C);
      r.q2 = A.alloc<bf16>(static_cast<size_t>(M) * C);
      r.o2 = A.alloc<bf16>(static_cast<size_t>(M) * C);
      ln("st.ln2b", r.x1, t.ln2, r.n2);
      if (!lin(F, "st.q2", r.n2, C, M, t.q2.pw, C, C, r.q2, C)) return false;
      attn(r.q2, t.kv2, r.o2);
      r.x2 = new_t(H, Wd, C);
      if (!lin(F, "st.o2", r.o2, C, M, t.o2.pw, C, C, r.x2->p, C, W(t.o2.b), r.x1->p, C)) return false;
      r.n3 = A.alloc<bf16>(static_cast<size_t>(M) * C);
      r.p = A.alloc<bf16>(static_cast<size_t>(M) * 8 * C);
      r.gg = A.alloc<bf16>(static_cast<size_t>(M) * 4 * C);
      ln("st.ln3", r.x2, t.ln3, r.n3);
      if (!lin(F, "st.ff_proj", r.n3, C, M, t.ffp.pw, 8 * C, C, r.p, 8 * C, W(t.ffp.b))) return false;
      if (!dry) {
        const bf16* pp = r.p;
        bf16* gg = r.gg;
        F.push_back(TOp{"st.geglu", [=](const TRun&, cudaStream_t st) { return geglu_fwd_launch(pp, gg, M, 4 * C, st); }});
      }
      r.x3 = new_t(H, Wd, C);
      if (!lin(F, "st.ff_out", r.gg, 4 * C, M, t.ffo.pw, C, 4 * C, r.x3->p, C, W(t.ffo.b), r.x2->p, C)) return false;
      x = r.x3;
      recs.push_back(r);
    }
    out = new_t(H, Wd, s.C);
    {
      GEpi ep;
      ep.bias = W(s.proj_out.b);
      ep.out = out->p;
      ep.out_ld = s.C;
      ep.residual = x_in->p;
      ep.res_ld = x_in->C;
      ep.rows_per_sample = HW;
      if (epilogue_stats_ok(HW, s.C)) { ep.gn_partial = out->stats; out->pslots = HW / 32; }
      if (!gemm(F, "st.proj_out", M, false, 0, 0, {GSrc{x->p, C, C, 1, 1, 1, 1}}, s.proj_out.pw, s.C, C, ep)) return false;
      if (!ensure_stats(F, out)) return false;
    }
    // ---------------- backward ----------------
    TT* x_last = x;
    tape.push_back([this, s, x_in, g, fag, recs, out, x_last, H, Wd, HW, M, C, L, KVLD, scale]() -> bool {
      auto& Bk = plan->bwd;
      const bf16* dout = out->g;
      if (!wgrad_lin(Bk, "st.proj_out.wgrad", x_last->p, C, C, dout, s.C, s.C, M, G(s.proj_out.w))) return false;
      colsum(Bk, dout, s.C, s.C, M, G(s.proj_out.b));
      if (!lin(Bk, "st.proj_out.dgrad", dout, s.C, M, s.proj_out.pwt, C, s.C, x_last->g, C)) return false;
      x_last->g_init = true;
      auto ln_bwd = [&](const char* what, TT* src, const TNorm& nw, const bf16* dy, const bf16* add, bf16* dx) {
        if (dry) return;
        const bf16* xp = src->p;
        const float* gw = W(nw.g);
        float* dgm = G(nw.g);
        float* dbt = G(nw.b);
        Bk.push_back(TOp{what, [=](const TRun&, cudaStream_t st) { return layernorm_bwd_launch(xp, dy, gw, add, dx, dgm, dbt, M, C, 1e-5f, st); }});
      };
      auto attn_bwd = [&](const bf16* q, int kvi, const bf16* d_o, bf16* dq) {
        if (dry) return;
        AttnSmallBwdArgs a;
        a.q = q; a.q_ld = C;
        a.k = kv_all + static_cast<size_t>(kvi) * 2 * C;
        a.v = a.k + C;
        a.kv_ld = KVLD;
        a.dout = d_o; a.do_ld = C;
        a.dq = dq; a.dq_ld = C;
        a.dk = d_kv_all + static_cast<size_t>(kvi) * 2 * C;
        a.dv = a.dk + C;
        a.dkv_ld = KVLD;
        a.Sq = HW; a.L = L; a.heads = s.heads; a.scale = scale;
        const int Bc = B;
        Bk.push_back(TOp{"st.attn_bwd", [a, Bc](const TRun&, cudaStream_t st) { return attn_small_bwd_launch(a, Bc, st); }});
      };
      for (int bi = static_cast<int>(recs.size()) - 1; bi >= 0; --bi) {
        const BlkRec& r = recs[bi];
        const TBlk& t = s.blocks[bi];
        const bf16* dx3 = r.x3->g;
        // ---- feed-forward ----
        if (!wgrad_lin(Bk, "st.ff_out.wgrad", r.gg, 4 * C, 4 * C, dx3, C, C, M, G(t.ffo.w))) return false;
        colsum(Bk, dx3, C, C, M, G(t.ffo.b));
        if (!lin(Bk, "st.ff_out.dgrad", dx3, C, M, t.ffo.pwt, 4 * C, C, scr_mid, 4 * C)) return false;
        if (!dry) {
          const bf16* pp = r.p;
          const bf16* dgg = scr_mid;
          bf16* dp = scr_wide;
          Bk.push_back(TOp{"st.geglu_bwd", [=](const TRun&, cudaStream_t st) { return geglu_bwd_launch(pp, dgg, dp, M, 4 * C, st); }});
        }
        if (!wgrad_lin(Bk, "st.ff_proj.wgrad", r.n3, C, C, scr_wide, 8 * C, 8 * C, M, G(t.ffp.w))) return false;
        colsum(Bk, scr_wide, 8 * C, 8 * C, M, G(t.ffp.b));
        if (!lin(Bk, "st.ff_proj.dgrad", scr_wide, 8 * C, M, t.ffp.pwt, C, 8 * C, scr_a, C)) return false;
        ln_bwd("st.ln3_bwd", r.x2, t.ln3, scr_a, dx3, r.x2->g);
        r.x2->g_init = true;
        // ---- attn2 ----
        if (!wgrad_lin(Bk, "st.o2.wgrad", r.o2, C, C, r.x2->g, C, C, M, G(t.o2.w))) return false;
        colsum(Bk, r.x2->g, C, C, M, G(t.o2.b));
        if (!lin(Bk, "st.o2.dgrad", r.x2->g, C, M, t.o2.pwt, C, C, scr_a, C)) return false;
        attn_bwd(r.q2, t.kv2, scr_a, scr_b);
        if (!wgrad_lin(Bk, "st.q2.wgrad", r.n2, C, C, scr_b, C, C, M, G(t.q2.w))) return false;
        if (!lin(Bk, "st.q2.dgrad", scr_b, C, M, t.q2.pwt, C, C, scr_c, C)) return false;
        ln_bwd("st.ln2b_bwd", r.x1, t.ln2, scr_c, r.x2->g, r.x1->g);
        r.x1->g_init = true;
        // ---- attn1 ----
        if (!wgrad_lin(Bk, "st.o1.wgrad", r.o1, C, C, r.x1->g, C, C, M, G(t.o1.w))) return false;
        colsum(Bk, r.x1->g, C, C, M, G(t.o1.b));
        if (!lin(Bk, "st.o1.dgrad", r.x1->g, C, M, t.o1.pwt, C, C, scr_a, C)) return false;
        attn_bwd(r.q1, t.kv1, scr_a, scr_b);
        if (!wgrad_lin(Bk, "st.q1.wgrad", r.n1, C, C, scr_b, C, C, M, G(t.q1.w))) return false;
        if (!lin(Bk, "st.q1.dgrad", scr_b, C, M, t.q1.pwt, C, C, scr_c, C)) return false;
        ln_bwd("st.ln2a_bwd", r.x0, t.ln2, scr_c, r.x1->g, r.x0->g);
        r.x0->g_init = true;
      }
      TT* x0 = recs.empty() ? x_last : recs[0].x0;
      if (!wgrad_lin(Bk, "st.proj_in.wgrad", g->p, s.C, s.C, x0->g, C, C, M, G(s.proj_in.w))) return false;
      colsum(Bk, x0->g, C, C, M, G(s.proj_in.b));
      if (!lin(Bk, "st.proj_in.dgrad", x0->g, C, M, s.proj_in.pwt, s.C, C, g->g, s.C)) return false;
      return gn_bwd(Bk, {x_in}, s.gn, fag, g, dout, s.C, nullptr, 0, gn_ws);
    });
    return true;
  }

  bool build() {
    const wd_config& c = e->cfg;
    const int mc = c.model_channels, ted = e->time_dim, D = c.context_dim;
    const int L = plan->L;
    auto& F = plan->fwd;
    const int H0 = c.latent_h, W0 = c.latent_w, HW0 = H0 * W0, M0 = B * HW0;
    const int Cmax = 2 * mc * 4;  // widest concat (bounded below by what the plan needs; checked by construction sizes)
    (void)Cmax;

    plan->in_x = A.alloc<float>(static_cast<size_t>(M0) * c.in_channels);
    plan->out_eps = A.alloc<float>(static_cast<size_t>(M0) * c.out_channels);
    plan->in_deps = A.alloc<float>(static_cast<size_t>(M0) * c.out_channels);
    plan->in_t = A.alloc<long long>(B);
    plan->in_y = A.alloc<long long>(B);
    plan->in_ctx = A.alloc<long long>(static_cast<size_t>(B) * L);
    // ---------------- shared scratch ----------------
    const size_t tokC = static_cast<size_t>(M0) * mc;
    scr_a = A.alloc<bf16>(tokC);
    scr_b = A.alloc<bf16>(tokC);
    scr_c = A.alloc<bf16>(tokC);
    scr_mid = A.alloc<bf16>(tokC * 4);
    scr_wide = A.alloc<bf16>(tokC * 8);
    scr_cat = A.alloc<bf16>(tokC * 2);
    gn_ws = A.alloc<float>(static_cast<size_t>(B) * 2 * mc * 4 * 2);
    d_emb_out = A.alloc<bf16>(static_cast<size_t>(B) * e->emb_cols);
    kv_all = A.alloc<bf16>(static_cast<size_t>(B) * L * e->kv_cols);
    d_kv_all = A.alloc<bf16>(static_cast<size_t>(B) * L * e->kv_cols);

    // ================= context encoder (fp32) + fused K/V projection =================
    float* emb = A.alloc<float>(static_cast<size_t>(B) * L * D);
    float* cq = A.alloc<float>(static_cast<size_t>(B) * L * D);
    float* ck = A.alloc<float>(static_cast<size_t>(B) * L * D);
    float* cv = A.alloc<float>(static_cast<size_t>(B) * L * D);
    bf16* ctx = A.alloc<bf16>(static_cast<size_t>(B) * L * D);
    bf16* d_ctx = A.alloc<bf16>(static_cast<size_t>(B) * L * D);
    bf16* emb_bf = A.alloc<bf16>(static_cast<size_t>(B) * L * D);
    bf16* d_qkv = A.alloc<bf16>(static_cast<size_t>(B) * L * 3 * D);
    bf16* d_emb = A.alloc<bf16>(static_cast<size_t>(B) * L * D);
    if (L > c.max_seq_len) { err = "context longer than max_seq_len"; return false; }
    if (L > 16) { err = "training supports a character context of at most 16 tokens"; return false; }
    if (!dry) {
      const float* E = W("word_emb.embedding.weight");
      const float* pe = e->pe;
      const int vocab = c.vocab_size, Bc = B;
      F.push_back(TOp{"ctx.embed", [=](const TRun& r, cudaStream_t s) { return embed_tokens_launch(r.ctx, 1, E, vocab, pe, 1, emb, Bc, L, D, s); }});
      const char* names[3] = {"word_emb.attention.linear_query", "word_emb.attention.linear_key", "word_emb.attention.linear_value"};
      float* outs[3] = {cq, ck, cv};
      for (int i = 0; i < 3; ++i) {
        const float* w = W(std::string(names[i]) + ".weight");
        const float* b = W(std::string(names[i]) + ".bias");
        float* o = outs[i];
        F.push_back(TOp{"ctx.linear", [=](const TRun&, cudaStream_t s) { return linear_f32_launch(emb, w, b, o, Bc * L, D, D, s); }});
      }
      F.push_back(TOp{"ctx.word_attn", [=](const TRun&, cudaStream_t s) { return word_attn_launch(cq, ck, cv, ctx, nullptr, Bc, L, D, L, 0, s); }});
    }
    if (!lin(F, "ctx.kv_proj", ctx, D, B * L, e->kv_all_w, e->kv_cols, D, kv_all, e->kv_cols)) return false;
    tape.push_back([=]() -> bool {
      auto& Bk = plan->bwd;
      const int inner = e->kv_cols / (2 * e->n_kv);
      std::vector<float*> dst;
      for (const std::string& n : e->kv_names) dst.push_back(G(n));
      if (inner != 320) { err = "training: attention inner dimension must be 320"; return false; }
      if (!wgrad(Bk, "ctx.kv_proj.wgrad", WX{ctx, D, D, false, 1, 1, 1, 1}, d_kv_all, e->kv_cols, e->kv_cols, B * L, 0, 0, dst, D, 1, 0))
        return false;
      if (!lin(Bk, "ctx.kv_proj.dgrad", d_kv_all, e->kv_cols, B * L, e->kv_all_wt, D, e->kv_cols, d_ctx, D)) return false;
      if (!dry) {
        const int Bc = B;
        Bk.push_back(TOp{"ctx.word_attn_bwd", [=](const TRun&, cudaStream_t s) { return word_attn_bwd_launch(cq, ck, cv, d_ctx, d_qkv, Bc, L, D, L, 0, s); }});
        Bk.push_back(TOp{"ctx.cast", [=](const TRun&, cudaStream_t s) { return f32_to_bf16_launch(emb, emb_bf, static_cast<size_t>(Bc) * L * D, s); }});
      }
      const char* names[3] = {"word_emb.attention.linear_query", "word_emb.attention.linear_key", "word_emb.attention.linear_value"};
      std::vector<float*> dq;
      for (int i = 0; i < 3; ++i) dq.push_back(G(std::string(names[i]) + ".weight"));
      if (D != 320) { err = "training: context_dim must be 320"; return false; }
      if (!wgrad(Bk, "ctx.qkv.wgrad", WX{emb_bf, D, D, false, 1, 1, 1, 1}, d_qkv, 3 * D, 3 * D, B * L, 0, 0, dq, D, 1, 0)) return false;
      for (int i = 0; i < 3; ++i) colsum(Bk, d_qkv + i * D, 3 * D, D, B * L, G(std::string(names[i]) + ".bias"));
      if (!lin(Bk, "ctx.qkv.dgrad", d_qkv, 3 * D, B * L, e->wqkv_t, D, 3 * D, d_emb, D)) return false;
      if (!dry) {
        float* dE = G("word_emb.embedding.weight");
        const int Bc = B, vocab = e->cfg.vocab_size;
        Bk.push_back(TOp{"ctx.embed_bwd", [=](const TRun& r, cudaStream_t s) { return scatter_add_rows_launch(d_emb, D, r.ctx, 1, dE, Bc * L, D, vocab, s); }});
      }
      return true;
    });

    // ================= timestep / label embedding =================
    bf16* temb = A.alloc<bf16>(static_cast<size_t>(B) * mc);
    bf16* h1p = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    bf16* h1 = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    bf16* embp = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    bf16* emb_act = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    float* emb_out = A.alloc<float>(static_cast<size_t>(B) * e->emb_cols);
    bf16* d_emb_act = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    bf16* d_embp = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    bf16* d_h1 = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    bf16* d_h1p = A.alloc<bf16>(static_cast<size_t>(B) * ted);
    const bool use_label = c.num_classes > 0 && c.add_label_emb;
    {
      if (!dry) {
        const int Bc = B;
        F.push_back(TOp{"temb", [=](const TRun& r, cudaStream_t s) { return timestep_embed_launch(r.t, 0, temb, Bc, mc, s); }});
      }
      if (!lin(F, "time_embed.0", temb, mc, B, e->te0.pw, ted, mc, h1p, ted, W(e->te0.b))) return false;
      if (!dry) {
        const size_t n = static_cast<size_t>(B) * ted;
        F.push_back(TOp{"silu", [=](const TRun&, cudaStream_t s) { return silu_fwd_launch(h1p, h1, n, s); }});
      }
      GEpi ep;
      ep.bias = W(e->te2.b);
      ep.out = embp;
      ep.out_ld = ted;
      if (use_label) {
        ep.rowbias = W("label_emb.weight");
        ep.rb_ld = ted;
        ep.rows_per_sample = 1;
        ep.patch_y = 1;
      }
      if (!gemm(F, "time_embed.2", B, false, 0, 0, {GSrc{h1, ted, ted, 1, 1, 1, 1}}, e->te2.pw, ted, ted, ep)) return false;
      if (!dry) {
        const size_t n = static_cast<size_t>(B) * ted;
        F.push_back(TOp{"silu", [=](const TRun&, cudaStream_t s) { return silu_fwd_launch(embp, emb_act, n, s); }});
      }
      GEpi ep3;
      ep3.bias = e->emb_all_b;
      ep3.out = emb_out;
      ep3.out_ld = e->emb_cols;
      ep3.out_f32 = 1;
      if (!gemm(F, "emb_layers", B, false, 0, 0, {GSrc{emb_act, ted, ted, 1, 1, 1, 1}}, e->emb_all_w, e->emb_cols, ted, ep3)) return false;
    }
    tape.push_back([=]() -> bool {
      auto& Bk = plan->bwd;
      std::vector<float*> dst;
      for (const TRes& r : e->res) {
        if (r.Cout != 320) { err = "training: ResBlock channels must be 320"; return false; }
        dst.push_back(G(r.emb_w));
      }
      if (!wgrad(Bk, "emb_layers.wgrad", WX{emb_act, ted, ted, false, 1, 1, 1, 1}, d_emb_out, e->emb_cols, e->emb_cols, B, 0, 0, dst, ted, 1, 0))
        return false;
      for (const TRes& r : e->res) colsum(Bk, d_emb_out + r.emb_off, e->emb_cols, r.Cout, B, G(r.emb_b), B);
      if (!lin(Bk, "emb_layers.dgrad", d_emb_out, e->emb_cols, B, e->emb_all_wt, ted, e->emb_cols, d_emb_act, ted)) return false;
      const size_t n = static_cast<size_t>(B) * ted;
      if (!dry) Bk.push_back(TOp{"silu_bwd", [=](const TRun&, cudaStream_t s) { return silu_bwd_launch(embp, d_emb_act, d_embp, n, s); }});
      if (use_label && !dry) {
        float* dL = G("label_emb.weight");
        const int Bc = B, ncls = e->cfg.num_classes;
        Bk.push_back(TOp{"label_emb_bwd", [=](const TRun& r, cudaStream_t s) { return scatter_add_rows_launch(d_embp, ted, r.y, 1, dL, Bc, ted, ncls, s); }});
      }
      if (!wgrad_lin(Bk, "time_embed.2.wgrad", h1, ted, ted, d_embp, ted, ted, B, G(e->te2.w))) return false;
      colsum(Bk, d_embp, ted, ted, B, G(e->te2.b), B);
      if (!lin(Bk, "time_embed.2.dgrad", d_embp, ted, B, e->te2.pwt, ted, ted, d_h1, ted)) return false;
      if (!dry) Bk.push_back(TOp{"silu_bwd", [=](const TRun&, cudaStream_t s) { return silu_bwd_launch(h1p, d_h1, d_h1p, n, s); }});
      if (!wgrad_lin(Bk, "time_embed.0.wgrad", temb, mc, mc, d_h1p, ted, ted, B, G(e->te0.w))) return false;
      colsum(Bk, d_h1p, ted, ted, B, G(e->te0.b), B);
      return true;
    });
    const int emb_ld = e->emb_cols;
    if (!dry) {
      plan->named["temb"] = {temb, static_cast<size_t>(B) * mc * 2};
      plan->named["h1p"] = {h1p, static_cast<size_t>(B) * ted * 2};
      plan->named["h1"] = {h1, static_cast<size_t>(B) * ted * 2};
      plan->named["embp"] = {embp, static_cast<size_t>(B) * ted * 2};
      plan->named["emb_act"] = {emb_act, static_cast<size_t>(B) * ted * 2};
      plan->named["emb_out"] = {emb_out, static_cast<size_t>(B) * e->emb_cols * 4};
      plan->named["ctx"] = {ctx, static_cast<size_t>(B) * L * D * 2};
      plan->named["kv_all"] = {kv_all, static_cast<size_t>(B) * L * e->kv_cols * 2};
      plan->named["d_emb_out"] = {d_emb_out, static_cast<size_t>(B) * e->emb_cols * 2};
    }

    // ================= UNet body =================
    std::vector<TT*> hs;
    TT* h = nullptr;
    auto run_block = [&](const TBlock& blk, std::vector<TT*> in) -> bool {
      for (const TLayer& l : blk) {
        TT* out = nullptr;
        switch (l.kind) {
          case TL_CONVIN: {
            out = new_t(H0, W0, mc);
            bf16* col = A.alloc<bf16>(static_cast<size_t>(M0) * 128);
            float* gtmp = A.alloc<float>(static_cast<size_t>(mc) * 128);
            if (!dry) {
              const int Bc = B;
              F.push_back(TOp{"conv_in.im2col", [=](const TRun& r, cudaStream_t s) { return conv_in_im2col_launch(r.x, col, Bc, H0, W0, s); }});
              plan->zero_on_bwd.push_back({gtmp, static_cast<size_t>(mc) * 128 * sizeof(float)});
            }
            GEpi ep;
            ep.bias = W("input_blocks.0.0.bias");
            ep.out = out->p;
            ep.out_ld = mc;
            ep.rows_per_sample = HW0;
            if (epilogue_stats_ok(HW0, mc)) { ep.gn_partial = out->stats; out->pslots = HW0 / 32; }
            if (!gemm(F, "conv_in", M0, false, 0, 0, {GSrc{col, 128, 128, 1, 1, 1, 1}}, e->conv_in_w, mc, 128, ep)) return false;
            if (!ensure_stats(F, out)) return false;
            TT* o = out;
            tape.push_back([=]() -> bool {
              auto& Bk = plan->bwd;
              if (mc != 320) { err = "training: model_channels must be 320"; return false; }
              if (!wgrad(Bk, "conv_in.wgrad", WX{col, 128, 128, false, 1, 1, 1, 1}, o->g, mc, mc, M0, 0, 0, {gtmp}, 128, 1, 0)) return false;
              if (!dry) {
                float* dW = G("input_blocks.0.0.weight");
                Bk.push_back(TOp{"conv_in.fold", [=](const TRun&, cudaStream_t s) { return conv_in_wgrad_fold_launch(gtmp, dW, mc, s); }});
              }
              colsum(Bk, o->g, mc, mc, M0, G("input_blocks.0.0.bias"));
              return true;
            });
            break;
          }
          case TL_RES:
            if (!res_block(e->res[l.idx], in, emb_out, emb_ld, out)) return false;
            break;
          case TL_ST:
            if (!st_block(e->st[l.idx], in[0], out)) return false;
            break;
          case TL_DOWN: {
            TT* x = in[0];
            const TSamp& sp = e->samp[l.idx];
            if (x->H % 2 || x->W % 2) { err = "downsample needs even spatial size"; return false; }
            out = new_t(x->H / 2, x->W / 2, x->C);
            const int Ho = out->H, Wo = out->W, Mo = B * Ho * Wo;
            GEpi ep;
            ep.bias = W(sp.conv.b);
            ep.out = out->p;
            ep.out_ld = x->C;
            ep.rows_per_sample = Ho * Wo;
            if (epilogue_stats_ok(Ho * Wo, x->C)) { ep.gn_partial = out->stats; out->pslots = Ho * Wo / 32; }
            if (!gemm(F, "down.conv", Mo, true, Ho, Wo, {GSrc{x->p, x->C, x->C, 9, 2, x->H, x->W}}, sp.conv.pw, x->C, 9 * x->C, ep)) return false;
            if (!ensure_stats(F, out)) return false;
            TT* o = out;
            tape.push_back([=]() -> bool {
              auto& Bk = plan->bwd;
              if (!wgrad(Bk, "down.wgrad", WX{x->p, x->C, x->C, true, 9, 2, x->H, x->W}, o->g, x->C, x->C, Mo, Ho, Wo, {G(sp.conv.w)},
                         static_cast<long long>(x->C) * 9, 9, 1))
                return false;
              colsum(Bk, o->g, x->C, x->C, Mo, G(sp.conv.b));
              // data gradient of the stride-2 conv = stride-1 transposed conv of the zero-dilated gradient
              bf16* dil = scr_mid;
              if (!dry) {
                const bf16* og = o->g;
                const int Bc = B, C = x->C;
                Bk.push_back(TOp{"down.dilate", [=](const TRun&, cudaStream_t s) { return dilate2x_launch(og, dil, Bc, Ho, Wo, C, s); }});
              }
              GEpi ep2;
              ep2.out = x->g;
              ep2.out_ld = x->C;
              if (x->g_init) { ep2.residual = x->g; ep2.res_ld = x->C; }
              x->g_init = true;
              return gemm(Bk, "down.dgrad", B * x->H * x->W, true, x->H, x->W, {GSrc{dil, x->C, x->C, 9, 1, x->H, x->W}}, sp.conv.pwt, x->C,
                          9 * x->C, ep2);
            });
            break;
          }
          case TL_UP: {
            TT* x = in[0];
            const TSamp& sp = e->samp[l.idx];
            TT* up = new_t(x->H * 2, x->W * 2, x->C);
            if (!dry) {
              const bf16* xp = x->p;
              bf16* upp = up->p;
              const int Bc = B, Hh = x->H, Ww = x->W, C = x->C;
              F.push_back(TOp{"upsample", [=](const TRun&, cudaStream_t s) { return upsample2x_launch(xp, upp, Bc, Hh, Ww, C, s); }});
            }
            out = new_t(up->H, up->W, x->C);
            const int Ho = up->H, Wo = up->W, Mo = B * Ho * Wo;
            GEpi ep;
            ep.bias = W(sp.conv.b);
            ep.out = out->p;
            ep.out_ld = x->C;
            ep.rows_per_sample = Ho * Wo;
            if (epilogue_stats_ok(Ho * Wo, x->C)) { ep.gn_partial = out->stats; out->pslots = Ho * Wo / 32; }
            if (!gemm(F, "up.conv", Mo, true, Ho, Wo, {GSrc{up->p, x->C, x->C, 9, 1, Ho, Wo}}, sp.conv.pw, x->C, 9 * x->C, ep)) return false;
            if (!ensure_stats(F, out)) return false;
            TT* o = out;
            tape.push_back([=]() -> bool {
              auto& Bk = plan->bwd;
              if (!wgrad(Bk, "up.wgrad", WX{up->p, x->C, x->C, true, 9, 1, Ho, Wo}, o->g, x->C, x->C, Mo, Ho, Wo, {G(sp.conv.w)},
                         static_cast<long long>(x->C) * 9, 9, 1))
                return false;
              colsum(Bk, o->g, x->C, x->C, Mo, G(sp.conv.b));
              GEpi ep2;
              ep2.out = up->g;
              ep2.out_ld = x->C;
              if (!gemm(Bk, "up.dgrad", Mo, true, Ho, Wo, {GSrc{o->g, x->C, x->C, 9, 1, Ho, Wo}}, sp.conv.pwt, x->C, 9 * x->C, ep2)) return false;
              if (!dry) {
                const bf16* ug = up->g;
                bf16* xg = x->g;
                const int Bc = B, Hh = x->H, Ww = x->W, C = x->C, acc = x->g_init ? 1 : 0;
                Bk.push_back(TOp{"up.sum2x2", [=](const TRun&, cudaStream_t s) { return upsample2x_bwd_launch(ug, xg, Bc, Hh, Ww, C, acc, s); }});
              }
              x->g_init = true;
              return true;
            });
            break;
          }
        }
        in = {out};
        h = out;
      }
      return true;
    };

    for (auto& blk : e->input_blocks) {
      if (!run_block(blk, h ? std::vector<TT*>{h} : std::vector<TT*>{})) return false;
      hs.push_back(h);
    }
    if (!run_block(e->middle, {h})) return false;
    for (auto& blk : e->output_blocks) {
      TT* skip = hs.back();
      hs.pop_back();
      if (skip->H != h->H || skip->W != h->W) { err = "skip connection spatial mismatch"; return false; }
      if (!run_block(blk, {h, skip})) return false;
    }
    // out: GN + SiLU + conv 320 -> 4 (fp32 NCHW prediction)
    {
      TT* a;
      GroupNormArgs fao;
      TT* hl = h;
      if (!gn_fwd(F, {hl}, e->out_gn, 1e-5f, 1, a, fao)) return false;
      GEpi ep;
      ep.epi = EPI_SAMPLER;
      ep.bias = W(e->conv_out.b);
      ep.rows_per_sample = a->H * a->W;
      if (!gemm(F, "conv_out", B * a->H * a->W, true, a->H, a->W, {GSrc{a->p, a->C, a->C, 9, 1, a->H, a->W}}, e->conv_out.pw,
                GEMM_BLOCK_N_OUT, 9 * a->C, ep))
        return false;
      bf16* d64 = A.alloc<bf16>(static_cast<size_t>(M0) * 64);
      float* btmp = A.alloc<float>(64);
      if (!dry) {
        T_ZERO_ONCE.push_back({d64, static_cast<size_t>(M0) * 64 * sizeof(bf16)});
        plan->zero_on_bwd.push_back({btmp, 64 * sizeof(float)});
      }
      tape.push_back([=]() -> bool {
        auto& Bk = plan->bwd;
        const int Ha = a->H, Wa = a->W, Ma = B * Ha * Wa;
        if (!dry) {
          const int Bc = B;
          Bk.push_back(TOp{"d_eps.layout", [=](const TRun& r, cudaStream_t s) { return nchw4_to_tok64_launch(r.d_eps, d64, Bc, Ha * Wa, s); }});
        }
        if (!wgrad(Bk, "conv_out.wgrad", WX{a->p, a->C, a->C, true, 9, 1, Ha, Wa}, d64, 64, 64, Ma, Ha, Wa, {G(e->conv_out.w)},
                   static_cast<long long>(a->C) * 9, 9, 1, 64, e->cfg.out_channels))
          return false;
        colsum(Bk, d64, 64, 64, Ma, btmp);
        if (!dry) {
          float* gb = G(e->conv_out.b);
          const int oc = e->cfg.out_channels;
          Bk.push_back(TOp{"conv_out.bias", [=](const TRun&, cudaStream_t s) { return repack_vec_launch(btmp, gb, oc, 0, 0, 1, s); }});
        }
        GEpi ep2;
        ep2.out = a->g;
        ep2.out_ld = a->C;
        if (!gemm(Bk, "conv_out.dgrad", Ma, true, Ha, Wa, {GSrc{d64, 64, 64, 9, 1, Ha, Wa}}, e->conv_out.pwt, a->C, 9 * 64, ep2)) return false;
        return gn_bwd(Bk, {hl}, e->out_gn, fao, a, nullptr, 0, nullptr, 0, gn_ws);
      });
    }
    // ================= backward plan: the tape in reverse =================
    g_touched.clear();
    for (auto it = tape.rbegin(); it != tape.rend(); ++it) {
      if (!(*it)()) return false;
      const int stage = static_cast<int>(plan->stage_end.size());
      for (const std::string& n : g_touched) plan->grad_stage[n] = stage;
      g_touched.clear();
      plan->stage_end.push_back(static_cast<int>(plan->bwd.size()));
    }
    plan->bytes = A.used;
    return true;
  }
  std::vector<std::pair<void*, size_t>> T_ZERO_ONCE;  // zeroed when the plan is created (columns the kernels never write)
};

int ensure_tplan(wd_trainer* e, int B, int L, TPlan** out) {
  if (B < 1 || L < 1) return tfail(WD_ERR_INVALID, "batch and context length must be >= 1");
  auto key = std::make_pair(B, L);
  auto it = e->plans.find(key);
  if (it != e->plans.end()) {
    *out = it->second.get();
    return WD_OK;
  }
  for (auto& kv : e->expected) {
    auto p = e->params.find(kv.first);
    if (p == e->params.end() || !p->second.w || !p->second.g) return tfail(WD_ERR_STATE, "parameter '%s' is not bound", kv.first.c_str());
  }
  if (!e->pe_set) return tfail(WD_ERR_STATE, "positional encoding not set");
  size_t need = 0;
  {
    TPlan tmp;
    tmp.B = B;
    tmp.L = L;
    TPlanBuilder pb{e, &tmp, TArena(), true, B};
    if (!pb.build()) return tfail(WD_ERR_UNSUPPORTED, "train plan: %s", pb.err.c_str());
    need = tmp.bytes;
  }
  if (need > e->acap) {
    T_CUDA_TRY(cudaDeviceSynchronize());
    if (e->abase) T_CUDA_TRY(cudaFree(e->abase));
    e->abase = nullptr;
    e->acap = 0;
    e->plans.clear();
    e->cur = nullptr;
    T_CUDA_TRY(cudaMalloc(&e->abase, need));
    e->acap = need;
  }
  std::unique_ptr<TPlan> p(new TPlan());
  p->B = B;
  p->L = L;
  TArena A;
  A.base = e->abase;
  TPlanBuilder pb{e, p.get(), A, false, B};
  if (!pb.build()) return tfail(WD_ERR_UNSUPPORTED, "train plan: %s", pb.err.c_str());
  for (auto& z : pb.T_ZERO_ONCE) T_CUDA_TRY(cudaMemset(z.first, 0, z.second));
  if (e->n_stages > 0) {  // the caller may already have laid its gradient buckets out by stage
    if (static_cast<int>(p->stage_end.size()) != e->n_stages) return tfail(WD_ERR_STATE, "backward stage count changed between plans");
    for (auto& kv : p->grad_stage) {
      auto it2 = e->grad_stage.find(kv.first);
      const int promised = it2 == e->grad_stage.end() ? e->n_stages - 1 : it2->second;
      if (kv.second > promised) return tfail(WD_ERR_STATE, "gradient of '%s' finishes in a later stage than announced", kv.first.c_str());
    }
  }
  *out = p.get();
  // plans of different batch sizes alias the same arena: only one is valid at a time
  e->plans.clear();
  e->plans[key] = std::move(p);
  return WD_OK;
}

// env WD_TRAIN_PROF=1 (tools/train_bench.py): CUDA events around every launch, per-kernel-name totals printed to stderr
bool train_prof_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_TRAIN_PROF");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}
int run_tops(const std::vector<TOp>& ops, const TRun& r, cudaStream_t s, const char* phase) {
  const bool prof = train_prof_enabled();
  std::vector<cudaEvent_t> ev;
  if (prof) {
    ev.resize(ops.size() + 1);
    for (auto& x : ev) cudaEventCreate(&x);
    cudaEventRecord(ev[0], s);
  }
  size_t i = 0;
  for (const TOp& op : ops) {
    const cudaError_t err = op.fn(r, s);
    if (err != cudaSuccess) return tfail(WD_ERR_CUDA, "launch of '%s' failed: %s", op.what, cudaGetErrorString(err));
    if (prof) cudaEventRecord(ev[i + 1], s);
    ++i;
  }
  if (prof) {
    cudaStreamSynchronize(s);
    std::map<std::string, std::pair<double, int>> agg;
    double total = 0;
    for (size_t k = 0; k < ops.size(); ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
      auto& a = agg[ops[k].what];
      a.first += ms;
      a.second += 1;
      total += ms;
    }
    fprintf(stderr, "[wd_train_prof] %s: %zu launches, %.3f ms\n", phase, ops.size(), total);
    std::vector<std::pair<double, std::string>> order;
    for (auto& kv : agg) order.push_back({kv.second.first, kv.first});
    std::sort(order.rbegin(), order.rend());
    for (auto& o : order)
      fprintf(stderr, "[wd_train_prof]   %-24s x%-3d %8.3f ms  %5.1f%%\n", o.second.c_str(), agg[o.second].second, o.first, 100.0 * o.first / total);
    for (auto& x : ev) cudaEventDestroy(x);
  }
  return WD_OK;
}

bool train_graph_enabled() {  // env WD_TRAIN_GRAPH (default on): replay the launch lists as CUDA graphs from the 2nd call on
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_TRAIN_GRAPH");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0 && !train_prof_enabled();
}
// first call of a plan: eager (kernel attributes get set, errors surface per launch); second call: captured; then replayed
// dry plan build -> which backward stage finishes each parameter's gradient (structure only: no device memory is touched)
int compute_grad_stages(wd_trainer* e) {
  if (e->n_stages > 0) return WD_OK;
  TPlan tmp;
  tmp.B = 2;
  tmp.L = std::max(1, std::min(e->cfg.max_seq_len, 16));
  TPlanBuilder pb{e, &tmp, TArena(), true, tmp.B};
  if (!pb.build()) return tfail(WD_ERR_UNSUPPORTED, "train plan: %s", pb.err.c_str());
  e->grad_stage = tmp.grad_stage;
  e->n_stages = static_cast<int>(tmp.stage_end.size());
  return WD_OK;
}

int run_list(const std::vector<TOp>& ops, const TRun& r, cudaStream_t s, const char* phase, cudaGraphExec_t* exec, int* calls,
             const std::vector<std::pair<void*, size_t>>* zero_first) {
  ++*calls;
  if (train_graph_enabled() && *exec) {
    T_CUDA_TRY(cudaGraphLaunch(*exec, s));
    return WD_OK;
  }
  const bool capture = train_graph_enabled() && *calls == 2;
  if (capture) {
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return run_list(ops, r, s, phase, exec, &(*calls = 2), zero_first);  // cannot capture on this stream: stay eager
    }
  }
  if (zero_first)
    for (auto& z : *zero_first) T_CUDA_TRY(cudaMemsetAsync(z.first, 0, z.second, s));
  const int rc = run_tops(ops, r, s, phase);
  if (capture) {
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (rc) return rc;
    if (ce != cudaSuccess || !g) return tfail(WD_ERR_CUDA, "graph capture of the %s list failed: %s", phase, cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) return tfail(WD_ERR_CUDA, "cudaGraphInstantiate (%s): %s", phase, cudaGetErrorString(ie));
    T_CUDA_TRY(cudaGraphLaunch(*exec, s));
  }
  return rc;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------
extern "C" int wd_trainer_create(const wd_config* cfg, wd_trainer** out) {
  if (!cfg || !out) return tfail(WD_ERR_INVALID, "null argument");
  if (cfg->variant != WD_VARIANT_UNET)
    return tfail(WD_ERR_UNSUPPORTED, "the training step is built for unet.UNetModel (train.py:403); the PHOSC variants are inference-only");
  if (cfg->in_channels != 4 || cfg->out_channels != 4) return tfail(WD_ERR_UNSUPPORTED, "in/out channels must be 4");
  if (cfg->model_channels != 320 || cfg->context_dim != 320)
    return tfail(WD_ERR_UNSUPPORTED, "training kernels are built for model_channels = context_dim = 320");
  for (int i = 0; i < cfg->n_channel_mult; ++i)
    if (cfg->channel_mult[i] != 1) return tfail(WD_ERR_UNSUPPORTED, "training kernels are built for channel_mult = 1");
  if (cfg->transformer_depth < 1 || cfg->n_channel_mult < 1 || cfg->n_channel_mult > 8) return tfail(WD_ERR_INVALID, "bad config");
  int dev = 0, major = 0;
  T_CUDA_TRY(cudaGetDevice(&dev));
  T_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return tfail(WD_ERR_UNSUPPORTED, "wd_b200 needs an sm_100a GPU (found compute capability %d.x)", major);
  std::unique_ptr<wd_trainer> e(new wd_trainer());
  e->cfg = *cfg;
  {
    TArena dryA;
    TBuilder b{e.get(), dryA, true};
    b.build();
    e->wbytes = dryA.used;
  }
  T_CUDA_TRY(cudaMalloc(&e->wbase, e->wbytes));
  T_CUDA_TRY(cudaMemset(e->wbase, 0, e->wbytes));
  TArena A;
  A.base = e->wbase;
  TBuilder b{e.get(), A, false};
  b.build();
  for (auto& s : e->st)
    if (s.dh != 80 || s.heads * s.dh != 320) return tfail(WD_ERR_UNSUPPORTED, "training attention kernels are built for 4 heads of 80 channels");
  *out = e.release();
  return WD_OK;
}

extern "C" void wd_trainer_destroy(wd_trainer* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  if (e->wbase) cudaFree(e->wbase);
  if (e->abase) cudaFree(e->abase);
  if (e->pack_tab_dev) cudaFree(e->pack_tab_dev);
  delete e;
}

extern "C" int wd_trainer_bind_param(wd_trainer* e, const char* name, const float* w, float* grad, const int64_t* shape, int ndim) {
  if (!e || !name || !w) return tfail(WD_ERR_INVALID, "null argument");
  auto it = e->expected.find(name);
  if (it == e->expected.end()) return WD_IGNORED;  // parameters the reference forward never reads receive no gradient
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  if (numel != it->second) return tfail(WD_ERR_INVALID, "parameter '%s': %lld elements, expected %lld", name, (long long)numel, (long long)it->second);
  if (!grad) return tfail(WD_ERR_INVALID, "parameter '%s' needs a gradient buffer", name);
  TParam& p = e->params[name];
  if (p.w != w || p.g != grad) {
    e->plans.clear();  // plans capture parameter pointers
    e->cur = nullptr;
    e->packs_valid = false;
  }
  p.w = w;
  p.g = grad;
  p.numel = numel;
  return WD_OK;
}

extern "C" int wd_trainer_set_pos_encoding(wd_trainer* e, const float* pe, void* stream) {
  if (!e || !pe) return tfail(WD_ERR_INVALID, "null argument");
  T_CUDA_TRY(cudaMemcpyAsync(e->pe, pe, static_cast<size_t>(e->cfg.max_seq_len) * e->cfg.context_dim * sizeof(float),
                             cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  e->pe_set = true;
  return WD_OK;
}

extern "C" int wd_trainer_sync_weights(wd_trainer* e, void* stream) {
  if (!e) return tfail(WD_ERR_INVALID, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // every matrix pack in one launch (device job table, re-uploaded only when a binding changed); the bias vectors keep their own
  // small launches because some of them accumulate into a slot another job wrote first
  static const bool multi = [] { const char* v = getenv("WD_TRAIN_MULTI_PACK"); return v ? atoi(v) != 0 : true; }();
  if (multi) {
    // table = [element-per-thread jobs (kinds 0, 2, 4)] ++ [tiled transposes (kinds 1, 3)]
    std::vector<PackDesc> tab, tabT;
    long long total = 0, total_tiles = 0;
    for (const PackJob& j : e->jobs) {
      if (j.kind == PK_VEC) continue;
      auto it = e->params.find(j.src);
      if (it == e->params.end() || !it->second.w) return tfail(WD_ERR_STATE, "parameter '%s' is not bound", j.src.c_str());
      PackDesc d{};
      d.src = it->second.w;
      d.dst = j.dst;
      d.N = j.N;
      d.K = j.K;
      d.ld = j.ld;
      d.off0 = j.off0;
      d.off1 = j.off1;
      if (j.kind == PK_LIN_T || j.kind == PK_CONV3_T) {
        d.kind = j.kind == PK_LIN_T ? 1 : 3;
        const long long C = j.kind == PK_LIN_T ? j.K : 9LL * j.K;
        d.start = total_tiles;
        total_tiles += ((j.N + 31) / 32) * ((C + 31) / 32);
        tabT.push_back(d);
        continue;
      }
      long long n = 0;
      switch (j.kind) {
        case PK_LIN: d.kind = 0; n = static_cast<long long>(j.N) * j.K; break;
        case PK_CONV3: d.kind = 2; n = 9LL * j.N * j.K; break;
        case PK_CONV_IN: d.kind = 4; n = 128LL * j.N; break;
        default: break;
      }
      d.start = total;
      total += n;
      tab.push_back(d);
    }
    const int n_plain = static_cast<int>(tab.size()), n_T = static_cast<int>(tabT.size());
    tab.insert(tab.end(), tabT.begin(), tabT.end());
    const size_t bytes = tab.size() * sizeof(PackDesc);
    if (bytes != e->pack_tab_host.size() || memcmp(tab.data(), e->pack_tab_host.data(), bytes) != 0) {
      if (e->pack_tab_dev) {
        cudaStreamSynchronize(s);  // a previous launch may still read the old table
        cudaFree(e->pack_tab_dev);
        e->pack_tab_dev = nullptr;
      }
      T_CUDA_TRY(cudaMalloc(&e->pack_tab_dev, bytes ? bytes : 16));
      T_CUDA_TRY(cudaMemcpy(e->pack_tab_dev, tab.data(), bytes, cudaMemcpyHostToDevice));
      e->pack_tab_host.assign(reinterpret_cast<const char*>(tab.data()), reinterpret_cast<const char*>(tab.data()) + bytes);
    }
    const PackDesc* dev = static_cast<const PackDesc*>(e->pack_tab_dev);
    T_CUDA_TRY(repack_multi_launch(dev, n_plain, total, s));
    T_CUDA_TRY(repack_multi_T_launch(dev + n_plain, n_T, total_tiles, s));
  }
  for (const PackJob& j : e->jobs) {
    if (multi && j.kind != PK_VEC) continue;
    auto it = e->params.find(j.src);
    if (it == e->params.end() || !it->second.w) return tfail(WD_ERR_STATE, "parameter '%s' is not bound", j.src.c_str());
    const float* src = it->second.w;
    switch (j.kind) {
      case PK_LIN:
        T_CUDA_TRY(repack_linear_launch(src, static_cast<bf16*>(j.dst), j.N, j.K, j.ld, j.off0, j.off1, 0, 0, s));
        break;
      case PK_LIN_T:
        T_CUDA_TRY(repack_linear_T_launch(src, static_cast<bf16*>(j.dst), j.N, j.K, j.ld, j.off0, j.off1, s));
        break;
      case PK_CONV3:
        T_CUDA_TRY(repack_conv3x3_launch(src, static_cast<bf16*>(j.dst), j.N, j.K, j.ld, j.off0, j.off1, s));
        break;
      case PK_CONV3_T:
        T_CUDA_TRY(repack_conv3x3_T_launch(src, static_cast<bf16*>(j.dst), j.N, j.K, j.ld, s));
        break;
      case PK_CONV_IN:
        T_CUDA_TRY(repack_conv_in_launch(src, static_cast<bf16*>(j.dst), j.N, j.K, s));
        break;
      case PK_VEC:
        T_CUDA_TRY(repack_vec_launch(src, static_cast<float*>(j.dst), j.N, j.off0, 0, j.off1, s));
        break;
    }
  }
  e->packs_valid = true;
  return WD_OK;
}

extern "C" int wd_trainer_forward(wd_trainer* e, int batch, const float* x, const int64_t* timesteps, const int64_t* y,
                                  const int64_t* ctx_tokens, int L, float* eps_out, void* stream) {
  if (!e || !x || !timesteps || !ctx_tokens || !eps_out) return tfail(WD_ERR_INVALID, "null argument");
  if (e->cfg.num_classes > 0 && e->cfg.add_label_emb && !y) return tfail(WD_ERR_INVALID, "y (writer ids) is required by this model");
  if (!e->packs_valid) return tfail(WD_ERR_STATE, "wd_trainer_sync_weights must run after binding / updating the parameters");
  TPlan* p = nullptr;
  int rc = ensure_tplan(e, batch, L, &p);
  if (rc) return rc;
  e->cur = p;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t HW = static_cast<size_t>(e->cfg.latent_h) * e->cfg.latent_w;
  T_CUDA_TRY(cudaMemcpyAsync(p->in_x, x, static_cast<size_t>(batch) * e->cfg.in_channels * HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  T_CUDA_TRY(cudaMemcpyAsync(p->in_t, timesteps, static_cast<size_t>(batch) * 8, cudaMemcpyDeviceToDevice, s));
  if (y) T_CUDA_TRY(cudaMemcpyAsync(p->in_y, y, static_cast<size_t>(batch) * 8, cudaMemcpyDeviceToDevice, s));
  T_CUDA_TRY(cudaMemcpyAsync(p->in_ctx, ctx_tokens, static_cast<size_t>(batch) * L * 8, cudaMemcpyDeviceToDevice, s));
  TRun r;
  r.x = p->in_x;
  r.t = p->in_t;
  r.y = p->in_y;
  r.ctx = p->in_ctx;
  r.eps_out = p->out_eps;
  rc = run_list(p->fwd, r, s, "forward", &p->g_fwd, &p->fwd_calls, nullptr);
  if (rc) return rc;
  T_CUDA_TRY(cudaMemcpyAsync(eps_out, p->out_eps, static_cast<size_t>(batch) * e->cfg.out_channels * HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  e->fwd_done = true;
  e->bwd_next_stage = 0;
  return WD_OK;
}

extern "C" int wd_trainer_backward(wd_trainer* e, const float* d_eps, const int64_t* y, const int64_t* ctx_tokens, void* stream) {
  if (!e || !d_eps) return tfail(WD_ERR_INVALID, "null argument");
  if (!e->cur || !e->fwd_done) return tfail(WD_ERR_STATE, "wd_trainer_forward must run before wd_trainer_backward");
  (void)y;           // the index tensors of the forward call are still staged in the plan
  (void)ctx_tokens;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TPlan* p = e->cur;
  const size_t HW = static_cast<size_t>(e->cfg.latent_h) * e->cfg.latent_w;
  T_CUDA_TRY(cudaMemcpyAsync(p->in_deps, d_eps, static_cast<size_t>(p->B) * e->cfg.out_channels * HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  TRun r;
  r.d_eps = p->in_deps;
  r.y = p->in_y;
  r.ctx = p->in_ctx;
  const int rc = run_list(p->bwd, r, s, "backward", &p->g_bwd, &p->bwd_calls, &p->zero_on_bwd);
  e->fwd_done = false;
  return rc;
}

// ---- staged backward: the same launch list cut at layer boundaries, so that the caller can start the all-reduce of the
// gradients a stage has finished while the later stages still run (bucketed, overlapped gradient exchange) ----
extern "C" int wd_trainer_num_grad_stages(wd_trainer* e, int* n) {
  if (!e || !n) return tfail(WD_ERR_INVALID, "null argument");
  const int rc = compute_grad_stages(e);
  if (rc) return rc;
  *n = e->n_stages;
  return WD_OK;
}
// *stage = the backward stage after which the gradient of parameter `name` is final (0-based, execution order)
extern "C" int wd_trainer_grad_stage(wd_trainer* e, const char* name, int* stage) {
  if (!e || !name || !stage) return tfail(WD_ERR_INVALID, "null argument");
  const int rc = compute_grad_stages(e);
  if (rc) return rc;
  if (!e->expected.count(name)) return tfail(WD_ERR_INVALID, "'%s' is not a parameter that receives a gradient", name);
  auto it = e->grad_stage.find(name);
  *stage = it == e->grad_stage.end() ? e->n_stages - 1 : it->second;  // not named by the dry build: final at the end
  return WD_OK;
}
// runs the backward stages [stage_begin, stage_end) of the last forward; a whole backward pass is the calls
// (0, a), (a, b), ..., (z, n) in order.  d_eps is read by the call that starts at stage 0.
extern "C" int wd_trainer_backward_stages(wd_trainer* e, const float* d_eps, int stage_begin, int stage_end, void* stream) {
  if (!e) return tfail(WD_ERR_INVALID, "null argument");
  if (!e->cur || !e->fwd_done) return tfail(WD_ERR_STATE, "wd_trainer_forward must run before wd_trainer_backward_stages");
  TPlan* p = e->cur;
  const int n = static_cast<int>(p->stage_end.size());
  if (stage_begin < 0 || stage_end <= stage_begin || stage_end > n) return tfail(WD_ERR_INVALID, "bad stage range [%d, %d) of %d", stage_begin, stage_end, n);
  if (stage_begin != e->bwd_next_stage) return tfail(WD_ERR_STATE, "backward stages must run in order: expected stage %d, got %d", e->bwd_next_stage, stage_begin);
  if (stage_begin == 0 && !d_eps) return tfail(WD_ERR_INVALID, "d_eps is required by the first stage");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (stage_begin == 0) {
    const size_t HW = static_cast<size_t>(e->cfg.latent_h) * e->cfg.latent_w;
    T_CUDA_TRY(cudaMemcpyAsync(p->in_deps, d_eps, static_cast<size_t>(p->B) * e->cfg.out_channels * HW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  TRun r;
  r.d_eps = p->in_deps;
  r.y = p->in_y;
  r.ctx = p->in_ctx;
  const int lo = stage_begin == 0 ? 0 : p->stage_end[stage_begin - 1], hi = p->stage_end[stage_end - 1];
  std::vector<TOp> ops(p->bwd.begin() + lo, p->bwd.begin() + hi);
  TPlan::Seg& seg = p->segs[std::make_pair(stage_begin, stage_end)];
  const int rc = run_list(ops, r, s, "backward stages", &seg.exec, &seg.calls, stage_begin == 0 ? &p->zero_on_bwd : nullptr);
  if (rc) {
    e->bwd_next_stage = 0;
    return rc;
  }
  e->bwd_next_stage = stage_end == n ? 0 : stage_end;
  if (stage_end == n) e->fwd_done = false;
  return WD_OK;
}

extern "C" int wd_trainer_launch_counts(const wd_trainer* e, int* fwd, int* bwd) {
  if (!e || !e->cur) return tfail(WD_ERR_STATE, "no plan");
  if (fwd) *fwd = static_cast<int>(e->cur->fwd.size());
  if (bwd) *bwd = static_cast<int>(e->cur->bwd.size());
  return WD_OK;
}
extern "C" int wd_trainer_read_tensor(const wd_trainer* e, const char* name, void* dst, size_t bytes, void* stream) {
  if (!e || !e->cur || !name || !dst) return tfail(WD_ERR_STATE, "no plan / null argument");
  auto it = e->cur->named.find(name);
  if (it == e->cur->named.end()) return tfail(WD_ERR_INVALID, "unknown tensor '%s'", name);
  if (bytes != it->second.second) return tfail(WD_ERR_INVALID, "tensor '%s' holds %zu bytes, not %zu", name, it->second.second, bytes);
  T_CUDA_TRY(cudaMemcpyAsync(dst, it->second.first, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}
extern "C" size_t wd_trainer_workspace_bytes(const wd_trainer* e) { return e ? e->acap : 0; }
extern "C" size_t wd_trainer_weight_bytes(const wd_trainer* e) { return e ? e->wbytes : 0; }

extern "C" int wd_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, int step, float ema_beta, int ema_mode,
                                 float grad_scale, void* stream) {
  if (!p || !g || !m || !v || (ema_mode && !ema)) return tfail(WD_ERR_INVALID, "null argument");
  T_CUDA_TRY(adamw_ema_launch(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, step, ema_beta, ema_mode, grad_scale,
                              static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

// ----------------------------------------------------------------------------------------------
// C ABI: single backward operators (parity tests; same kernels as the trainer)
// ----------------------------------------------------------------------------------------------
extern "C" int wd_op_wgrad_linear(const void* x, const void* dy, float* dw, int M, int N, int K, void* stream) {
  const int bn = (N % 320 == 0) ? 320 : 64;
  if (N % bn || K < WG_BLOCK_C || K % 64 || N / bn > WG_MAX_GROUPS) return tfail(WD_ERR_UNSUPPORTED, "wgrad_linear: N %% 320 (or N = 64), K >= 128, K %% 64");
  WgradLaunch L;
  memset(&L, 0, sizeof(L));
  L.bn = bn;
  WgradArgs& a = L.args;
  a.M = M;
  a.Cin = K;
  a.taps = 1;
  a.conv = 0;
  a.stride = 1;
  a.HWout = 1;
  a.Wout = 1;
  a.n_groups = N / bn;
  a.n_valid = bn;
  for (int g = 0; g < a.n_groups; ++g) a.dst[g] = dw + static_cast<size_t>(g) * bn * K;
  a.sN = K;
  a.sC = 1;
  a.sT = 0;
  a.splits = wgrad_pick_splits(M, a.n_groups * ((K + WG_BLOCK_C - 1) / WG_BLOCK_C));
  if (!tmap_encode_2d_bf16(&L.mapX, x, K, M, K, 64, WG_BLOCK_TOK)) return tfail(WD_ERR_CUDA, "tensor map X");
  if (!tmap_encode_2d_bf16(&L.mapDY, dy, N, M, N, 64, WG_BLOCK_TOK)) return tfail(WD_ERR_CUDA, "tensor map dY");
  T_CUDA_TRY(wgrad_tc_launch(L, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_wgrad_conv3x3(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout, int stride,
                                   void* stream) {
  const int bn = (Cout % 320 == 0) ? 320 : 64;
  if (Cout % bn || Cout / bn > WG_MAX_GROUPS || Cin < WG_BLOCK_C || Cin % 64) return tfail(WD_ERR_UNSUPPORTED, "wgrad_conv3x3: channel counts");
  if (stride != 1 && stride != 2) return tfail(WD_ERR_INVALID, "stride");
  const int Ho = H / stride, Wo = W / stride, HWo = Ho * Wo, M = B * HWo;
  WgradLaunch L;
  memset(&L, 0, sizeof(L));
  L.bn = bn;
  WgradArgs& a = L.args;
  a.M = M;
  a.Cin = Cin;
  a.taps = 9;
  a.conv = 1;
  a.stride = stride;
  a.HWout = HWo;
  a.Wout = Wo;
  a.n_groups = Cout / bn;
  a.n_valid = bn;
  for (int g = 0; g < a.n_groups; ++g) a.dst[g] = dw + static_cast<size_t>(g) * bn * Cin * 9;
  a.sN = static_cast<long long>(Cin) * 9;
  a.sC = 9;
  a.sT = 1;
  a.splits = wgrad_pick_splits(M, a.n_groups * ((Cin + WG_BLOCK_C - 1) / WG_BLOCK_C) * 9);
  uint32_t bw, bh, bnn;
  if (HWo >= WG_BLOCK_TOK) {
    if (HWo % WG_BLOCK_TOK || WG_BLOCK_TOK % Wo) return tfail(WD_ERR_UNSUPPORTED, "wgrad_conv3x3: spatial size");
    bw = Wo * stride;
    bh = (WG_BLOCK_TOK / Wo) * stride;
    bnn = 1;
  } else {
    if (WG_BLOCK_TOK % HWo) return tfail(WD_ERR_UNSUPPORTED, "wgrad_conv3x3: spatial size");
    bw = W;
    bh = H;
    bnn = WG_BLOCK_TOK / HWo;
  }
  if (!tmap_encode_4d_bf16(&L.mapX, x, Cin, W, H, B, Cin, 64, bw, bh, bnn, stride)) return tfail(WD_ERR_CUDA, "tensor map X");
  if (!tmap_encode_2d_bf16(&L.mapDY, dy, Cout, M, Cout, 64, WG_BLOCK_TOK)) return tfail(WD_ERR_CUDA, "tensor map dY");
  T_CUDA_TRY(wgrad_tc_launch(L, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_pack_conv3x3_t(const float* w_oihw, void* dst_bf16, int Cout, int Cin, void* stream) {
  T_CUDA_TRY(repack_conv3x3_T_launch(w_oihw, static_cast<bf16*>(dst_bf16), Cout, Cin, Cout, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}
extern "C" int wd_op_pack_linear_t(const float* w, void* dst_bf16, int N, int K, void* stream) {
  T_CUDA_TRY(repack_linear_T_launch(w, static_cast<bf16*>(dst_bf16), N, K, N, 0, 0, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_groupnorm_bwd(const void* x, const void* dy, const float* gamma, const float* beta, void* dx, float* dgamma,
                                   float* dbeta, int B, int HW, int C, int groups, float eps, int silu, void* stream) {
  if (C % groups || C % 8 || C > 1024) return tfail(WD_ERR_UNSUPPORTED, "groupnorm_bwd: unsupported channel count");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  const int slots = groupnorm_stats_slots(HW);
  float* partial = nullptr;
  float* ws = nullptr;
  T_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&partial), static_cast<size_t>(B) * groups * slots * 2 * sizeof(float), s));
  T_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&ws), static_cast<size_t>(B) * 4 * C * 2 * sizeof(float), s));
  GroupNormStatsArgs st{static_cast<const bf16*>(x), C, partial, HW, C, cpg, slots, 0};
  T_CUDA_TRY(groupnorm_stats_launch(st, B, s));
  GroupNormBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.x[0] = static_cast<const bf16*>(x);
  a.x_ld[0] = C;
  a.partial[0] = partial;
  a.pslots[0] = slots;
  a.dy = static_cast<const bf16*>(dy);
  a.dy_ld = C;
  a.gamma = gamma;
  a.beta = beta;
  a.ws = ws;
  a.dx[0] = static_cast<bf16*>(dx);
  a.dx_ld[0] = C;
  a.dgamma = dgamma;
  a.dbeta = dbeta;
  a.HW = HW;
  a.Cs = C;
  a.cpg = cpg;
  a.pcpg = cpg;
  a.eps = eps;
  a.silu = silu;
  T_CUDA_TRY(groupnorm_bwd_launch(a, B, 1, s));
  T_CUDA_TRY(cudaFreeAsync(partial, s));
  T_CUDA_TRY(cudaFreeAsync(ws, s));
  return WD_OK;
}

extern "C" int wd_op_layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* add, void* dx, float* dgamma,
                                   float* dbeta, int M, int C, float eps, void* stream) {
  T_CUDA_TRY(layernorm_bwd_launch(static_cast<const bf16*>(x), static_cast<const bf16*>(dy), gamma, static_cast<const bf16*>(add),
                                  static_cast<bf16*>(dx), dgamma, dbeta, M, C, eps, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_geglu_fwd(const void* p, void* out, int M, int H, void* stream) {
  T_CUDA_TRY(geglu_fwd_launch(static_cast<const bf16*>(p), static_cast<bf16*>(out), M, H, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}
extern "C" int wd_op_geglu_bwd(const void* p, const void* dout, void* dp, int M, int H, void* stream) {
  T_CUDA_TRY(geglu_bwd_launch(static_cast<const bf16*>(p), static_cast<const bf16*>(dout), static_cast<bf16*>(dp), M, H,
                              static_cast<cudaStream_t>(stream)));
  return WD_OK;
}

extern "C" int wd_op_attention_small_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv,
                                         int B, int Sq, int L, int heads, float scale, void* stream) {
  const int C = heads * 80;
  AttnSmallBwdArgs a;
  a.q = static_cast<const bf16*>(q); a.q_ld = C;
  a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v); a.kv_ld = C;
  a.dout = static_cast<const bf16*>(dout); a.do_ld = C;
  a.dq = static_cast<bf16*>(dq); a.dq_ld = C;
  a.dk = static_cast<bf16*>(dk); a.dv = static_cast<bf16*>(dv); a.dkv_ld = C;
  a.Sq = Sq; a.L = L; a.heads = heads; a.scale = scale;
  T_CUDA_TRY(attn_small_bwd_launch(a, B, static_cast<cudaStream_t>(stream)));
  return WD_OK;
}
