// fp32 mode of the hot path (BASELINE.json north_star: "per-step predicted noise within ... 1e-4 (fp32 mode)"; configs[1]
// "fp32 and bf16").  The reference computes everything in fp32 (use_fp16=False, unet.py:1193,1638): this file evaluates the same
// UNet with fp32 storage and fp32 FFMA arithmetic -- no tensor cores, no 16-bit tensor anywhere -- so that the result can be held
// against the reference at 1e-4.  It is the accuracy mode; the throughput mode is the tcgen05 engine (engine.cu).
//
// Layout: activations fp32 token-major [B*H*W, C] (NHWC), the latent in / eps out fp32 NCHW as at the reference seam.
// Weights: nn.Linear / 1x1 conv as stored ([N, K]); 3x3 conv repacked once to [Cout][tap][Cin] (k = tap*Cin + c) so that a
// 16-wide K block of the implicit GEMM is one contiguous channel run of one shifted pixel.
// The walker recovers the layer sequence from the loaded state_dict keys (unet.py:1248-1458 builds them in this order).
//
// Kernels: f32_gemm_kernel (every contraction: implicit-GEMM 3x3 conv with stride / nearest-2x / two-source concat / NCHW latent
// gather, Linear, 1x1 conv; fused bias, timestep-embedding row, residual, SiLU, NCHW store), f32_groupnorm_kernel (+SiLU),
// f32_layernorm_kernel, f32_attention_tq_kernel (80-wide heads, thread per query over shared-memory K/V tiles),
// f32_attention_kernel (warp per query: Word_Attention's single 320-wide head), f32_geglu_kernel, the embedding kernels, and for
// args.attentionMaps == 1 f32_attn_probs_kernel + f32_upsample_map_kernel.  Measured: DESIGN.md section 6 (fp32 mode).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wd_b200.h"
#include "engine_internal.h"
#include "f32_tc.h"

namespace {

// =====================================================================================================
// Implicit-GEMM / GEMM:  out[m, n] = act( sum_k A(m, k) W[n, k] + bias[n] + rowbias[sample(m), n] + residual[m, n] )
// A(m, k): k = tap*Cin + c, pixel = shifted (stride / nearest-2x aware) input pixel of output pixel m, channel c from the
// channel concatenation of up to two NHWC sources (torch.cat([h, hs.pop()], dim=1), unet.py:1750, is never materialised).
// 128 x 64 output tile, BK = 16, 256 threads, 8 x 4 accumulators per thread, register-staged prefetch of the next K block.
// =====================================================================================================
struct GemmF32 {
  const float* a1;
  const float* a2;
  int C1, C2;
  int taps;  // 1: plain GEMM (pixel = m), 9: 3x3 pad 1
  int Hin, Win, Hout, Wout;
  int stride;  // conv stride (Downsample: 2, unet.py:540)
  int up;      // nearest x2 before the conv (Upsample, unet.py:497)
  int a_nchw;  // a1 is the fp32 NCHW latent (conv_in)
  const float* w;
  const float* bias;
  const float* rowbias;
  int rb_ld;
  const float* residual;
  float* out;
  int out_nchw;
  int M, N, K;
  int act_silu;
};

constexpr int BM = 128, BN = 64, BK = 16, LDA_S = BM + 4, LDB_S = BN + 4;

struct RowCtx {  // per-thread decode of its A row
  int valid;
  int b, oh, ow;
  size_t pix;
};

__device__ __forceinline__ float4 load_a4(const GemmF32& g, const RowCtx& r, int k) {
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!r.valid || k >= g.K) return z;
  const int Cin = g.C1 + g.C2;
  int c = k;
  size_t pix = r.pix;
  int ih = 0, iw = 0;
  if (g.taps != 1) {
    const int tap = k / Cin;
    c = k - tap * Cin;
    const int kh = tap / 3, kw = tap - kh * 3;
    ih = r.oh * g.stride + kh - 1;
    iw = r.ow * g.stride + kw - 1;
    const int Hs = g.up ? 2 * g.Hin : g.Hin, Ws = g.up ? 2 * g.Win : g.Win;
    if (ih < 0 || ih >= Hs || iw < 0 || iw >= Ws) return z;
    if (g.up) {
      ih >>= 1;
      iw >>= 1;
    }
    pix = (static_cast<size_t>(r.b) * g.Hin + ih) * g.Win + iw;
  }
  if (g.a_nchw) {  // C1 == 4 channels of one pixel, plane stride Hin*Win
    const size_t plane = static_cast<size_t>(g.Hin) * g.Win;
    const float* p = g.a1 + (static_cast<size_t>(r.b) * g.C1 + c) * plane + static_cast<size_t>(ih) * g.Win + iw;
    return make_float4(p[0], p[plane], p[2 * plane], p[3 * plane]);
  }
  if (c < g.C1) return *reinterpret_cast<const float4*>(g.a1 + pix * g.C1 + c);
  return *reinterpret_cast<const float4*>(g.a2 + pix * g.C2 + (c - g.C1));
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

// ncu (profiles/r03a_ncu_f32_gemm.txt, 3x3 conv 320->320 at batch 256: 3.89 ms = 31 TFLOP/s): FMA pipe 50 % busy, issue slots 67 %,
// L1/shared 77 %, top stall short_scoreboard (shared-memory operand loads).  A 128-thread 8 x 8-accumulator variant (half the
// shared-memory wavefronts per FFMA, 128 registers, 4 CTAs/SM) measured no faster (89.8 vs 85.8 ms per batch-256 step) and was dropped;
// the division-free K loop below (FAST) brought 85.8 -> 82.9 ms.
// FAST (C1 % 16 == 0, C2 % 16 == 0, NHWC sources): a 16-wide K block lies inside one tap of one source, so the shifted-pixel
// pointers are recomputed only when the tap changes and the K loop carries no division.  (In the generic form the per-block index
// arithmetic was a third of all issued instructions: FFMA 64.5 % of 2.93 G warp instructions, ncu above.)
template <bool FAST>
__global__ void __launch_bounds__(256, 3) f32_gemm_kernel(const GemmF32 g) {
  __shared__ __align__(16) float As[BK][LDA_S];
  __shared__ __align__(16) float Bs[BK][LDB_S];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // A staging: row am, float4 columns aq and aq + 2 of the 16-wide K block
  const int am = tid & (BM - 1), aq = tid >> 7;
  RowCtx r;
  {
    const int m = m0 + am;
    r.valid = m < g.M;
    r.b = r.oh = r.ow = 0;
    r.pix = static_cast<size_t>(m);
    if (r.valid && g.taps != 1) {
      const int hw = g.Hout * g.Wout;
      r.b = m / hw;
      const int rem = m - r.b * hw;
      r.oh = rem / g.Wout;
      r.ow = rem - r.oh * g.Wout;
    }
  }
  // B staging: weight row bn, float4 column bq
  const int bn = tid & (BN - 1), bq = tid >> 6;
  const bool b_ok = (n0 + bn) < g.N;
  const float* wrow = g.w + static_cast<size_t>(b_ok ? n0 + bn : 0) * g.K;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // FAST-path cursor of the K block being fetched: tap, channel offset inside the tap, the pixel's row in each source
  const int Cin = g.C1 + g.C2;
  int cur_tap = 0, cur_c0 = 0;
  bool pv = false;
  const float* p1 = g.a1;
  const float* p2 = g.a2;
  auto set_tap = [&](int tp) {
    pv = r.valid != 0;
    size_t pix = r.pix;
    if (pv && g.taps != 1) {
      const int kh = tp / 3, kw = tp - kh * 3;
      int ih = r.oh * g.stride + kh - 1, iw = r.ow * g.stride + kw - 1;
      const int Hs = g.up ? 2 * g.Hin : g.Hin, Ws = g.up ? 2 * g.Win : g.Win;
      pv = ih >= 0 && ih < Hs && iw >= 0 && iw < Ws;
      if (pv) {
        if (g.up) {
          ih >>= 1;
          iw >>= 1;
        }
        pix = (static_cast<size_t>(r.b) * g.Hin + ih) * g.Win + iw;
      }
    }
    if (pv) {
      p1 = g.a1 + pix * g.C1;
      p2 = g.a2 + pix * g.C2;  // only dereferenced when C2 > 0
    }
  };
  auto fetch = [&](float4& x0, float4& x1) {
    x0 = make_float4(0.f, 0.f, 0.f, 0.f);
    x1 = x0;
    if (pv) {
      const float* base = cur_c0 < g.C1 ? p1 + cur_c0 : p2 + (cur_c0 - g.C1);
      x0 = *reinterpret_cast<const float4*>(base + aq * 4);
      x1 = *reinterpret_cast<const float4*>(base + aq * 4 + 8);
    }
  };

  float4 ra0, ra1;
  if (FAST) {
    set_tap(0);
    fetch(ra0, ra1);
  } else {
    ra0 = load_a4(g, r, aq * 4);
    ra1 = load_a4(g, r, (aq + 2) * 4);
  }
  float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b_ok && bq * 4 < g.K) rb = *reinterpret_cast<const float4*>(wrow + bq * 4);

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    As[aq * 4 + 0][am] = ra0.x;
    As[aq * 4 + 1][am] = ra0.y;
    As[aq * 4 + 2][am] = ra0.z;
    As[aq * 4 + 3][am] = ra0.w;
    As[aq * 4 + 8][am] = ra1.x;
    As[aq * 4 + 9][am] = ra1.y;
    As[aq * 4 + 10][am] = ra1.z;
    As[aq * 4 + 11][am] = ra1.w;
    Bs[bq * 4 + 0][bn] = rb.x;
    Bs[bq * 4 + 1][bn] = rb.y;
    Bs[bq * 4 + 2][bn] = rb.z;
    Bs[bq * 4 + 3][bn] = rb.w;
    __syncthreads();
    const int kn = k0 + BK;
    if (kn < g.K) {
      if (FAST) {
        cur_c0 += BK;
        if (cur_c0 >= Cin) {
          cur_c0 = 0;
          set_tap(++cur_tap);
        }
        fetch(ra0, ra1);
      } else {
        ra0 = load_a4(g, r, kn + aq * 4);
        ra1 = load_a4(g, r, kn + (aq + 2) * 4);
      }
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b_ok && kn + bq * 4 < g.K) rb = *reinterpret_cast<const float4*>(wrow + kn + bq * 4);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a_lo = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int hw_out = g.Hout * g.Wout;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
    const int sample = (g.rowbias || g.out_nchw) ? m / hw_out : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.rowbias) v += g.rowbias[static_cast<size_t>(sample) * g.rb_ld + n];
      if (g.residual) v += g.residual[static_cast<size_t>(m) * g.N + n];
      if (g.act_silu) v = silu_f(v);
      if (g.out_nchw)
        g.out[(static_cast<size_t>(sample) * g.N + n) * hw_out + (m - sample * hw_out)] = v;
      else
        g.out[static_cast<size_t>(m) * g.N + n] = v;
    }
  }
}

void launch_gemm(const GemmF32& g, cudaStream_t s) {
  const dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
  if (!g.a_nchw && (g.C1 % BK) == 0 && (g.C2 % BK) == 0)
    f32_gemm_kernel<true><<<grid, 256, 0, s>>>(g);
  else
    f32_gemm_kernel<false><<<grid, 256, 0, s>>>(g);
}

// =====================================================================================================
// GroupNorm(32 groups) (+SiLU), unet.py:429-431 / 161-162, over the channel concatenation of up to two NHWC sources.
// One CTA per (group, sample); mean, then the centred second moment (two passes), then the apply pass.
// =====================================================================================================
__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

__global__ void __launch_bounds__(256) f32_groupnorm_kernel(const float* __restrict__ a1, const float* __restrict__ a2, int C1,
                                                            int C2, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ out, int HW,
                                                            int groups, float eps, int silu) {
  __shared__ float red[32];
  const int C = C1 + C2, cg = C / groups;
  const int grp = blockIdx.x, b = blockIdx.y;
  const int c0 = grp * cg;
  const int n = HW * cg;
  auto at = [&](int i) -> float {
    const int p = i / cg, c = c0 + (i - p * cg);
    const size_t pix = static_cast<size_t>(b) * HW + p;
    return c < C1 ? a1[pix * C1 + c] : a2[pix * C2 + (c - C1)];
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += at(i);
  const float mean = block_sum(s, red) / static_cast<float>(n);
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = at(i) - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum(q, red) / static_cast<float>(n);
  const float rstd = 1.0f / sqrtf(var + eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / cg, c = c0 + (i - p * cg);
    float v = (at(i) - mean) * rstd * gamma[c] + beta[c];
    if (silu) v = silu_f(v);
    out[(static_cast<size_t>(b) * HW + p) * C + c] = v;
  }
}

// ---- coalesced two-kernel GroupNorm (the kernel above reads 40-80 byte channel runs three times: 1 TB/s) ----
// Statistics: a CTA takes `rpc` pixel rows of one sample and reads them as whole rows (float4 per thread, fixed channels per
// thread), sums x - shift and (x - shift)^2 per channel (shift = the sample's first value of the channel's group: the one-pass
// variance then has nothing to cancel), folds them per group in a fixed order and writes one partial per (sample, group, chunk).
// Apply: every CTA re-forms mean / rstd from the sample's partials (fixed order), then streams its rows once.
constexpr int GN32_T = 256;
__device__ __forceinline__ float gn32_shift(const float* a1, const float* a2, int C1, int C2, size_t pix0, int ch) {
  return ch < C1 ? __ldg(a1 + pix0 * C1 + ch) : __ldg(a2 + pix0 * C2 + (ch - C1));
}
__global__ void __launch_bounds__(GN32_T) f32_gn_stats_kernel(const float* __restrict__ a1, const float* __restrict__ a2, int C1, int C2,
                                                              int HW, int groups, int rpc, int nch, float2* __restrict__ partial) {
  __shared__ float ss[1024], sq[1024];
  const int C = C1 + C2, cg = C / groups;
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int r0 = chunk * rpc, r1 = min(r0 + rpc, HW);
  const size_t pix0 = static_cast<size_t>(b) * HW;
  for (int src = 0; src < 2; ++src) {
    const int Cs = src ? C2 : C1, coff = src ? C1 : 0;
    if (Cs == 0) continue;
    const float* ap = src ? a2 : a1;
    const int nv = Cs >> 2, R = GN32_T / nv;
    const int v = threadIdx.x % nv, rl = threadIdx.x / nv;
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = s4;
    if (rl < R) {
      float sh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) sh[j] = gn32_shift(a1, a2, C1, C2, pix0, ((coff + 4 * v + j) / cg) * cg);
      for (int p = r0 + rl; p < r1; p += R) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(ap + (pix0 + p) * Cs) + v);
        const float d0 = x.x - sh[0], d1 = x.y - sh[1], d2 = x.z - sh[2], d3 = x.w - sh[3];
        s4.x += d0; s4.y += d1; s4.z += d2; s4.w += d3;
        q4.x = fmaf(d0, d0, q4.x); q4.y = fmaf(d1, d1, q4.y); q4.z = fmaf(d2, d2, q4.z); q4.w = fmaf(d3, d3, q4.w);
      }
      float* ps = ss + rl * Cs + 4 * v;
      float* pq = sq + rl * Cs + 4 * v;
      ps[0] = s4.x; ps[1] = s4.y; ps[2] = s4.z; ps[3] = s4.w;
      pq[0] = q4.x; pq[1] = q4.y; pq[2] = q4.z; pq[3] = q4.w;
    }
    __syncthreads();
    const int gl = threadIdx.x;
    if (gl < Cs / cg) {
      float S = 0.f, Q = 0.f;
      for (int r = 0; r < R; ++r)
        for (int c = 0; c < cg; ++c) {
          S += ss[r * Cs + gl * cg + c];
          Q += sq[r * Cs + gl * cg + c];
        }
      partial[(static_cast<size_t>(b) * groups + coff / cg + gl) * nch + chunk] = make_float2(S, Q);
    }
    __syncthreads();
  }
}
// SPLIT: write the result as its TF32 split (hi = tf32_rn(y), lo = y - hi: the arithmetic of f32tc_split_kernel) into out / out_lo
__device__ __forceinline__ float gn32_tf32_rn(float a) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  return __uint_as_float(r & 0xFFFFE000u);
}
template <bool SPLIT>
__global__ void __launch_bounds__(GN32_T) f32_gn_apply_kernel(const float* __restrict__ a1, const float* __restrict__ a2, int C1, int C2,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float* __restrict__ out, float* __restrict__ out_lo, int HW, int groups,
                                                              int rpc, int nch, const float2* __restrict__ partial, float eps, int silu) {
  __shared__ float s_mean[128], s_rstd[128];
  const int C = C1 + C2, cg = C / groups;
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int r0 = chunk * rpc, r1 = min(r0 + rpc, HW);
  const size_t pix0 = static_cast<size_t>(b) * HW;
  if (static_cast<int>(threadIdx.x) < groups) {
    const int g = threadIdx.x;
    const float2* pg = partial + (static_cast<size_t>(b) * groups + g) * nch;
    float S = 0.f, Q = 0.f;
    for (int i = 0; i < nch; ++i) {
      const float2 t = __ldg(pg + i);
      S += t.x;
      Q += t.y;
    }
    const float inv_n = 1.0f / (static_cast<float>(HW) * static_cast<float>(cg));
    const float ms = S * inv_n;  // mean - shift
    const float var = fmaxf(fmaf(-ms, ms, Q * inv_n), 0.f);
    s_mean[g] = gn32_shift(a1, a2, C1, C2, pix0, g * cg) + ms;
    s_rstd[g] = 1.0f / sqrtf(var + eps);
  }
  __syncthreads();
  for (int src = 0; src < 2; ++src) {
    const int Cs = src ? C2 : C1, coff = src ? C1 : 0;
    if (Cs == 0) continue;
    const float* ap = src ? a2 : a1;
    const int nv = Cs >> 2, R = GN32_T / nv;
    const int v = threadIdx.x % nv, rl = threadIdx.x / nv;
    if (rl >= R) continue;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = coff + 4 * v + j, g = c / cg;
      sc[j] = s_rstd[g] * __ldg(gamma + c);
      sh[j] = fmaf(-s_mean[g], sc[j], __ldg(beta + c));
    }
    for (int p = r0 + rl; p < r1; p += R) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(ap + (pix0 + p) * Cs) + v);
      float4 y = make_float4(fmaf(x.x, sc[0], sh[0]), fmaf(x.y, sc[1], sh[1]), fmaf(x.z, sc[2], sh[2]), fmaf(x.w, sc[3], sh[3]));
      if (silu) {
        y.x = silu_f(y.x); y.y = silu_f(y.y); y.z = silu_f(y.z); y.w = silu_f(y.w);
      }
      if constexpr (SPLIT) {
        const float4 h = make_float4(gn32_tf32_rn(y.x), gn32_tf32_rn(y.y), gn32_tf32_rn(y.z), gn32_tf32_rn(y.w));
        reinterpret_cast<float4*>(out + (pix0 + p) * C + coff)[v] = h;
        reinterpret_cast<float4*>(out_lo + (pix0 + p) * C + coff)[v] = make_float4(y.x - h.x, y.y - h.y, y.z - h.z, y.w - h.w);
      } else {
        reinterpret_cast<float4*>(out + (pix0 + p) * C + coff)[v] = y;
      }
    }
  }
}

// nn.LayerNorm(C), eps 1e-5 (unet.py:314-316): one warp per token
// SPLIT: the result leaves as its TF32 split (out = hi, out_lo = lo) for the tensor-core Linear(s) that consume it
template <bool SPLIT>
__global__ void __launch_bounds__(256) f32_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ out,
                                                            float* __restrict__ out_lo, int M, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + static_cast<size_t>(row) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    q = fmaf(d, d, q);
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q / static_cast<float>(C) + eps);
  float* orow = out + static_cast<size_t>(row) * C;
  for (int c = lane; c < C; c += 32) {
    const float y = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if constexpr (SPLIT) {
      const float h = gn32_tf32_rn(y);
      orow[c] = h;
      out_lo[static_cast<size_t>(row) * C + c] = y - h;
    } else {
      orow[c] = y;
    }
  }
}

// =====================================================================================================
// softmax(q k^T scale) v (CrossAttention.forward unet.py:185-207; Word_Attention unet.py:825-836 with heads = 1, scale = 1).
// One warp per (query, head, sample); lane l owns channels l, l + 32, ...; keys in groups of four; online softmax in fp32.
// =====================================================================================================
template <int NPL>
__global__ void __launch_bounds__(256) f32_attention_kernel(const float* __restrict__ q, size_t q_bs, int ldq,
                                                            const float* __restrict__ k, const float* __restrict__ v,
                                                            size_t kv_bs, int ldkv, float* __restrict__ out, size_t o_bs, int ldo,
                                                            int B, int Sq, int Skv, int heads, int d, float scale) {
  const size_t wid = static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const size_t total = static_cast<size_t>(B) * heads * Sq;
  if (wid >= total) return;
  const int qi = static_cast<int>(wid % Sq);
  const int h = static_cast<int>((wid / Sq) % heads);
  const int b = static_cast<int>(wid / (static_cast<size_t>(Sq) * heads));
  const float* qp = q + b * q_bs + static_cast<size_t>(qi) * ldq + h * d;
  const float* kp = k + b * kv_bs + h * d;
  const float* vp = v + b * kv_bs + h * d;
  float qr[NPL], o[NPL];
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int c = lane + 32 * j;
    qr[j] = c < d ? qp[c] : 0.f;
    o[j] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < Skv; j0 += 4) {
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float part = 0.f;
      if (j0 + u < Skv) {
        const float* kr = kp + static_cast<size_t>(j0 + u) * ldkv;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int c = lane + 32 * j;
          if (c < d) part = fmaf(qr[j], kr[c], part);
        }
      }
      s[u] = part;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], off);
    float mnew = mx;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = (j0 + u < Skv) ? s[u] * scale : -INFINITY;
      mnew = fmaxf(mnew, s[u]);
    }
    const float corr = expf(mx - mnew);  // mx = -inf on the first group: exp(-inf) = 0
    l *= corr;
#pragma unroll
    for (int j = 0; j < NPL; ++j) o[j] *= corr;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 + u >= Skv) continue;
      const float p = expf(s[u] - mnew);
      l += p;
      const float* vr = vp + static_cast<size_t>(j0 + u) * ldkv;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        const int c = lane + 32 * j;
        if (c < d) o[j] = fmaf(p, vr[c], o[j]);
      }
    }
    mx = mnew;
  }
  float* op = out + b * o_bs + static_cast<size_t>(qi) * ldo + h * d;
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int c = lane + 32 * j;
    if (c < d) op[c] = o[j] / l;
  }
}

// Head width 80 (the UNet's 4 x 80 heads): one THREAD per query with q and the output row in registers, K/V tiles of 32 keys staged in
// shared memory once per 128 queries and read as warp-wide broadcasts -- no shuffles, and K/V leave L2 once per CTA instead of once
// per query (the warp-per-query kernel above re-read 640 B per (query, key) pair: ~0.5 TB of L1/L2 traffic per unetPhosc step).
// Keys in groups of eight: one running-max update and one rescale of the output row per group.
template <int D, int KT>
__global__ void __launch_bounds__(128) f32_attention_tq_kernel(const float* __restrict__ q, size_t q_bs, int ldq,
                                                               const float* __restrict__ k, const float* __restrict__ v, size_t kv_bs,
                                                               int ldkv, float* __restrict__ out, size_t o_bs, int ldo, int Sq, int Skv,
                                                               float scale) {
  constexpr int D4 = D / 4;
  __shared__ __align__(16) float Ks[KT][D];
  __shared__ __align__(16) float Vs[KT][D];
  const int tid = threadIdx.x;
  const int qi = blockIdx.x * 128 + tid, h = blockIdx.y, b = blockIdx.z;
  const bool active = qi < Sq;
  float4 qr[D4], o[D4];
  {
    const float4* qp = reinterpret_cast<const float4*>(q + b * q_bs + static_cast<size_t>(active ? qi : 0) * ldq + h * D);
#pragma unroll
    for (int c = 0; c < D4; ++c) {
      qr[c] = active ? qp[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float* kb = k + b * kv_bs + h * D;
  const float* vb = v + b * kv_bs + h * D;
  float mx = -INFINITY, l = 0.f;
  for (int t0 = 0; t0 < Skv; t0 += KT) {
    __syncthreads();
    for (int idx = tid; idx < KT * D4; idx += 128) {
      const int j = idx / D4, c = idx - j * D4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (t0 + j < Skv) {
        kk = *reinterpret_cast<const float4*>(kb + static_cast<size_t>(t0 + j) * ldkv + c * 4);
        vv = *reinterpret_cast<const float4*>(vb + static_cast<size_t>(t0 + j) * ldkv + c * 4);
      }
      *reinterpret_cast<float4*>(&Ks[j][c * 4]) = kk;
      *reinterpret_cast<float4*>(&Vs[j][c * 4]) = vv;
    }
    __syncthreads();
    const int nk = min(KT, Skv - t0);
    for (int j0 = 0; j0 < nk; j0 += 8) {
      float sc[8];
      float mnew = mx;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (j0 + u < nk) {
#pragma unroll
          for (int c = 0; c < D4; ++c) {
            const float4 kk = *reinterpret_cast<const float4*>(&Ks[j0 + u][c * 4]);
            a0 = fmaf(qr[c].x, kk.x, a0);
            a1 = fmaf(qr[c].y, kk.y, a1);
            a2 = fmaf(qr[c].z, kk.z, a2);
            a3 = fmaf(qr[c].w, kk.w, a3);
          }
          sc[u] = ((a0 + a1) + (a2 + a3)) * scale;
          mnew = fmaxf(mnew, sc[u]);
        } else {
          sc[u] = -INFINITY;
        }
      }
      const float corr = expf(mx - mnew);  // first group: exp(-inf) = 0
      l *= corr;
#pragma unroll
      for (int c = 0; c < D4; ++c) {
        o[c].x *= corr;
        o[c].y *= corr;
        o[c].z *= corr;
        o[c].w *= corr;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j0 + u < nk) {
          const float p = expf(sc[u] - mnew);
          l += p;
#pragma unroll
          for (int c = 0; c < D4; ++c) {
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[j0 + u][c * 4]);
            o[c].x = fmaf(p, vv.x, o[c].x);
            o[c].y = fmaf(p, vv.y, o[c].y);
            o[c].z = fmaf(p, vv.z, o[c].z);
            o[c].w = fmaf(p, vv.w, o[c].w);
          }
        }
      }
      mx = mnew;
    }
  }
  if (active) {
    float4* op = reinterpret_cast<float4*>(out + b * o_bs + static_cast<size_t>(qi) * ldo + h * D);
#pragma unroll
    for (int c = 0; c < D4; ++c) op[c] = make_float4(o[c].x / l, o[c].y / l, o[c].z / l, o[c].w / l);
  }
}

// ---- small elementwise kernels -------------------------------------------------------------------------
// timestep_embedding (unet.py:96-116): [B, dim] = cat(cos(t f), sin(t f)), f_i = exp(-ln(1e4) i / half)
__global__ void f32_timestep_kernel(const long long* __restrict__ t_dev, long long t_scalar, float* __restrict__ out, int B,
                                    int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx - b * half;
  const float t = static_cast<float>(t_dev ? t_dev[b] : t_scalar);
  const float freq = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
  const float arg = t * freq;
  out[static_cast<size_t>(b) * dim + i] = cosf(arg);
  out[static_cast<size_t>(b) * dim + half + i] = sinf(arg);
}

// emb = time_embed(...) (+ label_emb[y], unet.py:1578-1581); semb = SiLU(emb) (the first op of every ResBlock.emb_layers)
__global__ void f32_emb_finish_kernel(float* __restrict__ emb, const float* __restrict__ label_w, const long long* __restrict__ y,
                                      float* __restrict__ semb, int B, int D, int num_classes, int* __restrict__ bad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, c = idx - b * D;
  float v = emb[idx];
  if (label_w) {
    const long long cls = y[b];
    if (cls < 0 || cls >= num_classes) {
      *bad = 1;
    } else {
      v += label_w[static_cast<size_t>(cls) * D + c];
    }
  }
  emb[idx] = v;
  semb[idx] = silu_f(v);
}

// GEGLU (unet.py:122-130): out[m, j] = p[m, j] * gelu(p[m, H + j]), exact erf GELU (F.gelu default)
__global__ void f32_geglu_kernel(const float* __restrict__ p, float* __restrict__ out, size_t M, int H) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * H) return;
  const size_t m = idx / H;
  const int j = static_cast<int>(idx - m * H);
  const float a = p[m * 2 * H + j], gt = p[m * 2 * H + H + j];
  out[idx] = a * (0.5f * gt * (1.0f + erff(gt * 0.70710678118654752440f)));
}

// CharacterEncoder embedding (+ positional encoding), unet.py:851-872: tokens int64 or int32
__global__ void f32_embed_tokens_kernel(const long long* __restrict__ tok64, const int* __restrict__ tok32,
                                        const float* __restrict__ table, const float* __restrict__ pe, float* __restrict__ out,
                                        int B, int L, int D, int vocab, int* __restrict__ bad) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(B) * L * D) return;
  const int c = static_cast<int>(idx % D);
  const size_t row = idx / D;
  const int pos = static_cast<int>(row % L);
  const long long t = tok64 ? tok64[row] : static_cast<long long>(tok32[row]);
  if (t < 0 || t >= vocab) {
    *bad = 1;
    out[idx] = 0.f;
    return;
  }
  float v = table[static_cast<size_t>(t) * D + c];
  if (pe) v += pe[static_cast<size_t>(pos) * D + c];
  out[idx] = v;
}

// Attention probabilities summed over the heads (args.attentionMaps == 1: CrossAttention returns attn [B, heads, Sq, L],
// unet.py:276-279, and UNetModel.forward sums it over the heads, unet.py:1786): out[b, i, j] = sum_h softmax_j(q_h[i] . k_h[j] scale).
// One thread per (sample, query); L <= 32 (the character context).
__global__ void __launch_bounds__(128) f32_attn_probs_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                             float* __restrict__ out, int B, int Sq, int L, int heads, int d,
                                                             float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Sq) return;
  const int b = idx / Sq;
  const int C = heads * d;
  const float* qrow = q + static_cast<size_t>(idx) * C;
  const float* kb = k + static_cast<size_t>(b) * L * C;
  float acc[32], s[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int h = 0; h < heads; ++h) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = -INFINITY;
      if (j < L) {
        const float* kr = kb + static_cast<size_t>(j) * C + h * d;
        float dot = 0.f;
        for (int c = 0; c < d; ++c) dot = fmaf(qrow[h * d + c], kr[c], dot);
        s[j] = dot * scale;
        mx = fmaxf(mx, s[j]);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = (j < L) ? expf(s[j] - mx) : 0.f;
      sum += s[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += s[j] / sum;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < L) out[static_cast<size_t>(idx) * L + j] = acc[j];
}

// nearest-neighbour upsampling of a stored map (F.interpolate(..., scale_factor=(s, s), mode="nearest"), unet.py:1787-1797):
// src [B, H, W, L] -> dst [B, H s, W s, L]
__global__ void f32_upsample_map_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int H, int W, int L, int sc) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * H * sc * W * sc * L;
  if (idx >= total) return;
  const int j = static_cast<int>(idx % L);
  size_t r = idx / L;
  const int x = static_cast<int>(r % (W * sc));
  r /= (W * sc);
  const int y = static_cast<int>(r % (H * sc));
  const int b = static_cast<int>(r / (H * sc));
  dst[idx] = src[((static_cast<size_t>(b) * H + y / sc) * W + x / sc) * L + j];
}

// [Cout, Cin, 3, 3] -> [Cout][tap][Cin]
__global__ void f32_pack_conv_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(Cout) * Cin * 9;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % Cin);
  const int tap = static_cast<int>((idx / Cin) % 9);
  const size_t n = idx / (static_cast<size_t>(Cin) * 9);
  dst[idx] = w[(n * Cin + c) * 9 + tap];
}

// =====================================================================================================
// host side
// =====================================================================================================
struct Param {
  float* p = nullptr;
  size_t n = 0;
  std::vector<int64_t> shape;
  bool packed3x3 = false;
  float* hi = nullptr;  // TF32 split of p for the tensor-core GEMM (WD_F32_TC=1 only)
  float* lo = nullptr;
  // GEGLU.proj only (keys "...ff.net.0.proj.weight" / ".bias"): the rows in the 80-values-then-80-gates tile order of
  // f32tc_gemm_geglu (perm; for the weight also its TF32 split ghi / glo)
  float* perm = nullptr;
  float* ghi = nullptr;
  float* glo = nullptr;
  // Upsample convs only (keys "output_blocks.*.conv.weight", "...upsamplers.0.conv.weight"): the sub-pixel fold [4 Cout][4 Cin]
  // of f32tc_upconv (uf) and its TF32 split (uhi / ulo)
  float* uf = nullptr;
  float* uhi = nullptr;
  float* ulo = nullptr;
};

struct Act {  // token-major activation
  float* p = nullptr;
  int H = 0, W = 0, C = 0;
  // TF32 split of the tensor written by its producer (GroupNorm apply feeding a tensor-core convolution): when set, p is null
  // and the consumer skips its own split pass
  float* hi = nullptr;
  float* lo = nullptr;
};

}  // namespace

struct wd_f32 {
  wd_config cfg{};
  std::map<std::string, Param> params;
  float* pe = nullptr;
  // encoded context of the last wd_f32_encode_context
  float* ctx = nullptr;
  size_t ctx_cap = 0;
  int ctx_B = 0, ctx_L = 0;
  int* bad_flag = nullptr;
  // args.attentionMaps == 1: head-summed attn2 probabilities [B, H*W, L] of the last SpatialTransformer of the input blocks (0),
  // the middle block (1) and the output blocks (2); they live in the arena until the next evaluation
  struct MapRec {
    float* p = nullptr;
    int H = 0, W = 0, L = 0;
  } maps[3];
  int want_maps = 0, section = 0, maps_B = 0;
  // activation arena (per eval, bump allocated)
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0;
  bool dry = false;
  int launches = 0;
  cudaStream_t s = nullptr;
  std::string err;
};

namespace {

struct Fail {
  int code;
  std::string msg;
};

[[noreturn]] void fail(int code, const std::string& m) { throw Fail{code, m}; }

const Param& P(wd_f32* e, const std::string& k) {
  auto it = e->params.find(k);
  if (it == e->params.end()) fail(WD_ERR_STATE, "fp32 path: state_dict key not loaded: " + k);
  return it->second;
}
bool has(wd_f32* e, const std::string& k) { return e->params.count(k) != 0; }

float* alloc(wd_f32* e, size_t floats) {
  const size_t bytes = (floats * sizeof(float) + 255) & ~static_cast<size_t>(255);
  const size_t off = e->arena_off;
  e->arena_off += bytes;
  if (e->dry) return reinterpret_cast<float*>(static_cast<uintptr_t>(256));  // never dereferenced
  if (e->arena_off > e->arena_cap) fail(WD_ERR_STATE, "fp32 path: activation arena overflow");
  return reinterpret_cast<float*>(e->arena + off);
}

void after_launch(wd_f32* e, const char* what) {
  ++e->launches;
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) fail(WD_ERR_CUDA, std::string("fp32 path: ") + what + ": " + cudaGetErrorString(ce));
}

// generic implicit GEMM launch
struct ConvSpec {
  int taps = 1, stride = 1, up = 0, a_nchw = 0, out_nchw = 0, silu = 0;
};

Act gemm(wd_f32* e, const Act& a1, const Act* a2, int B, const float* w, int N, const float* bias, const float* rowbias,
         int rb_ld, const float* residual, const ConvSpec& cs, float* out_override = nullptr, const float* w_hi = nullptr,
         const float* w_lo = nullptr) {
  GemmF32 g{};
  g.a1 = a1.p;
  g.a2 = a2 ? a2->p : nullptr;
  g.C1 = a1.C;
  g.C2 = a2 ? a2->C : 0;
  if (a2 && (a2->H != a1.H || a2->W != a1.W)) fail(WD_ERR_INVALID, "fp32 path: concat sources differ in size");
  if ((g.C1 & 3) || (g.C2 & 3)) fail(WD_ERR_UNSUPPORTED, "fp32 path: channel counts must be multiples of 4");
  g.taps = cs.taps;
  g.Hin = a1.H;
  g.Win = a1.W;
  g.stride = cs.stride;
  g.up = cs.up;
  g.Hout = cs.up ? a1.H * 2 : (cs.stride == 2 ? a1.H / 2 : a1.H);
  g.Wout = cs.up ? a1.W * 2 : (cs.stride == 2 ? a1.W / 2 : a1.W);
  g.a_nchw = cs.a_nchw;
  g.w = w;
  g.bias = bias;
  g.rowbias = rowbias;
  g.rb_ld = rb_ld;
  g.residual = residual;
  g.out_nchw = cs.out_nchw;
  g.M = B * g.Hout * g.Wout;
  g.N = N;
  g.K = cs.taps * (g.C1 + g.C2);
  g.act_silu = cs.silu;
  Act o;
  o.H = g.Hout;
  o.W = g.Wout;
  o.C = N;
  o.p = out_override ? out_override : alloc(e, static_cast<size_t>(g.M) * N);
  g.out = o.p;
  // opt-in tensor-core route (split TF32, f32_gemm_tc.cu): operands are split into (hi, lo) pairs first -- a plain split for a
  // Linear / 1x1 conv input, the split patch matrix for a 3x3 conv
  // implicit 3x3 stride-1 convolution on the tensor cores: only the NHWC sources are split (no patch matrix)
  if (w_hi && w_lo && wd::f32tc_enabled() && cs.taps == 9 && cs.stride == 1 && !cs.up && !cs.a_nchw && !cs.out_nchw &&
      wd::f32tc_conv_ok(B, a1.H, a1.W, g.C1, g.C2, N)) {
    const size_t n1 = static_cast<size_t>(g.M) * g.C1, n2 = static_cast<size_t>(g.M) * g.C2;
    const bool presplit = a1.hi && a1.lo;  // written by the producer (GroupNorm apply)
    float* h1 = presplit ? a1.hi : alloc(e, n1);
    float* l1 = presplit ? a1.lo : alloc(e, n1);
    float* h2 = n2 ? alloc(e, n2) : nullptr;
    float* l2 = n2 ? alloc(e, n2) : nullptr;
    if (!e->dry) {
      cudaError_t ce = cudaSuccess;
      if (!presplit) {
        ce = wd::f32tc_split(a1.p, h1, l1, n1, e->s);
        ++e->launches;
      }
      if (ce == cudaSuccess && n2) {
        ce = wd::f32tc_split(a2->p, h2, l2, n2, e->s);
        ++e->launches;
      }
      if (ce == cudaSuccess) {
        ce = wd::f32tc_conv3x3(h1, l1, g.C1, h2, l2, g.C2, B, a1.H, a1.W, w_hi, w_lo, N, bias, rowbias, rb_ld, residual, o.p, cs.silu, e->s);
        ++e->launches;
      }
      if (ce != cudaSuccess) fail(WD_ERR_CUDA, std::string("fp32 path: tensor-core conv: ") + cudaGetErrorString(ce));
    }
    return o;
  }
  const bool presplit_lin = a1.hi && a1.lo && cs.taps == 1 && !a2;  // written by the producer (GEGLU epilogue)
  if (!a1.p && !(presplit_lin && w_hi && w_lo && wd::f32tc_enabled() && !cs.a_nchw && !cs.out_nchw && wd::f32tc_shape_ok(g.M, N, g.K)))
    fail(WD_ERR_STATE, "fp32 path: a split-only activation reached a contraction that cannot consume it");
  if (w_hi && w_lo && wd::f32tc_enabled() && !cs.a_nchw && !cs.out_nchw && wd::f32tc_shape_ok(g.M, N, g.K)) {
    const size_t nA = static_cast<size_t>(g.M) * g.K;
    float* a_hi = presplit_lin ? a1.hi : alloc(e, nA);
    float* a_lo = presplit_lin ? a1.lo : alloc(e, nA);
    const int splits = wd::f32tc_splits(g.K);
    float* ws = splits > 1 ? alloc(e, static_cast<size_t>(splits) * g.M * N) : nullptr;
    if (!e->dry) {
      cudaError_t ce = cudaSuccess;
      if (!presplit_lin) {
        ce = cs.taps == 9 ? wd::f32tc_im2col_split(a1.p, a2 ? a2->p : nullptr, g.C1, g.C2, B, a1.H, a1.W, cs.stride, cs.up, a_hi, a_lo, e->s)
             : a2       ? wd::f32tc_split_concat(a1.p, a2->p, g.C1, g.C2, static_cast<size_t>(g.M), a_hi, a_lo, e->s)
                        : wd::f32tc_split(a1.p, a_hi, a_lo, nA, e->s);
        ++e->launches;
      }
      if (ce == cudaSuccess) {
        ce = wd::f32tc_gemm(a_hi, a_lo, w_hi, w_lo, g.M, N, g.K, bias, rowbias, rb_ld, g.Hout * g.Wout, residual, o.p, cs.silu, ws, e->s);
        ++e->launches;
      }
      if (ce != cudaSuccess) fail(WD_ERR_CUDA, std::string("fp32 path: tensor-core gemm: ") + cudaGetErrorString(ce));
    }
    return o;
  }
  if (!e->dry) {
    launch_gemm(g, e->s);
    after_launch(e, "gemm");
  }
  return o;
}

// nn.Linear / 1x1 conv by state_dict prefix
Act linear(wd_f32* e, const std::string& pfx, const Act& a, int B, bool bias, const float* residual = nullptr, int silu = 0,
           const Act* a2 = nullptr) {
  const Param& w = P(e, pfx + ".weight");
  const int N = static_cast<int>(w.shape[0]);
  int K = 1;
  for (size_t i = 1; i < w.shape.size(); ++i) K *= static_cast<int>(w.shape[i]);
  if (K != a.C + (a2 ? a2->C : 0)) fail(WD_ERR_INVALID, "fp32 path: " + pfx + ": input width does not match the weight");
  ConvSpec cs;
  cs.silu = silu;
  return gemm(e, a, a2, B, w.p, N, bias ? P(e, pfx + ".bias").p : nullptr, nullptr, 0, residual, cs, nullptr, w.hi, w.lo);
}

static bool upconv_enabled() {  // env WD_F32_UPCONV (default on): Upsample + conv3x3 in sub-pixel form
  static int v = -1;
  if (v < 0) {
    const char* x = getenv("WD_F32_UPCONV");
    v = x ? (atoi(x) != 0) : 1;
  }
  return v != 0;
}

Act conv3x3(wd_f32* e, const std::string& pfx, const Act& a, const Act* a2, int B, const float* rowbias, int rb_ld,
            const float* residual, int stride, int up) {
  const Param& w = P(e, pfx + ".weight");
  if (!w.packed3x3) fail(WD_ERR_INVALID, "fp32 path: " + pfx + " is not a 3x3 convolution");
  if (w.shape[1] != a.C + (a2 ? a2->C : 0)) fail(WD_ERR_INVALID, "fp32 path: " + pfx + ": input channels do not match the weight");
  // nearest 2x upsample + conv in sub-pixel form on the tensor cores (f32tc_upconv): no upsampled tensor, no patch matrix
  if (up && stride == 1 && !a2 && !rowbias && !residual && w.uhi && w.ulo && a.p && upconv_enabled() && wd::f32tc_enabled() &&
      wd::f32tc_upconv_ok(B, a.H, a.W, a.C, static_cast<int>(w.shape[0]))) {
    const int Cout = static_cast<int>(w.shape[0]);
    const size_t nA = static_cast<size_t>(B) * a.H * a.W * a.C;
    float* a_hi = alloc(e, nA);
    float* a_lo = alloc(e, nA);
    Act o;
    o.H = 2 * a.H;
    o.W = 2 * a.W;
    o.C = Cout;
    o.p = alloc(e, static_cast<size_t>(B) * o.H * o.W * Cout);
    if (!e->dry) {
      cudaError_t ce = wd::f32tc_split(a.p, a_hi, a_lo, nA, e->s);
      ++e->launches;
      if (ce == cudaSuccess) {
        ce = wd::f32tc_upconv(a_hi, a_lo, a.C, B, a.H, a.W, w.uhi, w.ulo, Cout, P(e, pfx + ".bias").p, o.p, e->s);
        ++e->launches;
      }
      if (ce != cudaSuccess) fail(WD_ERR_CUDA, std::string("fp32 path: sub-pixel upsample conv: ") + cudaGetErrorString(ce));
    }
    return o;
  }
  ConvSpec cs;
  cs.taps = 9;
  cs.stride = stride;
  cs.up = up;
  return gemm(e, a, a2, B, w.p, static_cast<int>(w.shape[0]), P(e, pfx + ".bias").p, rowbias, rb_ld, residual, cs, nullptr, w.hi, w.lo);
}

static bool geglu_fused_enabled() {  // env WD_F32_GEGLU_FUSED (default on)
  static int v = -1;
  if (v < 0) {
    const char* x = getenv("WD_F32_GEGLU_FUSED");
    v = x ? (atoi(x) != 0) : 1;
  }
  return v != 0;
}
static bool gn_fast_enabled() {  // env WD_F32_GN_FAST (default on)
  static int v = -1;
  if (v < 0) {
    const char* x = getenv("WD_F32_GN_FAST");
    v = x ? (atoi(x) != 0) : 1;
  }
  return v != 0;
}

Act groupnorm(wd_f32* e, const std::string& pfx, const Act& a, const Act* a2, int B, float eps, int silu, bool split_out = false) {
  Act o;
  o.H = a.H;
  o.W = a.W;
  o.C = a.C + (a2 ? a2->C : 0);
  if (o.C % 32) fail(WD_ERR_UNSUPPORTED, "fp32 path: GroupNorm32 needs channels % 32 == 0");
  // coalesced statistics + apply kernels when every source holds whole groups and whole float4 vectors of at most 1024 channels
  const int HW = a.H * a.W, C1 = a.C, C2 = a2 ? a2->C : 0, cg = o.C / 32;
  const bool fast = gn_fast_enabled() && (C1 % 4) == 0 && (C2 % 4) == 0 && C1 % cg == 0 && C1 <= 1024 && C2 <= 1024;
  const int rpc = HW <= 512 ? 32 : (HW + 15) / 16, nch = (HW + rpc - 1) / rpc;
  float2* partial = fast ? reinterpret_cast<float2*>(alloc(e, static_cast<size_t>(B) * 32 * nch * 2)) : nullptr;
  const size_t n_out = static_cast<size_t>(B) * HW * o.C;
  const bool split = fast && split_out;  // the consumer is a tensor-core convolution: hand it the TF32 split, nothing else reads it
  if (split) {
    o.hi = alloc(e, n_out);
    o.lo = alloc(e, n_out);
  } else {
    o.p = alloc(e, n_out);
  }
  if (!e->dry) {
    const float* gm = P(e, pfx + ".weight").p;
    const float* bt = P(e, pfx + ".bias").p;
    if (fast) {
      f32_gn_stats_kernel<<<dim3(nch, B), GN32_T, 0, e->s>>>(a.p, a2 ? a2->p : nullptr, C1, C2, HW, 32, rpc, nch, partial);
      after_launch(e, "groupnorm statistics");
      if (split)
        f32_gn_apply_kernel<true><<<dim3(nch, B), GN32_T, 0, e->s>>>(a.p, a2 ? a2->p : nullptr, C1, C2, gm, bt, o.hi, o.lo, HW, 32, rpc, nch, partial, eps, silu);
      else
        f32_gn_apply_kernel<false><<<dim3(nch, B), GN32_T, 0, e->s>>>(a.p, a2 ? a2->p : nullptr, C1, C2, gm, bt, o.p, nullptr, HW, 32, rpc, nch, partial, eps, silu);
      after_launch(e, "groupnorm apply");
    } else {
      f32_groupnorm_kernel<<<dim3(32, B), 256, 0, e->s>>>(a.p, a2 ? a2->p : nullptr, C1, C2, gm, bt, o.p, HW, 32, eps, silu);
      after_launch(e, "groupnorm");
    }
  }
  return o;
}

// will linear(wpfx) over an [M, K] activation take the tensor-core route (the one that consumes a pre-split activation)?
bool lin_takes_split(wd_f32* e, const std::string& wpfx, int M, int K) {
  const Param& w = P(e, wpfx + ".weight");
  int Kw = 1;  // nn.Linear [N, K] or a 1x1 convolution [N, K, 1, 1] (as linear() reads it)
  for (size_t i = 1; i < w.shape.size(); ++i) Kw *= static_cast<int>(w.shape[i]);
  return w.hi && w.lo && !w.packed3x3 && Kw == K && wd::f32tc_enabled() && wd::f32tc_shape_ok(M, static_cast<int>(w.shape[0]), K);
}

Act layernorm(wd_f32* e, const std::string& pfx, const Act& a, int B, bool split_out = false) {
  Act o = a;
  const int M = B * a.H * a.W;
  const size_t n = static_cast<size_t>(M) * a.C;
  if (split_out) {  // every consumer is a tensor-core Linear: hand them the TF32 split, nothing else reads it
    o.p = nullptr;
    o.hi = alloc(e, n);
    o.lo = alloc(e, n);
  } else {
    o.p = alloc(e, n);
  }
  if (!e->dry) {
    if (split_out)
      f32_layernorm_kernel<true><<<(M + 7) / 8, 256, 0, e->s>>>(a.p, P(e, pfx + ".weight").p, P(e, pfx + ".bias").p, o.hi, o.lo, M, a.C, 1e-5f);
    else
      f32_layernorm_kernel<false><<<(M + 7) / 8, 256, 0, e->s>>>(a.p, P(e, pfx + ".weight").p, P(e, pfx + ".bias").p, o.p, nullptr, M, a.C, 1e-5f);
    after_launch(e, "layernorm");
  }
  return o;
}

void attention(wd_f32* e, const float* q, size_t q_bs, int ldq, const float* k, const float* v, size_t kv_bs, int ldkv, float* out,
               size_t o_bs, int ldo, int B, int Sq, int Skv, int heads, int d, float scale) {
  if (e->dry) return;
  const size_t warps = static_cast<size_t>(B) * heads * Sq;
  const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
  if (d == 80 && (ldq & 3) == 0 && (ldkv & 3) == 0 && (ldo & 3) == 0 && (q_bs & 3) == 0 && (kv_bs & 3) == 0 && (o_bs & 3) == 0)
    f32_attention_tq_kernel<80, 32><<<dim3((Sq + 127) / 128, heads, B), 128, 0, e->s>>>(q, q_bs, ldq, k, v, kv_bs, ldkv, out, o_bs, ldo, Sq,
                                                                                     Skv, scale);
  else if (d <= 96)
    f32_attention_kernel<3><<<grid, 256, 0, e->s>>>(q, q_bs, ldq, k, v, kv_bs, ldkv, out, o_bs, ldo, B, Sq, Skv, heads, d, scale);
  else if (d <= 160)
    f32_attention_kernel<5><<<grid, 256, 0, e->s>>>(q, q_bs, ldq, k, v, kv_bs, ldkv, out, o_bs, ldo, B, Sq, Skv, heads, d, scale);
  else if (d <= 320)
    f32_attention_kernel<10><<<grid, 256, 0, e->s>>>(q, q_bs, ldq, k, v, kv_bs, ldkv, out, o_bs, ldo, B, Sq, Skv, heads, d, scale);
  else if (d <= 512)  // the VAE decoder's single 512-wide head
    f32_attention_kernel<16><<<grid, 256, 0, e->s>>>(q, q_bs, ldq, k, v, kv_bs, ldkv, out, o_bs, ldo, B, Sq, Skv, heads, d, scale);
  else
    fail(WD_ERR_UNSUPPORTED, "fp32 path: attention head width > 512");
  after_launch(e, "attention");
}

int heads_of(const wd_config& c, int ch) { return c.num_head_channels > 0 ? ch / c.num_head_channels : c.num_heads; }

// CrossAttention.forward (unet.py:185-279 / unetPhosc.py:176-198); ctx == nullptr: self-attention.  Returns to_out(...) + residual
Act cross_attention(wd_f32* e, const std::string& pfx, const Act& xq, int B, const Act* ctx, const float* residual,
                    float* probs_out = nullptr) {
  const Act& kvsrc = ctx ? *ctx : xq;
  Act q = linear(e, pfx + "to_q", xq, B, false);
  Act k = linear(e, pfx + "to_k", kvsrc, B, false);
  Act v = linear(e, pfx + "to_v", kvsrc, B, false);
  const int Sq = xq.H * xq.W, Skv = kvsrc.H * kvsrc.W, inner = q.C;
  const int heads = heads_of(e->cfg, inner), d = inner / heads;
  if (probs_out) {
    if (Skv > 32) fail(WD_ERR_UNSUPPORTED, "fp32 path: attention maps need a context of at most 32 tokens");
    if (!e->dry) {
      f32_attn_probs_kernel<<<(B * Sq + 127) / 128, 128, 0, e->s>>>(q.p, k.p, probs_out, B, Sq, Skv, heads, d,
                                                                    1.0f / sqrtf(static_cast<float>(d)));
      after_launch(e, "attn_probs");
    }
  }
  Act o = q;
  o.p = alloc(e, static_cast<size_t>(B) * Sq * inner);
  attention(e, q.p, static_cast<size_t>(Sq) * inner, inner, k.p, v.p, static_cast<size_t>(Skv) * inner, inner, o.p,
            static_cast<size_t>(Sq) * inner, inner, B, Sq, Skv, heads, d, 1.0f / sqrtf(static_cast<float>(d)));
  return linear(e, pfx + "to_out.0", o, B, true, residual);
}

// SpatialTransformer.forward (unet.py:381-412 / unetPhosc.py:282-300) with BasicTransformerBlock (unet.py:337-345 / unetPhosc.py:241-246)
Act spatial_transformer(wd_f32* e, const std::string& pfx, const Act& x, int B, const Act& ctx) {
  Act n = groupnorm(e, pfx + "norm", x, nullptr, B, 1e-6f, 0, lin_takes_split(e, pfx + "proj_in", B * x.H * x.W, x.C));
  Act t = linear(e, pfx + "proj_in", n, B, true);
  for (int dpt = 0; dpt < e->cfg.transformer_depth; ++dpt) {
    const std::string bp = pfx + "transformer_blocks." + std::to_string(dpt) + ".";
    if (e->cfg.variant == WD_VARIANT_UNET) {
      const int Mt = B * t.H * t.W;
      Act l1 = layernorm(e, bp + "norm2", t, B, lin_takes_split(e, bp + "attn1.to_q", Mt, t.C));
      t = cross_attention(e, bp + "attn1.", l1, B, &ctx, t.p);
      Act l2 = layernorm(e, bp + "norm2", t, B, lin_takes_split(e, bp + "attn2.to_q", Mt, t.C));
      float* probs = nullptr;
      if (e->want_maps && dpt == e->cfg.transformer_depth - 1) {  // SpatialTransformer returns the LAST block's attn (unet.py:396-397)
        const int L = ctx.H * ctx.W;
        probs = alloc(e, static_cast<size_t>(B) * t.H * t.W * L);
        wd_f32::MapRec& mr = e->maps[e->section];  // a later SpatialTransformer of the section overwrites (unet.py:1660,1716)
        mr.p = probs;
        mr.H = t.H;
        mr.W = t.W;
        mr.L = L;
      }
      t = cross_attention(e, bp + "attn2.", l2, B, &ctx, t.p, probs);
    } else {
      const int Mt = B * t.H * t.W;
      Act l1 = layernorm(e, bp + "norm1", t, B,
                         lin_takes_split(e, bp + "attn1.to_q", Mt, t.C) && lin_takes_split(e, bp + "attn1.to_k", Mt, t.C) &&
                             lin_takes_split(e, bp + "attn1.to_v", Mt, t.C));
      t = cross_attention(e, bp + "attn1.", l1, B, nullptr, t.p);
      Act l2 = layernorm(e, bp + "norm2", t, B, lin_takes_split(e, bp + "attn2.to_q", Mt, t.C));
      t = cross_attention(e, bp + "attn2.", l2, B, &ctx, t.p);
    }
    Act l3 = layernorm(e, bp + "norm3", t, B, lin_takes_split(e, bp + "ff.net.0.proj", B * t.H * t.W, t.C));
    const size_t M = static_cast<size_t>(B) * t.H * t.W;
    const Param& wp = P(e, bp + "ff.net.0.proj.weight");
    const Param& bp_ = P(e, bp + "ff.net.0.proj.bias");
    const Param& w2 = P(e, bp + "ff.net.2.weight");
    const int Np = static_cast<int>(wp.shape[0]), Kp = static_cast<int>(wp.shape[1]), Hd = Np / 2;
    Act gg = t;
    gg.C = Hd;
    // GEGLU (unet.py:122-130) in the projection's own epilogue, the product written as the TF32 split the next Linear consumes:
    // the [M, 2 Hd] projection, the gating pass over it and the split pass never touch HBM
    const bool fused = geglu_fused_enabled() && wp.ghi && wp.glo && bp_.perm && w2.hi && w2.lo && wd::f32tc_enabled() && Kp == l3.C &&
                       wd::f32tc_shape_ok(static_cast<int>(M), Np, Kp) && Np % 320 == 0 &&
                       wd::f32tc_shape_ok(static_cast<int>(M), static_cast<int>(w2.shape[0]), Hd);
    if (fused) {
      const size_t nA = M * Kp;
      const bool l3_split = l3.hi && l3.lo;
      float* a_hi = l3_split ? l3.hi : alloc(e, nA);
      float* a_lo = l3_split ? l3.lo : alloc(e, nA);
      gg.p = nullptr;
      gg.hi = alloc(e, M * Hd);
      gg.lo = alloc(e, M * Hd);
      if (!e->dry) {
        cudaError_t ce = cudaSuccess;
        if (!l3_split) {
          ce = wd::f32tc_split(l3.p, a_hi, a_lo, nA, e->s);
          ++e->launches;
        }
        if (ce == cudaSuccess) {
          ce = wd::f32tc_gemm_geglu(a_hi, a_lo, wp.ghi, wp.glo, static_cast<int>(M), Np, Kp, bp_.perm, gg.hi, gg.lo, e->s);
          ++e->launches;
        }
        if (ce != cudaSuccess) fail(WD_ERR_CUDA, std::string("fp32 path: fused GEGLU projection: ") + cudaGetErrorString(ce));
      }
    } else {
      Act pr = linear(e, bp + "ff.net.0.proj", l3, B, true);
      gg.p = alloc(e, M * Hd);
      if (!e->dry) {
        f32_geglu_kernel<<<static_cast<unsigned>((M * Hd + 255) / 256), 256, 0, e->s>>>(pr.p, gg.p, M, Hd);
        after_launch(e, "geglu");
      }
    }
    t = linear(e, bp + "ff.net.2", gg, B, true, t.p);
  }
  return linear(e, pfx + "proj_out", t, B, true, x.p);
}

// will conv3x3(pfx) over a single [B, H, W, C] source run as the implicit tensor-core convolution (the route of gemm() that
// consumes a pre-split source)?  Same predicate as gemm().
bool conv_takes_split(wd_f32* e, const std::string& pfx, int H, int W, int C, int B) {
  const Param& w = P(e, pfx + ".weight");
  return w.packed3x3 && w.hi && w.lo && wd::f32tc_enabled() && wd::f32tc_conv_ok(B, H, W, C, 0, static_cast<int>(w.shape[0]));
}

// ResBlock._forward (unet.py:646-671), input = channel concat of a (and a2)
Act res_block(wd_f32* e, const std::string& pfx, const Act& a, const Act* a2, int B, const Act& semb) {
  Act g1 = groupnorm(e, pfx + "in_layers.0", a, a2, B, 1e-5f, 1, conv_takes_split(e, pfx + "in_layers.2", a.H, a.W, a.C + (a2 ? a2->C : 0), B));
  Act eo = linear(e, pfx + "emb_layers.1", semb, B, true);
  Act h1 = conv3x3(e, pfx + "in_layers.2", g1, nullptr, B, eo.p, eo.C, nullptr, 1, 0);
  Act g2 = groupnorm(e, pfx + "out_layers.0", h1, nullptr, B, 1e-5f, 1, conv_takes_split(e, pfx + "out_layers.3", h1.H, h1.W, h1.C, B));
  const float* skip;
  if (has(e, pfx + "skip_connection.weight")) {
    skip = linear(e, pfx + "skip_connection", a, B, true, nullptr, 0, a2).p;
  } else {
    if (a2) fail(WD_ERR_INVALID, "fp32 path: ResBlock without skip_connection fed by a concatenation");
    skip = a.p;
  }
  return conv3x3(e, pfx + "out_layers.3", g2, nullptr, B, nullptr, 0, skip, 1, 0);
}

// TimestepEmbedSequential.forward (unet.py:452-469): sub-layer kinds recovered from the keys
Act run_block(wd_f32* e, const std::string& pfx, Act h, const Act* cat, int B, const Act& semb, const Act& ctx, const float* x_nchw) {
  for (int j = 0;; ++j) {
    const std::string p = pfx + std::to_string(j) + ".";
    if (has(e, p + "in_layers.0.weight")) {
      h = res_block(e, p, h, j == 0 ? cat : nullptr, B, semb);
    } else if (has(e, p + "proj_in.weight")) {
      h = spatial_transformer(e, p, h, B, ctx);
    } else if (has(e, p + "op.weight")) {
      h = conv3x3(e, p + "op", h, nullptr, B, nullptr, 0, nullptr, 2, 0);
    } else if (has(e, p + "conv.weight")) {
      h = conv3x3(e, p + "conv", h, nullptr, B, nullptr, 0, nullptr, 1, 1);
    } else if (has(e, p + "weight")) {  // input_blocks.0.0: conv_in over the NCHW latent
      const Param& w = P(e, p + "weight");
      ConvSpec cs;
      cs.taps = 9;
      cs.a_nchw = 1;
      Act lat;
      lat.p = const_cast<float*>(x_nchw);
      lat.H = e->cfg.latent_h;
      lat.W = e->cfg.latent_w;
      lat.C = e->cfg.in_channels;
      if (lat.C != 4) fail(WD_ERR_UNSUPPORTED, "fp32 path: in_channels must be 4");
      h = gemm(e, lat, nullptr, B, w.p, static_cast<int>(w.shape[0]), P(e, p + "bias").p, nullptr, 0, nullptr, cs);
    } else {
      if (j == 0) fail(WD_ERR_STATE, "fp32 path: empty block " + pfx);
      break;
    }
  }
  return h;
}

void encode_one(wd_f32* e, int B, const long long* tok64, const int* tok32, int L, bool use_pe, float* out, size_t o_bs) {
  const int D = e->cfg.context_dim;
  Act x;
  x.H = 1;
  x.W = L;
  x.C = D;
  x.p = alloc(e, static_cast<size_t>(B) * L * D);
  if (!e->dry) {
    const size_t n = static_cast<size_t>(B) * L * D;
    f32_embed_tokens_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, e->s>>>(
        tok64, tok32, P(e, "word_emb.embedding.weight").p, use_pe ? e->pe : nullptr, x.p, B, L, D, e->cfg.vocab_size, e->bad_flag);
    after_launch(e, "embed_tokens");
  }
  Act q = linear(e, "word_emb.attention.linear_query", x, B, true);
  Act k = linear(e, "word_emb.attention.linear_key", x, B, true);
  Act v = linear(e, "word_emb.attention.linear_value", x, B, true);
  attention(e, q.p, static_cast<size_t>(L) * D, D, k.p, v.p, static_cast<size_t>(L) * D, D, out, o_bs, D, B, L, L, 1, D, 1.0f);
}

void encode_context_impl(wd_f32* e, int B, const long long* tokens, int L, const int* phosc) {
  const int D = e->cfg.context_dim;
  const bool is_unet = e->cfg.variant == WD_VARIANT_UNET;
  const int PL = (!is_unet && phosc) ? e->cfg.phosc_len : 0;
  const int Lt = L + PL;
  if (is_unet && L > e->cfg.max_seq_len) fail(WD_ERR_INVALID, "fp32 path: context longer than max_seq_len (unet.py:872 would fail to broadcast)");
  const size_t o_bs = static_cast<size_t>(Lt) * D;
  // CharacterEncoder (unet.py:851-874: PE always; unetPhosc.py:721-731: PE only if len <= max_seq_len)
  encode_one(e, B, tokens, nullptr, L, is_unet || L <= e->cfg.max_seq_len, e->ctx, o_bs);
  if (PL) encode_one(e, B, nullptr, phosc, PL, PL <= e->cfg.max_seq_len, e->ctx + static_cast<size_t>(L) * D, o_bs);
}

void unet_eval_impl(wd_f32* e, int B, const float* x, const long long* timesteps, long long t_scalar, const long long* y,
                    float* eps_out) {
  const wd_config& c = e->cfg;
  const int mc = c.model_channels, ted = 4 * mc;
  // a1 / a2: timestep embedding -> time_embed MLP (+ label embedding)
  Act temb;
  temb.H = temb.W = 1;
  temb.C = mc;
  temb.p = alloc(e, static_cast<size_t>(B) * mc);
  if (!e->dry) {
    f32_timestep_kernel<<<(B * (mc / 2) + 255) / 256, 256, 0, e->s>>>(timesteps, t_scalar, temb.p, B, mc);
    after_launch(e, "timestep");
  }
  Act e1 = linear(e, "time_embed.0", temb, B, true, nullptr, 1);
  Act emb = linear(e, "time_embed.2", e1, B, true);
  Act semb = emb;
  semb.p = alloc(e, static_cast<size_t>(B) * ted);
  const bool use_label = c.add_label_emb && c.num_classes > 0;
  if (use_label && !y) fail(WD_ERR_INVALID, "fp32 path: y is required (unet.py:1555)");
  if (!e->dry) {
    f32_emb_finish_kernel<<<(B * ted + 255) / 256, 256, 0, e->s>>>(emb.p, use_label ? P(e, "label_emb.weight").p : nullptr, y, semb.p,
                                                                   B, ted, c.num_classes, e->bad_flag);
    after_launch(e, "emb_finish");
  }
  Act ctx;
  ctx.H = 1;
  ctx.W = e->ctx_L;
  ctx.C = c.context_dim;
  ctx.p = e->ctx;

  std::vector<Act> hs;
  Act h{};
  int i = 0;
  e->section = 0;
  for (int w = 0; w < 3; ++w) e->maps[w] = wd_f32::MapRec{};
  for (;; ++i) {
    const std::string p = "input_blocks." + std::to_string(i) + ".";
    if (!(has(e, p + "0.weight") || has(e, p + "0.in_layers.0.weight") || has(e, p + "0.op.weight"))) break;
    h = run_block(e, p, h, nullptr, B, semb, ctx, x);
    hs.push_back(h);
  }
  if (hs.empty()) fail(WD_ERR_STATE, "fp32 path: no input_blocks loaded");
  e->section = 1;
  h = run_block(e, "middle_block.", h, nullptr, B, semb, ctx, x);
  e->section = 2;
  for (i = 0;; ++i) {
    const std::string p = "output_blocks." + std::to_string(i) + ".";
    if (!has(e, p + "0.in_layers.0.weight")) break;
    if (hs.empty()) fail(WD_ERR_STATE, "fp32 path: more output blocks than skip tensors");
    Act sk = hs.back();
    hs.pop_back();
    h = run_block(e, p, h, &sk, B, semb, ctx, x);
  }
  // out: GroupNorm32 -> SiLU -> conv3x3 (unet.py:1454-1458), written NCHW
  Act g = groupnorm(e, "out.0", h, nullptr, B, 1e-5f, 1);
  const Param& w = P(e, "out.2.weight");
  ConvSpec cs;
  cs.taps = 9;
  cs.out_nchw = 1;
  gemm(e, g, nullptr, B, w.p, static_cast<int>(w.shape[0]), P(e, "out.2.bias").p, nullptr, 0, nullptr, cs, eps_out);
}

int finish(wd_f32* e, const Fail& f) {
  e->err = f.msg;
  return wd_set_error(f.code, e->err.c_str());
}

int check_bad_flag(wd_f32* e, const char* what) {
  int bad = 0;
  if (cudaMemcpyAsync(&bad, e->bad_flag, sizeof(int), cudaMemcpyDeviceToHost, e->s) != cudaSuccess ||
      cudaStreamSynchronize(e->s) != cudaSuccess)
    return wd_set_error(WD_ERR_CUDA, "fp32 path: reading the index-check flag failed");
  if (bad) {
    cudaMemsetAsync(e->bad_flag, 0, sizeof(int), e->s);
    return wd_set_error(WD_ERR_INVALID, what);
  }
  return WD_OK;
}



// =====================================================================================================
// OCR head of args.ocrTraining == 1 (CTCtopC, unet.py:1054-1092) in eval mode: four (1 x 5) convolutions with BatchNorm (running
// statistics) + ReLU, a (1 x 5) convolution to the classes, Linear(32, 128) and Linear(128, 256) along the width, and
// `y.permute(2, 3, 0, 1)[0]`: only image row 0 leaves the head, and no layer mixes rows, so only row 0 is computed.
// Activations [B, W, C] (token-major); weights as stored ([Cout, Cin, 1, 5]).
// =====================================================================================================
__global__ void __launch_bounds__(256) ctc_conv1x5_kernel(const float* __restrict__ x, int x_nchw_H, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ bn_w,
                                                          const float* __restrict__ bn_b, const float* __restrict__ bn_rm,
                                                          const float* __restrict__ bn_rv, float* __restrict__ out, int B, int W,
                                                          int Cin, int Cout) {
  // one thread per (b, w, co); x_nchw_H > 0: x is the NCHW tensor [B, Cin, H, W], row 0 is read
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(B) * W * Cout) return;
  const int co = static_cast<int>(idx % Cout);
  const int wq = static_cast<int>((idx / Cout) % W);
  const int b = static_cast<int>(idx / (static_cast<size_t>(Cout) * W));
  float acc = bias[co];
  for (int tap = 0; tap < 5; ++tap) {
    const int wi = wq + tap - 2;
    if (wi < 0 || wi >= W) continue;
    const float* wr = w + static_cast<size_t>(co) * Cin * 5 + tap;
    if (x_nchw_H > 0) {
      for (int ci = 0; ci < Cin; ++ci)
        acc = fmaf(x[((static_cast<size_t>(b) * Cin + ci) * x_nchw_H) * W + wi], wr[ci * 5], acc);
    } else {
      const float* xr = x + (static_cast<size_t>(b) * W + wi) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(xr[ci], wr[ci * 5], acc);
    }
  }
  if (bn_w) {  // BatchNorm2d in eval mode (eps 1e-5) + ReLU
    acc = (acc - bn_rm[co]) / sqrtf(bn_rv[co] + 1e-5f) * bn_w[co] + bn_b[co];
    acc = fmaxf(acc, 0.f);
  }
  out[idx] = acc;
}
// t [B, W = 32, C] -> lin1 over the width -> lin2 -> out [256, B, C]; one CTA per (b, c)
__global__ void __launch_bounds__(256) ctc_lin_kernel(const float* __restrict__ t, const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out,
                                                      int B, int W, int C, int N1, int N2) {
  extern __shared__ float sm[];
  float* row = sm;        // [W]
  float* y1 = sm + W;     // [N1]
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  for (int i = threadIdx.x; i < W; i += blockDim.x) row[i] = t[(static_cast<size_t>(b) * W + i) * C + c];
  __syncthreads();
  for (int j = threadIdx.x; j < N1; j += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < W; ++i) acc = fmaf(row[i], w1[j * W + i], acc);
    y1[j] = acc + b1[j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N2; j += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < N1; ++i) acc = fmaf(y1[i], w2[j * N1 + i], acc);
    out[(static_cast<size_t>(j) * B + b) * C + c] = acc + b2[j];
  }
}

// =====================================================================================================
// VAE decode (reference train.py:239-247, regenerateFromtrain2.py:624-636):
//     latents = 1 / 0.18215 * x ; image = vae.decode(latents).sample ; image = (image / 2 + 0.5).clamp(0, 1)
// `vae` is diffusers' AutoencoderKL of Stable Diffusion v1 (train.py:415, AutoencoderKL.from_pretrained(..., subfolder="vae")).
// diffusers is not part of the reference tree (requirements only); its published decoder is restated here, layer sequence
// recovered from the loaded state_dict keys:
//     post_quant_conv (1x1) -> decoder.conv_in (3x3) -> mid_block: ResnetBlock2D, Attention (1 head over all pixels), ResnetBlock2D
//     -> up_blocks.i: resnets.j (ResnetBlock2D: GroupNorm32(eps 1e-6)+SiLU, conv3x3, GroupNorm32+SiLU, conv3x3, + shortcut /
//        1x1 conv_shortcut), upsamplers.0 (nearest x2 + conv3x3) -> conv_norm_out (GroupNorm32 + SiLU) -> conv_out (3x3)
// fp32 storage and arithmetic on the kernels of this file.
// =====================================================================================================
__global__ void vae_post_quant_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ b,
                                      float* __restrict__ out, int n, int C, size_t plane, float scale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(n) * plane) return;
  const size_t img = i / plane, p = i - img * plane;
  const float* zi = z + img * C * plane + p;
  float* oi = out + img * C * plane + p;
  for (int o = 0; o < C; ++o) {
    float acc = b ? b[o] : 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(w ? w[o * C + c] : (o == c ? 1.f : 0.f), zi[c * plane] * scale, acc);
    oi[o * plane] = acc;
  }
}
__global__ void vae_postprocess_kernel(float* __restrict__ img, size_t n) {  // (image / 2 + 0.5).clamp(0, 1), train.py:243
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) img[i] = fminf(fmaxf(img[i] / 2.f + 0.5f, 0.f), 1.f);
}

Act vae_resnet(wd_f32* e, const std::string& pfx, const Act& x, int B) {
  Act n1 = groupnorm(e, pfx + "norm1", x, nullptr, B, 1e-6f, 1);
  Act c1 = conv3x3(e, pfx + "conv1", n1, nullptr, B, nullptr, 0, nullptr, 1, 0);
  Act n2 = groupnorm(e, pfx + "norm2", c1, nullptr, B, 1e-6f, 1);
  const float* res = x.p;
  if (has(e, pfx + "conv_shortcut.weight")) res = linear(e, pfx + "conv_shortcut", x, B, true).p;
  else if (has(e, pfx + "nin_shortcut.weight")) res = linear(e, pfx + "nin_shortcut", x, B, true).p;
  return conv3x3(e, pfx + "conv2", n2, nullptr, B, nullptr, 0, res, 1, 0);
}

Act vae_attention(wd_f32* e, const std::string& pfx, const Act& x, int B) {
  // diffusers >= 0.15 names the projections to_q / to_k / to_v / to_out.0, older checkpoints query / key / value / proj_attn
  const bool new_names = has(e, pfx + "to_q.weight");
  const std::string nq = new_names ? "to_q" : "query", nk = new_names ? "to_k" : "key", nv = new_names ? "to_v" : "value",
                    no = new_names ? "to_out.0" : "proj_attn";
  Act n = groupnorm(e, pfx + "group_norm", x, nullptr, B, 1e-6f, 0);
  Act q = linear(e, pfx + nq, n, B, true);
  Act k = linear(e, pfx + nk, n, B, true);
  Act v = linear(e, pfx + nv, n, B, true);
  const int S = x.H * x.W, C = q.C;
  Act o = q;
  o.p = alloc(e, static_cast<size_t>(B) * S * C);
  attention(e, q.p, static_cast<size_t>(S) * C, C, k.p, v.p, static_cast<size_t>(S) * C, C, o.p, static_cast<size_t>(S) * C, C, B, S, S, 1,
            C, 1.0f / sqrtf(static_cast<float>(C)));
  return linear(e, pfx + no, o, B, true, x.p);
}

void vae_decode_impl(wd_f32* e, int B, const float* latents, int h, int w, float scale, int postprocess, float* images) {
  const Param& cin = P(e, "decoder.conv_in.weight");
  const int zc = static_cast<int>(cin.shape[1]);
  if (zc != 4) fail(WD_ERR_UNSUPPORTED, "vae decode: latent_channels must be 4");
  const size_t plane = static_cast<size_t>(h) * w;
  float* z = alloc(e, static_cast<size_t>(B) * zc * plane);
  if (!e->dry) {
    const bool pq = has(e, "post_quant_conv.weight");
    const size_t tot = static_cast<size_t>(B) * plane;
    vae_post_quant_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, e->s>>>(
        latents, pq ? P(e, "post_quant_conv.weight").p : nullptr, pq ? P(e, "post_quant_conv.bias").p : nullptr, z, B, zc, plane, scale);
    after_launch(e, "vae post_quant_conv");
  }
  Act zin;
  zin.p = z;
  zin.H = h;
  zin.W = w;
  zin.C = zc;
  ConvSpec cs;
  cs.taps = 9;
  cs.a_nchw = 1;
  Act x = gemm(e, zin, nullptr, B, cin.p, static_cast<int>(cin.shape[0]), P(e, "decoder.conv_in.bias").p, nullptr, 0, nullptr, cs);
  x = vae_resnet(e, "decoder.mid_block.resnets.0.", x, B);
  if (has(e, "decoder.mid_block.attentions.0.group_norm.weight")) x = vae_attention(e, "decoder.mid_block.attentions.0.", x, B);
  x = vae_resnet(e, "decoder.mid_block.resnets.1.", x, B);
  for (int i = 0; has(e, "decoder.up_blocks." + std::to_string(i) + ".resnets.0.norm1.weight"); ++i) {
    const std::string bp = "decoder.up_blocks." + std::to_string(i) + ".";
    for (int j = 0; has(e, bp + "resnets." + std::to_string(j) + ".norm1.weight"); ++j)
      x = vae_resnet(e, bp + "resnets." + std::to_string(j) + ".", x, B);
    if (has(e, bp + "upsamplers.0.conv.weight")) x = conv3x3(e, bp + "upsamplers.0.conv", x, nullptr, B, nullptr, 0, nullptr, 1, 1);
  }
  Act n = groupnorm(e, "decoder.conv_norm_out", x, nullptr, B, 1e-6f, 1);
  const Param& cout = P(e, "decoder.conv_out.weight");
  if (!cout.packed3x3 || cout.shape[1] != n.C) fail(WD_ERR_INVALID, "vae decode: decoder.conv_out does not match the last block");
  ConvSpec co;
  co.taps = 9;
  co.out_nchw = 1;
  Act img = gemm(e, n, nullptr, B, cout.p, static_cast<int>(cout.shape[0]), P(e, "decoder.conv_out.bias").p, nullptr, 0, nullptr, co, images);
  if (postprocess && !e->dry) {
    const size_t tot = static_cast<size_t>(B) * img.C * img.H * img.W;
    vae_postprocess_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, e->s>>>(images, tot);
    after_launch(e, "vae postprocess");
  }
}

}  // namespace

extern "C" {

int wd_f32_create(const wd_config* cfg, wd_f32** out) {
  if (!cfg || !out) return wd_set_error(WD_ERR_INVALID, "wd_f32_create: null argument");
  int dev = 0;
  cudaDeviceProp prop{};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return wd_set_error(WD_ERR_CUDA, "wd_f32_create: no CUDA device");
  if (prop.major != 10) return wd_set_error(WD_ERR_UNSUPPORTED, "wd_f32_create: libwd_b200 is built for sm_100a only");
  if (cfg->model_channels % 32 || cfg->context_dim % 4 || cfg->in_channels != 4)
    return wd_set_error(WD_ERR_UNSUPPORTED, "wd_f32_create: unsupported channel configuration");
  wd_f32* e = new wd_f32();
  e->cfg = *cfg;
  if (cudaMalloc(&e->bad_flag, sizeof(int)) != cudaSuccess || cudaMemset(e->bad_flag, 0, sizeof(int)) != cudaSuccess) {
    delete e;
    return wd_set_error(WD_ERR_CUDA, "wd_f32_create: cudaMalloc failed");
  }
  *out = e;
  return WD_OK;
}

void wd_f32_destroy(wd_f32* e) {
  if (!e) return;
  for (auto& kv : e->params) {
    cudaFree(kv.second.p);
    cudaFree(kv.second.hi);
    cudaFree(kv.second.lo);
    cudaFree(kv.second.perm);
    cudaFree(kv.second.ghi);
    cudaFree(kv.second.glo);
    cudaFree(kv.second.uf);
    cudaFree(kv.second.uhi);
    cudaFree(kv.second.ulo);
  }
  cudaFree(e->pe);
  cudaFree(e->ctx);
  cudaFree(e->arena);
  cudaFree(e->bad_flag);
  delete e;
}

int wd_f32_load_param(wd_f32* e, const char* name, const float* src, const int64_t* shape, int ndim, void* stream) {
  if (!e || !name || !src || (ndim > 0 && !shape)) return wd_set_error(WD_ERR_INVALID, "wd_f32_load_param: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  size_t n = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) {
    n *= static_cast<size_t>(shape[i]);
    shp.push_back(shape[i]);
  }
  Param& p = e->params[name];
  if (p.n != n) {
    cudaFree(p.p);
    p.p = nullptr;
    if (cudaMalloc(&p.p, n * sizeof(float)) != cudaSuccess) {
      e->params.erase(name);
      return wd_set_error(WD_ERR_CUDA, "wd_f32_load_param: cudaMalloc failed");
    }
    p.n = n;
  }
  p.shape = shp;
  p.packed3x3 = ndim == 4 && shape[2] == 3 && shape[3] == 3;
  cudaError_t ce;
  if (p.packed3x3) {
    f32_pack_conv_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, p.p, static_cast<int>(shape[0]),
                                                                               static_cast<int>(shape[1]));
    ce = cudaGetLastError();
  } else {
    ce = cudaMemcpyAsync(p.p, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s);
  }
  if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  if (wd::f32tc_enabled() && ndim >= 2 && (n & 3) == 0) {  // TF32 split of every weight matrix for the tensor-core GEMM
    cudaFree(p.hi);
    cudaFree(p.lo);
    p.hi = p.lo = nullptr;
    if (cudaMalloc(&p.hi, n * sizeof(float)) != cudaSuccess || cudaMalloc(&p.lo, n * sizeof(float)) != cudaSuccess)
      return wd_set_error(WD_ERR_CUDA, "wd_f32_load_param: cudaMalloc of the TF32 split failed");
    ce = wd::f32tc_split(p.p, p.hi, p.lo, n, s);
    if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  }
  // GEGLU.proj (unet.py:125): a second copy in the tile order of the fused GEGLU epilogue
  {
    const std::string nm(name);
    auto ends_with = [&](const char* suf) {
      const size_t l = strlen(suf);
      return nm.size() >= l && nm.compare(nm.size() - l, l, suf) == 0;
    };
    // Upsample convs (unet.py:472-500; diffusers Upsample2D): sub-pixel weights for f32tc_upconv
    if (wd::f32tc_enabled() && p.packed3x3 && ends_with(".conv.weight") &&
        (nm.find("output_blocks.") != std::string::npos || nm.find("upsamplers.") != std::string::npos)) {
      const int Co = static_cast<int>(shape[0]), Ci = static_cast<int>(shape[1]);
      const size_t nf = static_cast<size_t>(16) * Co * Ci;
      cudaFree(p.uf);
      cudaFree(p.uhi);
      cudaFree(p.ulo);
      p.uf = p.uhi = p.ulo = nullptr;
      if (cudaMalloc(&p.uf, nf * sizeof(float)) != cudaSuccess || cudaMalloc(&p.uhi, nf * sizeof(float)) != cudaSuccess ||
          cudaMalloc(&p.ulo, nf * sizeof(float)) != cudaSuccess)
        return wd_set_error(WD_ERR_CUDA, "wd_f32_load_param: cudaMalloc failed");
      ce = wd::f32tc_upconv_fold(p.p, p.uf, Co, Ci, s);
      if (ce == cudaSuccess) ce = wd::f32tc_split(p.uf, p.uhi, p.ulo, nf, s);
      if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
    }
    const bool gw = ends_with("ff.net.0.proj.weight") && ndim == 2, gb = ends_with("ff.net.0.proj.bias") && ndim == 1;
    if (wd::f32tc_enabled() && (gw || gb) && shape[0] % 320 == 0 && (n & 3) == 0) {
      cudaFree(p.perm);
      cudaFree(p.ghi);
      cudaFree(p.glo);
      p.perm = p.ghi = p.glo = nullptr;
      if (cudaMalloc(&p.perm, n * sizeof(float)) != cudaSuccess) return wd_set_error(WD_ERR_CUDA, "wd_f32_load_param: cudaMalloc failed");
      ce = wd::f32tc_geglu_permute(p.p, p.perm, static_cast<int>(shape[0]), gw ? static_cast<int>(shape[1]) : 1, s);
      if (ce == cudaSuccess && gw) {
        if (cudaMalloc(&p.ghi, n * sizeof(float)) != cudaSuccess || cudaMalloc(&p.glo, n * sizeof(float)) != cudaSuccess)
          return wd_set_error(WD_ERR_CUDA, "wd_f32_load_param: cudaMalloc failed");
        ce = wd::f32tc_split(p.perm, p.ghi, p.glo, n, s);
      }
      if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
    }
  }
  return WD_OK;
}

int wd_f32_set_pos_encoding(wd_f32* e, const float* pe, void* stream) {
  if (!e || !pe) return wd_set_error(WD_ERR_INVALID, "wd_f32_set_pos_encoding: null argument");
  const size_t n = static_cast<size_t>(e->cfg.max_seq_len) * e->cfg.context_dim;
  if (!e->pe && cudaMalloc(&e->pe, n * sizeof(float)) != cudaSuccess)
    return wd_set_error(WD_ERR_CUDA, "wd_f32_set_pos_encoding: cudaMalloc failed");
  if (cudaMemcpyAsync(e->pe, pe, n * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return wd_set_error(WD_ERR_CUDA, "wd_f32_set_pos_encoding: copy failed");
  return WD_OK;
}

static int ensure_arena(wd_f32* e, size_t need) {
  if (need <= e->arena_cap) return WD_OK;
  if (e->arena) {
    cudaDeviceSynchronize();
    cudaFree(e->arena);
    e->arena = nullptr;
    e->arena_cap = 0;
  }
  if (cudaMalloc(&e->arena, need) != cudaSuccess) return wd_set_error(WD_ERR_CUDA, "fp32 path: activation arena cudaMalloc failed");
  e->arena_cap = need;
  return WD_OK;
}

int wd_f32_encode_context(wd_f32* e, int batch, const int64_t* ctx_tokens, int L, const int32_t* phosc, void* stream) {
  if (!e || !ctx_tokens || batch <= 0 || L <= 0) return wd_set_error(WD_ERR_INVALID, "wd_f32_encode_context: invalid argument");
  if (!e->pe) return wd_set_error(WD_ERR_STATE, "wd_f32_encode_context: positional encoding not set");
  if (e->cfg.variant != WD_VARIANT_UNET && e->cfg.phosc_len > 0 && !phosc)
    return wd_set_error(WD_ERR_INVALID, "wd_f32_encode_context: this model needs the PHOSC labels");
  e->s = static_cast<cudaStream_t>(stream);
  const int Lt = L + ((e->cfg.variant != WD_VARIANT_UNET && phosc) ? e->cfg.phosc_len : 0);
  const size_t need = static_cast<size_t>(batch) * Lt * e->cfg.context_dim;
  if (need > e->ctx_cap) {
    cudaDeviceSynchronize();
    cudaFree(e->ctx);
    e->ctx = nullptr;
    e->ctx_cap = 0;
    if (cudaMalloc(&e->ctx, need * sizeof(float)) != cudaSuccess) return wd_set_error(WD_ERR_CUDA, "wd_f32_encode_context: cudaMalloc failed");
    e->ctx_cap = need;
  }
  try {
    e->dry = true;
    e->arena_off = 0;
    encode_context_impl(e, batch, reinterpret_cast<const long long*>(ctx_tokens), L, phosc);
    const size_t need_arena = e->arena_off;
    int rc = ensure_arena(e, need_arena);
    if (rc != WD_OK) return rc;
    e->dry = false;
    e->arena_off = 0;
    e->launches = 0;
    encode_context_impl(e, batch, reinterpret_cast<const long long*>(ctx_tokens), L, phosc);
  } catch (const Fail& f) {
    e->dry = false;
    return finish(e, f);
  } catch (const std::exception& ex) {  // nothing may cross the C ABI
    e->dry = false;
    return wd_set_error(WD_ERR_STATE, ex.what());
  }
  e->ctx_B = batch;
  e->ctx_L = Lt;
  return check_bad_flag(e, "wd_f32_encode_context: token id outside the embedding table");
}

static int unet_eval_common(wd_f32* e, int batch, const float* x, const int64_t* timesteps, int64_t t_scalar, const int64_t* y,
                            float* eps_out, void* stream, int want_maps) {
  if (!e || !x || !eps_out || batch <= 0) return wd_set_error(WD_ERR_INVALID, "wd_f32_unet_eval: invalid argument");
  if (!e->ctx || e->ctx_B != batch) return wd_set_error(WD_ERR_STATE, "wd_f32_unet_eval: call wd_f32_encode_context for this batch first");
  if (want_maps && e->cfg.variant != WD_VARIANT_UNET)
    return wd_set_error(WD_ERR_UNSUPPORTED, "wd_f32_unet_eval_maps: attention maps exist for unet.UNetModel only (unet.py:1645)");
  e->s = static_cast<cudaStream_t>(stream);
  e->want_maps = want_maps;
  e->maps_B = want_maps ? batch : 0;
  try {
    e->dry = true;
    e->arena_off = 0;
    unet_eval_impl(e, batch, x, reinterpret_cast<const long long*>(timesteps), t_scalar, reinterpret_cast<const long long*>(y), eps_out);
    const size_t need_arena = e->arena_off;
    int rc = ensure_arena(e, need_arena);
    if (rc != WD_OK) return rc;
    e->dry = false;
    e->arena_off = 0;
    e->launches = 0;
    unet_eval_impl(e, batch, x, reinterpret_cast<const long long*>(timesteps), t_scalar, reinterpret_cast<const long long*>(y), eps_out);
  } catch (const Fail& f) {
    e->dry = false;
    return finish(e, f);
  } catch (const std::exception& ex) {  // nothing may cross the C ABI
    e->dry = false;
    return wd_set_error(WD_ERR_STATE, ex.what());
  }
  return WD_OK;
}

int wd_f32_unet_eval(wd_f32* e, int batch, const float* x, const int64_t* timesteps, int64_t t_scalar, const int64_t* y,
                     float* eps_out, void* stream) {
  return unet_eval_common(e, batch, x, timesteps, t_scalar, y, eps_out, stream, 0);
}

int wd_f32_unet_eval_maps(wd_f32* e, int batch, const float* x, const int64_t* timesteps, int64_t t_scalar, const int64_t* y,
                          float* eps_out, void* stream) {
  return unet_eval_common(e, batch, x, timesteps, t_scalar, y, eps_out, stream, 1);
}

int wd_f32_read_attention_map(wd_f32* e, int which, int scale, float* dst, int* H, int* W, int* L, void* stream) {
  if (!e || which < 0 || which > 2 || scale < 1) return wd_set_error(WD_ERR_INVALID, "wd_f32_read_attention_map: invalid argument");
  const wd_f32::MapRec& m = e->maps[which];
  if (!e->maps_B || !m.p) return wd_set_error(WD_ERR_STATE, "wd_f32_read_attention_map: no map stored (call wd_f32_unet_eval_maps first)");
  if (H) *H = m.H;
  if (W) *W = m.W;
  if (L) *L = m.L;
  if (!dst) return WD_OK;
  const size_t total = static_cast<size_t>(e->maps_B) * m.H * scale * m.W * scale * m.L;
  f32_upsample_map_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      m.p, dst, e->maps_B, m.H, m.W, m.L, scale);
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  return WD_OK;
}

int wd_f32_read_context(wd_f32* e, float* dst, size_t bytes, void* stream) {
  if (!e || !dst) return wd_set_error(WD_ERR_INVALID, "wd_f32_read_context: null argument");
  if (!e->ctx || !e->ctx_B) return wd_set_error(WD_ERR_STATE, "wd_f32_read_context: no context encoded");
  const size_t have = static_cast<size_t>(e->ctx_B) * e->ctx_L * e->cfg.context_dim * sizeof(float);
  if (bytes != have) return wd_set_error(WD_ERR_INVALID, "wd_f32_read_context: size mismatch");
  if (cudaMemcpyAsync(dst, e->ctx, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return wd_set_error(WD_ERR_CUDA, "wd_f32_read_context: copy failed");
  return WD_OK;
}


/* ---- variants of unet.UNetModel: dense context, style interpolation, OCR head ---- */
int wd_f32_set_context(wd_f32* e, int batch, const float* ctx, int L, void* stream) {
  if (!e || !ctx || batch <= 0 || L <= 0) return wd_set_error(WD_ERR_INVALID, "wd_f32_set_context: invalid argument");
  const size_t need = static_cast<size_t>(batch) * L * e->cfg.context_dim;
  if (need > e->ctx_cap) {
    cudaDeviceSynchronize();
    cudaFree(e->ctx);
    e->ctx = nullptr;
    e->ctx_cap = 0;
    if (cudaMalloc(&e->ctx, need * sizeof(float)) != cudaSuccess) return wd_set_error(WD_ERR_CUDA, "wd_f32_set_context: cudaMalloc failed");
    e->ctx_cap = need;
  }
  if (cudaMemcpyAsync(e->ctx, ctx, need * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return wd_set_error(WD_ERR_CUDA, "wd_f32_set_context: copy failed");
  e->ctx_B = batch;
  e->ctx_L = L;
  return WD_OK;
}

__global__ void f32_label_mix_kernel(float* __restrict__ w, int D, int row, int s1, int s2, float mix) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  w[static_cast<size_t>(row) * D + c] =
      __fadd_rn(__fmul_rn(1.0f - mix, w[static_cast<size_t>(s1) * D + c]), __fmul_rn(mix, w[static_cast<size_t>(s2) * D + c]));
}
int wd_f32_set_label_mix(wd_f32* e, int row, int s1, int s2, float mix, void* stream) {
  if (!e || !e->params.count("label_emb.weight")) return wd_set_error(WD_ERR_STATE, "wd_f32_set_label_mix: no label embedding loaded");
  const Param& p = e->params["label_emb.weight"];
  const int n = static_cast<int>(p.shape[0]), D = static_cast<int>(p.shape[1]);
  if (row < 0 || row >= n || s1 < 0 || s1 >= n || s2 < 0 || s2 >= n) return wd_set_error(WD_ERR_INVALID, "wd_f32_set_label_mix: class out of range");
  f32_label_mix_kernel<<<(D + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(p.p, D, row, s1, s2, mix);
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  return WD_OK;
}

/* tdec = auxhead(eps) of args.ocrTraining == 1 (CTCtopC, unet.py:1054-1092,1829) in eval mode: eps fp32 NCHW [B, C, H, W = 32]
 * -> out fp32 [256, B, nclasses].  The auxhead.* entries must have been loaded with wd_f32_load_param. */
int wd_f32_ctc_head(wd_f32* e, int batch, const float* eps, int C, int H, int W, float* out, void* stream) {
  if (!e || !eps || !out || batch <= 0) return wd_set_error(WD_ERR_INVALID, "wd_f32_ctc_head: invalid argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float *a = nullptr, *b = nullptr;
  int rc = WD_OK;
  try {
    const Param& wi = P(e, "auxhead.temporal_i.0.weight");
    const int hid = static_cast<int>(wi.shape[0]);
    if (wi.shape.size() != 4 || wi.shape[1] != C || wi.shape[2] != 1 || wi.shape[3] != 5) fail(WD_ERR_INVALID, "ctc head: temporal_i shape");
    const Param& l1 = P(e, "auxhead.lin1.weight");
    const Param& l2 = P(e, "auxhead.lin2.weight");
    if (l1.shape[1] != W || l2.shape[1] != l1.shape[0]) fail(WD_ERR_INVALID, "ctc head: lin1 expects the latent width (32)");
    const Param& wo = P(e, "auxhead.temporal_o.weight");
    const int ncls = static_cast<int>(wo.shape[0]);
    const size_t n_act = static_cast<size_t>(batch) * W * (hid > ncls ? hid : ncls);
    if (cudaMalloc(&a, n_act * sizeof(float)) != cudaSuccess || cudaMalloc(&b, n_act * sizeof(float)) != cudaSuccess)
      fail(WD_ERR_CUDA, "ctc head: cudaMalloc failed");
    auto conv = [&](const float* x, int nchw_h, const std::string& cw, const std::string& bn, float* o, int cin, int cout) {
      const size_t tot = static_cast<size_t>(batch) * W * cout;
      const bool has_bn = !bn.empty();
      ctc_conv1x5_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(
          x, nchw_h, P(e, cw + ".weight").p, P(e, cw + ".bias").p, has_bn ? P(e, bn + ".weight").p : nullptr,
          has_bn ? P(e, bn + ".bias").p : nullptr, has_bn ? P(e, bn + ".running_mean").p : nullptr,
          has_bn ? P(e, bn + ".running_var").p : nullptr, o, batch, W, cin, cout);
      if (cudaGetLastError() != cudaSuccess) fail(WD_ERR_CUDA, "ctc head: conv launch failed");
    };
    conv(eps, H, "auxhead.temporal_i.0", "auxhead.temporal_i.1", a, C, hid);
    float *cur = a, *nxt = b;
    for (int i = 0; has(e, "auxhead.temporal_m." + std::to_string(i) + ".0.weight"); ++i) {
      const std::string pfx = "auxhead.temporal_m." + std::to_string(i);
      conv(cur, 0, pfx + ".0", pfx + ".1", nxt, hid, hid);
      std::swap(cur, nxt);
    }
    conv(cur, 0, "auxhead.temporal_o", "", nxt, hid, ncls);
    const int N1 = static_cast<int>(l1.shape[0]), N2 = static_cast<int>(l2.shape[0]);
    ctc_lin_kernel<<<batch * ncls, 256, (W + N1) * sizeof(float), s>>>(nxt, l1.p, P(e, "auxhead.lin1.bias").p, l2.p,
                                                                      P(e, "auxhead.lin2.bias").p, out, batch, W, ncls, N1, N2);
    if (cudaGetLastError() != cudaSuccess) fail(WD_ERR_CUDA, "ctc head: linear launch failed");
  } catch (const Fail& f) {
    rc = wd_set_error(f.code, f.msg.c_str());
  } catch (const std::exception& ex) {
    rc = wd_set_error(WD_ERR_STATE, ex.what());
  }
  cudaStreamSynchronize(s);
  cudaFree(a);
  cudaFree(b);
  return rc;
}

/* out[M, N] = x[M, K] w[N, K]^T + bias, fp32 FFMA (wrd_proj of args.wrdChrWrStyl == 1, unet.py:1590-1591) */
int wd_f32_op_linear(const float* x, const float* w, const float* bias, float* out, int M, int N, int K, void* stream) {
  if (!x || !w || !out || M <= 0 || N <= 0 || K <= 0 || (K & 3)) return wd_set_error(WD_ERR_INVALID, "wd_f32_op_linear: invalid argument");
  GemmF32 g{};
  g.a1 = x;
  g.C1 = K;
  g.taps = 1;
  g.Hin = g.Win = g.Hout = g.Wout = 1;
  g.stride = 1;
  g.w = w;
  g.bias = bias;
  g.out = out;
  g.M = M;
  g.N = N;
  g.K = K;
  launch_gemm(g, static_cast<cudaStream_t>(stream));
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  return WD_OK;
}

/* ---- VAE decode (AutoencoderKL decoder) ---- */
int wd_vae_create(wd_f32** out) {
  if (!out) return wd_set_error(WD_ERR_INVALID, "wd_vae_create: null argument");
  int dev = 0;
  cudaDeviceProp prop{};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return wd_set_error(WD_ERR_CUDA, "wd_vae_create: no CUDA device");
  if (prop.major != 10) return wd_set_error(WD_ERR_UNSUPPORTED, "wd_vae_create: libwd_b200 is built for sm_100a only");
  wd_f32* e = new wd_f32();
  *out = e;
  return WD_OK;
}

int wd_vae_decode(wd_f32* e, int n, const float* latents, int h, int w, float scale, int postprocess, float* images, int chunk,
                  void* stream) {
  if (!e || !latents || !images || n < 0 || h <= 0 || w <= 0) return wd_set_error(WD_ERR_INVALID, "wd_vae_decode: invalid argument");
  if (chunk <= 0) chunk = 32;
  e->s = static_cast<cudaStream_t>(stream);
  try {
    int ups = 0, out_c = 3;
    for (int i = 0; has(e, "decoder.up_blocks." + std::to_string(i) + ".resnets.0.norm1.weight"); ++i)
      if (has(e, "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv.weight")) ++ups;
    out_c = static_cast<int>(P(e, "decoder.conv_out.weight").shape[0]);
    const size_t in_stride = static_cast<size_t>(4) * h * w;
    const size_t out_stride = static_cast<size_t>(out_c) * (static_cast<size_t>(h) << ups) * (static_cast<size_t>(w) << ups);
    e->launches = 0;
    for (int b0 = 0; b0 < n; b0 += chunk) {
      const int B = n - b0 < chunk ? n - b0 : chunk;
      e->dry = true;
      e->arena_off = 0;
      vae_decode_impl(e, B, latents + b0 * in_stride, h, w, scale, postprocess, images + b0 * out_stride);
      const int rc = ensure_arena(e, e->arena_off);
      if (rc != WD_OK) {
        e->dry = false;
        return rc;
      }
      e->dry = false;
      e->arena_off = 0;
      vae_decode_impl(e, B, latents + b0 * in_stride, h, w, scale, postprocess, images + b0 * out_stride);
    }
  } catch (const Fail& f) {
    e->dry = false;
    return finish(e, f);
  } catch (const std::exception& ex) {
    e->dry = false;
    return wd_set_error(WD_ERR_STATE, ex.what());
  }
  return WD_OK;
}

int wd_f32_last_launch_count(const wd_f32* e) { return e ? e->launches : 0; }
size_t wd_f32_workspace_bytes(const wd_f32* e) { return e ? e->arena_cap : 0; }

/* single operators of the fp32 path for the parity tests */
int wd_f32_op_conv3x3(const float* x_nhwc, const float* w_oihw, const float* bias, float* out_nhwc, int B, int H, int W, int Cin,
                      int Cout, int stride, int up, void* stream) {
  if (!x_nhwc || !w_oihw || !out_nhwc || (Cin & 3)) return wd_set_error(WD_ERR_INVALID, "wd_f32_op_conv3x3: invalid argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* wp = nullptr;
  const size_t n = static_cast<size_t>(Cout) * Cin * 9;
  if (cudaMalloc(&wp, n * sizeof(float)) != cudaSuccess) return wd_set_error(WD_ERR_CUDA, "wd_f32_op_conv3x3: cudaMalloc failed");
  f32_pack_conv_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(w_oihw, wp, Cout, Cin);
  GemmF32 g{};
  g.a1 = x_nhwc;
  g.C1 = Cin;
  g.taps = 9;
  g.Hin = H;
  g.Win = W;
  g.stride = stride;
  g.up = up;
  g.Hout = up ? 2 * H : (stride == 2 ? H / 2 : H);
  g.Wout = up ? 2 * W : (stride == 2 ? W / 2 : W);
  g.w = wp;
  g.bias = bias;
  g.out = out_nhwc;
  g.M = B * g.Hout * g.Wout;
  g.N = Cout;
  g.K = 9 * Cin;
  launch_gemm(g, s);
  const cudaError_t ce = cudaGetLastError();
  cudaStreamSynchronize(s);
  cudaFree(wp);
  if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  return WD_OK;
}

/* 3x3 pad-1 stride-1 convolution over cat([x1, x2], channel) on the split-TF32 tensor-core kernel (implicit GEMM: 4-D TMA boxes of the
 * split NHWC sources, f32_gemm_tc.cu); x2 may be NULL (C2 = 0).  fp32 NHWC in / out, weights [Cout, C1 + C2, 3, 3]. */
int wd_f32_op_conv3x3_tc(const float* x1, const float* x2, const float* w_oihw, const float* bias, float* out_nhwc, int B, int H, int W,
                         int C1, int C2, int Cout, void* stream) {
  if (!x1 || !w_oihw || !out_nhwc || (C2 > 0 && !x2) || !wd::f32tc_conv_ok(B, H, W, C1, C2, Cout))
    return wd_set_error(WD_ERR_INVALID, "wd_f32_op_conv3x3_tc: unsupported shape (B H W % 128, channels % 32, Cout % 160 or % 128)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Cin = C1 + C2;
  const size_t nw = static_cast<size_t>(Cout) * Cin * 9, n1 = static_cast<size_t>(B) * H * W * C1, n2 = static_cast<size_t>(B) * H * W * C2;
  float* buf = nullptr;
  if (cudaMalloc(&buf, (3 * nw + 2 * (n1 + n2)) * sizeof(float)) != cudaSuccess) return wd_set_error(WD_ERR_CUDA, "wd_f32_op_conv3x3_tc: cudaMalloc failed");
  float *wp = buf, *wh = wp + nw, *wl = wh + nw, *h1 = wl + nw, *l1 = h1 + n1, *h2 = l1 + n1, *l2 = h2 + n2;
  f32_pack_conv_kernel<<<static_cast<unsigned>((nw + 255) / 256), 256, 0, s>>>(w_oihw, wp, Cout, Cin);
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = wd::f32tc_split(wp, wh, wl, nw, s);
  if (ce == cudaSuccess) ce = wd::f32tc_split(x1, h1, l1, n1, s);
  if (ce == cudaSuccess && n2) ce = wd::f32tc_split(x2, h2, l2, n2, s);
  if (ce == cudaSuccess)
    ce = wd::f32tc_conv3x3(h1, l1, C1, n2 ? h2 : nullptr, n2 ? l2 : nullptr, C2, B, H, W, wh, wl, Cout, bias, nullptr, 0, nullptr, out_nhwc, 0, s);
  cudaStreamSynchronize(s);
  cudaFree(buf);
  if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  return WD_OK;
}

/* out[M,N] = A[M,K] W[N,K]^T + bias on the split-TF32 tensor-core kernel (f32_gemm_tc.cu); splits both operands itself */
int wd_f32_op_gemm_tc(const float* a, const float* w, const float* bias, float* out, int M, int N, int K, void* stream) {
  if (!a || !w || !out || !wd::f32tc_shape_ok(M, N, K)) return wd_set_error(WD_ERR_INVALID, "wd_f32_op_gemm_tc: need M % 128 == 0, N % 160 == 0 or N % 128 == 0, K % 32 == 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* buf = nullptr;
  const size_t nA = static_cast<size_t>(M) * K, nW = static_cast<size_t>(N) * K;
  const int splits = wd::f32tc_splits(K);
  const size_t nP = splits > 1 ? static_cast<size_t>(splits) * M * N : 0;
  if (cudaMalloc(&buf, (2 * (nA + nW) + nP) * sizeof(float)) != cudaSuccess) return wd_set_error(WD_ERR_CUDA, "wd_f32_op_gemm_tc: cudaMalloc failed");
  float *ah = buf, *al = buf + nA, *wh = buf + 2 * nA, *wl = buf + 2 * nA + nW;
  float* ws = nP ? buf + 2 * (nA + nW) : nullptr;
  cudaError_t ce = wd::f32tc_split(a, ah, al, nA, s);
  if (ce == cudaSuccess) ce = wd::f32tc_split(w, wh, wl, nW, s);
  if (ce == cudaSuccess) ce = wd::f32tc_gemm(ah, al, wh, wl, M, N, K, bias, nullptr, 0, 1, nullptr, out, 0, ws, s);
  const cudaError_t se = cudaStreamSynchronize(s);
  cudaFree(buf);
  if (ce == cudaSuccess) ce = se;
  if (ce != cudaSuccess) return wd_set_error(WD_ERR_CUDA, cudaGetErrorString(ce));
  return WD_OK;
}

int wd_f32_op_attention(const float* q, const float* k, const float* v, float* out, int B, int Sq, int Skv, int heads, int d,
                        float scale, void* stream) {
  if (!q || !k || !v || !out || d > 320) return wd_set_error(WD_ERR_INVALID, "wd_f32_op_attention: invalid argument");
  wd_f32 tmp;
  tmp.s = static_cast<cudaStream_t>(stream);
  const int C = heads * d;
  try {
    attention(&tmp, q, static_cast<size_t>(Sq) * C, C, k, v, static_cast<size_t>(Skv) * C, C, out, static_cast<size_t>(Sq) * C, C, B, Sq,
              Skv, heads, d, scale);
  } catch (const Fail& f) {
    return wd_set_error(f.code, f.msg.c_str());
  }
  return WD_OK;
}

}  // extern "C"
