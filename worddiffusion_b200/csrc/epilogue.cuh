// Register-level epilogue arithmetic shared by gemm_tc.cu and gemm_pair.cu: one call handles the 80 accumulator columns a
// warp has just drained from TMEM (one row per thread) and writes the 16-bit results into the staging sub-tiles.
//
// The flavour (residual add, GroupNorm partial sums, fp16 / bf16 storage) is a template parameter: with run-time flags
// inside the 10x-unrolled column loop the kernels grew to ~5000 SASS instructions, each 8-column chunk carried four
// BSSY/BSYNC reconvergence pairs, and ncu showed the epilogue warps stalled on instruction fetch (`no_inst`) more than
// on anything else (profiles/).  One `switch` per round picks a straight-line variant instead.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace wd {

// v: 80 fp32 accumulator values (bit patterns), wvr: the warp's 80 additive column terms (shared memory, broadcast reads),
// srow: this thread's row in staging sub-tile 0 of the round (sub-tile 1 is `sub_stride` bytes further; a sub-tile row is
// 40 columns = 5 chunks of 16 bytes).  RES: the staging chunk holds the residual (TMA-prefetched) and is added in place.
// GN: accumulate GroupNorm partial sums gs[2g] += x, gs[2g+1] += x^2 over the 8 groups of 10 columns (invalid rows add 0).
// LNS: also return {sum x, sum x^2} of the thread's 80 values in gs[0], gs[1] (LayerNorm row statistics for the consumer GEMM).
template <bool RES, bool GN, bool F16, bool LNS = false>
WD_DEVINL void epi_round80(const uint32_t* v, const float* wvr, uint8_t* srow, int sub_stride, bool valid, float* gs) {
  float ls = 0.f, lq = 0.f;
#pragma unroll
  for (int c = 0; c < 10; ++c) {  // 8 columns = one 16-byte staging chunk
    float f[8];
    const float4 b0 = *reinterpret_cast<const float4*>(wvr + c * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(wvr + c * 8 + 4);
    f[0] = __uint_as_float(v[c * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[c * 8 + 1]) + b0.y;
    f[2] = __uint_as_float(v[c * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[c * 8 + 3]) + b0.w;
    f[4] = __uint_as_float(v[c * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[c * 8 + 5]) + b1.y;
    f[6] = __uint_as_float(v[c * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[c * 8 + 7]) + b1.w;
    uint4* sp = reinterpret_cast<uint4*>(srow + (c / 5) * sub_stride + (c % 5) * 16);
    if constexpr (RES) {
      const uint4 r4 = *sp;
      const uint32_t ru[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = F16 ? unpack_f16x2(ru[j]) : unpack_bf16x2(ru[j]);
        f[2 * j] += t.x;
        f[2 * j + 1] += t.y;
      }
    }
    if constexpr (GN) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = (c * 8 + j) / 10;  // compile-time after unrolling
        const float x = valid ? f[j] : 0.f;
        gs[2 * g] += x;
        gs[2 * g + 1] = fmaf(x, x, gs[2 * g + 1]);
      }
    }
    if constexpr (LNS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ls += f[j];
        lq = fmaf(f[j], f[j], lq);
      }
    }
    if constexpr (F16)
      *sp = make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
    else
      *sp = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
  if constexpr (LNS) {
    gs[0] = ls;
    gs[1] = lq;
  }
}

// ---- GroupNorm applied by the producer (GemmArgs::gn_apply, gemm_pair.cu) ----
// Round that stays in registers: f = acc + additive terms, GroupNorm partial sums as in epi_round80<.., GN = true, ..>, the 80 values
// packed as fp16 pairs into h[40].
WD_DEVINL void epi_round80_hold(const uint32_t* v, const float* wvr, uint32_t* h, bool valid, float* gs) {
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    float f[8];
    const float4 b0 = *reinterpret_cast<const float4*>(wvr + c * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(wvr + c * 8 + 4);
    f[0] = __uint_as_float(v[c * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[c * 8 + 1]) + b0.y;
    f[2] = __uint_as_float(v[c * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[c * 8 + 3]) + b0.w;
    f[4] = __uint_as_float(v[c * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[c * 8 + 5]) + b1.y;
    f[6] = __uint_as_float(v[c * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[c * 8 + 7]) + b1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c * 8 + j) / 10;
      const float x = valid ? f[j] : 0.f;
      gs[2 * g] += x;
      gs[2 * g + 1] = fmaf(x, x, gs[2 * g + 1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) h[c * 4 + j] = pack_f16x2(f[2 * j], f[2 * j + 1]);
  }
}
// 8 columns of y = silu(h * sc + sh), bf16, from 8 fp16 values; sc8 / sh8: shared memory, the pre-halved scale / shift of the 8
// columns (sc = rstd_g gamma / 2, sh = (beta - mean_g rstd_g gamma) / 2).  SiLU as y/2 + y/2 tanh(y/2) with one packed
// tanh.approx.f16x2 per two elements: the arithmetic of ops.cu gn_vec8<true, true>, bit for bit.
WD_DEVINL uint4 gn_apply_vec8(const uint4 hv, const float* sc8, const float* sh8) {
  const uint32_t u[4] = {hv.x, hv.y, hv.z, hv.w};
  const float4 s0 = *reinterpret_cast<const float4*>(sc8), s1 = *reinterpret_cast<const float4*>(sc8 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(sh8), b1 = *reinterpret_cast<const float4*>(sh8 + 4);
  const float2 sc2[4] = {make_float2(s0.x, s0.y), make_float2(s0.z, s0.w), make_float2(s1.x, s1.y), make_float2(s1.z, s1.w)};
  const float2 sh2[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 y = __ffma2_rn(unpack_f16x2(u[j]), sc2[j], sh2[j]);  // y / 2
    uint32_t h16 = pack_f16x2(y.x, y.y), t16;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t16) : "r"(h16));
    y = __ffma2_rn(y, unpack_f16x2(t16), y);
    o[j] = pack_bf16x2(y.x, y.y);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// LayerNorm-consuming flavour: y = rstd * acc + (b'[c] - rstd*mu * s[c]); wvr = b', wsr = s (shared memory); bf16 out
WD_DEVINL void epi_round80_lnc(const uint32_t* v, const float* wvr, const float* wsr, float rstd, float rstd_mu, uint8_t* srow,
                               int sub_stride) {
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    float f[8];
    const float4 b0 = *reinterpret_cast<const float4*>(wvr + c * 8), b1 = *reinterpret_cast<const float4*>(wvr + c * 8 + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(wsr + c * 8), s1 = *reinterpret_cast<const float4*>(wsr + c * 8 + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(rstd, __uint_as_float(v[c * 8 + j]), fmaf(-rstd_mu, ss[j], bb[j]));
    uint4* sp = reinterpret_cast<uint4*>(srow + (c / 5) * sub_stride + (c % 5) * 16);
    *sp = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

// run-time dispatch to the straight-line variants (flags are warp-uniform)
WD_DEVINL void epi_round80_dispatch(bool res, bool gn, bool f16, bool lns, const uint32_t* v, const float* wvr, uint8_t* srow,
                                    int sub_stride, bool valid, float* gs) {
  if (lns) {  // row statistics for a following LayerNorm: only produced for the fp16 token stream, never together with GN
    if (res) epi_round80<true, false, true, true>(v, wvr, srow, sub_stride, valid, gs);
    else epi_round80<false, false, true, true>(v, wvr, srow, sub_stride, valid, gs);
    return;
  }
  const int sel = (res ? 4 : 0) | (gn ? 2 : 0) | (f16 ? 1 : 0);
  switch (sel) {
    case 0: epi_round80<false, false, false>(v, wvr, srow, sub_stride, valid, gs); break;
    case 1: epi_round80<false, false, true>(v, wvr, srow, sub_stride, valid, gs); break;
    case 2: epi_round80<false, true, false>(v, wvr, srow, sub_stride, valid, gs); break;
    case 3: epi_round80<false, true, true>(v, wvr, srow, sub_stride, valid, gs); break;
    case 4: epi_round80<true, false, false>(v, wvr, srow, sub_stride, valid, gs); break;
    case 5: epi_round80<true, false, true>(v, wvr, srow, sub_stride, valid, gs); break;
    case 6: epi_round80<true, true, false>(v, wvr, srow, sub_stride, valid, gs); break;
    default: epi_round80<true, true, true>(v, wvr, srow, sub_stride, valid, gs); break;
  }
}

// Rarely used flavours (the three time-embedding GEMMs with M = batch rows): SiLU, fp32 output written straight to global
// memory, per-thread row-bias rows, residual read from global memory.  Run-time flags, not performance relevant.
WD_DEVINL void epi_round80_generic(const uint32_t* v, const float* wvr, const float* rb_cols /*or null*/, int act_silu,
                                   const __nv_bfloat16* res_row /*or null*/, bool res_f16, bool use_stg, uint8_t* srow,
                                   int sub_stride, bool out_f32, bool out_f16, void* out_row /*global row + column base*/,
                                   bool valid) {
#pragma unroll  // (a rolled loop would index v[] dynamically and push the whole accumulator array into local memory)
  for (int c = 0; c < 10; ++c) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[c * 8 + j]) + wvr[c * 8 + j];
    if (rb_cols) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += __ldg(rb_cols + c * 8 + j);
    }
    uint4* sp = reinterpret_cast<uint4*>(srow + (c / 5) * sub_stride + (c % 5) * 16);
    if (res_row && valid) {
      const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(res_row + c * 8));
      const uint32_t ru[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_16x2(ru[j], res_f16);
        f[2 * j] += t.x;
        f[2 * j + 1] += t.y;
      }
    }
    if (act_silu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
    }
    const uint4 o4 = make_uint4(pack_16x2(f[0], f[1], out_f16), pack_16x2(f[2], f[3], out_f16), pack_16x2(f[4], f[5], out_f16),
                                pack_16x2(f[6], f[7], out_f16));
    if (use_stg) {
      *sp = o4;
    } else if (valid) {
      if (out_f32) {
        float4* op = reinterpret_cast<float4*>(static_cast<float*>(out_row) + c * 8);
        op[0] = make_float4(f[0], f[1], f[2], f[3]);
        op[1] = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out_row) + c * 8) = o4;
      }
    }
  }
}

// GEGLU (unet.py:127-129): out = (value + bv) * gelu(gate + bg) for 40 output columns; v[0..39] values, v[40..79] gates,
// wv_val / wv_gate their biases; writes one staging sub-tile row (5 chunks), bf16.
// LNC: the A operand was the un-normalised tensor: value / gate = rstd * acc + (b' - rstd*mu * s) first (ws_* = s vectors).
template <bool LNC>
WD_DEVINL void epi_geglu40(const uint32_t* v, const float* wv_val, const float* wv_gate, uint8_t* srow, const float* ws_val = nullptr,
                           const float* ws_gate = nullptr, float rstd = 1.f, float rstd_mu = 0.f) {
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    float f[8];
    const float4 bv0 = *reinterpret_cast<const float4*>(wv_val + c * 8), bv1 = *reinterpret_cast<const float4*>(wv_val + c * 8 + 4);
    const float4 bg0 = *reinterpret_cast<const float4*>(wv_gate + c * 8), bg1 = *reinterpret_cast<const float4*>(wv_gate + c * 8 + 4);
    float bv[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
    float bg[8] = {bg0.x, bg0.y, bg0.z, bg0.w, bg1.x, bg1.y, bg1.z, bg1.w};
    float sc = 1.f;
    if constexpr (LNC) {
      const float4 sv0 = *reinterpret_cast<const float4*>(ws_val + c * 8), sv1 = *reinterpret_cast<const float4*>(ws_val + c * 8 + 4);
      const float4 sg0 = *reinterpret_cast<const float4*>(ws_gate + c * 8), sg1 = *reinterpret_cast<const float4*>(ws_gate + c * 8 + 4);
      const float sv[8] = {sv0.x, sv0.y, sv0.z, sv0.w, sv1.x, sv1.y, sv1.z, sv1.w};
      const float sg[8] = {sg0.x, sg0.y, sg0.z, sg0.w, sg1.x, sg1.y, sg1.z, sg1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bv[j] = fmaf(-rstd_mu, sv[j], bv[j]);
        bg[j] = fmaf(-rstd_mu, sg[j], bg[j]);
      }
      sc = rstd;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      f[j] = fmaf(sc, __uint_as_float(v[c * 8 + j]), bv[j]) * gelu_fast_f(fmaf(sc, __uint_as_float(v[40 + c * 8 + j]), bg[j]));
    *reinterpret_cast<uint4*>(srow + c * 16) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

// Fused short-context cross-attention on the to_q accumulator (CrossAttention.forward, unet.py:185-207, with the <= 16 key /
// value rows of the sample's character context).  v = the thread's 80 accumulator columns = q of one (row, head) before the
// additive terms.  The warp (32 rows x one head) writes q as bf16 into its own rows of the staging tile, reads it back as
// mma.sync A fragments (ldmatrix), forms S = q K^T (16 keys, padded rows are zero and masked), the softmax in the accumulator
// layout (a row lives in one quad), and O = P V, which overwrites the q rows: the staging tile then holds the output and is
// stored by TMA like any other epilogue.  sK / sV: bf16 [16][ATT_KP] rows of (sample, head) in shared memory.
// (A first version kept q in registers and read K / V with broadcast LDS.128: 400 per warp and tile; the shared-memory
// return path (128 B / clk / SM) made the launch 2.5x slower than the two kernels it replaces.)
constexpr int ATT_KP = 88;  // K / V row pitch in elements (176 B: conflict-free fragment loads)

WD_DEVINL void epi_mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// swarp: row 0 of this warp in staging sub-tile 0 (rows are 80 bytes = 40 columns; sub-tile 1 is `sub_stride` bytes further)
template <bool LNC>
WD_DEVINL void epi_ctx_attn80(const uint32_t* v, const float* wvr, const float* wsr, float rstd, float rstd_mu, float sl2,
                              const __nv_bfloat16* sK, const __nv_bfloat16* sV, int L, uint8_t* swarp, int sub_stride, int lane) {
  // ---- q (fp32, exact additive terms) -> bf16 -> own staging row ----
  uint8_t* const srow = swarp + lane * (GEMM_SUB_N * 2);
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    float f[8];
    const float4 b0 = *reinterpret_cast<const float4*>(wvr + c * 8), b1 = *reinterpret_cast<const float4*>(wvr + c * 8 + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    if constexpr (LNC) {
      const float4 s0 = *reinterpret_cast<const float4*>(wsr + c * 8), s1 = *reinterpret_cast<const float4*>(wsr + c * 8 + 4);
      const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(rstd, __uint_as_float(v[c * 8 + j]), fmaf(-rstd_mu, ss[j], bb[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[c * 8 + j]) + bb[j];
    }
    *reinterpret_cast<uint4*>(srow + (c / 5) * sub_stride + (c % 5) * 16) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
  __syncwarp();
  const int g = lane >> 2, tq = lane & 3;
  // ---- K fragments (B operand of S = q K^T): both key octets, 5 k-steps ----
  uint32_t kb[2][5][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const __nv_bfloat16* kr = sK + (nt * 8 + g) * ATT_KP + 2 * tq;
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      kb[nt][ks][0] = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
      kb[nt][ks][1] = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    uint8_t* const mrow = swarp + (mt * 16) * (GEMM_SUB_N * 2);
    // ---- q fragments of 16 rows (A operand): 8-column blocks never straddle the two 40-column sub-tiles ----
    uint32_t qf[5][4];
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      const int c8 = ks * 2 + (lane >> 4);
      const uint8_t* p = mrow + (lane & 15) * (GEMM_SUB_N * 2) + (c8 / 5) * sub_stride + (c8 % 5) * 16;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                   : "=r"(qf[ks][0]), "=r"(qf[ks][1]), "=r"(qf[ks][2]), "=r"(qf[ks][3])
                   : "r"(smem_u32(p)));
    }
    float sc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[nt][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) epi_mma_bf16_16816(sc[nt], qf[ks], kb[nt][ks][0], kb[nt][ks][1]);
    }
    // ---- softmax over the keys (rows g and g + 8 of the m-tile; a row lives in one quad) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int key = nt * 8 + 2 * tq;
      if (key >= L) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
      if (key + 1 >= L) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pf[4];  // P as one A fragment (k = 16 keys)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const float p0 = exp2f((sc[nt][0] - mx0) * sl2), p1 = exp2f((sc[nt][1] - mx0) * sl2);
      const float p2 = exp2f((sc[nt][2] - mx1) * sl2), p3 = exp2f((sc[nt][3] - mx1) * sl2);
      l0 += p0 + p1;
      l1 += p2 + p3;
      pf[nt * 2] = pack_bf16x2(p0, p1);
      pf[nt * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    // ---- O = P V over the q rows of this m-tile (their fragments are in registers) ----
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < 10; ++dt) {
      uint32_t b0, b1;
      asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                   : "=r"(b0), "=r"(b1)
                   : "r"(smem_u32(sV + (lane & 15) * ATT_KP + dt * 8)));
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      epi_mma_bf16_16816(o, pf, b0, b1);
      uint8_t* const op = mrow + (dt / 5) * sub_stride + (dt % 5) * 16 + tq * 4;
      *reinterpret_cast<uint32_t*>(op + g * (GEMM_SUB_N * 2)) = pack_bf16x2(o[0] * i0, o[1] * i0);
      *reinterpret_cast<uint32_t*>(op + (g + 8) * (GEMM_SUB_N * 2)) = pack_bf16x2(o[2] * i1, o[3] * i1);
    }
  }
}

// lane L (< 16) ends with the sum over the warp's 32 lanes of v[L] (v is destroyed): 16 + 15 shuffles
WD_DEVINL float warp_transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

}  // namespace wd
