// Register-level epilogue arithmetic shared by gemm_tc.cu and gemm_pair.cu: one call handles the 80 accumulator columns a
// warp has just drained from TMEM (one row per thread) and writes the 16-bit results into the staging sub-tiles.
//
// The flavour (residual add, GroupNorm partial sums, fp16 / bf16 storage) is a template parameter: with run-time flags
// inside the 10x-unrolled column loop the kernels grew to ~5000 SASS instructions, each 8-column chunk carried four
// BSSY/BSYNC reconvergence pairs, and ncu showed the epilogue warps stalled on instruction fetch (`no_inst`) more than
// on anything else (profiles/).  One `switch` per round picks a straight-line variant instead.
#pragma once
#include "common.cuh"

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
