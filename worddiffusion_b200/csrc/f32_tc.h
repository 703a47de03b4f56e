// Internal interface of the split-TF32 tensor-core GEMM (f32_gemm_tc.cu) used by the fp32 path (f32_path.cu).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace wd {
bool f32tc_enabled();                        // env WD_F32_TC (default on; 0 keeps every contraction on the FFMA kernel)
bool f32tc_shape_ok(int M, int N, int K);    // M % 128 == 0, N % 160 == 0, K % 32 == 0
// a[n] -> hi[n] = tf32_rn(a), lo[n] = a - hi   (n % 4 == 0)
cudaError_t f32tc_split(const float* a, float* hi, float* lo, size_t n, cudaStream_t s);
// [M, C1] ++ [M, C2] along the channels, split (C1, C2 % 4 == 0)
cudaError_t f32tc_split_concat(const float* a1, const float* a2, int C1, int C2, size_t M, float* hi, float* lo, cudaStream_t s);
// 3x3 pad-1 patch matrix [B Hout Wout, 9 (C1 + C2)] of NHWC source(s), already split (stride 1|2, or nearest-2x first)
cudaError_t f32tc_im2col_split(const float* a1, const float* a2, int C1, int C2, int B, int Hin, int Win, int stride, int up, float* hi,
                               float* lo, cudaStream_t s);
// number of K chunks whose partial tiles are summed in fp32 (1: K is short enough for one TMEM accumulation)
int f32tc_splits(int K);
// out[M,N] = act(A W^T + bias + rowbias[m / rows_per_sample] + residual), operands pre-split, all fp32 row-major.
// partial_ws: fp32 [f32tc_splits(K), M, N] workspace for the two-level accumulation, or null (single accumulation whatever K)
cudaError_t f32tc_gemm(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, int M, int N, int K, const float* bias,
                       const float* rowbias, int rb_ld, int rows_per_sample, const float* residual, float* out, int act_silu,
                       float* partial_ws, cudaStream_t s);
// GEGLU.proj with the gating in the GEMM's epilogue (unet.py:122-130): out[M, N / 2] = (A Wv^T + bv) * gelu_erf(A Wg^T + bg).
// w_hi / w_lo / bias_perm are in f32tc_geglu_permute order (160-row tiles of 80 values + their 80 gates; N % 320 == 0; a bias is a
// [N, 1] matrix for the permutation).  out_lo != null: the result is written as its TF32 split (out = hi, out_lo = lo).
cudaError_t f32tc_geglu_permute(const float* w, float* dst, int N, int K, cudaStream_t s);
cudaError_t f32tc_gemm_geglu(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, int M, int N, int K,
                             const float* bias_perm, float* out, float* out_lo, cudaStream_t s);
// "nearest 2x upsample, then conv3x3" (unet.py:497-499) in sub-pixel form: four 2 x 2 convolutions of the H x W input, one per output
// phase, as ONE implicit GEMM (9/4 fewer MACs, no upsampled tensor, no patch matrix).  w_hi / w_lo: the TF32 split of
// f32tc_upconv_fold(packed [Cout][9][C]) = [4 Cout][4 C]; out: [B, 2H, 2W, Cout].
cudaError_t f32tc_upconv_fold(const float* w_packed, float* dst, int Cout, int C, cudaStream_t s);
bool f32tc_upconv_ok(int B, int H, int W, int C, int Cout);
cudaError_t f32tc_upconv(const float* a_hi, const float* a_lo, int C, int B, int H, int W, const float* w_hi, const float* w_lo, int Cout,
                         const float* bias, float* out, cudaStream_t s);
// implicit 3x3 pad-1 stride-1 convolution over the channel concatenation of up to two split NHWC sources [B, H, W, C1 | C2]
// (C % 32 == 0; 128 % W == 0 and HW % 128 == 0, or 128 % HW == 0); weights [N, 9 (C1 + C2)] split, k = tap (C1 + C2) + c
bool f32tc_conv_ok(int B, int H, int W, int C1, int C2, int N);
cudaError_t f32tc_conv3x3(const float* a1_hi, const float* a1_lo, int C1, const float* a2_hi, const float* a2_lo, int C2, int B, int H, int W,
                          const float* w_hi, const float* w_lo, int N, const float* bias, const float* rowbias, int rb_ld,
                          const float* residual, float* out, int act_silu, cudaStream_t s);
}  // namespace wd
