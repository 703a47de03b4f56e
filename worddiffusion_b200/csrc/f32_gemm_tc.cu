// fp32-class GEMM on the 5th-gen tensor cores: C[M,N] = A[M,K] W[N,K]^T with every fp32 operand split into two TF32 terms,
//   a = a_hi + a_lo,  a_hi = tf32_rn(a) (low 13 mantissa bits zero),  a_lo = a - a_hi (exact in fp32),
//   C = A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T        (the A_lo W_lo^T term is ~2^-22 and dropped)
// as three tcgen05.mma.kind::tf32 per K step into one fp32 TMEM accumulator.  Per product the error is ~2^-21 (the hardware's own
// tf32 conversion of a_lo / w_lo), i.e. fp32-class: DESIGN.md section 10 item 4 has the error budget that rules out a 2-term bf16
// split for north_star's 1e-4.  The operands are PRE-SPLIT in memory (f32tc_split_kernel / f32tc_im2col_split_kernel write a_hi and
// a_lo; weights are split once per load), so the result does not depend on how kind::tf32 converts its inputs.
//
// Kernel: one CTA = one 128 x 160 output tile, 192 threads: warp 0 = TMA producer (four SWIZZLE_128B boxes per 32-wide K block:
// A_hi, A_lo 128 x 32, W_hi, W_lo 160 x 32 fp32 = 72 KB per stage, 3 stages), warp 1 = single-thread MMA issuer (12 MMAs
// 128 x 160 x 8 per K block), warps 2-5 = epilogue (one TMEM lane quarter each; bias / per-sample row bias / residual / SiLU, fp32
// stores of 64-byte row pieces).
//
// STATUS: opt-in (env WD_F32_TC=1 routes the fp32 mode's Linear / 1x1 / 3x3 contractions here); the default fp32 mode is the FFMA
// kernel of f32_path.cu.  First run on a B200 (profiles/r03h_split_tf32_gemm_first_run.log, tests/test_gpu_zfp32.py::
// test_f32_tc_gemm_operator under WD_F32_TC_TEST=1): correct; max-rel error vs fp64 2.4e-7 at K = 32, 1.6e-6 at K = 320, 1.3e-5 at
// K = 2880 -- the tensor core's fp32 accumulation truncates, so the error grows with K: the long-K convolutions need a two-level
// accumulation before this is fp32-class.  That second level was added AFTER the run (not yet executed on a GPU): K > 512 is cut into
// chunks of 320 (blockIdx.z), each chunk's raw tile goes to a workspace and f32tc_reduce_kernel adds the chunks in fp32, in order
// (deterministic), then applies the epilogue.  K <= 512 takes the code path that was run.  Not timed yet.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "f32_tc.h"

namespace wd {
namespace {

constexpr int TC_BM = 128, TC_BN = 160, TC_BK = 32;  // 32 fp32 = 128 B = one SWIZZLE_128B row
// The kernel is also built with 128-column tiles for widths that are multiples of 128 but not of 160 (the VAE decoder's 512 / 256 /
// 128 channels); tile width for a given N:
constexpr int tc_bn_for(int N) { return N % TC_BN == 0 ? TC_BN : (N % 128 == 0 ? 128 : 0); }
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
constexpr int TC_W_BYTES = TC_BN * TC_BK * 4;  // 20 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_W_BYTES;  // 72 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 256;  // + barriers / TMEM slot
constexpr uint32_t TC_TMEM_COLS = 512;  // two 160-column chunk accumulators at columns 0 and 256
// The tensor core's fp32 accumulation truncates: the error of one TMEM accumulation grows with K (measured 1.6e-6 at K = 320,
// 1.3e-5 at K = 2880).  Longer contractions are therefore cut into chunks of TC_SPLIT_KB K blocks (K = 320) whose partial tiles a
// second kernel adds in fp32 (two-level accumulation).
constexpr int TC_SPLIT_KB = 10;
constexpr int TC_SPLIT_MIN_K = 512;

// Instruction descriptor, kind::tf32: A, B = TF32 (format code 2), K-major both, D = fp32, shape M x N (K = 8)
__host__ __device__ constexpr uint32_t make_idesc_tf32_f32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

WD_DEVINL void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// a -> tf32_rn(a) (low 13 mantissa bits zero); a - tf32_rn(a) is exact in fp32
WD_DEVINL float tf32_rn(float a) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  return __uint_as_float(r & 0xFFFFE000u);
}

struct TcArgs {
  int M, N, K;
  // implicit 3x3 pad-1 stride-1 convolution: A rows are 4-D TMA boxes (c, w, h, n) of the split NHWC source(s), shifted per filter
  // tap (out-of-bounds zero fill = the padding); k = tap (C1 + C2) + c over the channel concatenation of up to two sources
  int conv;
  int HW, W;     // output (= input) pixels per image, row length
  int cb1, cb2;  // 32-channel K blocks of source 1 / source 2
  const float* bias;
  const float* rowbias;
  int rb_ld;
  int rows_per_sample;
  const float* residual;
  float* out;
  int act_silu;
  // two-level accumulation: the tensor core's fp32 accumulation truncates (error grows with K), so K is walked in chunks of
  // kb_per_chunk K blocks; each chunk accumulates in one of two TMEM buffers and the epilogue warps add the finished chunk to fp32
  // REGISTER accumulators (exact fp32 adds, in order: deterministic) while the next chunk's MMAs run in the other buffer
  int kb_per_chunk;
  // GEGLU epilogue (unet.py:122-130): the weight rows were permuted so that a 160-column tile holds 80 value columns and their 80
  // gates (f32tc_geglu_permute); out[m, n0 / 2 + j] = (acc[j] + b[j]) * gelu_erf(acc[80 + j] + b[80 + j]), row stride N / 2.
  // out_lo != null: the result leaves as its TF32 split (out = hi, out_lo = lo) for the Linear that consumes it.
  int geglu;
  float* out_lo;
  // Sub-pixel form of "nearest 2x upsample, then conv3x3" (unet.py:497-499): the four output phases (a, b) are N tiles of a
  // [4 up_cout, 4 C] weight (phase = n0 / up_cout); tap t of phase (a, b) reads input offset ((t >> 1) - 1 + a, (t & 1) - 1 + b);
  // row m = (image, y, x) of the H x W input grid is written to pixel (image, 2 y + a, 2 x + b) of the [B, 2H, 2W, up_cout] output.
  int up;
  int up_cout;
  int H;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) f32tc_gemm_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                                                            const __grid_constant__ CUtensorMap mapA2h, const __grid_constant__ CUtensorMap mapA2l,
                                                            const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
                                                            const TcArgs args) {
  constexpr int W_B = BN * TC_BK * 4;                // one weight box
  constexpr int STAGE_B = 2 * TC_A_BYTES + 2 * W_B;  // A_hi, A_lo, W_hi, W_lo
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment (no static shared memory in this kernel)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_B);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* acc_full = empty_bar + TC_STAGES;   // [2]: chunk accumulator complete (MMA -> epilogue)
  uint64_t* acc_empty = acc_full + 2;           // [2]: chunk accumulator drained (4 epilogue warps -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: CTA b walks tiles b, b + gridDim.x, ... (n-tile fastest, so concurrent CTAs share the A rows in L2); the chunk
  // accumulators alternate between the two TMEM buffers ACROSS tiles, so the epilogue of one tile (registers -> global) runs under
  // the MMAs of the next.  (The first build launched one CTA per tile: ~16 us of prologue / pipeline fill / epilogue per tile
  // against 9 us of MMA time at K = 320, profiles/R2q_fp32_launches.txt.)
  const int n_tiles = args.N / BN;
  const int total_tiles = n_tiles * (args.M / TC_BM);
  const int nkb = args.K / TC_BK;
  const int kpc = args.kb_per_chunk;
  const int nchunks = (nkb + kpc - 1) / kpc;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapWh);
    tma_prefetch_desc(&mapWl);
    if (args.cb2) {
      tma_prefetch_desc(&mapA2h);
      tma_prefetch_desc(&mapA2l);
    }
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TC_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int cbt = args.cb1 + args.cb2;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
      int img = 0, oh0 = 0, ow0 = 0;
      if (args.conv) {
        img = m0 / args.HW;
        oh0 = (m0 % args.HW) / args.W;
        ow0 = m0 % args.W;  // non-zero only for rows longer than a tile (W a multiple of 128)
      }
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], STAGE_B);
        uint8_t* s = smem + stage * STAGE_B;
        const int kc = kb * TC_BK;
        if (args.conv) {
          const int tap = kb / cbt, cb = kb - tap * cbt;
          int dy = tap / 3 - 1, dx = tap % 3 - 1;
          if (args.up) {
            const int ph = n0 / args.up_cout;
            dy = (tap >> 1) - 1 + (ph >> 1);
            dx = (tap & 1) - 1 + (ph & 1);
          }
          const bool second = cb >= args.cb1;
          const int c0 = (second ? cb - args.cb1 : cb) * TC_BK;
          tma_load_4d(s, second ? &mapA2h : &mapAh, &full_bar[stage], c0, ow0 + dx, oh0 + dy, img);
          tma_load_4d(s + TC_A_BYTES, second ? &mapA2l : &mapAl, &full_bar[stage], c0, ow0 + dx, oh0 + dy, img);
        } else {
          tma_load_2d(s, &mapAh, &full_bar[stage], kc, m0);
          tma_load_2d(s + TC_A_BYTES, &mapAl, &full_bar[stage], kc, m0);
        }
        tma_load_2d(s + 2 * TC_A_BYTES, &mapWh, &full_bar[stage], kc, n0);
        tma_load_2d(s + 2 * TC_A_BYTES + W_B, &mapWl, &full_bar[stage], kc, n0);
        if (++stage == TC_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32_f32(TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t gch = 0;  // chunks issued so far by this CTA (over all its tiles)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
      for (int ch = 0; ch < nchunks; ++ch, ++gch) {
        const int buf = gch & 1;
        if (gch >= 2) {  // the epilogue warps have added the chunk that used this buffer last to their registers
          mbar_wait(&acc_empty[buf], ((gch >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t d = tmem_base + buf * 256;
        const int kend = min(nkb, (ch + 1) * kpc);
        for (int kb = ch * kpc; kb < kend; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t s = smem_u32(smem + stage * STAGE_B);
          const uint64_t ah = make_smem_desc_sw128(s), al = make_smem_desc_sw128(s + TC_A_BYTES);
          const uint64_t wh = make_smem_desc_sw128(s + 2 * TC_A_BYTES), wl = make_smem_desc_sw128(s + 2 * TC_A_BYTES + W_B);
          const bool first = kb == ch * kpc;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            // advance 8 tf32 = 32 B along K inside the swizzled row: +2 in the (addr >> 4) field; small terms first
            umma_tf32_ss(d, al + 2 * k, wh + 2 * k, idesc, (!first || k != 0) ? 1u : 0u);
            umma_tf32_ss(d, ah + 2 * k, wl + 2 * k, idesc, 1u);
            umma_tf32_ss(d, ah + 2 * k, wh + 2 * k, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);  // frees the stage when these MMAs retire
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&acc_full[buf]);  // chunk accumulator complete
      }
    }
  } else {
    // epilogue: warp w may only read TMEM lanes [32 (w % 4), +32); thread = one output row
    const int q = warp & 3;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t gch = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
      const int m = m0 + q * 32 + lane;
      float acc[BN];  // second accumulation level, fp32 registers (every index below is a compile-time constant)
#pragma unroll 1
      for (int ch = 0; ch < nchunks; ++ch, ++gch) {
        const int buf = gch & 1;
        mbar_wait(&acc_full[buf], (gch >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + buf * 256 + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c * 16 + j] = ch == 0 ? __uint_as_float(v[j]) : acc[c * 16 + j] + __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      if constexpr (BN == TC_BN) {
        if (args.geglu) {
          if (m < args.M) {
            const size_t o0 = static_cast<size_t>(m) * (args.N / 2) + n0 / 2;
            const float4* bv = reinterpret_cast<const float4*>(args.bias + n0);
            const float4* bg = reinterpret_cast<const float4*>(args.bias + n0 + 80);
#pragma unroll
            for (int c4 = 0; c4 < 20; ++c4) {
              const float4 b1 = __ldg(bv + c4), b2 = __ldg(bg + c4);
              const float val[4] = {acc[c4 * 4] + b1.x, acc[c4 * 4 + 1] + b1.y, acc[c4 * 4 + 2] + b1.z, acc[c4 * 4 + 3] + b1.w};
              const float gt[4] = {acc[80 + c4 * 4] + b2.x, acc[80 + c4 * 4 + 1] + b2.y, acc[80 + c4 * 4 + 2] + b2.z, acc[80 + c4 * 4 + 3] + b2.w};
              float o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = val[j] * (0.5f * gt[j] * (1.0f + erff(gt[j] * 0.70710678118654752440f)));
              if (args.out_lo) {
                const float4 h = make_float4(tf32_rn(o[0]), tf32_rn(o[1]), tf32_rn(o[2]), tf32_rn(o[3]));
                *reinterpret_cast<float4*>(args.out + o0 + c4 * 4) = h;
                *reinterpret_cast<float4*>(args.out_lo + o0 + c4 * 4) = make_float4(o[0] - h.x, o[1] - h.y, o[2] - h.z, o[3] - h.w);
              } else {
                *reinterpret_cast<float4*>(args.out + o0 + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          continue;
        }
      }
      if (args.up) {
        if (m < args.M) {
          const int ph = n0 / args.up_cout, nc = n0 - ph * args.up_cout;
          const int img = m / args.HW, rem = m - img * args.HW, y = rem / args.W, x = rem - y * args.W;
          const size_t opix = (static_cast<size_t>(img) * 2 * args.H + 2 * y + (ph >> 1)) * (2 * args.W) + 2 * x + (ph & 1);
          float* orow_u = args.out + opix * args.up_cout + nc;
          const float4* bias_u = args.bias ? reinterpret_cast<const float4*>(args.bias + nc) : nullptr;
#pragma unroll
          for (int c4 = 0; c4 < BN / 4; ++c4) {
            float4 o = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
            if (bias_u) {
              const float4 b = __ldg(bias_u + c4);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            *reinterpret_cast<float4*>(orow_u + c4 * 4) = o;
          }
        }
        continue;
      }
      const int sample = args.rowbias ? m / args.rows_per_sample : 0;
      float* orow = args.out + static_cast<size_t>(m) * args.N + n0;
      const float4* rrow = args.residual ? reinterpret_cast<const float4*>(args.residual + static_cast<size_t>(m) * args.N + n0) : nullptr;
      const float4* rbrow = args.rowbias ? reinterpret_cast<const float4*>(args.rowbias + static_cast<size_t>(sample) * args.rb_ld + n0) : nullptr;
      const float4* bias = args.bias ? reinterpret_cast<const float4*>(args.bias + n0) : nullptr;
      const int act_silu = args.act_silu;
      if (m < args.M) {
#pragma unroll
        for (int c4 = 0; c4 < BN / 4; ++c4) {  // 16-byte loads of bias / row bias / residual (N, n0, rb_ld are multiples of 4)
          float4 o = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
          if (bias) {
            const float4 b = __ldg(bias + c4);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (rbrow) {
            const float4 b = __ldg(rbrow + c4);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (rrow) {
            const float4 b = __ldg(rrow + c4);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (act_silu) {
            o.x = o.x / (1.0f + expf(-o.x)); o.y = o.y / (1.0f + expf(-o.y));
            o.z = o.z / (1.0f + expf(-o.z)); o.w = o.w / (1.0f + expf(-o.w));
          }
          *reinterpret_cast<float4*>(orow + c4 * 4) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TC_TMEM_COLS>(tmem_base);
}


__global__ void f32tc_split_kernel(const float* __restrict__ a, float* __restrict__ hi, float* __restrict__ lo, size_t n4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(a)[i];
  float4 h, l;
  h.x = tf32_rn(v.x); l.x = v.x - h.x;
  h.y = tf32_rn(v.y); l.y = v.y - h.y;
  h.z = tf32_rn(v.z); l.z = v.z - h.z;
  h.w = tf32_rn(v.w); l.w = v.w - h.w;
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// channel concatenation of two row-major sources [M, C1], [M, C2] -> split [M, C1 + C2] (a 1x1 convolution over torch.cat([h, skip]))
__global__ void f32tc_split_concat_kernel(const float* __restrict__ a1, const float* __restrict__ a2, int C1, int C2, float* __restrict__ hi,
                                          float* __restrict__ lo, size_t total4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int K4 = (C1 + C2) / 4;
  const size_t m = i / K4;
  const int c = static_cast<int>(i - m * K4) * 4;
  const float4 v = c < C1 ? *reinterpret_cast<const float4*>(a1 + m * C1 + c) : *reinterpret_cast<const float4*>(a2 + m * C2 + (c - C1));
  float4 h, l;
  h.x = tf32_rn(v.x); l.x = v.x - h.x;
  h.y = tf32_rn(v.y); l.y = v.y - h.y;
  h.z = tf32_rn(v.z); l.z = v.z - h.z;
  h.w = tf32_rn(v.w); l.w = v.w - h.w;
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// 3x3 pad-1 patch matrix of an NHWC fp32 tensor (channel concat of up to two sources), split: hi, lo [M, 9 (C1 + C2)], k = tap (C1 + C2) + c
__global__ void f32tc_im2col_split_kernel(const float* __restrict__ a1, const float* __restrict__ a2, int C1, int C2, int Hin, int Win,
                                          int Hout, int Wout, int stride, int up, float* __restrict__ hi, float* __restrict__ lo,
                                          size_t total4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int Cin = C1 + C2, K4 = 9 * Cin / 4;
  const size_t m = i / K4;
  const int k = static_cast<int>(i - m * K4) * 4;
  const int tap = k / Cin, c = k - tap * Cin;
  const int hw = Hout * Wout;
  const int b = static_cast<int>(m / hw), rem = static_cast<int>(m - static_cast<size_t>(b) * hw);
  const int oh = rem / Wout, ow = rem - oh * Wout;
  int ih = oh * stride + tap / 3 - 1, iw = ow * stride + tap % 3 - 1;
  const int Hs = up ? 2 * Hin : Hin, Ws = up ? 2 * Win : Win;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ih >= 0 && ih < Hs && iw >= 0 && iw < Ws) {
    if (up) {
      ih >>= 1;
      iw >>= 1;
    }
    const size_t pix = (static_cast<size_t>(b) * Hin + ih) * Win + iw;
    v = c < C1 ? *reinterpret_cast<const float4*>(a1 + pix * C1 + c) : *reinterpret_cast<const float4*>(a2 + pix * C2 + (c - C1));
  }
  float4 h, l;
  h.x = tf32_rn(v.x); l.x = v.x - h.x;
  h.y = tf32_rn(v.y); l.y = v.y - h.y;
  h.z = tf32_rn(v.z); l.z = v.z - h.z;
  h.w = tf32_rn(v.w); l.w = v.w - h.w;
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// fp32 [rows, K] row-major, boxes of 32 columns x box_rows rows, SWIZZLE_128B
bool tmap_f32(CUtensorMap* m, const float* base, uint64_t K, uint64_t rows, uint32_t box_rows) {
  PFN_encodeTiled fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {K * 4};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 NHWC [N, H, W, C]: boxes of 32 channels x bw x bh x bn pixels, SWIZZLE_128B, out-of-bounds elements read as zero
bool tmap_f32_4d(CUtensorMap* m, const float* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint32_t bw, uint32_t bh, uint32_t bn) {
  PFN_encodeTiled fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 4, C * 4 * W, C * 4 * W * H};
  cuuint32_t box[4] = {TC_BK, bw, bh, bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_tc(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mA2h, const CUtensorMap& mA2l, const CUtensorMap& mWh,
                      const CUtensorMap& mWl, const TcArgs& a, cudaStream_t s) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(f32tc_gemm_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(f32tc_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  static int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  const int bn = tc_bn_for(a.N);
  if (!bn) return cudaErrorInvalidValue;
  const int tiles = (a.M / TC_BM) * (a.N / bn);
  if (bn == TC_BN) f32tc_gemm_kernel<160><<<dim3(tiles < sms ? tiles : sms), 192, TC_SMEM_BYTES, s>>>(mAh, mAl, mA2h, mA2l, mWh, mWl, a);
  else f32tc_gemm_kernel<128><<<dim3(tiles < sms ? tiles : sms), 192, TC_SMEM_BYTES, s>>>(mAh, mAl, mA2h, mA2l, mWh, mWl, a);
  return cudaGetLastError();
}

}  // namespace

bool f32tc_enabled() {  // default ON since round 2 (measured: profiles/R2o_*); WD_F32_TC=0 keeps every contraction on the FFMA kernel
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_F32_TC");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

bool f32tc_shape_ok(int M, int N, int K) { return M > 0 && M % TC_BM == 0 && tc_bn_for(N) != 0 && K % TC_BK == 0 && K >= TC_BK; }

cudaError_t f32tc_split(const float* a, float* hi, float* lo, size_t n, cudaStream_t s) {
  if (n & 3) return cudaErrorInvalidValue;
  const size_t n4 = n / 4;
  f32tc_split_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(a, hi, lo, n4);
  return cudaGetLastError();
}

cudaError_t f32tc_split_concat(const float* a1, const float* a2, int C1, int C2, size_t M, float* hi, float* lo, cudaStream_t s) {
  if (((C1 | C2) & 3) != 0) return cudaErrorInvalidValue;
  const size_t total4 = M * static_cast<size_t>(C1 + C2) / 4;
  f32tc_split_concat_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, s>>>(a1, a2, C1, C2, hi, lo, total4);
  return cudaGetLastError();
}

cudaError_t f32tc_im2col_split(const float* a1, const float* a2, int C1, int C2, int B, int Hin, int Win, int stride, int up, float* hi,
                               float* lo, cudaStream_t s) {
  const int Hout = up ? 2 * Hin : (stride == 2 ? Hin / 2 : Hin), Wout = up ? 2 * Win : (stride == 2 ? Win / 2 : Win);
  if (((C1 | C2) & 3) != 0) return cudaErrorInvalidValue;
  const size_t total4 = static_cast<size_t>(B) * Hout * Wout * 9 * (C1 + C2) / 4;
  f32tc_im2col_split_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, s>>>(a1, a2, C1, C2, Hin, Win, Hout, Wout, stride, up, hi, lo,
                                                                                        total4);
  return cudaGetLastError();
}

int f32tc_splits(int) { return 1; }  // the second accumulation level lives in the kernel's registers: no partial workspace

cudaError_t f32tc_gemm(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, int M, int N, int K, const float* bias,
                       const float* rowbias, int rb_ld, int rows_per_sample, const float* residual, float* out, int act_silu,
                       float* /*partial_ws*/, cudaStream_t s) {
  if (!f32tc_shape_ok(M, N, K)) return cudaErrorInvalidValue;
  CUtensorMap mAh, mAl, mWh, mWl;
  if (!tmap_f32(&mAh, a_hi, K, M, TC_BM) || !tmap_f32(&mAl, a_lo, K, M, TC_BM) || !tmap_f32(&mWh, w_hi, K, N, tc_bn_for(N)) ||
      !tmap_f32(&mWl, w_lo, K, N, tc_bn_for(N)))
    return cudaErrorInvalidValue;
  TcArgs a{};
  a.M = M;
  a.N = N;
  a.K = K;
  a.bias = bias;
  a.rowbias = rowbias;
  a.rb_ld = rb_ld;
  a.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  a.residual = residual;
  a.out = out;
  a.act_silu = act_silu;
  a.kb_per_chunk = TC_SPLIT_KB;
  return launch_tc(mAh, mAl, mAh, mAl, mWh, mWl, a, s);
}

// fp32 sub-pixel weights: packed [Cout][9 taps][C] -> [4 phases][Cout][4 taps][C] (taps that read the same input pixel summed)
__global__ void f32tc_upconv_fold_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int C) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(4) * Cout * 4 * C;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  size_t r = idx / C;
  const int t = static_cast<int>(r % 4);
  r /= 4;
  const int n = static_cast<int>(r % Cout);
  const int ph = static_cast<int>(r / Cout);
  const int a = ph >> 1, b = ph & 1, ty = t >> 1, tx = t & 1;
  const int ky0 = a == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = a == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
  const int kx0 = b == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = b == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
  float acc = 0.f;
  for (int ky = ky0; ky <= ky1; ++ky)
    for (int kx = kx0; kx <= kx1; ++kx) acc += w[(static_cast<size_t>(n) * 9 + ky * 3 + kx) * C + c];
  dst[idx] = acc;
}
cudaError_t f32tc_upconv_fold(const float* w_packed, float* dst, int Cout, int C, cudaStream_t s) {
  const size_t total = static_cast<size_t>(4) * Cout * 4 * C;
  f32tc_upconv_fold_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w_packed, dst, Cout, C);
  return cudaGetLastError();
}
bool f32tc_upconv_ok(int B, int H, int W, int C, int Cout) {
  const int bn = tc_bn_for(4 * Cout);
  return bn != 0 && Cout % bn == 0 && f32tc_conv_ok(B, H, W, C, 0, 4 * Cout);
}
cudaError_t f32tc_upconv(const float* a_hi, const float* a_lo, int C, int B, int H, int W, const float* w_hi, const float* w_lo, int Cout,
                         const float* bias, float* out, cudaStream_t s) {
  if (!f32tc_upconv_ok(B, H, W, C, Cout)) return cudaErrorInvalidValue;
  const int HW = H * W, K = 4 * C, N = 4 * Cout;
  uint32_t bw = W, bh, bn;
  if (W > TC_BM) {
    bw = TC_BM;
    bh = 1;
    bn = 1;
  } else if (HW >= TC_BM) {
    bh = TC_BM / W;
    bn = 1;
  } else {
    bh = H;
    bn = TC_BM / HW;
  }
  CUtensorMap mAh, mAl, mWh, mWl;
  if (!tmap_f32_4d(&mAh, a_hi, C, W, H, B, bw, bh, bn) || !tmap_f32_4d(&mAl, a_lo, C, W, H, B, bw, bh, bn) ||
      !tmap_f32(&mWh, w_hi, K, N, tc_bn_for(N)) || !tmap_f32(&mWl, w_lo, K, N, tc_bn_for(N)))
    return cudaErrorInvalidValue;
  TcArgs a{};
  a.M = B * HW;
  a.N = N;
  a.K = K;
  a.conv = 1;
  a.HW = HW;
  a.W = W;
  a.H = H;
  a.cb1 = C / TC_BK;
  a.cb2 = 0;
  a.bias = bias;
  a.rows_per_sample = HW;
  a.out = out;
  a.up = 1;
  a.up_cout = Cout;
  a.kb_per_chunk = TC_SPLIT_KB;
  return launch_tc(mAh, mAl, mAh, mAl, mWh, mWl, a, s);
}

// GEGLU projection with the gating in the epilogue: w_hi / w_lo and bias are in the f32tc_geglu_permute row order; out (and out_lo)
// are [M, N / 2].
cudaError_t f32tc_gemm_geglu(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, int M, int N, int K,
                             const float* bias_perm, float* out, float* out_lo, cudaStream_t s) {
  if (!f32tc_shape_ok(M, N, K) || N % (2 * TC_BN) || !bias_perm) return cudaErrorInvalidValue;
  CUtensorMap mAh, mAl, mWh, mWl;
  if (!tmap_f32(&mAh, a_hi, K, M, TC_BM) || !tmap_f32(&mAl, a_lo, K, M, TC_BM) || !tmap_f32(&mWh, w_hi, K, N, TC_BN) ||
      !tmap_f32(&mWl, w_lo, K, N, TC_BN))
    return cudaErrorInvalidValue;
  TcArgs a{};
  a.M = M;
  a.N = N;
  a.K = K;
  a.bias = bias_perm;
  a.rows_per_sample = 1;
  a.out = out;
  a.out_lo = out_lo;
  a.geglu = 1;
  a.kb_per_chunk = TC_SPLIT_KB;
  return launch_tc(mAh, mAl, mAh, mAl, mWh, mWl, a, s);
}

// rows of nn.Linear(K, N) of GEGLU.proj ([values (N / 2) ; gates (N / 2)], unet.py:125-128) -> 160-row tiles of 80 values + their 80 gates
__global__ void f32tc_geglu_permute_kernel(const float* __restrict__ w, float* __restrict__ dst, int N, size_t K) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(N) * K) return;
  const int n = static_cast<int>(idx / K);
  const size_t k = idx - static_cast<size_t>(n) * K;
  const int half = N / 2, gate = n >= half ? 1 : 0, j = gate ? n - half : n;
  const int row = (j / 80) * 160 + gate * 80 + (j % 80);
  dst[static_cast<size_t>(row) * K + k] = w[idx];
}
cudaError_t f32tc_geglu_permute(const float* w, float* dst, int N, int K, cudaStream_t s) {
  if (N % (2 * TC_BN)) return cudaErrorInvalidValue;
  const size_t n = static_cast<size_t>(N) * K;
  f32tc_geglu_permute_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(w, dst, N, static_cast<size_t>(K));
  return cudaGetLastError();
}

bool f32tc_conv_ok(int B, int H, int W, int C1, int C2, int N) {
  const int HW = H * W, M = B * HW;
  if (M <= 0 || M % TC_BM || tc_bn_for(N) == 0 || C1 % TC_BK || C2 % TC_BK || C1 <= 0) return false;
  if (W > TC_BM) return W % TC_BM == 0;  // a tile is a 128-pixel piece of one image row
  if (HW >= TC_BM) return HW % TC_BM == 0 && TC_BM % W == 0;
  return TC_BM % HW == 0 && H <= 256;
}

cudaError_t f32tc_conv3x3(const float* a1_hi, const float* a1_lo, int C1, const float* a2_hi, const float* a2_lo, int C2, int B, int H, int W,
                          const float* w_hi, const float* w_lo, int N, const float* bias, const float* rowbias, int rb_ld,
                          const float* residual, float* out, int act_silu, cudaStream_t s) {
  if (!f32tc_conv_ok(B, H, W, C1, C2, N)) return cudaErrorInvalidValue;
  const int HW = H * W, K = 9 * (C1 + C2);
  uint32_t bw = W, bh, bn;
  if (W > TC_BM) {
    bw = TC_BM;
    bh = 1;
    bn = 1;
  } else if (HW >= TC_BM) {
    bh = TC_BM / W;
    bn = 1;
  } else {
    bh = H;
    bn = TC_BM / HW;
  }
  CUtensorMap mAh, mAl, mA2h, mA2l, mWh, mWl;
  if (!tmap_f32_4d(&mAh, a1_hi, C1, W, H, B, bw, bh, bn) || !tmap_f32_4d(&mAl, a1_lo, C1, W, H, B, bw, bh, bn) ||
      !tmap_f32(&mWh, w_hi, K, N, tc_bn_for(N)) || !tmap_f32(&mWl, w_lo, K, N, tc_bn_for(N)))
    return cudaErrorInvalidValue;
  mA2h = mAh;
  mA2l = mAl;
  if (C2 > 0 && (!tmap_f32_4d(&mA2h, a2_hi, C2, W, H, B, bw, bh, bn) || !tmap_f32_4d(&mA2l, a2_lo, C2, W, H, B, bw, bh, bn)))
    return cudaErrorInvalidValue;
  TcArgs a{};
  a.M = B * HW;
  a.N = N;
  a.K = K;
  a.conv = 1;
  a.HW = HW;
  a.W = W;
  a.cb1 = C1 / TC_BK;
  a.cb2 = C2 / TC_BK;
  a.bias = bias;
  a.rowbias = rowbias;
  a.rb_ld = rb_ld;
  a.rows_per_sample = HW;
  a.residual = residual;
  a.out = out;
  a.act_silu = act_silu;
  a.kb_per_chunk = TC_SPLIT_KB;
  return launch_tc(mAh, mAl, mA2h, mA2l, mWh, mWl, a, s);
}

}  // namespace wd
