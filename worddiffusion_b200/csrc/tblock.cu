// Fused transformer block of unet.UNetModel's SpatialTransformer on tcgen05 / TMEM / TMA -- see tblock.cuh for the contract.
//
// Roles (320 threads, one persistent CTA per SM, 128-token tiles round-robin):
//   warp 0      TMA producer: the tile's g operand, then every weight "unit" of the tile in consumption order through a ring
//               of five 20 KB slots (K-major SWIZZLE_128B boxes; the per-sample attention operands as four 2 KB boxes per unit).
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue for the eight GEMM phases of a tile.
//   warps 2..9  epilogue: two warps per TMEM lane quarter, each owning one column half of its 32 rows.  Between GEMM phases they
//               turn the fp32 residual stream (TMEM) into the next 16-bit A operand in shared memory: LayerNorm-normalised copies
//               (the gamma / beta of the norms are folded into the weights that follow), softmax probabilities, GEGLU products.
// TMEM (512 columns): [0, 320) residual stream X (fp32, accumulated in place by every residual branch), [320, 448) GEGLU
// projection chunk, [448, 512) attention scores.
#include "tblock.cuh"
#include "epilogue.cuh"

#include <cstdlib>
#include <mutex>

namespace wd {
namespace {

constexpr int TB_THREADS = 320;
constexpr int TB_KB = TB_C / 64;                 // K blocks of a 320-wide operand
constexpr int TB_ABLK = TB_M * 128;              // one [128 rows x 64 cols] 16-bit K-major block: 16 KB
constexpr int TB_SLOT = 20480;                   // ring slot: [160 rows x 64] = 20 KB (W1 boxes use 16 KB, attention units 8 KB)
constexpr int TB_NSLOT = 5;
constexpr int TB_NCHUNK = TB_HID / TB_CHUNK;     // 20 feed-forward chunks
constexpr int OFF_A = 0;
constexpr int OFF_G = OFF_A + TB_KB * TB_ABLK;           // 81920: two [128 x 64] buffers (P of the attentions / GEGLU chunks)
constexpr int OFF_RING = OFF_G + 2 * TB_ABLK;            // 114688
constexpr int OFF_BFF = OFF_RING + TB_NSLOT * TB_SLOT;   // 217088
constexpr int OFF_CSM = OFF_BFF + 2 * TB_HID * 4;        // 227328
constexpr int OFF_STAT = OFF_CSM + 2 * 64 * 4;           // 227840
constexpr int OFF_BARS = OFF_STAT + 2 * TB_M * 8;        // 229888
constexpr int TB_SMEM = OFF_BARS + 256;                  // 230144
static_assert(TB_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t COL_X = 0, COL_G = 320, COL_S = 448;

enum Bar : int { B_RING_FULL = 0, B_RING_EMPTY = 5, B_A_FULL = 10, B_A_FREE = 11, B_ACC = 12, B_S = 13, B_A_READY = 14, B_P_READY = 15,
                 B_GACC_FULL = 16, B_GACC_FREE = 17, B_GBUF_FULL = 18, B_GBUF_EMPTY = 20, B_X_FREE = 22, B_COUNT = 23 };

WD_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// MN-major SWIZZLE_128B B operand: rows = K index (128 B = 64 N columns per row), 8-row atoms 1 KB apart; one 64-column N box
WD_DEVINL uint64_t desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// byte offset of the 16-byte chunk holding columns [col, col + 8) of `row` in a 320-wide operand made of five K-major
// SWIZZLE_128B blocks: chunk c16 of a row lives at position c16 ^ (row & 7)
WD_DEVINL uint32_t a_chunk_off(int row, int col) {
  return static_cast<uint32_t>((col >> 6) * TB_ABLK + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
}
WD_DEVINL float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// phases after which a debug launch stops (TBlockArgs::stage)
WD_DEVINL bool stop_after(int stage, int phase) { return stage != 0 && stage == phase; }

__global__ void __launch_bounds__(TB_THREADS, 1)
tblock_unet_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapWpi,
                   const __grid_constant__ CUtensorMap mapF0, const __grid_constant__ CUtensorMap mapF1,
                   const __grid_constant__ CUtensorMap mapF2, const __grid_constant__ CUtensorMap mapF3,
                   const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                   const __grid_constant__ CUtensorMap mapWpo, const __grid_constant__ CUtensorMap mapOut, const TBlockArgs args) {
  extern __shared__ __align__(1024) uint8_t tb_smem[];
  uint8_t* const smem = tb_smem;
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* const sA = smem + OFF_A;
  uint8_t* const sG = smem + OFF_G;
  uint8_t* const sRing = smem + OFF_RING;
  float* const sBff = reinterpret_cast<float*>(smem + OFF_BFF);
  float* const sC = reinterpret_cast<float*>(smem + OFF_CSM);
  float2* const sStat = reinterpret_cast<float2*>(smem + OFF_STAT);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = args.M / TB_M;
  const int stage = args.stage;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapG);
    tma_prefetch_desc(&mapWpi);
    tma_prefetch_desc(&mapF0);
    tma_prefetch_desc(&mapF1);
    tma_prefetch_desc(&mapF2);
    tma_prefetch_desc(&mapF3);
    tma_prefetch_desc(&mapW1);
    tma_prefetch_desc(&mapW2);
    tma_prefetch_desc(&mapWpo);
    tma_prefetch_desc(&mapOut);
    for (int i = 0; i < TB_NSLOT; ++i) {
      mbar_init(&bars[B_RING_FULL + i], 1);
      mbar_init(&bars[B_RING_EMPTY + i], 1);
    }
    mbar_init(&bars[B_A_FULL], 1);
    mbar_init(&bars[B_A_FREE], 1);
    mbar_init(&bars[B_ACC], 1);
    mbar_init(&bars[B_S], 1);
    mbar_init(&bars[B_A_READY], 8);
    mbar_init(&bars[B_P_READY], 8);
    mbar_init(&bars[B_GACC_FULL], 1);
    mbar_init(&bars[B_GACC_FREE], 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_GBUF_FULL + i], 8);
      mbar_init(&bars[B_GBUF_EMPTY + i], 1);
    }
    mbar_init(&bars[B_X_FREE], 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // static data: the folded GEGLU bias (weights only) -> shared memory
  for (int i = threadIdx.x; i < 2 * TB_HID / 4; i += TB_THREADS)
    reinterpret_cast<float4*>(sBff)[i] = __ldg(reinterpret_cast<const float4*>(args.b_ff) + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
      auto acquire = [&](uint32_t bytes) -> uint8_t* {
        mbar_wait(&bars[B_RING_EMPTY + slot], phase ^ 1);
        mbar_arrive_expect_tx(&bars[B_RING_FULL + slot], bytes);
        return sRing + slot * TB_SLOT;
      };
      auto advance = [&]() {
        if (++slot == TB_NSLOT) { slot = 0; phase ^= 1; }
      };
      auto weight_320_half = [&](const CUtensorMap* mp, int nh) {  // one N half of a [320 x 320] weight: 5 K blocks of [160 x 64]
        for (int kb = 0; kb < TB_KB; ++kb) {
          uint8_t* dst = acquire(160 * 128);
          tma_load_2d(dst, mp, &bars[B_RING_FULL + slot], kb * 64, nh * 160);
          advance();
        }
      };
      auto fold_units = [&](const CUtensorMap* mp, int sample) {  // five units of four [16 keys x 64] boxes (one per head)
        for (int b = 0; b < TB_KB; ++b) {
          uint8_t* dst = acquire(TB_HEADS * TB_KEYS * 128);
          for (int h = 0; h < TB_HEADS; ++h) tma_load_3d(dst + h * (TB_KEYS * 128), mp, &bars[B_RING_FULL + slot], h * TB_C + b * 64, 0, sample);
          advance();
        }
      };
      auto w1_units = [&](int c) {
        for (int kb = 0; kb < TB_KB; ++kb) {
          uint8_t* dst = acquire(2 * TB_CHUNK * 128);
          tma_load_2d(dst, &mapW1, &bars[B_RING_FULL + slot], kb * 64, c * 2 * TB_CHUNK);
          advance();
        }
      };
      auto w2_units = [&](int c) {
        for (int nh = 0; nh < 2; ++nh) {
          uint8_t* dst = acquire(160 * 128);
          tma_load_2d(dst, &mapW2, &bars[B_RING_FULL + slot], c * TB_CHUNK, nh * 160);
          advance();
        }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        const int m0 = tile * TB_M;
        const int sample = m0 / args.HW;
        // the first five weight units fill the ring while the previous tile finishes (exactly the ring's capacity: issuing more
        // before the g operand could block on a slot that only this tile's MMAs -- which wait for g -- can free)
        weight_320_half(&mapWpi, 0);
        // the A buffer is free once the previous tile's output store has read it
        mbar_wait(&bars[B_A_FREE], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[B_A_FULL], TB_KB * TB_ABLK);
        for (int kb = 0; kb < TB_KB; ++kb) tma_load_2d(sA + kb * TB_ABLK, &mapG, &bars[B_A_FULL], kb * 64, m0);
        weight_320_half(&mapWpi, 1);
        if (stop_after(stage, 1)) continue;
        fold_units(&mapF0, sample);
        fold_units(&mapF1, sample);
        if (stop_after(stage, 2)) continue;
        fold_units(&mapF2, sample);
        fold_units(&mapF3, sample);
        if (stop_after(stage, 3)) continue;
        for (int c = 0; c < TB_NCHUNK; ++c) {
          w1_units(c);
          if (c > 0) w2_units(c - 1);
        }
        w2_units(TB_NCHUNK - 1);
        if (stop_after(stage, 4)) continue;
        weight_320_half(&mapWpo, 0);
        weight_320_half(&mapWpo, 1);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      constexpr uint32_t ID_BF16_160 = make_idesc_bf16_f32(TB_M, 160);
      constexpr uint32_t ID_F16_160 = make_idesc_f16_f32(TB_M, 160);
      constexpr uint32_t ID_F16_128 = make_idesc_f16_f32(TB_M, 128);
      constexpr uint32_t ID_F16_64 = make_idesc_f16_f32(TB_M, 64);
      constexpr uint32_t ID_F16_64_MN = ID_F16_64 | (1u << 16);  // B operand MN-major
      int slot = 0;
      uint32_t phase = 0;
      uint32_t n_a_ready = 0, n_p_ready = 0, n_gchunk = 0;  // completed-phase counters of the barriers this thread waits on
      uint32_t n_gbuf[2] = {0, 0};
      auto ring_wait = [&]() -> uint32_t {
        mbar_wait(&bars[B_RING_FULL + slot], phase);
        tc_fence_after();
        return smem_u32(sRing + slot * TB_SLOT);
      };
      auto ring_release = [&]() {
        umma_commit(&bars[B_RING_EMPTY + slot]);
        if (++slot == TB_NSLOT) { slot = 0; phase ^= 1; }
      };
      // X[:, nh*160 ..] (+)= A[128 x 320] W^T for a [320 x 320] weight streamed as 2 x 5 units
      auto gemm_320 = [&](uint32_t idesc, bool fresh) {
        for (int nh = 0; nh < 2; ++nh)
          for (int kb = 0; kb < TB_KB; ++kb) {
            const uint64_t b_desc = make_smem_desc_sw128(ring_wait());
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + kb * TB_ABLK));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem_base + COL_X + nh * 160, a_desc + 2 * k, b_desc + 2 * k, idesc, (!fresh || (kb | k) != 0) ? 1u : 0u);
            ring_release();
          }
      };
      auto wait_a_ready = [&]() {
        mbar_wait(&bars[B_A_READY], n_a_ready & 1);
        ++n_a_ready;
        tc_fence_after();
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        // ---- proj_in: X = g Wpi^T ----
        mbar_wait(&bars[B_A_FULL], it & 1);
        mbar_wait(&bars[B_X_FREE], (it & 1) ^ 1);  // the previous tile's epilogue has read its last accumulator
        tc_fence_after();
        gemm_320(ID_BF16_160, true);
        umma_commit(&bars[B_ACC]);
        if (stop_after(stage, 1)) continue;
        // ---- two cross-attentions: S = xhat M^T ; X += P N^T ----
        for (int a = 0; a < 2; ++a) {
          wait_a_ready();
          for (int kb = 0; kb < TB_KB; ++kb) {
            const uint64_t b_desc = make_smem_desc_sw128(ring_wait());
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + kb * TB_ABLK));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + COL_S, a_desc + 2 * k, b_desc + 2 * k, ID_F16_64, (kb | k) != 0);
            ring_release();
          }
          umma_commit(&bars[B_S]);
          mbar_wait(&bars[B_P_READY], n_p_ready & 1);
          ++n_p_ready;
          tc_fence_after();
          const uint64_t p_desc = make_smem_desc_sw128(smem_u32(sG));
          for (int nb = 0; nb < TB_KB; ++nb) {
            const uint64_t b_desc = desc_mn_sw128(ring_wait());
#pragma unroll
            for (int k = 0; k < 4; ++k)  // K step k = the 16 key slots of head k: rows [16 k, 16 k + 16) of the MN-major unit
              umma_f16_ss(tmem_base + COL_X + nb * 64, p_desc + 2 * k, b_desc + 128 * k, ID_F16_64_MN, 1u);
            ring_release();
          }
          umma_commit(&bars[B_ACC]);
          if (stop_after(stage, 2 + a)) break;
        }
        if (stop_after(stage, 2) || stop_after(stage, 3)) continue;
        // ---- feed-forward: per 64-column hidden chunk  Gacc = xhat W1_c^T ; X += GEGLU(Gacc) W2_c^T ----
        wait_a_ready();
        auto mma2 = [&](int c) {
          const int b = c & 1;
          mbar_wait(&bars[B_GBUF_FULL + b], n_gbuf[b] & 1);
          ++n_gbuf[b];
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sG + b * TB_ABLK));
          for (int nh = 0; nh < 2; ++nh) {
            const uint64_t b_desc = make_smem_desc_sw128(ring_wait());
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + COL_X + nh * 160, a_desc + 2 * k, b_desc + 2 * k, ID_BF16_160, 1u);
            ring_release();
          }
          umma_commit(&bars[B_GBUF_EMPTY + b]);
        };
        for (int c = 0; c < TB_NCHUNK; ++c) {
          if (n_gchunk > 0) {  // the epilogue has drained the previous chunk's accumulator
            mbar_wait(&bars[B_GACC_FREE], (n_gchunk - 1) & 1);
            tc_fence_after();
          }
          ++n_gchunk;
          for (int kb = 0; kb < TB_KB; ++kb) {
            const uint64_t b_desc = make_smem_desc_sw128(ring_wait());
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + kb * TB_ABLK));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + COL_G, a_desc + 2 * k, b_desc + 2 * k, ID_F16_128, (kb | k) != 0);
            ring_release();
          }
          umma_commit(&bars[B_GACC_FULL]);
          if (c > 0) mma2(c - 1);
        }
        mma2(TB_NCHUNK - 1);
        umma_commit(&bars[B_ACC]);
        if (stop_after(stage, 4)) continue;
        // ---- proj_out: X = x3 Wpo^T (fresh accumulator; the epilogue adds bias and x_in) ----
        wait_a_ready();
        gemm_320(ID_F16_160, true);
        umma_commit(&bars[B_ACC]);
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // column half of the row
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int et = threadIdx.x - 64;   // 0 .. 255
    uint32_t n_acc = 0, n_s = 0, n_gacc = 0;
    uint32_t n_gbuf_empty[2] = {0, 0};

    auto wait_acc = [&]() {
      mbar_wait(&bars[B_ACC], n_acc & 1);
      ++n_acc;
      tc_fence_after();
    };
    auto arrive_warp = [&](int bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[bar]);
    };
    // X + cb -> 16-bit operand copy in the A buffer.  NORM: (x - mean) * rstd as fp16 (LayerNorm without its affine part, which
    // is folded into the weights that consume the copy); else the raw value as fp16.
    auto x_to_a = [&](const float* cb, bool norm) {
      float mu = 0.f, rstd = 1.f;
      const int c0 = half * 160;
      if (norm) {
        float s = 0.f, sq = 0.f;
#pragma unroll 1
        for (int g = 0; g < 5; ++g) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + COL_X + c0 + g * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(cb + c0 + g * 32 + i));
            const float x0 = __uint_as_float(v[i]) + b4.x, x1 = __uint_as_float(v[i + 1]) + b4.y;
            const float x2 = __uint_as_float(v[i + 2]) + b4.z, x3 = __uint_as_float(v[i + 3]) + b4.w;
            s += (x0 + x1) + (x2 + x3);
            sq = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, sq))));
          }
        }
        sStat[half * TB_M + row] = make_float2(s, sq);
        named_barrier_sync(1 + q, 64);  // the two warps that share this lane quarter
        const float2 o = sStat[(half ^ 1) * TB_M + row];
        mu = (s + o.x) * (1.0f / TB_C);
        const float var = fmaxf((sq + o.y) * (1.0f / TB_C) - mu * mu, 0.f);
        rstd = rsqrtf(var + args.ln_eps);
      }
#pragma unroll 1
      for (int g = 0; g < 5; ++g) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + COL_X + c0 + g * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const int col = c0 + g * 32 + c8 * 8;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(cb + col));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(cb + col + 4));
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = (__uint_as_float(v[c8 * 8 + j]) + bb[j] - mu) * rstd;
          *reinterpret_cast<uint4*>(sA + a_chunk_off(row, col)) =
              make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
        }
      }
      fence_proxy_async();  // generic-proxy writes of the operand -> visible to the tensor core / TMA (async proxy)
      tc_fence_before();
    };
    // scores (+ per-sample constants) -> softmax over the L keys of each head -> fp16 probabilities, K-major P tile
    auto softmax_to_p = [&](const float* cs) {
      mbar_wait(&bars[B_S], n_s & 1);
      ++n_s;
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + COL_S + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      uint32_t pk[16];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float sc[16];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sc[j] = (j < args.L) ? __uint_as_float(v[hh * 16 + j]) + cs[half * 32 + hh * 16 + j] : -INFINITY;
          mx = fmaxf(mx, sc[j]);
        }
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sc[j] = ex2_fast(sc[j] - mx);  // keys beyond L: ex2(-inf) = 0
          l += sc[j];
        }
        const float inv = rcp_fast(l);
#pragma unroll
        for (int j = 0; j < 16; j += 2) pk[hh * 8 + j / 2] = pack_f16x2(sc[j] * inv, sc[j + 1] * inv);
      }
      uint8_t* const prow = sG + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (row & 7)) << 4)) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      fence_proxy_async();
    };
    // the A buffer holds a finished [128 x 320] fp16 tile: store it to `out`, then free the buffer and the accumulator
    auto store_a_tile = [&](int m0) {
      arrive_warp(B_X_FREE);
      named_barrier_sync(5, 256);
      if (et == 0) {
        for (int kb = 0; kb < TB_KB; ++kb) tma_store_2d(&mapOut, sA + kb * TB_ABLK, kb * 64, m0);
        bulk_commit_group();
        bulk_wait_group_read<0>();
        mbar_arrive(&bars[B_A_FREE]);
      }
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      const int m0 = tile * TB_M;
      const int sample = m0 / args.HW;
      // per-sample score constants of both attentions -> shared memory, slot (h, j) = h * 16 + j
      if (et < 128) {
        const int a = et >> 6, n = et & 63, h = n >> 4, j = n & 15;
        const float* cv = a ? args.cvec2 : args.cvec1;
        sC[et] = (j < args.L) ? __ldg(cv + (static_cast<size_t>(sample) * args.L + j) * args.cvec_ld + h) : 0.f;
      }
      named_barrier_sync(6, 256);

      // ---- after proj_in ----
      wait_acc();
      x_to_a(args.cb, true);
      if (stop_after(stage, 1)) { store_a_tile(m0); continue; }
      arrive_warp(B_A_READY);
      // ---- attention 1 / 2 ----
      bool stopped = false;
      for (int a = 0; a < 2; ++a) {
        softmax_to_p(sC + a * 64);
        arrive_warp(B_P_READY);
        wait_acc();
        x_to_a(args.cb + (1 + a) * TB_C, true);
        if (stop_after(stage, 2 + a)) { store_a_tile(m0); stopped = true; break; }
        arrive_warp(B_A_READY);
      }
      if (stopped) continue;
      // ---- feed-forward chunks: GEGLU of the projection accumulator -> bf16 operand chunk ----
#pragma unroll 1
      for (int c = 0; c < TB_NCHUNK; ++c) {
        const int b = c & 1;
        mbar_wait(&bars[B_GACC_FULL], n_gacc & 1);
        ++n_gacc;
        tc_fence_after();
        uint32_t vv[32], vg[32];
        tmem_ld_32x32b_x32(t_row + COL_G + half * 32, vv);
        tmem_ld_32x32b_x32(t_row + COL_G + TB_CHUNK + half * 32, vg);
        tmem_ld_wait();
        tc_fence_before();
        arrive_warp(B_GACC_FREE);
        const float* bv = sBff + c * 2 * TB_CHUNK + half * 32;
        const float* bg = bv + TB_CHUNK;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float o0 = (__uint_as_float(vv[i]) + bv[i]) * gelu_fast_f(__uint_as_float(vg[i]) + bg[i]);
          const float o1 = (__uint_as_float(vv[i + 1]) + bv[i + 1]) * gelu_fast_f(__uint_as_float(vg[i + 1]) + bg[i + 1]);
          pk[i / 2] = pack_bf16x2(o0, o1);
        }
        if (c >= 2) {  // the MMAs of chunk c - 2 have read this operand buffer
          mbar_wait(&bars[B_GBUF_EMPTY + b], n_gbuf_empty[b] & 1);
          ++n_gbuf_empty[b];
        }
        uint8_t* const grow = sG + b * TB_ABLK + row * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(grow + (((half * 4 + k) ^ (row & 7)) << 4)) = make_uint4(pk[k * 4], pk[k * 4 + 1], pk[k * 4 + 2], pk[k * 4 + 3]);
        fence_proxy_async();
        arrive_warp(B_GBUF_FULL + b);
      }
      // the last two chunks' MMAs are covered by the accumulator barrier below; account for their buffer releases
      n_gbuf_empty[0] += 1;
      n_gbuf_empty[1] += 1;
      // ---- x3 (raw) -> operand of proj_out ----
      wait_acc();
      x_to_a(args.cb + 3 * TB_C, false);
      if (stop_after(stage, 4)) { store_a_tile(m0); continue; }
      arrive_warp(B_A_READY);
      // ---- proj_out accumulator + bias + x_in -> fp16 tile in the A buffer (its MMAs have retired), GroupNorm partials ----
      wait_acc();
      {
        const int c0 = half * 160;
        const __half* xr = args.x_in + static_cast<size_t>(m0 + row) * args.x_in_ld + c0;
        float gs[32];  // 16 groups of 10 columns: [2 g] = sum, [2 g + 1] = sum of squares
#pragma unroll
        for (int i = 0; i < 32; ++i) gs[i] = 0.f;
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + COL_X + c0 + g * 32, v);
          uint4 r4[4];
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) r4[c8] = __ldg(reinterpret_cast<const uint4*>(xr + g * 32 + c8 * 8));
          tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const int cl = g * 32 + c8 * 8;  // column inside this thread's 160
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.b_po + c0 + cl));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.b_po + c0 + cl + 4));
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const uint32_t ru[4] = {r4[c8].x, r4[c8].y, r4[c8].z, r4[c8].w};
            float f[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 t = unpack_f16x2(ru[j]);
              f[2 * j] = __uint_as_float(v[c8 * 8 + 2 * j]) + bb[2 * j] + t.x;
              f[2 * j + 1] = __uint_as_float(v[c8 * 8 + 2 * j + 1]) + bb[2 * j + 1] + t.y;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int grp = (cl + j) / 10;  // compile-time after unrolling
              gs[2 * grp] += f[j];
              gs[2 * grp + 1] = fmaf(f[j], f[j], gs[2 * grp + 1]);
            }
            *reinterpret_cast<uint4*>(sA + a_chunk_off(row, c0 + cl)) =
                make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
          }
        }
        tc_fence_before();
        fence_proxy_async();
        if (args.gn_partial) {
          // the 32 rows of a warp belong to one sample and one 32-row slot: reduce over the rows, 8 groups per pass
          const int mw = m0 + q * 32;
          const int slot = (mw % args.HW) >> 5, nslot = args.HW >> 5;
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            float part[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) part[i] = gs[p * 16 + i];
            const float tot = warp_transpose_reduce16(part, lane);
            if (lane < 16) {
              const int g = half * 16 + p * 8 + (lane >> 1);
              args.gn_partial[((static_cast<size_t>(sample) * 32 + g) * nslot + slot) * 2 + (lane & 1)] = tot;
            }
          }
        }
      }
      store_a_tile(m0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ----------------------------------------------------------------------------------------------
// weight folding (once per weight load) and per-trajectory score constants
// ----------------------------------------------------------------------------------------------
// W_fold rows of one attention, bf16 [2560, 320]:
//   rows [0, 1280):    (h, k) -> sl2 gamma[k] sum_d Wq[h 80 + d, k] Wk[h 80 + d, c]          ("M" part, sl2 = 80^-1/2 log2 e)
//   rows [1280, 2560): (h, n) -> sum_d Wout[n, h 80 + d] Wv[h 80 + d, c]                     ("N" part)
__global__ void __launch_bounds__(320) tblock_fold_weights_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                                                                 const float* __restrict__ wv, const float* __restrict__ wo,
                                                                 const float* __restrict__ gamma, __nv_bfloat16* __restrict__ w_fold) {
  const int r = blockIdx.x, c = threadIdx.x;  // one output row per block, one column per thread
  const int part = r / (TB_HEADS * TB_C), rr = r % (TB_HEADS * TB_C), h = rr / TB_C, k = rr % TB_C;
  float acc = 0.f;
  if (part == 0) {
    for (int d = 0; d < TB_DH; ++d) acc = fmaf(__ldg(wq + (h * TB_DH + d) * TB_C + k), __ldg(wk + (h * TB_DH + d) * TB_C + c), acc);
    acc *= __ldg(gamma + k) * (0.11180339887498949f * 1.4426950408889634f);  // 80^-1/2 * log2(e)
  } else {
    for (int d = 0; d < TB_DH; ++d) acc = fmaf(__ldg(wo + k * TB_C + h * TB_DH + d), __ldg(wv + (h * TB_DH + d) * TB_C + c), acc);
  }
  w_fold[static_cast<size_t>(r) * TB_C + c] = __float2bfloat16(acc);
}
// u[h][c] = sl2 sum_d (sum_k beta[k] Wq[h 80 + d, k]) Wk[h 80 + d, c]
__global__ void __launch_bounds__(320) tblock_fold_u_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                                                           const float* __restrict__ beta, float* __restrict__ u) {
  __shared__ float t[TB_C];
  const int n = threadIdx.x;
  float a = 0.f;
  for (int k = 0; k < TB_C; ++k) a = fmaf(__ldg(beta + k), __ldg(wq + n * TB_C + k), a);
  t[n] = a;
  __syncthreads();
  for (int h = 0; h < TB_HEADS; ++h) {
    float acc = 0.f;
    for (int d = 0; d < TB_DH; ++d) acc = fmaf(t[h * TB_DH + d], __ldg(wk + (h * TB_DH + d) * TB_C + n), acc);
    u[h * TB_C + n] = acc * (0.11180339887498949f * 1.4426950408889634f);
  }
}
// cvec[row][hh] = sum_c ctx[row][c] u[hh][c] for the `heads` (= attentions x 4) pooled u vectors; one warp per (row, 4 heads)
__global__ void __launch_bounds__(256) tblock_cvec_kernel(const __nv_bfloat16* __restrict__ ctx, const float* __restrict__ u,
                                                         float* __restrict__ cvec, int rows, int heads) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int groups = heads / TB_HEADS;
  const int row = w / groups, h0 = (w % groups) * TB_HEADS;
  if (row >= rows) return;
  float acc[TB_HEADS] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane; c < TB_C; c += 32) {
    const float x = __bfloat162float(ctx[static_cast<size_t>(row) * TB_C + c]);
#pragma unroll
    for (int h = 0; h < TB_HEADS; ++h) acc[h] = fmaf(x, __ldg(u + (h0 + h) * TB_C + c), acc[h]);
  }
#pragma unroll
  for (int h = 0; h < TB_HEADS; ++h) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < TB_HEADS; ++h) cvec[static_cast<size_t>(row) * heads + h0 + h] = acc[h];
  }
}

}  // namespace

cudaError_t tblock_fold_weights_launch(const float* wq, const float* wk, const float* wv, const float* wo, const float* gamma,
                                       const float* beta, __nv_bfloat16* w_fold, float* u, cudaStream_t s) {
  tblock_fold_weights_kernel<<<TB_FOLD_N, TB_C, 0, s>>>(wq, wk, wv, wo, gamma, w_fold);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  tblock_fold_u_kernel<<<1, TB_C, 0, s>>>(wq, wk, beta, u);
  return cudaGetLastError();
}

cudaError_t tblock_cvec_launch(const __nv_bfloat16* ctx, const float* u, float* cvec, int rows, int heads, cudaStream_t s) {
  if (heads < TB_HEADS || heads % TB_HEADS) return cudaErrorInvalidValue;
  const long long warps = static_cast<long long>(rows) * (heads / TB_HEADS);
  tblock_cvec_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, s>>>(ctx, u, cvec, rows, heads);
  return cudaGetLastError();
}

bool tblock_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_TBLOCK");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

cudaError_t tblock_launch(const TBlockLaunch& L, int num_sms, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tblock_unet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const TBlockArgs& a = L.args;
  if (a.M <= 0 || a.M % TB_M || a.HW % TB_M || a.L < 1 || a.L > TB_KEYS || a.x_in_ld % 8) return cudaErrorInvalidValue;
  const int tiles = a.M / TB_M;
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_pdl(tblock_unet_kernel, dim3(grid), dim3(TB_THREADS), TB_SMEM, stream, L.mapG, L.mapWpi, L.mapF[0], L.mapF[1],
                    L.mapF[2], L.mapF[3], L.mapW1, L.mapW2, L.mapWpo, L.mapOut, a);
}

}  // namespace wd
