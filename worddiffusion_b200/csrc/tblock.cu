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
//
// CTA-pair build (PAIR, cta_group::2; used when a sample holds a multiple of 256 tokens).  ncu of the single-CTA build
// (profiles/R2d_ncu_tblock.txt): tensor pipe 33 %, the MMA thread mostly waiting for weight units -- the feed-forward phase needs
// 120 KB of weights per 1920 MMA clocks (62 B/clk) and one SM ingests less than that with 100 KB of loads in flight.  Two CTAs of
// a cluster therefore share every weight unit: a tile is 256 tokens (128 rows per CTA: own operand copies, own TMEM lanes, own
// epilogue warps), each CTA loads HALF of every weight unit (the N / 2 rows that tcgen05.mma.cta_group::2 reads from its shared
// memory), the leader CTA issues all MMAs, commits are multicast to both CTAs and the epilogue -> MMA barriers live in the leader
// (16 warp arrivals).  Weight bytes per token halve.
#include "tblock.cuh"
#include "epilogue.cuh"

#include <cstdlib>
#include <mutex>

namespace wd {
// phase timeline (clock64) of CTA 0 under TBlockArgs::trace: [role 0 = MMA thread, 1 = epilogue warp 2 lane 0][tile < 8][stamp < 64]
__device__ unsigned long long g_tb_trace[2 * 8 * 64];
namespace {

#define TB_STAMP(role, it_, idx)                                                                      \
  do {                                                                                                \
    if (args.trace && blockIdx.x == 0 && (it_) < 8) g_tb_trace[((role) * 8 + (it_)) * 64 + (idx)] = clock64(); \
  } while (0)

constexpr int TB_THREADS = 320;
constexpr int TB_KB = TB_C / 64;                 // K blocks of a 320-wide operand
constexpr int TB_ABLK = TB_M * 128;              // one [128 rows x 64 cols] 16-bit K-major block: 16 KB
constexpr int TB_RING = 92160;                   // operand ring: four 20 KB slots ([160 rows x 64]; W1 boxes use 16 KB, attention units
                                                 // 8-16 KB), or nine 10 KB slots in the CTA-pair build (every unit is half as large per
                                                 // CTA).  (One slot fewer than the first build: the 10 KB went to the bias vectors below --
                                                 // the phase trace showed the hand-overs waiting on L2 latency of per-group bias loads.)
constexpr int TB_MAXSLOT = 10;
constexpr int TB_NCHUNK = TB_HID / TB_CHUNK;     // 20 feed-forward chunks
constexpr int OFF_A = 0;
constexpr int OFF_G = OFF_A + TB_KB * TB_ABLK;           // 81920: two [128 x 64] buffers (P of the attentions / GEGLU chunks)
constexpr int OFF_RING = OFF_G + 2 * TB_ABLK;            // 114688
constexpr int OFF_BFF = OFF_RING + TB_RING;              // 206848
constexpr int OFF_CB = OFF_BFF + 2 * TB_HID * 4;         // 217088: cumulative stream biases [4][320] ++ proj_out bias [320]
constexpr int OFF_CSM = OFF_CB + 5 * TB_C * 4;           // 223488
constexpr int OFF_STAT = OFF_CSM + 4 * 64 * 4;           // 228352 (score constants: 2 attentions x up to 2 samples x 64)
constexpr int OFF_BARS = OFF_STAT + 2 * TB_M * 8;        // 230400
constexpr int TB_SMEM = OFF_BARS + 320;                  // 230720
static_assert(TB_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t COL_X = 0, COL_G = 320, COL_S = 448;

enum Bar : int { B_RING_FULL = 0, B_RING_EMPTY = 10, B_A_FULL = 20, B_A_FREE = 21, B_ACC = 22, B_S = 23, B_A_READY = 24, B_P_READY = 25,
                 B_GACC_FULL = 26, B_GACC_FREE = 27, B_GBUF_FULL = 28, B_GBUF_EMPTY = 30, B_X_FREE = 32, B_COUNT = 33 };

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

WD_DEVINL void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// X + cb -> 16-bit operand copy in the A buffer.  NORM: (x - mean) * rstd as fp16 (LayerNorm without its affine part, which is folded
// into the weights that consume the copy); else the raw value as fp16.  One out-of-line copy for the four call sites: fully inlined
// the kernel was 148 KB of SASS, far beyond the instruction cache.  TMEM loads run one 32-column group ahead of the arithmetic; the
// biases come from shared memory (the first build read them from global memory per group: five exposed L2 round trips per pass).
template <int G>
WD_DEVINL void x_stats_group(const uint32_t (&v)[32], const float* cbg, float& s, float& sq) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(cbg + i);
    const float x0 = __uint_as_float(v[i]) + b4.x, x1 = __uint_as_float(v[i + 1]) + b4.y;
    const float x2 = __uint_as_float(v[i + 2]) + b4.z, x3 = __uint_as_float(v[i + 3]) + b4.w;
    s += (x0 + x1) + (x2 + x3);
    sq = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, sq))));
  }
}
WD_DEVINL void x_copy_group(const uint32_t (&v)[32], const float* cbg, uint8_t* sA, int row, int col0, float mu, float rstd) {
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) {
    const float4 b0 = *reinterpret_cast<const float4*>(cbg + c8 * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(cbg + c8 * 8 + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (__uint_as_float(v[c8 * 8 + j]) + bb[j] - mu) * rstd;
    *reinterpret_cast<uint4*>(sA + a_chunk_off(row, col0 + c8 * 8)) =
        make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
  }
}
__device__ __noinline__ void x_to_a_fn(uint32_t t_row, const float* cb, bool norm, uint8_t* sA, float2* sStat, int row, int half, int q,
                                       float ln_eps) {
  float mu = 0.f, rstd = 1.f;
  const int c0 = half * 160;
  const uint32_t t0 = t_row + COL_X + c0;
  uint32_t v0[32], v1[32];
  if (norm) {
    float s = 0.f, sq = 0.f;
    tmem_ld_32x32b_x32(t0, v0);
#pragma unroll 1
    for (int gp = 0; gp < 2; ++gp) {  // groups (0, 1), (2, 3)
      tmem_ld_wait();
      tmem_ld_32x32b_x32(t0 + (2 * gp + 1) * 32, v1);
      x_stats_group<0>(v0, cb + c0 + (2 * gp) * 32, s, sq);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(t0 + (2 * gp + 2) * 32, v0);
      x_stats_group<0>(v1, cb + c0 + (2 * gp + 1) * 32, s, sq);
    }
    tmem_ld_wait();
    x_stats_group<0>(v0, cb + c0 + 4 * 32, s, sq);
    sStat[half * TB_M + row] = make_float2(s, sq);
    named_barrier_sync(1 + q, 64);  // the two warps that share this lane quarter
    const float2 o = sStat[(half ^ 1) * TB_M + row];
    mu = (s + o.x) * (1.0f / TB_C);
    const float var = fmaxf((sq + o.y) * (1.0f / TB_C) - mu * mu, 0.f);
    rstd = rsqrtf(var + ln_eps);
  }
  tmem_ld_32x32b_x32(t0, v0);
#pragma unroll 1
  for (int gp = 0; gp < 2; ++gp) {
    tmem_ld_wait();
    tmem_ld_32x32b_x32(t0 + (2 * gp + 1) * 32, v1);
    x_copy_group(v0, cb + c0 + (2 * gp) * 32, sA, row, c0 + (2 * gp) * 32, mu, rstd);
    tmem_ld_wait();
    tmem_ld_32x32b_x32(t0 + (2 * gp + 2) * 32, v0);
    x_copy_group(v1, cb + c0 + (2 * gp + 1) * 32, sA, row, c0 + (2 * gp + 1) * 32, mu, rstd);
  }
  tmem_ld_wait();
  x_copy_group(v0, cb + c0 + 4 * 32, sA, row, c0 + 4 * 32, mu, rstd);
  fence_proxy_async();  // generic-proxy writes of the operand -> visible to the tensor core / TMA (async proxy)
  tc_fence_before();
}

// GroupNorm of the input tile in place (fp16 x_in -> bf16 g): per-sample scale / shift from the producer's partial statistics.
// Out of line (once per tile; the kernel's hot loops are sensitive to its code size).
template <int SPT>
__device__ __noinline__ void gn_in_fn(const TBlockArgs& args, uint8_t* sA, uint8_t* sG, uint64_t* a_full, uint32_t parity, int sample,
                                      int n_samples, int et, int row, int half, int q) {
    // ---- GroupNorm of the input tile in place (fp16 x_in -> bf16 g): per-sample scale / shift from the producer's partials ----
        float* const sGn = reinterpret_cast<float*>(sG);  // [SPT samples][mean/rstd 64 | scale 320 | shift 320] (sG is idle here)
        constexpr int GN_STRIDE = 64 + 2 * TB_C;
        if (et < 32 * SPT) {
          const int sl = et >> 5, g = et & 31;
          float S = 0.f, Q = 0.f;
          if (sample + sl < n_samples) {
            const float2* part = reinterpret_cast<const float2*>(args.gn_in_partial) + (static_cast<size_t>(sample + sl) * 32 + g) * args.gn_in_slots;
            for (int k = 0; k < args.gn_in_slots; ++k) {  // fixed order: bit-reproducible, the same fold as groupnorm_apply_bulk_kernel
              const float2 t = __ldg(part + k);
              S += t.x;
              Q += t.y;
            }
          }
          const float inv_n = 1.0f / static_cast<float>((TB_C / 32) * args.HW);
          const float mean = S * inv_n;
          const float var = fmaxf(Q * inv_n - mean * mean, 0.f);
          sGn[sl * GN_STRIDE + g] = mean;
          sGn[sl * GN_STRIDE + 32 + g] = rsqrtf(var + args.gn_eps);
        }
        named_barrier_sync(6, 256);
        for (int i = et; i < TB_C * SPT; i += 256) {
          const int sl = i / TB_C, c = i - sl * TB_C, g = c / (TB_C / 32);
          const float sc = sGn[sl * GN_STRIDE + 32 + g] * __ldg(args.gn_gamma + c);
          sGn[sl * GN_STRIDE + 64 + c] = sc;
          sGn[sl * GN_STRIDE + 64 + TB_C + c] = __ldg(args.gn_beta + c) - sGn[sl * GN_STRIDE + g] * sc;
        }
        named_barrier_sync(6, 256);
        mbar_wait(a_full, parity);  // this CTA's rows have landed
        {
          const int sl = SPT == 2 ? (q >> 1) : 0;
          const float* scp = sGn + sl * GN_STRIDE + 64 + half * 160;
          const float* shp = scp + TB_C;
#pragma unroll 4
          for (int c8 = 0; c8 < 20; ++c8) {
            uint4* const p = reinterpret_cast<uint4*>(sA + a_chunk_off(row, half * 160 + c8 * 8));
            const uint4 xv = *p;
            const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w};
            const float4 s0 = *reinterpret_cast<const float4*>(scp + c8 * 8), s1 = *reinterpret_cast<const float4*>(scp + c8 * 8 + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(shp + c8 * 8), h1 = *reinterpret_cast<const float4*>(shp + c8 * 8 + 4);
            const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack_f16x2(xu[j]);
              o[j] = pack_bf16x2(fmaf(f.x, scv[2 * j], shv[2 * j]), fmaf(f.y, scv[2 * j + 1], shv[2 * j + 1]));
            }
            *p = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
}

// proj_out accumulator + bias + x_in -> fp16 tile in the A buffer (its MMAs have retired), GroupNorm partials of the output.
// (An out-of-line copy of this one measured 2.8 k cycles slower per tile: profiles/R2y_trace.txt.)
WD_DEVINL void out_tile_fn(const TBlockArgs& args, uint32_t t_row, uint8_t* sA, const float* sCb, int m0, int row, int half, int q,
                                         int lane) {
    const int c0 = half * 160;
    const bool valid = m0 + row < args.M;
    const __half* xr = args.x_in + static_cast<size_t>(valid ? m0 + row : 0) * args.x_in_ld + c0;
    float gs[32];  // 16 groups of 10 columns: [2 g] = sum, [2 g + 1] = sum of squares
#pragma unroll
    for (int i = 0; i < 32; ++i) gs[i] = 0.f;
    // the TMEM group and the x_in row piece of group g + 1 are in flight while group g is combined
    const float* bpo = sCb + 4 * TB_C + c0;
    uint32_t v[2][32];
    uint4 r4[2][4];
    tmem_ld_32x32b_x32(t_row + COL_X + c0, v[0]);
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) r4[0][c8] = __ldg(reinterpret_cast<const uint4*>(xr + c8 * 8));
#pragma unroll
    for (int g = 0; g < 5; ++g) {
      tmem_ld_wait();
      if (g + 1 < 5) {
        tmem_ld_32x32b_x32(t_row + COL_X + c0 + (g + 1) * 32, v[(g + 1) & 1]);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) r4[(g + 1) & 1][c8] = __ldg(reinterpret_cast<const uint4*>(xr + (g + 1) * 32 + c8 * 8));
      }
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const int cl = g * 32 + c8 * 8;  // column inside this thread's 160
        const float4 b0 = *reinterpret_cast<const float4*>(bpo + cl);
        const float4 b1 = *reinterpret_cast<const float4*>(bpo + cl + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint32_t ru[4] = {r4[g & 1][c8].x, r4[g & 1][c8].y, r4[g & 1][c8].z, r4[g & 1][c8].w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_f16x2(ru[j]);
          f[2 * j] = __uint_as_float(v[g & 1][c8 * 8 + 2 * j]) + bb[2 * j] + t.x;
          f[2 * j + 1] = __uint_as_float(v[g & 1][c8 * 8 + 2 * j + 1]) + bb[2 * j + 1] + t.y;
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
      const int smp_w = mw / args.HW;  // the sample of this warp's 32 rows
      const int slot = (mw % args.HW) >> 5, nslot = args.HW >> 5;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        float part[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) part[i] = gs[p * 16 + i];
        const float tot = warp_transpose_reduce16(part, lane);
        if (lane < 16 && mw < args.M) {
          const int g = half * 16 + p * 8 + (lane >> 1);
          args.gn_partial[((static_cast<size_t>(smp_w) * 32 + g) * nslot + slot) * 2 + (lane & 1)] = tot;
        }
      }
    }
}

// SPT = samples per 128-token tile: 1 (a sample holds a multiple of 128 tokens) or 2 (64 tokens per sample, the 4 x 16 level).
// With two samples the score GEMM runs against both samples' keys (N = 128, into the idle GEGLU accumulator columns), each row
// soft-maxes its own sample's 64 score columns and writes zeros for the other sample's keys, and the output GEMM reduces over
// both samples' (head, key) rows (K = 128).
template <bool PAIR, int SPT>
WD_DEVINL void tblock_body(const CUtensorMap& mapG, const CUtensorMap& mapWpi, const CUtensorMap& mapF0, const CUtensorMap& mapF1,
                           const CUtensorMap& mapF2, const CUtensorMap& mapF3, const CUtensorMap& mapW1, const CUtensorMap& mapW2,
                           const CUtensorMap& mapWpo, const CUtensorMap& mapOut, const TBlockArgs& args) {
  constexpr int TB_NSLOT = PAIR ? 9 : 4;
  constexpr int TB_SLOT = PAIR ? 10240 : 20480;
  static_assert(TB_NSLOT * TB_SLOT <= TB_RING, "ring");
  constexpr int TILE_M = PAIR ? 2 * TB_M : TB_M;   // tokens per tile (both CTAs of a pair)
  constexpr uint32_t N_EPI = PAIR ? 16 : 8;        // epilogue-warp arrivals on an epilogue -> MMA barrier
  extern __shared__ __align__(1024) uint8_t tb_smem[];
  uint8_t* const smem = tb_smem;
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* const sA = smem + OFF_A;
  uint8_t* const sG = smem + OFF_G;
  uint8_t* const sRing = smem + OFF_RING;
  float* const sBff = reinterpret_cast<float*>(smem + OFF_BFF);
  float* const sC = reinterpret_cast<float*>(smem + OFF_CSM);
  float* const sCb = reinterpret_cast<float*>(smem + OFF_CB);  // [4][320] cb ++ [320] b_po
  float2* const sStat = reinterpret_cast<float2*>(smem + OFF_STAT);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs of the pair)
  static_assert(!(PAIR && SPT != 1), "the pair build takes whole samples per CTA");
  const int m_tiles = (args.M + TILE_M - 1) / TILE_M;  // a ragged last tile (M % 128 == 64) is zero-filled / clipped by TMA
  const int n_samples = args.M / args.HW;
  constexpr uint32_t COL_SC = SPT == 2 ? COL_G : COL_S;  // score accumulator: N = 64 SPT columns
  const int stage = args.stage;
  const bool gn_in = args.gn_in_partial != nullptr;
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int nworkers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // a barrier the peer CTA signals too lives in the leader: its shared::cluster address (the local one in the single-CTA build)
  auto leader_bar = [&](int b) -> uint32_t { return PAIR ? mapa_shared(smem_u32(&bars[b]), 0) : smem_u32(&bars[b]); };

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
    mbar_init(&bars[B_A_READY], N_EPI);
    mbar_init(&bars[B_P_READY], N_EPI);
    mbar_init(&bars[B_GACC_FULL], 1);
    mbar_init(&bars[B_GACC_FREE], N_EPI);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_GBUF_FULL + i], N_EPI);
      mbar_init(&bars[B_GBUF_EMPTY + i], 1);
    }
    mbar_init(&bars[B_X_FREE], N_EPI);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  // static data: the folded GEGLU bias (weights only) -> shared memory
  for (int i = threadIdx.x; i < 2 * TB_HID / 4; i += TB_THREADS)
    reinterpret_cast<float4*>(sBff)[i] = __ldg(reinterpret_cast<const float4*>(args.b_ff) + i);
  for (int i = threadIdx.x; i < 4 * TB_C / 4; i += TB_THREADS)
    reinterpret_cast<float4*>(sCb)[i] = __ldg(reinterpret_cast<const float4*>(args.cb) + i);
  for (int i = threadIdx.x; i < TB_C / 4; i += TB_THREADS)
    reinterpret_cast<float4*>(sCb + 4 * TB_C)[i] = __ldg(reinterpret_cast<const float4*>(args.b_po) + i);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
      // `bytes` = what BOTH CTAs of a pair load into this slot (the leader's full barrier counts them all)
      auto acquire = [&](uint32_t bytes) -> uint8_t* {
        mbar_wait(&bars[B_RING_EMPTY + slot], phase ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&bars[B_RING_FULL + slot], bytes);
        return sRing + slot * TB_SLOT;
      };
      auto advance = [&]() {
        if (++slot == TB_NSLOT) { slot = 0; phase ^= 1; }
      };
      auto load2 = [&](void* dst, const CUtensorMap* mp, int bar, int c0, int c1) {
        if (PAIR) tma_load_2d_pair(dst, mp, leader_bar(bar), c0, c1);
        else tma_load_2d(dst, mp, &bars[bar], c0, c1);
      };
      auto load3 = [&](void* dst, const CUtensorMap* mp, int bar, int c0, int c1, int c2) {
        if (PAIR) tma_load_3d_pair(dst, mp, leader_bar(bar), c0, c1, c2);
        else tma_load_3d(dst, mp, &bars[bar], c0, c1, c2);
      };
      // one N half of a [320 x 320] weight: 5 K blocks of [160 x 64] (pair: this CTA's 80 of the 160 rows)
      auto weight_320_units = [&](const CUtensorMap* mp, int nh, int kb0, int kb1) {
        for (int kb = kb0; kb < kb1; ++kb) {
          uint8_t* dst = acquire(160 * 128);
          load2(dst, mp, B_RING_FULL + slot, kb * 64, nh * 160 + (PAIR ? static_cast<int>(rank) * 80 : 0));
          advance();
        }
      };
      auto weight_320_half = [&](const CUtensorMap* mp, int nh) { weight_320_units(mp, nh, 0, TB_KB); };
      // score operand of one attention: five K blocks of four [16 keys x 64] boxes, one per head (pair: this CTA's two heads)
      auto fold_m_units = [&](const CUtensorMap* mp, int sample) {
        for (int b = 0; b < TB_KB; ++b) {
          uint8_t* dst = acquire(SPT * TB_HEADS * TB_KEYS * 128);
          if (PAIR) {
            for (int hl = 0; hl < 2; ++hl)
              load3(dst + hl * (TB_KEYS * 128), mp, B_RING_FULL + slot, (static_cast<int>(rank) * 2 + hl) * TB_C + b * 64, 0, sample);
          } else {
            for (int sl = 0; sl < SPT; ++sl)  // a sample index past the batch (ragged last tile) is zero-filled
              for (int h = 0; h < TB_HEADS; ++h)
                load3(dst + (sl * TB_HEADS + h) * (TB_KEYS * 128), mp, B_RING_FULL + slot, h * TB_C + b * 64, 0, sample + sl);
          }
          advance();
        }
      };
      // output operand of one attention (MN-major: rows = (head, key), 64 output columns per row).  Single CTA: five 64-column
      // units.  Pair: three N = 128 MMAs; this CTA supplies column block 2 nb + rank (the peer has none for the last one: its
      // half of that MMA lands in TMEM columns [320, 384), which belong to the idle GEGLU accumulator)
      auto fold_n_units = [&](const CUtensorMap* mp, int sample) {
        if (PAIR) {
          for (int nb = 0; nb < 3; ++nb) {
            const int blk = 2 * nb + static_cast<int>(rank);
            uint8_t* dst = acquire((nb < 2 ? 2 : 1) * TB_HEADS * TB_KEYS * 128);
            if (blk < TB_KB)
              for (int h = 0; h < TB_HEADS; ++h) load3(dst + h * (TB_KEYS * 128), mp, B_RING_FULL + slot, h * TB_C + blk * 64, 0, sample);
            advance();
          }
        } else {
          for (int b = 0; b < TB_KB; ++b) {
            uint8_t* dst = acquire(SPT * TB_HEADS * TB_KEYS * 128);
            for (int sl = 0; sl < SPT; ++sl)
              for (int h = 0; h < TB_HEADS; ++h)
                load3(dst + (sl * TB_HEADS + h) * (TB_KEYS * 128), mp, B_RING_FULL + slot, h * TB_C + b * 64, 0, sample + sl);
            advance();
          }
        }
      };
      auto w1_units = [&](int c) {  // pair: rank 0 holds the chunk's 64 value rows, rank 1 its 64 gate rows
        for (int kb = 0; kb < TB_KB; ++kb) {
          uint8_t* dst = acquire(2 * TB_CHUNK * 128);
          load2(dst, &mapW1, B_RING_FULL + slot, kb * 64, c * 2 * TB_CHUNK + (PAIR ? static_cast<int>(rank) * TB_CHUNK : 0));
          advance();
        }
      };
      auto w2_units = [&](int c) {
        for (int nh = 0; nh < 2; ++nh) {
          uint8_t* dst = acquire(160 * 128);
          load2(dst, &mapW2, B_RING_FULL + slot, c * TB_CHUNK, nh * 160 + (PAIR ? static_cast<int>(rank) * 80 : 0));
          advance();
        }
      };
      int it = 0;
      for (int tile = worker; tile < m_tiles; tile += nworkers, ++it) {
        const int m0 = tile * TILE_M + static_cast<int>(rank) * TB_M;  // this CTA's 128 rows
        const int sample = (tile * TILE_M) / args.HW;
        // the first weight units fill the ring while the previous tile finishes -- no more than the ring holds: issuing more
        // before the g operand could block on a slot that only this tile's MMAs (which wait for g) can free
        constexpr int PRE = TB_NSLOT < TB_KB ? TB_NSLOT : TB_KB;
        weight_320_units(&mapWpi, 0, 0, PRE);
        // the A buffer is free once the previous tile's output store has read it
        mbar_wait(&bars[B_A_FREE], (it & 1) ^ 1);
        if (gn_in) {  // the epilogue warps of THIS CTA normalise the tile first: its own barrier, its own 80 KB
          mbar_arrive_expect_tx(&bars[B_A_FULL], TB_KB * TB_ABLK);
          for (int kb = 0; kb < TB_KB; ++kb) {
            if (PAIR) tma_load_2d_pair(sA + kb * TB_ABLK, &mapG, mapa_shared(smem_u32(&bars[B_A_FULL]), rank), kb * 64, m0);
            else tma_load_2d(sA + kb * TB_ABLK, &mapG, &bars[B_A_FULL], kb * 64, m0);
          }
        } else {
          if (rank == 0) mbar_arrive_expect_tx(&bars[B_A_FULL], (PAIR ? 2 : 1) * TB_KB * TB_ABLK);
          for (int kb = 0; kb < TB_KB; ++kb) load2(sA + kb * TB_ABLK, &mapG, B_A_FULL, kb * 64, m0);
        }
        weight_320_units(&mapWpi, 0, PRE, TB_KB);
        weight_320_half(&mapWpi, 1);
        if (stop_after(stage, 1)) continue;
        fold_m_units(&mapF0, sample);
        fold_n_units(&mapF1, sample);
        if (stop_after(stage, 2)) continue;
        fold_m_units(&mapF2, sample);
        fold_n_units(&mapF3, sample);
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
    if (rank == 0 && elect_one()) {
      constexpr uint32_t ID_BF16_160 = make_idesc_bf16_f32(TILE_M, 160);
      constexpr uint32_t ID_F16_160 = make_idesc_f16_f32(TILE_M, 160);
      constexpr uint32_t ID_F16_128 = make_idesc_f16_f32(TILE_M, 128);
      constexpr uint32_t ID_F16_64 = make_idesc_f16_f32(TILE_M, 64);
      constexpr uint32_t ID_F16_64_MN = ID_F16_64 | (1u << 16);    // B operand MN-major
      constexpr uint32_t ID_F16_128_MN = ID_F16_128 | (1u << 16);
      auto mma = [&](uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
        if (PAIR) umma_f16_ss_pair(d, a_desc, b_desc, idesc, acc);
        else umma_f16_ss(d, a_desc, b_desc, idesc, acc);
      };
      auto commit = [&](int b) {  // pair: arrives on the barrier at this offset in BOTH CTAs
        if (PAIR) umma_commit_pair(&bars[b]);
        else umma_commit(&bars[b]);
      };
      auto wait_epi = [&](int b, uint32_t parity) {  // a barrier the epilogue warps (of both CTAs) arrive on
        mbar_wait(&bars[b], parity);  // CTA-scope acquire also for the peer's arrivals (see mbar_arrive_remote)
      };
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
        commit(B_RING_EMPTY + slot);
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
              mma(tmem_base + COL_X + nh * 160, a_desc + 2 * k, b_desc + 2 * k, idesc, (!fresh || (kb | k) != 0) ? 1u : 0u);
            ring_release();
          }
      };
      auto wait_a_ready = [&]() {
        wait_epi(B_A_READY, n_a_ready & 1);
        ++n_a_ready;
        tc_fence_after();
      };
      int it = 0;
      for (int tile = worker; tile < m_tiles; tile += nworkers, ++it) {
        // ---- proj_in: X = g Wpi^T ----
        TB_STAMP(0, it, 0);
        if (gn_in) wait_a_ready();  // the tile has landed AND the epilogue warps (of both CTAs) have normalised it in place
        else mbar_wait(&bars[B_A_FULL], it & 1);
        wait_epi(B_X_FREE, (it & 1) ^ 1);  // the previous tile's epilogue has read its last accumulator
        tc_fence_after();
        TB_STAMP(0, it, 1);
        gemm_320(args.mid ? ID_F16_160 : ID_BF16_160, true);  // mid: x (fp16) times the identity = the residual stream itself
        commit(B_ACC);
        TB_STAMP(0, it, 2);
        if (stop_after(stage, 1)) continue;
        // ---- two cross-attentions: S = xhat M^T ; X += P N^T ----
        for (int a = 0; a < 2; ++a) {
          wait_a_ready();
          TB_STAMP(0, it, 3 + 4 * a);
          for (int kb = 0; kb < TB_KB; ++kb) {
            const uint64_t b_desc = make_smem_desc_sw128(ring_wait());
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + kb * TB_ABLK));
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(tmem_base + COL_SC, a_desc + 2 * k, b_desc + 2 * k, SPT == 2 ? ID_F16_128 : ID_F16_64, (kb | k) != 0);
            ring_release();
          }
          commit(B_S);
          TB_STAMP(0, it, 4 + 4 * a);
          wait_epi(B_P_READY, n_p_ready & 1);
          ++n_p_ready;
          tc_fence_after();
          TB_STAMP(0, it, 5 + 4 * a);
          const uint64_t p_desc = make_smem_desc_sw128(smem_u32(sG));
          for (int nb = 0; nb < (PAIR ? 3 : TB_KB); ++nb) {
            const uint64_t b_desc = desc_mn_sw128(ring_wait());
#pragma unroll
            for (int k = 0; k < 4 * SPT; ++k)  // K step k = the 16 key slots of (sample k / 4, head k % 4): rows [16 k, +16) of the MN-major unit
              mma(tmem_base + COL_X + nb * (PAIR ? 128 : 64), p_desc + (k >> 2) * (TB_ABLK >> 4) + 2 * (k & 3), b_desc + 128 * k,
                  PAIR ? ID_F16_128_MN : ID_F16_64_MN, 1u);
            ring_release();
          }
          commit(B_ACC);
          TB_STAMP(0, it, 6 + 4 * a);
          if (stop_after(stage, 2 + a)) break;
        }
        if (stop_after(stage, 2) || stop_after(stage, 3)) continue;
        // ---- feed-forward: per 64-column hidden chunk  Gacc = xhat W1_c^T ; X += GEGLU(Gacc) W2_c^T ----
        wait_a_ready();
        TB_STAMP(0, it, 11);
        auto mma2 = [&](int c) {
          const int b = c & 1;
          wait_epi(B_GBUF_FULL + b, n_gbuf[b] & 1);
          ++n_gbuf[b];
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sG + b * TB_ABLK));
          for (int nh = 0; nh < 2; ++nh) {
            const uint64_t b_desc = make_smem_desc_sw128(ring_wait());
            // (handing the chunk over as a TMEM-resident A operand -- tcgen05.st by the epilogue, tcgen05.mma [d], [a_tmem], b_desc --
            //  was correct on its first run and measured SLOWER: 0.896 vs 0.846 ms for the four launches, profiles/R2v_*)
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(tmem_base + COL_X + nh * 160, a_desc + 2 * k, b_desc + 2 * k, ID_BF16_160, 1u);
            ring_release();
          }
          commit(B_GBUF_EMPTY + b);
        };
        for (int c = 0; c < TB_NCHUNK; ++c) {
          if (n_gchunk > 0) {  // the epilogue has drained the previous chunk's accumulator
            wait_epi(B_GACC_FREE, (n_gchunk - 1) & 1);
            tc_fence_after();
          }
          ++n_gchunk;
          for (int kb = 0; kb < TB_KB; ++kb) {
            const uint64_t b_desc = make_smem_desc_sw128(ring_wait());
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + kb * TB_ABLK));
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(tmem_base + COL_G, a_desc + 2 * k, b_desc + 2 * k, ID_F16_128, (kb | k) != 0);
            ring_release();
          }
          commit(B_GACC_FULL);
          if (c > 0) mma2(c - 1);
          if (c < 8) TB_STAMP(0, it, 16 + c);  // issue progress of the first chunks
        }
        mma2(TB_NCHUNK - 1);
        commit(B_ACC);
        TB_STAMP(0, it, 12);
        if (stop_after(stage, 4)) continue;
        // ---- proj_out: X = x3 Wpo^T (fresh accumulator; the epilogue adds bias and x_in) ----
        wait_a_ready();
        TB_STAMP(0, it, 13);
        gemm_320(ID_F16_160, true);
        commit(B_ACC);
        TB_STAMP(0, it, 14);
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
    auto arrive_warp = [&](int bar) {  // epilogue -> MMA barrier (in the leader CTA)
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_remote(leader_bar(bar));
        else mbar_arrive(&bars[bar]);
      }
    };
    // X + cb -> 16-bit operand copy in the A buffer.  NORM: (x - mean) * rstd as fp16 (LayerNorm without its affine part, which
    // is folded into the weights that consume the copy); else the raw value as fp16.
    auto x_to_a = [&](const float* cb, bool norm) { x_to_a_fn(t_row, cb, norm, sA, sStat, row, half, q, args.ln_eps); };
    // scores (+ per-sample constants) -> softmax over the L keys of each head -> fp16 probabilities, K-major P tile
    auto softmax_to_p = [&](const float* cs) {
      mbar_wait(&bars[B_S], n_s & 1);
      ++n_s;
      tc_fence_after();
      const int sel = SPT == 2 ? (q >> 1) : 0;  // which of the tile's samples this row belongs to (64 rows each)
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + COL_SC + sel * 64 + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      uint32_t pk[16];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float sc[16];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sc[j] = (j < args.L) ? __uint_as_float(v[hh * 16 + j]) + cs[sel * 64 + half * 32 + hh * 16 + j] : -INFINITY;
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
      uint8_t* const prow = sG + sel * TB_ABLK + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (row & 7)) << 4)) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      if (SPT == 2) {  // this row takes nothing from the other sample's keys
        uint8_t* const zrow = sG + (sel ^ 1) * TB_ABLK + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(zrow + (((half * 4 + c) ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
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
    for (int tile = worker; tile < m_tiles; tile += nworkers, ++it) {
      const int m0 = tile * TILE_M + static_cast<int>(rank) * TB_M;
      const int sample = (tile * TILE_M) / args.HW;
      // per-sample score constants of both attentions -> shared memory, slot (h, j) = h * 16 + j
      if (et < 128 * SPT) {  // layout [attention][sample of the tile][head][key]
        const int a = et / (64 * SPT), sl = (et / 64) % SPT, n = et & 63, h = n >> 4, j = n & 15;
        const float* cv = a ? args.cvec2 : args.cvec1;
        sC[et] = (j < args.L && sample + sl < n_samples) ? __ldg(cv + (static_cast<size_t>(sample + sl) * args.L + j) * args.cvec_ld + h) : 0.f;
      }
      named_barrier_sync(6, 256);

      if (gn_in) {
        gn_in_fn<SPT>(args, sA, sG, &bars[B_A_FULL], it & 1, sample, n_samples, et, row, half, q);
        fence_proxy_async();
        named_barrier_sync(6, 256);  // sGn (in the P / GEGLU buffer) has been read by everybody
        arrive_warp(B_A_READY);
      }

      // ---- after proj_in ----
      const bool tr = warp == 2 && lane == 0;
      if (tr) TB_STAMP(1, it, 0);
      wait_acc();
      if (tr) TB_STAMP(1, it, 1);
      x_to_a(sCb, true);
      if (tr) TB_STAMP(1, it, 2);
      if (stop_after(stage, 1)) { store_a_tile(m0); continue; }
      arrive_warp(B_A_READY);
      // ---- attention 1 / 2 ----
      bool stopped = false;
      for (int a = 0; a < 2; ++a) {
        softmax_to_p(sC + a * 64 * SPT);
        arrive_warp(B_P_READY);
        if (tr) TB_STAMP(1, it, 3 + 3 * a);
        wait_acc();
        if (tr) TB_STAMP(1, it, 4 + 3 * a);
        x_to_a(sCb + (1 + a) * TB_C, true);
        if (tr) TB_STAMP(1, it, 5 + 3 * a);
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
        if (tr && c < 8) TB_STAMP(1, it, 16 + 2 * c);
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
        if (tr && c < 8) TB_STAMP(1, it, 17 + 2 * c);
      }
      if (tr) TB_STAMP(1, it, 9);
      // the last two chunks' MMAs are covered by the accumulator barrier below; account for their buffer releases
      n_gbuf_empty[0] += 1;
      n_gbuf_empty[1] += 1;
      // ---- x3 (raw) -> operand of proj_out ----
      wait_acc();
      if (tr) TB_STAMP(1, it, 10);
      x_to_a(sCb + 3 * TB_C, false);
      if (tr) TB_STAMP(1, it, 11);
      if (stop_after(stage, 4)) { store_a_tile(m0); continue; }
      arrive_warp(B_A_READY);
      // ---- proj_out accumulator + bias + x_in -> fp16 tile in the A buffer (its MMAs have retired), GroupNorm partials ----
      wait_acc();
      if (tr) TB_STAMP(1, it, 12);
      out_tile_fn(args, t_row, sA, sCb, m0, row, half, q, lane);
      if (tr) TB_STAMP(1, it, 13);
      store_a_tile(m0);
      if (tr) TB_STAMP(1, it, 14);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while its peer may still read its smem / signal its barriers
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

__global__ void __launch_bounds__(TB_THREADS, 1)
tblock_unet_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapWpi,
                   const __grid_constant__ CUtensorMap mapF0, const __grid_constant__ CUtensorMap mapF1,
                   const __grid_constant__ CUtensorMap mapF2, const __grid_constant__ CUtensorMap mapF3,
                   const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                   const __grid_constant__ CUtensorMap mapWpo, const __grid_constant__ CUtensorMap mapOut, const TBlockArgs args) {
  tblock_body<false, 1>(mapG, mapWpi, mapF0, mapF1, mapF2, mapF3, mapW1, mapW2, mapWpo, mapOut, args);
}
__global__ void __launch_bounds__(TB_THREADS, 1)
tblock_unet_spt2_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapWpi,
                        const __grid_constant__ CUtensorMap mapF0, const __grid_constant__ CUtensorMap mapF1,
                        const __grid_constant__ CUtensorMap mapF2, const __grid_constant__ CUtensorMap mapF3,
                        const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                        const __grid_constant__ CUtensorMap mapWpo, const __grid_constant__ CUtensorMap mapOut, const TBlockArgs args) {
  tblock_body<false, 2>(mapG, mapWpi, mapF0, mapF1, mapF2, mapF3, mapW1, mapW2, mapWpo, mapOut, args);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TB_THREADS, 1)
tblock_unet_pair_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapWpi,
                        const __grid_constant__ CUtensorMap mapF0, const __grid_constant__ CUtensorMap mapF1,
                        const __grid_constant__ CUtensorMap mapF2, const __grid_constant__ CUtensorMap mapF3,
                        const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                        const __grid_constant__ CUtensorMap mapWpo, const __grid_constant__ CUtensorMap mapOut, const TBlockArgs args) {
  tblock_body<true, 1>(mapG, mapWpi, mapF0, mapF1, mapF2, mapF3, mapW1, mapW2, mapWpo, mapOut, args);
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

}  // namespace wd
// debug aid (not part of the product ABI): the phase stamps recorded under WD_TBLOCK_TRACE=1
extern "C" int wdx_tblock_trace_read(unsigned long long* host, int clear) {
  if (cudaMemcpyFromSymbol(host, wd::g_tb_trace, sizeof(unsigned long long) * 2 * 8 * 64) != cudaSuccess) return -1;
  if (clear) {
    static unsigned long long z[2 * 8 * 64];
    if (cudaMemcpyToSymbol(wd::g_tb_trace, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
namespace wd {

static bool tblock_trace_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_TBLOCK_TRACE");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}

bool tblock_gn_fused() {  // env WD_TBLOCK_GN (default on): the block's input GroupNorm runs inside the kernel
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_TBLOCK_GN");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

bool tblock_use_pair(int HW) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_TBLOCK_PAIR");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0 && HW % (2 * TB_M) == 0;
}

cudaError_t tblock_launch(const TBlockLaunch& L, int num_sms, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tblock_unet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(tblock_unet_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(tblock_unet_spt2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM);
  });
  if (attr_err != cudaSuccess) return attr_err;
  TBlockArgs a = L.args;
  a.trace = tblock_trace_enabled() ? 1 : 0;
  const bool spt2 = a.HW == TB_M / 2;  // two samples per tile (4 x 16 latents); a ragged last tile is allowed there
  if (a.M <= 0 || a.M % a.HW || (!spt2 && a.HW % TB_M) || a.L < 1 || a.L > TB_KEYS || a.x_in_ld % 8 || (spt2 && a.pair))
    return cudaErrorInvalidValue;
  if (a.mid && a.stage != 4) return cudaErrorInvalidValue;  // the mid form ends with the raw residual stream (no proj_out)
  if (spt2) {
    const int tiles = (a.M + TB_M - 1) / TB_M;
    const int grid = tiles < num_sms ? tiles : num_sms;
    return launch_pdl(tblock_unet_spt2_kernel, dim3(grid), dim3(TB_THREADS), TB_SMEM, stream, L.mapG, L.mapWpi, L.mapF[0], L.mapF[1],
                      L.mapF[2], L.mapF[3], L.mapW1, L.mapW2, L.mapWpo, L.mapOut, a);
  }
  if (a.pair) {  // the tensor maps of the weights were encoded with the half-unit boxes (tblock_use_pair)
    if (a.HW % (2 * TB_M)) return cudaErrorInvalidValue;
    const int tiles = a.M / (2 * TB_M);
    const int max_pairs = num_sms / 2;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    return launch_pdl(tblock_unet_pair_kernel, dim3(2 * pairs), dim3(TB_THREADS), TB_SMEM, stream, L.mapG, L.mapWpi, L.mapF[0],
                      L.mapF[1], L.mapF[2], L.mapF[3], L.mapW1, L.mapW2, L.mapWpo, L.mapOut, a);
  }
  const int tiles = a.M / TB_M;
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_pdl(tblock_unet_kernel, dim3(grid), dim3(TB_THREADS), TB_SMEM, stream, L.mapG, L.mapWpi, L.mapF[0], L.mapF[1],
                    L.mapF[2], L.mapF[3], L.mapW1, L.mapW2, L.mapWpo, L.mapOut, a);
}

}  // namespace wd
