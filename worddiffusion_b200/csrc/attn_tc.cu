// Fused attention on the 5th-gen tensor cores (tcgen05 + TMEM) for d_head = 80: softmax(q k^T * d^-0.5) v
// (reference unetPhosc.py:176-196).  Used for the 256 x 256 / 64 x 64 self-attention of UNetModelPhosc and for its
// cross-attention over the 779-token char + PHOSC context.  The score matrix never reaches HBM (the reference materialises
// [B*4, Sq, Skv] fp32).  The mma.sync flash kernel (attn_flash.cu) ran these launches at 96 TFLOP/s, 45 % of the unetPhosc step.
//
// One CTA = (MT x 128 query rows, head, sample).  MT = 1 (default): 4 ring slots, 112 KB of shared memory, two CTAs per SM, 192
// threads.  MT = 2 (env WD_ATTN_TC_MT2=1; both query tiles of a 256-token sample share every K / V tile, 8 slots, 224 KB, one CTA
// per SM, 320 threads) halves the L2 -> SM operand traffic but measured slower: a lone CTA cannot hide its softmax <-> MMA chain.
//   warp 0     TMA producer: Q once, then K(0), V(0), K(1), V(1), ... tiles of 64 keys as UNITS through a ring of 16 KB slots
//              (a K slot is released when its QK^T retires, a V slot after its PV).  Every operand tile is a pair of
//              SWIZZLE_128B boxes of 64 channels starting at the head's first channel: the second box over-fetches 48 channels of
//              the next head (zero-filled past the tensor's last column), only its first 16 are used.
//   warp 1     TMEM allocation + single-thread MMA issue.
//                S = Q K^T : 128 x 64 x 16, five K steps (four in box 0, one in box 1), both operands K-major.
//                O += P V  : 128 x 80 x 16, four K steps of 16 keys; P is K-major (written by the softmax warps), V is consumed
//                            as an MN-major operand straight from its [key][channel] boxes (descriptor LBO = box pitch).
//   warps 2..  softmax (four warps per M tile), one query row per thread (= TMEM lane).  ONE pass over the keys with a LAZY
//              running maximum: P = exp2(scale log2e (S - m)) where m only moves when a tile's maximum exceeds it by more than
//              2^8 in the exponent (P stays <= 256, exact in bf16 range; O / l does not depend on m).  When it moves, the warp
//              rescales its own 32 rows of O in TMEM (tcgen05.ld / st) between the retirement of PV(j-1) and its P(j) arrival
//              -- with bounded scores that is the first tile or two.  Round 1 ran TWO passes (row maxima first, then a second
//              QK^T and a second read of K): 1.5x the MMAs, 1.5x the K / V operand bytes (the binding resource: L2 -> SM
//              ingest), and ~1.7x the softmax-warp instructions of this form.
//              S is double-buffered in TMEM and handed back as soon as the scores are in registers, so QK^T of tile j+1 / j+2
//              overlaps the exponentials of tile j.
// Keys beyond Skv are zero-filled by TMA (3-D maps: channel, row, sample) and masked to -inf; query rows beyond Sq are
// computed on zero-filled operands and not stored.
#include "ops.cuh"

#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace wd {

bool tmap_encode_3d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t row_stride_elems,
                         uint64_t batch_stride_elems, uint32_t box_inner, uint32_t box_rows);  // gemm_tc.cu

namespace {
constexpr int AT_DH = 80;
constexpr int AT_BM = 128;  // query rows per M tile
constexpr int AT_BK = 64;   // keys per tile
constexpr int AT_KV_BOX = AT_BK * 128;   // one 64-channel box of 64 keys (8 KB)
constexpr int AT_SLOT_BYTES = 2 * AT_KV_BOX;  // a K tile or a V tile: two boxes (16 KB)
constexpr int AT_QT_BYTES = 2 * AT_BM * 128;  // Q of one M tile: two 64-channel boxes (32 KB)
constexpr int AT_PT_BYTES = AT_BM * 128;      // P of one M tile: [128 rows][64 keys] bf16, K-major SWIZZLE_128B (16 KB)

// MT = M tiles (of 128 query rows) per CTA.  MT = 1: 4 ring slots, 112 KB, two CTAs per SM (64-row levels, odd shapes).
// MT = 2: both query tiles of a 256-token sample share every K / V tile (half the L2 -> SM operand traffic, which is what
// bounds this kernel), 8 ring slots, 224 KB, one CTA per SM.
template <int MT>
struct ATCfg {
  static constexpr int SLOTS = MT == 1 ? 4 : 8;
  static constexpr int THREADS = 64 + 128 * MT;
  static constexpr int SMEM_BYTES = MT * AT_QT_BYTES + SLOTS * AT_SLOT_BYTES + MT * AT_PT_BYTES + 256;
  static constexpr int TMEM_COLS = MT == 1 ? 256 : 512;  // S[mt][buf] at (2 mt + buf) * 64, O[mt] at 2 MT * 64 + 80 mt
  static constexpr int O_COL = 2 * MT * AT_BK;
};

struct AttnTcArgs {
  __nv_bfloat16* out;
  int out_ld;
  int Sq, Skv, heads, batch;
  float sl2;  // scale * log2(e)
};

// MN-major SWIZZLE_128B operand made of [64 channel x 64 row] boxes (same encoding as wgrad_tc.cu): LBO = box pitch
WD_DEVINL uint64_t at_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(AT_KV_BOX >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

WD_DEVINL float at_max3(float a, float b, float c) {  // FMNMX3
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

WD_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Operand tiles travel through a ring of 16 KB slots as UNITS: K(0), V(0), K(1), V(1), ...
// Unit w lives in slot w % SLOTS (parity (w / SLOTS) & 1).  A K slot is released as soon as its QK^T has retired, a V slot after
// its PV -- SLOTS / 2 key tiles are in flight.
template <int MT>
__global__ void __launch_bounds__(ATCfg<MT>::THREADS, MT == 1 ? 2 : 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapV, const AttnTcArgs a) {
  using C = ATCfg<MT>;
  constexpr int NS = C::SLOTS;
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw;
  if (smem_u32(smem) & 1023) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment (no static shared memory in this kernel)
  uint8_t* sQ = smem;
  uint8_t* sRing = sQ + MT * AT_QT_BYTES;
  uint8_t* sP = sRing + NS * AT_SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + MT * AT_PT_BYTES);
  uint64_t* u_full = bars;             // [NS]
  uint64_t* u_empty = bars + NS;       // [NS]
  uint64_t* s_full = bars + 2 * NS;    // [2]
  uint64_t* s_empty = s_full + 2;      // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* q_full = p_empty + 1;
  uint64_t* o_full = q_full + 1;
  uint64_t* q_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (a.Skv + AT_BK - 1) / AT_BK;
  // PERSISTENT: the CTA walks work items (query tile, head, sample) = blockIdx.x, + gridDim.x, ... (query tile fastest: concurrent
  // CTAs share a (head, sample)'s K / V through L2).  Every barrier parity runs on counters that continue across items, so the
  // producer loads the next item's Q (as soon as the last QK^T of the current one has retired: q_empty) and K / V tiles, and the
  // MMA thread issues its first QK^T, while the softmax warps are still in the current item's last tiles and output epilogue --
  // a non-persistent CTA spent ~4 us on start-up and tear-down, half the time of a 256-key self-attention item.
  const int q_tiles = (a.Sq + AT_BM * MT - 1) / (AT_BM * MT);
  const int total_items = q_tiles * a.heads * a.batch;
  auto item_coords = [&](int item, int* q0_, int* c_head_, int* b_) {
    const int qt = item % q_tiles, hb = item / q_tiles;
    *q0_ = qt * (AT_BM * MT);
    *c_head_ = (hb % a.heads) * AT_DH;
    *b_ = hb / a.heads;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&u_full[i], 1);
      mbar_init(&u_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4 * MT);
    }
    mbar_init(p_full, 4 * MT);
    mbar_init(p_empty, 1);
    mbar_init(q_full, 1);
    mbar_init(o_full, 1);
    mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      int wg = 0;  // ring unit counter over all items
      int it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        int q0, c_head, b;
        item_coords(item, &q0, &c_head, &b);
        mbar_wait(q_empty, (it & 1) ^ 1);  // the previous item's last QK^T has read sQ (passes at once for the first item)
        mbar_arrive_expect_tx(q_full, MT * AT_QT_BYTES);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          tma_load_3d(sQ + mt * AT_QT_BYTES, &mapQ, q_full, c_head, q0 + mt * AT_BM, b);
          tma_load_3d(sQ + mt * AT_QT_BYTES + AT_BM * 128, &mapQ, q_full, c_head + 64, q0 + mt * AT_BM, b);
        }
        const int units = 2 * ntiles;  // K(0), V(0), K(1), V(1), ...
        for (int w = 0; w < units; ++w, ++wg) {
          const int sl = wg % NS;
          const bool is_v = (w & 1) != 0;
          const int j = w >> 1;
          mbar_wait(&u_empty[sl], ((wg / NS) & 1) ^ 1);
          uint8_t* dst = sRing + sl * AT_SLOT_BYTES;
          mbar_arrive_expect_tx(&u_full[sl], AT_SLOT_BYTES);
          const CUtensorMap* mp = is_v ? &mapV : &mapK;
          tma_load_3d(dst, mp, &u_full[sl], c_head, j * AT_BK, b);
          tma_load_3d(dst + AT_KV_BOX, mp, &u_full[sl], c_head + 64, j * AT_BK, b);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(AT_BM, AT_BK);
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32(AT_BM, AT_DH) | (1u << 16);  // B (= V) is MN-major
      int jg0 = 0;  // key-tile counter over all items (score buffer, P tile and ring parities)
      int it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it, jg0 += ntiles) {
        mbar_wait(q_full, it & 1);
        tc_fence_after();
        // S(j) = Q K(j)^T into score buffer (jg0 + j) & 1 from ring unit 2 (jg0 + j)
        auto issue_qk = [&](int j) {
          const int jg = jg0 + j;
          const int buf = jg & 1, w = 2 * jg, sl = w % NS;
          mbar_wait(&u_full[sl], (w / NS) & 1);
          mbar_wait(&s_empty[buf], ((jg >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sRing + sl * AT_SLOT_BYTES);
          const uint64_t k_desc0 = make_smem_desc_sw128(k_addr);
          const uint64_t k_desc1 = make_smem_desc_sw128(k_addr + AT_KV_BOX);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ + mt * AT_QT_BYTES));
            const uint64_t q_desc1 = make_smem_desc_sw128(smem_u32(sQ + mt * AT_QT_BYTES + AT_BM * 128));
            const uint32_t d_tmem = tmem_base + (2 * mt + buf) * AT_BK;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(d_tmem, q_desc0 + 2 * k, k_desc0 + 2 * k, idesc_qk, k != 0);
            umma_f16_ss(d_tmem, q_desc1, k_desc1, idesc_qk, 1u);
          }
          umma_commit(&s_full[buf]);
          umma_commit(&u_empty[sl]);  // the K tile is free once these MMAs retire
          if (j == ntiles - 1) umma_commit(q_empty);  // ... and so is sQ after the item's last QK^T
        };
        // S(j+1) is issued ahead of P(j) V(j)
        issue_qk(0);
        for (int j = 0; j < ntiles; ++j) {
          if (j + 1 < ntiles) issue_qk(j + 1);
          const int jg = jg0 + j;
          const int wv = 2 * jg + 1, sl = wv % NS;
          mbar_wait(&u_full[sl], (wv / NS) & 1);
          mbar_wait(p_full, jg & 1);  // P(j) is written -- and the softmax warps are done rescaling O / reading the previous item's O
          tc_fence_after();
          const uint64_t v_desc = at_desc_mn_sw128(smem_u32(sRing + sl * AT_SLOT_BYTES));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t p_desc = make_smem_desc_sw128(smem_u32(sP + mt * AT_PT_BYTES));
#pragma unroll
            for (int k = 0; k < AT_BK / 16; ++k)
              umma_f16_ss(tmem_base + C::O_COL + mt * AT_DH, p_desc + 2 * k, v_desc + 128 * k, idesc_pv, (j | k) != 0);
          }
          umma_commit(&u_empty[sl]);
          umma_commit(p_empty);
        }
        umma_commit(o_full);
      }
    }
  } else {
    // =========================== softmax warps (one query row per thread) ===========================
    const int mt = (warp - 2) >> 2;  // M tile of this warp
    const int qd = warp & 3;         // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(qd * 32) << 16;
    // ---- one pass over the keys: P = exp2(sl2 (S - m)) with a LAZY running maximum m.  m only moves when a tile's maximum
    // exceeds it by more than 2^8 in the exponent; then this warp rescales its 32 rows of O in TMEM (and l) before P(j) is
    // published -- PV(j-1) has retired by then (p_empty) and PV(j) is not issued before every warp's P(j) arrival.  With
    // bounded score ranges that happens on the first tiles only, and O / l is exact whatever m was used. ----
    uint8_t* const prow = sP + mt * AT_PT_BYTES + row * 128;
    const uint32_t o_addr = tmem_base + t_lane + C::O_COL + mt * AT_DH;
    const bool ragged = (a.Skv % AT_BK) != 0;
    int jg0 = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it, jg0 += ntiles) {
    int q0, c_head, b;
    item_coords(item, &q0, &c_head, &b);
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int jg = jg0 + j;
      const int buf = jg & 1;
      mbar_wait(&s_full[buf], (jg >> 1) & 1);
      tc_fence_after();
      uint32_t v[64];
      {
        uint32_t (&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(v);
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(v + 32);
        tmem_ld_32x32b_x32(tmem_base + t_lane + (2 * mt + buf) * AT_BK, v0);
        tmem_ld_32x32b_x32(tmem_base + t_lane + (2 * mt + buf) * AT_BK + 32, v1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[buf]);  // the scores are in registers: S(j+2) may overwrite the buffer
      const bool mask = ragged && j == ntiles - 1;
      const int nvalid = a.Skv - j * AT_BK;  // valid keys of this tile (only read when mask)
      float tm = -INFINITY;
      if (!mask) {
        // four independent chains of 3-input maxima (a single fmaxf chain is 64 dependent instructions deep)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
#pragma unroll
          for (int i = 0; i < 4; ++i) m4[i] = at_max3(m4[i], __uint_as_float(v[c + 2 * i]), __uint_as_float(v[c + 2 * i + 1]));
        }
        tm = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c < nvalid) tm = fmaxf(tm, __uint_as_float(v[c]));
      }
      bool waited = false;
      if (j == 0) {
        m_used = tm;
      } else {
        const bool need = (tm - m_used) * a.sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(p_empty, (jg - 1) & 1);  // PV(j-1) has retired: O holds tiles 0 .. j-1
          waited = true;
          tc_fence_after();
          float f = 1.0f;
          if (need) {
            f = exp2f((m_used - tm) * a.sl2);
            m_used = tm;
          }
#pragma unroll
          for (int cb = 0; cb < AT_DH / 16; ++cb) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(o_addr + cb * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32b_x16(o_addr + cb * 16, o);
          }
          tmem_st_wait();
          tc_fence_before();
          l *= f;
        }
      }
      const float mxs = m_used * a.sl2;
      uint32_t pk[32];  // 64 probabilities as bf16 pairs
      const float2 sl2_2 = make_float2(a.sl2, a.sl2), nm2 = make_float2(-mxs, -mxs);
      float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};  // packed fp32 arithmetic: half the issue slots
      auto exps = [&](auto masked) {  // two straight-line variants: only the last tile of a ragged key sequence masks
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), sl2_2, nm2);
          float2 pp;  // MUFU.EX2 directly: the arguments are <= 8, no range fix-up is needed
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pp.x) : "f"(x.x));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pp.y) : "f"(x.y));
          if constexpr (decltype(masked)::value) {
            if (c >= nvalid) pp.x = 0.f;
            if (c + 1 >= nvalid) pp.y = 0.f;
          }
          ls[(c >> 1) & 1] = __fadd2_rn(ls[(c >> 1) & 1], pp);
          pk[c / 2] = pack_bf16x2(pp.x, pp.y);
        }
      };
      if (mask) exps(std::true_type{});
      else exps(std::false_type{});
      l += (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
      if (jg > 0 && !waited) mbar_wait(p_empty, (jg - 1) & 1);  // P(j-1) V(j-1) (of the previous item when j == 0) has read the tile
      // K-major SWIZZLE_128B: 16-byte chunk c16 of row r lives at chunk (c16 ^ (r & 7))
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16)
        *reinterpret_cast<uint4*>(prow + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[c16 * 4], pk[c16 * 4 + 1], pk[c16 * 4 + 2], pk[c16 * 4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- O / l -> bf16 -> global ----
    mbar_wait(o_full, it & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int q = q0 + mt * AT_BM + row;
    __nv_bfloat16* orow = a.out + (static_cast<size_t>(b) * a.Sq + q) * a.out_ld + c_head;
#pragma unroll
    for (int cb = 0; cb < AT_DH / 16; ++cb) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + t_lane + C::O_COL + mt * AT_DH + cb * 16, v);
      tmem_ld_wait();
      if (q < a.Sq) {
        uint4 o0, o1;
        o0.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
        o0.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
        o0.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
        o0.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
        o1.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
        o1.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
        o1.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
        o1.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
        reinterpret_cast<uint4*>(orow + cb * 16)[0] = o0;
        reinterpret_cast<uint4*>(orow + cb * 16)[1] = o1;
      }
    }
    // PV(next item, tile 0) overwrites O: it is issued only after this warp's P arrival for that tile, i.e. after the loads above
    tc_fence_before();
    }  // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

bool attn_tc_enabled() {  // env WD_ATTN_TC (default on)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_ATTN_TC");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
}  // namespace

// true if the launch was taken (err holds its status); false -> the caller falls back to the mma.sync kernel
bool attn_tc_try_launch(const AttnFlashArgs& a, int B, cudaStream_t s, cudaError_t* err) {
  const int C = a.heads * AT_DH;
  if (!attn_tc_enabled() || a.Skv <= 16 || a.q_ld % 8 || a.kv_ld % 8 || a.out_ld % 8 ||
      (reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
       reinterpret_cast<uintptr_t>(a.out)) % 16)
    return false;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATCfg<1>::SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(attn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATCfg<2>::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) { *err = attr_err; return true; }
  // env WD_ATTN_TC_MT2=1 selects the 256-query CTA.  Default off: measured at batch 256 (unetPhosc, eight launches per step) the
  // two-CTAs-per-SM MT = 1 build takes 0.86 ms, MT = 2 0.97 ms -- one CTA per SM cannot hide its own softmax <-> MMA ping-pong.
  static int mt2 = -1;
  if (mt2 < 0) {
    const char* e = getenv("WD_ATTN_TC_MT2");
    mt2 = e ? (atoi(e) != 0) : 0;
  }
  const bool two = mt2 && a.Sq % (2 * AT_BM) == 0;
  CUtensorMap mq, mk, mv;
  // inner extent = the C channels of this operand (the over-fetching second box of the last head is zero-filled beyond it)
  if (!tmap_encode_3d_bf16(&mq, a.q, C, a.Sq, B, a.q_ld, static_cast<uint64_t>(a.Sq) * a.q_ld, 64, AT_BM) ||
      !tmap_encode_3d_bf16(&mk, a.k, C, a.Skv, B, a.kv_ld, static_cast<uint64_t>(a.Skv) * a.kv_ld, 64, AT_BK) ||
      !tmap_encode_3d_bf16(&mv, a.v, C, a.Skv, B, a.kv_ld, static_cast<uint64_t>(a.Skv) * a.kv_ld, 64, AT_BK)) {
    *err = cudaErrorInvalidValue;
    return true;
  }
  AttnTcArgs ta{a.out, a.out_ld, a.Sq, a.Skv, a.heads, B, a.scale * 1.4426950408889634f};
  static const int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  if (two) {
    const int items = (a.Sq / (2 * AT_BM)) * a.heads * B;
    *err = launch_pdl(attn_tc_kernel<2>, dim3(items < sms ? items : sms), dim3(ATCfg<2>::THREADS), ATCfg<2>::SMEM_BYTES, s, mq, mk, mv, ta);
  } else {
    const int items = ((a.Sq + AT_BM - 1) / AT_BM) * a.heads * B;
    *err = launch_pdl(attn_tc_kernel<1>, dim3(items < 2 * sms ? items : 2 * sms), dim3(ATCfg<1>::THREADS), ATCfg<1>::SMEM_BYTES, s, mq, mk,
                      mv, ta);
  }
  return true;
}

}  // namespace wd
