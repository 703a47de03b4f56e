// CTA-pair (tcgen05 cta_group::2) implicit-GEMM kernel: the main tensor-core kernel of the hot path.
//
// Why pairs.  Measured on B200 (tools/op_bench.py, tools/micro/mma_rate.cu, profiles/): one SM ingests ~64-70 B/clk from
// L2 through TMA, while a 128x160 single-CTA tile needs 36 KB of operands per 320 MMA clocks = 115 B/clk, so that kernel
// (gemm_tc_kernel<160,...>) is load-bound at ~55 % of the tensor pipe.  Here two CTAs of a cluster (one TPC) compute a
// 256 x 320 tile together: each CTA loads its own 128 rows of A (16 KB per 64-wide K block) and HALF of the weight tile
// (2 x 80 rows, 20 KB) and the leader issues tcgen05.mma.cta_group::2 256x160x16 pairs that read both CTAs' shared memory.
// The same 36 KB per stage now feeds 640 MMA clocks (56 B/clk), below the ingest limit, and N = 320 (= model_channels)
// is covered by ONE tile, so every A element is loaded once per 3x3 tap instead of twice.
//
// Roles per CTA (320 threads): warp 0 = TMA producer (own A rows + own half of B; transaction bytes are signalled on
// the LEADER's full barrier), warp 1 = MMA issuer (leader CTA only; tcgen05.commit multicasts the "stage free" /
// "accumulator ready" arrivals to both CTAs), warps 2..9 = epilogue on the CTA's own 128 accumulator rows.
// The 128 x 320 fp32 accumulator uses 320 of the 512 TMEM columns, so it is single-buffered: the epilogue hands TMEM back
// as soon as its second (last) tcgen05.ld round has completed and finishes its arithmetic / stores under the next tile's
// MMAs.  Epilogue structure is that of gemm_tc.cu (two warps per TMEM lane quarter, dense [128][40] bf16 staging
// sub-tiles, TMA stores, TMA-prefetched residual), in two rounds of 80 columns per warp that reuse one 40 KB staging
// buffer, so that five 36 KB operand stages fit (the ring must cover ~2 us of TMA latency at 36 KB per 640 MMA clocks).
// bias / per-sample row-bias are staged once per tile in a per-warp shared-memory vector (no global loads in the loop).
#include "gemm_tc.cuh"
#include "epilogue.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace wd {

namespace {

constexpr int PAIR_BN = 320;        // tile columns (both CTAs)
constexpr int PAIR_STAGES = 5;
constexpr int PAIR_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;      // 16 KB: this CTA's 128 rows
constexpr int PAIR_BH_BYTES = 80 * GEMM_BLOCK_K * 2;               // 10 KB: 80 weight rows (half of a 160-column MMA)
constexpr int PAIR_STAGE_BYTES = PAIR_A_BYTES + 2 * PAIR_BH_BYTES;  // 36 KB
constexpr int PAIR_SUB_BYTES = GEMM_BLOCK_M * GEMM_SUB_N * 2;       // dense [128][40] bf16
constexpr int PAIR_STG_BYTES = 4 * PAIR_SUB_BYTES;                  // 40 KB: one epilogue round = 128 rows x 2 halves x 80 columns
constexpr int PAIR_VEC_BYTES = GEMM_EPI_WARPS * 160 * 4;            // per-warp bias / row-bias vector of its 160 accumulator columns
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + PAIR_STG_BYTES + PAIR_VEC_BYTES + 256;
constexpr int PAIR_TMEM_COLS = 512;
static_assert(PAIR_SMEM_BYTES <= 227 * 1024, "shared memory budget");

}  // namespace

// Tile schedule of one CTA pair.  Full rounds hand out whole 256 x 320 tiles (tile = pair + it * pairs).  A last, partial round
// (rem < pairs tiles) would leave most pairs idle for a whole tile time -- 256 tiles on 74 pairs = 3.46 waves, 13.5 % of the launch --
// so when 2 * rem <= pairs each tail tile is cut into its two 160-column halves and handed to two neighbouring pairs: the same A rows
// (shared through L2), half of the weights, half of the MMAs; load-bound at ~0.6 of a whole tile's time, no cross-CTA exchange.
// nmask: bit h set = this pair computes columns [160 h, 160 h + 160) of the tile.
struct PairSched {
  int P, R, rem;
  bool split;
};
WD_DEVINL bool pair_sched(const PairSched& s, int pair, int it, int* tile, int* nmask) {
  if (it < s.R) {
    *tile = pair + it * s.P;
    *nmask = 3;
    return true;
  }
  if (it > s.R) return false;
  if (s.split) {
    if (pair >= 2 * s.rem) return false;
    *tile = s.R * s.P + (pair >> 1);
    *nmask = 1 << (pair & 1);
    return true;
  }
  if (pair >= s.rem) return false;
  *tile = s.R * s.P + pair;
  *nmask = 3;
  return true;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapRes,
                 const GemmArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment (no static shared memory in this kernel)
  uint8_t* stg = smem + PAIR_STAGES * PAIR_STAGE_BYTES;  // [2 halves][2 sub-tiles][128][40] bf16: one round of the tile
  float* vecs = reinterpret_cast<float*>(stg + PAIR_STG_BYTES);  // [8 epilogue warps][160]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + PAIR_STG_BYTES + PAIR_VEC_BYTES);  // leader's are used by both CTAs
  uint64_t* empty_bar = full_bar + PAIR_STAGES;                            // local (multicast commit)
  uint64_t* tmem_full_bar = empty_bar + PAIR_STAGES;                       // local (multicast commit)
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;                            // leader's: 16 epilogue-warp arrivals
  uint64_t* res_full_bar = tmem_empty_bar + 1;                             // [2 halves], local
  uint64_t* gn_bar = res_full_bar + 2;                                     // [2 rounds], local: 16 epilogue-warp arrivals (GemmArgs::gn_apply)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gn_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int n_tiles = args.N / PAIR_BN;
  const int m_tiles = (args.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
  const int total_tiles = n_tiles * m_tiles;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  PairSched sched;
  sched.P = npairs;
  sched.R = total_tiles / npairs;
  sched.rem = total_tiles % npairs;
  sched.split = args.tail_split && n_tiles == 1 && sched.rem > 0 && 2 * sched.rem <= npairs;

  int total_k = 0;
#pragma unroll
  for (int s = 0; s < GEMM_MAX_SRC; ++s)
    if (s < args.num_src) total_k += args.taps[s] * args.chunks[s];

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA0);
    if (args.num_src > 1) tma_prefetch_desc(&mapA1);
    if (args.num_src > 2) tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapB);
    if (!args.out_f32) tma_prefetch_desc(&mapOut);
    if (args.residual) tma_prefetch_desc(&mapRes);
    for (int i = 0; i < PAIR_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 2 * GEMM_EPI_WARPS);  // one arrival per epilogue warp of BOTH CTAs
    mbar_init(&res_full_bar[0], 1);
    mbar_init(&res_full_bar[1], 1);
    mbar_init(&gn_bar[0], 2 * GEMM_EPI_WARPS);
    mbar_init(&gn_bar[1], 2 * GEMM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<PAIR_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // PDL (common.cuh): the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    // (elect.sync, not `lane == 0`: ptxas then knows exactly one lane is active and issues UTMALDG / UTCHMMA / UTCBAR
    //  straight from uniform registers instead of wrapping each one in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int tile, nmask;
      for (int pit = 0; pair_sched(sched, pair, pit, &tile, &nmask); ++pit) {
        const int m0 = (tile / n_tiles) * (2 * GEMM_BLOCK_M) + static_cast<int>(rank) * GEMM_BLOCK_M;  // this CTA's rows
        const int n0 = (tile % n_tiles) * PAIR_BN;
        int img = 0, oh0 = 0;
        if (args.conv) {
          img = m0 / args.HWout;
          oh0 = (m0 % args.HWout) / args.Wout;
        }
        int kb = 0;
        for (int s = 0; s < args.num_src; ++s) {
          const CUtensorMap* mapA = (s == 0) ? &mapA0 : (s == 1 ? &mapA1 : &mapA2);
          const int taps = args.taps[s];
          const int chunks = args.chunks[s];
          const int st = args.stride[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int up_a = (args.up_phase - 1) >> 1, up_b = (args.up_phase - 1) & 1;
            const int dy = (taps == 9) ? tap / 3 - 1 : (taps == 4 ? (tap >> 1) - 1 + up_a : 0);
            const int dx = (taps == 9) ? tap % 3 - 1 : (taps == 4 ? (tap & 1) - 1 + up_b : 0);
            for (int ch = 0; ch < chunks; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);  // the leader's full barrier
              const int nb_boxes = nmask == 3 ? 2 : 1;  // weight boxes per CTA: one per 160-column half this pair computes
              if (is_leader)
                mbar_arrive_expect_tx(&full_bar[stage], 2 * (((args.dbg & 8) ? 0 : PAIR_A_BYTES) + ((args.dbg & 4) ? 0 : nb_boxes * PAIR_BH_BYTES)));
              uint8_t* sA = smem + stage * PAIR_STAGE_BYTES;
              uint8_t* sB = sA + PAIR_A_BYTES;
              if (!(args.dbg & 8)) {
                if (args.conv)
                  tma_load_4d_pair(sA, mapA, fb, ch * GEMM_BLOCK_K, dx, oh0 * st + dy, img);
                else
                  tma_load_2d_pair(sA, mapA, fb, ch * GEMM_BLOCK_K, m0);
              }
              // weight rows of MMA j (columns n0 + 160 j ..): this CTA supplies rows [80 rank, +80) of them
              if (!(args.dbg & 4)) {
                if (nmask & 1) tma_load_2d_pair(sB, &mapB, fb, kb * GEMM_BLOCK_K, n0 + static_cast<int>(rank) * 80);
                if (nmask & 2) tma_load_2d_pair(sB + PAIR_BH_BYTES, &mapB, fb, kb * GEMM_BLOCK_K, n0 + 160 + static_cast<int>(rank) * 80);
              }
              ++kb;
              if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA, single thread) ===========================
    if (is_leader && elect_one()) {
      constexpr uint32_t idesc_bf16 = make_idesc_bf16_f32(2 * GEMM_BLOCK_M, 160);
      constexpr uint32_t idesc_f16 = make_idesc_f16_f32(2 * GEMM_BLOCK_M, 160);
      int kend[GEMM_MAX_SRC];  // K-block index at which each source ends
      {
        int acc_k = 0;
#pragma unroll
        for (int s = 0; s < GEMM_MAX_SRC; ++s) {
          if (s < args.num_src) acc_k += args.taps[s] * args.chunks[s];
          kend[s] = acc_k;
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int tile, nmask;
      for (; pair_sched(sched, pair, it, &tile, &nmask); ++it) {
        // both CTAs' epilogues have drained the accumulator.  CTA-scope acquire / release on both sides (mbar_arrive_remote): the
        // cluster-scope forms put a MEMBAR.ALL.GPU in front of every epilogue warp's arrival (-2 % on the GEMM class, R2f)
        mbar_wait(tmem_empty_bar, (it & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < total_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);  // CTA-scope acquire: a cluster-scope one costs a CCTL.IVALL (L1 flush) per K block
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * PAIR_STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc0 = make_smem_desc_sw128(a_addr + PAIR_A_BYTES);
          const uint64_t b_desc1 = make_smem_desc_sw128(a_addr + PAIR_A_BYTES + PAIR_BH_BYTES);
          const int src = kb < kend[0] ? 0 : (kb < kend[1] ? 1 : 2);
          const uint32_t idesc = args.a_f16[src] ? idesc_f16 : idesc_bf16;
          if (!(args.dbg & 16)) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              if (nmask & 1) umma_f16_ss_pair(tmem_base, a_desc + 2 * k, b_desc0 + 2 * k, idesc, (kb | k) != 0);
              if (nmask & 2) umma_f16_ss_pair(tmem_base + 160, a_desc + 2 * k, b_desc1 + 2 * k, idesc, (kb | k) != 0);
            }
          }
          umma_commit_pair(&empty_bar[stage]);  // frees the stage in both CTAs when these MMAs retire
          if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tmem_full_bar);  // accumulator complete (both CTAs)
      }
    }
  } else {
    // =========================== epilogue (both CTAs, own 128 rows) ===========================
    const int q = warp & 3;            // TMEM lane quarter
    const int half = (warp - 2) >> 2;  // column half: tile columns [160 half, +160)
    const int row = q * 32 + lane;
    const bool leader_warp = (q == 0);  // its elected lane issues this half's TMA stores / residual loads (elect.sync is
                                        // deterministic, so the bulk async-groups always belong to the same thread)
    const int bar_id = 1 + half;
    const bool use_stg = !args.out_f32;
    // rare flavours (time-embedding GEMMs, operator tests): SiLU, fp32 output, per-thread row-bias rows, or a residual stored
    // in another 16-bit format than the output -> generic run-time-flag epilogue, residual read from global memory
    const bool slow_path = args.act != ACT_NONE || args.out_f32 || (args.rowbias && args.rows_per_sample % 32 != 0) ||
                           (args.residual && (args.res_f16 != 0) != (args.out_f16 != 0));
    const bool has_res = use_stg && args.residual != nullptr && !slow_path;  // residual TMA-prefetched into the staging tile
    uint8_t* const stg_half = stg + half * 2 * PAIR_SUB_BYTES;  // 2 sub-tiles = the 80 columns of one round
    float* const wv = vecs + (warp - 2) * 160;
    const uint32_t te_addr = mapa_shared(smem_u32(tmem_empty_bar), 0);
    const bool rb_warp_uniform = args.rowbias && (args.rows_per_sample % 32 == 0);
    const bool out_f16 = args.out_f16 != 0, res_f16 = args.res_f16 != 0;

    // residual sub-tiles of (tile_, round rnd_) of this half -> staging
    auto issue_res_load = [&](int tile_, int rnd_) {
      const int m0_ = (tile_ / n_tiles) * (2 * GEMM_BLOCK_M) + static_cast<int>(rank) * GEMM_BLOCK_M;
      const int c0_ = (tile_ % n_tiles) * PAIR_BN + half * 160 + rnd_ * 80;
      mbar_arrive_expect_tx(&res_full_bar[half], 2 * PAIR_SUB_BYTES);
      tma_load_2d(stg_half, &mapRes, &res_full_bar[half], c0_, m0_);
      tma_load_2d(stg_half + PAIR_SUB_BYTES, &mapRes, &res_full_bar[half], c0_ + GEMM_SUB_N, m0_);
    };
    {
      int t0_, nm0_;
      if (has_res && leader_warp && pair_sched(sched, pair, 0, &t0_, &nm0_) && ((nm0_ >> half) & 1)) {
        if (elect_one()) issue_res_load(t0_, 0);
      }
    }

    int it = 0;
    int tile, nmask;
    for (; pair_sched(sched, pair, it, &tile, &nmask); ++it) {
      const int n_tile = tile % n_tiles;
      const int m0 = (tile / n_tiles) * (2 * GEMM_BLOCK_M) + static_cast<int>(rank) * GEMM_BLOCK_M;
      const int n0 = n_tile * PAIR_BN;
      const int m = m0 + row;
      const bool valid = m < args.M;

      // ---- per-warp vector of the additive per-column terms (bias + the warp's sample row of the row-bias) ----
      const float* rb = nullptr;  // per-thread row-bias only when the rows of a warp can belong to different samples
      {
        const int mw = min(m0 + q * 32, args.M - 1);
        const float* rbw = nullptr;
        if (rb_warp_uniform) {
          const int sw = mw / args.rows_per_sample;
          rbw = args.rowbias + (args.rowbias_idx ? args.rowbias_idx[sw] : static_cast<long long>(sw)) * args.rb_ld;
        } else if (args.rowbias) {
          const int sample = valid ? (m / args.rows_per_sample) : 0;
          rb = args.rowbias + (args.rowbias_idx ? args.rowbias_idx[sample] : static_cast<long long>(sample)) * args.rb_ld;
        }
        __syncwarp();  // all lanes are done reading the previous tile's vector
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int c = lane + 32 * i;
          // STD: accumulator column n0 + 160 half + c.  GEGLU: c < 80 -> value column 80 half + c, else its gate (+160)
          const int col = args.geglu ? (n0 + half * 80 + (c < 80 ? c : c - 80 + 160)) : (n0 + half * 160 + c);
          float x = args.bias ? __ldg(args.bias + col) : 0.f;
          if (rbw) x += __ldg(rbw + col);
          wv[c] = x;
        }
        __syncwarp();
      }

      mbar_wait(tmem_full_bar, it & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      if (!((nmask >> half) & 1)) {
        // tail half-tile owned by the other column half: this warp only keeps the pair's barrier protocols whole
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_remote(te_addr);
          if (args.gn_apply) {
#pragma unroll
            for (int r_ = 0; r_ < 2; ++r_) {
              mbar_arrive_cluster(mapa_shared(smem_u32(&gn_bar[r_]), 0));
              mbar_arrive_cluster(mapa_shared(smem_u32(&gn_bar[r_]), 1));
            }
          }
        }
        continue;
      }
      if (args.dbg & 2) {  // experiment: no epilogue at all
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(te_addr);
        continue;
      }
      uint8_t* const srow = stg_half + row * (GEMM_SUB_N * 2);

      if (args.gn_apply) {
        // ===== GroupNorm + SiLU of the tile applied here (GemmArgs::gn_apply): out = silu(GN(acc + bias + row-bias)), bf16 =====
        // Each round's partial sums are published as soon as the round is drained (global memory, release / acquire at cluster
        // scope on gn_bar[rnd]); the exchange of round 0 travels under the TMEM drain of round 1 and that of round 1 under the
        // normalisation / store of round 0.
        const int rps = args.rows_per_sample;  // 256 % rps == 0, rps % 32 == 0: the pair tile holds whole samples
        const int mw = m0 + q * 32;            // the 32 rows of a warp lie in one sample
        const bool wvalid = mw < args.M;
        const int smp = wvalid ? mw / rps : 0;
        const int slot = (mw % rps) >> 5, nslot = rps >> 5;
        const int G = args.N / 10;
        float2* const part = reinterpret_cast<float2*>(args.gn_partial) + (static_cast<size_t>(smp) * G + half * 16) * nslot;
        float mean = 0.f, rstd = 0.f;  // lane L < 8: group (16 half + 8 rnd + L) of the warp's sample, re-formed per round
        auto publish = [&](int rnd_, float tot) {
          if (wvalid && lane < 16) reinterpret_cast<float*>(part + static_cast<size_t>(8 * rnd_ + (lane >> 1)) * nslot + slot)[lane & 1] = tot;
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(mapa_shared(smem_u32(&gn_bar[rnd_]), 0));
            mbar_arrive_cluster(mapa_shared(smem_u32(&gn_bar[rnd_]), 1));
          }
        };
        // every epilogue warp of the pair has published its partials of round rnd_: statistics + the round's 80 scale factors
        auto form_stats = [&](int rnd_) {
          mbar_wait_cluster(&gn_bar[rnd_], it & 1);
          if (wvalid && lane < 8) {
            const float2* pg = part + static_cast<size_t>(8 * rnd_ + lane) * nslot;
            float S = 0.f, Q = 0.f;
            for (int i = 0; i < nslot; ++i) {  // fixed order: bit-reproducible, the order of groupnorm_apply_bulk_kernel
              const float2 t = __ldcg(pg + i);
              S += t.x;
              Q += t.y;
            }
            const float inv_n = 1.0f / static_cast<float>(10 * rps);
            mean = S * inv_n;
            rstd = rsqrtf(fmaxf(fmaf(-mean, mean, Q * inv_n), 0.f) + args.gn_eps);
          }
        };
        float gam[3], bet[3];  // gamma / beta of columns lane, lane + 32, lane + 64 of the current round (prefetched)
        auto load_gb = [&](int rnd_) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c = lane + 32 * i;
            gam[i] = c < 80 ? __ldg(args.gn_gamma + n0 + half * 160 + rnd_ * 80 + c) : 0.f;
            bet[i] = c < 80 ? __ldg(args.gn_beta + n0 + half * 160 + rnd_ * 80 + c) : 0.f;
          }
        };
        // scale / shift of the round's 80 columns -> wv[0..79] = sc / 2, wv[80..159] = sh / 2: the warp's 160-float vector is free
        // once round 1's additive terms have been consumed.  sc = rstd gamma, sh = fma(-mean, sc, beta): the operations of
        // groupnorm_apply_bulk_kernel, so both forms of the ResBlock give the same bits
        auto build_table = [&]() {
          __syncwarp();  // every lane is done with the previous contents
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c = lane + 32 * i;
            const float r = __shfl_sync(0xffffffffu, rstd, (c < 80 ? c : 0) / 10);
            const float mu = __shfl_sync(0xffffffffu, mean, (c < 80 ? c : 0) / 10);
            if (c < 80) {
              const float sc = r * gam[i];
              const float sh = fmaf(-mu, sc, bet[i]);
              wv[c] = 0.5f * sc;
              wv[80 + c] = 0.5f * sh;
            }
          }
          __syncwarp();
        };
        uint32_t hreg[40];  // round 1 waits here (fp16 pairs); round 0 waits in the staging tile
        {
          uint32_t v[80];
#pragma unroll
          for (int c = 0; c < 5; ++c) tmem_ld_32x32b_x16p(t_row + half * 160 + c * 16, v + c * 16);
          tmem_ld_wait();
          if (leader_warp) {
            if (elect_one()) bulk_wait_group_read<0>();  // the previous tile's last store has read the staging tile
          }
          named_barrier_sync(bar_id, 128);
          float gs[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) gs[i] = 0.f;
          epi_round80<false, true, true>(v, wv, srow, PAIR_SUB_BYTES, valid, gs);
          publish(0, warp_transpose_reduce16(gs, lane));
        }
        {
          uint32_t v[80];
#pragma unroll
          for (int c = 0; c < 5; ++c) tmem_ld_32x32b_x16p(t_row + half * 160 + 80 + c * 16, v + c * 16);
          load_gb(0);
          tmem_ld_wait();
          tc_fence_before();  // accumulator fully read by this warp: hand TMEM back to the leader's MMA warp
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(te_addr);
          float gs[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) gs[i] = 0.f;
          epi_round80_hold(v, wv + 80, hreg, valid, gs);
          publish(1, warp_transpose_reduce16(gs, lane));
        }
        // ---- round 0: normalise the staged fp16 rows in place, store ----
        form_stats(0);
        build_table();
        load_gb(1);
#pragma unroll
        for (int c = 0; c < 10; ++c) {
          uint4* sp = reinterpret_cast<uint4*>(srow + (c / 5) * PAIR_SUB_BYTES + (c % 5) * 16);
          *sp = gn_apply_vec8(*sp, wv + c * 8, wv + 80 + c * 8);
        }
        fence_proxy_async();
        named_barrier_sync(bar_id, 128);
        if (leader_warp && elect_one()) {
          const int oc0 = n0 + half * 160;
          tma_store_2d(&mapOut, stg_half, oc0, m0);
          tma_store_2d(&mapOut, stg_half + PAIR_SUB_BYTES, oc0 + GEMM_SUB_N, m0);
          bulk_commit_group();
        }
        // ---- round 1: from the registers ----
        form_stats(1);
        build_table();
        if (leader_warp) {
          if (elect_one()) bulk_wait_group_read<0>();  // round 0's store has read the staging tile
        }
        named_barrier_sync(bar_id, 128);
#pragma unroll
        for (int c = 0; c < 10; ++c) {
          uint4* sp = reinterpret_cast<uint4*>(srow + (c / 5) * PAIR_SUB_BYTES + (c % 5) * 16);
          *sp = gn_apply_vec8(make_uint4(hreg[c * 4], hreg[c * 4 + 1], hreg[c * 4 + 2], hreg[c * 4 + 3]), wv + c * 8, wv + 80 + c * 8);
        }
        fence_proxy_async();
        named_barrier_sync(bar_id, 128);
        if (leader_warp && elect_one()) {
          const int oc0 = n0 + half * 160 + 80;
          tma_store_2d(&mapOut, stg_half, oc0, m0);
          tma_store_2d(&mapOut, stg_half + PAIR_SUB_BYTES, oc0 + GEMM_SUB_N, m0);
          bulk_commit_group();
        }
        continue;
      }

#pragma unroll 1
      for (int rnd = 0; rnd < 2; ++rnd) {
        // ---- drain 32 rows x 80 accumulator columns (GEGLU: 40 value + 40 gate columns) ----
        uint32_t v[80];
        if (!args.geglu) {
#pragma unroll
          for (int c = 0; c < 5; ++c) tmem_ld_32x32b_x16p(t_row + half * 160 + rnd * 80 + c * 16, v + c * 16);
        } else {
          const int vc = half * 80 + rnd * 40;  // value columns [vc, +40), gates at 160 + vc
          tmem_ld_32x32b_x16p(t_row + vc, v);
          tmem_ld_32x32b_x16p(t_row + vc + 16, v + 16);
          tmem_ld_32x32b_x8(t_row + vc + 32, v + 32);
          tmem_ld_32x32b_x16p(t_row + 160 + vc, v + 40);
          tmem_ld_32x32b_x16p(t_row + 160 + vc + 16, v + 56);
          tmem_ld_32x32b_x8(t_row + 160 + vc + 32, v + 72);
        }
        tmem_ld_wait();
        if (rnd == 1) {  // accumulator fully read by this warp: hand TMEM back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(te_addr);
        }

        // ---- the staging buffer of this half must be free: its previous TMA store has read it / the residual has landed ----
        if (use_stg) {
          if (has_res) {
            mbar_wait(&res_full_bar[half], (2 * it + rnd) & 1);
          } else {
            if (leader_warp) {
              if (elect_one()) bulk_wait_group_read<0>();
            }
            named_barrier_sync(bar_id, 128);
          }
        }

        if (args.geglu) {
          epi_geglu40<false>(v, wv + rnd * 40, wv + 80 + rnd * 40, srow);
        } else {
          const int nb = n0 + half * 160 + rnd * 80;
          const float* wvr = wv + rnd * 80;
          if (slow_path) {
            epi_round80_generic(v, wvr, rb ? rb + nb : nullptr, args.act == ACT_SILU,
                                (args.residual && !has_res) ? args.residual + static_cast<size_t>(m) * args.res_ld + nb : nullptr,
                                res_f16, use_stg, srow, PAIR_SUB_BYTES, args.out_f32 != 0, out_f16,
                                args.out_f32 ? static_cast<void*>(static_cast<float*>(args.out) + static_cast<size_t>(m) * args.out_ld + nb)
                                             : static_cast<void*>(static_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(m) * args.out_ld + nb),
                                valid);
          } else {
            float gs[16];  // GroupNorm partials: [2g] = sum, [2g+1] = sum of squares of group g (10 columns) of this row
#pragma unroll
            for (int i = 0; i < 16; ++i) gs[i] = 0.f;
            epi_round80_dispatch(has_res, args.gn_partial != nullptr, out_f16, false, v, wvr, srow, PAIR_SUB_BYTES, valid, gs);
            if (args.gn_partial) {
              // rows of a warp belong to one sample (rows_per_sample % 32 == 0): reduce over the 32 rows, lane L < 16 keeps entry L
              const float tot = warp_transpose_reduce16(gs, lane);
              const int mw = m0 + q * 32;
              if (mw < args.M && lane < 16) {
                const int smp = mw / args.rows_per_sample;
                const int slot = args.gn_slot_base + ((mw % args.rows_per_sample) >> 5);
                const int nslot = args.gn_nslot ? args.gn_nslot : args.rows_per_sample >> 5;
                const int G = args.N / 10;
                const int g = (nb / 10) + (lane >> 1);
                args.gn_partial[((static_cast<size_t>(smp) * G + g) * nslot + slot) * 2 + (lane & 1)] = tot;
              }
            }
          }
        }

        // ---- publish the staged round with TMA; prefetch the residual of the next round ----
        if (use_stg) {
          fence_proxy_async();  // generic-proxy smem writes -> visible to the async proxy (TMA)
          named_barrier_sync(bar_id, 128);
          if (leader_warp && elect_one()) {
            if (!args.geglu) {
              const int oc0 = n0 + half * 160 + rnd * 80;
              if (args.up_phase) {  // sub-pixel phase: (c, x, y, n) of the phase grid; this CTA's 128 rows hold whole images
                tma_store_4d(&mapOut, stg_half, oc0, 0, 0, m0 / args.HWout);
                tma_store_4d(&mapOut, stg_half + PAIR_SUB_BYTES, oc0 + GEMM_SUB_N, 0, 0, m0 / args.HWout);
              } else {
                tma_store_2d_keep(&mapOut, stg_half, oc0, m0, (args.dbg & 64) != 0);
                tma_store_2d_keep(&mapOut, stg_half + PAIR_SUB_BYTES, oc0 + GEMM_SUB_N, m0, (args.dbg & 64) != 0);
              }
            } else {
              tma_store_2d_keep(&mapOut, stg_half, n_tile * 160 + half * 80 + rnd * 40, m0, (args.dbg & 64) != 0);
            }
            bulk_commit_group();
            if (has_res) {
              int nt = tile, nm_ = nmask;
              const bool more = rnd == 0 ? true : (pair_sched(sched, pair, it + 1, &nt, &nm_) && ((nm_ >> half) & 1));
              if (more) {
                bulk_wait_group_read<0>();  // the store above has read the staging buffer
                issue_res_load(nt, rnd ^ 1);
              }
            }
          }
        }
      }
    }
    if (use_stg && leader_warp) {
      if (elect_one()) bulk_wait_group_read<0>();  // smem must outlive the last TMA store's read
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while its peer may still read its smem / signal its barriers
  if (warp == 1) tmem_dealloc_pair<PAIR_TMEM_COLS>(tmem_base);
}

bool gemm_pair_supported(const GemmArgs& a) {
  if (a.epi != EPI_STD) return false;
  if (a.N % PAIR_BN) return false;
  if (a.ln_out || a.ln_stats) return false;  // LayerNorm folding lives in the single-CTA kernel (K = 320 GEMMs)
  if (a.geglu) return false;  // the kernel implements it (value | gate per 320-column tile) but the weights are packed for 160-column tiles
  if (a.out_f32 && a.residual) return false;
  if (a.up_phase == 5) return false;  // all phases in one launch: single-CTA kernel only
  if (a.up_phase && (a.gn_apply || a.residual || a.out_f32 || a.geglu || a.num_src != 1 || a.taps[0] != 4 || !a.conv || a.HWout <= 0 ||
                     GEMM_BLOCK_M % a.HWout))
    return false;
  if (a.gn_apply && (a.N != PAIR_BN || !a.gn_partial || a.gn_cpg != 10 || !a.gn_gamma || !a.gn_beta || a.residual || a.out_f32 ||
                     a.act != ACT_NONE || a.rows_per_sample % 32 || 256 % a.rows_per_sample))
    return false;
  return true;
}

cudaError_t gemm_pair_launch(const GemmLaunch& L, int num_sms, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const GemmArgs& a = L.args;
  if (!gemm_pair_supported(a) || a.M <= 0) return cudaErrorInvalidValue;
  const int tiles = (a.N / PAIR_BN) * ((a.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M));
  const int max_pairs = num_sms / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  static const int tail_split = [] {  // env WD_PAIR_TAIL_SPLIT (default on)
    const char* e = getenv("WD_PAIR_TAIL_SPLIT");
    return e ? (atoi(e) != 0) : 1;
  }();
  GemmArgs a2 = a;
  a2.tail_split = tail_split && !a.geglu && !a.out_f32 && a.act == ACT_NONE;
  return launch_pdl(gemm_pair_kernel, dim3(2 * pairs), dim3(GEMM_THREADS), PAIR_SMEM_BYTES, stream, L.mapA[0], L.mapA[1], L.mapA[2],
                    L.mapB, L.mapOut, L.mapRes, a2);
}

}  // namespace wd
