// Persistent warp-specialised tcgen05 implicit-GEMM kernel, see gemm_tc.cuh for the contract.
#include "gemm_tc.cuh"
#include "epilogue.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace wd {

// dbg & 1024: clock64 stamps of the epilogue phases of (CTA 0, leader warp of column half 0), 12 per tile (tools/gemm_trace.py)
__device__ unsigned long long g_gemm_trace[12 * 64];
__device__ unsigned long long g_gemm_cta_times[4 * 160];  // per CTA: globaltimer at entry, after the prologue, after pdl_wait, at exit
WD_DEVINL unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define WD_TRACE(slot)                                                                                  \
  do {                                                                                                  \
    if ((args.dbg & 1024) && blockIdx.x == 0 && warp == 4 && lane == 0 && it < 64) g_gemm_trace[it * 12 + (slot)] = clock64(); \
  } while (0)

// WSK > 0: "weight-stationary" mode for short K (K <= 64 WSK): the CTA keeps its whole [BN x K] weight tile resident in
// shared memory (loaded once), works on ONE n-tile for all its m-tiles, and the ring only streams A (16 KB per K block).
// For K = 320 this halves the L2->SM operand traffic (100 KB of weights were re-streamed for every 80 KB of activations),
// which is what bounds the 1x1 / Linear GEMMs of the transformer blocks (the per-SM TMA ingest rate, see gemm_pair.cu).
template <int BN, int STAGES_, int NSTG_, int WSK_ = 0, int ATT_ = 0, int RESK_ = 0, int GRP_ = 0>
struct Cfg {
  static constexpr int STAGES = STAGES_;
  static constexpr int NSTG = NSTG_;  // staging buffers per column half (0: the epilogue writes global memory directly)
  static constexpr int WSK = WSK_;
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BN * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = WSK ? A_BYTES : A_BYTES + B_BYTES;
  // 64 x 64 fp16 identity tile (GemmArgs::res_k): resident beside the weights in the weight-stationary build, streamed into the
  // stage's B slot with each residual block otherwise
  static constexpr int IDENT_TILE = 64 * GEMM_BLOCK_K * 2;
  static constexpr int IDENT_BYTES = (RESK_ && WSK_) ? IDENT_TILE : 0;
  static constexpr int BRES_BYTES = WSK * B_BYTES + IDENT_BYTES;  // resident weight tile (+ identity)
  static constexpr int ACC_STRIDE = (BN == GEMM_BLOCK_N) ? 256 : 32;  // TMEM column offset of accumulator buffer 1
  static constexpr int TMEM_COLS = (BN == GEMM_BLOCK_N) ? 512 : 64;
  static constexpr int SUB_BYTES = GEMM_BLOCK_M * GEMM_SUB_N * 2;  // one dense [128][40] bf16 sub-tile
  static constexpr int HALF_STG_BYTES = 2 * SUB_BYTES;             // 80 columns of one column half
  static constexpr int STG_BYTES = 2 * NSTG * HALF_STG_BYTES;
  // per-warp bias / row-bias vector of its 80 accumulator columns (+ a second one, the LayerNorm column sums, in WS mode)
  // (GRP: a warp handles both column halves of its tiles -> 160 entries per vector)
  static constexpr int VEC_BYTES = (WSK ? 2 : 1) * GEMM_EPI_WARPS * 80 * 4 * (GRP_ ? 2 : 1);
  // fused context attention: bf16 K and V rows [16][ATT_KP] of the tile's sample, one (K, V) pair per column half (= head)
  static constexpr int KV_BYTES = ATT_ ? 2 * 2 * 16 * ATT_KP * 2 : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BRES_BYTES + STG_BYTES + VEC_BYTES + KV_BYTES + 256 /*barriers*/;
};

template <int BN, int EPI, int STAGES, int NSTG, int WSK, int ATT, int RESK, int GRP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapRes,
               const GemmArgs args) {
  using C = Cfg<BN, STAGES, NSTG, WSK, ATT, RESK, GRP>;
  constexpr int NB = NSTG > 0 ? NSTG : 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((args.dbg & 1024) && threadIdx.x == 64 && blockIdx.x < 160) g_gemm_cta_times[blockIdx.x * 4 + 0] = gtimer();
  if (smem_u32(smem) & 1023) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment (no static shared memory in this kernel)
  uint8_t* bres = smem + C::STAGES * C::STAGE_BYTES;  // [WSK][BN x 64] resident weight tile (weight-stationary mode)
  uint8_t* stg = bres + C::BRES_BYTES;                // [half][NSTG][2 sub-tiles][128][40] bf16
  float* vecs = reinterpret_cast<float*>(stg + C::STG_BYTES);  // [8 epilogue warps][80]
  __nv_bfloat16* kvs = reinterpret_cast<__nv_bfloat16*>(stg + C::STG_BYTES + C::VEC_BYTES);  // [half][K|V][16][ATT_KP] (ATT builds)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + C::STG_BYTES + C::VEC_BYTES + C::KV_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint64_t* res_full_bar = tmem_empty_bar + 2;      // [2 halves][2]
  uint64_t* bres_bar = res_full_bar + 4;            // resident weight tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = args.N / BN;
  const int m_tiles = (args.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int total_tiles = n_tiles * m_tiles;
  constexpr int EPI_ACTIVE_WARPS = (EPI == EPI_SAMPLER) ? 4 : GEMM_EPI_WARPS;

  int total_k = 0;
#pragma unroll
  for (int s = 0; s < GEMM_MAX_SRC; ++s)
    if (s < args.num_src) total_k += args.taps[s] * args.chunks[s];

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA0);
    if (args.num_src > 1) tma_prefetch_desc(&mapA1);
    if (args.num_src > 2) tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapB);
    if (NSTG > 0 && !args.out_f32) tma_prefetch_desc(&mapOut);
    if (NSTG > 0 && args.residual && !(RESK && args.res_k)) tma_prefetch_desc(&mapRes);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], GRP ? 4 : EPI_ACTIVE_WARPS);  // one arrival per epilogue warp that drains the buffer
    }
    for (int i = 0; i < 4; ++i) mbar_init(&res_full_bar[i], 1);
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if ((args.dbg & 1024) && threadIdx.x == 64 && blockIdx.x < 160) g_gemm_cta_times[blockIdx.x * 4 + 1] = gtimer();
  pdl_trigger();
  // PDL: everything above (and the resident weight tile below) touches only static data; every other global access of this
  // kernel -- operand loads, statistics / residual / row-bias reads, all output writes -- comes after pdl_wait()
  if (warp != 0) pdl_wait();
  if ((args.dbg & 1024) && threadIdx.x == 64 && blockIdx.x < 160) g_gemm_cta_times[blockIdx.x * 4 + 2] = gtimer();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // (elect.sync, not `lane == 0`: ptxas then knows exactly one lane is active and issues UTMALDG / UTCHMMA / UTCBAR
    //  straight from uniform registers instead of wrapping each one in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (WSK > 0 && static_cast<int>(blockIdx.x) < total_tiles) {
        // weight-stationary: gridDim.x is a multiple of n_tiles, so this CTA's n-tile never changes
        const int n0 = (blockIdx.x % n_tiles) * BN;
        mbar_arrive_expect_tx(bres_bar, total_k * C::B_BYTES + C::IDENT_BYTES);
        for (int kb = 0; kb < total_k; ++kb) tma_load_2d(bres + kb * C::B_BYTES, &mapB, bres_bar, kb * GEMM_BLOCK_K, n0);
        if constexpr (RESK && WSK > 0) tma_load_2d(bres + WSK * C::B_BYTES, &mapA2, bres_bar, 0, 0);
      }
      pdl_wait();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * GEMM_BLOCK_M;
        const int n0 = (tile % n_tiles) * BN;
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
            const int up_ph = args.up_phase == 5 ? n0 / 320 : args.up_phase - 1;
            const int up_a = up_ph >> 1, up_b = up_ph & 1;
            const int dy = (taps == 9) ? tap / 3 - 1 : (taps == 4 ? (tap >> 1) - 1 + up_a : 0);
            const int dx = (taps == 9) ? tap % 3 - 1 : (taps == 4 ? (tap & 1) - 1 + up_b : 0);
            for (int ch = 0; ch < chunks; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage],
                                    ((args.dbg & 8) ? 0 : C::A_BYTES) + ((WSK > 0 || (args.dbg & 4)) ? 0 : C::B_BYTES));
              uint8_t* sA = smem + stage * C::STAGE_BYTES;
              uint8_t* sB = sA + C::A_BYTES;
              if (!(args.dbg & 8)) {
                if (args.conv)
                  tma_load_4d(sA, mapA, &full_bar[stage], ch * GEMM_BLOCK_K, dx, oh0 * st + dy, img);
                else
                  tma_load_2d(sA, mapA, &full_bar[stage], ch * GEMM_BLOCK_K, m0);
              }
              if (WSK == 0 && !(args.dbg & 4)) tma_load_2d(sB, &mapB, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
              ++kb;
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
        if constexpr (RESK) {
          // residual K blocks: only the 64-channel blocks that overlap this tile's BN output columns
          for (int rb = n0 / GEMM_BLOCK_K; rb * GEMM_BLOCK_K < n0 + BN; ++rb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], C::A_BYTES + (WSK > 0 ? 0 : C::IDENT_TILE));
            tma_load_2d(smem + stage * C::STAGE_BYTES, &mapA1, &full_bar[stage], rb * GEMM_BLOCK_K, m0);
            if constexpr (WSK == 0) tma_load_2d(smem + stage * C::STAGE_BYTES + C::A_BYTES, &mapA2, &full_bar[stage], 0, 0);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (single elected thread) ===========================
    if (elect_one()) {
      constexpr uint32_t idesc_bf16 = make_idesc_bf16_f32(GEMM_BLOCK_M, BN);
      constexpr uint32_t idesc_f16 = make_idesc_f16_f32(GEMM_BLOCK_M, BN);
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
      if (WSK > 0 && static_cast<int>(blockIdx.x) < total_tiles) mbar_wait(bres_bar, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
        for (int kb = 0; kb < total_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (!(args.dbg & 512)) tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(WSK > 0 ? smem_u32(bres + kb * C::B_BYTES) : a_addr + C::A_BYTES);
          const int src = kb < kend[0] ? 0 : (kb < kend[1] ? 1 : 2);
          const uint32_t idesc = args.a_f16[src] ? idesc_f16 : idesc_bf16;
          if (!(args.dbg & 16)) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              // advance 16 bf16 = 32 B along K inside the swizzled row: +2 in the (addr >> 4) field
              umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (RESK) {
          // acc[:, c] += residual[:, n0 + c]: residual block rb (channels 64 rb ..) times the rows of the identity tile that
          // select the overlap with [n0, n0 + BN): an N = 64 or 32 MMA into the matching accumulator columns
          const int n0 = (tile % n_tiles) * BN;
          for (int rb = n0 / GEMM_BLOCK_K; rb * GEMM_BLOCK_K < n0 + BN; ++rb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const int lo = max(rb * GEMM_BLOCK_K, n0), hi = min(rb * GEMM_BLOCK_K + GEMM_BLOCK_K, n0 + BN);
            const uint32_t nn = static_cast<uint32_t>(hi - lo);  // 64 or 32
            const uint32_t idesc = (1u << 4) | ((nn >> 3) << 17) | ((GEMM_BLOCK_M >> 4) << 24);  // fp16 x fp16 -> fp32, 128 x nn
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + stage * C::STAGE_BYTES));
            const uint8_t* ident = WSK > 0 ? bres + WSK * C::B_BYTES : smem + stage * C::STAGE_BYTES + C::A_BYTES;
            const uint64_t b_desc = make_smem_desc_sw128(smem_u32(ident + (lo - rb * GEMM_BLOCK_K) * (GEMM_BLOCK_K * 2)));
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
              umma_f16_ss(d_tmem + static_cast<uint32_t>(lo - n0), a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
            umma_commit(&empty_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      }
    }
  } else if (warp - 2 < EPI_ACTIVE_WARPS) {
    // =========================== epilogue ===========================
    const int q = warp & 3;            // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int half = (warp - 2) >> 2;  // column half of the tile handled by this warp
    const int row = q * 32 + lane;
    int it = 0;

    if constexpr (GRP) {
      // ============ GEGLU projection, alternating warp groups (weight-stationary build, LayerNorm-consuming or plain) ============
      // The per-tile epilogue is a serial chain (accumulator ready -> TMEM drain -> GELU arithmetic -> staging -> TMA store,
      // ~4000 cycles, tools/gemm_trace.py) while the MMAs of a 128 x 160 x 320 tile take ~3000: with all eight warps on every
      // tile the chain bounded the launch.  Here warps 2-5 take the even tiles of the CTA and warps 6-9 the odd ones (TMEM
      // buffer = group), each warp draining BOTH column halves of its rows in two passes, so two chains are in flight.
      const int grp = half;
      constexpr int HC = BN / 2;
      uint8_t* const stg_g = stg + grp * C::HALF_STG_BYTES;  // [2 sub-tiles][128][40]: 80 output columns of the group's tile
      float* const wv = vecs + (warp - 2) * 160;
      float* const wv2 = vecs + GEMM_EPI_WARPS * 160 + (warp - 2) * 160;
      const bool ln_consume = args.ln_stats != nullptr;
      const int bar_id = 1 + grp;
      const bool leader_warp = (q == 0);
      float ln_S = 0.f, ln_Q = 0.f;
      auto ln_fetch = [&](int m_) {
        ln_S = 0.f;
        ln_Q = 0.f;
        if (ln_consume && m_ < args.M) {
          const float2* st = reinterpret_cast<const float2*>(args.ln_stats) + static_cast<size_t>(m_) * args.ln_slots;
          for (int i = 0; i < args.ln_slots; ++i) {
            const float2 t = __ldg(st + i);
            ln_S += t.x;
            ln_Q += t.y;
          }
        }
      };
      for (int tile = blockIdx.x + grp * gridDim.x, itg = grp; tile < total_tiles; tile += 2 * gridDim.x, itg += 2) {
        const int n_tile = tile % n_tiles;
        const int m0 = (tile / n_tiles) * GEMM_BLOCK_M;
        const int n0 = n_tile * BN;
        const int m = m0 + row;
        const bool valid = m < args.M;
        if (itg == grp) {  // the CTA keeps its n-tile (weight-stationary) and there is no row-bias: one vector for all tiles
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const int c = lane + 32 * i;  // pass p = c / 80: values 40 p + [0, 40), then their gates (+80)
            const int pp = c / 80, cc = c % 80;
            const int col = n0 + pp * 40 + (cc < 40 ? cc : cc - 40 + HC);
            wv[c] = args.bias ? __ldg(args.bias + col) : 0.f;
            wv2[c] = ln_consume ? __ldg(args.ln_s + col) : 0.f;
          }
          __syncwarp();
        }
        // LayerNorm row statistics of this tile: fetched one tile ahead (their load latency was part of every tile's chain)
        if (itg == grp) ln_fetch(m);
        float ln_rstd = 1.f, ln_rstd_mu = 0.f;
        if (ln_consume && valid) {
          const float inv = 1.0f / static_cast<float>(args.ln_dim);
          const float mu = ln_S * inv;
          ln_rstd = rsqrtf(fmaxf(ln_Q * inv - mu * mu, 0.f) + args.ln_eps);
          ln_rstd_mu = ln_rstd * mu;
        }
        {
          const int next = tile + 2 * gridDim.x;
          if (next < total_tiles) ln_fetch((next / n_tiles) * GEMM_BLOCK_M + row);
        }
        mbar_wait(&tmem_full_bar[grp], (itg >> 1) & 1);
        tc_fence_after();
        const uint32_t t_row = tmem_base + grp * C::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
        // the group's previous TMA store (two tiles ago) has read the staging buffer
        if (leader_warp) {
          if (elect_one()) bulk_wait_group_read<0>();
        }
        named_barrier_sync(bar_id, 128);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          uint32_t v[HC];
          tmem_ld_32x32b_x16p(t_row + p * 40, v);
          tmem_ld_32x32b_x16p(t_row + p * 40 + 16, v + 16);
          tmem_ld_32x32b_x8(t_row + p * 40 + 32, v + 32);
          tmem_ld_32x32b_x16p(t_row + HC + p * 40, v + 40);
          tmem_ld_32x32b_x16p(t_row + HC + p * 40 + 16, v + 56);
          tmem_ld_32x32b_x8(t_row + HC + p * 40 + 32, v + 72);
          tmem_ld_wait();
          if (p == 1) {  // accumulator fully drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[grp]);
          }
          uint8_t* const srow = stg_g + p * C::SUB_BYTES + row * (GEMM_SUB_N * 2);
          if (ln_consume) epi_geglu40<true>(v, wv + p * 80, wv + p * 80 + 40, srow, wv2 + p * 80, wv2 + p * 80 + 40, ln_rstd, ln_rstd_mu);
          else epi_geglu40<false>(v, wv + p * 80, wv + p * 80 + 40, srow);
        }
        fence_proxy_async();
        named_barrier_sync(bar_id, 128);
        if (leader_warp && elect_one()) {
          tma_store_2d_keep(&mapOut, stg_g, n_tile * HC, m0, (args.dbg & 64) != 0);
          tma_store_2d_keep(&mapOut, stg_g + C::SUB_BYTES, n_tile * HC + GEMM_SUB_N, m0, (args.dbg & 64) != 0);
          bulk_commit_group();
        }
      }
      if (leader_warp) {
        if (elect_one()) bulk_wait_group_read<0>();  // smem must outlive the last TMA store's read
      }
    } else if constexpr (EPI == EPI_SAMPLER) {
      // per-step quantities: kernel arguments, or (graph replay) the device-resident StepParams
      const float4 coef = args.sp ? args.sp->coef : args.coef;
      const int mode = args.sp ? args.sp->mode : args.mode;
      const int use_philox = args.sp ? args.sp->use_philox : args.use_philox;
      const unsigned long long seed = args.sp ? args.sp->seed : args.seed;
      const unsigned long long sample_offset = args.sp ? args.sp->sample_offset : args.sample_offset;
      const int step_index = args.sp ? args.sp->step_index : args.step_index;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const int m = (tile / n_tiles) * GEMM_BLOCK_M + row;
        const bool valid = m < args.M;
        mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * C::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
        // ---- output conv: columns 0..3 = predicted noise of pixel m; fused sampler update (train.py:229-236) ----
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (valid) {
          const int HW = args.HWout;
          const int b = m / HW, pix = m % HW;
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            // columns 0..3: bf16 "hi" part of the weights, 4..7: their "lo" part (engine packs out.2.weight that way)
            const float eps = (__uint_as_float(v[o]) + __uint_as_float(v[4 + o])) + __ldg(args.bias + o);
            const size_t idx = (static_cast<size_t>(b) * 4 + o) * HW + pix;
            if (args.eps_out) args.eps_out[idx] = eps;
            if (mode == STEP_DDPM) {
              float z = 0.f;
              if (args.noise)
                z = __ldg(args.noise + idx);
              else if (use_philox)
                z = philox_normal(seed, sample_offset * (4ull * HW) + idx, static_cast<uint32_t>(step_index));
              const float xv = args.x[idx];
              // same op order as the reference expression, no FMA contraction
              const float inner = __fsub_rn(xv, __fmul_rn(coef.y, eps));
              args.x[idx] = __fadd_rn(__fmul_rn(coef.x, inner), __fmul_rn(coef.z, z));
            } else if (mode == STEP_DDIM) {
              const float xv = args.x[idx];
              const float x0 = __fmul_rn(__fsub_rn(xv, __fmul_rn(coef.y, eps)), coef.x);
              args.x[idx] = __fadd_rn(__fmul_rn(coef.z, x0), __fmul_rn(coef.w, eps));
            }
          }
        }
      }
    } else {
      constexpr int HC = BN / 2;             // accumulator columns per warp (80)
      constexpr int NSUB = HC / GEMM_SUB_N;  // staging sub-tiles per half: 2 (GEGLU fills only the first)
      const bool leader_warp = (q == 0);  // its elected lane issues this half's TMA stores / residual loads
      const int bar_id = 1 + half;
      const bool use_stg = (NSTG > 0) && !args.out_f32;
      // rare flavours (time-embedding GEMMs, operator tests): SiLU, fp32 output, per-thread row-bias rows, or a residual stored
      // in another 16-bit format than the output -> generic run-time-flag epilogue, residual read from global memory
      const bool slow_path = !use_stg || args.act != ACT_NONE || (args.rowbias && args.rows_per_sample % 32 != 0) ||
                             (args.residual && !(RESK && args.res_k) && (args.res_f16 != 0) != (args.out_f16 != 0));
      const bool has_res = use_stg && args.residual != nullptr && !slow_path && !(RESK && args.res_k);
      uint8_t* const stg_half = stg + half * NB * C::HALF_STG_BYTES;
      uint64_t* const res_bar = res_full_bar + half * 2;
      float* const wv = vecs + (warp - 2) * 80;
      float* const wv2 = vecs + GEMM_EPI_WARPS * 80 + (warp - 2) * 80;  // WS mode only: ln_s of this warp's columns
      const bool ln_consume = (WSK > 0) && args.ln_stats != nullptr;
      const bool out_f16 = args.out_f16 != 0, res_f16 = args.res_f16 != 0;

      auto issue_res_load = [&](int tile_, int sb_) {
        const int m0_ = (tile_ / n_tiles) * GEMM_BLOCK_M;
        const int c0_ = (tile_ % n_tiles) * BN + half * HC;
        mbar_arrive_expect_tx(&res_bar[sb_], C::HALF_STG_BYTES);
#pragma unroll
        for (int s = 0; s < NSUB; ++s)
          tma_load_2d(stg_half + sb_ * C::HALF_STG_BYTES + s * C::SUB_BYTES, &mapRes, &res_bar[sb_],
                      c0_ + s * GEMM_SUB_N, m0_);
      };
      if (has_res && leader_warp && static_cast<int>(blockIdx.x) < total_tiles) {
        if (elect_one()) issue_res_load(blockIdx.x, 0);
      }
      float std_ln_S = 0.f, std_ln_Q = 0.f;
      auto std_ln_fetch = [&](int m_) {
        std_ln_S = 0.f;
        std_ln_Q = 0.f;
        if (m_ < args.M) {
          const float2* st = reinterpret_cast<const float2*>(args.ln_stats) + static_cast<size_t>(m_) * args.ln_slots;
          for (int i = 0; i < args.ln_slots; ++i) {
            const float2 t = __ldg(st + i);
            std_ln_S += t.x;
            std_ln_Q += t.y;
          }
        }
      };
      uint4 att_pf[3];  // ATT builds: this thread's share of the next tile's K / V rows
      auto att_fetch = [&](int m0_, int n0_) {
        if constexpr (ATT) {
          const int L = args.att_L;
          const int sample = m0_ / args.rows_per_sample;
          const int head = (n0_ + half * HC) / 80;
          const __nv_bfloat16* src = args.att_kv + static_cast<size_t>(sample) * L * args.att_ld + head * 80;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int idx = q * 32 + lane + i * 128;
            att_pf[i] = make_uint4(0u, 0u, 0u, 0u);
            if (idx < 2 * 16 * 10) {
              const int sel = idx / 160, r = idx % 160, j = r / 10, ch = r % 10;
              if (j < L) att_pf[i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(j) * args.att_ld + sel * args.att_voff) + ch);
            }
          }
        }
      };
      if (ATT && static_cast<int>(blockIdx.x) < total_tiles)
        att_fetch((static_cast<int>(blockIdx.x) / n_tiles) * GEMM_BLOCK_M, (static_cast<int>(blockIdx.x) % n_tiles) * BN);

      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const int sb = it % NB;
        const int n_tile = tile % n_tiles;
        const int m0 = (tile / n_tiles) * GEMM_BLOCK_M;
        const int n0 = n_tile * BN;
        const int m = m0 + row;
        const bool valid = m < args.M;
        WD_TRACE(0);
        // ---- per-warp vector of the additive per-column terms (bias + the warp's sample row of the row-bias) ----
        const float* rb = nullptr;  // per-thread row-bias only when the rows of a warp can belong to different samples
        {
          const int mw = min(m0 + q * 32, args.M - 1);
          const float* rbw = nullptr;
          if (args.rowbias && args.rows_per_sample % 32 == 0) {
            const int sw = mw / args.rows_per_sample;
            rbw = args.rowbias + (args.rowbias_idx ? args.rowbias_idx[sw] : static_cast<long long>(sw)) * args.rb_ld;
          } else if (args.rowbias) {
            const int sample = valid ? (m / args.rows_per_sample) : 0;
            rb = args.rowbias + (args.rowbias_idx ? args.rowbias_idx[sample] : static_cast<long long>(sample)) * args.rb_ld;
          }
          // weight-stationary CTAs keep their n-tile: without a row-bias the vector is the same for every tile of the CTA
          const bool vec_static = (WSK > 0) && args.rowbias == nullptr;
          __syncwarp();  // all lanes are done reading the previous tile's vector
          if (!vec_static || it == 0) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c = lane + 32 * i;
            if (c < HC) {
              // STD: accumulator column n0 + 80 half + c.  GEGLU: c < 40 -> value column 40 half + c, else its gate (+80)
              const int col = args.geglu ? (n0 + half * 40 + (c < 40 ? c : c - 40 + HC)) : (n0 + half * HC + c);
              float x = args.bias ? __ldg(args.bias + col) : 0.f;
              if (rbw) x += __ldg(rbw + col);
              wv[c] = x;
              if constexpr (WSK > 0) {
                if (ln_consume) wv2[c] = __ldg(args.ln_s + col);
              }
            }
          }
          }
          __syncwarp();
        }

        WD_TRACE(1);
        // LayerNorm row statistics: this tile's were fetched one tile ahead (std_ln_fetch), the next tile's are requested now
        float ln_rstd = 0.f, ln_rstd_mu = 0.f;
        if (ln_consume) {
          if (it == 0) std_ln_fetch(m);
          if (valid) {
            const float inv = 1.0f / static_cast<float>(args.ln_dim);
            const float mu = std_ln_S * inv;
            ln_rstd = rsqrtf(fmaxf(std_ln_Q * inv - mu * mu, 0.f) + args.ln_eps);
            ln_rstd_mu = ln_rstd * mu;
          }
          const int next = tile + gridDim.x;
          if (next < total_tiles) std_ln_fetch((next / n_tiles) * GEMM_BLOCK_M + row);
        }

        if constexpr (ATT) {
          // K / V rows of (sample of this tile, head of this column half) -> shared memory, rows >= L zeroed.  The four warps of
          // the half passed the post-staging barrier of the previous tile after their last read of the buffer; the barrier in
          // front of the epilogue arithmetic below publishes the new contents.  The rows were fetched into registers one tile
          // ahead (att_pf: 3 x 16 bytes per thread), so their global-load latency is not part of this tile's chain.
          __nv_bfloat16* dst = kvs + half * (2 * 16 * ATT_KP);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int idx = q * 32 + lane + i * 128;
            if (idx < 2 * 16 * 10) {
              const int sel = idx / 160, r = idx % 160, j = r / 10, ch = r % 10;
              *reinterpret_cast<uint4*>(dst + (sel * 16 + j) * ATT_KP + ch * 8) = att_pf[i];
            }
          }
          const int next = tile + gridDim.x;
          if (next < total_tiles) att_fetch((next / n_tiles) * GEMM_BLOCK_M, (next % n_tiles) * BN);
        }

        WD_TRACE(2);
        mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
        tc_fence_after();
        WD_TRACE(3);
        const uint32_t t_row = tmem_base + acc * C::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);

        // ---- drain this warp's 32 x 80 accumulator block, then hand the TMEM buffer back to the MMA warp ----
        uint32_t v[HC];
        if (args.dbg & 256) {
#pragma unroll
          for (int c = 0; c < HC; ++c) v[c] = 0;
        } else if (!args.geglu) {
#pragma unroll
          for (int c = 0; c < HC / 16; ++c) tmem_ld_32x32b_x16p(t_row + half * HC + c * 16, v + c * 16);
        } else {
          // values: tile columns [half*40, +40), gates: [80 + half*40, +40)
          tmem_ld_32x32b_x16p(t_row + half * 40, v);
          tmem_ld_32x32b_x16p(t_row + half * 40 + 16, v + 16);
          tmem_ld_32x32b_x8(t_row + half * 40 + 32, v + 32);
          tmem_ld_32x32b_x16p(t_row + HC + half * 40, v + 40);
          tmem_ld_32x32b_x16p(t_row + HC + half * 40 + 16, v + 56);
          tmem_ld_32x32b_x8(t_row + HC + half * 40 + 32, v + 72);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        WD_TRACE(4);

        if (args.dbg & 2) continue;

        // ---- staging buffer `sb` of this half must be free (its previous TMA store has read it) ----
        if (use_stg) {
          if (has_res) {
            mbar_wait(&res_bar[sb], (it / NB) & 1);  // the residual tile has landed in the staging buffer
          } else if (ATT || !(args.dbg & 64)) {
            if (leader_warp) {
              if (elect_one()) {
                if (NSTG > 1) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
              }
            }
            named_barrier_sync(bar_id, 128);
          }
        }
        uint8_t* const srow = stg_half + sb * C::HALF_STG_BYTES + row * (GEMM_SUB_N * 2);
        WD_TRACE(5);

        if (args.dbg & 128) {
        } else if (ATT) {
          if constexpr (ATT) {
            const __nv_bfloat16* sK = kvs + half * (2 * 16 * ATT_KP);
            const float sl2 = args.att_scale * 1.4426950408889634f;  // scale * log2(e): the softmax runs on exp2
            uint8_t* const swarp = stg_half + sb * C::HALF_STG_BYTES + (q * 32) * (GEMM_SUB_N * 2);
            if (ln_consume) epi_ctx_attn80<true>(v, wv, wv2, ln_rstd, ln_rstd_mu, sl2, sK, sK + 16 * ATT_KP, args.att_L, swarp, C::SUB_BYTES, lane);
            else epi_ctx_attn80<false>(v, wv, wv2, 1.f, 0.f, sl2, sK, sK + 16 * ATT_KP, args.att_L, swarp, C::SUB_BYTES, lane);
          }
        } else if (args.geglu) {
          // values of this warp: tile columns [40 half, +40), gates 80 further; 40 output columns = one staging sub-tile
          // (every GEGLU launch stages: gemm_tc_launch rejects geglu + fp32 output)
          if (ln_consume) epi_geglu40<true>(v, wv, wv + 40, srow, wv2, wv2 + 40, ln_rstd, ln_rstd_mu);
          else if (use_stg) epi_geglu40<false>(v, wv, wv + 40, srow);
        } else {
          const int nb = n0 + half * HC;
          if (ln_consume) {
            epi_round80_lnc(v, wv, wv2, ln_rstd, ln_rstd_mu, srow, C::SUB_BYTES);
          } else if (slow_path) {
            epi_round80_generic(v, wv, rb ? rb + nb : nullptr, args.act == ACT_SILU,
                                (args.residual && !has_res && !(RESK && args.res_k)) ? args.residual + static_cast<size_t>(m) * args.res_ld + nb : nullptr,
                                res_f16, use_stg, srow, C::SUB_BYTES, args.out_f32 != 0, out_f16,
                                args.out_f32 ? static_cast<void*>(static_cast<float*>(args.out) + static_cast<size_t>(m) * args.out_ld + nb)
                                             : static_cast<void*>(static_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(m) * args.out_ld + nb),
                                valid);
          } else {
            float gs[16];  // GroupNorm partials: [2g] = sum, [2g+1] = sum of squares of group g (10 columns) of this row
#pragma unroll
            for (int i = 0; i < 16; ++i) gs[i] = 0.f;
            epi_round80_dispatch(has_res, args.gn_partial != nullptr, out_f16, args.ln_out != nullptr, v, wv, srow, C::SUB_BYTES,
                                 valid, gs);
            if (args.ln_out && valid)  // LayerNorm row statistics of the tensor being written (this thread's 80 columns)
              *reinterpret_cast<float2*>(args.ln_out + (static_cast<size_t>(m) * (args.N / 80) + nb / 80) * 2) = make_float2(gs[0], gs[1]);
            if (args.gn_partial) {
              // rows of a warp belong to one sample (rows_per_sample % 32 == 0): reduce over the 32 rows, lane L < 16 keeps entry L
              const float tot = warp_transpose_reduce16(gs, lane);
              const int mw = m0 + q * 32;
              if (mw < args.M && lane < 16) {
                const int smp = mw / args.rows_per_sample;
                const bool up5 = args.up_phase == 5;  // phases as N tiles: phase = n0 / 320, 32 groups per phase
                const int slot = (up5 ? (n0 / 320) * (args.rows_per_sample >> 5) : args.gn_slot_base) + ((mw % args.rows_per_sample) >> 5);
                const int nslot = args.gn_nslot ? args.gn_nslot : args.rows_per_sample >> 5;
                const int G = up5 ? 32 : args.N / 10;
                const int g = (up5 ? (n0 % 320) / 10 : n_tile * (BN / 10)) + half * (HC / 10) + (lane >> 1);
                args.gn_partial[((static_cast<size_t>(smp) * G + g) * nslot + slot) * 2 + (lane & 1)] = tot;
              }
            }
          }
        }

        // ---- publish the staged half tile with TMA; prefetch the residual of this CTA's next tile ----
        WD_TRACE(6);
        if (use_stg) {
          if (!(args.dbg & 32)) fence_proxy_async();  // generic-proxy smem writes -> visible to the async proxy (TMA)
          named_barrier_sync(bar_id, 128);
          WD_TRACE(7);
          if (leader_warp && !(args.dbg & 1) && elect_one()) {
            const uint8_t* src = stg_half + sb * C::HALF_STG_BYTES;
            if (!args.geglu) {
#pragma unroll
              for (int s = 0; s < NSUB; ++s) {
                if (args.up_phase == 5)  // (c, b, x, a, image-row) with the phase (a, b) = n0 / 320
                  tma_store_5d(&mapOut, src + s * C::SUB_BYTES, n0 % 320 + half * HC + s * GEMM_SUB_N, (n0 / 320) & 1, 0, (n0 / 320) >> 1,
                               (m0 / args.HWout) * (args.HWout / args.Wout));
                else if (args.up_phase)  // sub-pixel phase: (c, x, y, n) of the phase grid; the tile holds whole images
                  tma_store_4d(&mapOut, src + s * C::SUB_BYTES, n0 + half * HC + s * GEMM_SUB_N, 0, 0, m0 / args.HWout);
                else
                  tma_store_2d_keep(&mapOut, src + s * C::SUB_BYTES, n0 + half * HC + s * GEMM_SUB_N, m0, (args.dbg & 64) != 0);
              }
            } else {
              tma_store_2d_keep(&mapOut, src, n_tile * HC + half * 40, m0, (args.dbg & 64) != 0);
            }
            bulk_commit_group();
            const int next = tile + gridDim.x;
            if (has_res && next < total_tiles) {
              // the buffer the next tile uses was last read by the store issued NSTG tiles before it
              if (NSTG > 1) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
              issue_res_load(next, (it + 1) % NB);
            }
          }
          WD_TRACE(8);
        }
      }
      if (use_stg && leader_warp) {
        if (elect_one()) bulk_wait_group_read<0>();  // smem must outlive the last TMA store's read
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if ((args.dbg & 1024) && threadIdx.x == 64 && blockIdx.x < 160) g_gemm_cta_times[blockIdx.x * 4 + 3] = gtimer();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

bool tmap_encode_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                         uint32_t box_inner, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[wd_b200] cuTensorMapEncodeTiled(2d) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

bool tmap_encode_4d_bf16(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N,
                         uint64_t pix_stride_elems, uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n,
                         uint32_t stride_wh) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {pix_stride_elems * 2, pix_stride_elems * 2 * W, pix_stride_elems * 2 * W * H};
  cuuint32_t box[4] = {box_c, box_w, box_h, box_n};
  cuuint32_t es[4] = {1, stride_wh, stride_wh, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[wd_b200] cuTensorMapEncodeTiled(4d) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

bool tmap_encode_out_phase_bf16(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn || W * H == 0 || GEMM_BLOCK_M % (W * H)) return false;
  cuuint64_t dims[4] = {C, W, H, N};
  // the phase grid steps two pixels / two rows of the [N, 2H, 2W, C] tensor
  cuuint64_t strides[3] = {2 * C * 2, 2 * (2 * W) * C * 2, (2 * H) * (2 * W) * C * 2};
  cuuint32_t box[4] = {GEMM_SUB_N, static_cast<cuuint32_t>(W), static_cast<cuuint32_t>(H), static_cast<cuuint32_t>(GEMM_BLOCK_M / (W * H))};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[wd_b200] cuTensorMapEncodeTiled(out phase) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

bool tmap_encode_out_phase5_bf16(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn || W == 0 || H == 0 || GEMM_BLOCK_M % W || GEMM_BLOCK_M % (W * H)) return false;
  // pixel (n, 2 y + a, 2 x + b): address / (2 C bytes) = b/2 ... written as ascending strides: b: C, x: 2 C, a: 2 W C, (n H + y): 4 W C
  cuuint64_t dims[5] = {C, 2, W, 2, N * H};
  cuuint64_t strides[4] = {C * 2, 2 * C * 2, 2 * W * C * 2, 4 * W * C * 2};
  cuuint32_t box[5] = {GEMM_SUB_N, 1, static_cast<cuuint32_t>(W), 1, static_cast<cuuint32_t>(GEMM_BLOCK_M / W)};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[wd_b200] cuTensorMapEncodeTiled(out phase5) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

bool tmap_encode_3d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t row_stride_elems,
                         uint64_t batch_stride_elems, uint32_t box_inner, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[wd_b200] cuTensorMapEncodeTiled(3d) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

bool tmap_encode_out_bf16(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_elems) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {GEMM_SUB_N, GEMM_BLOCK_M};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[wd_b200] cuTensorMapEncodeTiled(out) failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int EPI, int STAGES, int NSTG, int WSK = 0, int ATT = 0, int RESK = 0, int GRP = 0>
static cudaError_t launch_impl(const GemmLaunch& L, cudaStream_t stream) {
  using C = Cfg<BN, STAGES, NSTG, WSK, ATT, RESK, GRP>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, STAGES, NSTG, WSK, ATT, RESK, GRP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const GemmArgs& a = L.args;
  if (a.N % BN != 0 || a.M <= 0) return cudaErrorInvalidValue;
  const int n_tiles = a.N / BN;
  const int tiles = n_tiles * ((a.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (WSK > 0) {
    // one n-tile per CTA for its whole life: the grid is a multiple of n_tiles (tiles is one already)
    if (n_tiles > num_sms()) return cudaErrorInvalidValue;
    grid = (grid / n_tiles) * n_tiles;
  }
  return launch_pdl(gemm_tc_kernel<BN, EPI, STAGES, NSTG, WSK, ATT, RESK, GRP>, dim3(grid), dim3(GEMM_THREADS), C::SMEM_BYTES, stream, L.mapA[0],
                    L.mapA[1], L.mapA[2], L.mapB, L.mapOut, L.mapRes, a);
}

static bool gemm_ws_enabled();
// 64 x 64 fp16 identity (one per device), the B operand of the residual K blocks
static const void* gemm_identity_f16() {
  static void* ident[64] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!ident[dev]) {
    std::vector<uint16_t> h(64 * 64, 0);
    for (int i = 0; i < 64; ++i) h[i * 64 + i] = 0x3C00;  // 1.0 in fp16
    void* d = nullptr;
    if (cudaMalloc(&d, h.size() * 2) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    ident[dev] = d;
  }
  return ident[dev];
}

static bool gemm_grp_enabled() {  // env WD_GEMM_GRP (default on): alternating epilogue warp groups for the GEGLU projection
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_GEMM_GRP");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
static bool gemm_res_k_enabled() {  // env WD_GEMM_RESK (default on)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_GEMM_RESK");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

bool gemm_prepare_res_k(GemmLaunch& L) {
  GemmArgs& a = L.args;
  a.res_k = 0;
  int total_k = 0;
  for (int s = 0; s < a.num_src; ++s) total_k += a.taps[s] * a.chunks[s];
  // the weight-stationary launch of gemm_tc_launch() (K <= 320) or its long-K single-CTA launch (K >= 1024: ff_out), plain Linear,
  // an fp16 residual covering the N output columns, 16-bit staged output
  const bool ws_shape = gemm_ws_enabled() && total_k <= 5 && a.N / GEMM_BLOCK_N <= num_sms();
  const bool long_shape = total_k >= 16;
  if (!gemm_res_k_enabled() || !(ws_shape || long_shape) || !a.residual || !a.res_f16 || a.epi != EPI_STD || a.conv || a.num_src != 1 ||
      a.N % GEMM_BLOCK_K % 32 || a.out_f32 || a.geglu || a.act != ACT_NONE || a.ln_stats || a.att_kv || a.res_ld % 8 ||
      (a.rowbias && a.rows_per_sample % 32 != 0) || gemm_uses_pair(a))
    return true;
  const void* ident = gemm_identity_f16();
  if (!ident) return false;
  // residual [M, N] (row stride res_ld) as an A operand: 64-channel x 128-row SWIZZLE_128B boxes; columns beyond N are zero-filled
  if (!tmap_encode_2d_bf16(&L.mapA[1], a.residual, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M), static_cast<uint64_t>(a.res_ld),
                           GEMM_BLOCK_K, GEMM_BLOCK_M))
    return false;
  if (!tmap_encode_2d_bf16(&L.mapA[2], ident, 64, 64, 64, GEMM_BLOCK_K, 64)) return false;
  a.res_k = 1;
  return true;
}

}  // namespace wd
// debug aid (not part of the product ABI): copies the clock64 stamps recorded under WD_GEMM_DBG & 1024
extern "C" int wdx_gemm_trace_read(unsigned long long* host, int n) {
  if (n > 12 * 64) n = 12 * 64;
  return cudaMemcpyFromSymbol(host, wd::g_gemm_trace, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : -1;
}
extern "C" int wdx_gemm_cta_times_read(unsigned long long* host) {
  return cudaMemcpyFromSymbol(host, wd::g_gemm_cta_times, sizeof(unsigned long long) * 4 * 160) == cudaSuccess ? 0 : -1;
}
extern "C" int wdx_gemm_trace_clear(void) {
  static unsigned long long z[12 * 64] = {0};
  return cudaMemcpyToSymbol(wd::g_gemm_trace, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
namespace wd {
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
static int gemm_dbg_flags() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_GEMM_DBG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static bool gemm_ws_enabled() {  // env WD_GEMM_WS (default on): weight-stationary mode for K <= 320
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_GEMM_WS");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
bool gemm_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_GEMM_PAIR");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
// Which GEMMs run on the CTA-pair kernel (256 x 320 tiles): by the number of 64-wide K blocks (env WD_GEMM_PAIR_MINK overrides).
// Measured per op inside the step (bench.py --ops-out, r02a): the pair kernel wins from 50 K blocks on (conv2 + fused skip conv:
// 123 -> 116 us, the 640-channel convs), and already from 40 on the 4 x 16 level (M = 16384), where 256 single-CTA tiles fill the
// 148 SMs 1.73 times but 64 pair tiles fit one wave (36 -> 34 us, 41 -> 37 us); the 320-channel 3 x 3 convs of the 8 x 32 level
// (45 K blocks, 6.9 waves of single-CTA tiles) stay on the single-CTA kernel.
static int gemm_pair_min_kblocks(int M) {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("WD_GEMM_PAIR_MINK");
    v = e ? atoi(e) : -1;
  }
  if (v >= 0) return v;
  // fewer than 48 pair tiles (0.65 of a wave of 74 pairs): the single-CTA kernel's 128 x 160 tiles fill 4x as many SMs and walk a K
  // block in half the time -- unet step at batch 32 / 8 / 1: 0.989 -> 0.840, 0.947 -> 0.796, 0.872 -> 0.709 ms (tools/ab_r4B.sh)
  if (M < 48 * 256) return 1 << 30;
  return M <= 32768 ? 40 : 50;
}
bool gemm_uses_pair(const GemmArgs& a) {
  if (!gemm_pair_enabled() || !gemm_pair_supported(a)) return false;
  if (a.gn_apply) return true;  // the producer-side GroupNorm exists in the pair kernel only (the caller checked the shape)
  int total_k = 0;
  for (int s = 0; s < a.num_src; ++s) total_k += a.taps[s] * a.chunks[s];
  return total_k >= gemm_pair_min_kblocks(a.M);
}
int gemm_b_box_rows(const GemmArgs& a) {
  if (a.epi == EPI_SAMPLER) return GEMM_BLOCK_N_OUT;
  return gemm_uses_pair(a) ? 80 : GEMM_BLOCK_N;
}
int gemm_geglu_block(int) { return GEMM_BLOCK_N; }  // GEGLU projections (K = 320) run on the single-CTA kernel: value | gate per 160-column tile

cudaError_t gemm_tc_launch(const GemmLaunch& L0, cudaStream_t stream) {
  GemmLaunch L = L0;
  L.args.dbg = gemm_dbg_flags();
  const GemmArgs& a = L.args;
  if (a.epi == EPI_SAMPLER) {
    if (a.N != GEMM_BLOCK_N_OUT || !a.bias || !a.conv) return cudaErrorInvalidValue;
    return launch_impl<GEMM_BLOCK_N_OUT, EPI_SAMPLER, 8, 0>(L, stream);
  }
  if (a.gn_partial && (a.gn_cpg != 10 || a.rows_per_sample % 32 || a.geglu || a.N % 10)) return cudaErrorInvalidValue;
  if (a.gn_apply && !gemm_uses_pair(a)) return cudaErrorInvalidValue;  // the producer-side GroupNorm lives in the pair kernel only
  if (a.geglu && (a.residual || a.out_f32 || a.out_f16)) return cudaErrorInvalidValue;
  if (a.ln_out && (a.gn_partial || !a.out_f16 || a.out_f32 || a.geglu || a.act != ACT_NONE || a.N % 80)) return cudaErrorInvalidValue;
  if (a.ln_stats && (a.residual || a.gn_partial || a.out_f32 || a.out_f16 || a.rowbias || a.act != ACT_NONE || !a.ln_s ||
                     a.ln_slots < 1 || a.conv || a.num_src != 1))
    return cudaErrorInvalidValue;
  int total_k = 0;
  for (int s = 0; s < a.num_src; ++s) total_k += a.taps[s] * a.chunks[s];
  if (a.att_kv) {
    // to_q projection with the context attention in its epilogue: weight-stationary build only, whole tiles inside one sample
    if (a.conv || a.num_src != 1 || total_k > 5 || a.N % 80 || a.N / GEMM_BLOCK_N > num_sms() || a.M % GEMM_BLOCK_M ||
        a.rows_per_sample % GEMM_BLOCK_M || a.att_L < 1 || a.att_L > GEMM_ATT_MAXL || a.att_ld % 8 || a.att_voff % 8 || a.residual ||
        a.gn_partial || a.ln_out || a.out_f32 || a.out_f16 || a.rowbias || a.geglu || a.act != ACT_NONE || (a.ln_stats && !a.ln_s))
      return cudaErrorInvalidValue;
    return launch_impl<GEMM_BLOCK_N, EPI_STD, 4, 1, 5, 1>(L, stream);
  }
  if (gemm_uses_pair(a)) return gemm_pair_launch(L, num_sms(), stream);
  // long K loops hide the epilogue behind the MMAs of the next tile: spend shared memory on operand stages;
  // short K loops are epilogue / store bound: spend it on a second staging buffer
  if (total_k >= 16) {
    if (a.res_k) return launch_impl<GEMM_BLOCK_N, EPI_STD, 5, 1, 0, 0, 1>(L, stream);  // + residual as identity K blocks (streamed identity)
    return launch_impl<GEMM_BLOCK_N, EPI_STD, 5, 1>(L, stream);
  }
  if ((gemm_ws_enabled() || a.ln_stats) && !a.conv && a.num_src == 1 && total_k <= 5 && a.N / GEMM_BLOCK_N <= num_sms()) {
    if (a.res_k) return launch_impl<GEMM_BLOCK_N, EPI_STD, 4, 1, 5, 0, 1>(L, stream);  // + residual as identity K blocks
    if (a.geglu && !a.rowbias && gemm_grp_enabled()) return launch_impl<GEMM_BLOCK_N, EPI_STD, 4, 1, 5, 0, 0, 1>(L, stream);
    return launch_impl<GEMM_BLOCK_N, EPI_STD, 5, 1, 5>(L, stream);  // weight-stationary short-K GEMM
  }
  if (a.res_k) return cudaErrorInvalidValue;
  if (a.ln_stats) return cudaErrorInvalidValue;  // the LayerNorm-consuming epilogue exists in the weight-stationary build only
  return launch_impl<GEMM_BLOCK_N, EPI_STD, 4, 2>(L, stream);
}

}  // namespace wd
