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
// sub-tiles, TMA stores, TMA-prefetched residual), with two rounds of 80 columns per warp.
#include "gemm_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace wd {

namespace {

constexpr int PAIR_BN = 320;        // tile columns (both CTAs)
constexpr int PAIR_STAGES = 4;
constexpr int PAIR_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;      // 16 KB: this CTA's 128 rows
constexpr int PAIR_BH_BYTES = 80 * GEMM_BLOCK_K * 2;               // 10 KB: 80 weight rows (half of a 160-column MMA)
constexpr int PAIR_STAGE_BYTES = PAIR_A_BYTES + 2 * PAIR_BH_BYTES;  // 36 KB
constexpr int PAIR_SUB_BYTES = GEMM_BLOCK_M * GEMM_SUB_N * 2;       // dense [128][40] bf16
constexpr int PAIR_STG_BYTES = 8 * PAIR_SUB_BYTES;                  // 80 KB: the CTA's whole 128 x 320 bf16 output tile
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + PAIR_STG_BYTES + 1024 + 256;
constexpr int PAIR_TMEM_COLS = 512;
static_assert(PAIR_SMEM_BYTES <= 227 * 1024, "shared memory budget");

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

}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapRes,
                 const GemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + PAIR_STAGES * PAIR_STAGE_BYTES;  // [8 sub-tiles][128][40] bf16 (sub-tile s = columns 40 s ..)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + PAIR_STG_BYTES);  // leader's are used by both CTAs
  uint64_t* empty_bar = full_bar + PAIR_STAGES;                            // local (multicast commit)
  uint64_t* tmem_full_bar = empty_bar + PAIR_STAGES;                       // local (multicast commit)
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;                            // leader's: 16 epilogue-warp arrivals
  uint64_t* res_full_bar = tmem_empty_bar + 1;                             // [2 halves], local
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int n_tiles = args.N / PAIR_BN;
  const int m_tiles = (args.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
  const int total_tiles = n_tiles * m_tiles;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

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
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<PAIR_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
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
            const int dy = (taps == 9) ? tap / 3 - 1 : 0;
            const int dx = (taps == 9) ? tap % 3 - 1 : 0;
            for (int ch = 0; ch < chunks; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);  // the leader's full barrier
              if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * PAIR_STAGE_BYTES);
              uint8_t* sA = smem + stage * PAIR_STAGE_BYTES;
              uint8_t* sB = sA + PAIR_A_BYTES;
              if (args.conv)
                tma_load_4d_pair(sA, mapA, fb, ch * GEMM_BLOCK_K, dx, oh0 * st + dy, img);
              else
                tma_load_2d_pair(sA, mapA, fb, ch * GEMM_BLOCK_K, m0);
              // weight rows of MMA j (columns n0 + 160 j ..): this CTA supplies rows [80 rank, +80) of them
              tma_load_2d_pair(sB, &mapB, fb, kb * GEMM_BLOCK_K, n0 + static_cast<int>(rank) * 80);
              tma_load_2d_pair(sB + PAIR_BH_BYTES, &mapB, fb, kb * GEMM_BLOCK_K, n0 + 160 + static_cast<int>(rank) * 80);
              ++kb;
              if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA, single thread) ===========================
    if (is_leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * GEMM_BLOCK_M, 160);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
        mbar_wait_cluster(tmem_empty_bar, (it & 1) ^ 1);  // both CTAs' epilogues have drained the accumulator
        tc_fence_after();
        for (int kb = 0; kb < total_k; ++kb) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * PAIR_STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc0 = make_smem_desc_sw128(a_addr + PAIR_A_BYTES);
          const uint64_t b_desc1 = make_smem_desc_sw128(a_addr + PAIR_A_BYTES + PAIR_BH_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            umma_f16_ss_pair(tmem_base, a_desc + 2 * k, b_desc0 + 2 * k, idesc, (kb | k) != 0);
            umma_f16_ss_pair(tmem_base + 160, a_desc + 2 * k, b_desc1 + 2 * k, idesc, (kb | k) != 0);
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
    const bool leader_thr = (q == 0) && (lane == 0);  // issues this half's TMA stores / residual loads
    const int bar_id = 1 + half;
    const bool use_stg = !args.out_f32;
    const bool has_res = use_stg && args.residual != nullptr;
    uint8_t* const stg_half = stg + half * 4 * PAIR_SUB_BYTES;  // 4 sub-tiles = 160 columns
    const uint32_t te_addr = mapa_shared(smem_u32(tmem_empty_bar), 0);
    const int out_half_cols = args.geglu ? 80 : 160;  // output columns written by this half per tile

    auto issue_res_load = [&](int tile_) {
      const int m0_ = (tile_ / n_tiles) * (2 * GEMM_BLOCK_M) + static_cast<int>(rank) * GEMM_BLOCK_M;
      const int c0_ = (tile_ % n_tiles) * PAIR_BN + half * 160;
      mbar_arrive_expect_tx(&res_full_bar[half], 4 * PAIR_SUB_BYTES);
#pragma unroll
      for (int s = 0; s < 4; ++s)
        tma_load_2d(stg_half + s * PAIR_SUB_BYTES, &mapRes, &res_full_bar[half], c0_ + s * GEMM_SUB_N, m0_);
    };
    if (has_res && leader_thr && pair < total_tiles) issue_res_load(pair);

    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
      const int n_tile = tile % n_tiles;
      const int m0 = (tile / n_tiles) * (2 * GEMM_BLOCK_M) + static_cast<int>(rank) * GEMM_BLOCK_M;
      const int n0 = n_tile * PAIR_BN;
      const int m = m0 + row;
      const bool valid = m < args.M;
      const int sample = valid ? (m / args.rows_per_sample) : 0;
      const float* rb = nullptr;
      if (args.rowbias) {
        const long long r = args.rowbias_idx ? args.rowbias_idx[sample] : static_cast<long long>(sample);
        rb = args.rowbias + r * args.rb_ld;
      }

      mbar_wait(tmem_full_bar, it & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

      // the staging buffer of this half must be free: its previous TMA store has read it (or the residual has landed)
      if (use_stg) {
        if (has_res) {
          mbar_wait(&res_full_bar[half], it & 1);
        } else {
          if (leader_thr) bulk_wait_group_read<0>();
          named_barrier_sync(bar_id, 128);
        }
      }
      uint8_t* const srow = stg_half + row * (GEMM_SUB_N * 2);

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
          if (lane == 0) mbar_arrive_cluster(te_addr);
        }

        if (!args.geglu) {
          float gs[16];  // GroupNorm partials: [2g] = sum, [2g+1] = sum of squares of group g (10 columns) of this row
          if (args.gn_partial) {
#pragma unroll
            for (int i = 0; i < 16; ++i) gs[i] = 0.f;
          }
          const int nb = n0 + half * 160 + rnd * 80;
#pragma unroll
          for (int c = 0; c < 10; ++c) {  // 8 columns = one 16-byte staging chunk
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[c * 8 + j]);
            if (args.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.bias + nb + c * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.bias + nb + c * 8 + 4));
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (rb) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb + nb + c * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb + nb + c * 8 + 4));
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            uint4* sp = reinterpret_cast<uint4*>(srow + (rnd * 2 + c / 5) * PAIR_SUB_BYTES + (c % 5) * 16);
            if (has_res) {
              const uint4 r4 = *sp;
              const uint32_t ru[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t = unpack_bf16x2(ru[j]);
                f[2 * j] += t.x;
                f[2 * j + 1] += t.y;
              }
            }
            if (args.act == ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
            }
            if (args.gn_partial) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int g = (c * 8 + j) / 10;  // compile-time after unrolling (80 columns -> 8 groups of 10)
                const float x = valid ? f[j] : 0.f;
                gs[2 * g] += x;
                gs[2 * g + 1] = fmaf(x, x, gs[2 * g + 1]);
              }
            }
            if (use_stg) {
              *sp = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            } else if (valid) {  // fp32 output (emb_layers GEMM): direct stores
              float4* op = reinterpret_cast<float4*>(static_cast<float*>(args.out) + static_cast<size_t>(m) * args.out_ld + nb + c * 8);
              op[0] = make_float4(f[0], f[1], f[2], f[3]);
              op[1] = make_float4(f[4], f[5], f[6], f[7]);
            }
          }
          if (args.gn_partial) {
            // rows of a warp belong to one sample (rows_per_sample % 32 == 0): reduce over the 32 rows, lane L < 16 keeps entry L
            const float tot = warp_transpose_reduce16(gs, lane);
            const int mw = m0 + q * 32;
            if (mw < args.M && lane < 16) {
              const int smp = mw / args.rows_per_sample;
              const int slot = (mw % args.rows_per_sample) >> 5;
              const int nslot = args.rows_per_sample >> 5;
              const int G = args.N / 10;
              const int g = (nb / 10) + (lane >> 1);
              args.gn_partial[((static_cast<size_t>(smp) * G + g) * nslot + slot) * 2 + (lane & 1)] = tot;
            }
          }
        } else {
          // GEGLU: out = (value + bv) * gelu(gate + bg); 40 output columns per round = one staging sub-tile
          const int nbv = n0 + half * 80 + rnd * 40;  // bias index of the value columns inside the permuted layout
          const int nbg = nbv + 160;                  // ... of the gate columns
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            float f[8];
            float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f), bv1 = bv0, bg0 = bv0, bg1 = bv0;
            if (args.bias) {
              bv0 = __ldg(reinterpret_cast<const float4*>(args.bias + nbv + c * 8));
              bv1 = __ldg(reinterpret_cast<const float4*>(args.bias + nbv + c * 8 + 4));
              bg0 = __ldg(reinterpret_cast<const float4*>(args.bias + nbg + c * 8));
              bg1 = __ldg(reinterpret_cast<const float4*>(args.bias + nbg + c * 8 + 4));
            }
            const float bv[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
            const float bg[8] = {bg0.x, bg0.y, bg0.z, bg0.w, bg1.x, bg1.y, bg1.z, bg1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
              f[j] = (__uint_as_float(v[c * 8 + j]) + bv[j]) * gelu_fast_f(__uint_as_float(v[40 + c * 8 + j]) + bg[j]);
            *reinterpret_cast<uint4*>(srow + rnd * PAIR_SUB_BYTES + c * 16) =
                make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          }
        }
      }

      // ---- publish the staged half tile with TMA; prefetch the residual of this CTA's next tile ----
      if (use_stg) {
        fence_proxy_async();  // generic-proxy smem writes -> visible to the async proxy (TMA)
        named_barrier_sync(bar_id, 128);
        if (leader_thr) {
          const int oc0 = args.geglu ? (n_tile * 160 + half * 80) : (n0 + half * 160);
          const int nsub = out_half_cols / GEMM_SUB_N;
          for (int s = 0; s < nsub; ++s) tma_store_2d(&mapOut, stg_half + s * PAIR_SUB_BYTES, oc0 + s * GEMM_SUB_N, m0);
          bulk_commit_group();
          const int next = tile + npairs;
          if (has_res && next < total_tiles) {
            bulk_wait_group_read<0>();  // the store above has read the staging buffer
            issue_res_load(next);
          }
        }
      }
    }
    if (use_stg && leader_thr) bulk_wait_group_read<0>();  // smem must outlive the last TMA store's read
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while its peer may still read its smem / signal its barriers
  if (warp == 1) tmem_dealloc_pair<PAIR_TMEM_COLS>(tmem_base);
}

bool gemm_pair_supported(const GemmArgs& a) {
  if (a.epi != EPI_STD) return false;
  if (a.N % PAIR_BN) return false;
  if (a.geglu && (a.residual || a.out_f32 || a.gn_partial)) return false;
  if (a.out_f32 && a.residual) return false;
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
  gemm_pair_kernel<<<2 * pairs, GEMM_THREADS, PAIR_SMEM_BYTES, stream>>>(L.mapA[0], L.mapA[1], L.mapA[2], L.mapB, L.mapOut,
                                                                        L.mapRes, a);
  return cudaGetLastError();
}

}  // namespace wd
