// tcgen05 weight-gradient kernel, see wgrad_tc.cuh for the contract.
#include "wgrad_tc.cuh"

#include <cstdlib>
#include <mutex>

namespace wd {

namespace {

constexpr int WG_THREADS = 192;  // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-5: epilogue
constexpr int WG_BOX_BYTES = 64 * WG_BLOCK_TOK * 2;  // one [64 channel x 64 token] box = 8 KB

template <int BN>
struct WCfg {
  static constexpr int STAGES = (BN == 320) ? 3 : 6;
  static constexpr int X_BYTES = (WG_BLOCK_C / 64) * WG_BOX_BYTES;  // 16 KB
  static constexpr int DY_BYTES = (BN / 64) * WG_BOX_BYTES;         // 40 KB (BN = 320) / 8 KB (BN = 64)
  static constexpr int STAGE_BYTES = X_BYTES + DY_BYTES;
  static constexpr int TMEM_COLS = (BN == 320) ? 512 : 64;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;
};

// Shared-memory matrix descriptor, MN-major operand, SWIZZLE_128B, 16-bit elements.  The tile is a row of [64 channel x 64
// token] TMA boxes: inside a box a token is one 128-byte row (64 channels, swizzled in 8-row / 1024-byte atoms).
//   canonical layout (units of 16 B): ((8, n), (8, k)) : ((1, LBO), (8, SBO))
//   LBO = distance between 64-channel chunks (= one box, 8192 B), SBO = distance between 8-token groups (1024 B)
WD_DEVINL uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(WG_BOX_BYTES >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32_mn(uint32_t M, uint32_t N) {
  return make_idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);  // a_major = b_major = MN
}

WD_DEVINL void red_add_f32(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(p), "f"(v) : "memory"); }

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY, const WgradArgs args) {
  using C = WCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* acc_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- decode the work item ----
  const int cin_tiles = (args.Cin + WG_BLOCK_C - 1) / WG_BLOCK_C;
  int item = blockIdx.x;
  const int group = item % args.n_groups;
  item /= args.n_groups;
  const int ctile = item % cin_tiles;
  item /= cin_tiles;
  const int tap = item % args.taps;
  const int split = item / args.taps;
  // the last channel slice is shifted back so that it stays inside the tensor; its already-covered rows are masked below
  const int c_nom = ctile * WG_BLOCK_C;
  const int c0 = min(c_nom, args.Cin - WG_BLOCK_C);
  const int kb_total = (args.M + WG_BLOCK_TOK - 1) / WG_BLOCK_TOK;
  const int kb_per = (kb_total + args.splits - 1) / args.splits;
  const int kb_begin = split * kb_per;
  const int kb_end = min(kb_begin + kb_per, kb_total);
  const int nkb = kb_end - kb_begin;
  if (nkb <= 0) return;  // uniform over the CTA

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapDY);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const int dy = (args.taps == 9) ? tap / 3 - 1 : 0;
      const int dx = (args.taps == 9) ? tap % 3 - 1 : 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int tok0 = kb * WG_BLOCK_TOK;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
        uint8_t* sX = smem + stage * C::STAGE_BYTES;
        uint8_t* sDY = sX + C::X_BYTES;
        if (args.conv) {
          const int img = tok0 / args.HWout;
          const int oh0 = (tok0 % args.HWout) / args.Wout;
#pragma unroll
          for (int j = 0; j < WG_BLOCK_C / 64; ++j)
            tma_load_4d(sX + j * WG_BOX_BYTES, &mapX, &full_bar[stage], c0 + j * 64, dx, oh0 * args.stride + dy, img);
        } else {
#pragma unroll
          for (int j = 0; j < WG_BLOCK_C / 64; ++j)
            tma_load_2d(sX + j * WG_BOX_BYTES, &mapX, &full_bar[stage], c0 + j * 64, tok0);
        }
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_2d(sDY + j * WG_BOX_BYTES, &mapDY, &full_bar[stage], group * BN + j * 64, tok0);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t x_addr = smem_u32(smem + stage * C::STAGE_BYTES);
        const uint64_t a_desc = make_smem_desc_mn_sw128(x_addr);
        const uint64_t b_desc = make_smem_desc_mn_sw128(x_addr + C::X_BYTES);
#pragma unroll
        for (int k = 0; k < WG_BLOCK_TOK / 16; ++k) {
          // 16 tokens = two 8-token swizzle atoms = 2048 B further along K: +128 in the (addr >> 4) field
          const uint32_t accum = (i | k) != 0;
          if constexpr (BN == 320) {
            umma_f16_ss(tmem_base, a_desc + 128 * k, b_desc + 128 * k, make_idesc_bf16_f32_mn(128, 192), accum);
            umma_f16_ss(tmem_base + 192, a_desc + 128 * k, b_desc + 128 * k + ((3 * WG_BOX_BYTES) >> 4),
                        make_idesc_bf16_f32_mn(128, 128), accum);
          } else {
            umma_f16_ss(tmem_base, a_desc + 128 * k, b_desc + 128 * k, make_idesc_bf16_f32_mn(128, BN), accum);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_bar);
    }
  } else {
    // ---- epilogue: TMEM lane = X channel, TMEM column = dY channel ----
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int c = c0 + r;
    const bool valid = c >= c_nom && c < args.Cin;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* const dst = args.dst[group] + static_cast<long long>(c) * args.sC + static_cast<long long>(tap) * args.sT;
#pragma unroll 1
    for (int cb = 0; cb < BN / 16; ++cb) {
      if (cb * 16 >= args.n_valid) break;  // warp-uniform
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_row + cb * 16, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = cb * 16 + j;
          if (n < args.n_valid) red_add_f32(dst + static_cast<long long>(n) * args.sN, __uint_as_float(v[j]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN>
cudaError_t launch_bn(const WgradLaunch& L, cudaStream_t stream) {
  using C = WCfg<BN>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const WgradArgs& a = L.args;
  const int cin_tiles = (a.Cin + WG_BLOCK_C - 1) / WG_BLOCK_C;
  const int items = a.n_groups * cin_tiles * a.taps * a.splits;
  wgrad_tc_kernel<BN><<<items, WG_THREADS, C::SMEM_BYTES, stream>>>(L.mapX, L.mapDY, a);
  return cudaGetLastError();
}

}  // namespace

int wgrad_pick_splits(int M, int items_base) {
  // One full wave: the largest token split whose work items still fit the 148 SMs at once (each item ends with 128 x 320 scattered
  // fp32 atomics, a fixed cost per item, and a second, partial wave doubles the launch's MMA time), but at least 8 token blocks
  // (512 tokens) per item.  Measured (tools/ab_r4w.sh): the former "2 items per SM" (297 items for a 320 -> 320 conv) against 135
  // items: training step 15.4 -> 14.4 ms at batch 224, 6.19 -> 5.89 ms at batch 28.  WD_WGRAD_TARGET_ITEMS overrides the item budget.
  const int kb_total = (M + WG_BLOCK_TOK - 1) / WG_BLOCK_TOK;
  static const int target = [] {
    const char* e = getenv("WD_WGRAD_TARGET_ITEMS");
    return e && atoi(e) > 0 ? atoi(e) : 0;
  }();
  int s = target > 0 ? (target + items_base - 1) / items_base : 148 / (items_base > 0 ? items_base : 1);
  const int max_s = kb_total / 8 > 0 ? kb_total / 8 : 1;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}

cudaError_t wgrad_tc_launch(const WgradLaunch& L, cudaStream_t stream) {
  const WgradArgs& a = L.args;
  if (a.M <= 0 || a.Cin < WG_BLOCK_C || a.Cin % 64 || a.n_groups < 1 || a.n_groups > WG_MAX_GROUPS || a.splits < 1 ||
      (a.taps != 1 && a.taps != 9))
    return cudaErrorInvalidValue;
  if (L.bn == 320) return launch_bn<320>(L, stream);
  if (L.bn == 64) return launch_bn<64>(L, stream);
  return cudaErrorInvalidValue;
}

}  // namespace wd
