// Persistent warp-specialised tcgen05 implicit-GEMM kernel, see gemm_tc.cuh for the contract.
#include "gemm_tc.cuh"

#include <cstdio>
#include <mutex>

namespace wd {

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == GEMM_BLOCK_N) ? 6 : 8;
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BN * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = (BN == GEMM_BLOCK_N) ? 256 : 32;  // TMEM column offset of accumulator buffer 1
  static constexpr int TMEM_COLS = (BN == GEMM_BLOCK_N) ? 512 : 64;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// lane L ends with the sum over the warp's 32 lanes of v[L] (v is destroyed): 31 shuffles
WD_DEVINL float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
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

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
               const GemmArgs args) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = args.N / BN;
  const int m_tiles = (args.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int total_tiles = n_tiles * m_tiles;

  int total_k = 0;
#pragma unroll
  for (int s = 0; s < GEMM_MAX_SRC; ++s)
    if (s < args.num_src) total_k += args.taps[s] * args.chunks[s];

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA0);
    if (args.num_src > 1) tma_prefetch_desc(&mapA1);
    if (args.num_src > 2) tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapB);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
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
            const int dy = (taps == 9) ? tap / 3 - 1 : 0;
            const int dx = (taps == 9) ? tap % 3 - 1 : 0;
            for (int ch = 0; ch < chunks; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
              uint8_t* sA = smem + stage * C::STAGE_BYTES;
              uint8_t* sB = sA + C::A_BYTES;
              if (args.conv)
                tma_load_4d(sA, mapA, &full_bar[stage], ch * GEMM_BLOCK_K, dx, oh0 * st + dy, img);
              else
                tma_load_2d(sA, mapA, &full_bar[stage], ch * GEMM_BLOCK_K, m0);
              tma_load_2d(sB, &mapB, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
              ++kb;
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (single thread) ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(GEMM_BLOCK_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
        for (int kb = 0; kb < total_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzled row: +2 in the (addr >> 4) field
            umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // =========================== epilogue (4 warps, one TMEM lane quarter each) ===========================
    const int q = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int n_tile = tile % n_tiles;
      const int m0 = (tile / n_tiles) * GEMM_BLOCK_M;
      const int n0 = n_tile * BN;
      const int m = m0 + row;
      const bool valid = m < args.M;
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * C::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);

      if constexpr (EPI == EPI_SAMPLER) {
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
            const float eps = __uint_as_float(v[o]) + __ldg(args.bias + o);
            const size_t idx = (static_cast<size_t>(b) * 4 + o) * HW + pix;
            if (args.eps_out) args.eps_out[idx] = eps;
            if (args.mode == STEP_DDPM) {
              float z = 0.f;
              if (args.noise)
                z = __ldg(args.noise + idx);
              else if (args.use_philox)
                z = philox_normal(args.seed, args.sample_offset * (4ull * HW) + idx, static_cast<uint32_t>(args.step_index));
              const float xv = args.x[idx];
              // same op order as the reference expression, no FMA contraction
              const float inner = __fsub_rn(xv, __fmul_rn(args.coef.y, eps));
              args.x[idx] = __fadd_rn(__fmul_rn(args.coef.x, inner), __fmul_rn(args.coef.z, z));
            } else if (args.mode == STEP_DDIM) {
              const float xv = args.x[idx];
              const float x0 = __fmul_rn(__fsub_rn(xv, __fmul_rn(args.coef.y, eps)), args.coef.x);
              args.x[idx] = __fadd_rn(__fmul_rn(args.coef.z, x0), __fmul_rn(args.coef.w, eps));
            }
          }
        }
      } else {
        const int sample = valid ? (m / args.rows_per_sample) : 0;
        const float* rb = nullptr;
        if (args.rowbias) {
          const long long r = args.rowbias_idx ? args.rowbias_idx[sample] : static_cast<long long>(sample);
          rb = args.rowbias + r * args.rb_ld;
        }
        if (!args.geglu) {
          constexpr int NCH = BN / 32;
          float gs[32];  // GroupNorm partials: [2g] = sum, [2g+1] = sum of squares of group g (10 columns) of this row
          if (args.gn_partial) {
#pragma unroll
            for (int i = 0; i < 32; ++i) gs[i] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + c * 32, v);
            tmem_ld_wait();
            if (c == NCH - 1) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            const int n = n0 + c * 32;
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (args.bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias + n + j));
                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
              }
            }
            if (rb) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(rb + n + j));
                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
              }
            }
            if (args.residual && valid) {
              const uint4* rp = reinterpret_cast<const uint4*>(args.residual + static_cast<size_t>(m) * args.res_ld + n);
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                const uint4 r4 = __ldg(rp + g4);
                const uint32_t ru[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 t = unpack_bf16x2(ru[j]);
                  f[g4 * 8 + 2 * j] += t.x;
                  f[g4 * 8 + 2 * j + 1] += t.y;
                }
              }
            }
            if (args.act == ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
            }
            if (args.gn_partial) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int g = (c * 32 + j) / 10;  // compile-time after unrolling (BN = 160 -> 16 groups of 10)
                const float x = valid ? f[j] : 0.f;
                gs[2 * g] += x;
                gs[2 * g + 1] = fmaf(x, x, gs[2 * g + 1]);
              }
            }
            if (valid) {
              if (args.out_f32) {
                float4* op = reinterpret_cast<float4*>(static_cast<float*>(args.out) + static_cast<size_t>(m) * args.out_ld + n);
#pragma unroll
                for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              } else {
                uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(m) * args.out_ld + n);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  op[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                     pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
              }
            }
          }
          if (args.gn_partial) {
            // rows of a warp belong to one sample (rows_per_sample % 32 == 0): reduce over the 32 rows, lane L keeps entry L
            const float tot = warp_transpose_reduce32(gs, lane);
            const int mw = m0 + q * 32;
            if (mw < args.M) {
              const int smp = mw / args.rows_per_sample;
              const int slot = (mw % args.rows_per_sample) >> 5;
              const int nslot = args.rows_per_sample >> 5;
              const int G = args.N / 10;
              const int g = n_tile * (BN / 10) + (lane >> 1);
              args.gn_partial[((static_cast<size_t>(smp) * G + g) * nslot + slot) * 2 + (lane & 1)] = tot;
            }
          }
        } else {
          constexpr int HALF = BN / 2;
#pragma unroll
          for (int c = 0; c < HALF / 16; ++c) {
            uint32_t va[16], vg[16];
            tmem_ld_32x32b_x16(t_row + c * 16, va);
            tmem_ld_32x32b_x16(t_row + HALF + c * 16, vg);
            tmem_ld_wait();
            if (c == HALF / 16 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            const int nb = n0 + c * 16;  // bias index of the value columns inside the permuted layout
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
              if (args.bias) {
                ba = __ldg(reinterpret_cast<const float4*>(args.bias + nb + j));
                bg = __ldg(reinterpret_cast<const float4*>(args.bias + nb + HALF + j));
              }
              f[j] = (__uint_as_float(va[j]) + ba.x) * gelu_fast_f(__uint_as_float(vg[j]) + bg.x);
              f[j + 1] = (__uint_as_float(va[j + 1]) + ba.y) * gelu_fast_f(__uint_as_float(vg[j + 1]) + bg.y);
              f[j + 2] = (__uint_as_float(va[j + 2]) + ba.z) * gelu_fast_f(__uint_as_float(vg[j + 2]) + bg.z);
              f[j + 3] = (__uint_as_float(va[j + 3]) + ba.w) * gelu_fast_f(__uint_as_float(vg[j + 3]) + bg.w);
            }
            if (valid) {
              const int n_out = n_tile * HALF + c * 16;
              uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(m) * args.out_ld + n_out);
              op[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
              op[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
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

template <int BN, int EPI>
static cudaError_t launch_impl(const GemmLaunch& L, cudaStream_t stream) {
  using C = Cfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const GemmArgs& a = L.args;
  if (a.N % BN != 0 || a.M <= 0) return cudaErrorInvalidValue;
  const int tiles = (a.N / BN) * ((a.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tc_kernel<BN, EPI><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(L.mapA[0], L.mapA[1], L.mapA[2], L.mapB, a);
  return cudaGetLastError();
}

cudaError_t gemm_tc_launch(const GemmLaunch& L, cudaStream_t stream) {
  const GemmArgs& a = L.args;
  if (a.epi == EPI_SAMPLER) {
    if (a.N != GEMM_BLOCK_N_OUT || !a.bias || !a.conv) return cudaErrorInvalidValue;
    return launch_impl<GEMM_BLOCK_N_OUT, EPI_SAMPLER>(L, stream);
  }
  if (a.gn_partial && (a.gn_cpg != 10 || a.rows_per_sample % 32 || a.geglu || a.N % 10)) return cudaErrorInvalidValue;
  return launch_impl<GEMM_BLOCK_N, EPI_STD>(L, stream);
}

}  // namespace wd
