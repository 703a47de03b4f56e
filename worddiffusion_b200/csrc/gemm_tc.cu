// tcgen05 implicit-GEMM kernel, see gemm_tc.cuh for the contract.
#include "gemm_tc.cuh"

#include <cstdio>
#include <mutex>

namespace wd {

constexpr int BLOCK_N = 160;  // 320 = 2 x 160; UMMA shape 128 x 160 x 16 (N % 16 == 0, <= 256)
constexpr int STAGES = 3;     // 3 x 36 KB: two CTAs stay resident per SM (smem 2 x 112 KB, TMEM 2 x 256 cols)
constexpr int A_STAGE_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
constexpr int B_STAGE_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int TMEM_COLS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

int gemm_tc_block_n() { return BLOCK_N; }

__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
               const GemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int m0 = m_tile * GEMM_BLOCK_M;
  const int n0 = n_tile * BLOCK_N;

  int total_k = 0;
#pragma unroll
  for (int s = 0; s < GEMM_MAX_SRC; ++s)
    if (s < args.num_src) total_k += args.taps[s] * args.chunks[s];

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA0);
    if (args.num_src > 1) tma_prefetch_desc(&mapA1);
    if (args.num_src > 2) tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int kb = 0;
      int img = 0, oh0 = 0;
      if (args.conv) {
        img = m0 / args.HWout;
        oh0 = (m0 % args.HWout) / args.Wout;
      }
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
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            uint8_t* sA = smem + stage * STAGE_BYTES;
            uint8_t* sB = sA + A_STAGE_BYTES;
            if (args.conv)
              tma_load_4d(sA, mapA, &full_bar[stage], ch * GEMM_BLOCK_K, dx, oh0 * st + dy, img);
            else
              tma_load_2d(sA, mapA, &full_bar[stage], ch * GEMM_BLOCK_K, m0);
            tma_load_2d(sB, &mapB, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
            ++kb;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (single thread) ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(GEMM_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < total_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t a_desc = make_smem_desc_sw128(a_addr);
        const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzled row: +2 in the (addr >> 4) field
          umma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // =========================== epilogue (4 warps, one TMEM lane quarter each) ===========================
    const int q = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool valid = m < args.M;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    const int sample = valid ? (m / args.rows_per_sample) : 0;
    const float* rb = nullptr;
    if (args.rowbias) {
      const long long r = args.rowbias_idx ? args.rowbias_idx[sample] : static_cast<long long>(sample);
      rb = args.rowbias + r * args.rb_ld;
    }

    if (!args.geglu) {
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + c * 16, v);
        tmem_ld_wait();
        if (valid) {
          const int n = n0 + c * 16;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (args.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += __ldg(args.bias + n + j);
          }
          if (rb) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += __ldg(rb + n + j);
          }
          if (args.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(args.residual + static_cast<size_t>(m) * args.res_ld + n);
            uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            const uint32_t ru[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float2 t = unpack_bf16x2(ru[j]);
              f[2 * j] += t.x;
              f[2 * j + 1] += t.y;
            }
          }
          if (args.act == ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = silu_f(f[j]);
          }
          if (args.out_f32) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(args.out) + static_cast<size_t>(m) * args.out_ld + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(m) * args.out_ld + n);
            op[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            op[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
          }
        }
      }
    } else {
      constexpr int HALF = BLOCK_N / 2;
#pragma unroll 1
      for (int c = 0; c < HALF / 16; ++c) {
        uint32_t va[16], vg[16];
        tmem_ld_32x32b_x16(t_row + c * 16, va);
        tmem_ld_32x32b_x16(t_row + HALF + c * 16, vg);
        tmem_ld_wait();
        if (valid) {
          const int nb = n0 + c * 16;  // bias index of the value columns inside the permuted layout
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(va[j]);
            float g = __uint_as_float(vg[j]);
            if (args.bias) {
              a += __ldg(args.bias + nb + j);
              g += __ldg(args.bias + nb + HALF + j);
            }
            f[j] = a * gelu_erf_f(g);
          }
          const int n_out = n_tile * HALF + c * 16;
          uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) + static_cast<size_t>(m) * args.out_ld + n_out);
          op[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          op[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
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

cudaError_t gemm_tc_launch(const GemmLaunch& L, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const GemmArgs& a = L.args;
  if (a.N % BLOCK_N != 0 || a.M <= 0) return cudaErrorInvalidValue;
  dim3 grid(a.N / BLOCK_N, (a.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  gemm_tc_kernel<<<grid, GEMM_THREADS, SMEM_BYTES, stream>>>(L.mapA[0], L.mapA[1], L.mapA[2], L.mapB, a);
  return cudaGetLastError();
}

}  // namespace wd
