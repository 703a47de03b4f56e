// Flash-style fused attention for d_head = 80 (reference unetPhosc.py:176-196: softmax(q k^T * d^-0.5) v).
// Used for the 256x256 / 64x64 self-attention of UNetModelPhosc and for cross-attention over the 779-token
// char+PHOSC context.  The score matrix never touches HBM (the reference materialises [B*4, Sq, Skv] fp32).
//
// Round-1 implementation: one CTA = (64 queries, head, sample), 4 warps x 16 query rows, K/V streamed through
// shared memory in 64-key tiles, QK^T and PV on the warp-level bf16 tensor-core path (mma.sync.m16n8k16, fp32
// accumulate), online softmax in registers.  Attention is <= 10% of the step FLOPs; the tcgen05
// version of this kernel is a later-round item.
#include "ops.cuh"

#include <mutex>

namespace wd {

namespace {
constexpr int DH = 80;
constexpr int BQ = 64;       // queries per CTA
constexpr int KS = DH + 8;   // K smem row stride (bf16) -> conflict-free B-fragment loads

WD_DEVINL void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
}  // namespace

// BK = keys per tile: 64 for long key sequences, 16 for the 10-token character context (one tile, no rescaling pass)
template <int BK>
__global__ void __launch_bounds__(128) attn_flash_kernel(const AttnFlashArgs a) {
  constexpr int VS = BK + 8;  // Vt smem row stride (bf16)
  __shared__ __align__(16) __nv_bfloat16 sK[BK * KS];
  __shared__ __align__(16) __nv_bfloat16 sVt[DH * VS];

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;  // fragment row group / thread-in-quad

  const __nv_bfloat16* qb = a.q + static_cast<size_t>(b) * a.Sq * a.q_ld + h * DH;
  const __nv_bfloat16* kb = a.k + static_cast<size_t>(b) * a.Skv * a.kv_ld + h * DH;
  const __nv_bfloat16* vb = a.v + static_cast<size_t>(b) * a.Skv * a.kv_ld + h * DH;

  // ---- Q fragments (A operand, 16 rows x 80 dims per warp) ----
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const int r0c = min(r0, a.Sq - 1), r1c = min(r1, a.Sq - 1);
  uint32_t qf[DH / 16][4];
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) {
    const int d = ks * 16 + 2 * tq;
    qf[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(qb + static_cast<size_t>(r0c) * a.q_ld + d));
    qf[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(qb + static_cast<size_t>(r1c) * a.q_ld + d));
    qf[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(qb + static_cast<size_t>(r0c) * a.q_ld + d + 8));
    qf[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(qb + static_cast<size_t>(r1c) * a.q_ld + d + 8));
  }

  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sl2 = a.scale * 1.4426950408889634f;  // scale * log2(e)

  for (int k0 = 0; k0 < a.Skv; k0 += BK) {
    __syncthreads();  // previous tile fully consumed
    // ---- stage K [key][dim] and V^T [dim][key] (zero rows beyond Skv) ----
    for (int i = threadIdx.x; i < BK * (DH / 8); i += blockDim.x) {
      const int key = i / (DH / 8), vec = i % (DH / 8);
      uint4 kv4 = make_uint4(0, 0, 0, 0), vv4 = make_uint4(0, 0, 0, 0);
      if (k0 + key < a.Skv) {
        kv4 = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<size_t>(k0 + key) * a.kv_ld) + vec);
        vv4 = __ldg(reinterpret_cast<const uint4*>(vb + static_cast<size_t>(k0 + key) * a.kv_ld) + vec);
      }
      *reinterpret_cast<uint4*>(sK + key * KS + vec * 8) = kv4;
      const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv4);
#pragma unroll
      for (int j = 0; j < 8; ++j) sVt[(vec * 8 + j) * VS + key] = ve[j];
    }
    __syncthreads();

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[BK / 8][4];
#pragma unroll
    for (int nt = 0; nt < BK / 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
      const __nv_bfloat16* kr = sK + (nt * 8 + g) * KS + 2 * tq;
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
        mma_bf16_16816(s[nt], qf[ks], b0, b1);
      }
    }
    // ---- mask the tail, online softmax ----
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < BK / 8; ++nt) {
      const int key = k0 + nt * 8 + 2 * tq;
      if (key >= a.Skv) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (key + 1 >= a.Skv) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f((m0 - mx0) * sl2), c1 = exp2f((m1 - mx1) * sl2);  // first tile: exp2(-inf) = 0
    m0 = mx0;
    m1 = mx1;
    l0 *= c0;
    l1 *= c1;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      o[i][0] *= c0; o[i][1] *= c0;
      o[i][2] *= c1; o[i][3] *= c1;
    }
    uint32_t pf[BK / 16][4];  // P as A fragments (k = keys)
#pragma unroll
    for (int nt = 0; nt < BK / 8; ++nt) {
      const float p0 = exp2f((s[nt][0] - m0) * sl2), p1 = exp2f((s[nt][1] - m0) * sl2);
      const float p2 = exp2f((s[nt][2] - m1) * sl2), p3 = exp2f((s[nt][3] - m1) * sl2);
      l0 += p0 + p1;
      l1 += p2 + p3;
      const int kt = nt >> 1;
      if ((nt & 1) == 0) {
        pf[kt][0] = pack_bf16x2(p0, p1);
        pf[kt][1] = pack_bf16x2(p2, p3);
      } else {
        pf[kt][2] = pack_bf16x2(p0, p1);
        pf[kt][3] = pack_bf16x2(p2, p3);
      }
    }
    // ---- O += P V ----
#pragma unroll
    for (int dt = 0; dt < DH / 8; ++dt) {
      const __nv_bfloat16* vr = sVt + (dt * 8 + g) * VS + 2 * tq;
#pragma unroll
      for (int kt = 0; kt < BK / 16; ++kt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vr + kt * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vr + kt * 16 + 8);
        mma_bf16_16816(o[dt], pf[kt], b0, b1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = a.out + static_cast<size_t>(b) * a.Sq * a.out_ld + h * DH;
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) {
    const int d = dt * 8 + 2 * tq;
    if (r0 < a.Sq) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * a.out_ld + d) = pack_bf16x2(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < a.Sq) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * a.out_ld + d) = pack_bf16x2(o[dt][2] * i1, o[dt][3] * i1);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Cross-attention over a short context (Skv <= 16: the 10 character tokens of unet.UNetModel, unet.py:337-345).
// The flash kernel above is latency-bound there (one 64-query x 1-head CTA moves 10 KB with 4-byte accesses).  Here one
// CTA takes 64 query rows x ALL heads: Q rows and the whole K / V of the sample are staged with coalesced 16-byte loads
// into padded shared memory (row stride 656 B -> conflict-free ldmatrix / fragment reads), every warp owns 16 rows and
// loops over the heads (QK^T and PV on mma.sync.m16n8k16, softmax over the <= 16 keys in registers), writes its output
// tile over its own Q rows in shared memory and stores it with coalesced 16-byte writes.
// ----------------------------------------------------------------------------------------------------------------
namespace {
WD_DEVINL void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
WD_DEVINL void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
WD_DEVINL void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}
}  // namespace

__global__ void __launch_bounds__(128) attn_ctx_kernel(const AttnFlashArgs a) {
  extern __shared__ __align__(16) uint8_t actx_smem[];
  const int C = a.heads * DH;
  const int RS = C + 8;  // padded row stride (elements): 656 B for C = 320
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(actx_smem);  // [64][RS]  (Q, then the output tile)
  __nv_bfloat16* sK = sQ + 64 * RS;                                 // [16][RS]
  __nv_bfloat16* sV = sK + 16 * RS;                                 // [16][RS]
  const int b = blockIdx.y, q0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int nv = C >> 3;  // 16-byte vectors per row

  const __nv_bfloat16* qb = a.q + static_cast<size_t>(b) * a.Sq * a.q_ld;
  const __nv_bfloat16* kb = a.k + static_cast<size_t>(b) * a.Skv * a.kv_ld;
  const __nv_bfloat16* vb = a.v + static_cast<size_t>(b) * a.Skv * a.kv_ld;
  // asynchronous 16-byte copies (LDGSTS): all of a thread's ~25 loads are in flight at once; src-size 0 zero-fills
  for (int i = threadIdx.x; i < 64 * nv; i += 128) {
    const int r = i / nv, vec = i % nv;
    const bool ok = q0 + r < a.Sq;
    cp_async_16(sQ + r * RS + vec * 8, qb + static_cast<size_t>(ok ? q0 + r : 0) * a.q_ld + vec * 8, ok);
  }
  for (int i = threadIdx.x; i < 16 * nv; i += 128) {
    const int r = i / nv, vec = i % nv;
    const bool ok = r < a.Skv;
    cp_async_16(sK + r * RS + vec * 8, kb + static_cast<size_t>(ok ? r : 0) * a.kv_ld + vec * 8, ok);
    cp_async_16(sV + r * RS + vec * 8, vb + static_cast<size_t>(ok ? r : 0) * a.kv_ld + vec * 8, ok);
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();

  const float sl2 = a.scale * 1.4426950408889634f;  // scale * log2(e)
  const int r0 = warp * 16;
  for (int h = 0; h < a.heads; ++h) {
    // ---- Q fragments of this warp's 16 rows (A operand) ----
    uint32_t qf[DH / 16][4];
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks)
      ldmatrix_x4(qf[ks], sQ + (r0 + (lane & 15)) * RS + h * DH + ks * 16 + (lane >> 4) * 8);
    // ---- S = Q K^T (16 x 16 keys) ----
    float sc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[nt][j] = 0.f;
      const __nv_bfloat16* kr = sK + (nt * 8 + g) * RS + h * DH + 2 * tq;
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
        mma_bf16_16816(sc[nt], qf[ks], b0, b1);
      }
    }
    // ---- softmax over the keys (rows g and g + 8 of the warp tile; a row lives in one quad) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int key = nt * 8 + 2 * tq;
      if (key >= a.Skv) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
      if (key + 1 >= a.Skv) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pf[4];  // P as one A fragment (k = 16 keys)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const float p0 = exp2f((sc[nt][0] - mx0) * sl2), p1 = exp2f((sc[nt][1] - mx0) * sl2);
      const float p2 = exp2f((sc[nt][2] - mx1) * sl2), p3 = exp2f((sc[nt][3] - mx1) * sl2);
      l0 += p0 + p1;
      l1 += p2 + p3;
      pf[nt * 2] = pack_bf16x2(p0, p1);
      pf[nt * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    // ---- O = P V, written over this warp's Q rows of head h (their fragments are already in registers) ----
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < DH / 8; ++dt) {
      uint32_t b0, b1;
      ldmatrix_x2_trans(b0, b1, sV + (lane & 15) * RS + h * DH + dt * 8);
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o, pf, b0, b1);
      const int d = h * DH + dt * 8 + 2 * tq;
      *reinterpret_cast<uint32_t*>(sQ + (r0 + g) * RS + d) = pack_bf16x2(o[0] * i0, o[1] * i0);
      *reinterpret_cast<uint32_t*>(sQ + (r0 + g + 8) * RS + d) = pack_bf16x2(o[2] * i1, o[3] * i1);
    }
  }
  __syncwarp();
  // ---- coalesced store of the warp's 16 output rows ----
  __nv_bfloat16* ob = a.out + static_cast<size_t>(b) * a.Sq * a.out_ld;
  for (int i = lane; i < 16 * nv; i += 32) {
    const int r = r0 + i / nv, vec = i % nv;
    if (q0 + r < a.Sq)
      *(reinterpret_cast<uint4*>(ob + static_cast<size_t>(q0 + r) * a.out_ld) + vec) = *reinterpret_cast<const uint4*>(sQ + r * RS + vec * 8);
  }
}

cudaError_t attn_flash_launch(const AttnFlashArgs& a, int B, cudaStream_t s) {
  if (a.Sq < 1 || a.Skv < 1 || a.q_ld % 8 || a.kv_ld % 8 || a.out_ld % 2) return cudaErrorInvalidValue;
  {
    cudaError_t tc_err = cudaSuccess;
    if (attn_tc_try_launch(a, B, s, &tc_err)) return tc_err;  // tcgen05 kernel for the long key sequences
  }
  dim3 grid((a.Sq + BQ - 1) / BQ, a.heads, B);
  const int C = a.heads * DH;
  const size_t ctx_smem = static_cast<size_t>(64 + 32) * (C + 8) * 2;
  if (a.Skv <= 16 && ctx_smem <= 100 * 1024 && a.out_ld % 8 == 0) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
      attr_err = cudaFuncSetAttribute(attn_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    });
    if (attr_err != cudaSuccess) return attr_err;
    attn_ctx_kernel<<<dim3((a.Sq + 63) / 64, B), 128, ctx_smem, s>>>(a);
    return cudaGetLastError();
  }
  if (a.Skv <= 16)
    attn_flash_kernel<16><<<grid, 128, 0, s>>>(a);
  else
    attn_flash_kernel<64><<<grid, 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace wd
