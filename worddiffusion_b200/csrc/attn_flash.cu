// Flash-style fused attention for d_head = 80 (reference unetPhosc.py:176-196: softmax(q k^T * d^-0.5) v).
// Used for the 256x256 / 64x64 self-attention of UNetModelPhosc and for cross-attention over the 779-token
// char+PHOSC context.  The score matrix never touches HBM (the reference materialises [B*4, Sq, Skv] fp32).
//
// Round-1 implementation: one CTA = (64 queries, head, sample), 4 warps x 16 query rows, K/V streamed through
// shared memory in 64-key tiles, QK^T and PV on the warp-level bf16 tensor-core path (mma.sync.m16n8k16, fp32
// accumulate), online softmax in registers.  Attention is <= 10% of the step FLOPs; the tcgen05
// version of this kernel is a later-round item.
#include "ops.cuh"

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

cudaError_t attn_flash_launch(const AttnFlashArgs& a, int B, cudaStream_t s) {
  if (a.Sq < 1 || a.Skv < 1 || a.q_ld % 8 || a.kv_ld % 8 || a.out_ld % 2) return cudaErrorInvalidValue;
  dim3 grid((a.Sq + BQ - 1) / BQ, a.heads, B);
  if (a.Skv <= 16)
    attn_flash_kernel<16><<<grid, 128, 0, s>>>(a);
  else
    attn_flash_kernel<64><<<grid, 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace wd
