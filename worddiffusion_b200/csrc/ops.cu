// HBM-/latency-bound kernels of the WordDiffusion hot path.  All activations are NHWC bf16
// ([B, H*W, C] == token-major), statistics and accumulators are fp32.
#include "ops.cuh"
#include "phosc_table.cuh"

#include <cstdlib>
#include <mutex>

namespace wd {

// =====================================================================================================
// GroupNorm (+SiLU) in two parts, no atomics (bit-reproducible):
//   statistics : per-tensor partial sums  partial[sample][32 groups][slots][{sum, sum of squares}]  (fp32) at the
//                granularity of C/32 channels.  They are written by the epilogue of the tcgen05 GEMM that produces
//                the tensor (gemm_tc.cu), or by groupnorm_stats_kernel for tensors produced elsewhere (conv_in).
//   apply      : groupnorm_apply_kernel folds the slots of its (<= 2) groups in a fixed order, forms the
//                per-channel scale/shift once and streams its rows: y = silu(x * sc + sh), 16-byte loads/stores,
//                4 independent rows in flight per thread.  A GroupNorm over the channel concatenation of two tensors
//                (decoder ResBlocks, unet.py:1750) reads each tensor's own partials and merges adjacent groups.
// Thread t owns vector column t % (Cs/8) (8 channels) and pixel rows t / (Cs/8), stepping by R.
// =====================================================================================================
constexpr int GN_R = 8;

__global__ void __launch_bounds__(1024) groupnorm_stats_kernel(const GroupNormStatsArgs a) {
  extern __shared__ float gn_part[];  // [2][R][C]
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int C = a.C, cpg = a.pcpg;
  const int nv = C >> 3;
  const int R = blockDim.x / nv;
  const int col = threadIdx.x % nv, rl = threadIdx.x / nv;
  const int ng = C / cpg;
  const int P = a.HW / a.pslots;
  const __nv_bfloat16* xb = a.x + (static_cast<size_t>(b) * a.HW + static_cast<size_t>(chunk) * P) * a.ld;

  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  for (int p0 = rl; p0 < P; p0 += 4 * R) {
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + i * R;
      v[i] = (p < P) ? __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * a.ld) + col) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_16x2(u[j], a.x_f16 != 0);
        s[2 * j] += f.x;
        s[2 * j + 1] += f.y;
        q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
        q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
      }
    }
  }
  float* ps = gn_part;
  float* pq = gn_part + R * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ps[rl * C + col * 8 + j] = s[j];
    pq[rl * C + col * 8 + j] = q[j];
  }
  __syncthreads();
  // one thread per (group, {sum, sumsq}): fixed summation order
  for (int i = threadIdx.x; i < 2 * ng; i += blockDim.x) {
    const int g = i >> 1, which = i & 1;
    const float* src = which ? pq : ps;
    float t = 0.f;
    for (int c = 0; c < cpg; ++c) {
      float tc = 0.f;
      for (int r = 0; r < R; ++r) tc += src[r * C + g * cpg + c];
      t += tc;
    }
    a.partial[((static_cast<size_t>(b) * ng + g) * a.pslots + chunk) * 2 + which] = t;
  }
}

int groupnorm_stats_slots(int HW) {
  int n = 1;
  while (n < 8 && HW % (2 * n) == 0 && HW / (2 * n) >= 4 * GN_R) n *= 2;
  return n;
}

cudaError_t groupnorm_stats_launch(const GroupNormStatsArgs& a, int B, cudaStream_t s) {
  const int nv = a.C / 8;
  if (a.C % 8 || a.C % a.pcpg || nv * GN_R > 1024 || !a.partial || a.pslots < 1 || a.HW % a.pslots)
    return cudaErrorInvalidValue;
  const size_t smem = static_cast<size_t>(2) * GN_R * a.C * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  groupnorm_stats_kernel<<<dim3(B, a.pslots), nv * GN_R, smem, s>>>(a);
  return cudaGetLastError();
}

constexpr int GN_APPLY_R = 16;  // pixel rows in flight per CTA pass (threads = Cs/8 * GN_APPLY_R)

// 640 threads (C/8 x 16 rows) and <= 51 registers: two CTAs per SM.  (With the default 1024-thread bound ptxas took 64
// registers -> one CTA per SM, 31 % occupancy, and the kernel ran latency-bound at 1.3-2.3 TB/s: profiles/r01k.)
__global__ void __launch_bounds__(640, 2) groupnorm_apply_kernel(const GroupNormArgs a) {
  __shared__ float s_mean[128], s_rstd[128];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, slab = blockIdx.y, chunk = blockIdx.z;
  const int Cs = a.Cs, cpg = a.cpg;
  const int nv = Cs >> 3;
  const int R = blockDim.x / nv;
  const int col = threadIdx.x % nv, rl = threadIdx.x / nv;
  const int P = a.HW / a.nchunk;
  const int ld = a.x_ld[slab];
  const int merge = cpg / a.pcpg;  // partial groups per GroupNorm group (1, or 2 for the concat GroupNorm)
  const int PG = Cs / a.pcpg;      // partial groups of this source tensor
  const int slots = a.pslots[slab];
  const int ng = Cs / cpg;         // GroupNorm groups inside this slab (<= 128, checked by the launcher)
  const bool xf16 = a.x_f16[slab] != 0;

  const __nv_bfloat16* xb = a.x[slab] + (static_cast<size_t>(b) * a.HW + static_cast<size_t>(chunk) * P) * ld;
  __nv_bfloat16* ob = a.out + (static_cast<size_t>(b) * a.HW + static_cast<size_t>(chunk) * P) * a.out_ld + slab * Cs;
  constexpr int INFL = 2;  // independent 16-byte loads in flight per thread (x 1280 resident threads per SM)

  // the first rows do not depend on the statistics: get them in flight before the reduction below
  uint4 v[INFL];
#pragma unroll
  for (int i = 0; i < INFL; ++i) {
    const int p = rl + i * R;
    if (p < P) v[i] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * ld) + col);
  }

  // one thread per group folds the partial sums in a fixed order (bit-reproducible), then everybody forms scale/shift
  if (static_cast<int>(threadIdx.x) < ng) {
    const float2* part = reinterpret_cast<const float2*>(a.partial[slab]) +
                         (static_cast<size_t>(b) * PG + static_cast<size_t>(threadIdx.x) * merge) * slots;
    float S = 0.f, Q = 0.f;
    const int n = merge * slots;
    for (int k = 0; k < n; ++k) {  // partial groups g*merge .. are contiguous: [pg][slot][2]
      const float2 t = __ldg(part + k);
      S += t.x;
      Q += t.y;
    }
    const float inv_n = 1.0f / static_cast<float>(cpg * a.HW);
    const float mean = S * inv_n;
    const float var = fmaxf(fmaf(-mean, mean, Q * inv_n), 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + a.eps);
  }
  __syncthreads();
  float sc[8], sh[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + slab * Cs) + col * 2);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + slab * Cs) + col * 2 + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + slab * Cs) + col * 2);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta + slab * Cs) + col * 2 + 1);
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (col * 8 + j) / cpg;
      sc[j] = s_rstd[g] * gm[j];
      sh[j] = be[j] - s_mean[g] * sc[j];
    }
  }
  for (int p0 = rl; p0 < P; p0 += INFL * R) {
    if (p0 != rl) {
#pragma unroll
      for (int i = 0; i < INFL; ++i) {
        const int p = p0 + i * R;
        if (p < P) v[i] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * ld) + col);
      }
    }
#pragma unroll
    for (int i = 0; i < INFL; ++i) {
      const int p = p0 + i * R;
      if (p >= P) break;
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_16x2(u[j], xf16);
        float y0 = fmaf(f.x, sc[2 * j], sh[2 * j]);
        float y1 = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
        if (a.silu) {
          y0 = silu_f(y0);
          y1 = silu_f(y1);
        }
        o[j] = pack_bf16x2(y0, y1);
      }
      reinterpret_cast<uint4*>(ob + static_cast<size_t>(p) * a.out_ld)[col] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Bulk-copy variant (the one the engine launches): persistent CTAs (two per SM), each walking a contiguous range of work
// items = GNB_P pixel rows of one (sample, source tensor).  The rows of an item are contiguous in global memory (row stride
// == channels), so ONE thread fetches each 20 KB item with cp.async.bulk into a 4-stage shared-memory ring, three items
// ahead; the rows are normalised in place in shared memory and leave through bulk stores (one per item, or one per row when
// the output interleaves two sources).  Loads of later items and stores of earlier ones overlap the arithmetic of the
// current one, ~120 KB of loads are in flight per SM, and no registers are spent on staging.  (The register-staged kernel
// above held 40 KB per SM and ran at 2.4 TB/s; a one-item-per-CTA bulk version moved its loads and stores in lockstep
// waves and reached 2.7 TB/s: profiles/.)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GNB_P = 32;      // pixel rows per work item
constexpr int GNB_R = 8;       // rows per pass (threads = Cs/8 * GNB_R)
constexpr int GNB_STAGES = 4;

// one 16-byte vector (8 channels): y = x * sc + sh (packed fp32 FMAs); SiLU as h + h tanh(h), h = y / 2, with ONE packed
// tanh.approx.f16x2 per two elements (sc / sh arrive pre-halved when SILU): 3.5 instructions and half a MUFU operation per
// element.  History: ~20 instructions per element (run-time format / activation switches, scalar arithmetic) ran issue-bound
// at half the HBM rate; the ex2 + rcp form (5.5 instructions, two MUFU operations per element) kept the XU pipe 55 % busy
// and the SiLU launches at 3.0-3.4 TB/s against 4.7 TB/s without SiLU (profiles/R2d_ncu_gn.txt).  tanh's absolute error
// (2^-11, fp16 result) leaves |error| <= |h| 2^-11 on the output, below the bf16 rounding of the stored value for y > -4.
template <bool SILU, bool XF16>
WD_DEVINL uint4 gn_vec8(const uint4 v, const float2 (&sc2)[4], const float2 (&sh2)[4]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = XF16 ? unpack_f16x2(u[j]) : unpack_bf16x2(u[j]);
    float2 y = __ffma2_rn(f, sc2[j], sh2[j]);  // SILU: this is h = y / 2
    if constexpr (SILU) {
      uint32_t h16 = pack_f16x2(y.x, y.y), t16;
      asm("tanh.approx.f16x2 %0, %1;" : "=r"(t16) : "r"(h16));
      y = __ffma2_rn(y, unpack_f16x2(t16), y);
    }
    o[j] = pack_bf16x2(y.x, y.y);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

template <bool SILU>
__global__ void __launch_bounds__(512, 2) groupnorm_apply_bulk_kernel(const GroupNormArgs a, int nslab, int n_items, int per_cta) {
  extern __shared__ __align__(128) uint8_t gnb_smem[];
  __shared__ float s_mean[128], s_rstd[128];
  __shared__ __align__(8) uint64_t bars[GNB_STAGES];
  const int Cs = a.Cs, cpg = a.cpg;
  const int nv = Cs >> 3;
  const int col = threadIdx.x % nv, rl = threadIdx.x / nv;
  const int R = blockDim.x / nv;
  const int merge = cpg / a.pcpg;
  const int PG = Cs / a.pcpg;
  const int ng = Cs / cpg;
  const uint32_t row_bytes = static_cast<uint32_t>(Cs) * 2;
  const uint32_t item_bytes = GNB_P * row_bytes;
  // WD_GN_REVERSE experiment (a.reverse): CTA 0 takes the LAST items -- the rows the producing GEMM wrote most recently and the
  // likeliest to still sit in L2
  const int cta = a.reverse ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
  const int first = cta * per_cta;
  const int last = min(first + per_cta, n_items);

  if (threadIdx.x == 0) {
    for (int i = 0; i < GNB_STAGES; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  pdl_trigger();
  pdl_wait();  // x and its statistics come from the previous kernel; our output may still be read by an earlier one
  __syncthreads();
  auto issue_load = [&](int item, int stage) {  // item -> (sample, slab, chunk); its rows are contiguous
    const int chunk = item % a.nchunk, bs = item / a.nchunk;
    const int slab = bs % nslab, b = bs / nslab;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(a.x[slab] + (static_cast<size_t>(b) * a.HW + static_cast<size_t>(chunk) * GNB_P) * Cs);
    mbar_arrive_expect_tx(&bars[stage], item_bytes);
    bulk_load_1d(gnb_smem + stage * item_bytes, src, item_bytes, &bars[stage]);
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < GNB_STAGES - 1 && first + i < last; ++i) issue_load(first + i, i);
  }
  float2 sc2[4], sh2[4];
  int cur_bs = -1;
  for (int item = first, k = 0; item < last; ++item, ++k) {
    const int chunk = item % a.nchunk, bs = item / a.nchunk;
    const int slab = bs % nslab, b = bs / nslab;
    if (bs != cur_bs) {
      cur_bs = bs;
      __syncthreads();  // everybody has formed scale/shift from the previous (sample, slab) statistics
      // one thread per group folds the partial sums in a fixed order (bit-reproducible), then everybody forms scale/shift
      if (static_cast<int>(threadIdx.x) < ng) {
        const int slots = a.pslots[slab];
        const float2* part = reinterpret_cast<const float2*>(a.partial[slab]) +
                             (static_cast<size_t>(b) * PG + static_cast<size_t>(threadIdx.x) * merge) * slots;
        float S = 0.f, Q = 0.f;
        const int n = merge * slots;
        for (int i = 0; i < n; ++i) {
          const float2 t = __ldg(part + i);
          S += t.x;
          Q += t.y;
        }
        const float inv_n = 1.0f / static_cast<float>(cpg * a.HW);
        const float mean = S * inv_n;
        const float var = fmaxf(fmaf(-mean, mean, Q * inv_n), 0.f);
        s_mean[threadIdx.x] = mean;
        s_rstd[threadIdx.x] = rsqrtf(var + a.eps);
      }
      __syncthreads();
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + slab * Cs) + col * 2);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + slab * Cs) + col * 2 + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + slab * Cs) + col * 2);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta + slab * Cs) + col * 2 + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g0i = (col * 8 + 2 * j) / cpg, g1i = (col * 8 + 2 * j + 1) / cpg;
        sc2[j] = make_float2(s_rstd[g0i] * gm[2 * j], s_rstd[g1i] * gm[2 * j + 1]);
        // (explicit fma: gemm_pair.cu's producer-side GroupNorm forms the same scale / shift and must give the same bits)
        sh2[j] = make_float2(fmaf(-s_mean[g0i], sc2[j].x, be[2 * j]), fmaf(-s_mean[g1i], sc2[j].y, be[2 * j + 1]));
      }
      if constexpr (SILU) {  // gn_vec8 wants h = y / 2
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sc2[j] = make_float2(0.5f * sc2[j].x, 0.5f * sc2[j].y);
          sh2[j] = make_float2(0.5f * sh2[j].x, 0.5f * sh2[j].y);
        }
      }
    }
    const int stage = k % GNB_STAGES;
    const bool xf16 = a.x_f16[slab] != 0;
    uint8_t* const sbase = gnb_smem + stage * item_bytes;
    mbar_wait(&bars[stage], (k / GNB_STAGES) & 1);  // the item has landed
    if (xf16) {
      for (int p = rl; p < GNB_P; p += R) {
        uint4* sp = reinterpret_cast<uint4*>(sbase + p * row_bytes) + col;
        *sp = gn_vec8<SILU, true>(*sp, sc2, sh2);
      }
    } else {
      for (int p = rl; p < GNB_P; p += R) {
        uint4* sp = reinterpret_cast<uint4*>(sbase + p * row_bytes) + col;
        *sp = gn_vec8<SILU, false>(*sp, sc2, sh2);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      __nv_bfloat16* ob = a.out + (static_cast<size_t>(b) * a.HW + static_cast<size_t>(chunk) * GNB_P) * a.out_ld + slab * Cs;
      if (a.out_ld == Cs) {
        bulk_store_1d(ob, sbase, item_bytes);
      } else {
        for (int r = 0; r < GNB_P; ++r) bulk_store_1d(ob + static_cast<size_t>(r) * a.out_ld, sbase + r * row_bytes, row_bytes);
      }
      bulk_commit_group();
      const int nxt = item + GNB_STAGES - 1;
      if (nxt < last) {
        bulk_wait_group_read<1>();  // the ring slot of item k+3 held item k-1: its store (the group before this one) has read it
        issue_load(nxt, (k + GNB_STAGES - 1) % GNB_STAGES);
      }
    }
  }
  if (threadIdx.x == 0) bulk_wait_group_read<0>();  // shared memory must outlive the last stores' reads
}

static bool groupnorm_bulk_enabled() {  // env WD_GN_BULK (default on)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WD_GN_BULK");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

int groupnorm_apply_chunks(int HW) {
  // 8 pixel rows per thread (four passes of 2 independent 16-byte loads) when the image is large enough: finer chunks
  // (more CTAs, each repeating the statistics preamble) measured slower (profiles/)
  const int per = 8 * GN_APPLY_R;
  return (HW >= 2 * per && HW % per == 0) ? HW / per : 1;
}

cudaError_t groupnorm_launch(const GroupNormArgs& a, int B, int nslab, cudaStream_t s) {
  const int nv = a.Cs / 8;
  if (a.Cs % 8 || a.Cs % a.cpg || a.cpg % a.pcpg || a.Cs / a.cpg > 128 || a.nchunk < 1 || a.HW % a.nchunk)
    return cudaErrorInvalidValue;
  // bulk-copy kernel: contiguous source rows, whole 32-row items, a ring of <= 100 KB (two CTAs per SM)
  const size_t ring = static_cast<size_t>(GNB_STAGES) * GNB_P * a.Cs * 2;
  bool bulk = groupnorm_bulk_enabled() && a.HW % GNB_P == 0 && nv * GNB_R <= 512 && nv * GNB_R >= a.Cs / a.cpg && ring <= 100 * 1024 &&
              a.out_ld % 8 == 0;
  for (int i = 0; i < nslab; ++i) bulk = bulk && a.x_ld[i] == a.Cs && a.partial[i] && a.pslots[i] >= 1;
  if (bulk) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static int sms = 148;
    std::call_once(once, [] {
      attr_err = cudaFuncSetAttribute(groupnorm_apply_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute(groupnorm_apply_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      int dev = 0;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
    });
    if (attr_err != cudaSuccess) return attr_err;
    GroupNormArgs a2 = a;
    a2.nchunk = a.HW / GNB_P;
    static const int rev = [] { const char* e = getenv("WD_GN_REVERSE"); return e ? atoi(e) : 0; }();
    a2.reverse = rev;
    const int n_items = B * nslab * a2.nchunk;
    const int per = (n_items + 2 * sms - 1) / (2 * sms);  // contiguous items per CTA: consecutive chunks share their statistics
    const int grid = (n_items + per - 1) / per;
    if (a.silu) return launch_pdl(groupnorm_apply_bulk_kernel<true>, dim3(grid), dim3(nv * GNB_R), ring, s, a2, nslab, n_items, per);
    return launch_pdl(groupnorm_apply_bulk_kernel<false>, dim3(grid), dim3(nv * GNB_R), ring, s, a2, nslab, n_items, per);
  }
  int R = GN_APPLY_R;
  while (R > 1 && nv * R > 640) R >>= 1;
  if (nv * R > 640 || nv * R < a.Cs / a.cpg) return cudaErrorInvalidValue;
  for (int i = 0; i < nslab; ++i)
    if (!a.partial[i] || a.pslots[i] < 1) return cudaErrorInvalidValue;
  return launch_pdl(groupnorm_apply_kernel, dim3(B, nslab, a.nchunk), dim3(nv * R), 0, s, a);
}

// Forward noising of the training step (train.py:190-194): x_t = sqrt(ah[t]) x + sqrt(1 - ah[t]) eps, eps ~ N(0, I)
// from the caller (eps_in) or from the sampler's Philox4x32-10 stream keyed by (seed, global element, stream id) -- one pass that
// writes both x_t and eps.  alpha_hat: the schedule's cumulative products [T] on the device; t: int64 [n].
__global__ void noise_images_kernel(const float* __restrict__ x, const long long* __restrict__ t, const float* __restrict__ alpha_hat,
                                    int T, const float* __restrict__ eps_in, unsigned long long seed, unsigned long long elem_offset,
                                    uint32_t stream_id, float* __restrict__ x_t, float* __restrict__ eps_out, size_t n, int per,
                                    int* __restrict__ bad) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long ti = t[i / per];
  if (ti < 0 || ti >= T) {  // the reference's alpha_hat[t] raises IndexError
    if (bad) atomicExch(bad, 1);
    return;
  }
  const float ah = alpha_hat[ti];
  const float e = eps_in ? eps_in[i] : philox_normal(seed, elem_offset + i, stream_id);
  eps_out[i] = e;
  x_t[i] = __fadd_rn(__fmul_rn(sqrtf(ah), x[i]), __fmul_rn(sqrtf(1.0f - ah), e));
}
cudaError_t noise_images_launch(const float* x, const long long* t, const float* alpha_hat, int T, const float* eps_in,
                                unsigned long long seed, unsigned long long elem_offset, uint32_t stream_id, float* x_t, float* eps_out,
                                size_t n, int per, int* bad, cudaStream_t s) {
  if (!n) return cudaSuccess;
  noise_images_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x, t, alpha_hat, T, eps_in, seed, elem_offset, stream_id, x_t,
                                                                             eps_out, n, per, bad);
  return cudaGetLastError();
}

// nn.MSELoss (train.py:287) and its gradient in one pass: d[i] = 2 (pred[i] - target[i]) / n, loss = mean((pred - target)^2).
// Deterministic: per-CTA partial sums in fixed order, folded by the last CTA to finish (ticket counter, no float atomics).
constexpr int MSE_THREADS = 256, MSE_PER_THREAD = 8;
__global__ void __launch_bounds__(MSE_THREADS) mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                               float* __restrict__ d, float* __restrict__ partial,
                                                               unsigned int* __restrict__ ticket, float* __restrict__ loss, size_t n) {
  __shared__ float red[MSE_THREADS / 32];
  __shared__ bool last;
  const float inv = 2.0f / static_cast<float>(n);
  float acc = 0.f;
  const size_t base = static_cast<size_t>(blockIdx.x) * MSE_THREADS * MSE_PER_THREAD;
#pragma unroll
  for (int k = 0; k < MSE_PER_THREAD; ++k) {
    const size_t i = base + static_cast<size_t>(k) * MSE_THREADS + threadIdx.x;
    if (i < n) {
      const float df = pred[i] - target[i];
      d[i] = df * inv;
      acc = fmaf(df, df, acc);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
    for (int w = 0; w < MSE_THREADS / 32; ++w) tsum += red[w];
    partial[blockIdx.x] = tsum;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) tot += static_cast<double>(partial[b]);
    *loss = static_cast<float>(tot / static_cast<double>(n));
    *ticket = 0u;  // ready for the next call
  }
}
size_t mse_grad_workspace_bytes(size_t n) {
  const size_t blocks = (n + MSE_THREADS * MSE_PER_THREAD - 1) / (MSE_THREADS * MSE_PER_THREAD);
  return 16 + blocks * sizeof(float);
}
cudaError_t mse_grad_launch(const float* pred, const float* target, float* d, float* loss, void* workspace, size_t n, cudaStream_t s) {
  if (!n) return cudaErrorInvalidValue;
  const unsigned blocks = static_cast<unsigned>((n + MSE_THREADS * MSE_PER_THREAD - 1) / (MSE_THREADS * MSE_PER_THREAD));
  unsigned int* ticket = static_cast<unsigned int*>(workspace);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + 16);
  mse_grad_kernel<<<blocks, MSE_THREADS, 0, s>>>(pred, target, d, partial, ticket, loss, n);
  return cudaGetLastError();
}

// torch.lerp(start, end, weight) with torch's formula (weight < 0.5 ? start + weight (end - start) : end - (end - start)(1 - weight)):
// the classifier-free-guidance mix of train.py:226-228
__global__ void lerp_kernel(const float* __restrict__ a, const float* __restrict__ b, float w, float* __restrict__ out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = __fsub_rn(b[i], a[i]);
  out[i] = (fabsf(w) < 0.5f) ? __fadd_rn(a[i], __fmul_rn(w, d)) : __fsub_rn(b[i], __fmul_rn(d, 1.0f - w));
}
cudaError_t lerp_launch(const float* a, const float* b, float w, float* out, size_t n, cudaStream_t s) {
  lerp_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(a, b, w, out, n);
  return cudaGetLastError();
}

// label_emb.weight[row] = (1 - mix) * w[s1] + mix * w[s2] with the reference's roundings (unet.py:1568)
__global__ void label_mix_kernel(float* __restrict__ w, int D, int row, int s1, int s2, float mix) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  const float a = __fmul_rn(1.0f - mix, w[static_cast<size_t>(s1) * D + c]);
  const float b = __fmul_rn(mix, w[static_cast<size_t>(s2) * D + c]);
  w[static_cast<size_t>(row) * D + c] = __fadd_rn(a, b);
}
cudaError_t label_mix_launch(float* table, int D, int row, int s1, int s2, float mix, cudaStream_t s) {
  label_mix_kernel<<<(D + 255) / 256, 256, 0, s>>>(table, D, row, s1, s2, mix);
  return cudaGetLastError();
}

// =====================================================================================================
// LayerNorm: one warp per token, fp32 two-pass in registers, bf16 in/out.
// =====================================================================================================
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                        __nv_bfloat16* __restrict__ out,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int M, int C, float eps,
                                                        int x_f16) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nv = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(warp) * C);
  float f[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nv) {
      const uint4 v = __ldg(xr + vi);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_16x2(u[j], x_f16 != 0);
        f[i][2 * j] = t.x;
        f[i][2 * j + 1] = t.y;
        sum += t.x + t.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(C);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + 32 * i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        var += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / static_cast<float>(C) + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<size_t>(warp) * C);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nv) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + vi * 2);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + vi * 2 + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + vi * 2);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + vi * 2 + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (f[i][j] - mean) * rstd * g[j] + bb[j];
      orow[vi] = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                            pack_bf16x2(y[6], y[7]));
    }
  }
}

cudaError_t layernorm_launch(const __nv_bfloat16* x, __nv_bfloat16* out, const float* gamma, const float* beta, int M,
                             int C, float eps, int x_f16, cudaStream_t s) {
  if (C % 8 || C > 8 * 32 * 4) return cudaErrorInvalidValue;
  const int warps_per_block = 8;
  const int blocks = (M + warps_per_block - 1) / warps_per_block;
  if (C <= 8 * 32 * 2)
    layernorm_kernel<2><<<blocks, 256, 0, s>>>(x, out, gamma, beta, M, C, eps, x_f16);
  else
    layernorm_kernel<4><<<blocks, 256, 0, s>>>(x, out, gamma, beta, M, C, eps, x_f16);
  return cudaGetLastError();
}

// =====================================================================================================
// Attention over a short context (L <= 16): one thread per (query token, head); K/V of the sample staged
// in shared memory as fp32.  softmax(q k^T * scale) v, exactly the reference op order (unet.py:195-270).
// =====================================================================================================
template <int DH>
__global__ void __launch_bounds__(128) attn_small_kernel(const AttnSmallArgs a) {
  extern __shared__ float as_smem[];
  const int b = blockIdx.y;
  const int C = a.heads * DH;
  float* sk = as_smem;           // [L][C]
  float* sv = as_smem + a.L * C;  // [L][C]
  const __nv_bfloat16* kb = a.k + static_cast<size_t>(b) * a.L * a.kv_ld;
  const __nv_bfloat16* vb = a.v + static_cast<size_t>(b) * a.L * a.kv_ld;
  for (int i = threadIdx.x; i < a.L * C; i += blockDim.x) {
    const int l = i / C, c = i % C;
    sk[i] = __bfloat162float(kb[static_cast<size_t>(l) * a.kv_ld + c]);
    sv[i] = __bfloat162float(vb[static_cast<size_t>(l) * a.kv_ld + c]);
  }
  __syncthreads();
  const int tokens_per_block = blockDim.x / a.heads;
  const int tok = blockIdx.x * tokens_per_block + threadIdx.x / a.heads;
  const int h = threadIdx.x % a.heads;
  if (tok >= a.Sq) return;

  float q[DH];
  const uint4* qp = reinterpret_cast<const uint4*>(a.q + (static_cast<size_t>(b) * a.Sq + tok) * a.q_ld + h * DH);
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    const uint4 v = __ldg(qp + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_bf16x2(u[j]);
      q[i * 8 + 2 * j] = t.x;
      q[i * 8 + 2 * j + 1] = t.y;
    }
  }
  float sc[16];
  float mx = -INFINITY;
#pragma unroll
  for (int l = 0; l < 16; ++l) {
    if (l < a.L) {
      const float* kr = sk + l * C + h * DH;
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < DH; ++i) d += q[i] * kr[i];
      sc[l] = d * a.scale;
      mx = fmaxf(mx, sc[l]);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int l = 0; l < 16; ++l) {
    if (l < a.L) {
      sc[l] = __expf(sc[l] - mx);
      den += sc[l];
    }
  }
  const float inv = 1.0f / den;
  if (a.probs) {
    float* pp = a.probs + ((static_cast<size_t>(b) * a.heads + h) * a.Sq + tok) * a.L;
#pragma unroll
    for (int l = 0; l < 16; ++l)
      if (l < a.L) pp[l] = sc[l] * inv;
  }
  float o[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) o[i] = 0.f;
#pragma unroll
  for (int l = 0; l < 16; ++l) {
    if (l < a.L) {
      const float p = sc[l] * inv;
      const float* vr = sv + l * C + h * DH;
#pragma unroll
      for (int i = 0; i < DH; ++i) o[i] += p * vr[i];
    }
  }
  uint4* op = reinterpret_cast<uint4*>(a.out + (static_cast<size_t>(b) * a.Sq + tok) * a.out_ld + h * DH);
#pragma unroll
  for (int i = 0; i < DH / 8; ++i)
    op[i] = make_uint4(pack_bf16x2(o[i * 8], o[i * 8 + 1]), pack_bf16x2(o[i * 8 + 2], o[i * 8 + 3]),
                       pack_bf16x2(o[i * 8 + 4], o[i * 8 + 5]), pack_bf16x2(o[i * 8 + 6], o[i * 8 + 7]));
}

cudaError_t attn_small_launch(const AttnSmallArgs& a, int B, cudaStream_t s) {
  if (a.L > 16 || a.L < 1 || 128 % a.heads) return cudaErrorInvalidValue;
  const int C = a.heads * 80;
  const size_t smem = static_cast<size_t>(2) * a.L * C * sizeof(float);
  const int tpb = 128 / a.heads;
  dim3 grid((a.Sq + tpb - 1) / tpb, B);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_small_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  if (smem > 100 * 1024) return cudaErrorInvalidValue;
  attn_small_kernel<80><<<grid, 128, smem, s>>>(a);
  return cudaGetLastError();
}

// =====================================================================================================
// Sinusoidal timestep embedding: out[b] = [cos(t f_0..f_{h-1}) | sin(t f_0..f_{h-1})], f_i = exp(-ln(1e4) i / h)
// =====================================================================================================
__global__ void timestep_embed_kernel(const long long* __restrict__ t_dev, long long t_scalar,
                                      __nv_bfloat16* __restrict__ out, int B, int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  pdl_trigger();
  pdl_wait();
  if (idx >= B * half) return;
  const int b = idx / half, i = idx % half;
  const float t = static_cast<float>(t_dev ? t_dev[b] : (t_scalar < 0 ? static_cast<long long>(b) : t_scalar));  // t_scalar < 0: row b embeds t = b
  const float freq = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
  const float arg = t * freq;
  out[static_cast<size_t>(b) * dim + i] = __float2bfloat16(cosf(arg));
  out[static_cast<size_t>(b) * dim + half + i] = __float2bfloat16(sinf(arg));
}

cudaError_t timestep_embed_launch(const long long* t_dev, long long t_scalar, __nv_bfloat16* out, int B, int dim,
                                  cudaStream_t s) {
  if (dim % 2) return cudaErrorInvalidValue;
  const int n = B * (dim / 2);
  return launch_pdl(timestep_embed_kernel, dim3((n + 255) / 256), dim3(256), 0, s, t_dev, t_scalar, out, B, dim);
}

// =====================================================================================================
// conv_in (unet.py:1251): 3x3 pad 1, Cin = 4 on the tensor cores with fp32-class accuracy (the first layer's rounding
// error propagates through the whole net: bf16 weights here alone cost 1.4e-3 of the 1e-2 budget).  This kernel builds
// the im2col operand A[m, 128] bf16 from the fp32 NCHW latent, with x = x_hi + x_lo and w = w_hi + w_lo (bf16 pairs):
//   columns  0..35  x_hi (j = c*9 + ky*3 + kx)   against weight columns w_hi
//   columns 36..71  x_lo                          against w_hi
//   columns 72..107 x_hi                          against w_lo
// (the x_lo*w_lo term is ~2^-18 and dropped); columns 108..127 are zero.  One thread = 8 columns (16 B).
// =====================================================================================================
// Four threads per pixel: each loads the pixel's 3 x 3 x 4 patch once (36 predicated loads, L1 hits across neighbours) and
// writes 64 contiguous bytes of the 256-byte row.  (The first version recomputed tap coordinates and issued dependent loads per
// output element: 29 us for a 17 MB tensor.)
__global__ void __launch_bounds__(256) conv_in_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                             int B, int H, int W) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * H * W * 4;
  pdl_trigger();
  pdl_wait();
  if (idx >= total) return;
  const int sub = idx & 3;
  const size_t m = idx >> 2;
  const int px = m % W;
  const int py = (m / W) % H;
  const int b = m / (static_cast<size_t>(W) * H);
  float v[36];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = py + ky - 1, xx = px + kx - 1;
        float xv = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) xv = __ldg(x + ((static_cast<size_t>(b) * 4 + c) * H + yy) * W + xx);
        v[c * 9 + ky * 3 + kx] = xv;
      }
  // column k of the row: k < 36 hi(v[k]); 36 <= k < 72 lo(v[k-36]); 72 <= k < 108 hi(v[k-72]); else 0
  auto colval = [&](int k) -> float {
    if (k >= 108) return 0.f;
    const float xv = v[k % 36];
    const float hi = __bfloat162float(__float2bfloat16(xv));
    return (k >= 36 && k < 72) ? xv - hi : hi;
  };
  uint4* dst = reinterpret_cast<uint4*>(out + m * 128);
#pragma unroll
  for (int s4 = 0; s4 < 4; ++s4) {
    if (s4 != sub) continue;  // (compile-time column indices inside each branch: v[] stays in registers)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int k0 = (s4 * 4 + ch) * 8;
      dst[s4 * 4 + ch] = make_uint4(pack_bf16x2(colval(k0), colval(k0 + 1)), pack_bf16x2(colval(k0 + 2), colval(k0 + 3)),
                                    pack_bf16x2(colval(k0 + 4), colval(k0 + 5)), pack_bf16x2(colval(k0 + 6), colval(k0 + 7)));
    }
  }
}

cudaError_t conv_in_im2col_launch(const float* x, __nv_bfloat16* out, int B, int H, int W, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * H * W * 4;
  return launch_pdl(conv_in_im2col_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, x, out, B, H, W);
}

// =====================================================================================================
// PHOSC tokenizer (reference ResPhoSCNetZSL/modules/utils/phos_generator.py:59-78, phoc_generator.py:17-90, 'eng'): the step
// in front of UNetModelPhosc -- words -> [B, 769] integer labels -- as integer work on the device.
//   out[0,165)   PHOS: 11 shape counts of the whole word and of the segments of its 2..5-way splits (15 segments)
//   out[165,669) PHOC: 36 presence bits [0-9a-z] of the lower-cased word per segment of the 2..5-way splits (14 segments)
//   out[669,769) the 2 x 50 bigram bits, which the reference never sets (it looks single characters up in a bigram list)
// words: [B, max_len] bytes, zero padded (the caller has removed ' ' and '_', trainGWModifyCondition.py:394).  One CTA per word.
// =====================================================================================================
__global__ void __launch_bounds__(256) phosc_tokenize_kernel(const unsigned char* __restrict__ words, int max_len,
                                                             int* __restrict__ out, int* __restrict__ bad) {
  const unsigned char* w = words + static_cast<size_t>(blockIdx.x) * max_len;
  int L = 0;
  while (L < max_len && w[L] != 0) ++L;
  for (int e = threadIdx.x; e < 769; e += blockDim.x) {
    int val = 0;
    if (e < 669) {
      const bool is_phos = e < 165;
      const int ee = is_phos ? e : e - 165;
      const int width = is_phos ? PHOS_SHAPES : 36;
      int seg = ee / width;
      const int colx = ee % width;
      if (!is_phos) ++seg;  // PHOC has no level-1 segment
      // segment index -> (split, mul): seg 0 = whole word; then splits 2,3,4,5 with 2,3,4,5 segments each
      int a = 0, b = L;
      if (seg > 0) {
        int split = 2, base = 1;
        while (seg >= base + split) { base += split; ++split; }
        const int mul = seg - base;
        const int parts = L / split;
        a = mul * parts;
        b = (mul == split - 1) ? L : a + parts;
      }
      for (int i = a; i < b; ++i) {
        const int ch = w[i];
        if (is_phos) {
          int li = -1;
          if (ch >= 'a' && ch <= 'z') li = ch - 'a';
          else if (ch >= 'A' && ch <= 'Z') li = 26 + ch - 'A';
          if (li < 0) { atomicOr(bad, 1); continue; }  // the reference raises KeyError for a character outside a-zA-Z
          val += PHOS_TABLE[li][colx];
        } else {
          const int lc = (ch >= 'A' && ch <= 'Z') ? ch - 'A' + 'a' : ch;
          if (lc >= '0' && lc <= '9') { if (lc - '0' == colx) val = 1; }
          else if (lc >= 'a' && lc <= 'z') { if (10 + lc - 'a' == colx) val = 1; }
        }
      }
    }
    out[static_cast<size_t>(blockIdx.x) * 769 + e] = val;
  }
}
cudaError_t phosc_tokenize_launch(const unsigned char* words, int B, int max_len, int* out, int* bad_flag, cudaStream_t s) {
  if (B < 1 || max_len < 1) return cudaErrorInvalidValue;
  phosc_tokenize_kernel<<<B, 256, 0, s>>>(words, max_len, out, bad_flag);
  return cudaGetLastError();
}

// =====================================================================================================
// emb_act[b] = SiLU(time_embed_table[t] + label_emb[y[b]])  (unet.py:1575-1581 + the SiLU every emb_layers starts with, :609-610).
// The sampling loop evaluates all latents at ONE timestep: time_embed(t) comes from a per-trajectory table over all timesteps
// (two small GEMMs per trajectory instead of two latency-bound GEMMs + the sinusoid kernel per step).
// =====================================================================================================
__global__ void emb_from_table_kernel(const float* __restrict__ table, long long t, const StepParams* __restrict__ sp,
                                      const float* __restrict__ label_emb, const long long* __restrict__ y,
                                      __nv_bfloat16* __restrict__ out, int B, int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (idx >= B * dim) return;
  if (sp) t = sp->t;  // graph replay: the timestep lives in device memory
  const int b = idx / dim, d = idx % dim;
  float v = __ldg(table + static_cast<size_t>(t) * dim + d);
  if (label_emb) v += __ldg(label_emb + static_cast<size_t>(y[b]) * dim + d);
  out[idx] = __float2bfloat16(silu_f(v));
}
cudaError_t emb_from_table_launch(const float* table, long long t, const StepParams* sp, const float* label_emb, const long long* y,
                                  __nv_bfloat16* out, int B, int dim, cudaStream_t s) {
  if (label_emb && !y) return cudaErrorInvalidValue;
  return launch_pdl(emb_from_table_kernel, dim3((B * dim + 255) / 256), dim3(256), 0, s, table, t, sp, label_emb, y, out, B, dim);
}
// refresh the device-resident step parameters (by-value kernel argument: no host staging buffer to keep alive)
__global__ void set_step_params_kernel(StepParams* dst, const StepParams v) { *dst = v; }
cudaError_t set_step_params_launch(StepParams* dst, const StepParams& v, cudaStream_t s) {
  set_step_params_kernel<<<1, 1, 0, s>>>(dst, v);
  return cudaGetLastError();
}

// =====================================================================================================
// Stand-alone sampler update for steps that re-use a predicted noise (the reference's reduced-call generator,
// regenerateFromtrain2.py:536,615-618): the same fp32 arithmetic, in the same order, as the fused epilogue of the output
// convolution (gemm_tc.cu, EPI_SAMPLER).  x, eps, noise: fp32 NCHW, elementwise.
// =====================================================================================================
__global__ void sampler_update_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                      int use_philox, unsigned long long seed, unsigned long long elem_offset, int step_index,
                                      float4 coef, int mode, size_t n) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (idx >= n) return;
  const float e = eps[idx];
  const float xv = x[idx];
  if (mode == STEP_DDPM) {
    float z = 0.f;
    if (noise)
      z = __ldg(noise + idx);
    else if (use_philox)
      z = philox_normal(seed, elem_offset + idx, static_cast<uint32_t>(step_index));
    const float inner = __fsub_rn(xv, __fmul_rn(coef.y, e));
    x[idx] = __fadd_rn(__fmul_rn(coef.x, inner), __fmul_rn(coef.z, z));
  } else if (mode == STEP_DDIM) {
    const float x0 = __fmul_rn(__fsub_rn(xv, __fmul_rn(coef.y, e)), coef.x);
    x[idx] = __fadd_rn(__fmul_rn(coef.z, x0), __fmul_rn(coef.w, e));
  }
}
cudaError_t sampler_update_launch(float* x, const float* eps, const float* noise, int use_philox, unsigned long long seed,
                                  unsigned long long elem_offset, int step_index, float4 coef, int mode, size_t n, cudaStream_t s) {
  if (mode != STEP_DDPM && mode != STEP_DDIM) return cudaErrorInvalidValue;
  return launch_pdl(sampler_update_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, s, x, eps, noise, use_philox, seed,
                    elem_offset, step_index, coef, mode, n);
}

// =====================================================================================================
// nearest 2x upsample, NHWC bf16
// =====================================================================================================
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int nv) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * 4 * H * W * nv;
  pdl_trigger();
  pdl_wait();
  if (idx >= total) return;
  const int v = idx % nv;
  size_t p = idx / nv;
  const int ox = p % (2 * W);
  p /= (2 * W);
  const int oy = p % (2 * H);
  const int b = p / (2 * H);
  out[idx] = __ldg(x + ((static_cast<size_t>(b) * H + (oy >> 1)) * W + (ox >> 1)) * nv + v);
}
cudaError_t upsample2x_launch(const __nv_bfloat16* x, __nv_bfloat16* out, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(B) * 4 * H * W * (C / 8);
  return launch_pdl(upsample2x_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s,
                    reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), B, H, W, C / 8);
}

// =====================================================================================================
// output head: GroupNorm32 + SiLU + conv3x3 (C -> 4) + sampler update, one CTA per sample (ops.cuh: OutHeadArgs)
// =====================================================================================================
constexpr int OH_THREADS = 256;
constexpr int OH_WROWS = 8;  // weight rows used: 4 hi + 4 lo
WD_DEVINL void oh_mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
static size_t out_head_smem_bytes(int H, int W, int C) {
  const size_t pitch = static_cast<size_t>(C) * 2 + 16;           // bytes per pixel row (+16: ldmatrix rows land in distinct banks)
  const size_t wpitch = (static_cast<size_t>(9) * C + 8) * 2;     // bytes per weight row (+16)
  return (static_cast<size_t>(H) * W + 1) * pitch + OH_WROWS * wpitch + static_cast<size_t>(C) * 8 + 256 + 64 +  // tables, statistics, mbarriers
         static_cast<size_t>(H) * W * 32;  // [HW][8] fp32 output tile of the shifted accumulation
}
bool out_head_supported(int H, int W, int C) {
  return H >= 1 && W >= 1 && (H * W) % 16 == 0 && C % 32 == 0 && C % 16 == 0 && (C / 32) % 2 == 0 && C <= 1024 &&
         out_head_smem_bytes(H, W, C) <= 227 * 1024;
}

__global__ void __launch_bounds__(OH_THREADS, 1) out_head_kernel(const OutHeadArgs a) {
  extern __shared__ __align__(128) uint8_t oh_smem[];
  const int HW = a.H * a.W, C = a.C, cpg = C / 32;
  const int pitch = C * 2 + 16, wpitch = (9 * C + 8) * 2;
  uint8_t* s_act = oh_smem;                                  // [HW + 1][pitch]: row HW is the zero (padding) row
  uint8_t* s_w = s_act + static_cast<size_t>(HW + 1) * pitch;  // [8][wpitch]
  float* s_sc = reinterpret_cast<float*>(s_w + OH_WROWS * wpitch);  // [C] scale / 2, then [C] shift / 2
  float* s_sh = s_sc + C;
  float* s_mean = s_sh + C;  // [32], [32]: the 256 spare bytes of out_head_smem_bytes (no static shared memory: the dynamic
  float* s_rstd = s_mean + 32;  // allocation may then use the whole 227 KB)
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nv = C >> 3;  // 16-byte vectors per pixel row

  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rstd + 32);  // [3]: weights landed, first / second half of the image landed
  // ---- everything arrives by bulk copies (one per weight row / pixel row, straight into the padded rows): 210 KB in flight
  // per SM instead of a few dependent 16-byte loads per thread ----
  const uint32_t row_bytes = static_cast<uint32_t>(C) * 2, wrow_bytes = static_cast<uint32_t>(9 * C) * 2;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  for (int i = tid; i < pitch / 16; i += OH_THREADS) *reinterpret_cast<uint4*>(s_act + static_cast<size_t>(HW) * pitch + i * 16) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (warp == 0) {  // the weights do not depend on the previous kernel
    if (lane == 0) mbar_arrive_expect_tx(&bars[0], OH_WROWS * wrow_bytes);
    __syncwarp();
    if (lane < OH_WROWS) bulk_load_1d(s_w + lane * wpitch, a.w + static_cast<size_t>(lane) * 9 * C, wrow_bytes, &bars[0]);
  }
  pdl_trigger();
  pdl_wait();  // h and its statistics come from the previous kernel
  if (warp == 0) {
    const int half_rows = HW / 2;  // HW % 16 == 0
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[1], static_cast<uint32_t>(half_rows) * row_bytes);
      mbar_arrive_expect_tx(&bars[2], static_cast<uint32_t>(HW - half_rows) * row_bytes);
    }
    __syncwarp();
    const uint8_t* src = reinterpret_cast<const uint8_t*>(a.h + static_cast<size_t>(b) * HW * C);
    for (int p = lane; p < HW; p += 32)
      bulk_load_1d(s_act + static_cast<size_t>(p) * pitch, src + static_cast<size_t>(p) * row_bytes, row_bytes, &bars[p < half_rows ? 1 : 2]);
  }
  // ---- statistics of this sample (the arithmetic of groupnorm_apply_bulk_kernel) -> scale / shift tables ----
  if (tid >= 32 && tid < 64) {
    const int g = tid - 32;
    const float2* part = reinterpret_cast<const float2*>(a.partial) + (static_cast<size_t>(b) * 32 + g) * a.pslots;
    float S = 0.f, Q = 0.f;
    for (int i = 0; i < a.pslots; ++i) {
      const float2 t = __ldg(part + i);
      S += t.x;
      Q += t.y;
    }
    const float inv_n = 1.0f / static_cast<float>(cpg * HW);
    const float mean = S * inv_n;
    const float var = fmaxf(fmaf(-mean, mean, Q * inv_n), 0.f);
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + a.gn_eps);
  }
  __syncthreads();
  for (int c = tid; c < C; c += OH_THREADS) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * __ldg(a.gamma + c);
    const float sh = fmaf(-s_mean[g], sc, __ldg(a.beta + c));
    s_sc[c] = 0.5f * sc;
    s_sh[c] = 0.5f * sh;
  }
  __syncthreads();
  // ---- h -> silu(GroupNorm(h)) as bf16, in place in the shared-memory image: a thread keeps ONE 8-channel column (its scale /
  // shift pairs stay in registers) and walks the rows; the first half is normalised while the second is still landing ----
  {
    const int rows_par = OH_THREADS / nv;  // rows in flight (threads beyond rows_par * nv idle: 16 of 256 at C = 320)
    const int col = tid % nv, r0 = tid / nv;
    if (r0 < rows_par) {
      const float4 c0 = *reinterpret_cast<const float4*>(s_sc + col * 8), c1 = *reinterpret_cast<const float4*>(s_sc + col * 8 + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(s_sh + col * 8), h1 = *reinterpret_cast<const float4*>(s_sh + col * 8 + 4);
      const float2 sc2[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w)};
      const float2 sh2[4] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y), make_float2(h1.z, h1.w)};
      const int half_rows = HW / 2;
      mbar_wait(&bars[1], 0);
      int p = r0;
      for (; p < half_rows; p += rows_par) {
        uint4* sp = reinterpret_cast<uint4*>(s_act + static_cast<size_t>(p) * pitch + col * 16);
        *sp = gn_vec8<true, true>(*sp, sc2, sh2);
      }
      mbar_wait(&bars[2], 0);
      for (; p < HW; p += rows_par) {
        uint4* sp = reinterpret_cast<uint4*>(s_act + static_cast<size_t>(p) * pitch + col * 16);
        *sp = gn_vec8<true, true>(*sp, sc2, sh2);
      }
    }
  }
  mbar_wait(&bars[0], 0);
  __syncthreads();

  // ---- 3 x 3 convolution on mma.sync, SOURCE-tile form: a warp loads the A fragments of a 16-pixel source tile ONCE per K step
  // (no shift, no padding row) and multiplies them by all nine taps' weight fragments (nine accumulators of 16 x 8 per tile);
  // tap (dy, dx)'s accumulator of source pixel (y, x) then belongs to output pixel (y - dy, x - dx).  The first version walked
  // output tiles and re-read the image through ldmatrix once per tap: 1.5 MB of shared-memory reads per CTA, the kernel's bound
  // (profiles/R4p_out_head.txt); this form reads it once.  The shifted accumulators are added into a [HW][8] fp32 tile tap by
  // tap (within a tap every output pixel receives from exactly one thread: plain adds, fixed order, deterministic). ----
  const float4 coef = a.sp ? a.sp->coef : a.coef;
  const int mode = a.sp ? a.sp->mode : a.mode;
  const int use_philox = a.sp ? a.sp->use_philox : a.use_philox;
  const unsigned long long seed = a.sp ? a.sp->seed : a.seed;
  const unsigned long long sample_offset = a.sp ? a.sp->sample_offset : a.sample_offset;
  const int step_index = a.sp ? a.sp->step_index : a.step_index;
  const int g = lane >> 2, tq = lane & 3;
  const int ntile = HW / 16;
  const uint32_t act_base = smem_u32(s_act);
  const int ksteps = C / 16;
  float* s_out = reinterpret_cast<float*>(bars + 4);  // [HW][8] fp32
  for (int i = tid; i < HW * 8; i += OH_THREADS) s_out[i] = 0.f;
  constexpr int OH_WARPS = OH_THREADS / 32;
  const int rounds = (ntile + 2 * OH_WARPS - 1) / (2 * OH_WARPS);
  for (int rd = 0; rd < rounds; ++rd) {
    const int t0 = (rd * OH_WARPS + warp) * 2;
    const bool have0 = t0 < ntile, have1 = t0 + 1 < ntile;
    float acc0[9][4], acc1[9][4];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc0[tp][i] = acc1[tp][i] = 0.f;
    if (have0) {
      // this lane's ldmatrix row: source pixel (tile * 16 + r), 8-column half kh
      const int r = (lane & 7) + ((lane >> 3) & 1) * 8, kh = lane >> 4;
      const uint32_t ra0 = act_base + static_cast<uint32_t>(t0 * 16 + r) * pitch + kh * 16;
      const uint32_t ra1 = act_base + static_cast<uint32_t>((have1 ? t0 + 1 : t0) * 16 + r) * pitch + kh * 16;
      const uint8_t* wrow = s_w + g * wpitch + (2 * tq) * 2;
#pragma unroll 1
      for (int ks = 0; ks < ksteps; ++ks) {
        uint32_t f0[4], f1[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                     : "=r"(f0[0]), "=r"(f0[1]), "=r"(f0[2]), "=r"(f0[3])
                     : "r"(ra0 + ks * 32));
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                     : "=r"(f1[0]), "=r"(f1[1]), "=r"(f1[2]), "=r"(f1[3])
                     : "r"(ra1 + ks * 32));
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
          const uint8_t* wp = wrow + (static_cast<size_t>(tp) * C + ks * 16) * 2;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wp);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wp + 16);
          oh_mma_16816(acc0[tp], f0, b0, b1);
          oh_mma_16816(acc1[tp], f1, b0, b1);
        }
      }
    }
    // shifted accumulation, tap by tap (every thread takes part in the barriers)
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      const int dy = tp / 3 - 1, dx = tp % 3 - 1;
      __syncthreads();
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        if (!(tt ? have1 : have0)) continue;
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int p = (t0 + tt) * 16 + g + hr * 8;  // source pixel
          const int y = p / a.W, x = p - y * a.W;
          const int oy = y - dy, ox = x - dx;
          if (oy >= 0 && oy < a.H && ox >= 0 && ox < a.W) {
            float2* dst = reinterpret_cast<float2*>(s_out + static_cast<size_t>(oy * a.W + ox) * 8 + 2 * tq);
            float2 v = *dst;
            v.x += tt ? acc1[tp][2 * hr] : acc0[tp][2 * hr];
            v.y += tt ? acc1[tp][2 * hr + 1] : acc0[tp][2 * hr + 1];
            *dst = v;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- epilogue: eps = (hi + lo) + bias, sampler update (the arithmetic of gemm_tc.cu's EPI_SAMPLER); one pixel per thread and
  // pass, so every global access of a warp is 32 consecutive floats ----
  for (int pix = tid; pix < HW; pix += OH_THREADS) {
    const float4 hi = *reinterpret_cast<const float4*>(s_out + static_cast<size_t>(pix) * 8);
    const float4 lo = *reinterpret_cast<const float4*>(s_out + static_cast<size_t>(pix) * 8 + 4);
    const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
    size_t idx[4];
    float eps[4], xv[4], z[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      idx[o] = (static_cast<size_t>(b) * 4 + o) * HW + pix;
      eps[o] = (hv[o] + lv[o]) + __ldg(a.bias + o);
      xv[o] = 0.f;
      z[o] = 0.f;
      if (mode != STEP_EPS_ONLY) {
        xv[o] = a.x[idx[o]];
        if (mode == STEP_DDPM && a.noise) z[o] = __ldg(a.noise + idx[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (a.eps_out) a.eps_out[idx[o]] = eps[o];
      if (mode == STEP_DDPM) {
        float zz = z[o];
        if (!a.noise && use_philox) zz = philox_normal(seed, sample_offset * (4ull * HW) + idx[o], static_cast<uint32_t>(step_index));
        const float inner = __fsub_rn(xv[o], __fmul_rn(coef.y, eps[o]));
        a.x[idx[o]] = __fadd_rn(__fmul_rn(coef.x, inner), __fmul_rn(coef.z, zz));
      } else if (mode == STEP_DDIM) {
        const float x0v = __fmul_rn(__fsub_rn(xv[o], __fmul_rn(coef.y, eps[o])), coef.x);
        a.x[idx[o]] = __fadd_rn(__fmul_rn(coef.z, x0v), __fmul_rn(coef.w, eps[o]));
      }
    }
  }
}

cudaError_t out_head_launch(const OutHeadArgs& a, cudaStream_t s) {
  if (!out_head_supported(a.H, a.W, a.C) || a.B < 1 || a.pslots < 1 || !a.h || !a.partial || !a.w || !a.bias) return cudaErrorInvalidValue;
  const size_t smem = out_head_smem_bytes(a.H, a.W, a.C);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(out_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  if (attr_err != cudaSuccess) return attr_err;
  return launch_pdl(out_head_kernel, dim3(a.B), dim3(OH_THREADS), smem, s, a);
}

// =====================================================================================================
// sub-pixel weights of upsample + conv3x3 (ops.cuh)
// =====================================================================================================
__global__ void upconv_phase_fold_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin, int as_f16) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(4) * Cout * 4 * Cin;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % Cin);
  size_t r = idx / Cin;
  const int t = static_cast<int>(r % 4);
  r /= 4;
  const int n = static_cast<int>(r % Cout);
  const int ph = static_cast<int>(r / Cout);
  const int a = ph >> 1, b = ph & 1, ty = t >> 1, tx = t & 1;
  // 3x3 taps that read input row (ty - 1 + a): a = 0: ty = 0 <- {0}, ty = 1 <- {1, 2};  a = 1: ty = 0 <- {0, 1}, ty = 1 <- {2}
  const int ky0 = a == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = a == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
  const int kx0 = b == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = b == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
  const float* wn = w + (static_cast<size_t>(n) * Cin + c) * 9;
  float acc = 0.f;
  for (int ky = ky0; ky <= ky1; ++ky)
    for (int kx = kx0; kx <= kx1; ++kx) acc += wn[ky * 3 + kx];
  if (as_f16) {
    const __half h = __float2half_rn(acc);
    dst[idx] = *reinterpret_cast<const __nv_bfloat16*>(&h);
  } else {
    dst[idx] = __float2bfloat16(acc);
  }
}
cudaError_t upconv_phase_fold_launch(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int as_f16, cudaStream_t s) {
  const size_t total = static_cast<size_t>(4) * Cout * 4 * Cin;
  upconv_phase_fold_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, dst, Cout, Cin, as_f16);
  return cudaGetLastError();
}

// =====================================================================================================
// weight repacking
// =====================================================================================================
WD_DEVINL int geglu_perm(int n, int N, int bn) {
  // rows [0, N/2) are values, [N/2, N) gates (torch chunk(2), unet.py:128).  Tile t of width bn holds
  // values t*bn/2 .. in its first half and the matching gates in its second half.
  const int half = bn / 2, Nh = N / 2;
  const bool gate = n >= Nh;
  const int j = gate ? n - Nh : n;
  return (j / half) * bn + (gate ? half : 0) + (j % half);
}

WD_DEVINL __nv_bfloat16 to_16(float v, int as_f16) {  // bf16, or the fp16 bit pattern in a 16-bit slot
  if (!as_f16) return __float2bfloat16(v);
  const __half h = __float2half_rn(v);
  return *reinterpret_cast<const __nv_bfloat16*>(&h);
}

__global__ void repack_conv3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin,
                                      int ldk, int k_off, int as_f16) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(Cout) * Cin * 9;
  if (idx >= total) return;
  const int tap = idx % 9;
  const int c = (idx / 9) % Cin;
  const int n = idx / (9 * static_cast<size_t>(Cin));
  if (as_f16 == 2) {
    // output conv (N = 4 of a 16-column tile): bf16 hi part in row n, bf16 lo part (w - hi) in row n + 4; the epilogue
    // adds the two accumulator columns, so the weights act with ~16 mantissa bits
    const float hi = __bfloat162float(__float2bfloat16(w[idx]));
    dst[static_cast<size_t>(n) * ldk + k_off + tap * Cin + c] = __float2bfloat16(hi);
    dst[static_cast<size_t>(n + 4) * ldk + k_off + tap * Cin + c] = __float2bfloat16(w[idx] - hi);
    return;
  }
  dst[static_cast<size_t>(n) * ldk + k_off + tap * Cin + c] = to_16(w[idx], as_f16);
}
cudaError_t repack_conv3x3_launch(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int ldk, int k_off, int as_f16,
                                  cudaStream_t s) {
  const size_t total = static_cast<size_t>(Cout) * Cin * 9;
  repack_conv3x3_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, dst, Cout, Cin, ldk, k_off, as_f16);
  return cudaGetLastError();
}

__global__ void repack_linear_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int N, int K,
                                     int ldk, int k_off, int n_off, int geglu_bn, int as_f16) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(N) * K) return;
  const int k = idx % K, n = idx / K;
  const int nn = geglu_bn ? geglu_perm(n, N, geglu_bn) : n;
  dst[static_cast<size_t>(nn + n_off) * ldk + k_off + k] = to_16(w[idx], as_f16);
}
cudaError_t repack_linear_launch(const float* w, __nv_bfloat16* dst, int N, int K, int ldk, int k_off, int n_off,
                                 int geglu_bn, int as_f16, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * K;
  repack_linear_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, dst, N, K, ldk, k_off, n_off,
                                                                                    geglu_bn, as_f16);
  return cudaGetLastError();
}

__global__ void repack_vec_kernel(const float* __restrict__ v, float* __restrict__ dst, int N, int n_off, int geglu_bn,
                                  int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int nn = (geglu_bn ? geglu_perm(n, N, geglu_bn) : n) + n_off;
  dst[nn] = accumulate ? dst[nn] + v[n] : v[n];
}
cudaError_t repack_vec_launch(const float* v, float* dst, int N, int n_off, int geglu_bn, int accumulate,
                              cudaStream_t s) {
  repack_vec_kernel<<<(N + 255) / 256, 256, 0, s>>>(v, dst, N, n_off, geglu_bn, accumulate);
  return cudaGetLastError();
}

// LayerNorm folded into the Linear that follows it (gemm_tc.cuh, GemmArgs::ln_stats): one warp per weight row n
//   W'[n,k] = fp16(gamma[k] W[n,k]),  s[n] = sum_k W'[n,k] (of the ROUNDED values the MMA will use),  b'[n] = sum_k beta[k] W[n,k] + b[n]
__global__ void __launch_bounds__(256) fold_ln_linear_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ bias,
                                                             __nv_bfloat16* __restrict__ dst, float* __restrict__ s_out,
                                                             float* __restrict__ b_out, int N, int K, int ldk, int n_off,
                                                             int geglu_bn) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int nn = (geglu_bn ? geglu_perm(n, N, geglu_bn) : n) + n_off;
  float s = 0.f, b = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[static_cast<size_t>(n) * K + k];
    const __half h = __float2half_rn(gamma[k] * wv);
    dst[static_cast<size_t>(nn) * ldk + k] = *reinterpret_cast<const __nv_bfloat16*>(&h);
    s += __half2float(h);
    b = fmaf(beta[k], wv, b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    s_out[nn] = s;
    b_out[nn] = b + (bias ? bias[n] : 0.f);
  }
}
cudaError_t fold_ln_linear_launch(const float* w, const float* gamma, const float* beta, const float* bias, __nv_bfloat16* dst,
                                  float* s_out, float* b_out, int N, int K, int ldk, int n_off, int geglu_bn, cudaStream_t s) {
  fold_ln_linear_kernel<<<(N * 32 + 255) / 256, 256, 0, s>>>(w, gamma, beta, bias, dst, s_out, b_out, N, K, ldk, n_off, geglu_bn);
  return cudaGetLastError();
}

// conv_in weight [Cout, 4, 3, 3] fp32 -> bf16 [Cout, 128]: k = j and 36 + j hold w_hi[n][j], 72 + j holds w_lo, rest 0
__global__ void repack_conv_in_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * 128) return;
  const int k = idx % 128, n = idx / 128;
  float v = 0.f;
  if (k < 108) {
    const float wv = w[n * 36 + k % 36];
    const float hi = __bfloat162float(__float2bfloat16(wv));
    v = k < 72 ? hi : wv - hi;
  }
  dst[idx] = __float2bfloat16(v);
}
cudaError_t repack_conv_in_launch(const float* w, __nv_bfloat16* dst, int Cout, int Cin, cudaStream_t s) {
  if (Cin != 4) return cudaErrorInvalidValue;
  const int total = Cout * 128;
  repack_conv_in_kernel<<<(total + 255) / 256, 256, 0, s>>>(w, dst, Cout);
  return cudaGetLastError();
}

// =====================================================================================================
// fp32 context encoder (time-invariant; runs once per trajectory).  Kept in fp32 because the reference's
// Word_Attention softmax is unscaled (unet.py:831-832) and therefore close to one-hot.
// =====================================================================================================
__global__ void embed_tokens_kernel(const void* __restrict__ tokens, int is_i64, const float* __restrict__ E, int vocab,
                                    const float* __restrict__ pe, int add_pe, float* __restrict__ out, int B, int L,
                                    int D) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(B) * L * D) return;
  const int d = idx % D;
  const size_t bl = idx / D;
  const int l = bl % L;
  long long tok = is_i64 ? static_cast<const long long*>(tokens)[bl] : static_cast<const int*>(tokens)[bl];
  if (tok < 0 || tok >= vocab) __trap();  // nn.Embedding raises on out-of-range ids
  float v = __ldg(E + tok * D + d);
  if (add_pe) v += __ldg(pe + static_cast<size_t>(l) * D + d);
  out[idx] = v;
}
cudaError_t embed_tokens_launch(const void* tokens, int tokens_are_i64, const float* E, int vocab, const float* pe,
                                int add_pe, float* out, int B, int L, int D, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * L * D;
  embed_tokens_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(tokens, tokens_are_i64, E, vocab, pe,
                                                                                  add_pe, out, B, L, D);
  return cudaGetLastError();
}

// 64x64 output tile, 256 threads, 4x4 outputs per thread, K tile 16.
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ Wt,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int M, int N, int K) {
  __shared__ float sx[16][64 + 1];
  __shared__ float sw[16][64 + 1];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i / 16, kk = i % 16;
      const int m = m0 + r, n = n0 + r, k = k0 + kk;
      sx[kk][r] = (m < M && k < K) ? __ldg(x + static_cast<size_t>(m) * K + k) : 0.f;
      sw[kk][r] = (n < N && k < K) ? __ldg(Wt + static_cast<size_t>(n) * K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float xa[4], wb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xa[i] = sx[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wb[j] = sw[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += xa[i] * wb[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) out[static_cast<size_t>(m) * N + n] = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
    }
  }
}
cudaError_t linear_f32_launch(const float* x, const float* W, const float* b, float* out, int M, int N, int K,
                              cudaStream_t s) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  linear_f32_kernel<<<grid, 256, 0, s>>>(x, W, b, out, M, N, K);
  return cudaGetLastError();
}

// One warp per query row; scores kept in shared memory; unscaled softmax(q k^T) v.
__global__ void __launch_bounds__(128) word_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                        const float* __restrict__ v, __nv_bfloat16* __restrict__ ctx,
                                                        float* __restrict__ ctx_f32, int L, int D, int Ltot,
                                                        int row_off) {
  extern __shared__ float wa_smem[];  // [4][L] scores + [4][D] query
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int qi = blockIdx.x * 4 + warp;
  if (qi >= L) return;
  float* sc = wa_smem + warp * L;
  float* sq = wa_smem + 4 * L + warp * D;
  const float* qr = q + (static_cast<size_t>(b) * L + qi) * D;
  for (int d = lane; d < D; d += 32) sq[d] = __ldg(qr + d);
  __syncwarp();
  float mx = -INFINITY;
  for (int j = 0; j < L; ++j) {
    const float* kr = k + (static_cast<size_t>(b) * L + j) * D;
    float dsum = 0.f;
    for (int d = lane; d < D; d += 32) dsum += sq[d] * __ldg(kr + d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) sc[j] = dsum;
    mx = fmaxf(mx, dsum);
  }
  __syncwarp();
  float den = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    den += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  __syncwarp();
  const float inv = 1.0f / den;
  const size_t orow = (static_cast<size_t>(b) * Ltot + row_off + qi) * D;
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int d = d0 + lane;
    float acc = 0.f;
    if (d < D) {
      for (int j = 0; j < L; ++j) acc += sc[j] * __ldg(v + (static_cast<size_t>(b) * L + j) * D + d);
      acc *= inv;
      if (ctx) ctx[orow + d] = __float2bfloat16(acc);
      if (ctx_f32) ctx_f32[orow + d] = acc;
    }
  }
}
// Word_Attention over a segment WITHOUT positional encoding (the 769 PHOSC tokens of UNetModelPhosc, unetPhosc.py:726-729):
// q_i, k_j, v_j depend only on the token values, q_i = TQ[tok_i] etc., so softmax_j(q_i . k_j) v_j depends only on tok_i and on
// how often each token value occurs in the sample:
//   out_i = R[tok_i],  R[a] = sum_t n_t exp(G[a,t] - m_a) TV[t] / sum_t n_t exp(G[a,t] - m_a),  G = TQ TK^T  [vocab, vocab].
// The same sums as the reference's 769 x 769 attention with equal terms grouped: O(vocab^2 D) per sample instead of O(L^2 D)
// (the straightforward kernel took 57 ms at batch 256 -- a fifth of a 50-step DDIM trajectory).  One CTA per sample.
__global__ void __launch_bounds__(320) word_attn_hist_kernel(const void* __restrict__ tokens, int is_i64, const float* __restrict__ G,
                                                             const float* __restrict__ TV, int vocab, __nv_bfloat16* __restrict__ ctx,
                                                             int L, int D, int Ltot, int row_off) {
  extern __shared__ float wh_smem[];  // [vocab][D] R rows, [vocab] weights, [vocab] counts
  float* R = wh_smem;
  float* w = R + static_cast<size_t>(vocab) * D;
  int* cnt = reinterpret_cast<int*>(w + vocab);
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < vocab; t += blockDim.x) cnt[t] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const size_t bl = static_cast<size_t>(b) * L + i;
    const long long tok = is_i64 ? static_cast<const long long*>(tokens)[bl] : static_cast<const int*>(tokens)[bl];
    if (tok < 0 || tok >= vocab) __trap();  // nn.Embedding raises on out-of-range ids
    atomicAdd(&cnt[tok], 1);
  }
  __syncthreads();
  for (int a = 0; a < vocab; ++a) {
    if (cnt[a] == 0) continue;  // block-uniform
    float mx = -INFINITY;
    for (int t = 0; t < vocab; ++t)
      if (cnt[t] > 0) mx = fmaxf(mx, __ldg(G + a * vocab + t));
    __syncthreads();  // the previous row's weights are no longer read
    for (int t = threadIdx.x; t < vocab; t += blockDim.x)
      w[t] = cnt[t] > 0 ? static_cast<float>(cnt[t]) * expf(__ldg(G + a * vocab + t) - mx) : 0.f;
    __syncthreads();
    float den = 0.f;
    for (int t = 0; t < vocab; ++t) den += w[t];
    const float inv = 1.0f / den;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float acc = 0.f;
      for (int t = 0; t < vocab; ++t)
        if (w[t] != 0.f) acc = fmaf(w[t], __ldg(TV + static_cast<size_t>(t) * D + d), acc);
      R[static_cast<size_t>(a) * D + d] = acc * inv;
    }
  }
  __syncthreads();
  const int nv = D >> 1;  // bf16 pairs
  for (size_t idx = threadIdx.x; idx < static_cast<size_t>(L) * nv; idx += blockDim.x) {
    const int i = idx / nv, dp = idx % nv;
    const size_t bl = static_cast<size_t>(b) * L + i;
    const long long tok = is_i64 ? static_cast<const long long*>(tokens)[bl] : static_cast<const int*>(tokens)[bl];
    const float2 r = *reinterpret_cast<const float2*>(R + static_cast<size_t>(tok) * D + dp * 2);
    *reinterpret_cast<uint32_t*>(ctx + (static_cast<size_t>(b) * Ltot + row_off + i) * D + dp * 2) = pack_bf16x2(r.x, r.y);
  }
}
cudaError_t word_attn_hist_launch(const void* tokens, int tokens_are_i64, const float* G, const float* TV, int vocab,
                                  __nv_bfloat16* ctx_out, int B, int L, int D, int Ltot, int row_off, cudaStream_t s) {
  if (D % 2) return cudaErrorInvalidValue;
  const size_t smem = (static_cast<size_t>(vocab) * D + 2 * vocab) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(word_attn_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  word_attn_hist_kernel<<<B, 320, smem, s>>>(tokens, tokens_are_i64, G, TV, vocab, ctx_out, L, D, Ltot, row_off);
  return cudaGetLastError();
}
// G[a, t] = TQ[a] . TK[t]  (vocab x vocab, fp32; built once per weight load)
__global__ void word_attn_gram_kernel(const float* __restrict__ TQ, const float* __restrict__ TK, float* __restrict__ G, int vocab, int D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= vocab * vocab) return;
  const int a = idx / vocab, t = idx % vocab;
  float acc = 0.f;
  for (int d = 0; d < D; ++d) acc = fmaf(__ldg(TQ + static_cast<size_t>(a) * D + d), __ldg(TK + static_cast<size_t>(t) * D + d), acc);
  G[idx] = acc;
}
cudaError_t word_attn_gram_launch(const float* TQ, const float* TK, float* G, int vocab, int D, cudaStream_t s) {
  word_attn_gram_kernel<<<(vocab * vocab + 127) / 128, 128, 0, s>>>(TQ, TK, G, vocab, D);
  return cudaGetLastError();
}

// Short sequences (the 10 character tokens): one CTA per sample, one warp per query row, K and V of the sample staged in shared
// memory once (the general kernel below re-reads them from global memory inside its key loop: 41 us at batch 256 for 20 MFLOP).
__global__ void __launch_bounds__(512) word_attn_short_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ v, __nv_bfloat16* __restrict__ ctx,
                                                              float* __restrict__ ctx_f32, int L, int D, int Ltot, int row_off) {
  extern __shared__ float ws_smem[];  // K [L][D], V [L][D], Q [L][D], scores [L][L]
  float* sk = ws_smem;
  float* sv = sk + L * D;
  float* sq = sv + L * D;
  float* ssc = sq + L * D;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = static_cast<size_t>(b) * L * D;
  for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
    sk[i] = __ldg(k + base + i);
    sv[i] = __ldg(v + base + i);
    sq[i] = __ldg(q + base + i);
  }
  __syncthreads();
  if (warp >= L) return;
  const float* qr = sq + warp * D;
  float* sc = ssc + warp * L;
  float mx = -INFINITY;
  for (int j = 0; j < L; ++j) {
    float dsum = 0.f;
    for (int d = lane; d < D; d += 32) dsum += qr[d] * sk[j * D + d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) sc[j] = dsum;
    mx = fmaxf(mx, dsum);
  }
  __syncwarp();
  float den = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    den += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  __syncwarp();
  const float inv = 1.0f / den;
  const size_t orow = (static_cast<size_t>(b) * Ltot + row_off + warp) * D;
  for (int d = lane; d < D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc += sc[j] * sv[j * D + d];
    acc *= inv;
    if (ctx) ctx[orow + d] = __float2bfloat16(acc);
    if (ctx_f32) ctx_f32[orow + d] = acc;
  }
}

cudaError_t word_attn_launch(const float* q, const float* k, const float* v, __nv_bfloat16* ctx_out, float* ctx_out_f32,
                             int B, int L, int D, int Ltot, int row_off, cudaStream_t s) {
  if (L <= 16 && (static_cast<size_t>(3) * L * D + L * L) * sizeof(float) <= 96 * 1024) {
    const size_t sm = (static_cast<size_t>(3) * L * D + L * L) * sizeof(float);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
      attr_err = cudaFuncSetAttribute(word_attn_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    if (attr_err != cudaSuccess) return attr_err;
    word_attn_short_kernel<<<B, 32 * L, sm, s>>>(q, k, v, ctx_out, ctx_out_f32, L, D, Ltot, row_off);
    return cudaGetLastError();
  }
  const size_t smem = (static_cast<size_t>(4) * L + 4 * D) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  dim3 grid((L + 3) / 4, B);
  word_attn_kernel<<<grid, 128, smem, s>>>(q, k, v, ctx_out, ctx_out_f32, L, D, Ltot, row_off);
  return cudaGetLastError();
}

}  // namespace wd
