// Backward / training-only kernels, see ops_bwd.cuh.  All HBM-/latency-bound: 16-byte accesses, fp32 arithmetic.
#include "ops_bwd.cuh"

#include <cstdlib>

#include <mutex>

namespace wd {

namespace {

WD_DEVINL void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = unpack_bf16x2(u[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
WD_DEVINL uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
WD_DEVINL void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// d/dz silu(z) = s (1 + z (1 - s)),  s = sigmoid(z)
WD_DEVINL float silu_grad_f(float z) {
  // MUFU.EX2 + MUFU.RCP (rel. error ~2^-22): the IEEE division expands to ~8 instructions plus a slow-path call, and ncu
  // showed the GroupNorm backward passes issue-bound, not HBM-bound
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  const float s = rcp_fast(1.0f + e);
  return s * fmaf(z, 1.0f - s, 1.0f);
}

}  // namespace

// =====================================================================================================
// GroupNorm (+SiLU) backward.  With g = dz * gamma (dz = dy * act'(z)), n = cpg * HW elements per (sample, group):
//   dx = rstd * (g - mean_n(g) - xhat * mean_n(g * xhat)),   dgamma_c = sum dz * xhat,   dbeta_c = sum dz
// Pass 1 reduces {sum dz, sum dz*xhat} per (sample, channel) into `ws` (no atomics); the group means are gamma-weighted
// sums of those, formed in the preamble of pass 2, whose first pixel chunk also adds its sample's share to dgamma / dbeta.
// =====================================================================================================
constexpr int GNB_R = 16;
constexpr int GNB_MAX_RCHUNKS = 4;  // row chunks of the reduce pass (workspace = B * 4 * C * 2 floats)

// mean / rstd of the groups of one slab from the forward partial statistics (same fold order as groupnorm_apply_kernel)
WD_DEVINL void gn_group_stats(const GroupNormBwdArgs& a, int b, int slab, int g, float& mean, float& rstd) {
  const int merge = a.cpg / a.pcpg;
  const int PG = a.Cs / a.pcpg;
  const int slots = a.pslots[slab];
  const float2* part = reinterpret_cast<const float2*>(a.partial[slab]) +
                       (static_cast<size_t>(b) * PG + static_cast<size_t>(g) * merge) * slots;
  float S = 0.f, Q = 0.f;
  const int n = merge * slots;
  for (int k = 0; k < n; ++k) {
    const float2 t = __ldg(part + k);
    S += t.x;
    Q += t.y;
  }
  const float inv_n = 1.0f / static_cast<float>(a.cpg * a.HW);
  mean = S * inv_n;
  rstd = rsqrtf(fmaxf(Q * inv_n - mean * mean, 0.f) + a.eps);
}

__global__ void __launch_bounds__(640) gn_bwd_reduce_kernel(const GroupNormBwdArgs a, int rchunks) {
  extern __shared__ float gnb_smem[];  // [R][Cs][2]
  __shared__ float s_mean[128], s_rstd[128];
  const int b = blockIdx.x, slab = blockIdx.y, rc = blockIdx.z;
  const int Cs = a.Cs, cpg = a.cpg;
  const int nv = Cs >> 3;
  const int R = blockDim.x / nv;
  const int col = threadIdx.x % nv, rl = threadIdx.x / nv;
  const int ng = Cs / cpg;
  if (static_cast<int>(threadIdx.x) < ng) gn_group_stats(a, b, slab, threadIdx.x, s_mean[threadIdx.x], s_rstd[threadIdx.x]);
  __syncthreads();
  float gm[8], be[8], mu[8], rs[8];
  load8f(a.gamma + slab * Cs + col * 8, gm);
  load8f(a.beta + slab * Cs + col * 8, be);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (col * 8 + j) / cpg;
    mu[j] = s_mean[g];
    rs[j] = s_rstd[g];
  }
  const int P = a.HW / rchunks;  // pixel rows of this CTA
  const size_t row0 = static_cast<size_t>(b) * a.HW + static_cast<size_t>(rc) * P;
  const bf16_t* xb = a.x[slab] + row0 * a.x_ld[slab];
  const bf16_t* dyb = a.dy + row0 * a.dy_ld + slab * Cs;
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.f;
  if (rl < R) {
    for (int p = rl; p < P; p += R) {
      float x[8], dy[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * a.x_ld[slab]) + col), x);
      unpack8(__ldg(reinterpret_cast<const uint4*>(dyb + static_cast<size_t>(p) * a.dy_ld) + col), dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mu[j]) * rs[j];
        float dz = dy[j];
        if (a.silu) dz *= silu_grad_f(fmaf(xh, gm[j], be[j]));
        sa[j] += dz;
        sb[j] = fmaf(dz, xh, sb[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gnb_smem[(rl * Cs + col * 8 + j) * 2] = sa[j];
      gnb_smem[(rl * Cs + col * 8 + j) * 2 + 1] = sb[j];
    }
  }
  __syncthreads();
  const int Ctot = gridDim.y * Cs;
  for (int i = threadIdx.x; i < 2 * Cs; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < R; ++r) t += gnb_smem[r * Cs * 2 + i];
    a.ws[((static_cast<size_t>(b) * rchunks + rc) * Ctot + slab * Cs) * 2 + i] = t;  // ws[b][rchunk][channel][2]
  }
}

__global__ void __launch_bounds__(640) gn_bwd_apply_kernel(const GroupNormBwdArgs a, int nchunk, int rchunks) {
  __shared__ float s_mean[128], s_rstd[128], s_c1[128], s_c2[128];
  __shared__ float s_wa[1024], s_wb[1024];
  const int b = blockIdx.x, slab = blockIdx.y, chunk = blockIdx.z;
  const int Cs = a.Cs, cpg = a.cpg;
  const int nv = Cs >> 3;
  const int R = blockDim.x / nv;
  const int col = threadIdx.x % nv, rl = threadIdx.x / nv;
  const int ng = Cs / cpg;
  const int Ctot = gridDim.y * Cs;
  // fold the row chunks of pass 1: one thread per channel (independent loads), gamma-weighted values to shared memory
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    float wa = 0.f, wb = 0.f;
    for (int r = 0; r < rchunks; ++r) {
      const float2 w = __ldg(reinterpret_cast<const float2*>(a.ws) + (static_cast<size_t>(b) * rchunks + r) * Ctot + slab * Cs + c);
      wa += w.x;
      wb += w.y;
    }
    if (chunk == 0) {  // parameter gradients: dbeta_c += sum dz, dgamma_c += sum dz * xhat (this sample's share)
      atomicAdd(a.dbeta + slab * Cs + c, wa);
      atomicAdd(a.dgamma + slab * Cs + c, wb);
    }
    const float gmm = __ldg(a.gamma + slab * Cs + c);
    s_wa[c] = gmm * wa;
    s_wb[c] = gmm * wb;
  }
  if (static_cast<int>(threadIdx.x) < ng) gn_group_stats(a, b, slab, threadIdx.x, s_mean[threadIdx.x], s_rstd[threadIdx.x]);
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < ng) {
    const int g = threadIdx.x;
    float S1 = 0.f, S2 = 0.f;
    for (int c = 0; c < cpg; ++c) {
      S1 += s_wa[g * cpg + c];
      S2 += s_wb[g * cpg + c];
    }
    const float inv_n = 1.0f / static_cast<float>(cpg * a.HW);
    s_c1[g] = S1 * inv_n;
    s_c2[g] = S2 * inv_n;
  }
  __syncthreads();
  if (rl >= R) return;
  float gm[8], be[8], mu[8], rs[8], c1[8], c2[8];
  load8f(a.gamma + slab * Cs + col * 8, gm);
  load8f(a.beta + slab * Cs + col * 8, be);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (col * 8 + j) / cpg;
    mu[j] = s_mean[g];
    rs[j] = s_rstd[g];
    c1[j] = s_c1[g];
    c2[j] = s_c2[g];
  }
  const int P = a.HW / nchunk;
  const size_t row0 = static_cast<size_t>(b) * a.HW + static_cast<size_t>(chunk) * P;
  const bf16_t* xb = a.x[slab] + row0 * a.x_ld[slab];
  const bf16_t* dyb = a.dy + row0 * a.dy_ld + slab * Cs;
  bf16_t* dxb = a.dx[slab] + row0 * a.dx_ld[slab];
  const bf16_t* addb = a.add[slab] ? a.add[slab] + row0 * a.add_ld[slab] : nullptr;
  const bool acc = a.accumulate[slab] != 0;
  for (int p = rl; p < P; p += R) {
    float x[8], dy[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * a.x_ld[slab]) + col), x);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dyb + static_cast<size_t>(p) * a.dy_ld) + col), dy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - mu[j]) * rs[j];
      float dz = dy[j];
      if (a.silu) dz *= silu_grad_f(fmaf(xh, gm[j], be[j]));
      o[j] = rs[j] * (dz * gm[j] - c1[j] - xh * c2[j]);
    }
    if (addb) {
      float t[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(addb + static_cast<size_t>(p) * a.add_ld[slab]) + col), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += t[j];
    }
    uint4* dst = reinterpret_cast<uint4*>(dxb + static_cast<size_t>(p) * a.dx_ld[slab]) + col;
    if (acc) {
      float t[8];
      unpack8(*dst, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += t[j];
    }
    *dst = pack8(o);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused single-pass version (the one the trainer launches when the shape allows): one CTA = (sample, source tensor, 40-channel
// slice = whole groups).  x and dy of the slice (HW rows x 80 bytes each) are read ONCE into registers (<= 6 rows per thread),
// the per-channel {sum dz, sum dz xhat} are reduced inside the CTA, and dx is formed from the registers: 3 tensor passes of
// HBM traffic instead of 5, one launch instead of two, no workspace round trip.  Measured: only 0.1 ms of the 16.2 ms training
// step (batch 224) -- the 80-byte row pieces of a channel slice use sectors and DRAM pages poorly, which eats most of what the
// saved passes give; a row-sliced two-phase kernel with a cluster-level reduction is the better shape for a later round.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GNF_C = 40;      // channels per CTA
constexpr int GNF_ROWS = 51;   // row lanes: 255 of the 256 threads = 51 rows x 5 vectors of 8 channels
constexpr int GNF_MAXIT = 6;   // rows per thread (HW <= 306)

__global__ void __launch_bounds__(256, 2) gn_bwd_fused_kernel(const GroupNormBwdArgs a) {
  __shared__ float s_red[GNF_ROWS][GNF_C][2];  // per row-lane partial sums (16 KB)
  __shared__ float s_sum[GNF_C][2];            // {sum dz, sum dz xhat} per channel of this (sample, slice)
  __shared__ float s_mean[4], s_rstd[4], s_c1[4], s_c2[4];
  const int b = blockIdx.x, slab = blockIdx.y, slice = blockIdx.z;
  const int Cs = a.Cs, cpg = a.cpg;
  const int c0 = slice * GNF_C;          // first channel of the slice inside the slab
  const int ngc = GNF_C / cpg;           // groups in this slice
  const int t = threadIdx.x;
  const int vc = t % 5, rl = t / 5;      // vector column (8 channels), row lane
  const bool active = rl < GNF_ROWS;
  if (t < ngc) gn_group_stats(a, b, slab, c0 / cpg + t, s_mean[t], s_rstd[t]);
  // ---- the slice of x and dy into registers ----
  const size_t row0 = static_cast<size_t>(b) * a.HW;
  const bf16_t* xb = a.x[slab] + row0 * a.x_ld[slab] + c0;
  const bf16_t* dyb = a.dy + row0 * a.dy_ld + slab * Cs + c0;
  uint4 xr[GNF_MAXIT], dr[GNF_MAXIT];
#pragma unroll
  for (int i = 0; i < GNF_MAXIT; ++i) {
    const int p = rl + i * GNF_ROWS;
    if (active && p < a.HW) {
      xr[i] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * a.x_ld[slab]) + vc);
      dr[i] = __ldg(reinterpret_cast<const uint4*>(dyb + static_cast<size_t>(p) * a.dy_ld) + vc);
    }
  }
  __syncthreads();
  float gm[8], be[8], mu[8], rs[8];
  load8f(a.gamma + slab * Cs + c0 + vc * 8, gm);
  load8f(a.beta + slab * Cs + c0 + vc * 8, be);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (vc * 8 + j) / cpg;
    mu[j] = s_mean[g];
    rs[j] = s_rstd[g];
  }
  // ---- pass 1 (registers): per-channel sums of dz and dz * xhat over this thread's rows ----
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.f;
#pragma unroll
  for (int i = 0; i < GNF_MAXIT; ++i) {
    const int p = rl + i * GNF_ROWS;
    if (active && p < a.HW) {
      float x[8], dy[8];
      unpack8(xr[i], x);
      unpack8(dr[i], dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mu[j]) * rs[j];
        float dz = dy[j];
        if (a.silu) dz *= silu_grad_f(fmaf(xh, gm[j], be[j]));
        sa[j] += dz;
        sb[j] = fmaf(dz, xh, sb[j]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_red[rl][vc * 8 + j][0] = sa[j];
      s_red[rl][vc * 8 + j][1] = sb[j];
    }
  }
  __syncthreads();
  if (t < 2 * GNF_C) {  // fixed-order fold over the row lanes (bit-reproducible)
    const int c = t >> 1, k = t & 1;
    float acc = 0.f;
    for (int r = 0; r < GNF_ROWS; ++r) acc += s_red[r][c][k];
    s_sum[c][k] = acc;
    // parameter gradients: dbeta_c += sum dz, dgamma_c += sum dz * xhat (this sample's share)
    atomicAdd((k ? a.dgamma : a.dbeta) + slab * Cs + c0 + c, acc);
  }
  __syncthreads();
  if (t < ngc) {
    float S1 = 0.f, S2 = 0.f;
    for (int c = 0; c < cpg; ++c) {
      const float gmm = __ldg(a.gamma + slab * Cs + c0 + t * cpg + c);
      S1 = fmaf(gmm, s_sum[t * cpg + c][0], S1);
      S2 = fmaf(gmm, s_sum[t * cpg + c][1], S2);
    }
    const float inv_n = 1.0f / static_cast<float>(cpg * a.HW);
    s_c1[t] = S1 * inv_n;
    s_c2[t] = S2 * inv_n;
  }
  __syncthreads();
  if (!active) return;
  float c1[8], c2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (vc * 8 + j) / cpg;
    c1[j] = s_c1[g];
    c2[j] = s_c2[g];
  }
  // ---- pass 2 (registers): dx = rstd (dz gamma - c1 - xhat c2) (+ add) (+= existing) ----
  bf16_t* dxb = a.dx[slab] + row0 * a.dx_ld[slab] + c0;
  const bf16_t* addb = a.add[slab] ? a.add[slab] + row0 * a.add_ld[slab] + c0 : nullptr;
  const bool acc = a.accumulate[slab] != 0;
#pragma unroll
  for (int i = 0; i < GNF_MAXIT; ++i) {
    const int p = rl + i * GNF_ROWS;
    if (p < a.HW) {
      float x[8], dy[8], o[8];
      unpack8(xr[i], x);
      unpack8(dr[i], dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mu[j]) * rs[j];
        float dz = dy[j];
        if (a.silu) dz *= silu_grad_f(fmaf(xh, gm[j], be[j]));
        o[j] = rs[j] * (dz * gm[j] - c1[j] - xh * c2[j]);
      }
      if (addb) {
        float tt[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(addb + static_cast<size_t>(p) * a.add_ld[slab]) + vc), tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += tt[j];
      }
      uint4* dst = reinterpret_cast<uint4*>(dxb + static_cast<size_t>(p) * a.dx_ld[slab]) + vc;
      if (acc) {
        float tt[8];
        unpack8(*dst, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += tt[j];
      }
      *dst = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-sliced version (the one the trainer launches when the shape allows; round 2, session 3): a cluster of CL CTAs per
// (sample, source tensor), each CTA owning HW / CL whole pixel rows.  x and dy of its rows are bulk-copied ONCE into shared
// memory (one cp.async.bulk per row: fully coalesced, no 80-byte channel pieces), the per-channel {sum dz, sum dz xhat} are
// reduced inside the CTA, exchanged with the peer through distributed shared memory behind one cluster barrier, and dx is
// formed from the shared-memory copies: 3 tensor passes, every access a whole row; pass 1 leaves dz (bf16, like dy itself) in
// place of dy so that the SiLU gradient (two MUFU operations) is paid once.  Measured at batch 224: the 21 launches of a
// backward pass 1.79 -> 1.51 ms.  The first build (320 threads, SiLU gradient in both passes) took 1.85 ms -- exactly the
// channel-sliced kernel's time: this operator is bound by its arithmetic (37 instructions and 4 MUFU operations per element at
// 10-20 warps per SM), not by HBM or by the access pattern.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GNR_T = 640;  // 16 row lanes at 320 channels: 20 warps per SM (the arithmetic, not HBM, bounds this kernel)
struct GnrLayout {
  size_t xs, ds, red, sum, bar, total;
  int RL;
};
static __host__ __device__ GnrLayout gnr_layout(int rows, int Cs, int CL) {
  GnrLayout L;
  const int nv = Cs / 8;
  L.RL = GNR_T / nv;
  L.xs = 0;
  L.ds = static_cast<size_t>(rows) * Cs * 2;
  L.red = 2 * L.ds;                                        // [RL][Cs][2] fp32, later the totals [Cs][2]
  L.sum = L.red + static_cast<size_t>(L.RL) * Cs * 8;      // [CL][Cs][2] fp32: every rank's partial sums
  L.bar = L.sum + static_cast<size_t>(CL) * Cs * 8;        // mbarrier + group tables
  L.total = L.bar + 16 + 4 * 128 * 4;
  return L;
}
WD_DEVINL void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(cluster_addr), "f"(v) : "memory");
}

template <int CL>
WD_DEVINL void gn_bwd_rows_body(const GroupNormBwdArgs& a) {
  extern __shared__ __align__(128) uint8_t gnr_smem[];
  const int Cs = a.Cs, cpg = a.cpg, nv = Cs >> 3, ng = Cs / cpg;
  const int rows = a.HW / CL;
  const GnrLayout L = gnr_layout(rows, Cs, CL);
  const int RL = L.RL;
  uint8_t* xs = gnr_smem + L.xs;
  uint8_t* ds = gnr_smem + L.ds;
  float* s_red = reinterpret_cast<float*>(gnr_smem + L.red);
  float* s_sum = reinterpret_cast<float*>(gnr_smem + L.sum);
  uint64_t* bar = reinterpret_cast<uint64_t*>(gnr_smem + L.bar);
  float* s_mean = reinterpret_cast<float*>(gnr_smem + L.bar + 16);
  float* s_rstd = s_mean + 128;
  float* s_c1 = s_rstd + 128;
  float* s_c2 = s_c1 + 128;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int b = blockIdx.x / CL, slab = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int vc = t % nv, rl = t / nv;
  const bool active = rl < RL;
  const size_t row0 = static_cast<size_t>(b) * a.HW + static_cast<size_t>(rank) * rows;
  const uint32_t row_bytes = static_cast<uint32_t>(Cs) * 2;

  if (t == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) mbar_arrive_expect_tx(bar, 2u * static_cast<uint32_t>(rows) * row_bytes);
    __syncwarp();
    const bf16_t* xb = a.x[slab] + row0 * a.x_ld[slab];
    const bf16_t* dyb = a.dy + row0 * a.dy_ld + slab * Cs;
    for (int p = lane; p < rows; p += 32) {
      bulk_load_1d(xs + static_cast<size_t>(p) * row_bytes, xb + static_cast<size_t>(p) * a.x_ld[slab], row_bytes, bar);
      bulk_load_1d(ds + static_cast<size_t>(p) * row_bytes, dyb + static_cast<size_t>(p) * a.dy_ld, row_bytes, bar);
    }
  } else if (t >= 32 && t < 32 + ng) {
    gn_group_stats(a, b, slab, t - 32, s_mean[t - 32], s_rstd[t - 32]);
  }
  __syncthreads();
  float gm[8], be[8], mu[8], rs[8];
  if (active) {
    load8f(a.gamma + slab * Cs + vc * 8, gm);
    load8f(a.beta + slab * Cs + vc * 8, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (vc * 8 + j) / cpg;
      mu[j] = s_mean[g];
      rs[j] = s_rstd[g];
    }
  }
  mbar_wait(bar, 0);
  // ---- pass 1: per-channel sums of dz and dz * xhat over this CTA's rows ----
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.f;
  if (active) {
    for (int p = rl; p < rows; p += RL) {
      float x[8], dy[8];
      uint4* dp = reinterpret_cast<uint4*>(ds + static_cast<size_t>(p) * row_bytes + vc * 16);
      unpack8(*reinterpret_cast<const uint4*>(xs + static_cast<size_t>(p) * row_bytes + vc * 16), x);
      unpack8(*dp, dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mu[j]) * rs[j];
        float dz = dy[j];
        if (a.silu) dz *= silu_grad_f(fmaf(xh, gm[j], be[j]));
        dy[j] = dz;
        sa[j] += dz;
        sb[j] = fmaf(dz, xh, sb[j]);
      }
      if (a.silu) *dp = pack8(dy);  // pass 2 reads dz (bf16, like dy itself) instead of paying the SiLU gradient twice
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_red[(static_cast<size_t>(rl) * Cs + vc * 8 + j) * 2] = sa[j];
      s_red[(static_cast<size_t>(rl) * Cs + vc * 8 + j) * 2 + 1] = sb[j];
    }
  }
  __syncthreads();
  // fixed-order fold over the row lanes -> this rank's slot of s_sum in EVERY CTA of the cluster
  for (int i = t; i < 2 * Cs; i += GNR_T) {
    float acc = 0.f;
    for (int r = 0; r < RL; ++r) acc += s_red[static_cast<size_t>(r) * Cs * 2 + i];
    float* slot = s_sum + static_cast<size_t>(rank) * Cs * 2 + i;
    *slot = acc;
    if (CL > 1) {
#pragma unroll
      for (int pr = 0; pr < CL; ++pr)
        if (pr != static_cast<int>(rank)) st_cluster_f32(mapa_shared(smem_u32(slot), pr), acc);
    }
  }
  if (CL > 1) cluster_sync_all();
  else __syncthreads();
  // totals of the sample (fixed rank order: every CTA of the cluster forms the same bits) -> s_red[0 .. 2 Cs)
  for (int i = t; i < 2 * Cs; i += GNR_T) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) acc += s_sum[static_cast<size_t>(r) * Cs * 2 + i];
    s_red[i] = acc;
    // parameter gradients: dbeta_c += sum dz, dgamma_c += sum dz * xhat (this sample's share, added once per cluster)
    if (rank == 0) atomicAdd(((i & 1) ? a.dgamma : a.dbeta) + slab * Cs + (i >> 1), acc);
  }
  __syncthreads();
  if (t < ng) {
    float S1 = 0.f, S2 = 0.f;
    for (int c = 0; c < cpg; ++c) {
      const float gmm = __ldg(a.gamma + slab * Cs + t * cpg + c);
      S1 = fmaf(gmm, s_red[(t * cpg + c) * 2], S1);
      S2 = fmaf(gmm, s_red[(t * cpg + c) * 2 + 1], S2);
    }
    const float inv_n = 1.0f / static_cast<float>(cpg * a.HW);
    s_c1[t] = S1 * inv_n;
    s_c2[t] = S2 * inv_n;
  }
  __syncthreads();
  if (active) {
    float c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (vc * 8 + j) / cpg;
      c1[j] = s_c1[g];
      c2[j] = s_c2[g];
    }
    // ---- pass 2: dx = rstd (dz gamma - c1 - xhat c2) (+ add) (+= existing) ----
    bf16_t* dxb = a.dx[slab] + row0 * a.dx_ld[slab];
    const bf16_t* addb = a.add[slab] ? a.add[slab] + row0 * a.add_ld[slab] : nullptr;
    const bool acc = a.accumulate[slab] != 0;
    for (int p = rl; p < rows; p += RL) {
      float x[8], dy[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(xs + static_cast<size_t>(p) * row_bytes + vc * 16), x);
      unpack8(*reinterpret_cast<const uint4*>(ds + static_cast<size_t>(p) * row_bytes + vc * 16), dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[j] - mu[j]) * rs[j];
        o[j] = rs[j] * (dy[j] * gm[j] - c1[j] - xh * c2[j]);  // dy holds dz since pass 1
      }
      if (addb) {
        float tt[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(addb + static_cast<size_t>(p) * a.add_ld[slab]) + vc), tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += tt[j];
      }
      uint4* dst = reinterpret_cast<uint4*>(dxb + static_cast<size_t>(p) * a.dx_ld[slab]) + vc;
      if (acc) {
        float tt[8];
        unpack8(*dst, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += tt[j];
      }
      *dst = pack8(o);
    }
  }
  if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still write its s_sum slot (it cannot: all writes precede
                                   // the first barrier; this keeps the exit order simple for the DSMEM rules)
}
__global__ void __launch_bounds__(GNR_T, 1) gn_bwd_rows1_kernel(const GroupNormBwdArgs a) { gn_bwd_rows_body<1>(a); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GNR_T, 1) gn_bwd_rows2_kernel(const GroupNormBwdArgs a) { gn_bwd_rows_body<2>(a); }

cudaError_t groupnorm_bwd_launch(const GroupNormBwdArgs& a, int B, int nslab, cudaStream_t s) {
  const int nv = a.Cs / 8;
  if (a.Cs % 8 || a.Cs % a.cpg || a.cpg % a.pcpg || a.Cs / a.cpg > 128 || a.Cs > 1024 || nslab < 1 || nslab > 2 || !a.ws)
    return cudaErrorInvalidValue;
  {
    static int fused = -1;  // env WD_GN_BWD_FUSED=0: the two-kernel path
    if (fused < 0) {
      const char* e = getenv("WD_GN_BWD_FUSED");
      fused = e ? (atoi(e) != 0) : 1;
    }
    // row-sliced cluster kernel first (env WD_GN_BWD_ROWS=0 switches it off)
    static int rows_on = -1;
    if (rows_on < 0) {
      const char* e = getenv("WD_GN_BWD_ROWS");
      rows_on = e ? (atoi(e) != 0) : 1;
    }
    {
      const int nvr = a.Cs / 8;
      bool okr = rows_on && nvr >= 1 && nvr <= GNR_T && a.Cs / a.cpg <= 128 && a.dy_ld % 8 == 0 && a.cpg == a.pcpg * (a.cpg / a.pcpg);
      for (int i = 0; i < nslab; ++i)
        okr = okr && a.x_ld[i] % 8 == 0 && a.dx_ld[i] % 8 == 0 && (!a.add[i] || a.add_ld[i] % 8 == 0);
      if (okr) {
        int CL = 0;
        if (gnr_layout(a.HW, a.Cs, 1).total <= 226 * 1024) CL = 1;
        else if (a.HW % 2 == 0 && gnr_layout(a.HW / 2, a.Cs, 2).total <= 226 * 1024) CL = 2;
        // small batches leave most SMs without a CTA (one 210 KB CTA per SM, B * nslab * CL of them): the channel-sliced kernel's
        // finer grid wins there (batch 28: 5.92 vs 6.07 ms per training step; batch 224: 14.31 vs 14.06)
        if (CL && B * nslab * CL < 148) CL = 0;
        if (CL) {
          static std::once_flag once;
          static cudaError_t attr_err = cudaSuccess;
          std::call_once(once, [] {
            attr_err = cudaFuncSetAttribute(gn_bwd_rows1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
            if (attr_err == cudaSuccess)
              attr_err = cudaFuncSetAttribute(gn_bwd_rows2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
          });
          if (attr_err != cudaSuccess) return attr_err;
          const size_t smem = gnr_layout(a.HW / CL, a.Cs, CL).total;
          if (CL == 1) gn_bwd_rows1_kernel<<<dim3(B, nslab), GNR_T, smem, s>>>(a);
          else gn_bwd_rows2_kernel<<<dim3(B * 2, nslab), GNR_T, smem, s>>>(a);
          return cudaGetLastError();
        }
      }
    }
    bool ok = fused && a.Cs % GNF_C == 0 && GNF_C % a.cpg == 0 && GNF_C / a.cpg <= 4 && a.HW <= GNF_ROWS * GNF_MAXIT && a.dy_ld % 8 == 0;
    for (int i = 0; i < nslab; ++i)
      ok = ok && a.x_ld[i] % 8 == 0 && a.dx_ld[i] % 8 == 0 && (!a.add[i] || a.add_ld[i] % 8 == 0);
    if (ok) {
      gn_bwd_fused_kernel<<<dim3(B, nslab, a.Cs / GNF_C), 256, 0, s>>>(a);
      return cudaGetLastError();
    }
  }
  int R = GNB_R;
  while (R > 1 && nv * R > 640) R >>= 1;
  if (nv * R > 640 || nv * R < a.Cs / a.cpg) return cudaErrorInvalidValue;
  const size_t smem = static_cast<size_t>(R) * a.Cs * 2 * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  const int per = 4 * R;
  const int nchunk = (a.HW >= 2 * per && a.HW % per == 0) ? a.HW / per : 1;
  int rchunks = nchunk < GNB_MAX_RCHUNKS ? nchunk : GNB_MAX_RCHUNKS;  // ws holds [B][rchunks][C][2] floats
  while (a.HW % rchunks) --rchunks;
  gn_bwd_reduce_kernel<<<dim3(B, nslab, rchunks), nv * R, smem, s>>>(a, rchunks);
  gn_bwd_apply_kernel<<<dim3(B, nslab, nchunk), nv * R, 0, s>>>(a, nchunk, rchunks);
  return cudaGetLastError();
}

// =====================================================================================================
// LayerNorm backward: one warp per token (looping), per-lane register accumulators for dgamma / dbeta
// =====================================================================================================
constexpr int LNB_STAGES = 4;
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16_t* __restrict__ x, const bf16_t* __restrict__ dy,
                                                            const float* __restrict__ gamma, const bf16_t* __restrict__ add,
                                                            bf16_t* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int M, int C, float eps) {
  // [2][C] fp32 column sums, then per warp LNB_STAGES x (x row, dy row) bf16 staging buffers filled by cp.async.bulk (one token per
  // stage, three tokens in flight per warp: with plain loads a warp had ONE token = 1.3 KB in flight and the 12 launches of a
  // batch-224 backward pass ran at 1.45 TB/s), then the stages' mbarriers
  extern __shared__ __align__(128) float lnb_smem[];
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int warps_total = gridDim.x * nwarps;
  const int nv = C >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(C) * 2;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(lnb_smem + 2 * C);
  uint8_t* wbuf = stage_base + static_cast<size_t>(warp_in_block) * LNB_STAGES * 2 * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + static_cast<size_t>(nwarps) * LNB_STAGES * 2 * row_bytes);
  uint64_t* wbar = bars + warp_in_block * LNB_STAGES;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) lnb_smem[i] = 0.f;
  if (threadIdx.x < nwarps * LNB_STAGES) mbar_init(&bars[threadIdx.x], 1);
  fence_barrier_init();
  __syncthreads();
  const int tok0 = blockIdx.x * nwarps + warp_in_block;
  const int ntok = tok0 < M ? (M - tok0 + warps_total - 1) / warps_total : 0;  // tokens of this warp: tok0 + k * warps_total
  auto issue = [&](int k) {  // lane 0 only
    const int st = k % LNB_STAGES;
    const size_t tokk = static_cast<size_t>(tok0) + static_cast<size_t>(k) * warps_total;
    mbar_arrive_expect_tx(&wbar[st], 2 * row_bytes);
    bulk_load_1d(wbuf + static_cast<size_t>(st) * 2 * row_bytes, x + tokk * C, row_bytes, &wbar[st]);
    bulk_load_1d(wbuf + static_cast<size_t>(st) * 2 * row_bytes + row_bytes, dy + tokk * C, row_bytes, &wbar[st]);
  };
  if (lane == 0)
    for (int k = 0; k < LNB_STAGES - 1 && k < ntok; ++k) issue(k);
  float gm[MAXV][8], ag[MAXV][8], ab[MAXV][8];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nv) load8f(gamma + vi * 8, gm[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = 0.f;
  }
  for (int k = 0; k < ntok; ++k) {
    const int tok = tok0 + k * warps_total;
    const int st = k % LNB_STAGES;
    if (lane == 0 && k + LNB_STAGES - 1 < ntok) issue(k + LNB_STAGES - 1);  // its stage was drained in iteration k - 1
    mbar_wait(&wbar[st], (k / LNB_STAGES) & 1);
    const uint4* xr = reinterpret_cast<const uint4*>(wbuf + static_cast<size_t>(st) * 2 * row_bytes);
    const uint4* dr = reinterpret_cast<const uint4*>(wbuf + static_cast<size_t>(st) * 2 * row_bytes + row_bytes);
    float f[MAXV][8], d[MAXV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nv) {
        unpack8(xr[vi], f[i]);
        unpack8(dr[vi], d[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[i][j];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(C);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (lane + 32 * i < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = f[i][j] - mean;
          var += t * t;
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / static_cast<float>(C) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (lane + 32 * i < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (f[i][j] - mean) * rstd;
          f[i][j] = xh;
          ab[i][j] += d[i][j];
          ag[i][j] = fmaf(d[i][j], xh, ag[i][j]);
          const float g = d[i][j] * gm[i][j];
          d[i][j] = g;
          s1 += g;
          s2 = fmaf(g, xh, s2);
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= static_cast<float>(C);
    s2 /= static_cast<float>(C);
    uint4* orow = reinterpret_cast<uint4*>(dx + static_cast<size_t>(tok) * C);
    const uint4* arow = add ? reinterpret_cast<const uint4*>(add + static_cast<size_t>(tok) * C) : nullptr;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nv) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = rstd * (d[i][j] - s1 - f[i][j] * s2);
        if (arow) {
          float t[8];
          unpack8(__ldg(arow + vi), t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] += t[j];
        }
        orow[vi] = pack8(o8);
      }
    }
    __syncwarp();  // every lane has read the stage (f / d are in registers) before lane 0 refills it
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&lnb_smem[vi * 8 + j], ag[i][j]);
        atomicAdd(&lnb_smem[C + vi * 8 + j], ab[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, lnb_smem[i]);
    atomicAdd(dbeta + i, lnb_smem[C + i]);
  }
}

cudaError_t layernorm_bwd_launch(const bf16_t* x, const bf16_t* dy, const float* gamma, const bf16_t* add, bf16_t* dx,
                                 float* dgamma, float* dbeta, int M, int C, float eps, cudaStream_t s) {
  if (C % 8 || C > 8 * 32 * 4) return cudaErrorInvalidValue;
  int blocks = (M + 31) / 32;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  if (C % 8) return cudaErrorInvalidValue;  // bulk copies of whole rows: 16-byte multiples
  const size_t smem = static_cast<size_t>(2) * C * sizeof(float) + static_cast<size_t>(8) * LNB_STAGES * 2 * C * 2 + 8 * LNB_STAGES * 8;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(layernorm_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(layernorm_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  if (C <= 8 * 32 * 2)
    layernorm_bwd_kernel<2><<<blocks, 256, smem, s>>>(x, dy, gamma, add, dx, dgamma, dbeta, M, C, eps);
  else
    layernorm_bwd_kernel<4><<<blocks, 256, smem, s>>>(x, dy, gamma, add, dx, dgamma, dbeta, M, C, eps);
  return cudaGetLastError();
}

// =====================================================================================================
// GEGLU forward / backward (exact erf GELU, unet.py:127-129)
// =====================================================================================================
__global__ void geglu_fwd_kernel(const bf16_t* __restrict__ p, bf16_t* __restrict__ out, size_t total, int hv) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t m = idx / hv;
  const int v = idx % hv;
  const uint4* row = reinterpret_cast<const uint4*>(p) + m * 2 * hv;
  float a[8], g[8], o[8];
  unpack8(__ldg(row + v), a);
  unpack8(__ldg(row + hv + v), g);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = a[j] * gelu_erf_f(g[j]);
  reinterpret_cast<uint4*>(out)[idx] = pack8(o);
}
__global__ void geglu_bwd_kernel(const bf16_t* __restrict__ p, const bf16_t* __restrict__ dout, bf16_t* __restrict__ dp,
                                 size_t total, int hv) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t m = idx / hv;
  const int v = idx % hv;
  const uint4* row = reinterpret_cast<const uint4*>(p) + m * 2 * hv;
  float a[8], g[8], d[8], da[8], dg[8];
  unpack8(__ldg(row + v), a);
  unpack8(__ldg(row + hv + v), g);
  unpack8(__ldg(reinterpret_cast<const uint4*>(dout) + idx), d);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float cdf = 0.5f * (1.0f + erff(g[j] * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * g[j] * g[j]);
    da[j] = d[j] * g[j] * cdf;
    dg[j] = d[j] * a[j] * (cdf + g[j] * pdf);
  }
  uint4* orow = reinterpret_cast<uint4*>(dp) + m * 2 * hv;
  orow[v] = pack8(da);
  orow[hv + v] = pack8(dg);
}
cudaError_t geglu_fwd_launch(const bf16_t* p, bf16_t* out, int M, int H, cudaStream_t s) {
  if (H % 8) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(M) * (H / 8);
  geglu_fwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(p, out, total, H / 8);
  return cudaGetLastError();
}
cudaError_t geglu_bwd_launch(const bf16_t* p, const bf16_t* dout, bf16_t* dp, int M, int H, cudaStream_t s) {
  if (H % 8) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(M) * (H / 8);
  geglu_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(p, dout, dp, total, H / 8);
  return cudaGetLastError();
}

// =====================================================================================================
// SiLU forward / backward
// =====================================================================================================
__global__ void silu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, size_t nv) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nv) return;
  float f[8];
  unpack8(__ldg(x + idx), f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
  y[idx] = pack8(f);
}
__global__ void silu_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx, size_t nv) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nv) return;
  float f[8], d[8];
  unpack8(__ldg(x + idx), f);
  unpack8(__ldg(dy + idx), d);
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j] *= silu_grad_f(f[j]);
  dx[idx] = pack8(d);
}
cudaError_t silu_fwd_launch(const bf16_t* x, bf16_t* y, size_t n, cudaStream_t s) {
  if (n % 8) return cudaErrorInvalidValue;
  silu_fwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(x),
                                                                              reinterpret_cast<uint4*>(y), n / 8);
  return cudaGetLastError();
}
cudaError_t silu_bwd_launch(const bf16_t* x, const bf16_t* dy, bf16_t* dx, size_t n, cudaStream_t s) {
  if (n % 8) return cudaErrorInvalidValue;
  silu_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), n / 8);
  return cudaGetLastError();
}

// =====================================================================================================
// Short-context cross-attention backward: one CTA per (head, sample).  Phase 1: one thread per query recomputes its
// softmax row, forms dS and writes dq; phase 2: threads own (key, channel) outputs and reduce dK / dV over the queries.
// =====================================================================================================
constexpr int ASB_DH = 80;
constexpr int ASB_ROW = 88;  // padded bf16 row (176 B): conflict-free 16-byte row reads
constexpr int ASB_T = 128;   // queries per chunk = threads per CTA (71 KB of shared memory: 3 CTAs / SM)
constexpr int ASB_OWN = 16 * ASB_DH / ASB_T;  // (key, channel) outputs owned by a thread in phase 2

__global__ void __launch_bounds__(ASB_T) attn_small_bwd_kernel(const AttnSmallBwdArgs a) {
  extern __shared__ __align__(16) uint8_t asb_smem[];
  bf16_t* sQ = reinterpret_cast<bf16_t*>(asb_smem);                 // [ASB_T][88]
  bf16_t* sDO = sQ + ASB_T * ASB_ROW;                              // [ASB_T][88]
  float* sP = reinterpret_cast<float*>(sDO + ASB_T * ASB_ROW);     // [ASB_T][16]
  float* sDS = sP + ASB_T * 16;                                    // [ASB_T][16]
  float* sK = sDS + ASB_T * 16;                                    // [16][80]
  float* sV = sK + 16 * ASB_DH;                                    // [16][80]
  const int h = blockIdx.x, b = blockIdx.y;
  const int L = a.L;
  const int t = threadIdx.x;
  for (int i = t; i < 16 * ASB_DH; i += ASB_T) {
    const int l = i / ASB_DH, d = i % ASB_DH;
    float kv = 0.f, vv = 0.f;
    if (l < L) {
      kv = __bfloat162float(a.k[(static_cast<size_t>(b) * L + l) * a.kv_ld + h * ASB_DH + d]);
      vv = __bfloat162float(a.v[(static_cast<size_t>(b) * L + l) * a.kv_ld + h * ASB_DH + d]);
    }
    sK[i] = kv;
    sV[i] = vv;
  }
  // phase-2 ownership: output o = t + ASB_T * i  ->  (l, d) = (o / 80, o % 80), o < L * 80
  float accK[ASB_OWN], accV[ASB_OWN];
#pragma unroll
  for (int i = 0; i < ASB_OWN; ++i) accK[i] = accV[i] = 0.f;

  for (int q0 = 0; q0 < a.Sq; q0 += ASB_T) {
    const int nq = min(ASB_T, a.Sq - q0);
    __syncthreads();  // previous chunk's phase 2 is done with the staging buffers (and sK/sV are written)
    // cooperative, coalesced staging of the Q and dO rows of this (sample, head)
    for (int i = t; i < nq * (ASB_DH / 8); i += ASB_T) {
      const int r = i / (ASB_DH / 8), vcol = i % (ASB_DH / 8);
      const size_t tok = static_cast<size_t>(b) * a.Sq + q0 + r;
      *reinterpret_cast<uint4*>(sQ + r * ASB_ROW + vcol * 8) =
          __ldg(reinterpret_cast<const uint4*>(a.q + tok * a.q_ld + h * ASB_DH) + vcol);
      *reinterpret_cast<uint4*>(sDO + r * ASB_ROW + vcol * 8) =
          __ldg(reinterpret_cast<const uint4*>(a.dout + tok * a.do_ld + h * ASB_DH) + vcol);
    }
    __syncthreads();
    if (t < nq) {
      float sc[16], dp[16];
#pragma unroll
      for (int l = 0; l < 16; ++l) sc[l] = dp[l] = 0.f;
#pragma unroll 2
      for (int vcol = 0; vcol < ASB_DH / 8; ++vcol) {
        float qf[8], df[8];
        unpack8(*reinterpret_cast<const uint4*>(sQ + t * ASB_ROW + vcol * 8), qf);
        unpack8(*reinterpret_cast<const uint4*>(sDO + t * ASB_ROW + vcol * 8), df);
#pragma unroll
        for (int l = 0; l < 16; ++l) {
          if (l < L) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              sc[l] = fmaf(qf[j], sK[l * ASB_DH + vcol * 8 + j], sc[l]);
              dp[l] = fmaf(df[j], sV[l * ASB_DH + vcol * 8 + j], dp[l]);
            }
          }
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int l = 0; l < 16; ++l)
        if (l < L) {
          sc[l] *= a.scale;
          mx = fmaxf(mx, sc[l]);
        }
      float den = 0.f;
#pragma unroll
      for (int l = 0; l < 16; ++l)
        if (l < L) {
          sc[l] = __expf(sc[l] - mx);
          den += sc[l];
        }
      const float inv = 1.0f / den;
      float delta = 0.f;
#pragma unroll
      for (int l = 0; l < 16; ++l)
        if (l < L) {
          sc[l] *= inv;
          delta = fmaf(sc[l], dp[l], delta);
        }
#pragma unroll
      for (int l = 0; l < 16; ++l) {
        const float p = (l < L) ? sc[l] : 0.f;
        const float ds = (l < L) ? p * (dp[l] - delta) * a.scale : 0.f;  // scale folded in: dq = ds K, dk = ds^T q
        sP[t * 16 + l] = p;
        sDS[t * 16 + l] = ds;
        dp[l] = ds;
      }
      // dq = dS K
      bf16_t* dqr = a.dq + (static_cast<size_t>(b) * a.Sq + q0 + t) * a.dq_ld + h * ASB_DH;
#pragma unroll 2
      for (int vcol = 0; vcol < ASB_DH / 8; ++vcol) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
        for (int l = 0; l < 16; ++l)
          if (l < L) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(dp[l], sK[l * ASB_DH + vcol * 8 + j], o[j]);
          }
        reinterpret_cast<uint4*>(dqr)[vcol] = pack8(o);
      }
    }
    __syncthreads();
    // phase 2
#pragma unroll
    for (int i = 0; i < ASB_OWN; ++i) {
      const int o = t + ASB_T * i;
      if (o < L * ASB_DH) {
        const int l = o / ASB_DH, d = o % ASB_DH;
        float ak0 = 0.f, av0 = 0.f, ak1 = 0.f, av1 = 0.f;
        int r = 0;
        for (; r + 1 < nq; r += 2) {  // two independent accumulation chains
          ak0 = fmaf(sDS[r * 16 + l], __bfloat162float(sQ[r * ASB_ROW + d]), ak0);
          av0 = fmaf(sP[r * 16 + l], __bfloat162float(sDO[r * ASB_ROW + d]), av0);
          ak1 = fmaf(sDS[(r + 1) * 16 + l], __bfloat162float(sQ[(r + 1) * ASB_ROW + d]), ak1);
          av1 = fmaf(sP[(r + 1) * 16 + l], __bfloat162float(sDO[(r + 1) * ASB_ROW + d]), av1);
        }
        if (r < nq) {
          ak0 = fmaf(sDS[r * 16 + l], __bfloat162float(sQ[r * ASB_ROW + d]), ak0);
          av0 = fmaf(sP[r * 16 + l], __bfloat162float(sDO[r * ASB_ROW + d]), av0);
        }
        accK[i] += ak0 + ak1;
        accV[i] += av0 + av1;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ASB_OWN; ++i) {
    const int o = t + ASB_T * i;
    if (o < L * ASB_DH) {
      const int l = o / ASB_DH, d = o % ASB_DH;
      const size_t off = (static_cast<size_t>(b) * L + l) * a.dkv_ld + h * ASB_DH + d;
      a.dk[off] = __float2bfloat16(accK[i]);
      a.dv[off] = __float2bfloat16(accV[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core version (the one the trainer launches).  The kernel above reads K / V with broadcast shared-memory loads in its
// per-query phase and walks all queries per (key, channel) output in its second phase: 224 us per launch at batch 224, 13 % of
// the training step.  Here one CTA = (head, sample), four warps x 16 query rows per 64-row chunk, everything on
// mma.sync.m16n8k16 (bf16, fp32 accumulate):
//   S = Q K^T, dP = dO V^T (A fragments by ldmatrix from the staged rows), softmax / delta / dS in the accumulator layout,
//   dq = dS K (dS as A fragment straight from registers, K by ldmatrix.trans),
//   dK += dS^T Q, dV += P^T dO: P and dS are written as bf16 into a per-warp 16 x 16 tile and read back TRANSPOSED with
//   ldmatrix.trans as A fragments; Q / dO rows are the B operands (ldmatrix.trans).  dK / dV stay in registers over all chunks
//   and are reduced across the four warps in a fixed order at the end (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int ABM_ROW = 88;  // bf16 row pitch of the staged Q / dO / K / V rows (176 B: conflict-free ldmatrix)
constexpr int ABM_TP = 24;   // bf16 row pitch of the per-warp P / dS tiles (48 B)
constexpr int ABM_Q = 64;    // query rows per chunk
WD_DEVINL void abm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
WD_DEVINL void abm_ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
WD_DEVINL void abm_ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
WD_DEVINL void abm_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}
}  // namespace

__global__ void __launch_bounds__(128) attn_small_bwd_mma_kernel(const AttnSmallBwdArgs a) {
  extern __shared__ __align__(16) uint8_t abm_smem[];
  bf16_t* sQ = reinterpret_cast<bf16_t*>(abm_smem);      // [64][88]
  bf16_t* sDO = sQ + ABM_Q * ABM_ROW;                    // [64][88]
  bf16_t* sK = sDO + ABM_Q * ABM_ROW;                    // [16][88]
  bf16_t* sV = sK + 16 * ABM_ROW;                        // [16][88]
  bf16_t* sT = sV + 16 * ABM_ROW;                        // [4 warps][P | dS][16][24]
  float* red = reinterpret_cast<float*>(sT + 4 * 2 * 16 * ABM_TP);  // [dK | dV][16][80]
  const int h = blockIdx.x, b = blockIdx.y;
  const int L = a.L;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int g = lane >> 2, tq = lane & 3;
  // K / V rows of (sample, head); rows >= L are zero
  for (int i = t; i < 2 * 16 * 10; i += 128) {
    const int sel = i / 160, r = (i % 160) / 10, vc = i % 10;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (r < L) u = __ldg(reinterpret_cast<const uint4*>((sel ? a.v : a.k) + (static_cast<size_t>(b) * L + r) * a.kv_ld + h * ASB_DH) + vc);
    *reinterpret_cast<uint4*>((sel ? sV : sK) + r * ABM_ROW + vc * 8) = u;
  }
  float accK[10][4], accV[10][4];
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accK[i][j] = accV[i][j] = 0.f;
  bf16_t* tP = sT + warp * (2 * 16 * ABM_TP);
  bf16_t* tS = tP + 16 * ABM_TP;
  const int r0 = warp * 16;

  for (int q0 = 0; q0 < a.Sq; q0 += ABM_Q) {
    __syncthreads();  // the previous chunk's B-operand reads of sQ / sDO are done (and sK / sV are written)
    for (int i = t; i < ABM_Q * 10; i += 128) {
      const int r = i / 10, vc = i % 10;
      uint4 uq = make_uint4(0u, 0u, 0u, 0u), ud = uq;
      if (q0 + r < a.Sq) {
        const size_t tok = static_cast<size_t>(b) * a.Sq + q0 + r;
        uq = __ldg(reinterpret_cast<const uint4*>(a.q + tok * a.q_ld + h * ASB_DH) + vc);
        ud = __ldg(reinterpret_cast<const uint4*>(a.dout + tok * a.do_ld + h * ASB_DH) + vc);
      }
      *reinterpret_cast<uint4*>(sQ + r * ABM_ROW + vc * 8) = uq;
      *reinterpret_cast<uint4*>(sDO + r * ABM_ROW + vc * 8) = ud;
    }
    __syncthreads();
    // ---- S = Q K^T, dP = dO V^T (16 rows x 16 keys per warp) ----
    float sc[2][4], dp[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[nt][j] = dp[nt][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      uint32_t qf[4], df[4];
      abm_ldsm_x4(qf, sQ + (r0 + (lane & 15)) * ABM_ROW + ks * 16 + (lane >> 4) * 8);
      abm_ldsm_x4(df, sDO + (r0 + (lane & 15)) * ABM_ROW + ks * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const bf16_t* kr = sK + (nt * 8 + g) * ABM_ROW + ks * 16 + 2 * tq;
        const bf16_t* vr = sV + (nt * 8 + g) * ABM_ROW + ks * 16 + 2 * tq;
        abm_mma(sc[nt], qf, *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 8));
        abm_mma(dp[nt], df, *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
      }
    }
    // ---- softmax over the keys, delta = sum_l P dP, dS = P (dP - delta) scale  (rows g and g + 8; a row lives in one quad) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int key = nt * 8 + 2 * tq;
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[nt][j] *= a.scale;
      if (key >= L) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
      if (key + 1 >= L) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      sc[nt][0] = __expf(sc[nt][0] - mx0);
      sc[nt][1] = __expf(sc[nt][1] - mx0);
      sc[nt][2] = __expf(sc[nt][2] - mx1);
      sc[nt][3] = __expf(sc[nt][3] - mx1);
      l0 += sc[nt][0] + sc[nt][1];
      l1 += sc[nt][2] + sc[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      sc[nt][0] *= i0; sc[nt][1] *= i0; sc[nt][2] *= i1; sc[nt][3] *= i1;
      d0 = fmaf(sc[nt][0], dp[nt][0], fmaf(sc[nt][1], dp[nt][1], d0));
      d1 = fmaf(sc[nt][2], dp[nt][2], fmaf(sc[nt][3], dp[nt][3], d1));
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    uint32_t dsf[4];  // dS as an A fragment (m = query, k = key)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const float s0 = sc[nt][0] * (dp[nt][0] - d0) * a.scale, s1 = sc[nt][1] * (dp[nt][1] - d0) * a.scale;
      const float s2 = sc[nt][2] * (dp[nt][2] - d1) * a.scale, s3 = sc[nt][3] * (dp[nt][3] - d1) * a.scale;
      dsf[nt * 2] = pack_bf16x2(s0, s1);
      dsf[nt * 2 + 1] = pack_bf16x2(s2, s3);
      // P and dS tiles [query][key] for the transposed reads below
      const int key = nt * 8 + 2 * tq;
      *reinterpret_cast<uint32_t*>(tP + g * ABM_TP + key) = pack_bf16x2(sc[nt][0], sc[nt][1]);
      *reinterpret_cast<uint32_t*>(tP + (g + 8) * ABM_TP + key) = pack_bf16x2(sc[nt][2], sc[nt][3]);
      *reinterpret_cast<uint32_t*>(tS + g * ABM_TP + key) = dsf[nt * 2];
      *reinterpret_cast<uint32_t*>(tS + (g + 8) * ABM_TP + key) = dsf[nt * 2 + 1];
    }
    // ---- dq = dS K ----
    {
      const int row_a = q0 + r0 + g, row_b = row_a + 8;
      bf16_t* dqa = a.dq + (static_cast<size_t>(b) * a.Sq + row_a) * a.dq_ld + h * ASB_DH + 2 * tq;
      bf16_t* dqb = a.dq + (static_cast<size_t>(b) * a.Sq + row_b) * a.dq_ld + h * ASB_DH + 2 * tq;
#pragma unroll
      for (int dt = 0; dt < 10; ++dt) {
        uint32_t b0, b1;
        abm_ldsm_x2_trans(b0, b1, sK + (lane & 15) * ABM_ROW + dt * 8);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        abm_mma(o, dsf, b0, b1);
        if (row_a < a.Sq) *reinterpret_cast<uint32_t*>(dqa + dt * 8) = pack_bf16x2(o[0], o[1]);
        if (row_b < a.Sq) *reinterpret_cast<uint32_t*>(dqb + dt * 8) = pack_bf16x2(o[2], o[3]);
      }
    }
    __syncwarp();
    // ---- dK += dS^T Q, dV += P^T dO  (A = transposed tile: m = key, k = query; B = the warp's Q / dO rows) ----
    {
      uint32_t stf[4], ptf[4];
      const int mid = lane >> 3, rr = lane & 7;
      abm_ldsm_x4_trans(stf, tS + ((mid >> 1) * 8 + rr) * ABM_TP + (mid & 1) * 8);
      abm_ldsm_x4_trans(ptf, tP + ((mid >> 1) * 8 + rr) * ABM_TP + (mid & 1) * 8);
#pragma unroll
      for (int dt = 0; dt < 10; ++dt) {
        uint32_t b0, b1;
        abm_ldsm_x2_trans(b0, b1, sQ + (r0 + (lane & 15)) * ABM_ROW + dt * 8);
        abm_mma(accK[dt], stf, b0, b1);
        abm_ldsm_x2_trans(b0, b1, sDO + (r0 + (lane & 15)) * ABM_ROW + dt * 8);
        abm_mma(accV[dt], ptf, b0, b1);
      }
    }
    __syncwarp();  // the tiles are rewritten in the next chunk
  }
  // ---- reduce dK / dV over the four warps in a fixed order; accumulator rows = keys g, g + 8, columns 8 dt + 2 tq (+1) ----
  for (int w = 0; w < 4; ++w) {
    __syncthreads();
    if (warp == w) {
#pragma unroll
      for (int dt = 0; dt < 10; ++dt) {
        float* ka = red + g * ASB_DH + dt * 8 + 2 * tq;
        float* kb = red + (g + 8) * ASB_DH + dt * 8 + 2 * tq;
        float* va = ka + 16 * ASB_DH;
        float* vb = kb + 16 * ASB_DH;
        if (w == 0) {
          ka[0] = accK[dt][0]; ka[1] = accK[dt][1]; kb[0] = accK[dt][2]; kb[1] = accK[dt][3];
          va[0] = accV[dt][0]; va[1] = accV[dt][1]; vb[0] = accV[dt][2]; vb[1] = accV[dt][3];
        } else {
          ka[0] += accK[dt][0]; ka[1] += accK[dt][1]; kb[0] += accK[dt][2]; kb[1] += accK[dt][3];
          va[0] += accV[dt][0]; va[1] += accV[dt][1]; vb[0] += accV[dt][2]; vb[1] += accV[dt][3];
        }
      }
    }
  }
  __syncthreads();
  for (int o = t; o < L * ASB_DH; o += 128) {
    const int l = o / ASB_DH, d = o % ASB_DH;
    const size_t off = (static_cast<size_t>(b) * L + l) * a.dkv_ld + h * ASB_DH + d;
    a.dk[off] = __float2bfloat16(red[l * ASB_DH + d]);
    a.dv[off] = __float2bfloat16(red[16 * ASB_DH + l * ASB_DH + d]);
  }
}

cudaError_t attn_small_bwd_launch(const AttnSmallBwdArgs& a, int B, cudaStream_t s) {
  if (a.L < 1 || a.L > 16 || a.Sq < 1) return cudaErrorInvalidValue;
  {
    static int use_mma = -1;  // env WD_ATTN_BWD_MMA=0: the SIMT kernel
    if (use_mma < 0) {
      const char* e = getenv("WD_ATTN_BWD_MMA");
      use_mma = e ? (atoi(e) != 0) : 1;
    }
    if (use_mma && a.q_ld % 8 == 0 && a.kv_ld % 8 == 0 && a.do_ld % 8 == 0 && a.dq_ld % 2 == 0) {
      const size_t sm = static_cast<size_t>(2 * ABM_Q + 32) * ABM_ROW * 2 + static_cast<size_t>(4) * 2 * 16 * ABM_TP * 2 +
                        static_cast<size_t>(2) * 16 * ASB_DH * 4;
      attn_small_bwd_mma_kernel<<<dim3(a.heads, B), 128, sm, s>>>(a);
      return cudaGetLastError();
    }
  }
  const size_t smem = static_cast<size_t>(2) * ASB_T * ASB_ROW * 2 + static_cast<size_t>(2) * ASB_T * 16 * 4 +
                      static_cast<size_t>(2) * 16 * ASB_DH * 4;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [smem] {
    attr_err = cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  });
  if (attr_err != cudaSuccess) return attr_err;
  attn_small_bwd_kernel<<<dim3(a.heads, B), ASB_T, smem, s>>>(a);
  return cudaGetLastError();
}

// =====================================================================================================
// column sums
// =====================================================================================================
__global__ void __launch_bounds__(512) colsum_kernel(const bf16_t* __restrict__ dy, int ld, int N, int rows_total,
                                                     int rows_per_group, float* __restrict__ total,
                                                     bf16_t* __restrict__ per_group, int pg_ld, int R) {
  extern __shared__ float cs_smem[];  // [R][N]
  const int nv = N >> 3;
  const int col = threadIdx.x % nv, rl = threadIdx.x / nv;
  const int g = blockIdx.x;
  const int r0 = g * rows_per_group;
  const int r1 = min(r0 + rows_per_group, rows_total);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (rl < R) {
    for (int r = r0 + rl; r < r1; r += R) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + static_cast<size_t>(r) * ld) + col), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) cs_smem[rl * N + col * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < R; ++r) t += cs_smem[r * N + c];
    if (per_group) per_group[static_cast<size_t>(g) * pg_ld + c] = __float2bfloat16(t);
    if (total) atomicAdd(total + c, t);
  }
}
cudaError_t colsum_launch(const bf16_t* dy, int ld, int N, int groups, int rows_per_group, float* total, bf16_t* per_group,
                          int pg_ld, cudaStream_t s) {
  if (N % 8 || N / 8 > 512 || groups < 1) return cudaErrorInvalidValue;
  const int nv = N / 8;
  int R = 512 / nv;
  if (R < 1) R = 1;
  while (R > 1 && static_cast<size_t>(R) * N * 4 > 40 * 1024) --R;
  if (static_cast<size_t>(R) * N * 4 > 48 * 1024) return cudaErrorInvalidValue;
  if (R > rows_per_group) R = rows_per_group;
  int threads = nv * R;
  threads = (threads + 31) / 32 * 32;
  colsum_kernel<<<groups, threads, static_cast<size_t>(R) * N * 4, s>>>(dy, ld, N, groups * rows_per_group, rows_per_group,
                                                                        total, per_group, pg_ld, R);
  return cudaGetLastError();
}

// =====================================================================================================
// resampling
// =====================================================================================================
__global__ void upsample2x_bwd_kernel(const uint4* __restrict__ dup, uint4* __restrict__ dx, int B, int H, int W, int nv,
                                      int accumulate) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * H * W * nv;
  if (idx >= total) return;
  const int v = idx % nv;
  size_t p = idx / nv;
  const int x = p % W;
  p /= W;
  const int y = p % H;
  const int b = p / H;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dxx = 0; dxx < 2; ++dxx) {
      float f[8];
      unpack8(__ldg(dup + ((static_cast<size_t>(b) * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dxx) * nv + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += f[j];
    }
  if (accumulate) {
    float f[8];
    unpack8(dx[idx], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += f[j];
  }
  dx[idx] = pack8(o);
}
cudaError_t upsample2x_bwd_launch(const bf16_t* dup, bf16_t* dx, int B, int H, int W, int C, int accumulate, cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(B) * H * W * (C / 8);
  upsample2x_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const uint4*>(dup), reinterpret_cast<uint4*>(dx), B, H, W, C / 8, accumulate);
  return cudaGetLastError();
}

__global__ void dilate2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int nv) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * 4 * H * W * nv;
  if (idx >= total) return;
  const int v = idx % nv;
  size_t p = idx / nv;
  const int ox = p % (2 * W);
  p /= (2 * W);
  const int oy = p % (2 * H);
  const int b = p / (2 * H);
  uint4 r = make_uint4(0, 0, 0, 0);
  if (!(ox & 1) && !(oy & 1)) r = __ldg(x + ((static_cast<size_t>(b) * H + (oy >> 1)) * W + (ox >> 1)) * nv + v);
  out[idx] = r;
}
cudaError_t dilate2x_launch(const bf16_t* x, bf16_t* out, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(B) * 4 * H * W * (C / 8);
  dilate2x_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(x),
                                                                              reinterpret_cast<uint4*>(out), B, H, W, C / 8);
  return cudaGetLastError();
}

// =====================================================================================================
// layout / dtype glue
// =====================================================================================================
__global__ void nchw4_to_tok64_kernel(const float* __restrict__ g, bf16_t* __restrict__ out, int B, int HW) {
  const size_t m = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= static_cast<size_t>(B) * HW) return;
  const size_t b = m / HW, pix = m % HW;
  float f[8];
#pragma unroll
  for (int o = 0; o < 4; ++o) f[o] = __ldg(g + (b * 4 + o) * HW + pix);
#pragma unroll
  for (int o = 4; o < 8; ++o) f[o] = 0.f;
  *reinterpret_cast<uint4*>(out + m * 64) = pack8(f);
}
cudaError_t nchw4_to_tok64_launch(const float* g, bf16_t* out, int B, int HW, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * HW;
  nchw4_to_tok64_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(g, out, B, HW);
  return cudaGetLastError();
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, bf16_t* __restrict__ out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(x[i]);
}
cudaError_t f32_to_bf16_launch(const float* x, bf16_t* out, size_t n, cudaStream_t s) {
  f32_to_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x, out, n);
  return cudaGetLastError();
}

__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ y, size_t nv,
                                int accumulate) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  float o[8];
  unpack8(__ldg(a + i), o);
  if (b) {
    float f[8];
    unpack8(__ldg(b + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += f[j];
  }
  if (accumulate) {
    float f[8];
    unpack8(y[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += f[j];
  }
  y[i] = pack8(o);
}
cudaError_t add_bf16_launch(const bf16_t* a, const bf16_t* b, bf16_t* y, size_t n, int accumulate, cudaStream_t s) {
  if (n % 8) return cudaErrorInvalidValue;
  add_bf16_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(y), n / 8, accumulate);
  return cudaGetLastError();
}

__global__ void scatter_add_rows_kernel(const bf16_t* __restrict__ rows, int ld, const void* __restrict__ idx, int idx_i64,
                                        float* __restrict__ table, int nrows, int D, int table_rows) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(nrows) * D) return;
  const int r = i / D, d = i % D;
  const long long t = idx_i64 ? static_cast<const long long*>(idx)[r] : static_cast<const int*>(idx)[r];
  if (t < 0 || t >= table_rows) __trap();
  atomicAdd(table + t * D + d, __bfloat162float(rows[static_cast<size_t>(r) * ld + d]));
}
cudaError_t scatter_add_rows_launch(const bf16_t* rows, int ld, const void* idx, int idx_i64, float* table, int nrows, int D,
                                    int table_rows, cudaStream_t s) {
  const size_t total = static_cast<size_t>(nrows) * D;
  scatter_add_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(rows, ld, idx, idx_i64, table, nrows, D,
                                                                                      table_rows);
  return cudaGetLastError();
}

__global__ void conv_in_wgrad_fold_kernel(const float* __restrict__ g, float* __restrict__ dW, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 36) return;
  const int n = i / 36, j = i % 36;
  dW[i] += g[n * 128 + j] + g[n * 128 + 36 + j];
}
cudaError_t conv_in_wgrad_fold_launch(const float* g, float* dW, int N, cudaStream_t s) {
  conv_in_wgrad_fold_kernel<<<(N * 36 + 255) / 256, 256, 0, s>>>(g, dW, N);
  return cudaGetLastError();
}

// =====================================================================================================
// Word_Attention backward (fp32): ctx = softmax(q k^T) v, unscaled.  One CTA per sample, L <= 16.
// =====================================================================================================
__global__ void __launch_bounds__(512) word_attn_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, const bf16_t* __restrict__ dctx,
                                                            bf16_t* __restrict__ d_qkv, int L, int D, int Ltot, int row_off) {
  __shared__ float sP[16][16], sDS[16][16];
  const int b = blockIdx.x;
  const float* qb = q + static_cast<size_t>(b) * L * D;
  const float* kb = k + static_cast<size_t>(b) * L * D;
  const float* vb = v + static_cast<size_t>(b) * L * D;
  const bf16_t* db = dctx + (static_cast<size_t>(b) * Ltot + row_off) * D;
  // scores and dP, one thread per (i, j)
  for (int ij = threadIdx.x; ij < L * L; ij += blockDim.x) {
    const int i = ij / L, j = ij % L;
    float sdot = 0.f, pdot = 0.f;
    for (int d = 0; d < D; ++d) {
      sdot = fmaf(qb[i * D + d], kb[j * D + d], sdot);
      pdot = fmaf(__bfloat162float(db[static_cast<size_t>(i) * D + d]), vb[j * D + d], pdot);
    }
    sP[i][j] = sdot;
    sDS[i][j] = pdot;
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < L) {
    const int i = threadIdx.x;
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) mx = fmaxf(mx, sP[i][j]);
    float den = 0.f;
    for (int j = 0; j < L; ++j) {
      const float e = expf(sP[i][j] - mx);
      sP[i][j] = e;
      den += e;
    }
    const float inv = 1.0f / den;
    float delta = 0.f;
    for (int j = 0; j < L; ++j) {
      sP[i][j] *= inv;
      delta = fmaf(sP[i][j], sDS[i][j], delta);
    }
    for (int j = 0; j < L; ++j) sDS[i][j] = sP[i][j] * (sDS[i][j] - delta);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    for (int i = 0; i < L; ++i) {
      float dq = 0.f, dk = 0.f, dv = 0.f;
      for (int j = 0; j < L; ++j) {
        dq = fmaf(sDS[i][j], kb[j * D + d], dq);
        dk = fmaf(sDS[j][i], qb[j * D + d], dk);
        dv = fmaf(sP[j][i], __bfloat162float(db[static_cast<size_t>(j) * D + d]), dv);
      }
      bf16_t* o = d_qkv + (static_cast<size_t>(b) * L + i) * 3 * D + d;
      o[0] = __float2bfloat16(dq);
      o[D] = __float2bfloat16(dk);
      o[2 * D] = __float2bfloat16(dv);
    }
  }
}
cudaError_t word_attn_bwd_launch(const float* q, const float* k, const float* v, const bf16_t* dctx, bf16_t* d_qkv, int B,
                                 int L, int D, int Ltot, int row_off, cudaStream_t s) {
  if (L < 1 || L > 16) return cudaErrorInvalidValue;
  int threads = D < 512 ? (D + 31) / 32 * 32 : 512;
  if (threads < L * L) threads = (L * L + 31) / 32 * 32;
  word_attn_bwd_kernel<<<B, threads, 0, s>>>(q, k, v, dctx, d_qkv, L, D, Ltot, row_off);
  return cudaGetLastError();
}

// =====================================================================================================
// transposed weight packs
// =====================================================================================================
__global__ void repack_linear_T_kernel(const float* __restrict__ w, bf16_t* __restrict__ dst, int N, int K, int ldn, int n_off,
                                       int k_row_off) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(N) * K) return;
  // consecutive threads walk n (coalesced writes; reads hit L2)
  const int n = idx % N, k = idx / N;
  dst[static_cast<size_t>(k_row_off + k) * ldn + n_off + n] = __float2bfloat16(w[static_cast<size_t>(n) * K + k]);
}
cudaError_t repack_linear_T_launch(const float* w, bf16_t* dst, int N, int K, int ldn, int n_off, int k_row_off, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * K;
  repack_linear_T_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, dst, N, K, ldn, n_off, k_row_off);
  return cudaGetLastError();
}

__global__ void repack_conv3x3_T_kernel(const float* __restrict__ w, bf16_t* __restrict__ dst, int Cout, int Cin, int cout_pad) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(Cout) * Cin * 9;
  if (idx >= total) return;
  const int n = idx % Cout;
  const int tap = (idx / Cout) % 9;
  const int c = idx / (static_cast<size_t>(Cout) * 9);
  dst[static_cast<size_t>(c) * 9 * cout_pad + (8 - tap) * cout_pad + n] =
      __float2bfloat16(w[(static_cast<size_t>(n) * Cin + c) * 9 + tap]);
}
cudaError_t repack_conv3x3_T_launch(const float* w, bf16_t* dst, int Cout, int Cin, int cout_pad, cudaStream_t s) {
  const size_t total = static_cast<size_t>(Cout) * Cin * 9;
  repack_conv3x3_T_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, dst, Cout, Cin, cout_pad);
  return cudaGetLastError();
}

// Transposed packs (kind 1 and 3) as 32 x 32 shared-memory tiles: coalesced reads along the source row and coalesced writes along
// the destination row.  (The element-per-thread form read the fp32 source with a stride of one weight row per thread: the merged
// pack took 0.57 ms per training step, ten times its memory traffic.)  A job is the transpose of a row-major [R, C] matrix:
//   kind 1: R = N, C = K,      dst[(off1 + c) ld + off0 + r]
//   kind 3: R = N, C = 9 K,    c = ch * 9 + tap -> dst[(ch * 9 + 8 - tap) ld + r]
// `start` counts 32 x 32 tiles here.
__global__ void __launch_bounds__(256) repack_multi_T_kernel(const PackDesc* __restrict__ jobs, int njobs, long long total_tiles) {
  __shared__ float tile[32][33];
  const long long t = blockIdx.x;
  if (t >= total_tiles) return;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].start <= t) lo = mid;
    else hi = mid - 1;
  }
  const PackDesc j = jobs[lo];
  const int R = j.N, C = j.kind == 3 ? 9 * j.K : j.K;
  const int tiles_c = (C + 31) >> 5;
  const int lt = static_cast<int>(t - j.start);
  const int r0 = (lt / tiles_c) << 5, c0 = (lt % tiles_c) << 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < R && c < C) ? j.src[static_cast<size_t>(r) * C + c] : 0.f;
  }
  __syncthreads();
  bf16_t* d = static_cast<bf16_t*>(j.dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (r >= R || c >= C) continue;
    size_t o;
    if (j.kind == 3) {
      const int ch = c / 9, tap = c - ch * 9;
      o = static_cast<size_t>(ch * 9 + 8 - tap) * j.ld + r;
    } else {
      o = static_cast<size_t>(j.off1 + c) * j.ld + j.off0 + r;
    }
    d[o] = __float2bfloat16(tile[tx][ty + 8 * i]);
  }
}
cudaError_t repack_multi_T_launch(const PackDesc* jobs_dev, int njobs, long long total_tiles, cudaStream_t s) {
  if (njobs <= 0 || total_tiles <= 0) return cudaSuccess;
  repack_multi_T_kernel<<<static_cast<unsigned>(total_tiles), 256, 0, s>>>(jobs_dev, njobs, total_tiles);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) repack_multi_kernel(const PackDesc* __restrict__ jobs, int njobs, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int lo = 0, hi = njobs - 1;  // last job whose start <= idx
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].start <= idx) lo = mid;
    else hi = mid - 1;
  }
  const PackDesc j = jobs[lo];
  const long long i = idx - j.start;
  bf16_t* d = static_cast<bf16_t*>(j.dst);
  auto to16 = [](float v, int as_f16) -> bf16_t {
    if (!as_f16) return __float2bfloat16(v);
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const bf16_t*>(&h);
  };
  switch (j.kind) {
    case 0: {
      const int k = static_cast<int>(i % j.K), n = static_cast<int>(i / j.K);
      d[static_cast<size_t>(n + j.off1) * j.ld + j.off0 + k] = __float2bfloat16(j.src[i]);
      break;
    }
    case 1: {
      const int n = static_cast<int>(i % j.N), k = static_cast<int>(i / j.N);
      d[static_cast<size_t>(j.off1 + k) * j.ld + j.off0 + n] = __float2bfloat16(j.src[static_cast<size_t>(n) * j.K + k]);
      break;
    }
    case 2: {
      const int tap = static_cast<int>(i % 9), c = static_cast<int>((i / 9) % j.K), n = static_cast<int>(i / (9LL * j.K));
      const float w = j.src[i];
      if (j.off1 == 2) {
        const float h = __bfloat162float(__float2bfloat16(w));
        d[static_cast<size_t>(n) * j.ld + j.off0 + tap * j.K + c] = __float2bfloat16(h);
        d[static_cast<size_t>(n + 4) * j.ld + j.off0 + tap * j.K + c] = __float2bfloat16(w - h);
      } else {
        d[static_cast<size_t>(n) * j.ld + j.off0 + tap * j.K + c] = to16(w, j.off1);
      }
      break;
    }
    case 3: {
      const int n = static_cast<int>(i % j.N), tap = static_cast<int>((i / j.N) % 9), c = static_cast<int>(i / (9LL * j.N));
      d[static_cast<size_t>(c) * 9 * j.ld + (8 - tap) * j.ld + n] = __float2bfloat16(j.src[(static_cast<size_t>(n) * j.K + c) * 9 + tap]);
      break;
    }
    default: {  // conv_in
      const int k = static_cast<int>(i % 128), n = static_cast<int>(i / 128);
      float v = 0.f;
      if (k < 108) {
        const float wv = j.src[n * 36 + k % 36];
        const float h = __bfloat162float(__float2bfloat16(wv));
        v = k < 72 ? h : wv - h;
      }
      d[i] = __float2bfloat16(v);
    }
  }
}
cudaError_t repack_multi_launch(const PackDesc* jobs_dev, int njobs, long long total, cudaStream_t s) {
  if (njobs <= 0 || total <= 0) return cudaSuccess;
  repack_multi_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(jobs_dev, njobs, total);
  return cudaGetLastError();
}

// =====================================================================================================
// AdamW + EMA
// =====================================================================================================
__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, size_t n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, float bc1, float bc2_sqrt, float ema_beta, int ema_mode,
                                 float grad_scale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = g[i] * grad_scale;
  float pv = p[i];
  pv *= (1.0f - lr * weight_decay);
  const float mi = beta1 * m[i] + (1.0f - beta1) * gr;
  const float vi = beta2 * v[i] + (1.0f - beta2) * gr * gr;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pv -= (lr / bc1) * (mi / denom);
  p[i] = pv;
  if (ema_mode == 1) ema[i] = pv;
  else if (ema_mode == 2) ema[i] = ema[i] * ema_beta + (1.0f - ema_beta) * pv;
}
cudaError_t adamw_ema_launch(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, float ema_beta, int ema_mode,
                             float grad_scale, cudaStream_t s) {
  if (step < 1) return cudaErrorInvalidValue;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_ema_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps,
                                                                          weight_decay, bc1, sqrtf(bc2), ema_beta, ema_mode,
                                                                          grad_scale);
  return cudaGetLastError();
}

}  // namespace wd
