// Micro-benchmark (GPU box): issue rate of tcgen05.mma.cta_group::1.kind::f16 128xNx16 (SS operands, SWIZZLE_128B
// K-major tiles in shared memory, garbage data) as a function of N, one CTA per SM.  Prints cycles per MMA and the
// implied fraction of the 4096 MAC/clk/SM peak, plus the SM clock seen (clock64 vs globaltimer).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../../worddiffusion_b200/csrc/common.cuh"
using namespace wd;

template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int M, int n_mma, int stages_rt, long long* out_clk, long long* out_ns) {
  constexpr int stages = 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < stages * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32 && elect_one()) {  // elect.sync: no ELECT/BRA.U.ANY loop around each UTCHMMA
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(N) >> 3) << 17) | ((uint32_t(M) >> 4) << 24);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      unsigned long long g0;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
      if (MODE == 0) {
        // back-to-back issue, one commit at the end
        for (int i = 0; i < n_mma; i += 4) {
          const int st = (i >> 2) & (stages - 1);
          const uint32_t a_addr = smem_u32(smem + st * 49152);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + 16384);
          const uint32_t d = tmem + ((i >> 8) & 1) * 256;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d, a_desc + 2 * k, b_desc + 2 * k, idesc, ((i & 255) | k) != 0);
        }
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
      } else {
        // the real kernel's pattern: commit per k-block (4 MMAs) to a ring of barriers, but never wait inside the loop
        for (int i = 0; i < n_mma; i += 4) {
          const int st = (i >> 2) & (stages - 1);
          const uint32_t a_addr = smem_u32(smem + st * 49152);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + 16384);
          const uint32_t d = tmem + ((i >> 8) & 1) * 256;
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d, a_desc + 2 * k, b_desc + 2 * k, idesc, ((i & 255) | k) != 0);
          umma_commit(&bar2[st]);
        }
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
      long long t1 = clock64();
      unsigned long long g1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
      out_clk[blockIdx.x] = t1 - t0;
      out_ns[blockIdx.x] = (long long)(g1 - g0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int stages = 4, n_mma = 4096;
  const int smem = stages * 49152 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *d_clk, *d_ns;
  cudaMalloc(&d_clk, sms * 8);
  cudaMalloc(&d_ns, sms * 8);
  std::vector<long long> clk(sms), ns(sms);
  printf("%4s %4s %6s %10s %10s %8s %8s\n", "M", "N", "grid", "clk/mma", "ideal", "eff", "GHz");
  for (int mode : {0, 1}) {
    const int grid = sms;
    printf("mode %d (%s)\n", mode, mode ? "commit per 4 MMAs" : "single commit");
    for (int M : {128}) {
      for (int N : {16, 32, 64, 80, 96, 128, 160, 192, 224, 256}) {
        if (M == 64 && N % 8) continue;
        if (mode == 0) mma_rate_kernel<0><<<grid, 128, smem>>>(N, M, n_mma, stages, d_clk, d_ns);
        else mma_rate_kernel<1><<<grid, 128, smem>>>(N, M, n_mma, stages, d_clk, d_ns);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(clk.data(), d_clk, grid * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(ns.data(), d_ns, grid * 8, cudaMemcpyDeviceToHost);
        double c = 0, t = 0;
        for (int i = 0; i < grid; ++i) { c += clk[i]; t += ns[i]; }
        c /= grid; t /= grid;
        const double per = c / n_mma;
        const double ideal = double(M) * N * 16 / 4096.0;  // 4096 MAC/clk/SM
        printf("%4d %4d %6d %10.1f %10.1f %8.3f %8.3f\n", M, N, grid, per, ideal, ideal / per, c / t);
      }
    }
  }
  return 0;
}
