// Micro-benchmark (GPU box): issue rate of tcgen05.mma.cta_group::2.kind::f16 256xNx16 (CTA pair, SS operands, garbage
// data), one cluster of 2 per TPC.  Prints cycles per MMA vs the ideal N/2 (4096 MAC/clk/SM on both SMs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_pair mma_rate_pair.cu && ./mma_rate_pair
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../../worddiffusion_b200/csrc/common.cuh"
using namespace wd;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_rate_kernel(int N, int n_mma, int split, long long* out_clk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  constexpr int stages = 4;
  for (int i = threadIdx.x; i < stages * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc_pair<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  const bool leader = cluster_ctarank() == 0;
  if (threadIdx.x < 32 && leader && elect_one()) {
    // split == 0: one MMA of N columns per K step;  split == 1: N = 320 as 160 + 160;  split == 2: 320 as 256 + 64
    const uint32_t idN = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(N) >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t id160 = (1u << 4) | (1u << 7) | (1u << 10) | ((160u >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t id256 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t id64 = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((256u >> 4) << 24);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
        const int st = (i >> 2) & (stages - 1);
        const uint32_t a_addr = smem_u32(smem + st * 49152);
        const uint64_t a_desc = make_smem_desc_sw128(a_addr);
        const uint64_t b_desc = make_smem_desc_sw128(a_addr + 16384);
        const uint64_t b_desc2 = make_smem_desc_sw128(a_addr + 16384 + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = ((i & 255) | k) != 0;
          if (split == 0) {
            umma_f16_ss_pair(tmem, a_desc + 2 * k, b_desc + 2 * k, idN, acc);
          } else if (split == 1) {
            umma_f16_ss_pair(tmem, a_desc + 2 * k, b_desc + 2 * k, id160, acc);
            umma_f16_ss_pair(tmem + 160, a_desc + 2 * k, b_desc2 + 2 * k, id160, acc);
          } else {
            umma_f16_ss_pair(tmem, a_desc + 2 * k, b_desc + 2 * k, id256, acc);
            umma_f16_ss_pair(tmem + 256, a_desc + 2 * k, b_desc2 + 2 * k, id64, acc);
          }
        }
      }
      umma_commit_pair(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      out_clk[blockIdx.x >> 1] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_pair<512>(tmem);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n_mma = 4096;
  const int smem = 4 * 49152 + 1024;
  cudaFuncSetAttribute(pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d_clk;
  cudaMalloc(&d_clk, sms * 8);
  std::vector<long long> clk(sms);
  printf("%6s %5s %10s %10s %8s\n", "split", "N", "clk/kstep", "ideal", "eff");
  for (int split : {0, 1, 2}) {
    for (int N : {32, 64, 96, 128, 160, 192, 224, 256}) {
      if (split && N != 160) continue;
      pair_rate_kernel<<<sms, 128, smem>>>(N, n_mma, split, d_clk);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("split=%d N=%d: %s\n", split, N, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(clk.data(), d_clk, (sms / 2) * 8, cudaMemcpyDeviceToHost);
      double c = 0;
      for (int i = 0; i < sms / 2; ++i) c += clk[i];
      c /= (sms / 2);
      const double per = c / n_mma;
      const double ideal = split ? 160.0 : N / 2.0;
      printf("%6d %5d %10.1f %10.1f %8.3f\n", split, split ? 320 : N, per, ideal, ideal / per);
    }
  }
  return 0;
}
