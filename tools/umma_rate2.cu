// MMA throughput vs issue pacing, with a realistic smem footprint (A: 8 resident 16 KB chunks,
// B: 3 stages of 32 KB). Run: tools/bin/umma_rate2
#include <cstdio>
#include <vector>
#include "../mmlearn_b200/csrc/umma.cuh"
using namespace mmcl;

template <int CTAS>
__global__ void __launch_bounds__(128, 1)
rate_kernel(int N, int a_tmem, int iters, int pace, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t cta = CTAS == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (131072 + 98304) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { if (CTAS == 2) tmem_alloc_2cta<512>(&tmem_slot); else tmem_alloc<512>(&tmem_slot); }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (cta == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128 * CTAS, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 131072);
    int ac = 0, bs = 0;
    long long tl = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (pace > 0) { while (clock64() - tl < pace) {} tl = clock64(); }
        const uint64_t bdesc = make_smem_desc_sw128(sb + bs * 32768 + k * 32, 0, 1024);
        if (a_tmem) {
          if (CTAS == 2) umma_ts_2cta(tmem, tmem + 256 + k * 8, bdesc, idesc, 1);
          else umma_ts(tmem, tmem + 256 + k * 8, bdesc, idesc, 1);
        } else {
          const uint64_t adesc = make_smem_desc_sw128(sa + ac * 16384 + k * 32, 0, 1024);
          if (CTAS == 2) umma_ss_2cta(tmem, adesc, bdesc, idesc, 1);
          else umma_ss(tmem, adesc, bdesc, idesc, 1);
        }
      }
      if (++ac == 8) ac = 0;
      if (++bs == 3) bs = 0;
    }
    if (CTAS == 2) umma_commit_2cta(&bar, 3); else umma_commit(&bar);
  }
  if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
    cycles_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { if (CTAS == 2) tmem_dealloc_2cta<512>(tmem); else tmem_dealloc<512>(tmem); }
}

template <int CTAS>
void run(int N, int a_tmem, int pace, long long* d_cyc) {
  const int iters = 2048, ctas = 148;
  int smem_bytes = 131072 + 98304 + 1024;
  cudaFuncSetAttribute(rate_kernel<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    int its = rep ? iters : 64;
    cudaLaunchKernelEx(&cfg, rate_kernel<CTAS>, N, a_tmem, its, pace, d_cyc);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> cyc(ctas);
  cudaMemcpy(cyc.data(), d_cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto c : cyc) mx = c > mx ? c : mx;
  double macs = double(iters) * 4 * 128.0 * N * 16;
  printf("%dCTA A=%s N=%3d pace=%3d : %8lld cyc  %7.1f MAC/clk/SM  (%.1f clk/MMA)\n", CTAS,
         a_tmem ? "tmem" : "smem", N, pace, mx, macs / mx, double(mx) / (iters * 4));
}

int main() {
  long long* d_cyc;
  cudaMalloc(&d_cyc, 148 * sizeof(long long));
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int N : {128, 256})
      for (int pace : {0, 32, 64, 100, 128, 140}) {
        run<1>(N, a_tmem, pace, d_cyc);
        run<2>(N, a_tmem, pace, d_cyc);
      }
  return 0;
}
