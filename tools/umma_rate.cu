// Measures tcgen05.mma issue/throughput per SM for the tile shapes the kernels may use:
// M=128, N in {64,128,256}, A from shared memory (SS) or tensor memory (TS), B K-major or MN-major.
// 148 CTAs each issue a long chain of MMAs on fixed operands; cycles are taken with clock64.
// Prints MACs/clk/SM (peak is 4096 for bf16).   Run: tools/bin/umma_rate
#include <cstdio>
#include <vector>
#include "../mmlearn_b200/csrc/umma.cuh"

using namespace mmcl;

__global__ void __launch_bounds__(128, 1)
rate_kernel(int N, int a_tmem, int b_mn, int iters, long long* cycles_out, int nacc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // zero operands (values do not matter for timing, but avoid NaN garbage)
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, b_mn);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t bdesc = b_mn ? make_smem_desc_sw128(sb + k * 2048, 64 * 128, 1024)
                                    : make_smem_desc_sw128(sb + k * 32, 0, 1024);
        if (a_tmem)
          umma_ts(tmem + ((it % nacc) * N) % 256, tmem + 256 + k * 8, bdesc, idesc, 1);
        else
          umma_ss(tmem + ((it % nacc) * N) % 256, make_smem_desc_sw128(sa + k * 32, 0, 1024), bdesc, idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}


// CTA-pair version: M=256 (128 rows per CTA), N=256, each CTA supplies half of B.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
rate_kernel_2cta(int N, int a_tmem, int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t cta = cluster_ctarank();
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_2cta<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (cta == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t bdesc = make_smem_desc_sw128(sb + k * 32, 0, 1024);
        if (a_tmem)
          umma_ts_2cta(tmem, tmem + 256 + k * 8, bdesc, idesc, 1);
        else
          umma_ss_2cta(tmem, make_smem_desc_sw128(sa + k * 32, 0, 1024), bdesc, idesc, 1);
      }
    }
    umma_commit_2cta(&bar, 3);
  }
  if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2cta<512>(tmem);
}

int main() {
  long long* d_cyc;
  cudaMalloc(&d_cyc, 148 * sizeof(long long));
  int smem_bytes = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 2048;
  for (int ctas : {148})
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
      for (int b_mn = 0; b_mn < 2; ++b_mn)
        for (int N : {64, 128, 256}) {
          for (int nacc : {1, 2}) {
          rate_kernel<<<ctas, 128, smem_bytes>>>(N, a_tmem, b_mn, 64, d_cyc, nacc);  // warm
          rate_kernel<<<ctas, 128, smem_bytes>>>(N, a_tmem, b_mn, iters, d_cyc, nacc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<long long> cyc(ctas);
          cudaMemcpy(cyc.data(), d_cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (auto c : cyc) mx = c > mx ? c : mx;
          double macs = double(iters) * 4 * 128.0 * N * 16;
          printf("ctas=%3d A=%s B=%s N=%3d nacc=%d : %8lld cyc  %7.1f MAC/clk/SM  (%.1f clk/MMA)\n", ctas,
                 a_tmem ? "tmem" : "smem", b_mn ? "MN" : "K ", N, nacc, mx, macs / mx,
                 double(mx) / (iters * 4));
          }
        }
  cudaFuncSetAttribute(rate_kernel_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int ctas : {2, 148})
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
      for (int N : {128, 256}) {
        rate_kernel_2cta<<<ctas, 128, smem_bytes>>>(N, a_tmem, 64, d_cyc);
        rate_kernel_2cta<<<ctas, 128, smem_bytes>>>(N, a_tmem, iters, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("2cta error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> cyc(ctas);
        cudaMemcpy(cyc.data(), d_cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (auto c : cyc) mx = c > mx ? c : mx;
        double macs = double(iters) * 4 * 256.0 * N * 16 / 2;  // per SM
        printf("2CTA ctas=%3d A=%s N=%3d : %8lld cyc  %7.1f MAC/clk/SM  (%.1f clk/MMA)\n", ctas,
               a_tmem ? "tmem" : "smem", N, mx, macs / mx, double(mx) / (iters * 4));
      }
  return 0;
}
