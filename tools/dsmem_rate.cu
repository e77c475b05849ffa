// Distributed-shared-memory store throughput between the two CTAs of a cluster on sm_100a:
// st.async (16 B per lane, 64-byte or 128-byte runs) against cp.async.bulk shared::cta ->
// shared::cluster of 1/2/4/16 KB pieces, one direction and both directions at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/dsmem_rate tools/dsmem_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int TILE = 65536;  // bytes per round
// mode 0: st.async, 8 rows x 64 B per warp instruction; 1: st.async, 4 rows x 128 B; 2..5: bulk
// copies of 1, 2, 4, 16 KB.  senders: bit 0 = CTA 0 sends, bit 1 = CTA 1 sends.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
rate_kernel(int mode, int senders, int rounds, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* src = smem;          // 64 KB source
  uint8_t* dst = smem + TILE;   // 64 KB landing tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * TILE);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(src)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync();
  const bool sender = (senders >> rank) & 1, receiver = (senders >> peer) & 1;
  const uint32_t pdst = mapa(smem_u32(dst), peer), pbar = mapa(smem_u32(bar), peer);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if (receiver && threadIdx.x == 0) mbar_expect(bar, TILE);
    if (sender) {
      if (mode <= 1) {
        // 8 warps x 16 instructions x 512 B
        for (int i = 0; i < 16; ++i) {
          const uint32_t off = uint32_t(warp * 16 + i) * 512u;
          uint32_t a;
          if (mode == 0) a = off + uint32_t(lane >> 2) * 64u + uint32_t(lane & 3) * 16u;  // contiguous anyway
          else a = off + uint32_t(lane) * 16u;
          // scatter rows like the kernel does: 64-byte runs 128 B apart (mode 0), 128-byte runs (mode 1)
          if (mode == 0) a = uint32_t(warp * 16 + i) * 512u % TILE, a = (uint32_t(warp) * 8192u + uint32_t(i >> 1) * 1024u + uint32_t(lane >> 2) * 128u + uint32_t(i & 1) * 64u + uint32_t(lane & 3) * 16u);
          const uint4 v = *reinterpret_cast<const uint4*>(src + off + lane * 16);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                       ::"r"(pdst + a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(pbar) : "memory");
        }
      } else {
        const uint32_t piece = mode == 2 ? 1024u : mode == 3 ? 2048u : mode == 4 ? 4096u : 16384u;
        const int n = TILE / piece;
        // pieces are issued by lane 0 of the warps in turn, like one copy per warp and chunk
        for (int i = warp; i < n; i += 8)
          if (lane == 0)
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(pdst + i * piece), "r"(smem_u32(src) + i * piece), "r"(piece), "r"(pbar) : "memory");
      }
    }
    if (receiver) mbar_wait(bar, r & 1);
    // lock-step rounds: the sender may not overwrite before the receiver has seen the tile
    cluster_sync();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 8);
  const int smem = 2 * TILE + 64;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[6] = {"st.async 64-byte runs", "st.async 128-byte runs", "bulk 1 KB", "bulk 2 KB", "bulk 4 KB", "bulk 16 KB"};
  const int rounds = 200;
  for (int grid : {2, 148})
    for (int senders : {1, 3})
      for (int mode = 0; mode < 6; ++mode) {
        rate_kernel<<<grid, 256, smem>>>(mode, senders, rounds, out);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        // overhead of an empty round (cluster_sync) measured with senders = 0
        rate_kernel<<<grid, 256, smem>>>(mode, 0, rounds, out);
        cudaDeviceSynchronize();
        unsigned long long h0[2];
        cudaMemcpy(h0, out, 16, cudaMemcpyDeviceToHost);
        const double clk = double(h[0]) / rounds, base = double(h0[0]) / rounds;
        printf("grid %3d  %s  %-24s %8.0f clk/round (empty round %5.0f)  -> %5.1f B/clk per direction  %s\n", grid,
               senders == 1 ? "one way  " : "both ways", names[mode], clk, base, TILE / (clk - base),
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
