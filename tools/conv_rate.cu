// Throughput of the logit -> G conversion of the pair backward kernel in isolation (tensor-memory
// load, conv16_fast, tensor-memory store), for 8 epilogue warps working on 32-column chunks with a
// two-deep load pipeline (the kernel's arrangement) against 16 warps working on 16-column pieces.
// One CTA per SM; clocks per 128 x 256 tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/conv_rate tools/conv_rate.cu
#include <cstdio>
#include "../mmlearn_b200/csrc/sim_bwd_pair.cu"
namespace mmcl { void count_launches(int) {} long long launches_so_far() { return 0; } }
using namespace mmcl;

// MODE 0: full conversion; 1: without the exponentials; 2: tensor-memory load + store only
template <int MODE>
__device__ __forceinline__ void conv_probe(const uint32_t (&v)[16], uint32_t* g, float c, float offr,
                                           float R, const float4* cy4) {
  if (MODE == 0) {
    conv16_fast(v, g, c, offr, R, cy4);
  } else if (MODE == 1) {
    const float2 c2 = make_float2(c, c), o2 = make_float2(-offr, -offr), r2 = make_float2(R, R);
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const float4 f = cy4[e >> 1];
      const float2 yi = make_float2(f.z, f.w);
      const float2 t = __fmul2_rn(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), yi);
      const float2 l = __ffma2_rn(t, c2, o2);
      const float2 w = __ffma2_rn(r2, make_float2(f.x, f.y), yi);
      const float2 gg = __fmul2_rn(l, w);
      g[e >> 1] = pack_bf16x2(gg.x, gg.y);
    }
  } else if (MODE == 2) {
#pragma unroll
    for (int e = 0; e < 16; e += 2) g[e >> 1] = v[e] ^ v[e + 1];
  } else {  // 3: scalar fp32 arithmetic, full; 4: scalar, without the exponentials
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const float4 f = cy4[e >> 1];
      const float t0 = __uint_as_float(v[e]) * f.z, t1 = __uint_as_float(v[e + 1]) * f.w;
      const float l0 = fmaf(t0, c, -offr), l1 = fmaf(t1, c, -offr);
      const float e0 = MODE == 3 ? ex2f(l0) : l0, e1 = MODE == 3 ? ex2f(l1) : l1;
      const float w0 = fmaf(R, f.x, f.z), w1 = fmaf(R, f.y, f.w);
      g[e >> 1] = pack_bf16x2(e0 * w0, e1 * w1);
    }
  }
}

// mma: 0 = tensor pipe idle; 1 = warp 1 keeps issuing dX-shaped MMAs (A in tensor memory, B from
// shared memory, N = 256) into the accumulator columns while the conversion runs; 2 = SS form
template <int NW, int MODE>
__global__ void __launch_bounds__(NW * 32 + 128, 1) conv_kernel(int tiles, long long* out, int mma) {
  extern __shared__ __align__(1024) uint8_t opnd[];  // 48 KB of operands (values do not matter)
  __shared__ uint64_t done_bar;
  __shared__ volatile int stop;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float2 sCol[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sCol[i] = make_float2(0.5f, 1.f);
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(opnd)[i] = 0;
  if (threadIdx.x == 0) {
    stop = 0;
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_s = tmem_slot + 256;
  long long t0 = clock64();
  if (warp == 1 && mma) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, 256, 0, 1);
      const uint32_t sb = smem_u32(opnd);
      uint32_t n = 0;
      while (!stop) {
        for (int k = 0; k < 4; ++k) {
          const uint64_t bdesc = make_smem_desc_sw128(sb + 16384 + k * 2048, 8192, 1024);
          if (mma == 1) umma_ts(tmem_slot, tm_s + k * 8, bdesc, idesc, 1);
          else umma_ss(tmem_slot, make_smem_desc_sw128(sb + k * 32, 0, 1024), bdesc, idesc, 1);
        }
        umma_commit(&done_bar);
        mbar_wait(&done_bar, n & 1);  // one group of 4 in flight at a time keeps the pipe ~busy
        ++n;
      }
    }
  }
  if (warp >= 4) {
    const int q = warp & 3, j = (warp - 4) >> 2;  // quadrant, piece index within a quarter
    const uint32_t sbase = tm_s + (uint32_t(q * 32) << 16);
    const float cc = 0.01f, offr = 3.f, R = 0.25f;
    for (int t = 0; t < tiles; ++t) {
      if (NW == 8) {
        uint32_t va[16], vb[16], g[16];
        tmem_ld_x16(sbase + uint32_t(j) * 32u, va);
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int ch = ci * 2 + j;
          tmem_ld_wait();
          tmem_ld_x16(sbase + ch * 32 + 16, vb);
          conv_probe<MODE>(va, g, cc, offr, R, reinterpret_cast<const float4*>(sCol) + ((ch * 32) >> 1));
          tmem_ld_wait();
          if (ci < 3) tmem_ld_x16(sbase + ((ci + 1) * 2 + j) * 32, va);
          conv_probe<MODE>(vb, g + 8, cc, offr, R, reinterpret_cast<const float4*>(sCol) + ((ch * 32 + 16) >> 1));
          tmem_st_x16(sbase + ch * 32, g);
          tmem_st_wait();
          tc_fence_before();
          named_bar_sync(1, NW * 32);
        }
      } else {
        uint32_t v[16], g[8];
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int p = ci * 4 + j;
          tmem_ld_x16(sbase + p * 16, v);
          tmem_ld_wait();
          conv_probe<MODE>(v, g, cc, offr, R, reinterpret_cast<const float4*>(sCol) + ((p * 16) >> 1));
          tmem_st_x8(sbase + p * 16, g);
          tmem_st_wait();
          tc_fence_before();
          named_bar_sync(1, NW * 32);
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 128) {
    out[blockIdx.x] = t1 - t0;
    stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_slot);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  const int tiles = 400;
  auto report = [&](const char* name, cudaError_t e) {
    long long h;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-58s %7.0f clk per tile  %s\n", name, double(h) / tiles, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
#define RUN(NW, MODE, NAME)                                                   \
  cudaFuncSetAttribute(conv_kernel<NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024); \
  conv_kernel<NW, MODE><<<148, NW * 32 + 128, 49152 + 1024>>>(tiles, out, 0); \
  report(NAME, cudaDeviceSynchronize());                                      \
  conv_kernel<NW, MODE><<<148, NW * 32 + 128, 49152 + 1024>>>(tiles, out, 1); \
  report(NAME " + TS MMAs", cudaDeviceSynchronize());                         \
  conv_kernel<NW, MODE><<<148, NW * 32 + 128, 49152 + 1024>>>(tiles, out, 2); \
  report(NAME " + SS MMAs", cudaDeviceSynchronize());
  RUN(8, 0, "8 warps, 32-column chunks, full conversion");
  RUN(16, 0, "16 warps, 16-column pieces, full conversion");
  RUN(8, 1, "8 warps, without the exponentials");
  RUN(16, 1, "16 warps, without the exponentials");
  RUN(8, 3, "8 warps, scalar fp32 arithmetic, full conversion");
  RUN(16, 3, "16 warps, scalar fp32 arithmetic, full conversion");
  RUN(8, 4, "8 warps, scalar, without the exponentials");
  RUN(16, 4, "16 warps, scalar, without the exponentials");
  RUN(8, 2, "8 warps, tensor-memory load + store only");
  RUN(16, 2, "16 warps, tensor-memory load + store only");
  return 0;
}
