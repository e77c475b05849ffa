// Hardware probe for the tcgen05 operand encodings this repo relies on. One CTA, one MMA chain,
// checked against a host fp32 product of the same bf16-rounded values. Run on a B200:
//   tools/bin/umma_probe
// Variants: A from smem K-major / smem MN-major / TMEM; B from smem K-major / MN-major;
// N in {64,128,256}; K in {64,128}; TMA box larger than the tensor (zero fill).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../mmlearn_b200/csrc/umma.cuh"
#include "../mmlearn_b200/csrc/tma_host.h"

using namespace mmcl;

enum { A_SMEM_K = 0, A_SMEM_MN = 1, A_TMEM = 2 };
enum { B_K = 0, B_MN = 1 };

struct ProbeParams {
  int a_mode, b_mode, N, K;
  const float* a_f32;  // [128, K] fp32 (bf16-representable), used for A_TMEM
  float* d_out;        // [128, N]
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;                 // up to 128 x 128 x 2 = 32 KB
  uint8_t* sB = smem + 32768;         // up to 256 x 128 x 2 = 64 KB
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, K = p.K;

  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem;          // columns [0, N)
  const uint32_t tmem_a = tmem + 256;    // columns [256, 256 + K/2)

  if (p.a_mode == A_TMEM) {
    // thread = row m; pack (k, k+1) into one 32-bit column
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float lo = p.a_f32[tid * K + 2 * (c0 + c)];
        float hi = p.a_f32[tid * K + 2 * (c0 + c) + 1];
        v[c] = pack_bf16x2(lo, hi);
      }
      tmem_st_x16(tmem_a + (uint32_t(warp * 32) << 16) + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    uint32_t bytes = 0;
    if (p.a_mode == A_SMEM_K) {
      for (int kb = 0; kb < K / 64; ++kb) tma_load_2d(sA + kb * 16384, &tmA, kb * 64, 0, &bar_full);
      bytes += 128 * K * 2;
    } else if (p.a_mode == A_SMEM_MN) {
      for (int mb = 0; mb < 2; ++mb) tma_load_2d(sA + mb * K * 128, &tmA, mb * 64, 0, &bar_full);
      bytes += 128 * K * 2;
    }
    if (p.b_mode == B_K) {
      for (int kb = 0; kb < K / 64; ++kb)
        tma_load_2d(sB + kb * N * 128, &tmB, kb * 64, 0, &bar_full);
    } else {
      for (int nb = 0; nb < N / 64; ++nb)
        tma_load_2d(sB + nb * K * 128, &tmB, nb * 64, 0, &bar_full);
    }
    bytes += N * K * 2;
    mbar_arrive_expect_tx(&bar_full, bytes);
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc =
        make_idesc_bf16(128, N, p.a_mode == A_SMEM_MN ? 1 : 0, p.b_mode == B_MN ? 1 : 0);
    for (int kk = 0; kk < K / 16; ++kk) {
      uint64_t bdesc;
      if (p.b_mode == B_K)
        bdesc = make_smem_desc_sw128(smem_u32(sB) + (kk / 4) * N * 128 + (kk % 4) * 32, 0, 1024);
      else
        bdesc = make_smem_desc_sw128(smem_u32(sB) + kk * 2048, K * 128, 1024);
      if (p.a_mode == A_TMEM) {
        umma_ts(tmem_d, tmem_a + kk * 8, bdesc, idesc, kk > 0);
      } else {
        uint64_t adesc;
        if (p.a_mode == A_SMEM_K)
          adesc = make_smem_desc_sw128(smem_u32(sA) + (kk / 4) * 16384 + (kk % 4) * 32, 0, 1024);
        else
          adesc = make_smem_desc_sw128(smem_u32(sA) + kk * 2048, K * 128, 1024);
        umma_ss(tmem_d, adesc, bdesc, idesc, kk > 0);
      }
    }
    umma_commit(&bar_done);
  }
  __syncwarp();
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(tmem_d + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) p.d_out[tid * N + c0 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

static int run_case(int a_mode, int b_mode, int N, int K, int n_valid) {
  // logical A [128, K], B [N, K]; rows of B >= n_valid do not exist in global memory (TMA fills 0)
  std::vector<float> A(128 * K), B(size_t(N) * K);
  for (auto& v : A) v = bf16r((rand() % 2001 - 1000) / 500.0f);
  for (auto& v : B) v = bf16r((rand() % 2001 - 1000) / 500.0f);
  for (int n = n_valid; n < N; ++n)
    for (int k = 0; k < K; ++k) B[size_t(n) * K + k] = 0.f;
  std::vector<__nv_bfloat16> Ag(128 * K), Bg(size_t(N) * K);
  // global layouts
  if (a_mode == A_SMEM_MN) {
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) Ag[k * 128 + m] = __float2bfloat16(A[m * K + k]);
  } else {
    for (int i = 0; i < 128 * K; ++i) Ag[i] = __float2bfloat16(A[i]);
  }
  if (b_mode == B_MN) {
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) Bg[size_t(k) * N + n] = __float2bfloat16(B[size_t(n) * K + k]);
  } else {
    for (size_t i = 0; i < size_t(N) * K; ++i) Bg[i] = __float2bfloat16(B[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float *dAf, *dD;
  CK(cudaMalloc(&dA, Ag.size() * 2));
  CK(cudaMalloc(&dB, Bg.size() * 2));
  CK(cudaMalloc(&dAf, A.size() * 4));
  CK(cudaMalloc(&dD, size_t(128) * N * 4));
  CK(cudaMemcpy(dA, Ag.data(), Ag.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bg.data(), Bg.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAf, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, size_t(128) * N * 4));
  CUtensorMap tmA, tmB;
  int rc;
  if (a_mode == A_SMEM_MN)
    rc = make_tmap_bf16_sw128(&tmA, dA, K, 128, 128, K);
  else
    rc = make_tmap_bf16_sw128(&tmA, dA, 128, K, K, 128);
  if (rc) { printf("tmap A failed %d\n", rc); return 1; }
  if (b_mode == B_MN)
    rc = make_tmap_bf16_sw128(&tmB, dB, K, N, N, K);
  else
    rc = make_tmap_bf16_sw128(&tmB, dB, n_valid, K, K, N);  // box rows N > tensor rows n_valid
  if (rc) { printf("tmap B failed %d\n", rc); return 1; }
  ProbeParams p{a_mode, b_mode, N, K, dAf, dD};
  int smem_bytes = 32768 + 65536 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_kernel<<<1, 128, smem_bytes>>>(tmA, tmB, p);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(size_t(128) * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += double(A[m * K + k]) * B[size_t(n) * K + k];
      maxerr = fmax(maxerr, fabs(ref - D[size_t(m) * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  const char* an[] = {"smemK", "smemMN", "tmem"};
  const char* bn[] = {"K", "MN"};
  printf("A=%-6s B=%-2s N=%3d K=%3d nvalid=%3d : maxerr=%.4g (maxref %.3g) %s\n", an[a_mode],
         bn[b_mode], N, K, n_valid, maxerr, maxref, maxerr < 1e-2 * fmax(1.0, maxref) ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dAf); cudaFree(dD);
  return maxerr < 1e-2 * fmax(1.0, maxref) ? 0 : 2;
}

int main() {
  int fails = 0;
  for (int K : {64, 128})
    for (int N : {64, 128, 256})
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 2; ++b) fails += run_case(a, b, N, K, N) != 0;
  // zero-fill: TMA box (128 rows) larger than the tensor (40 rows)
  fails += run_case(A_SMEM_K, B_K, 128, 64, 40) != 0;
  printf("probe: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}
