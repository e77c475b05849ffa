// Timeline of the pair backward kernel: builds sim_bwd_pair.cu with MMCL_TRACE and dumps, for one
// CTA pair, the SM-clock stamps of the TMA producer, the MMA issuer and the epilogue warpgroups
// per step (CSV: cta,event,step,clk), plus the launch time.  tools/bwd_trace_summary.py reads it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/bwd_trace tools/bwd_trace.cu
//   tools/bin/bwd_trace [out.csv]
#define MMCL_TRACE 1
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../mmlearn_b200/csrc/sim_bwd_pair.cu"
namespace mmcl { void count_launches(int) {} long long launches_so_far() { return 0; } }

int main(int argc, char** argv) {
  const long long nx = 32768, ny = 32768, d = 512;
  __nv_bfloat16 *x, *y; float *dx, *offr, *offc, *scale;
  cudaMalloc(&x, nx * d * 2); cudaMalloc(&y, ny * d * 2); cudaMalloc(&dx, nx * d * 4);
  cudaMalloc(&offr, nx * 4); cudaMalloc(&offc, ny * 4); cudaMalloc(&scale, 4);
  std::vector<__nv_bfloat16> h(nx * d);
  unsigned s = 1;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = __float2bfloat16(((s >> 8) & 0xffff) / 65536.f * 0.09f - 0.045f); }
  cudaMemcpy(x, h.data(), nx * d * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(y, h.data(), ny * d * 2, cudaMemcpyHostToDevice);
  std::vector<float> off(nx, 30.f);
  cudaMemcpy(offr, off.data(), nx * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(offc, off.data(), ny * 4, cudaMemcpyHostToDevice);
  float sc = 14.285f; cudaMemcpy(scale, &sc, 4, cudaMemcpyHostToDevice);
  cudaMemset(dx, 0, nx * d * 4);
  const char* err = nullptr;
  for (int rep = 0; rep < 5; ++rep) {
    void* tp; cudaGetSymbolAddress(&tp, mmcl::g_trace); cudaMemset(tp, 0, sizeof(mmcl::g_trace));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = mmcl::launch_sim_grad_bwd_pair(148, x, nx, y, ny, d, scale, nullptr, nullptr, offr, offc, 1.f, 0.f, 0.f, dx, 0, &err);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d: launch %s sync %s  %.3f ms\n", rep, cudaGetErrorString(e), cudaGetErrorString(e2), ms);
  }
  static unsigned long long tr[2][4][mmcl::kTraceCap][2];
  cudaMemcpyFromSymbol(tr, mmcl::g_trace, sizeof(tr));
  FILE* f = fopen(argc > 1 ? argv[1] : "bwd_trace.csv", "w");
  fprintf(f, "cta,event,step,clk\n");
  unsigned n = 0;
  for (int c = 0; c < 2; ++c)
    for (int r = 0; r < 4; ++r)
      for (unsigned i = 0; i < mmcl::kTraceCap && tr[c][r][i][1]; ++i, ++n)
        fprintf(f, "%d,%u,%u,%llu\n", c, unsigned(tr[c][r][i][0] >> 32), unsigned(tr[c][r][i][0] & 0xffffffffu), tr[c][r][i][1]);
  fclose(f);
  printf("events: %u\n", n);
  return 0;
}
