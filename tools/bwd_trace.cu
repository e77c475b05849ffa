// Timeline of the pair backward kernel: builds sim_bwd_pair.cu with MMCL_TRACE and prints, for
// pair 0, the globaltimer stamps of the MMA issuer and the epilogue warpgroups per step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/bwd_trace tools/bwd_trace.cu
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
  for (int rep = 0; rep < 3; ++rep) {
    unsigned zero[2] = {0, 0};
    cudaMemcpyToSymbol(mmcl::g_trace_n, zero, sizeof(zero));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = mmcl::launch_sim_grad_bwd_pair(148, x, nx, y, ny, d, scale, nullptr, nullptr, offr, offc, 1.f, 0.f, 0.f, dx, 0, &err);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d: launch %s sync %s  %.3f ms\n", rep, cudaGetErrorString(e), cudaGetErrorString(e2), ms);
  }
  static unsigned long long tr[2][4096][2]; unsigned n[2];
  cudaMemcpyFromSymbol(tr, mmcl::g_trace, sizeof(tr)); cudaMemcpyFromSymbol(n, mmcl::g_trace_n, sizeof(n));
  unsigned long long t0 = ~0ull;
  for (int c = 0; c < 2; ++c) for (unsigned i = 0; i < std::min(n[c], 4096u); ++i) t0 = std::min(t0, tr[c][i][1]);
  const char* names[16] = {"S.issue", "S.issued", "dXp.wait", "dXp.go", "dXp.issued", "dXo.h0", "dXo.h1", "dXo.issued", "", "", "epi0.S", "epi1.S", "epi0.free", "epi1.free", "epi0.done", "epi1.done"};
  for (int c = 0; c < 2; ++c) {
    printf("---- CTA %d: %u events\n", c, n[c]);
    std::vector<std::pair<unsigned long long, unsigned long long>> ev;
    for (unsigned i = 0; i < std::min(n[c], 4096u); ++i) ev.push_back({tr[c][i][1], tr[c][i][0]});
    std::sort(ev.begin(), ev.end());
    int shown = 0;
    for (auto& e : ev) {
      unsigned evid = unsigned(e.second >> 32), step = unsigned(e.second & 0xffffffff);
      if (step >= 20 && step < 26) { printf("  t=%8llu ns  %-10s step %u\n", e.first - t0, names[evid & 15], step); ++shown; }
    }
  }
  return 0;
}
