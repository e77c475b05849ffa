// HBM-bound helper kernels around the two tensor-core kernels: row packing (select + L2
// normalise + bf16 cast), positive logits, loss assembly, offset vectors, gradient unpack
// (positive-label term, normalise backward, scatter-add).  One warp per embedding row,
// 16-byte vector accesses.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "kernels.h"

namespace mmcl {

namespace {

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// load 8 consecutive elements (d % 8 == 0, rows 16-byte aligned) as fp32
template <int DT>
__device__ __forceinline__ void load8(const void* base, long long elem, float (&f)[8]) {
  if (DT == 0) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else if (DT == 1) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __half*>(base) + elem));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int DT>
__global__ void pack_rows_kernel(const void* __restrict__ src, long long src_rows, int d,
                                 const int64_t* __restrict__ idx, long long n_out, int normalize,
                                 __nv_bfloat16* __restrict__ dst, float* __restrict__ inv_norm) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_out) return;
  long long srow = idx ? idx[row] : row;
  if (srow < 0 || srow >= src_rows) srow = 0;  // defensive; the matcher only emits valid rows
  float inv = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
      float f[8];
      load8<DT>(src, srow * d + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
    }
    __syncwarp();  // (lanes leave the loop after different trip counts)
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
  }
  for (int c = lane * 8; c < d; c += 256) {
    float f[8];
    load8<DT>(src, srow * d + c, f);
    uint4 o;
    // raw values: the 1/||x|| factor is applied in fp32 inside the GEMM epilogues
    o.x = bf16x2(f[0], f[1]);
    o.y = bf16x2(f[2], f[3]);
    o.z = bf16x2(f[4], f[5]);
    o.w = bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + row * d + c) = o;
  }
  if (inv_norm && lane == 0) inv_norm[row] = inv;
}

// log_s <- clamp(log_s, 0, log(max)) in place; s = exp(log_s)   (the caller glue of the task)
__global__ void logit_scale_kernel(float* __restrict__ log_s, long long n, float max_log,
                                   float* __restrict__ s_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fminf(fmaxf(log_s[i], 0.f), max_log);
  log_s[i] = v;
  s_out[i] = expf(v);
}

__global__ void positive_logits_kernel(const __nv_bfloat16* __restrict__ q, long long nq,
                                       const __nv_bfloat16* __restrict__ k, long long nk, int d,
                                       long long label_offset, const float* __restrict__ scale,
                                       const float* __restrict__ qinv,
                                       const float* __restrict__ kinv, float* __restrict__ pos,
                                       const int64_t* __restrict__ index, float add) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nq) return;
  const long long krow = index ? index[row] : label_offset + row;
  float acc = 0.f;
  if (krow >= 0 && krow < nk) {
    for (int c = lane * 8; c < d; c += 256) {
      float a[8], b[8];
      load8<1>(q, row * d + c, a);
      load8<1>(k, krow * d + c, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(a[i], b[i], acc);
    }
  }
  __syncwarp();
  acc = warp_sum(acc);
  if (lane == 0) {
    float f = scale ? *scale : 1.f;
    if (qinv) f *= qinv[row];
    if (kinv && krow >= 0 && krow < nk) f *= kinv[krow];
    pos[row] = acc * f + add;
  }
}

__global__ void __launch_bounds__(512)
pair_loss_kernel(const float* __restrict__ lse_a, const float* __restrict__ lse_b,
                                 const float* __restrict__ esum_a, const float* __restrict__ esum_b,
                                 const float* __restrict__ pos, long long n, double coef,
                                 const float* __restrict__ scale, float* __restrict__ out2) {
  // one block (fixed summation order => bit-reproducible loss); 8 independent loads per array
  // in flight per thread
  __shared__ double sh[2][32];
  double l = 0.0, e = 0.0;
  constexpr int U = 8;
  for (long long base = 0; base < n; base += (long long)blockDim.x * U) {
    float la[U], lb[U], ea[U], eb[U], pp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + (long long)u * blockDim.x + threadIdx.x;
      const bool ok = i < n;
      la[u] = ok ? lse_a[i] : 0.f;
      lb[u] = ok ? lse_b[i] : 0.f;
      ea[u] = ok ? esum_a[i] : 0.f;
      eb[u] = ok ? esum_b[i] : 0.f;
      pp[u] = ok ? pos[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double p2 = 2.0 * (double)pp[u];
      l += (double)la[u] + (double)lb[u] - p2;
      e += (double)ea[u] + (double)eb[u] - p2;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][w] = l; sh[1][w] = e; }
  __syncthreads();
  if (w == 0) {
    l = lane < (blockDim.x >> 5) ? sh[0][lane] : 0.0;
    e = lane < (blockDim.x >> 5) ? sh[1][lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if (lane == 0) {
      out2[0] += (float)(coef * l);
      out2[1] += (float)(coef * e / (double)(*scale));
    }
  }
}

__global__ void make_offsets_kernel(const float* __restrict__ lse, long long n,
                                    const float* __restrict__ denom_vec, float denom_log2,
                                    float* __restrict__ off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float dl = denom_vec ? log2f(denom_vec[i]) : denom_log2;
  off[i] = fmaf(lse[i], LOG2E, dl);
}

// store 8 consecutive fp32 values as dtype ODT (0 f32, 1 bf16, 2 f16)
template <int ODT>
__device__ __forceinline__ void store8(void* base, long long elem, const float (&o)[8]) {
  if (ODT == 0) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + elem);
    p[0] = make_float4(o[0], o[1], o[2], o[3]);
    p[1] = make_float4(o[4], o[5], o[6], o[7]);
  } else if (ODT == 1) {
    uint4 u;
    u.x = bf16x2(o[0], o[1]);
    u.y = bf16x2(o[2], o[3]);
    u.z = bf16x2(o[4], o[5]);
    u.w = bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + elem) = u;
  } else {
    uint4 u;
    __half2 h;
    h = __floats2half2_rn(o[0], o[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(o[2], o[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(o[4], o[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(o[6], o[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(static_cast<__half*>(base) + elem) = u;
  }
}

// kDirect: idx == identity and this is the only contribution to the rows -> plain stores in the
// output dtype ODT; otherwise fp32 atomics into `grad` (ODT == 0).
// MAXV = ceil(d / 256) register-resident 8-wide column groups per lane (dispatch below): sizing the
// arrays to the actual d keeps the register count low enough for full occupancy at d = 512.
template <int DT, int ODT, bool kDirect, int MAXV>
__global__ void unpack_grad_kernel(const float* __restrict__ dacc,
                                   const __nv_bfloat16* __restrict__ partner,
                                   const float* __restrict__ partner_inv, long long n, int d,
                                   float weight, float pos_coef, const float* __restrict__ gout,
                                   const float* __restrict__ scale, const void* __restrict__ src,
                                   const float* __restrict__ inv_norm, int normalize,
                                   const int64_t* __restrict__ idx, void* __restrict__ grad) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float coef = weight * (gout ? *gout : 1.f) * (*scale);
  const long long srow = idx ? idx[row] : row;
  float g[MAXV][8];
  float x[MAXV][8];
  float dot = 0.f;
  const float inv = normalize ? inv_norm[row] : 1.f;
  const float pc = pos_coef * (partner_inv ? partner_inv[row] : 1.f);
  // Every lane walks every column group with a clamped address and a 0/1 mask instead of
  // branching on `c < d`: the warp stays converged up to the shuffle reduction below (a divergent
  // `if` in front of it gave lane-dependent sums for d % 256 not in {0, 64, 128}).
#pragma unroll
  for (int nv = 0; nv < MAXV; ++nv) {
    const int c = lane * 8 + nv * 256;
    const bool act = c < d;
    const int cc = act ? c : 0;
    float pa[8];
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(dacc + row * d + cc));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(dacc + row * d + cc + 4));
    load8<1>(partner, row * d + cc, pa);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) g[nv][i] = act ? coef * (a[i] - pc * pa[i]) : 0.f;
    if (normalize) {
      load8<DT>(src, srow * d + cc, x[nv]);
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(g[nv][i], x[nv][i] * inv, dot);
    }
  }
  __syncwarp();
  if (normalize) dot = warp_sum(dot);
  const bool clamped = inv >= 0.99e12f;  // ||x|| < eps: x_hat = x / eps, a plain scaling
#pragma unroll
  for (int nv = 0; nv < MAXV; ++nv) {
    const int c = lane * 8 + nv * 256;
    if (c < d) {
      float o[8];
      if (normalize) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = clamped ? g[nv][i] * inv : inv * (g[nv][i] - x[nv][i] * inv * dot);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = g[nv][i];
      }
      if (kDirect) {
        store8<ODT>(grad, srow * d + c, o);
      } else {
        float* dst = static_cast<float*>(grad) + srow * d + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dst + i, o[i]);
      }
    }
  }
}

}  // namespace

cudaError_t launch_pack_rows(const void* src, int src_dtype, long long src_rows, long long d,
                             const int64_t* idx, long long n_out, int normalize, void* dst_bf16,
                             float* inv_norm, cudaStream_t stream) {
  if (n_out == 0) return cudaSuccess;
  const int wpb = 8;
  const unsigned blocks = unsigned((n_out + wpb - 1) / wpb);
  auto* dst = static_cast<__nv_bfloat16*>(dst_bf16);
  if (src_dtype == 0)
    pack_rows_kernel<0><<<blocks, wpb * 32, 0, stream>>>(src, src_rows, int(d), idx, n_out, normalize, dst, inv_norm);
  else if (src_dtype == 1)
    pack_rows_kernel<1><<<blocks, wpb * 32, 0, stream>>>(src, src_rows, int(d), idx, n_out, normalize, dst, inv_norm);
  else
    pack_rows_kernel<2><<<blocks, wpb * 32, 0, stream>>>(src, src_rows, int(d), idx, n_out, normalize, dst, inv_norm);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_positive_logits(const void* q, long long nq, const void* k, long long nk,
                                   long long d, long long label_offset, const float* scale,
                                   const float* qinv, const float* kinv, float* pos,
                                   cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  const int wpb = 8;
  positive_logits_kernel<<<unsigned((nq + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(q), nq, static_cast<const __nv_bfloat16*>(k), nk, int(d),
      label_offset, scale, qinv, kinv, pos, nullptr, 0.f);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_logit_scale(float* log_s, long long n, double max_scale, float* s_out,
                               cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  logit_scale_kernel<<<unsigned((n + 127) / 128), 128, 0, stream>>>(log_s, n, float(log(max_scale)),
                                                                   s_out);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_indexed_similarity(const void* q, long long nq, const void* k, long long nk,
                                      long long d, const int64_t* index, const float* qinv,
                                      const float* kinv, float add, float* out,
                                      cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  const int wpb = 8;
  positive_logits_kernel<<<unsigned((nq + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(q), nq, static_cast<const __nv_bfloat16*>(k), nk, int(d), 0,
      nullptr, qinv, kinv, out, index, add);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_pair_loss(const float* lse_a, const float* lse_b, const float* esum_a,
                             const float* esum_b, const float* pos, long long n, double weight,
                             double n_norm, const float* scale, float* out2, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  pair_loss_kernel<<<1, 512, 0, stream>>>(lse_a, lse_b, esum_a, esum_b, pos, n,
                                           weight / (2.0 * n_norm), scale, out2);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_make_offsets(const float* lse, long long n, const float* denom_vec,
                                double denom_scalar, float* off, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  make_offsets_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(
      lse, n, denom_vec, denom_vec ? 0.f : (float)log2(denom_scalar), off);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_unpack_grad(const float* dacc, const void* partner_bf16,
                               const float* partner_inv, long long n,
                               long long d, double weight, double pos_coef, const float* gout,
                               const float* scale, const void* src, int src_dtype,
                               const float* inv_norm, int normalize, const int64_t* idx,
                               void* grad, int grad_dtype, int direct, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int wpb = 4;
  const unsigned blocks = unsigned((n + wpb - 1) / wpb);
  auto* pb = static_cast<const __nv_bfloat16*>(partner_bf16);
  const int nv = int((d + 255) / 256);  // d <= 2048 -> 1..8
#define MMCL_UNPACK_NV(DT, ODT, DIRECT, NV)                                                       \
  unpack_grad_kernel<DT, ODT, DIRECT, NV><<<blocks, wpb * 32, 0, stream>>>(                       \
      dacc, pb, partner_inv, n, int(d), (float)weight, (float)pos_coef, gout, scale, src,         \
      inv_norm, normalize, idx, grad)
#define MMCL_UNPACK(DT, ODT, DIRECT)                                                              \
  do {                                                                                            \
    if (nv <= 1) MMCL_UNPACK_NV(DT, ODT, DIRECT, 1);                                              \
    else if (nv == 2) MMCL_UNPACK_NV(DT, ODT, DIRECT, 2);                                         \
    else if (nv == 3) MMCL_UNPACK_NV(DT, ODT, DIRECT, 3);                                         \
    else if (nv == 4) MMCL_UNPACK_NV(DT, ODT, DIRECT, 4);                                         \
    else MMCL_UNPACK_NV(DT, ODT, DIRECT, 8);                                                      \
  } while (0)
#define MMCL_UNPACK_SRC(ODT, DIRECT)                                                              \
  do {                                                                                            \
    if (src_dtype == 0) MMCL_UNPACK(0, ODT, DIRECT);                                              \
    else if (src_dtype == 1) MMCL_UNPACK(1, ODT, DIRECT);                                         \
    else MMCL_UNPACK(2, ODT, DIRECT);                                                             \
  } while (0)
  if (!direct) {
    MMCL_UNPACK_SRC(0, false);
  } else if (grad_dtype == 0) {
    MMCL_UNPACK_SRC(0, true);
  } else if (grad_dtype == 1) {
    MMCL_UNPACK_SRC(1, true);
  } else {
    MMCL_UNPACK_SRC(2, true);
  }
#undef MMCL_UNPACK_SRC
#undef MMCL_UNPACK
#undef MMCL_UNPACK_NV
  count_launches(1);
  return cudaGetLastError();
}

}  // namespace mmcl
