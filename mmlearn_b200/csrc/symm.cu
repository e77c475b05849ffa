// Symmetric peer-memory arena: the all-gathers of the loss as direct NVLink stores.
//
// Every rank owns one device allocation  [flags: 1 KB][buffer 0][buffer 1]  and maps the
// allocations of all other ranks of the node through CUDA IPC.  A "gather" is then a kernel that
// writes the rank's contribution to the same offset of EVERY rank's buffer (NVSwitch gives each
// GPU full bandwidth to each peer), followed by a fence kernel: flag[my_rank] = seq is stored
// (release, system scope) into every peer's flag array and the kernel spins until all of its own
// flags reached seq.  Work enqueued after the fence on the same stream sees every rank's data.
// No NCCL kernel, no host round trip, one small kernel per fence.
//
// The two buffers alternate between consecutive uses (parity): a rank that has passed fence i
// knows that every peer has finished the stream work it enqueued before its own fence i, which
// includes all reads of the buffer written two uses ago (the caller guarantees that ordering, see
// mmlearn_b200/loss.py).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace mmcl {

namespace {

// flag words: fused.h (kFlagBytes, kMaxRanks; channels 0..kFenceChannels-1 are fence channels)

template <typename V>  // uint4 (16-byte units) or uint32_t (4-byte units)
__global__ void symm_put_kernel(char* const* __restrict__ peers, int world, size_t dst_off,
                                const V* __restrict__ src, size_t n, int skip_rank) {
  const size_t total = n * size_t(world);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += size_t(gridDim.x) * blockDim.x) {
    const int r = int(i / n);
    if (r == skip_rank) continue;  // (the source already is that rank's copy)
    const size_t e = i % n;
    reinterpret_cast<V*>(peers[r] + dst_off)[e] = src[e];
  }
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// one block, one thread per rank
// `slot0` = first flag word of the fence channel (kMaxRanks words per channel): fences of
// different channels may be in flight at the same time on different streams.
__global__ void symm_fence_kernel(char* const* __restrict__ peers, int world, int rank,
                                  unsigned long long seq, int slot0) {
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();  // everything this stream wrote to peer memory before the fence
    st_release_sys(reinterpret_cast<unsigned long long*>(peers[r]) + slot0 + rank, seq);
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(peers[rank]) + slot0 + r;
    while (ld_acquire_sys(mine) < seq) {
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

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
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
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

// pack_rows (aux_kernels.cu) with the store fanned out to every rank's buffer
template <int DT>
__global__ void pack_rows_symm_kernel(const void* __restrict__ src, long long src_rows, int d,
                                      const int64_t* __restrict__ idx, long long n_out,
                                      int normalize, char* const* __restrict__ peers, int world,
                                      size_t off_rows, long long off_inv) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_out) return;
  long long srow = idx ? idx[row] : row;
  if (srow < 0 || srow >= src_rows) srow = 0;
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
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  for (int c = lane * 8; c < d; c += 256) {
    float f[8];
    load8<DT>(src, srow * d + c, f);
    uint4 o;
    o.x = bf16x2(f[0], f[1]);
    o.y = bf16x2(f[2], f[3]);
    o.z = bf16x2(f[4], f[5]);
    o.w = bf16x2(f[6], f[7]);
    for (int r = 0; r < world; ++r)
      *reinterpret_cast<uint4*>(peers[r] + off_rows + (size_t(row) * d + c) * 2) = o;
  }
  if (off_inv >= 0 && lane < world)
    reinterpret_cast<float*>(peers[lane] + off_inv)[row] = inv;
}

}  // namespace

int symm_reserve(SymmArena* a, void* comm_v, int rank, int world, size_t bytes,
                 cudaStream_t stream, std::string* err) {
  auto comm = static_cast<ncclComm_t>(comm_v);
  bytes = (bytes + 4095) & ~size_t(4095);
  if (a->local && bytes <= a->bytes) return 0;
  if (a->failed) {
    *err = "symmetric arena unavailable (an earlier CUDA IPC setup failed)";
    return -1;
  }
  // Every rank walks through the same collectives whatever happens locally; local failures are
  // recorded in `ok` and agreed on at the end, so that all ranks fall back together.
  bool ok = true;
  auto cu = [&](cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    if (ok) *err = std::string(what) + ": " + cudaGetErrorString(e);
    ok = false;
    cudaGetLastError();
    return false;
  };
  auto all_sum = [&](float mine) -> float {  // also a barrier: every rank drained `stream`
    float v = mine;
    float* d = static_cast<float*>(a->scratch);
    cu(cudaMemcpyAsync(d, &v, 4, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync");
    ncclResult_t r = ncclAllReduce(d, d, 1, ncclFloat, ncclSum, comm, stream);
    if (r != ncclSuccess && ok) {
      *err = std::string("ncclAllReduce: ") + ncclGetErrorString(r);
      ok = false;
    }
    cu(cudaMemcpyAsync(&v, d, 4, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    cu(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    return v;
  };
  if (!a->scratch) {
    if (cudaMalloc(&a->scratch, 64 + 64 * size_t(world)) != cudaSuccess) {
      // without the staging buffer no collective is possible; this can only fail everywhere
      *err = "cudaMalloc(scratch) failed";
      a->failed = true;
      return -1;
    }
  }
  all_sum(0.f);  // nobody still works on the old arena
  if (a->local) {  // grow path: unmap, agree, free
    for (int r = 0; r < world; ++r)
      if (r != rank && a->peers[r]) cudaIpcCloseMemHandle(a->peers[r]);
    all_sum(0.f);
    cudaFree(a->local);
    a->local = nullptr;
    a->peers.clear();
  }
  if (bytes < 2 * a->bytes) bytes = 2 * a->bytes;  // geometric growth: reserve is collective
  char* base = nullptr;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (cu(cudaMalloc(&base, kFlagBytes + 2 * bytes), "cudaMalloc(arena)")) {
    cu(cudaMemsetAsync(base, 0, kFlagBytes, stream), "cudaMemsetAsync");
    cu(cudaIpcGetMemHandle(&mine, base), "cudaIpcGetMemHandle");
  }
  char* send = static_cast<char*>(a->scratch);
  char* recv = send + 64;
  std::vector<cudaIpcMemHandle_t> all(world);
  cu(cudaMemcpyAsync(send, &mine, 64, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync");
  ncclResult_t nr = ncclAllGather(send, recv, 64, ncclChar, comm, stream);
  if (nr != ncclSuccess && ok) {
    *err = std::string("ncclAllGather(ipc handles): ") + ncclGetErrorString(nr);
    ok = false;
  }
  cu(cudaMemcpyAsync(all.data(), recv, 64 * size_t(world), cudaMemcpyDeviceToHost, stream),
     "cudaMemcpyAsync");
  cu(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  const bool everyone_allocated = all_sum(ok ? 1.f : 0.f) > float(world) - 0.5f;
  std::vector<char*> peers(world, nullptr);
  if (everyone_allocated) {
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) {
        peers[r] = base;
        continue;
      }
      void* p = nullptr;
      if (cu(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess),
             "cudaIpcOpenMemHandle"))
        peers[r] = static_cast<char*>(p);
    }
    if (ok && !a->d_peers) cu(cudaMalloc(&a->d_peers, sizeof(char*) * size_t(world)), "cudaMalloc");
    if (ok)
      cu(cudaMemcpyAsync(a->d_peers, peers.data(), sizeof(char*) * size_t(world),
                         cudaMemcpyHostToDevice, stream),
         "cudaMemcpyAsync");
  }
  // every rank has mapped every arena and zeroed its flags -- or nobody uses the arena at all
  const bool everyone_ok = all_sum(ok ? 1.f : 0.f) > float(world) - 0.5f;
  if (!everyone_ok) {
    for (int r = 0; r < world; ++r)
      if (r != rank && peers[r]) cudaIpcCloseMemHandle(peers[r]);
    all_sum(0.f);
    if (base) cudaFree(base);
    a->failed = true;
    a->bytes = 0;
    if (err->empty()) *err = "CUDA IPC setup failed on another rank";
    return -1;
  }
  a->peers = peers;
  a->local = base;
  a->bytes = bytes;
  for (auto& q : a->seq) q = 0;
  a->step_seq = 0;
  return 0;
}

void symm_destroy(SymmArena* a, int rank) {
  if (a->local) {
    for (size_t r = 0; r < a->peers.size(); ++r)
      if (int(r) != rank && a->peers[r]) cudaIpcCloseMemHandle(a->peers[r]);
    cudaFree(a->local);
  }
  if (a->d_peers) cudaFree(a->d_peers);
  if (a->scratch) cudaFree(a->scratch);
  *a = SymmArena();
}

char* symm_buffer(const SymmArena* a, int parity) {
  return a->local + kFlagBytes + size_t(parity & 1) * a->bytes;
}

cudaError_t launch_symm_put(const SymmArena* a, int world, int parity, const void* src,
                            size_t bytes, size_t dst_off, cudaStream_t stream, int skip_rank) {
  if (bytes == 0) return cudaSuccess;
  const size_t off = kFlagBytes + size_t(parity & 1) * a->bytes + dst_off;
  const bool v16 = bytes % 16 == 0 && dst_off % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const size_t n = bytes / (v16 ? 16 : 4);
  const size_t total = n * size_t(world);
  unsigned blocks = unsigned((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  if (v16)
    symm_put_kernel<uint4><<<blocks, 256, 0, stream>>>(a->d_peers, world, off,
                                                       static_cast<const uint4*>(src), n, skip_rank);
  else
    symm_put_kernel<uint32_t><<<blocks, 256, 0, stream>>>(a->d_peers, world, off,
                                                          static_cast<const uint32_t*>(src), n,
                                                          skip_rank);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_pack_rows_symm(const SymmArena* a, int world, int parity, const void* src,
                                  int src_dtype, long long src_rows, long long d,
                                  const int64_t* idx, long long n_out, int normalize,
                                  size_t off_rows, long long off_inv, cudaStream_t stream) {
  if (n_out == 0) return cudaSuccess;
  const int wpb = 8;
  const unsigned blocks = unsigned((n_out + wpb - 1) / wpb);
  const size_t base = kFlagBytes + size_t(parity & 1) * a->bytes;
  const long long oi = off_inv >= 0 ? (long long)(base + size_t(off_inv)) : -1;
  if (src_dtype == 0)
    pack_rows_symm_kernel<0><<<blocks, wpb * 32, 0, stream>>>(src, src_rows, int(d), idx, n_out, normalize, a->d_peers, world, base + off_rows, oi);
  else if (src_dtype == 1)
    pack_rows_symm_kernel<1><<<blocks, wpb * 32, 0, stream>>>(src, src_rows, int(d), idx, n_out, normalize, a->d_peers, world, base + off_rows, oi);
  else
    pack_rows_symm_kernel<2><<<blocks, wpb * 32, 0, stream>>>(src, src_rows, int(d), idx, n_out, normalize, a->d_peers, world, base + off_rows, oi);
  count_launches(1);
  return cudaGetLastError();
}

// Diagnostic (tools/push_bench.py): the same bytes pushed three ways, to pick the transport.
//   mode 0: store kernel with a caller-chosen grid (how many threads does NVLink need?)
//   mode 1: one cudaMemcpyAsync per peer on `stream` (copy engines, serial)
//   mode 2: one cudaMemcpyAsync per peer on per-peer side streams (copy engines, parallel)
cudaError_t debug_push(SymmArena* a, int world, int rank, int parity, size_t off, size_t bytes,
                       int mode, int blocks, int threads, cudaStream_t stream) {
  const size_t base = kFlagBytes + size_t(parity & 1) * a->bytes + off;
  if (mode == 0) {
    symm_put_kernel<uint4><<<blocks, threads, 0, stream>>>(
        a->d_peers, world, base, reinterpret_cast<const uint4*>(a->local + base), bytes / 16, rank);
    count_launches(1);
    return cudaGetLastError();
  }
  static cudaStream_t side[kMaxRanks] = {};
  static cudaEvent_t ev_fork = nullptr, ev_join[kMaxRanks] = {};
  cudaError_t e = cudaSuccess;
  if (mode == 2) {
    if (!ev_fork) e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev_fork, stream);
  }
  for (int r = 0; r < world && e == cudaSuccess; ++r) {
    if (r == rank) continue;
    cudaStream_t st = stream;
    if (mode == 2) {
      if (!side[r]) {
        e = cudaStreamCreateWithFlags(&side[r], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_join[r], cudaEventDisableTiming);
        if (e != cudaSuccess) break;
      }
      st = side[r];
      e = cudaStreamWaitEvent(st, ev_fork, 0);
      if (e != cudaSuccess) break;
    }
    e = cudaMemcpyAsync(a->peers[r] + base, a->local + base, bytes, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && mode == 2) {
      e = cudaEventRecord(ev_join[r], st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, ev_join[r], 0);
    }
  }
  return e;
}

cudaError_t launch_symm_fence(SymmArena* a, int world, int rank, cudaStream_t stream,
                              int channel) {
  if (channel < 0 || channel >= kFenceChannels || world > kMaxRanks) return cudaErrorInvalidValue;
  a->seq[channel] += 1;
  symm_fence_kernel<<<1, 32 * ((world + 31) / 32), 0, stream>>>(a->d_peers, world, rank,
                                                                a->seq[channel], channel * kMaxRanks);
  count_launches(1);
  return cudaGetLastError();
}

}  // namespace mmcl
