// C ABI (include/mmcl.h): handle, error reporting, NCCL collectives, and thin argument-checking
// wrappers around the kernel launchers.  No torch types, no device allocation, no device sync.
#include <cuda_runtime.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <atomic>

#include "handle.h"

using namespace mmcl;

namespace mmcl {
std::string& last_error_slot() {
  thread_local std::string g_err;
  return g_err;
}
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace mmcl

extern "C" {

int mmcl_version(void) { return MMCL_VERSION; }

int64_t mmcl_kernel_launches(void) { return launches_so_far(); }

const char* mmcl_last_error(void) { return last_error_slot().c_str(); }

int mmcl_nccl_unique_id(void* uid) {
  REQUIRE(uid, "mmcl_nccl_unique_id: null uid");
  static_assert(sizeof(ncclUniqueId) == MMCL_NCCL_UID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  ncclResult_t r = ncclGetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(uid, &id, sizeof(id));
  return MMCL_OK;
}

int mmcl_create(mmcl_handle** out, int device, int rank, int world, const void* nccl_uid) {
  REQUIRE(out, "mmcl_create: null out");
  REQUIRE(world >= 1 && rank >= 0 && rank < world, "mmcl_create: bad rank/world");
  REQUIRE(world == 1 || nccl_uid, "mmcl_create: world > 1 needs an NCCL unique id");
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(MMCL_ERR_ARCH, "mmcl_create: device is sm_" + std::to_string(prop.major) +
                                   std::to_string(prop.minor) +
                                   "; this library only contains sm_100a code");
  auto* h = new mmcl_handle();
  h->device = device;
  h->rank = rank;
  h->world = world;
  h->num_sms = prop.multiProcessorCount;
  if (world > 1) {
    DeviceGuard g(device);
    ncclUniqueId id;
    memcpy(&id, nccl_uid, sizeof(id));
    ncclResult_t r = ncclCommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) {
      delete h;
      return nccl_fail(r, "ncclCommInitRank");
    }
  }
  *out = h;
  return MMCL_OK;
}

int mmcl_destroy(mmcl_handle* h) {
  if (!h) return MMCL_OK;
  {
    DeviceGuard g(h->device);
    symm_destroy(&h->arena, h->rank);
  }
  if (h->comm) ncclCommDestroy(h->comm);
  delete h;
  return MMCL_OK;
}

size_t mmcl_match_workspace_bytes(int64_t n, int64_t m) { return match_workspace_bytes(n, m); }

int mmcl_match_ids(mmcl_handle* h, const int64_t* ids_a, int64_t n, const int64_t* ids_b,
                   int64_t m, int64_t* out_a, int64_t* out_b, int64_t capacity, int64_t* out_info,
                   void* workspace, size_t workspace_bytes, mmcl_stream stream) {
  REQUIRE(h && out_info, "mmcl_match_ids: null handle/out_info");
  REQUIRE(n >= 0 && m >= 0 && capacity >= 0, "mmcl_match_ids: negative size");
  REQUIRE(n < (1ll << 31) && m < (1ll << 31), "mmcl_match_ids: more than 2^31 rows");
  REQUIRE((n == 0 || ids_a) && (m == 0 || ids_b), "mmcl_match_ids: null ids");
  REQUIRE(capacity == 0 || (out_a && out_b), "mmcl_match_ids: null outputs");
  if (n > 0 && m > 0 && workspace_bytes < match_workspace_bytes(n, m))
    return fail(MMCL_ERR_WORKSPACE, "mmcl_match_ids: workspace too small");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_match_ids(ids_a, n, ids_b, m, out_a, out_b, capacity, out_info, workspace,
                                   workspace_bytes, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_match_ids");
}

size_t mmcl_identity_check_workspace_bytes(int64_t m) {
  return identity_check_workspace_bytes(m);
}

int mmcl_identity_check(mmcl_handle* h, const int64_t* ids_a, int64_t n, const int64_t* ids_b,
                        int64_t m, int64_t* out_info, void* workspace, size_t workspace_bytes,
                        mmcl_stream stream) {
  REQUIRE(h && out_info, "mmcl_identity_check: null handle/out_info");
  REQUIRE(n >= 0 && m >= 0 && n < (1ll << 31) && m < (1ll << 31), "mmcl_identity_check: bad size");
  REQUIRE((n == 0 || ids_a) && (m == 0 || ids_b), "mmcl_identity_check: null ids");
  if (n == m && n > 0 && (workspace_bytes < identity_check_workspace_bytes(m) || !workspace))
    return fail(MMCL_ERR_WORKSPACE, "mmcl_identity_check: workspace too small");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_identity_check(ids_a, n, ids_b, m, out_info, workspace, workspace_bytes,
                                        static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_identity_check");
}

int mmcl_pack_rows(mmcl_handle* h, const void* src, int src_dtype, int64_t src_rows, int64_t d,
                   const int64_t* idx, int64_t n_out, int normalize, void* dst_bf16,
                   float* inv_norm, mmcl_stream stream) {
  REQUIRE(h, "mmcl_pack_rows: null handle");
  REQUIRE(src_dtype >= 0 && src_dtype <= 2, "mmcl_pack_rows: dtype must be f32/bf16/f16");
  REQUIRE(d > 0 && d % 8 == 0, "mmcl_pack_rows: d must be a positive multiple of 8");
  REQUIRE(n_out >= 0 && src_rows >= 0, "mmcl_pack_rows: negative size");
  REQUIRE(n_out == 0 || (src && dst_bf16), "mmcl_pack_rows: null buffer");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_pack_rows(src, src_dtype, src_rows, d, idx, n_out, normalize, dst_bf16,
                                   inv_norm, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_pack_rows");
}

int mmcl_group_start(mmcl_handle* h) {
  REQUIRE(h, "mmcl_group_start: null handle");
  if (h->world == 1) return MMCL_OK;
  ncclResult_t r = ncclGroupStart();
  return r == ncclSuccess ? MMCL_OK : nccl_fail(r, "ncclGroupStart");
}

int mmcl_group_end(mmcl_handle* h) {
  REQUIRE(h, "mmcl_group_end: null handle");
  if (h->world == 1) return MMCL_OK;
  ncclResult_t r = ncclGroupEnd();
  return r == ncclSuccess ? MMCL_OK : nccl_fail(r, "ncclGroupEnd");
}

int mmcl_symm_reserve(mmcl_handle* h, size_t bytes_per_buffer, mmcl_stream stream) {
  REQUIRE(h, "mmcl_symm_reserve: null handle");
  REQUIRE(h->world > 1 && h->world <= 32, "mmcl_symm_reserve: needs 2..32 ranks");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  std::string err;
  if (symm_reserve(&h->arena, h->comm, h->rank, h->world, bytes_per_buffer,
                   static_cast<cudaStream_t>(stream), &err))
    return fail(MMCL_ERR_CUDA, "mmcl_symm_reserve: " + err);
  return MMCL_OK;
}

int mmcl_symm_buffer(mmcl_handle* h, int parity, void** out, size_t* bytes) {
  REQUIRE(h && out, "mmcl_symm_buffer: null argument");
  REQUIRE(h->arena.local, "mmcl_symm_buffer: no arena reserved");
  *out = symm_buffer(&h->arena, parity);
  if (bytes) *bytes = h->arena.bytes;
  return MMCL_OK;
}

int mmcl_symm_put(mmcl_handle* h, int parity, const void* src, size_t bytes, size_t dst_off,
                  mmcl_stream stream) {
  REQUIRE(h && h->arena.local, "mmcl_symm_put: no arena reserved");
  REQUIRE(bytes % 4 == 0 && dst_off % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0,
          "mmcl_symm_put: size, offset and source must be multiples of 4 bytes");
  REQUIRE(dst_off + bytes <= h->arena.bytes, "mmcl_symm_put: beyond the reserved buffer");
  REQUIRE(bytes == 0 || src, "mmcl_symm_put: null source");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_symm_put(&h->arena, h->world, parity, src, bytes, dst_off,
                                  static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_symm_put");
}

int mmcl_pack_rows_symm(mmcl_handle* h, int parity, const void* src, int src_dtype,
                        int64_t src_rows, int64_t d, const int64_t* idx, int64_t n_out,
                        int normalize, size_t dst_off_rows, int64_t dst_off_inv,
                        mmcl_stream stream) {
  REQUIRE(h && h->arena.local, "mmcl_pack_rows_symm: no arena reserved");
  REQUIRE(src_dtype >= 0 && src_dtype <= 2, "mmcl_pack_rows_symm: dtype must be f32/bf16/f16");
  REQUIRE(d > 0 && d % 8 == 0, "mmcl_pack_rows_symm: d must be a positive multiple of 8");
  REQUIRE(n_out >= 0 && src_rows >= 0, "mmcl_pack_rows_symm: negative size");
  REQUIRE(n_out == 0 || src, "mmcl_pack_rows_symm: null source");
  REQUIRE(dst_off_rows % 16 == 0 && dst_off_rows + size_t(n_out) * d * 2 <= h->arena.bytes,
          "mmcl_pack_rows_symm: rows beyond the reserved buffer");
  REQUIRE(dst_off_inv < 0 || (dst_off_inv % 4 == 0 &&
                              size_t(dst_off_inv) + size_t(n_out) * 4 <= h->arena.bytes),
          "mmcl_pack_rows_symm: inverse norms beyond the reserved buffer");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_pack_rows_symm(&h->arena, h->world, parity, src, src_dtype, src_rows, d,
                                        idx, n_out, normalize, dst_off_rows, dst_off_inv,
                                        static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_pack_rows_symm");
}

int mmcl_debug_push(mmcl_handle* h, int parity, size_t off, size_t bytes, int mode, int blocks,
                    int threads, mmcl_stream stream) {
  REQUIRE(h && h->world > 1 && h->arena.local, "mmcl_debug_push: no arena");
  REQUIRE(bytes % 16 == 0 && off % 16 == 0 && off + bytes <= h->arena.bytes,
          "mmcl_debug_push: bad range");
  REQUIRE(mode >= 0 && mode <= 2 && (mode != 0 || (blocks > 0 && threads > 0 && threads <= 1024)),
          "mmcl_debug_push: bad mode / grid");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = debug_push(&h->arena, h->world, h->rank, parity, off, bytes, mode, blocks, threads,
                             static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_debug_push");
}

int mmcl_symm_fence(mmcl_handle* h, mmcl_stream stream) {
  REQUIRE(h && h->arena.local, "mmcl_symm_fence: no arena reserved");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_symm_fence(&h->arena, h->world, h->rank,
                                    static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_symm_fence");
}

int mmcl_all_gather_rows(mmcl_handle* h, const void* send, void* recv, const int64_t* counts,
                         int64_t row_bytes, mmcl_stream stream) {
  REQUIRE(h && counts && row_bytes > 0, "mmcl_all_gather_rows: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  auto st = static_cast<cudaStream_t>(stream);
  if (h->world == 1) {
    if (counts[0] > 0 && send != recv) {
      cudaError_t e = cudaMemcpyAsync(recv, send, size_t(counts[0]) * row_bytes,
                                      cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return cuda_fail(e, "mmcl_all_gather_rows");
    }
    return MMCL_OK;
  }
  bool even = true;
  for (int r = 1; r < h->world; ++r) even = even && counts[r] == counts[0];
  ncclResult_t r;
  if (even) {
    r = ncclAllGather(send, recv, size_t(counts[0]) * row_bytes, ncclChar, h->comm, st);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
    return MMCL_OK;
  }
  r = ncclGroupStart();
  if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
  size_t off = 0;
  for (int root = 0; root < h->world; ++root) {
    const size_t bytes = size_t(counts[root]) * row_bytes;
    if (bytes > 0) {
      r = ncclBroadcast(root == h->rank ? send : nullptr, static_cast<char*>(recv) + off, bytes,
                        ncclChar, root, h->comm, st);
      if (r != ncclSuccess) {
        ncclGroupEnd();
        return nccl_fail(r, "ncclBroadcast");
      }
    }
    off += bytes;
  }
  r = ncclGroupEnd();
  if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
  return MMCL_OK;
}

int mmcl_all_reduce_sum_f32(mmcl_handle* h, float* buf, int64_t n, mmcl_stream stream) {
  REQUIRE(h && (n == 0 || buf), "mmcl_all_reduce_sum_f32: bad arguments");
  if (h->world == 1 || n == 0) return MMCL_OK;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  ncclResult_t r = ncclAllReduce(buf, buf, size_t(n), ncclFloat, ncclSum, h->comm,
                                 static_cast<cudaStream_t>(stream));
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  return MMCL_OK;
}

size_t mmcl_sim_lse_workspace_bytes(mmcl_handle* h, int64_t nq, int64_t nk, int64_t d) {
  if (!h || nq <= 0 || nk <= 0 || d <= 0) return 0;
  return sim_lse_workspace_bytes(h->num_sms, nq, nk, d);
}

static int check_sim_args(const char* fn, const void* a, int64_t na, const void* b, int64_t nb,
                          int64_t d) {
  if (na < 0 || nb < 0) return fail(MMCL_ERR_INVALID, std::string(fn) + ": negative size");
  if (d <= 0 || d % 8 != 0 || d > 2048)
    return fail(MMCL_ERR_INVALID, std::string(fn) + ": d must be a multiple of 8 in [8, 2048]");
  if (na >= (1ll << 31) || nb >= (1ll << 31))
    return fail(MMCL_ERR_INVALID, std::string(fn) + ": more than 2^31 rows");
  if ((na > 0 && !a) || (nb > 0 && !b)) return fail(MMCL_ERR_INVALID, std::string(fn) + ": null matrix");
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
    return fail(MMCL_ERR_INVALID, std::string(fn) + ": matrices must be 16-byte aligned");
  return MMCL_OK;
}

int mmcl_sim_lse_fwd(mmcl_handle* h, const void* q_bf16, int64_t nq, const void* k_bf16,
                     int64_t nk, int64_t d, const float* scale, const float* qinv,
                     const float* kinv, float* lse, float* esum, void* workspace,
                     size_t workspace_bytes, mmcl_stream stream) {
  REQUIRE(h && scale, "mmcl_sim_lse_fwd: null handle/scale");
  int rc = check_sim_args("mmcl_sim_lse_fwd", q_bf16, nq, k_bf16, nk, d);
  if (rc) return rc;
  if (nq == 0) return MMCL_OK;
  REQUIRE(nk > 0, "mmcl_sim_lse_fwd: nk must be positive (log-sum-exp over an empty set)");
  REQUIRE(lse, "mmcl_sim_lse_fwd: null lse");
  if (workspace_bytes < sim_lse_workspace_bytes(h->num_sms, nq, nk, d) || !workspace)
    return fail(MMCL_ERR_WORKSPACE, "mmcl_sim_lse_fwd: workspace too small");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const char* err = nullptr;
  cudaError_t e = launch_sim_lse_fwd(h->num_sms, q_bf16, nq, k_bf16, nk, d, scale, qinv, kinv, lse,
                                     esum, workspace, workspace_bytes,
                                     static_cast<cudaStream_t>(stream), &err);
  if (e != cudaSuccess) return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_sim_lse_fwd");
  return MMCL_OK;
}

int mmcl_positive_logits(mmcl_handle* h, const void* q_bf16, int64_t nq, const void* k_bf16,
                         int64_t nk, int64_t d, int64_t label_offset, const float* scale,
                         const float* qinv, const float* kinv, float* pos, mmcl_stream stream) {
  REQUIRE(h && scale, "mmcl_positive_logits: null handle/scale");
  int rc = check_sim_args("mmcl_positive_logits", q_bf16, nq, k_bf16, nk, d);
  if (rc) return rc;
  REQUIRE(nq == 0 || pos, "mmcl_positive_logits: null pos");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_positive_logits(q_bf16, nq, k_bf16, nk, d, label_offset, scale, qinv,
                                         kinv, pos, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_positive_logits");
}

int mmcl_logit_scale(mmcl_handle* h, float* log_scale, int64_t n, double max_scale,
                     float* scale_out, mmcl_stream stream) {
  REQUIRE(h, "mmcl_logit_scale: null handle");
  REQUIRE(n >= 0 && (n == 0 || (log_scale && scale_out)), "mmcl_logit_scale: bad arguments");
  REQUIRE(max_scale >= 1.0, "mmcl_logit_scale: max_scale must be >= 1");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_logit_scale(log_scale, n, max_scale, scale_out,
                                     static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_logit_scale");
}

int mmcl_indexed_similarity(mmcl_handle* h, const void* q_bf16, int64_t nq, const void* k_bf16,
                            int64_t nk, int64_t d, const int64_t* index, const float* qinv,
                            const float* kinv, double add, float* out, mmcl_stream stream) {
  REQUIRE(h, "mmcl_indexed_similarity: null handle");
  int rc = check_sim_args("mmcl_indexed_similarity", q_bf16, nq, k_bf16, nk, d);
  if (rc) return rc;
  REQUIRE(nq == 0 || (out && index), "mmcl_indexed_similarity: null index/out");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_indexed_similarity(q_bf16, nq, k_bf16, nk, d, index, qinv, kinv,
                                            float(add), out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_indexed_similarity");
}

int mmcl_sim_rank_fwd(mmcl_handle* h, const void* q_bf16, int64_t nq, const void* k_bf16,
                      int64_t nk, int64_t d, const float* qinv, const float* kinv,
                      const float* thr, const int64_t* skip, int64_t* rank, void* workspace,
                      size_t workspace_bytes, mmcl_stream stream) {
  REQUIRE(h, "mmcl_sim_rank_fwd: null handle");
  int rc = check_sim_args("mmcl_sim_rank_fwd", q_bf16, nq, k_bf16, nk, d);
  if (rc) return rc;
  if (nq == 0) return MMCL_OK;
  REQUIRE(nk > 0, "mmcl_sim_rank_fwd: nk must be positive");
  REQUIRE(thr && rank, "mmcl_sim_rank_fwd: null thr/rank");
  if (workspace_bytes < sim_lse_workspace_bytes(h->num_sms, nq, nk, d) || !workspace)
    return fail(MMCL_ERR_WORKSPACE, "mmcl_sim_rank_fwd: workspace too small");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const char* err = nullptr;
  cudaError_t e = launch_sim_rank_fwd(h->num_sms, q_bf16, nq, k_bf16, nk, d, qinv, kinv, thr,
                                      reinterpret_cast<const long long*>(skip),
                                      reinterpret_cast<long long*>(rank), workspace,
                                      workspace_bytes, static_cast<cudaStream_t>(stream), &err);
  if (e != cudaSuccess) return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_sim_rank_fwd");
  return MMCL_OK;
}

int mmcl_sim_logits_fwd(mmcl_handle* h, const void* q_bf16, int64_t nq, const void* k_bf16,
                        int64_t nk, int64_t d, double scale, const float* qinv, const float* kinv,
                        float* out, int64_t out_ld, mmcl_stream stream) {
  REQUIRE(h, "mmcl_sim_logits_fwd: null handle");
  int rc = check_sim_args("mmcl_sim_logits_fwd", q_bf16, nq, k_bf16, nk, d);
  if (rc) return rc;
  if (nq == 0 || nk == 0) return MMCL_OK;
  REQUIRE(out && out_ld >= nk, "mmcl_sim_logits_fwd: null output / row pitch smaller than nk");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const char* err = nullptr;
  cudaError_t e = launch_sim_logits_fwd(h->num_sms, q_bf16, nq, k_bf16, nk, d, float(scale), qinv,
                                        kinv, out, out_ld, static_cast<cudaStream_t>(stream), &err);
  if (e != cudaSuccess) return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_sim_logits_fwd");
  return MMCL_OK;
}

size_t mmcl_sim_bce_workspace_bytes(mmcl_handle* h, int64_t nq, int64_t nk, int64_t d) {
  if (!h || nq <= 0 || nk <= 0 || d <= 0) return 0;
  return sim_bce_workspace_bytes(h->num_sms, nq, nk, d);
}

int mmcl_sim_bce_fwd(mmcl_handle* h, const void* q_bf16, int64_t nq, const void* k_bf16, int64_t nk,
                     int64_t d, const float* scale, const float* qinv, const float* kinv,
                     int64_t row_base, const int32_t* band_hi, float* stats, void* workspace,
                     size_t workspace_bytes, mmcl_stream stream) {
  REQUIRE(h && scale, "mmcl_sim_bce_fwd: null handle/scale");
  int rc = check_sim_args("mmcl_sim_bce_fwd", q_bf16, nq, k_bf16, nk, d);
  if (rc) return rc;
  if (nq == 0) return MMCL_OK;
  REQUIRE(nk > 0 && band_hi && stats, "mmcl_sim_bce_fwd: null band_hi / stats or no keys");
  REQUIRE(row_base >= 0 && row_base + nq <= nk, "mmcl_sim_bce_fwd: query rows must be key rows");
  if (workspace_bytes < sim_bce_workspace_bytes(h->num_sms, nq, nk, d) || !workspace)
    return fail(MMCL_ERR_WORKSPACE, "mmcl_sim_bce_fwd: workspace too small");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const char* err = nullptr;
  cudaError_t e = launch_sim_bce_fwd(h->num_sms, q_bf16, nq, k_bf16, nk, d, scale, qinv, kinv,
                                     row_base, band_hi, stats, workspace, workspace_bytes,
                                     static_cast<cudaStream_t>(stream), &err);
  if (e != cudaSuccess) return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_sim_bce_fwd");
  return MMCL_OK;
}

int mmcl_sim_bce_bwd(mmcl_handle* h, const void* x_bf16, int64_t nx, const void* y_bf16, int64_t ny,
                     int64_t d, const float* scale, const float* xinv, const float* yinv,
                     int64_t row_base, const int32_t* band_hi, float* dx, mmcl_stream stream) {
  REQUIRE(h && scale, "mmcl_sim_bce_bwd: null handle/scale");
  int rc = check_sim_args("mmcl_sim_bce_bwd", x_bf16, nx, y_bf16, ny, d);
  if (rc) return rc;
  if (nx == 0 || ny == 0) return MMCL_OK;
  REQUIRE(band_hi && dx && (reinterpret_cast<uintptr_t>(dx) & 15) == 0,
          "mmcl_sim_bce_bwd: null band_hi / dx (16-byte aligned)");
  REQUIRE(row_base >= 0 && row_base + nx <= ny, "mmcl_sim_bce_bwd: x rows must be y rows");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const char* err = nullptr;
  cudaError_t e = launch_sim_bce_bwd_pair(h->num_sms, x_bf16, nx, y_bf16, ny, d, scale, xinv, yinv,
                                          row_base, band_hi, dx, static_cast<cudaStream_t>(stream),
                                          &err);
  if (e == cudaErrorNotSupported)
    return fail(MMCL_ERR_INVALID, "mmcl_sim_bce_bwd: d must be a multiple of 128 (256 above 512)");
  if (e != cudaSuccess) return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_sim_bce_bwd");
  return MMCL_OK;
}

int mmcl_pair_loss(mmcl_handle* h, const float* lse_a, const float* lse_b, const float* esum_a,
                   const float* esum_b, const float* pos, int64_t n, double weight, double n_norm,
                   const float* scale, float* out2, mmcl_stream stream) {
  REQUIRE(h && scale && out2, "mmcl_pair_loss: null handle/scale/out");
  REQUIRE(n >= 0 && n_norm > 0, "mmcl_pair_loss: bad n / n_norm");
  REQUIRE(n == 0 || (lse_a && lse_b && esum_a && esum_b && pos), "mmcl_pair_loss: null vector");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_pair_loss(lse_a, lse_b, esum_a, esum_b, pos, n, weight, n_norm, scale,
                                   out2, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_pair_loss");
}

int mmcl_sim_grad_bwd(mmcl_handle* h, const void* x_bf16, int64_t nx, const void* y_bf16,
                      int64_t ny, int64_t d, const float* scale, const float* xinv,
                      const float* yinv, const float* off_row, const float* off_col, float* dx,
                      mmcl_stream stream) {
  REQUIRE(h && scale, "mmcl_sim_grad_bwd: null handle/scale");
  int rc = check_sim_args("mmcl_sim_grad_bwd", x_bf16, nx, y_bf16, ny, d);
  if (rc) return rc;
  if (nx == 0 || ny == 0) return MMCL_OK;
  REQUIRE(dx && (reinterpret_cast<uintptr_t>(dx) & 15) == 0,
          "mmcl_sim_grad_bwd: dx must be non-null and 16-byte aligned");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const char* err = nullptr;
  cudaError_t e = launch_sim_grad_bwd(h->num_sms, x_bf16, nx, y_bf16, ny, d, scale, xinv, yinv,
                                      off_row, off_col, 1.f, 0.f, 0.f, dx,
                                      static_cast<cudaStream_t>(stream), &err);
  if (e != cudaSuccess) return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_sim_grad_bwd");
  return MMCL_OK;
}

int mmcl_sim_grad_bwd_lse(mmcl_handle* h, const void* x_bf16, int64_t nx, const void* y_bf16,
                          int64_t ny, int64_t d, const float* scale, const float* xinv,
                          const float* yinv, const float* lse_row, double row_den,
                          const float* lse_col, double col_den, float* dx, int zero_dx,
                          mmcl_stream stream) {
  REQUIRE(h && scale, "mmcl_sim_grad_bwd_lse: null handle/scale");
  int rc = check_sim_args("mmcl_sim_grad_bwd_lse", x_bf16, nx, y_bf16, ny, d);
  if (rc) return rc;
  if (nx == 0) return MMCL_OK;
  REQUIRE(dx && (reinterpret_cast<uintptr_t>(dx) & 15) == 0,
          "mmcl_sim_grad_bwd_lse: dx must be non-null and 16-byte aligned");
  REQUIRE((!lse_row || row_den > 0) && (!lse_col || col_den > 0),
          "mmcl_sim_grad_bwd_lse: denominators must be positive");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  auto st = static_cast<cudaStream_t>(stream);
  if (zero_dx) {
    cudaError_t e = cudaMemsetAsync(dx, 0, size_t(nx) * size_t(d) * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "mmcl_sim_grad_bwd_lse (memset)");
  }
  if (ny == 0) return MMCL_OK;
  const char* err = nullptr;
  const float log2e = 1.4426950408889634f;
  cudaError_t e = launch_sim_grad_bwd(
      h->num_sms, x_bf16, nx, y_bf16, ny, d, scale, xinv, yinv, lse_row, lse_col, log2e,
      lse_row ? float(log2(row_den)) : 0.f, lse_col ? float(log2(col_den)) : 0.f, dx, st, &err);
  if (e != cudaSuccess)
    return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_sim_grad_bwd_lse");
  return MMCL_OK;
}

int mmcl_make_offsets(mmcl_handle* h, const float* lse, int64_t n, const float* denom_vec,
                      double denom_scalar, float* off, mmcl_stream stream) {
  REQUIRE(h, "mmcl_make_offsets: null handle");
  REQUIRE(n >= 0 && (n == 0 || (lse && off)), "mmcl_make_offsets: bad arguments");
  REQUIRE(denom_vec || denom_scalar > 0, "mmcl_make_offsets: denominator must be positive");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_make_offsets(lse, n, denom_vec, denom_scalar, off,
                                      static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_make_offsets");
}

int mmcl_unpack_grad(mmcl_handle* h, const float* dacc, const void* partner_bf16,
                     const float* partner_inv, int64_t n,
                     int64_t d, double weight, double pos_coef, const float* gout,
                     const float* scale, const void* src, int src_dtype, const float* inv_norm,
                     int normalize, const int64_t* idx, void* grad, int grad_dtype, int direct,
                     mmcl_stream stream) {
  REQUIRE(h && scale, "mmcl_unpack_grad: null handle/scale");
  REQUIRE(n >= 0 && d > 0 && d % 8 == 0 && d <= 2048, "mmcl_unpack_grad: bad n/d");
  REQUIRE(src_dtype >= 0 && src_dtype <= 2, "mmcl_unpack_grad: dtype must be f32/bf16/f16");
  REQUIRE(grad_dtype >= 0 && grad_dtype <= 2, "mmcl_unpack_grad: grad dtype must be f32/bf16/f16");
  REQUIRE(direct || grad_dtype == MMCL_F32, "mmcl_unpack_grad: accumulation needs an fp32 buffer");
  REQUIRE(!direct || !idx, "mmcl_unpack_grad: direct stores need the identity row map");
  REQUIRE(n == 0 || (dacc && partner_bf16 && grad), "mmcl_unpack_grad: null buffer");
  REQUIRE(!normalize || (src && inv_norm), "mmcl_unpack_grad: normalize needs src and inv_norm");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaError_t e = launch_unpack_grad(dacc, partner_bf16, partner_inv, n, d, weight, pos_coef, gout,
                                     scale, src, src_dtype, inv_norm, normalize, idx, grad,
                                     grad_dtype, direct, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? MMCL_OK : cuda_fail(e, "mmcl_unpack_grad");
}

}  // extern "C"
