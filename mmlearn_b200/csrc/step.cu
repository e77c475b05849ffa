// Steady-state step executor (include/mmcl.h, mmcl_step_*): the whole launch sequence of one
// loss evaluation issued from native code.  It contains no arithmetic of its own -- every kernel
// is one of the launchers in kernels.h, called in the order mmlearn_b200/loss.py calls them on its
// general path; what it removes is the interpreter time between the launches.
//
// Buffers.  `workspace` (caller-owned, one per forward) holds everything local:
//   [info: P x 4 int64][kg: M x 4 int64][out2: P x 2 f32][vec: P x 5 x rows f32]
//   [identity-check table][forward split-K partials][rows + inverse norms (world == 1)]
//   [dacc: one rows x d f32 block per (pair, side) that needs a gradient]
// and, for world > 1, buffer `parity` of the symmetric arena holds what crosses NVLink:
//   [table: W x P x 4 int64][ids of every key modality: N x 2 int64][rows: N x d bf16 and
//   inverse norms: N f32 per modality][lse_a, lse_b: N f32 per pair]
// with N = world * rows.  Offsets are fixed at mmcl_step_create and identical on all ranks.
//
// Overlap (world > 1, opt-in with MMCL_OVERLAP=1).  The first kernel of the step (direction a of pair 0) needs the local rows
// of its query modality and ALL rows of its key modality.  Only the key modality is therefore
// pushed to the peers before the first fence; every other modality is packed into the LOCAL slot
// only, and a side stream pushes it to the peers (and runs its own fence, on a separate flag
// channel) while the main stream already computes direction a.  The main stream joins the side
// stream before the first kernel that reads remote rows of those modalities.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "handle.h"

using namespace mmcl;

namespace {

constexpr size_t kAlign = 1024;
inline size_t align_up(size_t x) { return (x + kAlign - 1) & ~(kAlign - 1); }

struct PairOff {
  size_t vec;         // workspace: 5 x rows floats
  size_t dacc[2];     // workspace: rows x d floats (or SIZE_MAX)
  size_t lse_all[2];  // arena: N floats each (world > 1, need_grad, has_col)
};

}  // namespace

struct mmcl_step {
  mmcl_handle* h = nullptr;
  int M = 0, P = 0;
  bool normalize = false, all_reduce_loss = false, need_grad = false;
  long long rows = 0, d = 0, N = 0;
  std::vector<int32_t> dtype, needs_grad;
  std::vector<mmcl_step_pair> pairs;
  std::vector<char> is_key;  // modality is the second entry of some pair (its ids are checked)
  // workspace layout
  size_t off_info = 0, off_kg = 0, off_out2 = 0, off_idws = 0, idws_bytes = 0, off_fws = 0,
         fws_bytes = 0, ws_bytes = 0;
  std::vector<size_t> ws_rows, ws_inv;  // world == 1: packed rows / inverse norms per modality
  std::vector<PairOff> po;
  // arena layout (offsets inside one parity buffer)
  size_t ar_table = 0, ar_bytes = 0;
  std::vector<size_t> ar_ids, ar_rows, ar_inv;
  // verification
  int64_t* host = nullptr;  // pinned: [W x P x 4 table][M kg counts]
  size_t host_words = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
  bool last_local_ok = false;
  // overlap of the second-wave pushes with the first similarity kernel
  bool overlap = false;
  int first_key = -1;  // modality whose rows every rank needs before anything can start
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

struct Views {  // device pointers of one step
  char* ws;
  char* ar;  // arena buffer (world > 1) or nullptr
};

inline const void* rows_ptr(const mmcl_step* s, const Views& v, int m) {
  return s->h->world > 1 ? static_cast<const void*>(v.ar + s->ar_rows[m])
                         : static_cast<const void*>(v.ws + s->ws_rows[m]);
}
inline const float* inv_ptr(const mmcl_step* s, const Views& v, int m) {
  if (!s->normalize) return nullptr;
  return reinterpret_cast<const float*>(s->h->world > 1 ? v.ar + s->ar_inv[m]
                                                        : v.ws + s->ws_inv[m]);
}

int enqueue_grad_kernels(mmcl_step* s, const float* scales, const Views& v, cudaStream_t st) {
  mmcl_handle* h = s->h;
  const long long b = s->rows, N = s->N, d = s->d;
  const long long off = (long long)h->rank * b;
  const float log2e = 1.4426950408889634f;
  for (int p = 0; p < s->P; ++p) {
    const mmcl_step_pair& pr = s->pairs[p];
    float* vec = reinterpret_cast<float*>(v.ws + s->po[p].vec);
    float* lse[2] = {vec, vec + b};
    for (int side = 0; side < 2; ++side) {
      const int mx = side == 0 ? pr.mod_a : pr.mod_b;
      const int my = side == 0 ? pr.mod_b : pr.mod_a;
      if (!s->needs_grad[mx]) continue;
      float* dacc = reinterpret_cast<float*>(v.ws + s->po[p].dacc[side]);
      cudaError_t e = cudaMemsetAsync(dacc, 0, size_t(b) * size_t(d) * sizeof(float), st);
      if (e != cudaSuccess) return cuda_fail(e, "mmcl_step (memset dacc)");
      const char* x = static_cast<const char*>(rows_ptr(s, v, mx)) + size_t(off) * d * 2;
      const float* xinv = inv_ptr(s, v, mx);
      if (xinv) xinv += off;
      // column term: log-sum-exp of the OTHER direction over all N rows of S
      const float* lse_col = nullptr;
      if (pr.has_col)
        lse_col = h->world > 1 ? reinterpret_cast<const float*>(v.ar + s->po[p].lse_all[1 - side])
                               : lse[1 - side];
      const char* err = nullptr;
      e = launch_sim_grad_bwd(h->num_sms, x, b, rows_ptr(s, v, my), N, d,
                              scales + pr.scale_index, xinv, inv_ptr(s, v, my), lse[side], lse_col,
                              log2e, float(log2(pr.row_den)),
                              lse_col ? float(log2(pr.col_den)) : 0.f, dacc, st, &err);
      if (e != cudaSuccess) return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_step (bwd)");
    }
  }
  return MMCL_OK;
}

}  // namespace

extern "C" {

int mmcl_step_create(mmcl_handle* h, const mmcl_step_desc* desc, mmcl_step** out) {
  REQUIRE(h && desc && out, "mmcl_step_create: null argument");
  REQUIRE(desc->n_modalities >= 1 && desc->n_modalities <= 64 && desc->n_pairs >= 1 &&
              desc->n_pairs <= 4096,
          "mmcl_step_create: bad modality / pair count");
  REQUIRE(desc->rows > 0 && desc->rows < (1ll << 31) / (h->world > 0 ? h->world : 1),
          "mmcl_step_create: bad row count");
  REQUIRE(desc->d > 0 && desc->d % 8 == 0 && desc->d <= 2048,
          "mmcl_step_create: d must be a multiple of 8 in [8, 2048]");
  REQUIRE(desc->mod_dtype && desc->mod_needs_grad && desc->pairs,
          "mmcl_step_create: null array");
  REQUIRE(h->world <= 32, "mmcl_step_create: at most 32 ranks");
  auto* s = new mmcl_step();
  s->h = h;
  s->M = desc->n_modalities;
  s->P = desc->n_pairs;
  s->normalize = desc->normalize != 0;
  s->all_reduce_loss = desc->all_reduce_loss != 0 && h->world > 1;
  s->need_grad = desc->need_grad != 0;
  s->rows = desc->rows;
  s->d = desc->d;
  s->N = desc->rows * h->world;
  s->dtype.assign(desc->mod_dtype, desc->mod_dtype + s->M);
  s->needs_grad.assign(desc->mod_needs_grad, desc->mod_needs_grad + s->M);
  s->pairs.assign(desc->pairs, desc->pairs + s->P);
  s->is_key.assign(s->M, 0);
  for (int m = 0; m < s->M; ++m)
    if (s->dtype[m] < 0 || s->dtype[m] > 2) {
      delete s;
      return fail(MMCL_ERR_INVALID, "mmcl_step_create: dtype must be f32/bf16/f16");
    }
  for (const auto& pr : s->pairs) {
    if (pr.mod_a < 0 || pr.mod_a >= s->M || pr.mod_b < 0 || pr.mod_b >= s->M ||
        pr.scale_index < 0 || !(pr.row_den > 0) || (pr.has_col && !(pr.col_den > 0)) ||
        !(pr.loss_rows > 0)) {
      delete s;
      return fail(MMCL_ERR_INVALID, "mmcl_step_create: bad pair descriptor");
    }
    s->is_key[pr.mod_b] = 1;
  }
  const size_t b = size_t(s->rows), N = size_t(s->N), d = size_t(s->d);
  const int W = h->world;
  // ---- workspace ----
  size_t o = 0;
  s->off_info = o;
  o = align_up(o + size_t(s->P) * 32);
  s->off_kg = o;
  o = align_up(o + size_t(s->M) * 32);
  s->off_out2 = o;
  o = align_up(o + size_t(s->P) * 8);
  s->po.resize(s->P);
  const size_t off_vec0 = o;
  for (int p = 0; p < s->P; ++p) {
    s->po[p].vec = o;
    o += 5 * b * sizeof(float);  // contiguous [P][5][rows]
  }
  (void)off_vec0;
  o = align_up(o);
  s->off_idws = o;
  s->idws_bytes = identity_check_workspace_bytes((long long)(W > 1 ? N : b));
  o = align_up(o + s->idws_bytes);
  s->off_fws = o;
  s->fws_bytes = sim_lse_workspace_bytes(h->num_sms, (long long)b, (long long)N, (long long)d);
  o = align_up(o + s->fws_bytes);
  s->ws_rows.assign(s->M, 0);
  s->ws_inv.assign(s->M, 0);
  if (W == 1) {
    for (int m = 0; m < s->M; ++m) {
      s->ws_rows[m] = o;
      o = align_up(o + b * d * 2);
      s->ws_inv[m] = o;
      if (s->normalize) o = align_up(o + b * 4);
    }
  }
  for (int p = 0; p < s->P; ++p) {
    const int mods[2] = {s->pairs[p].mod_a, s->pairs[p].mod_b};
    for (int side = 0; side < 2; ++side) {
      s->po[p].dacc[side] = SIZE_MAX;
      if (s->need_grad && s->needs_grad[mods[side]]) {
        s->po[p].dacc[side] = o;
        o = align_up(o + b * d * sizeof(float));
      }
    }
  }
  s->ws_bytes = o;
  // ---- arena ----
  s->ar_ids.assign(s->M, 0);
  s->ar_rows.assign(s->M, 0);
  s->ar_inv.assign(s->M, 0);
  if (W > 1) {
    size_t a = 0;
    s->ar_table = a;
    a = align_up(a + size_t(W) * s->P * 32);
    for (int m = 0; m < s->M; ++m) {
      if (s->is_key[m]) {
        s->ar_ids[m] = a;
        a = align_up(a + N * 16);
      }
      s->ar_rows[m] = a;
      a = align_up(a + N * d * 2);
      s->ar_inv[m] = a;
      if (s->normalize) a = align_up(a + N * 4);
    }
    for (int p = 0; p < s->P; ++p)
      for (int side = 0; side < 2; ++side) {
        s->po[p].lse_all[side] = a;
        if (s->need_grad && s->pairs[p].has_col) a = align_up(a + N * 4);
      }
    s->ar_bytes = a;
  }
  // ---- verification state ----
  s->host_words = size_t(W) * s->P * 4 + size_t(s->M);
  DeviceGuard g(h->device);
  cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&s->host), s->host_words * 8);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming);
  s->first_key = s->pairs[0].mod_b;
  // Opt-in (MMCL_OVERLAP=1): measured at 8 x B200 (b = 4096) it is neutral -- 0.988 ms/step with,
  // 0.980 ms without -- the pushes of both modalities share the same NVLink ports, so moving one
  // of them behind the first kernel does not shorten the first fence.
  const char* ov = getenv("MMCL_OVERLAP");
  s->overlap = W > 1 && s->M > 1 && ov && ov[0] == '1';
  if (e == cudaSuccess && s->overlap) {
    e = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    if (s->host) cudaFreeHost(s->host);
    if (s->ev) cudaEventDestroy(s->ev);
    if (s->side) cudaStreamDestroy(s->side);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    delete s;
    return cuda_fail(e, "mmcl_step_create");
  }
  *out = s;
  return MMCL_OK;
}

int mmcl_step_destroy(mmcl_step* s) {
  if (!s) return MMCL_OK;
  DeviceGuard g(s->h->device);
  if (s->ev) cudaEventDestroy(s->ev);
  if (s->side) {
    cudaStreamSynchronize(s->side);
    cudaStreamDestroy(s->side);
  }
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  if (s->host) cudaFreeHost(s->host);
  delete s;
  return MMCL_OK;
}

size_t mmcl_step_workspace_bytes(const mmcl_step* s) { return s ? s->ws_bytes : 0; }
size_t mmcl_step_arena_bytes(const mmcl_step* s) { return s ? s->ar_bytes : 0; }

int mmcl_step_layout(const mmcl_step* s, mmcl_step_layout_t* out) {
  REQUIRE(s && out, "mmcl_step_layout: null argument");
  out->off_out2 = s->off_out2;
  out->off_vec = s->po[0].vec;
  out->off_info = s->off_info;
  return MMCL_OK;
}

int mmcl_step_forward(mmcl_step* s, int parity, int local_ok, const void* const* emb,
                      const int64_t* const* ids, const float* scales, void* workspace,
                      int launch_grad, mmcl_stream stream) {
  REQUIRE(s && workspace && scales, "mmcl_step_forward: null argument");
  REQUIRE(!local_ok || (emb && ids), "mmcl_step_forward: null embedding / id arrays");
  mmcl_handle* h = s->h;
  const int W = h->world;
  REQUIRE(W == 1 || (h->arena.local && h->arena.bytes >= s->ar_bytes),
          "mmcl_step_forward: symmetric arena not reserved (mmcl_symm_reserve)");
  if (local_ok)
    for (int m = 0; m < s->M; ++m) {
      REQUIRE(emb[m] && ids[m], "mmcl_step_forward: null embedding / ids");
      REQUIRE((reinterpret_cast<uintptr_t>(emb[m]) & 15) == 0,
              "mmcl_step_forward: embeddings must be 16-byte aligned");
    }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  auto st = static_cast<cudaStream_t>(stream);
  Views v{static_cast<char*>(workspace), W > 1 ? symm_buffer(&h->arena, parity) : nullptr};
  const long long b = s->rows, N = s->N, d = s->d;
  const long long off = (long long)h->rank * b;
  cudaError_t e;
  s->last_local_ok = local_ok != 0;
#define CU(call, what)                                   \
  do {                                                   \
    e = (call);                                          \
    if (e != cudaSuccess) return cuda_fail(e, what);     \
  } while (0)

  auto* info = reinterpret_cast<int64_t*>(v.ws + s->off_info);
  auto* kg = reinterpret_cast<int64_t*>(v.ws + s->off_kg);
  float* out2 = reinterpret_cast<float*>(v.ws + s->off_out2);
  void* idws = v.ws + s->off_idws;
  // info / kg / out2 are adjacent 1 KB-aligned blocks: one memset clears all three
  CU(cudaMemsetAsync(v.ws, 0, s->po[0].vec, st), "mmcl_step_forward (memset)");

  // ---- 1. local id matching: is every pair still the identity? ----
  if (local_ok)
    for (int p = 0; p < s->P; ++p)
      CU(launch_identity_check(ids[s->pairs[p].mod_a], b, ids[s->pairs[p].mod_b], b, info + 4 * p,
                               idws, s->idws_bytes, st),
         "mmcl_step_forward (identity check)");
  // (local_ok == 0: the zeroed rows never equal the assumed {rows, 1, rows, rows})

  // ---- 2. pack rows (and publish table, ids, rows, norms to every rank) ----
  bool joined = true;  // false while the side stream still pushes second-wave modalities
  if (W > 1) {
    CU(launch_symm_put(&h->arena, W, parity, info, size_t(s->P) * 32,
                       s->ar_table + size_t(h->rank) * s->P * 32, st),
       "mmcl_step_forward (table)");
    if (local_ok)
      for (int m = 0; m < s->M; ++m) {
        if (s->is_key[m])
          CU(launch_symm_put(&h->arena, W, parity, ids[m], size_t(b) * 16,
                             s->ar_ids[m] + size_t(off) * 16, st),
             "mmcl_step_forward (ids)");
        const size_t o_rows = s->ar_rows[m] + size_t(off) * d * 2;
        const size_t o_inv = s->ar_inv[m] + size_t(off) * 4;
        if (!s->overlap || m == s->first_key) {
          CU(launch_pack_rows_symm(&h->arena, W, parity, emb[m], s->dtype[m], b, d, nullptr, b,
                                   s->normalize, o_rows, s->normalize ? (long long)o_inv : -1, st),
             "mmcl_step_forward (pack)");
        } else {  // second wave: local slot now, peers from the side stream
          CU(launch_pack_rows(emb[m], s->dtype[m], b, d, nullptr, b, s->normalize, v.ar + o_rows,
                              s->normalize ? reinterpret_cast<float*>(v.ar + o_inv) : nullptr, st),
             "mmcl_step_forward (pack local)");
        }
      }
    if (s->overlap) {
      // every rank walks through the side-stream fence, whatever its local state
      CU(cudaEventRecord(s->ev_fork, st), "mmcl_step_forward (fork)");
      CU(cudaStreamWaitEvent(s->side, s->ev_fork, 0), "mmcl_step_forward (fork wait)");
      if (local_ok)
        for (int m = 0; m < s->M; ++m) {
          if (m == s->first_key) continue;
          const size_t o_rows = s->ar_rows[m] + size_t(off) * d * 2;
          const size_t o_inv = s->ar_inv[m] + size_t(off) * 4;
          CU(launch_symm_put(&h->arena, W, parity, v.ar + o_rows, size_t(b) * d * 2, o_rows,
                             s->side, h->rank),
             "mmcl_step_forward (push rows)");
          if (s->normalize)
            CU(launch_symm_put(&h->arena, W, parity, v.ar + o_inv, size_t(b) * 4, o_inv, s->side,
                               h->rank),
               "mmcl_step_forward (push norms)");
        }
      CU(launch_symm_fence(&h->arena, W, h->rank, s->side, 1), "mmcl_step_forward (side fence)");
      CU(cudaEventRecord(s->ev_join, s->side), "mmcl_step_forward (join record)");
      joined = false;
    }
    CU(launch_symm_fence(&h->arena, W, h->rank, st), "mmcl_step_forward (fence)");
  } else if (local_ok) {
    for (int m = 0; m < s->M; ++m)
      CU(launch_pack_rows(emb[m], s->dtype[m], b, d, nullptr, b, s->normalize,
                          v.ws + s->ws_rows[m],
                          s->normalize ? reinterpret_cast<float*>(v.ws + s->ws_inv[m]) : nullptr,
                          st),
         "mmcl_step_forward (pack)");
  }
  auto join = [&]() -> cudaError_t {
    if (joined) return cudaSuccess;
    joined = true;
    return cudaStreamWaitEvent(st, s->ev_join, 0);
  };

  if (local_ok) {
    // ---- 3. cross-rank duplicate check on the gathered ids of every key modality ----
    if (W > 1)
      for (int m = 0; m < s->M; ++m)
        if (s->is_key[m]) {
          const auto* g_ids = reinterpret_cast<const int64_t*>(v.ar + s->ar_ids[m]);
          CU(launch_identity_check(g_ids, N, g_ids, N, kg + 4 * m, idws, s->idws_bytes, st),
             "mmcl_step_forward (global id check)");
        }

    // ---- 4. fused similarity + log-sum-exp, both directions; positives; pair loss ----
    for (int p = 0; p < s->P; ++p) {
      const mmcl_step_pair& pr = s->pairs[p];
      const float* sp = scales + pr.scale_index;
      float* vec = reinterpret_cast<float*>(v.ws + s->po[p].vec);
      float *lse_a = vec, *lse_b = vec + b, *es_a = vec + 2 * b, *es_b = vec + 3 * b,
            *pos = vec + 4 * b;
      const char* fa_all = static_cast<const char*>(rows_ptr(s, v, pr.mod_a));
      const char* fb_all = static_cast<const char*>(rows_ptr(s, v, pr.mod_b));
      const float* ia_all = inv_ptr(s, v, pr.mod_a);
      const float* ib_all = inv_ptr(s, v, pr.mod_b);
      const char* fa = fa_all + size_t(off) * d * 2;
      const char* fb = fb_all + size_t(off) * d * 2;
      const float* ia = ia_all ? ia_all + off : nullptr;
      const float* ib = ib_all ? ib_all + off : nullptr;
      const char* err = nullptr;
      // direction a of pair 0 reads remote rows of the first-wave modality only
      if (!(p == 0 && pr.mod_b == s->first_key)) CU(join(), "mmcl_step_forward (join)");
      e = launch_sim_lse_fwd(h->num_sms, fa, b, fb_all, N, d, sp, ia, ib_all, lse_a, es_a,
                             v.ws + s->off_fws, s->fws_bytes, st, &err);
      if (e == cudaSuccess) e = join();
      if (e == cudaSuccess)
        e = launch_sim_lse_fwd(h->num_sms, fb, b, fa_all, N, d, sp, ib, ia_all, lse_b, es_b,
                               v.ws + s->off_fws, s->fws_bytes, st, &err);
      if (e != cudaSuccess)
        return err ? fail(MMCL_ERR_CUDA, err) : cuda_fail(e, "mmcl_step_forward (fwd)");
      CU(launch_positive_logits(fa, b, fb_all, N, d, off, sp, ia, ib_all, pos, st),
         "mmcl_step_forward (positives)");
      CU(launch_pair_loss(lse_a, lse_b, es_a, es_b, pos, b, pr.weight, pr.loss_rows, sp,
                          out2 + 2 * p, st),
         "mmcl_step_forward (pair loss)");
    }
  }
  CU(join(), "mmcl_step_forward (join)");  // (also on ranks that only go through the motions)
  if (s->all_reduce_loss) {
    ncclResult_t r = ncclAllReduce(out2, out2, size_t(s->P) * 2, ncclFloat, ncclSum, h->comm, st);
    if (r != ncclSuccess) return nccl_fail(r, "mmcl_step_forward (all-reduce)");
  }

  // ---- 5. log-sum-exp exchange for the column terms of backward ----
  if (W > 1 && s->need_grad) {
    bool any = false;
    for (int p = 0; p < s->P; ++p) {
      if (!s->pairs[p].has_col) continue;
      any = true;
      if (!local_ok) continue;
      const float* vec = reinterpret_cast<const float*>(v.ws + s->po[p].vec);
      for (int side = 0; side < 2; ++side)
        CU(launch_symm_put(&h->arena, W, parity, vec + side * b, size_t(b) * 4,
                           s->po[p].lse_all[side] + size_t(off) * 4, st),
           "mmcl_step_forward (lse)");
    }
    if (any) CU(launch_symm_fence(&h->arena, W, h->rank, st), "mmcl_step_forward (fence)");
  }

  // ---- 6. table read-back for verification: the step's only host wait ----
  const size_t tab_words = size_t(W) * s->P * 4;
  CU(cudaMemcpyAsync(s->host, W > 1 ? static_cast<const void*>(v.ar + s->ar_table)
                                    : static_cast<const void*>(info),
                     tab_words * 8, cudaMemcpyDeviceToHost, st),
     "mmcl_step_forward (table copy)");
  if (W > 1)
    CU(cudaMemcpy2DAsync(s->host + tab_words, 8, kg, 32, 8, size_t(s->M), cudaMemcpyDeviceToHost,
                         st),
       "mmcl_step_forward (kg copy)");
  CU(cudaEventRecord(s->ev, st), "mmcl_step_forward (event)");
  s->pending = true;

  // ---- 7. the heavy half of backward does not depend on anything backward() receives ----
  if (launch_grad && s->need_grad && local_ok) return enqueue_grad_kernels(s, scales, v, st);
  return MMCL_OK;
#undef CU
}

int mmcl_step_verify(mmcl_step* s, int* ok) {
  REQUIRE(s && ok, "mmcl_step_verify: null argument");
  REQUIRE(s->pending, "mmcl_step_verify: no forward in flight");
  cudaError_t e = cudaEventSynchronize(s->ev);
  s->pending = false;
  if (e != cudaSuccess) return cuda_fail(e, "mmcl_step_verify");
  const int W = s->h->world;
  bool good = s->last_local_ok;
  const int64_t b = s->rows;
  for (size_t i = 0; good && i < size_t(W) * s->P; ++i) {
    const int64_t* r = s->host + 4 * i;
    good = r[0] == b && r[1] == 1 && r[2] == b && r[3] == b;
  }
  if (W > 1)
    for (int m = 0; good && m < s->M; ++m)
      if (s->is_key[m]) good = s->host[size_t(W) * s->P * 4 + m] == s->N;
  *ok = good ? 1 : 0;
  return MMCL_OK;
}

int mmcl_step_grad_kernels(mmcl_step* s, int parity, const float* scales, void* workspace,
                           mmcl_stream stream) {
  REQUIRE(s && scales && workspace, "mmcl_step_grad_kernels: null argument");
  REQUIRE(s->need_grad, "mmcl_step_grad_kernels: the step was created without need_grad");
  mmcl_handle* h = s->h;
  REQUIRE(h->world == 1 || h->arena.local, "mmcl_step_grad_kernels: no arena reserved");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  Views v{static_cast<char*>(workspace),
          h->world > 1 ? symm_buffer(&h->arena, parity) : nullptr};
  return enqueue_grad_kernels(s, scales, v, static_cast<cudaStream_t>(stream));
}

int mmcl_step_backward(mmcl_step* s, int parity, const void* const* emb, const float* scales,
                       void* workspace, const float* gout, void* const* grads,
                       const int32_t* grad_direct, mmcl_stream stream) {
  REQUIRE(s && emb && scales && workspace && gout && grads && grad_direct,
          "mmcl_step_backward: null argument");
  REQUIRE(s->need_grad, "mmcl_step_backward: the step was created without need_grad");
  mmcl_handle* h = s->h;
  REQUIRE(h->world == 1 || h->arena.local, "mmcl_step_backward: no arena reserved");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  auto st = static_cast<cudaStream_t>(stream);
  Views v{static_cast<char*>(workspace),
          h->world > 1 ? symm_buffer(&h->arena, parity) : nullptr};
  const long long b = s->rows, d = s->d;
  const long long off = (long long)h->rank * b;
  for (int p = 0; p < s->P; ++p) {
    const mmcl_step_pair& pr = s->pairs[p];
    for (int side = 0; side < 2; ++side) {
      const int mx = side == 0 ? pr.mod_a : pr.mod_b;
      const int my = side == 0 ? pr.mod_b : pr.mod_a;
      if (!s->needs_grad[mx] || !grads[mx]) continue;
      const float* dacc = reinterpret_cast<const float*>(v.ws + s->po[p].dacc[side]);
      const char* partner = static_cast<const char*>(rows_ptr(s, v, my)) + size_t(off) * d * 2;
      const float* pinv = inv_ptr(s, v, my);
      const float* inv = inv_ptr(s, v, mx);
      cudaError_t e = launch_unpack_grad(
          dacc, partner, pinv ? pinv + off : nullptr, b, d, pr.weight, pr.pos_coef, gout,
          scales + pr.scale_index, emb[mx], s->dtype[mx], inv ? inv + off : nullptr, s->normalize,
          nullptr, grads[mx], grad_direct[mx] ? s->dtype[mx] : MMCL_F32, grad_direct[mx] ? 1 : 0,
          st);
      if (e != cudaSuccess) return cuda_fail(e, "mmcl_step_backward (unpack)");
    }
  }
  return MMCL_OK;
}

}  // extern "C"
