// Steady-state step executor (include/mmcl.h, mmcl_step_*): one loss evaluation as FIVE kernel
// launches, issued from native code.
//
//   forward   1. step_pack_kernel      all modalities: bf16 rows + fp32 inverse norms + ids into
//                                      this rank's slot of the gather buffers, elementwise id check
//             2. sim_lse_fwd_kernel    every (pair, direction) in one persistent launch; the
//                                      positive logit is picked out of the tile that holds it.
//                                      Several ranks: the launch also IS the all-gather -- its two
//                                      spare warps per CTA store this rank's slot into every peer's
//                                      arena (NVLink) while the tensor cores start on the local
//                                      keys, and the TMA producer acquires the per-source,
//                                      per-modality ready flag before its first load of a remote
//                                      key block (local keys first, then remote ones in ring order)
//             3. step_post_fwd_kernel  merge partial statistics, loss + d(loss)/d(scale), push
//                                      the log-sum-exps to every rank, distinctness of the gathered
//                                      key ids, verification word, lse ready flags
//             (8-byte copy of the verification word to pinned memory: the step's only host wait)
//   backward  4. sim_grad_bwd_pair_kernel  every (pair, side) in one launch; waits for the lse
//                                      flags; results stored once (no memset, no atomics)
//             5. step_unpack_kernel    positive term, normalise backward, output dtype
//
// Buffers.  `workspace` (caller-owned, one per forward) holds everything local:
//   [header words][out2: P x 2 f32][vec: P x 5 x rows f32][block partial sums][id hash tables]
//   [forward partial statistics][rows + inverse norms (world == 1)]
//   [dacc: one rows x d f32 block per (pair, side) that needs a gradient][backward scratch tiles]
// and, for world > 1, buffer `parity` of the symmetric arena holds what crosses NVLink:
//   [status: W words][ids of every key modality: N x 2 int64][rows: N x d bf16 and inverse norms:
//   N f32 per modality][lse_a, lse_b: N f32 per pair]
// with N = world * rows.  Offsets are fixed at mmcl_step_create and identical on all ranks.
//
// Buffer reuse across steps needs no fence either: a rank's forward launch of step i+1 ends only
// after it has seen every peer's step-(i+1) flags, which those peers release after (in stream
// order) all of their step-i work; so when the rank writes buffer (i+2) & 1 = i & 1 again, nobody
// reads it any more.
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "handle.h"
#include "step_kernels.h"

using namespace mmcl;

namespace {

constexpr size_t kAlign = 1024;
inline size_t align_up(size_t x) { return (x + kAlign - 1) & ~(kAlign - 1); }

struct PairOff {
  size_t vec;         // workspace: 5 x rows floats
  size_t dacc[2];     // workspace: rows x d floats (or SIZE_MAX)
  size_t lse_all[2];  // arena: N floats each (world > 1, need_grad, has_col)
  size_t range;       // workspace: (min, max) of lse_a, lse_b of this rank (4 floats)
  size_t range_all;   // arena: [world][4] floats
};

struct Range {
  explicit Range(const char* name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};

unsigned long long next_pow2(unsigned long long x) {
  unsigned long long p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace

struct mmcl_step {
  mmcl_handle* h = nullptr;
  int M = 0, P = 0;
  bool normalize = false, all_reduce_loss = false, need_grad = false;
  long long rows = 0, d = 0, N = 0;
  std::vector<int32_t> dtype, needs_grad;
  std::vector<mmcl_step_pair> pairs;
  std::vector<char> is_key;  // modality is the second entry of some pair (its ids are checked)
  bool pair_bwd = false;     // the CTA-pair backward kernel supports this d
  // workspace layout
  size_t off_out2 = 0, off_blk = 0, off_blkr = 0, off_fws = 0, fws_bytes = 0, off_scratch = 0, ws_bytes = 0;
  std::vector<size_t> ws_rows, ws_inv, ws_idtab;
  unsigned long long idtab_size = 0;
  std::vector<PairOff> po;
  // arena layout (offsets inside one parity buffer)
  size_t ar_status = 0, ar_bytes = 0;
  std::vector<size_t> ar_ids, ar_rows, ar_inv;
  // launch plans (shapes are fixed; pointers are filled in per call)
  FwdLaunch fwd;
  std::vector<size_t> fwd_part_off;
  std::vector<int> sub_idx;  // [(pair * 2 + direction) * 2 + part]: forward sub-problem indices
  int nsub_dir = 1;          // parts per (pair, direction): local keys [+ remote keys]
  BwdLaunch bwd;
  int post_blocks_per_pair = 0, post_main_blocks = 0, post_total_blocks = 0;
  std::vector<int> push_order;  // modalities in the order their rows are pushed to the peers
  // device state owned by the step: completion counters of the push warps / post-forward blocks
  unsigned long long* d_tickets = nullptr;  // [kFlagChannels] + [1]
  unsigned long long n_push_launches = 0, n_post_launches = 0, n_pack_launches = 0;
  int first_wave = -1;  // modality the pack kernel itself stores into every rank (world > 1)
  bool fanout_all = false;  // ... or every modality (MMCL_PACK_FANOUT_ALL=1): no pushes in the forward kernel
  unsigned gen = 0, salt = 0;
  unsigned long long seq = 0;  // flag sequence number of the last forward
  // verification
  unsigned long long* host = nullptr;  // pinned: [verify word][err word]
  cudaEvent_t ev = nullptr;
  bool pending = false;
  // tensor-map cache per parity (the arena and, usually, the workspace do not move)
  TensorMaps maps_f[2], maps_b[2];
  const void* map_key[2] = {nullptr, nullptr};
  int map_victim = 0;
};

namespace {

struct Views {  // device pointers of one step
  char* ws;
  char* ar;  // arena buffer (world > 1) or nullptr
};

inline char* rows_ptr(const mmcl_step* s, const Views& v, int m) {
  return s->h->world > 1 ? v.ar + s->ar_rows[m] : v.ws + s->ws_rows[m];
}
inline float* inv_ptr(const mmcl_step* s, const Views& v, int m) {
  if (!s->normalize) return nullptr;
  return reinterpret_cast<float*>(s->h->world > 1 ? v.ar + s->ar_inv[m] : v.ws + s->ws_inv[m]);
}
inline unsigned long long* hdr_ptr(const Views& v) {
  return reinterpret_cast<unsigned long long*>(v.ws);
}
inline unsigned long long buf_off(const mmcl_handle* h, int parity) {
  return kFlagBytes + size_t(parity & 1) * h->arena.bytes;
}
inline int row_chan(int m) { return kChanRows + (m < kMaxRowChannels ? m : kMaxRowChannels - 1); }

// tensor maps of the gathered rows of every modality, cached while the base pointers stay put
// (world > 1: one cache slot per arena parity; world == 1: the rows live in the workspace, of
// which a training loop typically alternates two -- a two-entry cache)
int ensure_maps(mmcl_step* s, const Views& v, int parity, int* slot_out) {
  const void* key = s->h->world > 1 ? static_cast<const void*>(v.ar) : static_cast<const void*>(v.ws);
  int slot = parity & 1;
  if (s->h->world == 1) {
    if (s->map_key[0] == key) slot = 0;
    else if (s->map_key[1] == key) slot = 1;
    else slot = s->map_victim ^= 1;
  }
  *slot_out = slot;
  if (s->map_key[slot] == key) return MMCL_OK;
  // per modality m: forward [2m] = query box over THIS RANK's rows, [2m+1] = key box over all N
  // rows; backward [3m] = x box over this rank's rows, [3m+1], [3m+2] = y boxes over all N rows
  const size_t loc = size_t(s->h->rank) * size_t(s->rows) * size_t(s->d) * 2;
  for (int m = 0; m < s->M; ++m) {
    TensorMaps tf, tb;
    if (fwd_make_maps(&s->maps_f[slot], m, rows_ptr(s, v, m), s->N, s->d) ||
        bwd_make_maps(&s->maps_b[slot], m, rows_ptr(s, v, m), s->N, s->d) ||
        fwd_make_maps(&tf, 0, rows_ptr(s, v, m) + loc, s->rows, s->d) ||
        bwd_make_maps(&tb, 0, rows_ptr(s, v, m) + loc, s->rows, s->d))
      return fail(MMCL_ERR_CUDA, "mmcl_step: cuTensorMapEncodeTiled failed");
    s->maps_f[slot].tm[2 * m] = tf.tm[0];
    s->maps_b[slot].tm[3 * m] = tb.tm[0];
  }
  s->map_key[slot] = key;
  return MMCL_OK;
}

int enqueue_grad_kernels(mmcl_step* s, const float* scales, const Views& v, int parity,
                         cudaStream_t st) {
  Range nv("mmcl.bwd");
  mmcl_handle* h = s->h;
  const long long b = s->rows, N = s->N, d = s->d;
  const long long off = (long long)h->rank * b;
  const float log2e = 1.4426950408889634f;
  int slot = 0;
  if (int rc = ensure_maps(s, v, parity, &slot)) return rc;
  unsigned long long* hdr = hdr_ptr(v);
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(h->arena.local);
  if (s->pair_bwd) {
    BwdLaunch L = s->bwd;
    int k = 0;
    for (int p = 0; p < s->P; ++p) {
      const mmcl_step_pair& pr = s->pairs[p];
      float* vec = reinterpret_cast<float*>(v.ws + s->po[p].vec);
      float* lse[2] = {vec, vec + b};
      for (int side = 0; side < 2; ++side) {
        const int mx = side == 0 ? pr.mod_a : pr.mod_b;
        const int my = side == 0 ? pr.mod_b : pr.mod_a;
        if (!s->needs_grad[mx]) continue;
        BwdProb& q = L.prob[k++];
        q.map_x = 3 * mx;
        q.map_y256 = 3 * my + 1;
        q.map_y64 = 3 * my + 2;
        q.scale = scales + pr.scale_index;
        q.xinv = inv_ptr(s, v, mx);
        if (q.xinv) q.xinv += off;
        q.yinv = inv_ptr(s, v, my);
        q.off_row = lse[side];
        q.off_col = nullptr;
        if (pr.has_col)
          q.off_col = h->world > 1
                          ? reinterpret_cast<const float*>(v.ar + s->po[p].lse_all[1 - side])
                          : lse[1 - side];
        q.off_mul = log2e;
        q.row_add = float(log2(pr.row_den));
        q.col_add = q.off_col ? float(log2(pr.col_den)) : 0.f;
        q.dx = reinterpret_cast<float*>(v.ws + s->po[p].dacc[side]);
        // offset ranges from the post-forward kernel (own rows: this rank's; columns: every rank's)
        const float* rng = reinterpret_cast<const float*>(v.ws + s->po[p].range);
        q.rng_row = rng + 2 * side;
        if (q.off_col && h->world > 1) {
          q.rng_col = reinterpret_cast<const float*>(v.ar + s->po[p].range_all) + 2 * (1 - side);
          q.rng_col_stride = 4;
          q.rng_col_count = h->world;
        } else {
          q.rng_col = rng + 2 * (1 - side);
          q.rng_col_stride = 4;
          q.rng_col_count = 1;
        }
      }
    }
    if (k == 0) return MMCL_OK;
    L.accumulate = 0;
    L.scratch = reinterpret_cast<float*>(v.ws + s->off_scratch);
    L.world = h->world;
    L.rank = h->rank;
    L.flags = flags;
    L.seq = s->seq;
    L.err = hdr + kHdrErr;
    cudaError_t e = launch_bwd_fused(s->maps_b[slot], L, st);
    if (e != cudaSuccess) return cuda_fail(e, "mmcl_step (bwd)");
    return MMCL_OK;
  }
  // d does not fit the CTA-pair kernel: one single-CTA launch per (pair, side), with a memset
  if (h->world > 1) {
    cudaError_t e = launch_step_wait_flags(flags, kChanLse, s->seq, h->world, h->rank,
                                           hdr + kHdrErr, st);
    if (e != cudaSuccess) return cuda_fail(e, "mmcl_step (wait lse)");
  }
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
      const char* x = rows_ptr(s, v, mx) + size_t(off) * d * 2;
      const float* xinv = inv_ptr(s, v, mx);
      if (xinv) xinv += off;
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
  REQUIRE(desc->n_modalities >= 1 && desc->n_modalities <= kStepMaxMods && desc->n_pairs >= 1 &&
              desc->n_pairs <= kStepMaxPairs,
          "mmcl_step_create: the fused step handles at most 6 modalities and 4 pairs");
  REQUIRE(desc->rows > 0 && desc->rows < (1ll << 31) / (h->world > 0 ? h->world : 1),
          "mmcl_step_create: bad row count");
  REQUIRE(desc->d > 0 && desc->d % 8 == 0 && desc->d <= 2048,
          "mmcl_step_create: d must be a multiple of 8 in [8, 2048]");
  REQUIRE(desc->mod_dtype && desc->mod_needs_grad && desc->pairs,
          "mmcl_step_create: null array");
  REQUIRE(h->world <= kMaxRanks, "mmcl_step_create: at most 32 ranks");
  REQUIRE(h->world == 1 || desc->rows % 4 == 0,
          "mmcl_step_create: with several ranks the fused step needs rows % 4 == 0 (16-byte peer "
          "stores of the inverse norms)");
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
  s->pair_bwd = bwd_pair_supported(h->num_sms, s->d);
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
  const long long off = (long long)h->rank * s->rows;

  // ---- forward plan: local key block of every (pair, direction) first, then the remote keys ----
  fwd_begin(&s->fwd, s->d);
  const int nkt_all = int((N + 255) / 256);
  int kt_lo = 0, kt_hi = nkt_all;
  if (W > 1) {
    kt_lo = int(off / 256);
    kt_hi = int((off + s->rows + 255) / 256);
  }
  const int n_local = kt_hi - kt_lo;
  const bool split = W > 1 && n_local < nkt_all;
  // phase 0: the local key block of every (pair, direction); phase 1: the remote keys of every
  // pair's direction a (needs the peers' rows of the pairs' second modalities); phase 2: the
  // remote keys of direction b.  world == 1: one phase with everything.
  s->sub_idx.assign(size_t(s->P) * 4, -1);
  s->nsub_dir = split ? 2 : 1;
  bool plan_ok = true;
  for (int p = 0; p < s->P && plan_ok; ++p)
    for (int dir = 0; dir < 2; ++dir) {
      // (one rank: every (pair, direction) is its own phase, so one key matrix at a time streams
      // through L2; several ranks: phase 0 holds the local key blocks of all of them)
      const int k = fwd_add_sub(&s->fwd, s->rows, s->N, split ? kt_lo : 0, split ? n_local : -1,
                                split ? 0 : p * 2 + dir);
      plan_ok = plan_ok && k >= 0;
      s->sub_idx[(p * 2 + dir) * 2] = k;
    }
  if (split)
    for (int dir = 0; dir < 2 && plan_ok; ++dir)
      for (int p = 0; p < s->P; ++p) {
        const int k = fwd_add_sub(&s->fwd, s->rows, s->N, kt_hi % nkt_all, nkt_all - n_local,
                                  1 + dir);
        plan_ok = plan_ok && k >= 0;
        s->sub_idx[(p * 2 + dir) * 2 + 1] = k;
      }
  if (!plan_ok) {
    delete s;
    return fail(MMCL_ERR_INVALID, "mmcl_step_create: too many forward sub-problems");
  }
  s->fwd_part_off.assign(s->fwd.nsub, 0);
  s->fws_bytes = fwd_finish(&s->fwd, h->num_sms, s->fwd_part_off.data());
  // modalities in the order the remote phases need them as keys: the second modality of every
  // pair (direction a), then the first ones (direction b), then whatever is left
  for (int which = 0; which < 2; ++which)
    for (int p = 0; p < s->P; ++p) {
      const int m = which == 0 ? s->pairs[p].mod_b : s->pairs[p].mod_a;
      bool seen = false;
      for (int x : s->push_order) seen = seen || x == m;
      if (!seen) s->push_order.push_back(m);
    }
  for (int m = 0; m < s->M; ++m) {
    bool seen = false;
    for (int x : s->push_order) seen = seen || x == m;
    if (!seen) s->push_order.push_back(m);
  }

  // The first modality of that order is needed before any remote tile can run: the pack kernel
  // stores it into every rank at the full-GPU store rate (~550 GB/s measured at 8 GPUs); the
  // others travel under the forward kernel's tensor-core work (its two spare warps per CTA
  // sustain ~270 GB/s, enough to arrive before the phase that needs them).
  s->first_wave = (W > 1 && s->push_order.size() > 1) ? s->push_order[0] : -1;
  {
    const char* e = std::getenv("MMCL_PACK_FANOUT_ALL");
    s->fanout_all = W > 1 && e && std::atoi(e) != 0;
  }

  // ---- backward plan ----
  bwd_begin(&s->bwd, s->d);
  size_t scratch_bytes = 0;
  if (s->need_grad && s->pair_bwd) {
    for (int p = 0; p < s->P; ++p)
      for (int side = 0; side < 2; ++side) {
        const int mx = side == 0 ? s->pairs[p].mod_a : s->pairs[p].mod_b;
        if (s->needs_grad[mx]) bwd_add_prob(&s->bwd, s->rows, s->N);
      }
    scratch_bytes = bwd_finish(&s->bwd, h->num_sms);
  }

  // ---- workspace ----
  size_t o = 0;
  o = align_up(o + kHdrWords * 8);
  s->off_out2 = o;
  o = align_up(o + size_t(s->P) * 8);
  s->po.resize(s->P);
  for (int p = 0; p < s->P; ++p) {
    s->po[p].vec = o;
    o += 5 * b * sizeof(float);  // contiguous [P][5][rows]
  }
  o = align_up(o);
  s->post_blocks_per_pair = int((b + 255) / 256);
  s->post_main_blocks = s->post_blocks_per_pair * s->P;
  s->off_blk = o;
  o = align_up(o + size_t(s->post_main_blocks) * 16);
  s->off_blkr = o;
  o = align_up(o + size_t(s->post_main_blocks) * 16);
  for (int p = 0; p < s->P; ++p) {
    s->po[p].range = o;
    o += 16;
  }
  o = align_up(o);
  s->idtab_size = 4 * next_pow2(2 * (W > 1 ? N : b));
  s->ws_idtab.assign(s->M, 0);
  int nkeys = 0;
  for (int m = 0; m < s->M; ++m)
    if (s->is_key[m]) {
      s->ws_idtab[m] = o;
      o = align_up(o + size_t(s->idtab_size) * 8);
      ++nkeys;
    }
  s->post_total_blocks =
      s->post_main_blocks + nkeys * int(((W > 1 ? N : b) + 255) / 256);
  s->off_fws = o;
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
  s->off_scratch = o;
  o = align_up(o + scratch_bytes);
  s->ws_bytes = o;
  // ---- arena ----
  s->ar_ids.assign(s->M, 0);
  s->ar_rows.assign(s->M, 0);
  s->ar_inv.assign(s->M, 0);
  if (W > 1) {
    size_t a = 0;
    s->ar_status = a;
    a = align_up(a + size_t(W) * 8);
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
    for (int p = 0; p < s->P; ++p) {
      s->po[p].range_all = a;
      a = align_up(a + size_t(W) * 16);
    }
    s->ar_bytes = a;
  }
  // ---- device counters + verification state ----
  DeviceGuard g(h->device);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->d_tickets), (kFlagChannels + 2) * 8);
  if (e == cudaSuccess) e = cudaMemset(s->d_tickets, 0, (kFlagChannels + 2) * 8);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&s->host), 16);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    if (s->d_tickets) cudaFree(s->d_tickets);
    if (s->host) cudaFreeHost(s->host);
    if (s->ev) cudaEventDestroy(s->ev);
    delete s;
    return cuda_fail(e, "mmcl_step_create");
  }
  // generation stamps start at a per-object salt so that two step objects whose workspaces
  // recycle the same memory never read each other's stamps as their own
  s->salt = unsigned(reinterpret_cast<uintptr_t>(s) >> 4) * 2654435761u;
  s->gen = s->salt | 1u;
  *out = s;
  return MMCL_OK;
}

int mmcl_step_destroy(mmcl_step* s) {
  if (!s) return MMCL_OK;
  DeviceGuard g(s->h->device);
  if (s->ev) cudaEventDestroy(s->ev);
  if (s->host) cudaFreeHost(s->host);
  if (s->d_tickets) cudaFree(s->d_tickets);
  delete s;
  return MMCL_OK;
}

size_t mmcl_step_workspace_bytes(const mmcl_step* s) { return s ? s->ws_bytes : 0; }
size_t mmcl_step_arena_bytes(const mmcl_step* s) { return s ? s->ar_bytes : 0; }

int mmcl_step_layout(const mmcl_step* s, mmcl_step_layout_t* out) {
  REQUIRE(s && out, "mmcl_step_layout: null argument");
  out->off_out2 = s->off_out2;
  out->off_vec = s->po[0].vec;
  out->off_info = 0;
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
      REQUIRE((reinterpret_cast<uintptr_t>(emb[m]) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(ids[m]) & 15) == 0,
              "mmcl_step_forward: embeddings and ids must be 16-byte aligned");
    }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  auto st = static_cast<cudaStream_t>(stream);
  Views v{static_cast<char*>(workspace), W > 1 ? symm_buffer(&h->arena, parity) : nullptr};
  const long long b = s->rows, N = s->N, d = s->d;
  const long long off = (long long)h->rank * b;
  cudaError_t e;
#define CU(call, what)                                   \
  do {                                                   \
    e = (call);                                          \
    if (e != cudaSuccess) return cuda_fail(e, what);     \
  } while (0)

  unsigned long long* hdr = hdr_ptr(v);
  s->gen += 2;  // (stays odd: never 0)
  const unsigned gen = s->gen;
  if (W > 1) s->seq = ++h->arena.step_seq;
  const unsigned long long status_ok = (s->seq << 1) | 1ull;
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(h->arena.local);
  float* out2 = reinterpret_cast<float*>(v.ws + s->off_out2);

  if (!local_ok) {
    // this rank's batch does not have the assumed structure: publish that, keep the flag
    // protocol going, fail verification (every rank sees the status word and fails with it)
    Range nv("mmcl.motions");
    MotionsLaunch ML{};
    ML.world = W;
    ML.rank = h->rank;
    ML.gen = gen;
    ML.hdr = hdr;
    if (W > 1) {
      ML.peers = h->arena.d_peers;
      ML.flags = flags;
      ML.buf_off = buf_off(h, parity);
      ML.status_off = s->ar_status;
      ML.status_val = s->seq << 1;
      ML.seq = s->seq;
      for (int m = 0; m < s->M; ++m) ML.chan_mask |= 1u << row_chan(m);
      ML.wait_mask = ML.chan_mask;
      if (s->need_grad) ML.chan_mask |= 1u << kChanLse;
    } else {
      ML.world = 0;
    }
    CU(launch_step_motions(ML, st), "mmcl_step_forward (motions)");
  } else {
    // ---- 1. pack ----
    {
      Range nv("mmcl.pack");
      PackLaunch PL{};
      PL.M = s->M;
      PL.P = s->P;
      PL.d = int(d);
      PL.normalize = s->normalize;
      PL.rows = b;
      PL.gen = gen;
      PL.hdr = hdr;
      for (int p = 0; p < s->P; ++p) {
        PL.pair_a[p] = s->pairs[p].mod_a;
        PL.pair_b[p] = s->pairs[p].mod_b;
      }
      for (int m = 0; m < s->M; ++m) {
        PackMod& pm = PL.mod[m];
        pm.src = emb[m];
        pm.dtype = s->dtype[m];
        pm.dst = reinterpret_cast<__nv_bfloat16*>(rows_ptr(s, v, m) + size_t(off) * d * 2);
        pm.inv = inv_ptr(s, v, m);
        if (pm.inv) pm.inv += off;
        pm.ids = ids[m];
        pm.ids_dst = (W > 1 && s->is_key[m])
                         ? reinterpret_cast<int64_t*>(v.ar + s->ar_ids[m] + size_t(off) * 16)
                         : nullptr;
        pm.fanout = (m == s->first_wave || s->fanout_all) ? 1 : 0;
        pm.off_rows = s->ar_rows[m] + size_t(off) * d * 2;
        pm.off_inv = s->ar_inv[m] + size_t(off) * 4;
        pm.off_ids = s->ar_ids[m] + size_t(off) * 16;
        pm.chan = row_chan(m);
      }
      if (W > 1) {
        PL.world = W;
        PL.rank = h->rank;
        PL.peers = h->arena.d_peers;
        PL.buf_off = buf_off(h, parity);
        PL.tickets = s->d_tickets;
        s->n_pack_launches += 1;
        PL.ticket_target = s->n_pack_launches * (unsigned long long)((b + 7) / 8);
        PL.seq = s->seq;
        PL.status_chan = -1;
        if (s->fanout_all) {
          // a channel every rank's forward kernel waits on (a key modality), released last
          int ms = s->push_order[0];
          for (int m : s->push_order)
            if (s->is_key[m]) ms = m;
          PL.status_chan = row_chan(ms);
          PL.status_off = s->ar_status;
          PL.status_val = status_ok;
          PL.all_ticket = s->d_tickets + kFlagChannels + 1;
          PL.all_target = PL.ticket_target * (unsigned long long)s->M;
        }
      }
      CU(launch_step_pack(PL, st), "mmcl_step_forward (pack)");
    }
    // ---- 2. fused similarity + log-sum-exp of every (pair, direction), and the all-gather ----
    int slot = 0;
    if (int rc = ensure_maps(s, v, parity, &slot)) return rc;
    FwdLaunch FL = s->fwd;
    const TensorMaps& fmaps = s->maps_f[slot];
    {
      Range nv("mmcl.fwd");
      float* fws = reinterpret_cast<float*>(v.ws + s->off_fws);
      for (int p = 0; p < s->P; ++p) {
        const mmcl_step_pair& pr = s->pairs[p];
        float* vec = reinterpret_cast<float*>(v.ws + s->po[p].vec);
        for (int dir = 0; dir < 2; ++dir)
          for (int part = 0; part < s->nsub_dir; ++part) {
            const int k = s->sub_idx[(p * 2 + dir) * 2 + part];
            const int mq = dir == 0 ? pr.mod_a : pr.mod_b;
            const int mk = dir == 0 ? pr.mod_b : pr.mod_a;
            FwdSub& S = FL.sub[k];
            S.map_q = 2 * mq;
            S.map_k = 2 * mk + 1;
            S.chan = row_chan(mk);
            S.scale = scales + pr.scale_index;
            S.qinv = inv_ptr(s, v, mq);
            if (S.qinv) S.qinv += off;
            S.kinv = inv_ptr(s, v, mk);
            S.part = fws + s->fwd_part_off[k];
            S.label_off = dir == 0 ? off : -1;
            S.pos = dir == 0 ? vec + 4 * b : nullptr;
          }
      }
      FL.err = hdr + kHdrErr;
      if (W > 1) {
        FL.world = W;
        FL.rank = h->rank;
        FL.rows_per_rank = b;
        FL.flags = flags;
        FL.seq = s->seq;
        FL.peers = h->arena.d_peers;
        FL.buf_off = buf_off(h, parity);
        FL.tickets = s->d_tickets;
        s->n_push_launches += 1;
        FL.ticket_target = s->n_push_launches * 2ull * (unsigned long long)fwd_grid(FL);
        FL.status_off = s->ar_status;
        FL.status_val = status_ok;
        FL.bad = hdr + kHdrBad;
        FL.nbad = s->P;
        FL.gen = gen;
        int np = 0;
        auto seg = [&](size_t o_, size_t bytes, int chan, int last) {
          PushSeg& sg = FL.push[np++];
          sg.off = o_;
          sg.bytes = (bytes + 15) & ~size_t(15);
          sg.chan = chan;
          sg.last_of_chan = last;
        };
        for (size_t i = 0; i < s->push_order.size(); ++i) {
          const int m = s->push_order[i];
          if (m == s->first_wave || s->fanout_all) continue;  // already in every rank's arena (pack kernel)
          const int ch = row_chan(m);
          // a channel shared by several modalities (m >= kMaxRowChannels) is released once, after
          // the last of them
          bool last_user = true;
          for (size_t j = i + 1; j < s->push_order.size(); ++j)
            if (row_chan(s->push_order[j]) == ch) last_user = false;
          if (FL.status_chan < 0) FL.status_chan = ch;
          if (s->is_key[m]) seg(s->ar_ids[m] + size_t(off) * 16, size_t(b) * 16, ch, 0);
          if (s->normalize) seg(s->ar_inv[m] + size_t(off) * 4, size_t(b) * 4, ch, 0);
          seg(s->ar_rows[m] + size_t(off) * d * 2, size_t(b) * d * 2, ch, last_user ? 1 : 0);
        }
        FL.npush = np;
      }
      CU(launch_fwd_fused(fmaps, FL, st), "mmcl_step_forward (fwd)");
    }
    // ---- 3. merge, loss, log-sum-exp exchange, id distinctness, verification word ----
    {
      Range nv("mmcl.post");
      PostLaunch QL{};
      QL.P = s->P;
      QL.world = W;
      QL.rank = h->rank;
      QL.local_ok = 1;
      QL.rows = b;
      QL.row_off = off;
      QL.blocks_per_pair = s->post_blocks_per_pair;
      QL.main_blocks = s->post_main_blocks;
      QL.total_blocks = s->post_total_blocks;
      QL.gen = gen;
      QL.hdr = hdr;
      QL.blk_part = reinterpret_cast<double*>(v.ws + s->off_blk);
      QL.blk_range = reinterpret_cast<float*>(v.ws + s->off_blkr);
      QL.out2 = out2;
      QL.ticket = s->d_tickets + kFlagChannels;
      s->n_post_launches += 1;
      QL.ticket_target = s->n_post_launches * (unsigned long long)s->post_total_blocks;
      QL.seq = s->seq;
      for (int p = 0; p < s->P; ++p) {
        const mmcl_step_pair& pr = s->pairs[p];
        PostPair& pp = QL.pair[p];
        for (int dir = 0; dir < 2; ++dir) {
          pp.nsub[dir] = s->nsub_dir;
          for (int part = 0; part < s->nsub_dir; ++part)
            pp.sub[dir][part] = s->sub_idx[(p * 2 + dir) * 2 + part];
        }
        pp.vec = reinterpret_cast<float*>(v.ws + s->po[p].vec);
        pp.coef = pr.weight / (2.0 * pr.loss_rows);
        pp.scale = scales + pr.scale_index;
        pp.push_lse = (W > 1 && s->need_grad && pr.has_col) ? 1 : 0;
        pp.lse_off[0] = s->po[p].lse_all[0];
        pp.lse_off[1] = s->po[p].lse_all[1];
        pp.range = reinterpret_cast<float*>(v.ws + s->po[p].range);
        pp.range_off = s->po[p].range_all;
      }
      if (W > 1) {
        QL.peers = h->arena.d_peers;
        QL.buf_off = buf_off(h, parity);
        QL.status_off = s->ar_status;
        QL.status_ok = status_ok;
        QL.release_lse = s->need_grad ? 1 : 0;
      }
      int nk = 0;
      for (int m = 0; m < s->M; ++m)
        if (s->is_key[m]) {
          PostIds& I = QL.ids[nk++];
          I.ids = W > 1 ? reinterpret_cast<const int64_t*>(v.ar + s->ar_ids[m]) : ids[m];
          I.n = W > 1 ? N : b;
          I.table = reinterpret_cast<unsigned long long*>(v.ws + s->ws_idtab[m]);
          I.table_size = s->idtab_size;
        }
      QL.nids = nk;
      CU(launch_step_post_fwd(FL, QL, st), "mmcl_step_forward (post)");
    }
  }
  if (s->all_reduce_loss) {
    ncclResult_t r = ncclAllReduce(out2, out2, size_t(s->P) * 2, ncclFloat, ncclSum, h->comm, st);
    if (r != ncclSuccess) return nccl_fail(r, "mmcl_step_forward (all-reduce)");
  }

  // ---- verification word read-back: the step's only host wait ----
  CU(cudaMemcpyAsync(s->host, hdr, 16, cudaMemcpyDeviceToHost, st),
     "mmcl_step_forward (verification copy)");
  CU(cudaEventRecord(s->ev, st), "mmcl_step_forward (event)");
  s->pending = true;

  // ---- the heavy half of backward does not depend on anything backward() receives ----
  if (launch_grad && s->need_grad && local_ok) return enqueue_grad_kernels(s, scales, v, parity, st);
  return MMCL_OK;
#undef CU
}

int mmcl_step_verify(mmcl_step* s, int* ok) {
  REQUIRE(s && ok, "mmcl_step_verify: null argument");
  REQUIRE(s->pending, "mmcl_step_verify: no forward in flight");
  cudaError_t e = cudaEventSynchronize(s->ev);
  s->pending = false;
  if (e != cudaSuccess) return cuda_fail(e, "mmcl_step_verify");
  if (s->host[kHdrErr] != 0)
    return fail(MMCL_ERR_CUDA,
                "mmcl_step_verify: a rank's data did not arrive within 4 s (peer-flag wait timed "
                "out); the step's results are invalid");
  *ok = s->host[kHdrVerify] == ((static_cast<unsigned long long>(s->gen) << 1) | 1ull) ? 1 : 0;
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
  return enqueue_grad_kernels(s, scales, v, parity, static_cast<cudaStream_t>(stream));
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
  Range nv("mmcl.unpack");
  auto st = static_cast<cudaStream_t>(stream);
  Views v{static_cast<char*>(workspace),
          h->world > 1 ? symm_buffer(&h->arena, parity) : nullptr};
  const long long b = s->rows, d = s->d;
  const long long off = (long long)h->rank * b;
  // problems that share (direct) go into one launch; direct and accumulating ones are separate
  // only because the caller zero-fills accumulation buffers itself -- one launch handles both
  UnpackLaunch UL{};
  UL.d = int(d);
  UL.normalize = s->normalize;
  UL.rows = b;
  UL.gout = gout;
  UL.ndc = s->bwd.ndc;
  UL.dp = s->bwd.dp;
  int k = 0, bk = 0;
  for (int p = 0; p < s->P; ++p) {
    const mmcl_step_pair& pr = s->pairs[p];
    for (int side = 0; side < 2; ++side) {
      const int mx = side == 0 ? pr.mod_a : pr.mod_b;
      const int my = side == 0 ? pr.mod_b : pr.mod_a;
      if (!s->needs_grad[mx]) continue;
      const int this_bk = bk++;  // index of the matching backward problem
      if (!grads[mx]) continue;
      UnpackProb& q = UL.prob[k++];
      q.dacc = reinterpret_cast<const float*>(v.ws + s->po[p].dacc[side]);
      q.partner = reinterpret_cast<const __nv_bfloat16*>(rows_ptr(s, v, my) + size_t(off) * d * 2);
      const float* pinv = inv_ptr(s, v, my);
      q.partner_inv = pinv ? pinv + off : nullptr;
      q.src = emb[mx];
      const float* inv = inv_ptr(s, v, mx);
      q.inv_norm = inv ? inv + off : nullptr;
      q.scale = scales + pr.scale_index;
      q.grad = grads[mx];
      q.weight = float(pr.weight);
      q.pos_coef = float(pr.pos_coef);
      q.src_dtype = s->dtype[mx];
      q.direct = grad_direct[mx] ? 1 : 0;
      q.grad_dtype = grad_direct[mx] ? s->dtype[mx] : MMCL_F32;
      q.blocks = int((b + 3) / 4);
      q.tile_begin = 0;
      q.nyt = s->pair_bwd ? s->bwd.prob[this_bk].nyt : 1;
      q.per_pair = s->pair_bwd ? s->bwd.prob[this_bk].per_pair : 1;
      q.scratch = s->pair_bwd
                      ? reinterpret_cast<const float*>(v.ws + s->off_scratch) +
                            size_t(this_bk) * size_t(s->bwd.pairs) * 128 * size_t(s->bwd.dp)
                      : nullptr;
    }
  }
  UL.nprob = k;
  cudaError_t e = launch_step_unpack(UL, st);
  if (e != cudaSuccess) return cuda_fail(e, "mmcl_step_backward (unpack)");
  return MMCL_OK;
}

}  // extern "C"
