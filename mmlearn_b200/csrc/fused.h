// Launch descriptors of the fused step (step.cu): one forward launch covers every (pair,
// direction) of the loss, one backward launch every (pair, side); with several ranks the forward
// launch also carries the all-gather (peer stores by otherwise idle warps) and both launches wait
// on per-source-rank ready flags instead of a global fence.  Internal; not part of the ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mmcl {

// ---- symmetric-arena flag words ----------------------------------------------------------------
// [channel][source rank] 64-bit words at the start of every rank's arena allocation.  A producer
// stores its sequence number into word [channel][my_rank] of EVERY rank (release, system scope)
// after its data; a consumer spins on its OWN copy of [channel][source] (acquire, system scope).
constexpr int kMaxRanks = 32;
constexpr int kFlagChannels = 16;
constexpr size_t kFlagBytes = size_t(kFlagChannels) * kMaxRanks * 8;  // 4 KB
constexpr int kChanFence = 0;   // mmcl_symm_fence (channels 0..3 are fence channels)
constexpr int kChanLse = 4;     // log-sum-exp vectors of a step (written by the post-forward kernel)
constexpr int kChanRows = 5;    // kChanRows + m: rows / norms / ids of modality m (m clamped)
constexpr int kMaxRowChannels = kFlagChannels - kChanRows;

// Device-side error word (workspace): 0 = fine; set by a spin loop that timed out.
constexpr unsigned long long kErrFlagTimeout = 1;

// ---- forward ---------------------------------------------------------------------------------
constexpr int kMaxFwdSubs = 16;
constexpr int kMaxFwdMaps = 12;  // per modality: Q box (128 rows) and K box (256 rows)

struct FwdSub {  // one direction of one pair, restricted to a rotated range of key tiles
  int nq, nk, nqt;
  int nkt_all, kt0, nkt_sub;  // key tile j in [0, nkt_sub) is global tile (kt0 + j) % nkt_all
  long long tile_begin;       // position in the launch's linear tile space
  long long phase_begin;      // first tile of the phase this sub-problem belongs to
  int tpc;                    // tiles per CTA of that phase
  int map_q, map_k;
  int max_slots;
  int chan;                   // flag channel that guards remote rows of K (world > 1)
  long long label_off;        // key row of query row 0's positive; < 0: not extracted
  const float* scale;
  const float* qinv;
  const float* kinv;
  float* part;                // [3][max_slots][nqt * 128] partial (max, sum, weighted sum)
  float* pos;                 // [nq] positive logits (written by the tile that holds them) or null
  const float* thr;           // rank mode
  const long long* tgt;       // rank mode
  float* out;                 // store mode: [nq, out_ld] fp32 logits
  long long out_ld;
  float out_scale;
  long long row_base;         // BCE mode: key index of query row 0
  const int* band_hi;         // BCE mode: [nk] end of every key row's positive band
};

constexpr int kMaxPushSegs = 16;
struct PushSeg {
  unsigned long long off;    // byte offset inside the arena buffer (same on every rank)
  unsigned long long bytes;  // multiple of 16
  int chan;                  // flag channel released once every segment of the channel is out
  int last_of_chan;          // 1: release after this segment
};

// The tile space is a sequence of PHASES; every CTA takes an equal contiguous share of each phase
// and walks the phases in order ("local keys of every direction", then "remote keys of direction
// a", ...), so all SMs work on what is available first.
constexpr int kMaxFwdPhases = 8;
struct FwdPhase {
  long long tile_begin, total;
  int tiles_per_cta;
};

struct FwdLaunch {
  int nsub, kchunks, nstages, nphase, ctas;
  long long total_tiles;
  FwdPhase phase[kMaxFwdPhases];
  // several ranks
  int world, rank;
  long long rows_per_rank;              // source rank of key row g = g / rows_per_rank
  const unsigned long long* flags;      // this rank's flag words
  unsigned long long seq;               // value the flags must reach
  unsigned long long* err;              // device error word
  // peer pushes by warps 2-3 (world > 1)
  int npush;
  char* const* peers;                   // arena base of every rank
  unsigned long long buf_off;           // arena buffer offset (flags + parity * bytes)
  unsigned long long* tickets;          // [kFlagChannels] monotonic completion counters
  unsigned long long ticket_target;     // counter value after this launch's last push warp
  unsigned long long status_off;        // arena-buffer offset of status[world] (8 bytes each)
  unsigned long long status_val;        // this rank's status word for the step (ok form; bit 0
                                        // is cleared when the pack kernel flagged a pair)
  int status_chan;                      // channel whose release publishes the status word
  const unsigned long long* bad;        // [nbad] words that equal `gen` for a flagged pair
  int nbad;
  unsigned gen;
  PushSeg push[kMaxPushSegs];
  FwdSub sub[kMaxFwdSubs];
};

constexpr int kMaxMaps = 18;  // >= kMaxFwdMaps, kMaxBwdMaps
struct alignas(64) TensorMaps {
  CUtensorMap tm[kMaxMaps];
};
constexpr int kStepMaxMods = 6;   // modalities / pairs one fused step handles (more: general path)
constexpr int kStepMaxPairs = 4;

// ---- backward --------------------------------------------------------------------------------
constexpr int kMaxBwdProbs = 8;
constexpr int kMaxBwdMaps = 18;  // per modality: X box (128 rows), Y box (256 rows), Y box (64 rows)

struct BwdProb {  // one side of one pair: dX = sum_j G_ij yinv_j Y_j
  int nx, ny;
  int nxt, nyt;
  long long tile_begin;  // position in the launch's linear (x tile, d pass, y tile) space
  long long tiles;       // nxt * ndc * nyt
  int per_pair;          // tile steps of this problem per CTA pair (every pair takes an equal
                         // share of every problem, one problem after the other: the problems'
                         // key matrices then pass through L2 one at a time)
  int map_x, map_y256, map_y64;
  const float* scale;
  const float* xinv;
  const float* yinv;
  const float* off_row;
  const float* off_col;
  float off_mul, row_add, col_add;
  float* dx;             // [nx, d] fp32
  // (min, max) of the raw row offsets and, per rank, of the raw column offsets, when the caller
  // already knows them (the step's post-forward kernel does); null: the kernel scans the vectors
  const float* rng_row;  // [2]
  const float* rng_col;  // [rng_col_count][rng_col_stride floats], (min, max) first
  int rng_col_stride, rng_col_count;
  // binary-cross-entropy form of G (modality-alignment term, contrastive.py:344-413):
  //   G_ij = sigmoid(x_ij) (a_ij + a_ji) - (t_ij p_i + t_ji p_j),  t_ij = j in [i, band_hi[i]),
  //   a_ij = t_ij ? p_i : n_i,  p_i = 1 / (T |band_i|),  n_i = 1 / (T (T - |band_i|))
  int bce;
  long long row_base;    // index of x row 0 among the y rows
  const int* band_hi;    // [ny]
};

struct BwdLaunch {
  int nprob, d, kchunks, ndc, dp, dcc;
  long long total;
  int pairs;             // CTA pairs of the launch
  int accumulate;        // 1: red.add into dx (caller zero-fills); 0: stores + scratch (below)
  float* scratch;        // [problem][cta pair][128][dp] fp32: head segments that continue an item
  // several ranks: off_col of remote sources becomes valid when flags[kChanLse][r] >= seq
  int world, rank;
  const unsigned long long* flags;
  unsigned long long seq;
  unsigned long long* err;
  BwdProb prob[kMaxBwdProbs];
};

}  // namespace mmcl
