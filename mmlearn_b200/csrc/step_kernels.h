// Launch descriptors of the fused step's small kernels (step_kernels.cu).  Internal.
#pragma once
#include <cuda_bf16.h>

#include "fused.h"

namespace mmcl {

// workspace header: 64-bit words at the start of the caller's workspace
constexpr int kHdrVerify = 0;  // gen << 1 | ok, written by the post-forward / motions kernel
constexpr int kHdrErr = 1;     // device error word (kErrFlagTimeout)
constexpr int kHdrDup = 2;     // == gen: a duplicate example id was found among the key ids
constexpr int kHdrBad = 3;     // [kStepMaxPairs] == gen: ids of the pair differ row by row
constexpr int kHdrWords = 16;

struct PackMod {
  const void* src;         // [rows, d] embedding tensor of the caller
  int dtype;               // mmcl_dtype
  __nv_bfloat16* dst;      // this rank's slot of the gathered rows
  float* inv;              // this rank's slot of the inverse norms (null without l2_normalize)
  const int64_t* ids;      // [rows, 2]
  int64_t* ids_dst;        // this rank's slot of the gathered ids, or null
  // first-wave modality (several ranks): the pack kernel itself stores rows / norms / ids into
  // EVERY rank's arena (full-GPU store rate) and releases the modality's ready flags
  int fanout;
  unsigned long long off_rows, off_inv, off_ids;  // arena-buffer offsets of this rank's slots
  int chan;
};

struct PackLaunch {
  int M, P, d, normalize;
  long long rows;
  unsigned gen;
  unsigned long long* hdr;
  int pair_a[kStepMaxPairs], pair_b[kStepMaxPairs];
  // fan-out of first-wave modalities
  int world, rank;
  char* const* peers;
  unsigned long long buf_off;
  unsigned long long* tickets;        // [kFlagChannels]
  unsigned long long ticket_target;   // pack launches so far x blocks per modality
  unsigned long long seq;
  // Every modality fanned out by this kernel (no pushes left for the forward kernel): the block
  // that finishes last OVERALL publishes this rank's status word and only then releases the flags
  // of `status_chan` (the consumers read the status after they have seen that channel).
  int status_chan;                    // -1: the forward kernel's push warps publish the status
  unsigned long long status_off, status_val;
  unsigned long long* all_ticket;     // counts the blocks of all modalities
  unsigned long long all_target;
  PackMod mod[kStepMaxMods];
};

struct PostPair {
  int nsub[2];       // forward sub-problems of direction a / b (local + remote key ranges)
  int sub[2][2];
  float* vec;        // [5][rows]: lse_a, lse_b, esum_a, esum_b, pos
  double coef;       // weight / (2 * loss_rows)
  const float* scale;
  int push_lse;
  unsigned long long lse_off[2];  // arena-buffer offsets of the gathered lse_a / lse_b
  float* range;                   // [4] (min, max) of lse_a, (min, max) of lse_b of this rank
  unsigned long long range_off;   // arena-buffer offset of range[world][4] (push_lse)
};

struct PostIds {
  const int64_t* ids;
  long long n;
  unsigned long long* table;
  unsigned long long table_size;  // power of two
};

struct PostLaunch {
  int P, nids, world, rank, local_ok, release_lse;
  long long rows, row_off;
  int blocks_per_pair, main_blocks, total_blocks;
  unsigned gen;
  unsigned long long* hdr;
  double* blk_part;   // [main blocks][2] loss / esum partial sums
  float* blk_range;   // [main blocks][4] (min, max) of lse_a, lse_b
  float* out2;
  unsigned long long* ticket;
  unsigned long long ticket_target;
  char* const* peers;
  unsigned long long buf_off, status_off, status_ok, seq;
  PostPair pair[kStepMaxPairs];
  PostIds ids[kStepMaxMods];
};

struct MotionsLaunch {
  int world, rank;
  unsigned gen;
  unsigned chan_mask, wait_mask;
  char* const* peers;
  const unsigned long long* flags;
  unsigned long long buf_off, status_off, status_val, seq;
  unsigned long long* hdr;
};

struct UnpackProb {
  const float* dacc;
  const __nv_bfloat16* partner;
  const float* partner_inv;
  const void* src;
  const float* inv_norm;
  const float* scale;
  void* grad;
  float weight, pos_coef;
  int src_dtype, grad_dtype, direct;
  int blocks;
  long long tile_begin;  // of the matching backward problem (relative to its own start: 0)
  int nyt, per_pair;
  const float* scratch;  // that problem's scratch tiles ([cta pair][128][dp]) or null
};

struct UnpackLaunch {
  int nprob, d, normalize;
  long long rows;
  const float* gout;
  int ndc, dp;
  UnpackProb prob[kMaxBwdProbs];
};

cudaError_t launch_step_pack(const PackLaunch& L, cudaStream_t stream);
cudaError_t launch_step_post_fwd(const FwdLaunch& F, const PostLaunch& L, cudaStream_t stream);
cudaError_t launch_step_motions(const MotionsLaunch& L, cudaStream_t stream);
cudaError_t launch_step_wait_flags(const unsigned long long* flags, int chan,
                                   unsigned long long seq, int world, int rank,
                                   unsigned long long* err, cudaStream_t stream);
cudaError_t launch_step_unpack(const UnpackLaunch& L, cudaStream_t stream);

}  // namespace mmcl
