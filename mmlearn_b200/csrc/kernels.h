// Internal launcher declarations shared by the .cu files and api.cu (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "fused.h"

namespace mmcl {

// process-wide count of kernel launches (diagnostic, see mmcl_kernel_launches)
void count_launches(int n);
long long launches_so_far();

// sim_fwd.cu
size_t sim_lse_workspace_bytes(int num_sms, long long nq, long long nk, long long d);
cudaError_t launch_sim_lse_fwd(int num_sms, const void* q, long long nq, const void* k,
                               long long nk, long long d, const float* scale, const float* qinv,
                               const float* kinv, float* lse, float* esum, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream, const char** err);

cudaError_t launch_sim_rank_fwd(int num_sms, const void* q, long long nq, const void* k,
                                long long nk, long long d, const float* qinv, const float* kinv,
                                const float* thr, const long long* tgt, long long* rank,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream,
                                const char** err);

cudaError_t launch_sim_logits_fwd(int num_sms, const void* q, long long nq, const void* k,
                                  long long nk, long long d, float scale, const float* qinv,
                                  const float* kinv, float* out, long long out_ld,
                                  cudaStream_t stream, const char** err);

size_t sim_bce_workspace_bytes(int num_sms, long long nq, long long nk, long long d);
cudaError_t launch_sim_bce_fwd(int num_sms, const void* q, long long nq, const void* k,
                               long long nk, long long d, const float* scale, const float* qinv,
                               const float* kinv, long long row_base, const int* band_hi,
                               float* stats, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream, const char** err);

// sim_fwd.cu, fused launches: fwd_begin, fwd_add_sub per (pair, direction[, key range]),
// fwd_finish (CTA split + partial-statistics sizes), then fill the pointers and launch
void fwd_begin(FwdLaunch* L, long long d);
int fwd_add_sub(FwdLaunch* L, long long nq, long long nk, int kt0, int nkt_sub, int phase);
size_t fwd_finish(FwdLaunch* L, int num_sms, size_t* part_off_floats);
int fwd_grid(const FwdLaunch& L);
int fwd_make_maps(TensorMaps* maps, int slot, const void* rows, long long n, long long d);
cudaError_t launch_fwd_fused(const TensorMaps& maps, const FwdLaunch& L, cudaStream_t stream);

// sim_bwd_pair.cu, fused launches
bool bwd_pair_supported(int num_sms, long long d);
void bwd_begin(BwdLaunch* L, long long d);
int bwd_add_prob(BwdLaunch* L, long long nx, long long ny);
size_t bwd_finish(BwdLaunch* L, int num_sms);
int bwd_cta_pairs(const BwdLaunch& L);
int bwd_make_maps(TensorMaps* maps, int slot, const void* rows, long long n, long long d);
cudaError_t launch_bwd_fused(const TensorMaps& maps, const BwdLaunch& L, cudaStream_t stream);

cudaError_t launch_sim_bce_bwd_pair(int num_sms, const void* x, long long nx, const void* y,
                                    long long ny, long long d, const float* scale,
                                    const float* xinv, const float* yinv, long long row_base,
                                    const int* band_hi, float* dx, cudaStream_t stream,
                                    const char** err);

// sim_bwd.cu
cudaError_t launch_sim_grad_bwd(int num_sms, const void* x, long long nx, const void* y,
                                long long ny, long long d, const float* scale, const float* xinv,
                                const float* yinv, const float* off_row, const float* off_col,
                                float off_mul, float row_add, float col_add, float* dx,
                                cudaStream_t stream, const char** err);

// sim_bwd_pair.cu (cudaErrorNotSupported = shape does not fit, nothing launched)
cudaError_t launch_sim_grad_bwd_pair(int num_sms, const void* x, long long nx, const void* y,
                                     long long ny, long long d, const float* scale,
                                     const float* xinv, const float* yinv, const float* off_row,
                                     const float* off_col, float off_mul, float row_add,
                                     float col_add, float* dx, cudaStream_t stream,
                                     const char** err);

// aux_kernels.cu
cudaError_t launch_pack_rows(const void* src, int src_dtype, long long src_rows, long long d,
                             const int64_t* idx, long long n_out, int normalize, void* dst_bf16,
                             float* inv_norm, cudaStream_t stream);
cudaError_t launch_positive_logits(const void* q, long long nq, const void* k, long long nk,
                                   long long d, long long label_offset, const float* scale,
                                   const float* qinv, const float* kinv, float* pos,
                                   cudaStream_t stream);
cudaError_t launch_logit_scale(float* log_s, long long n, double max_scale, float* s_out,
                               cudaStream_t stream);
cudaError_t launch_indexed_similarity(const void* q, long long nq, const void* k, long long nk,
                                      long long d, const int64_t* index, const float* qinv,
                                      const float* kinv, float add, float* out,
                                      cudaStream_t stream);
cudaError_t launch_pair_loss(const float* lse_a, const float* lse_b, const float* esum_a,
                             const float* esum_b, const float* pos, long long n, double weight,
                             double n_norm, const float* scale, float* out2, cudaStream_t stream);
cudaError_t launch_make_offsets(const float* lse, long long n, const float* denom_vec,
                                double denom_scalar, float* off, cudaStream_t stream);
cudaError_t launch_unpack_grad(const float* dacc, const void* partner_bf16,
                               const float* partner_inv, long long n,
                               long long d, double weight, double pos_coef, const float* gout,
                               const float* scale, const void* src, int src_dtype,
                               const float* inv_norm, int normalize, const int64_t* idx,
                               void* grad, int grad_dtype, int direct, cudaStream_t stream);

// match.cu
size_t match_workspace_bytes(long long n, long long m);
cudaError_t launch_match_ids(const int64_t* ids_a, long long n, const int64_t* ids_b, long long m,
                             int64_t* out_a, int64_t* out_b, long long capacity, int64_t* out_info,
                             void* workspace, size_t workspace_bytes, cudaStream_t stream);

// symm.cu: symmetric peer-memory arena (CUDA IPC over NVLink)
constexpr int kFenceChannels = 4;
struct SymmArena {
  char* local = nullptr;       // [flags 1 KB][buffer 0][buffer 1]
  std::vector<char*> peers;    // base of every rank's allocation (own entry = local)
  char** d_peers = nullptr;    // the same table on the device
  void* scratch = nullptr;     // setup-time staging (IPC handle exchange, barriers)
  size_t bytes = 0;            // usable bytes per buffer
  unsigned long long seq[kFlagChannels] = {};  // sequence number per flag channel (fused.h)
  unsigned long long step_seq = 0;  // sequence number of the fused step's row / lse channels
  bool failed = false;
};
int symm_reserve(SymmArena* a, void* nccl_comm, int rank, int world, size_t bytes,
                 cudaStream_t stream, std::string* err);
void symm_destroy(SymmArena* a, int rank);
char* symm_buffer(const SymmArena* a, int parity);
// skip_rank >= 0: that rank's buffer is not written (its copy is the source itself)
cudaError_t launch_symm_put(const SymmArena* a, int world, int parity, const void* src,
                            size_t bytes, size_t dst_off, cudaStream_t stream, int skip_rank = -1);
cudaError_t launch_pack_rows_symm(const SymmArena* a, int world, int parity, const void* src,
                                  int src_dtype, long long src_rows, long long d,
                                  const int64_t* idx, long long n_out, int normalize,
                                  size_t off_rows, long long off_inv, cudaStream_t stream);
cudaError_t debug_push(SymmArena* a, int world, int rank, int parity, size_t off, size_t bytes,
                       int mode, int blocks, int threads, cudaStream_t stream);
// Fences of different channels use different flag words and may overlap (different streams).
cudaError_t launch_symm_fence(SymmArena* a, int world, int rank, cudaStream_t stream,
                              int channel = 0);

size_t identity_check_workspace_bytes(long long m);
cudaError_t launch_identity_check(const int64_t* ids_a, long long n, const int64_t* ids_b,
                                  long long m, int64_t* out_info, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream);

}  // namespace mmcl
