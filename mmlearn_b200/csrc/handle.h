// The handle behind include/mmcl.h and the error helpers shared by api.cu and step.cu
// (internal; not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <mutex>
#include <string>

#include "../../include/mmcl.h"
#include "kernels.h"

struct mmcl_handle {
  int device = 0;
  int rank = 0;
  int world = 1;
  int num_sms = 0;
  ncclComm_t comm = nullptr;
  mmcl::SymmArena arena;
  std::mutex mu;
};

namespace mmcl {

std::string& last_error_slot();  // thread-local message behind mmcl_last_error()

inline int fail(int code, const std::string& msg) {
  last_error_slot() = msg;
  return code;
}
inline int cuda_fail(cudaError_t e, const char* where) {
  return fail(MMCL_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
inline int nccl_fail(ncclResult_t r, const char* where) {
  return fail(MMCL_ERR_NCCL, std::string(where) + ": " + ncclGetErrorString(r));
}

#define REQUIRE(cond, msg)                                    \
  do {                                                        \
    if (!(cond)) return mmcl::fail(MMCL_ERR_INVALID, msg);    \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace mmcl
