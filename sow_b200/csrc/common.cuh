// Host-side plumbing shared by the C-ABI translation units: error reporting, device checks, and the
// TMA tensor-map encoder (driver entry point fetched at run time so that the library loads on GPU-less hosts).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sow_b200.h"

namespace sowb {

int set_error(int code, const char* fmt, ...);

#define SOWB_CHECK_CUDA(expr)                                                                          \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return ::sowb::set_error(SOWB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                               __FILE__, __LINE__);                                                    \
  } while (0)

#define SOWB_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::sowb::set_error(SOWB_EINVAL, __VA_ARGS__); \
  } while (0)

// Number of SMs of the current device (cached per device).
int num_sms();
// 0 if the current device is sm_100, else an error code (message set).
int require_sm100();

// Bind the primary context of the device owning `ptr` to the calling thread if the thread has none (see common.cu).
int ensure_context_for(const void* ptr);

// 2-D row-major tensor map with 128B swizzle.  `inner` is the contiguous dimension (elements), `outer` the row
// count, `pitch_bytes` the row pitch; box = (box_inner x box_outer) elements.  elem_bytes: 2 (bf16) or 4 (fp32).
int make_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                       uint32_t box_inner, uint32_t box_outer, int elem_bytes);

// ------------------------------------------------------------------------------------------------
// Live per-kernel timing (bench.py roofline): when enabled, launches are bracketed by CUDA events recorded on the
// launching stream; sow_profile_read() sums elapsed time / algorithmic work per kernel class.
// ------------------------------------------------------------------------------------------------
enum ProfClass : int {
  PROF_GEMM_FWD = 0,     // y  = x.W + t.B          (flops)
  PROF_GEMM_DX = 1,      // dX = dY.W^T + dt.A^T    (flops)
  PROF_GEMM_SKINNY = 2,  // t_cat = x.[A_q|A_k|..]  (flops)
  PROF_GEMM_SPLITK = 3,  // dA_cat = x^T.dt_cat     (flops)
  PROF_MERGE = 4,        // grouped merge           (bytes)
  PROF_ADAM = 5,         // multi-tensor Adam       (bytes)
  PROF_GEMM_K2 = 6,      // fused dt + dB pass over dY (flops)
  PROF_NUM = 7,
};
bool profile_enabled();
struct ProfileScope {
  cudaStream_t stream;
  cudaEvent_t start = nullptr;
  int klass;
  double work;
  ProfileScope(cudaStream_t s, int klass, double work);
  ~ProfileScope();
};

__host__ __device__ static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
__host__ __device__ static inline int ceil_div(int x, int m) { return (x + m - 1) / m; }

}  // namespace sowb
