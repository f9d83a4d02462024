// Shared helpers for the omnitok_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/omnitok_b200.h"

namespace omt {

void set_error(const char* fmt, ...);
int check_device();   // OMT_OK when the current device is sm_100; caches per device
int sm_count();

#define OMT_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      omt::set_error(__VA_ARGS__);              \
      return OMT_E_ARG;                         \
    }                                           \
  } while (0)

#define OMT_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      omt::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                     __FILE__, __LINE__);                                         \
      return OMT_E_CUDA;                                                          \
    }                                                                             \
  } while (0)

#define OMT_ENTER()                      \
  do {                                   \
    int rc__ = omt::check_device();      \
    if (rc__ != OMT_OK) return rc__;     \
  } while (0)

#define OMT_LAUNCH_CHECK() OMT_CUDA(cudaGetLastError())

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every kernel calls pdl_sync() before its first global-memory access: griddepcontrol.wait blocks until the
// previous kernel in the stream has completed and flushed (a no-op when the launch carried no PDL attribute);
// launch_dependents then lets the NEXT kernel's CTAs be scheduled as soon as all of ours are resident, so its
// prologue (barrier init, TMEM alloc, descriptor prefetch, launch latency) overlaps our last wave.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
extern int g_pdl;   // omt_set_option("pdl", 0|1)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// logical GEMM row -> physical row (see include/omnitok_b200.h)
__host__ __device__ __forceinline__ long long map_row(int r, int seg, int seg_stride, int seg_off) {
  if (seg <= 0) return r;
  return (long long)(r / seg) * seg_stride + seg_off + (r % seg);
}

struct GemmArgs {
  const float* A; int lda; int a_seg, a_seg_stride, a_seg_off;
  const float* W;
  float* C; int ldc; int c_seg, c_seg_stride, c_seg_off;
  int M, N, K;
  const float* bias; const float* residual; int ldr;
  const float* A2; int n_split;   // dual-A form: output columns >= n_split are computed from A2 (same lda / row map)
  // OMT_EPI_QKV: heads (64 columns) below qk_cols get rope + l2norm + per-dim scale in the epilogue
  const float* rope_cos; const float* rope_sin; const float* q_scale; const float* k_scale;
  int qk_cols; int tokens;        // rope position of row m is m % tokens; columns < qk_cols/2 use q_scale, the rest k_scale
};

}  // namespace omt
