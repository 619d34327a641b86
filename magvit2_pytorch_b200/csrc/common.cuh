// Shared helpers for libmagvit2_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <mutex>
#include "../../include/magvit2_b200.h"

namespace mv2 {

void set_error(const char* fmt, ...);

#define MV2_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) {                                                \
      mv2::set_error("%s:%d: argument check failed: %s", __FILE__, __LINE__, #cond); \
      return MV2_E_ARG;                                           \
    }                                                             \
  } while (0)

#define MV2_CHECK_LAUNCH()                                        \
  do {                                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) {                                     \
      mv2::set_error("%s:%d: CUDA launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return MV2_E_CUDA;                                          \
    }                                                             \
  } while (0)

#define MV2_CHECK_CUDA(expr)                                      \
  do {                                                            \
    cudaError_t e__ = (expr);                                     \
    if (e__ != cudaSuccess) {                                     \
      mv2::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return MV2_E_CUDA;                                          \
    }                                                             \
  } while (0)

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
// MV2_U8 sources (decoded video frames): the data loaders' normalisation x / 255 (reference data.py:103 ToTensor,
// data.py:188 `frames_torch /= 255.`) -- a correctly rounded fp32 division, so the result is bit-identical to theirs
template <> __device__ __forceinline__ float to_f32<uint8_t>(uint8_t v) { return __fdiv_rn((float)v, 255.f); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Activations.  expm1f / expf (not the fast intrinsics): the fp32 path must track the
// reference's libm-based CPU results to ~1 ulp.
__device__ __forceinline__ float act_elu(float v) { return v > 0.f ? v : expm1f(v); }
__device__ __forceinline__ float act_silu(float v) { return v / (1.f + expf(-v)); }
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == MV2_ACT_ELU) return act_elu(v);
  if (act == MV2_ACT_SILU) return act_silu(v);
  return v;
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

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- per-device one-time initialisation ----------------------------------------------------------------------
// cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in) applies to the CURRENT device only, so a process that
// drives several GPUs must repeat it on each of them: the flag is kept per device ordinal, not per process.
struct PerDeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  cudaError_t err[64] = {};
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (!done[dev]) { err[dev] = f(); done[dev] = true; }
    return err[dev];
  }
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// Every kernel of the library starts with pdl_wait() (everything before it -- barrier init, TMEM allocation, bias
// staging -- may overlap the tail of the previous kernel in the stream) and signals pdl_launch_dependents() right
// after, so the next kernel's CTAs move in as this kernel's CTAs retire.  Without the launch attribute both
// instructions are no-ops.  mv2_set_pdl(1) turns the attribute on for all launches.
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline void launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (g_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through MV2_CHECK_LAUNCH
}
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_kc(kernel, grid, block, smem, st, 1, static_cast<Args&&>(args)...);
}

}  // namespace mv2
