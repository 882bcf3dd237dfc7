// Shared helpers for the rcgan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rcgan_b200.h"

#define RCGAN_NUM_SMS 148

void rcgan_set_error(const char* fmt, ...);
void rcgan_set_conv_variant(const char* fmt, ...);
void rcgan_count_launch();  // every kernel launch of this library bumps rcgan_launch_count()

#define RCGAN_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      rcgan_set_error(__VA_ARGS__);           \
      return RCGAN_EBADSHAPE;                 \
    }                                         \
  } while (0)

#define RCGAN_LAUNCH_CHECK(name)                                                     \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      rcgan_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return RCGAN_ECUDA;                                                            \
    }                                                                                \
    rcgan_count_launch();                                                            \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL).  A training step is hundreds of short dependent kernels on one stream (422 per
// MNIST iteration, ~12 us each), so the launch gap between them is a first-order cost.  Every kernel of the library is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with pdl_sync(): griddepcontrol.wait blocks
// until the previous kernel has fully completed and flushed (so no dependency analysis is needed -- nothing is read or
// written earlier), griddepcontrol.launch_dependents lets the NEXT kernel's CTAs be scheduled while this one drains.
// The conv kernels run their prologue (barrier init, TMEM allocation, descriptor prefetch) before the wait.
// Captured into the step's CUDA graph as programmatic edges.  RCGAN_PDL=0 launches without the attribute.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool rcgan_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = rcgan_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float act_fwd(float x, int act, float leak) {
  switch (act) {
    case RCGAN_ACT_RELU: return fmaxf(x, 0.f);
    case RCGAN_ACT_LRELU: return fmaxf(x, leak * x);
    case RCGAN_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case RCGAN_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
// derivative expressed through the OUTPUT y = act(x)
__device__ __forceinline__ float act_bwd_from_y(float y, int act, float leak) {
  switch (act) {
    case RCGAN_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case RCGAN_ACT_LRELU: return y > 0.f ? 1.f : leak;
    case RCGAN_ACT_SIGMOID: return y * (1.f - y);
    case RCGAN_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in every thread; `sh` needs 33 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? sh[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
static inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }
