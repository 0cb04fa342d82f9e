// Shared device helpers for libyad.so (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/yad.h"

typedef __nv_bfloat16 bf16;

// ---- host side error plumbing -------------------------------------------------------------------
void yad_set_error(const char* fmt, ...);
#define YAD_CHECK(cond, ...)          \
  do {                                \
    if (!(cond)) {                    \
      yad_set_error(__VA_ARGS__);     \
      return 1;                       \
    }                                 \
  } while (0)
#define YAD_LAUNCH_CHECK(name)                                              \
  do {                                                                      \
    cudaError_t _e = cudaGetLastError();                                    \
    if (_e != cudaSuccess) {                                                \
      yad_set_error("%s: launch failed: %s", name, cudaGetErrorString(_e)); \
      return 2;                                                             \
    }                                                                       \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the library starts with pdl_sync() (kernels with an on-chip prologue -- barrier init, TMEM allocation, tensor-map prefetch --
// call it after that prologue): `griddepcontrol.wait` blocks until the grids this launch depends on have completed and their memory is visible,
// `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream be scheduled while this one is still running, so that its launch latency
// and prologue hide under this kernel's tail.  Launches go through YAD_LAUNCH, which sets the programmatic-stream-serialization attribute
// (a no-op when the predecessor is not a kernel).  Correct by construction: no kernel touches global memory before its own wait, and a wait
// covers the whole chain because each predecessor finished only after its own wait.  The attribute is OFF by default (measured neutral); yad_set_pdl(1) / YAD_PDL=1 turns it on.
int yad_pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_trigger();
}
template <typename... KArgs, typename... Args>
inline cudaError_t yad_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = yad_pdl_enabled();
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}
#define YAD_LAUNCH(kernel, grid, block, smem, st, ...) yad_launch(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)

// ---- vector access: 8 consecutive channels ------------------------------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename T>
__device__ __forceinline__ float round_to(float v);  // v rounded to the storage type T
template <>
__device__ __forceinline__ float round_to<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_to<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- activations (match torch fp32 semantics) ----------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case YAD_ACT_SILU: return x * sigmoidf_(x);
    case YAD_ACT_RELU: return fmaxf(x, 0.0f);
    case YAD_ACT_SIGMOID: return sigmoidf_(x);
    case YAD_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case YAD_ACT_HARDSWISH: return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    default: return x;
  }
}

// fast variants for the fused epilogues (ex2.approx / rcp.approx: ~2 ulp, far below bf16 and the 1e-3 fp32 bound) with the activation
// switch hoisted out of the element loop: one warp-uniform branch per N values instead of one per value
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
template <int N>
__device__ __forceinline__ void apply_act_n(float (&v)[N], int act) {
  switch (act) {
    case YAD_ACT_SILU:
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = v[i] * sigmoid_fast(v[i]);
      break;
    case YAD_ACT_RELU:
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = fmaxf(v[i], 0.0f);
      break;
    case YAD_ACT_SIGMOID:
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = sigmoid_fast(v[i]);
      break;
    case YAD_ACT_GELU:
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752440f));
      break;
    case YAD_ACT_HARDSWISH:
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = v[i] * fminf(fmaxf(v[i] + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
      break;
    default:
      break;
  }
}

// ---- reductions ------------------------------------------------------------------------------------
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
// block-wide sum; `red` is >= 32 floats of shared memory; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// ---- fused conv epilogue on 4 consecutive output channels ------------------------------------------
template <typename T>
__device__ __forceinline__ void epilogue4(float (&acc)[4], const yad_epilogue& e, int img, int64_t pix, int co) {
  float s = 1.0f;
  if (e.img_scale) s = e.img_scale[img];
  if (e.pix_scale) s *= ld1(reinterpret_cast<const T*>(e.pix_scale) + pix * e.pix_scale_ld);
  float b[4] = {0.f, 0.f, 0.f, 0.f};
  if (e.bias) {
    float4 bb = *reinterpret_cast<const float4*>(e.bias + co);
    b[0] = bb.x; b[1] = bb.y; b[2] = bb.z; b[3] = bb.w;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) acc[i] = acc[i] * s + b[i];
  apply_act_n<4>(acc, e.act);
#pragma unroll
  for (int i = 0; i < 4; i++) acc[i] *= e.alpha;
  if (e.mul) {
    float m[4];
    load4(reinterpret_cast<const T*>(e.mul) + pix * e.mul_ld + co, m);
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] *= m[i];
  }
  if (e.gate_h) {  // separable gate; the product is rounded to the activation dtype first, as the materialised gate map would have been
    const int64_t rem = pix - (int64_t)img * e.gate_hm * e.gate_wm;
    const int oy = (int)(rem / e.gate_wm), ox = (int)(rem - (int64_t)oy * e.gate_wm);
    float gh[4], gw[4];
    load4(reinterpret_cast<const T*>(e.gate_h) + ((int64_t)img * e.gate_hm + oy) * e.gate_ld + co, gh);
    load4(reinterpret_cast<const T*>(e.gate_w) + ((int64_t)img * e.gate_wm + ox) * e.gate_ld + co, gw);
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] *= round_to<T>(gh[i] * gw[i]);
  }
  if (e.add) {
    float a[4];
    load4(reinterpret_cast<const T*>(e.add) + pix * e.add_ld + co, a);
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] += a[i];
  }
}

#define YAD_DISPATCH_DTYPE(dtype, ...)                    \
  if ((dtype) == YAD_F32) {                               \
    typedef float T;                                      \
    __VA_ARGS__                                           \
  } else if ((dtype) == YAD_BF16) {                       \
    typedef bf16 T;                                       \
    __VA_ARGS__                                           \
  } else {                                                \
    yad_set_error("unsupported dtype %d", (int)(dtype));  \
    return 1;                                             \
  }
