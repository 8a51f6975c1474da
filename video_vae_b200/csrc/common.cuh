// Shared device/host helpers for libvvae (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vvae.h"

namespace vvae {

using bf16 = __nv_bfloat16;

// ---- error plumbing (thread-local message, never throws across the ABI) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> status

#define VVAE_REQUIRE(cond, ...)              \
  do {                                       \
    if (!(cond)) {                           \
      ::vvae::set_error(__VA_ARGS__);        \
      return VVAE_ERR_INVALID;               \
    }                                        \
  } while (0)

#define VVAE_DISPATCH_DTYPE(dtype, T, ...)                 \
  do {                                                     \
    if ((dtype) == VVAE_F32) {                             \
      using T = float;                                     \
      __VA_ARGS__;                                         \
    } else if ((dtype) == VVAE_BF16) {                     \
      using T = ::vvae::bf16;                              \
      __VA_ARGS__;                                         \
    } else {                                               \
      ::vvae::set_error("unsupported dtype %d", (int)(dtype)); \
      return VVAE_ERR_INVALID;                             \
    }                                                      \
  } while (0)

inline cudaStream_t as_stream(vvae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int num_sms();  // cudaDevAttrMultiProcessorCount of the current device, queried once (api.cu); 148 on B200
inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ----
// A kernel launched through launch_pdl may begin while the previous kernel in the stream is still draining: its
// prologue (barrier init, TMEM allocation, descriptor prefetch, smem staging of constants that no kernel writes) overlaps
// the predecessor's tail, and pdl_wait() -- which every such kernel executes BEFORE its first access to global memory
// another kernel may have produced -- blocks until the predecessor has completed and flushed.  pdl_launch_dependents()
// in a kernel whose CTAs are all resident lets the successor's CTAs take over SMs as they become free.  Inside a captured
// CUDA graph these launches become programmatic-dependency edges.  vvae_debug_set(11, 1) turns the attribute off.
extern long long g_dbg[32];
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dbg[11] ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- scalar conversion ----
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
// round-trip through T (models a rounding point of the compute dtype)
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f<T>(from_f<T>(v)); }

// ---- 16-byte vectors of T ----
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const {
    uint32_t w = (&raw.x)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
  __device__ __forceinline__ void set(int i, float v) {
    uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    uint32_t& w = (&raw.x)[i >> 1];
    w = (i & 1) ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
  }
};

// ---- warp / block reductions ----
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
// sum over a block of up to 1024 threads; result valid in every thread. scratch: >= 33 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) scratch[32] = r;
  }
  __syncthreads();
  return scratch[32];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float dsiluf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}

}  // namespace vvae
