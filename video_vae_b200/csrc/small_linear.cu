// Per-row linears with tiny feature counts (K, N <= 16): the un-embedding's per-pixel Linear 12 -> 3
// (train/layers.py:52), the U-Net's 1x1x1 output conv 16 -> 3 (train/unet.py:144-153) and their backward passes.
// With 8.4 M voxels per step and a dozen channels these are HBM-bound streams, not GEMMs: one thread per row, the
// weights in shared memory, fp32 accumulation in registers; the weight gradient keeps its K x N accumulators in
// registers across a grid-stride loop and reduces them with warp shuffles + one atomic per element per block.
#include <algorithm>

#include "common.cuh"

namespace vvae {

struct SmallLinArgs {
  const bf16* x; long long x_ld;
  const bf16* w; long long wk, wn;        // W(k, n) = w[k*wk + n*wn]
  const float* bias;
  bf16* y; long long y_ld;
  const bf16* aux; long long aux_ld;      // optional residual added to the output
  long long M; int K, N;
};

template <int P>
__device__ __forceinline__ void load_row(const bf16* p, int n, bool vec, float (&v)[P]) {
  if (vec) {                                // n % 4 == 0, 8-byte aligned
#pragma unroll
    for (int i = 0; i < P; i += 4) {
      if (i < n) {
        const uint2 u = *reinterpret_cast<const uint2*>(p + i);
        v[i] = __uint_as_float(u.x << 16); v[i + 1] = __uint_as_float(u.x & 0xffff0000u);
        v[i + 2] = __uint_as_float(u.y << 16); v[i + 3] = __uint_as_float(u.y & 0xffff0000u);
      } else {
        v[i] = v[i + 1] = v[i + 2] = v[i + 3] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < P; ++i) v[i] = i < n ? __bfloat162float(p[i]) : 0.f;
  }
}

template <int KP, int NP>
__global__ void __launch_bounds__(256)
small_linear_fwd_kernel(const SmallLinArgs a, int xvec, int yvec, int avec) {
  __shared__ float sw[KP * NP], sb[NP];
  for (int i = threadIdx.x; i < KP * NP; i += blockDim.x) {
    const int k = i / NP, n = i % NP;
    sw[i] = (k < a.K && n < a.N) ? __bfloat162float(a.w[k * a.wk + n * a.wn]) : 0.f;
  }
  for (int i = threadIdx.x; i < NP; i += blockDim.x) sb[i] = (a.bias && i < a.N) ? a.bias[i] : 0.f;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += stride) {
    float xv[KP], acc[NP];
    load_row<KP>(a.x + m * a.x_ld, a.K, xvec != 0, xv);
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[n] = sb[n];
#pragma unroll
    for (int k = 0; k < KP; ++k)
#pragma unroll
      for (int n = 0; n < NP; ++n) acc[n] = fmaf(xv[k], sw[k * NP + n], acc[n]);
    if (a.aux) {
      float av[NP];
      load_row<NP>(a.aux + m * a.aux_ld, a.N, avec != 0, av);
#pragma unroll
      for (int n = 0; n < NP; ++n) acc[n] += av[n];
    }
    bf16* yr = a.y + m * a.y_ld;
    if (yvec) {
#pragma unroll
      for (int n = 0; n < NP; n += 4) {
        if (n < a.N) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[n], acc[n + 1]), p1 = __floats2bfloat162_rn(acc[n + 2], acc[n + 3]);
          *reinterpret_cast<uint2*>(yr + n) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
        }
      }
    } else {
#pragma unroll
      for (int n = 0; n < NP; ++n)
        if (n < a.N) yr[n] = __float2bfloat16_rn(acc[n]);
    }
  }
}

template <int KP, int NP>
__global__ void __launch_bounds__(256)
small_linear_wgrad_kernel(const bf16* __restrict__ x, long long x_ld, const bf16* __restrict__ dy, long long dy_ld,
                          float* __restrict__ dw, long long dk, long long dn, long long M, int K, int N, int xvec,
                          int dvec) {
  __shared__ float red[KP * NP];
  for (int i = threadIdx.x; i < KP * NP; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[KP][NP];
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[k][n] = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
    float xv[KP], dv[NP];
    load_row<KP>(x + m * x_ld, K, xvec != 0, xv);
    load_row<NP>(dy + m * dy_ld, N, dvec != 0, dv);
#pragma unroll
    for (int k = 0; k < KP; ++k)
#pragma unroll
      for (int n = 0; n < NP; ++n) acc[k][n] = fmaf(xv[k], dv[n], acc[k][n]);
  }
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      const float s = warp_sum(acc[k][n]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&red[k * NP + n], s);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < KP * NP; i += blockDim.x) {
    const int k = i / NP, n = i % NP;
    if (k < K && n < N) atomicAdd(dw + k * dk + n * dn, red[i]);
  }
}

static bool al8(const void* p) { return ((uintptr_t)p % 8) == 0; }
static int pad4(int v) { return v <= 4 ? 4 : (v <= 12 ? 12 : 16); }

bool small_linear_ok(int K, int N) { return K >= 1 && N >= 1 && K <= 16 && N <= 16; }

int small_linear_fwd(const SmallLinArgs& a, cudaStream_t s) {
  const int xvec = (a.K % 4 == 0 && a.x_ld % 4 == 0 && al8(a.x)) ? 1 : 0;
  const int yvec = (a.N % 4 == 0 && a.y_ld % 4 == 0 && al8(a.y)) ? 1 : 0;
  const int avec = (a.aux && a.N % 4 == 0 && a.aux_ld % 4 == 0 && al8(a.aux)) ? 1 : 0;
  const int blocks = (int)std::min<long long>(cdiv(a.M, 256), 148LL * 8);
  const int kp = pad4(a.K), np = pad4(a.N);
#define SL_FWD(KP, NP) small_linear_fwd_kernel<KP, NP><<<blocks, 256, 0, s>>>(a, xvec, yvec, avec)
  if (kp == 4 && np == 4) SL_FWD(4, 4);
  else if (kp == 4 && np == 12) SL_FWD(4, 12);
  else if (kp == 4 && np == 16) SL_FWD(4, 16);
  else if (kp == 12 && np == 4) SL_FWD(12, 4);
  else if (kp == 16 && np == 4) SL_FWD(16, 4);
  else if (kp == 12 && np == 12) SL_FWD(12, 12);
  else if (kp == 16 && np == 16) SL_FWD(16, 16);
  else if (kp == 12 && np == 16) SL_FWD(12, 16);
  else SL_FWD(16, 12);
#undef SL_FWD
  return check_launch("small_linear_fwd");
}

bool small_linear_wgrad_ok(int K, int N) { return small_linear_ok(K, N) && pad4(K) * pad4(N) <= 64; }

int small_linear_wgrad(const bf16* x, long long x_ld, const bf16* dy, long long dy_ld, float* dw, long long dk, long long dn,
                       long long M, int K, int N, cudaStream_t s) {
  const int xvec = (K % 4 == 0 && x_ld % 4 == 0 && al8(x)) ? 1 : 0;
  const int dvec = (N % 4 == 0 && dy_ld % 4 == 0 && al8(dy)) ? 1 : 0;
  const int blocks = (int)std::min<long long>(cdiv(M, 256), 148LL * 4);
  const int kp = pad4(K), np = pad4(N);
#define SL_WG(KP, NP) small_linear_wgrad_kernel<KP, NP><<<blocks, 256, 0, s>>>(x, x_ld, dy, dy_ld, dw, dk, dn, M, K, N, xvec, dvec)
  if (kp == 4 && np == 4) SL_WG(4, 4);
  else if (kp == 4 && np == 12) SL_WG(4, 12);
  else if (kp == 4 && np == 16) SL_WG(4, 16);
  else if (kp == 12 && np == 4) SL_WG(12, 4);
  else SL_WG(16, 4);
#undef SL_WG
  return check_launch("small_linear_wgrad");
}

}  // namespace vvae
