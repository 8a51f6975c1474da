// Per-row linears with tiny feature counts (K, N <= 16): the un-embedding's per-pixel Linear 12 -> 3
// (train/layers.py:52), the U-Net's 1x1x1 output conv 16 -> 3 (train/unet.py:144-153) and their backward passes.
// With 8.4 M voxels per step and a dozen channels these are HBM-bound streams, not GEMMs: one thread per row, the
// weights in shared memory, fp32 accumulation in registers; the weight gradient keeps its K x N accumulators in
// registers across a grid-stride loop and reduces them with warp shuffles + one atomic per element per block.
#include <algorithm>

#include "common.cuh"

namespace vvae {

__device__ __forceinline__ long long cdiv_dev(long long a, long long b) { return (a + b - 1) / b; }

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

// Packed rows of an odd handful of elements (the 3-channel RGB maps: 6 bytes per row) cannot be read or written with
// vector accesses per thread: a block's 256 consecutive rows are staged through shared memory with 16-byte accesses.
__device__ __forceinline__ void stage_in(bf16* sdst, const bf16* g, int n) {       // g 16-byte aligned
  const int nv = n >> 3;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) reinterpret_cast<uint4*>(sdst)[i] = reinterpret_cast<const uint4*>(g)[i];
  for (int i = (nv << 3) + threadIdx.x; i < n; i += blockDim.x) sdst[i] = g[i];
}
__device__ __forceinline__ void stage_out(bf16* g, const bf16* ssrc, int n) {      // g 16-byte aligned
  const int nv = n >> 3;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(ssrc)[i];
  for (int i = (nv << 3) + threadIdx.x; i < n; i += blockDim.x) g[i] = ssrc[i];
}

// xvec / yvec / avec: 0 = element-wise, 1 = 8-byte vectors per thread, 2 = packed rows staged through shared memory
template <int KP, int NP>
__global__ void __launch_bounds__(256)
small_linear_fwd_kernel(const SmallLinArgs a, int xvec, int yvec, int avec) {
  __shared__ float sw[KP * NP], sb[NP];
  __shared__ __align__(16) bf16 s_x[256 * KP], s_y[256 * NP], s_a[256 * NP];
  for (int i = threadIdx.x; i < KP * NP; i += blockDim.x) {
    const int k = i / NP, n = i % NP;
    sw[i] = (k < a.K && n < a.N) ? __bfloat162float(a.w[k * a.wk + n * a.wn]) : 0.f;
  }
  for (int i = threadIdx.x; i < NP; i += blockDim.x) sb[i] = (a.bias && i < a.N) ? a.bias[i] : 0.f;
  __syncthreads();
  for (long long m0 = (long long)blockIdx.x * 256; m0 < a.M; m0 += (long long)gridDim.x * 256) {   // block-uniform
    const long long m = m0 + threadIdx.x;
    const bool live = m < a.M;
    const int nrows = (int)(a.M - m0 < 256 ? a.M - m0 : 256);
    if (xvec == 2) stage_in(s_x, a.x + m0 * a.K, nrows * a.K);
    if (avec == 2) stage_in(s_a, a.aux + m0 * a.N, nrows * a.N);
    if (xvec == 2 || avec == 2) __syncthreads();
    float xv[KP], acc[NP];
    if (xvec == 2) {
#pragma unroll
      for (int k = 0; k < KP; ++k) xv[k] = (k < a.K && live) ? __bfloat162float(s_x[threadIdx.x * a.K + k]) : 0.f;
    } else if (live) {
      load_row<KP>(a.x + m * a.x_ld, a.K, xvec != 0, xv);
    }
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[n] = sb[n];
    if (live) {
#pragma unroll
      for (int k = 0; k < KP; ++k)
#pragma unroll
        for (int n = 0; n < NP; ++n) acc[n] = fmaf(xv[k], sw[k * NP + n], acc[n]);
      if (a.aux) {
        float av[NP];
        if (avec == 2) {
#pragma unroll
          for (int n = 0; n < NP; ++n) av[n] = n < a.N ? __bfloat162float(s_a[threadIdx.x * a.N + n]) : 0.f;
        } else {
          load_row<NP>(a.aux + m * a.aux_ld, a.N, avec != 0, av);
        }
#pragma unroll
        for (int n = 0; n < NP; ++n) acc[n] += av[n];
      }
    }
    if (yvec == 2) {
      if (live) {
#pragma unroll
        for (int n = 0; n < NP; ++n)
          if (n < a.N) s_y[threadIdx.x * a.N + n] = __float2bfloat16_rn(acc[n]);
      }
      __syncthreads();
      stage_out(a.y + m0 * a.N, s_y, nrows * a.N);
    } else if (live) {
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
    __syncthreads();                                  // the staging tiles are reused by the next slab
  }
}

template <int KP, int NP>
__global__ void __launch_bounds__(256)
small_linear_wgrad_kernel(const bf16* __restrict__ x, long long x_ld, const bf16* __restrict__ dy, long long dy_ld,
                          float* __restrict__ dw, long long dk, long long dn, long long M, int K, int N, int xvec,
                          int dvec) {
  __shared__ float red[KP * NP];
  __shared__ __align__(16) bf16 s_x[256 * KP], s_d[256 * NP];
  for (int i = threadIdx.x; i < KP * NP; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[KP][NP];
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[k][n] = 0.f;
  for (long long m0 = (long long)blockIdx.x * 256; m0 < M; m0 += (long long)gridDim.x * 256) {   // block-uniform
    const long long m = m0 + threadIdx.x;
    const bool live = m < M;
    const int nrows = (int)(M - m0 < 256 ? M - m0 : 256);
    if (xvec == 2) stage_in(s_x, x + m0 * K, nrows * K);
    if (dvec == 2) stage_in(s_d, dy + m0 * N, nrows * N);
    if (xvec == 2 || dvec == 2) __syncthreads();
    if (live) {
      float xv[KP], dv[NP];
      if (xvec == 2) {
#pragma unroll
        for (int k = 0; k < KP; ++k) xv[k] = k < K ? __bfloat162float(s_x[threadIdx.x * K + k]) : 0.f;
      } else {
        load_row<KP>(x + m * x_ld, K, xvec != 0, xv);
      }
      if (dvec == 2) {
#pragma unroll
        for (int n = 0; n < NP; ++n) dv[n] = n < N ? __bfloat162float(s_d[threadIdx.x * N + n]) : 0.f;
      } else {
        load_row<NP>(dy + m * dy_ld, N, dvec != 0, dv);
      }
#pragma unroll
      for (int k = 0; k < KP; ++k)
#pragma unroll
        for (int n = 0; n < NP; ++n) acc[k][n] = fmaf(xv[k], dv[n], acc[k][n]);
    }
    if (xvec == 2 || dvec == 2) __syncthreads();
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
// access mode of a [M, n] operand with row stride ld: 1 = 8-byte vectors per thread, 2 = packed rows staged through
// shared memory (16-byte aligned base; a slab of 256 rows is 512*n bytes, so every slab starts aligned), 0 = element-wise
static int access_mode(const void* p, int n, long long ld) {
  if (n % 4 == 0 && ld % 4 == 0 && al8(p)) return 1;
  if (ld == n && ((uintptr_t)p % 16) == 0) return 2;
  return 0;
}
static int pad4(int v) { return v <= 4 ? 4 : (v <= 12 ? 12 : 16); }

bool small_linear_ok(int K, int N) { return K >= 1 && N >= 1 && K <= 16 && N <= 16; }

int small_linear_fwd(const SmallLinArgs& a, cudaStream_t s) {
  const int xvec = access_mode(a.x, a.K, a.x_ld);
  const int yvec = access_mode(a.y, a.N, a.y_ld);
  const int avec = a.aux ? access_mode(a.aux, a.N, a.aux_ld) : 0;
  const int blocks = (int)std::min<long long>(cdiv(a.M, 256), (long long)num_sms() * 8);
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
  const int xvec = access_mode(x, K, x_ld);
  const int dvec = access_mode(dy, N, dy_ld);
  const int blocks = (int)std::min<long long>(cdiv(M, 256), (long long)num_sms() * 4);
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

// ---------------------------------------------------------------------------------------------------------------------
// Rank-1 shapes of the encoder's frame-selection head (train/model.py:56-58: Linear 96 -> 1 over every token):
//   fwd    y[m]     = x[m,:] . w + b                      (M x K) . (K x 1)
//   dgrad  dx[m,k]  = dy[m] * w[k] (+ aux[m,k])           (M x 1) . (1 x K)
//   wgrad  dw[k]   += sum_m x[m,k] * dy[m]                (K x M) . (M x 1)
// K = 96 (a multiple of 8, <= 256).  8 lanes per row, 16-byte loads; these are 6 MB streams, not GEMMs.
__global__ void __launch_bounds__(256)
rank1_fwd_kernel(const bf16* __restrict__ x, long long x_ld, const bf16* __restrict__ w, long long w_st,
                 const float* __restrict__ bias, bf16* __restrict__ y, long long y_ld, long long M, int K) {
  const int sub = threadIdx.x & 7, nch = K >> 3;
  const long long row0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long rstride = ((long long)gridDim.x * blockDim.x) >> 3;
  const long long iters = cdiv_dev(M, rstride);
  for (long long it = 0; it < iters; ++it) {           // whole warps iterate together (shuffles below)
    const long long m = row0 + it * rstride;
    float acc = 0.f;
    if (m < M) {
      for (int c = sub; c < nch; c += 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(x + m * x_ld + c * 8);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc = fmaf(__uint_as_float(u[t] << 16), __bfloat162float(w[(long long)(c * 8 + 2 * t) * w_st]), acc);
          acc = fmaf(__uint_as_float(u[t] & 0xffff0000u), __bfloat162float(w[(long long)(c * 8 + 2 * t + 1) * w_st]), acc);
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (m < M && sub == 0) y[m * y_ld] = __float2bfloat16_rn(acc + (bias ? bias[0] : 0.f));
  }
}

__global__ void __launch_bounds__(256)
rank1_dgrad_kernel(const bf16* __restrict__ dy, long long dy_ld, const bf16* __restrict__ w, long long w_st,
                   const bf16* __restrict__ aux, long long aux_ld, bf16* __restrict__ dx, long long dx_ld, long long M,
                   int K) {
  const int nch = K >> 3;
  const long long total = M * nch;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long m = i / nch;
    const int c = (int)(i - m * nch);
    const float d = __bfloat162float(dy[m * dy_ld]);
    float o[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = d * __bfloat162float(w[(long long)(c * 8 + t) * w_st]);
    if (aux) {
      const uint4 v = *reinterpret_cast<const uint4*>(aux + m * aux_ld + c * 8);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        o[2 * t] += __uint_as_float(u[t] << 16);
        o[2 * t + 1] += __uint_as_float(u[t] & 0xffff0000u);
      }
    }
    uint4 r;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(o[4], o[5]), p3 = __floats2bfloat162_rn(o[6], o[7]);
    r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
    r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(dx + m * dx_ld + c * 8) = r;
  }
}

// blockDim = 256 = 8 row slots x 32 column chunks (K <= 256); per-thread partials over a grid-stride loop of rows
__global__ void __launch_bounds__(256)
rank1_wgrad_kernel(const bf16* __restrict__ x, long long x_ld, const bf16* __restrict__ dy, long long dy_ld,
                   float* __restrict__ dw, long long dw_st, long long M, int K) {
  __shared__ float red[256];
  const int nch = K >> 3, c = threadIdx.x & 31, slot = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  if (c < nch) {
    for (long long m = (long long)blockIdx.x * 8 + slot; m < M; m += (long long)gridDim.x * 8) {
      const float d = __bfloat162float(dy[m * dy_ld]);
      const uint4 v = *reinterpret_cast<const uint4*>(x + m * x_ld + c * 8);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc[2 * t] = fmaf(__uint_as_float(u[t] << 16), d, acc[2 * t]);
        acc[2 * t + 1] = fmaf(__uint_as_float(u[t] & 0xffff0000u), d, acc[2 * t + 1]);
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) atomicAdd(&red[c * 8 + t], acc[t]);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) atomicAdd(dw + (long long)k * dw_st, red[k]);
}

bool rank1_ok(int K) { return K >= 8 && K <= 256 && K % 8 == 0; }

int rank1_fwd(const bf16* x, long long x_ld, const bf16* w, long long w_st, const float* bias, bf16* y, long long y_ld,
              long long M, int K, cudaStream_t s) {
  const int blocks = (int)std::min<long long>(cdiv(M * 8, 256), (long long)num_sms() * 8);
  rank1_fwd_kernel<<<blocks, 256, 0, s>>>(x, x_ld, w, w_st, bias, y, y_ld, M, K);
  return check_launch("rank1_fwd");
}
int rank1_dgrad(const bf16* dy, long long dy_ld, const bf16* w, long long w_st, const bf16* aux, long long aux_ld, bf16* dx,
                long long dx_ld, long long M, int K, cudaStream_t s) {
  const int blocks = (int)std::min<long long>(cdiv(M * (K / 8), 256), (long long)num_sms() * 8);
  rank1_dgrad_kernel<<<blocks, 256, 0, s>>>(dy, dy_ld, w, w_st, aux, aux_ld, dx, dx_ld, M, K);
  return check_launch("rank1_dgrad");
}
int rank1_wgrad(const bf16* x, long long x_ld, const bf16* dy, long long dy_ld, float* dw, long long dw_st, long long M, int K,
                cudaStream_t s) {
  const int blocks = (int)std::min<long long>(cdiv(M, 8 * 16), (long long)num_sms() * 4);
  rank1_wgrad_kernel<<<blocks, 256, 0, s>>>(x, x_ld, dy, dy_ld, dw, dw_st, M, K);
  return check_launch("rank1_wgrad");
}

}  // namespace vvae
