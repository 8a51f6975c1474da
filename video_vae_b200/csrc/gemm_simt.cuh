// Generic SIMT GEMM with functor operand loaders and epilogues (fp32 accumulate).
//
// This is the any-shape / fp32 path: it serves the fp32 correctness configuration (cfg 1), odd shapes
// (K=12, N=3, N=1 ...), the implicit-GEMM convolutions in fp32 and ConvTranspose, and is the on-device
// cross-check for the tcgen05 kernels.  The bf16 production shapes go through gemm_sm100.cu.
#pragma once
#include "common.cuh"

namespace vvae {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

// ---- operand loaders: float operator()(row, col) ----
template <typename T>
struct RowMajorLoader {  // X[r*ld + c]
  const T* p; long long ld;
  __device__ __forceinline__ float operator()(long long r, long long c) const { return to_f(p[r * ld + c]); }
};
template <typename T>
struct ColMajorLoader {  // X[c*ld + r]
  const T* p; long long ld;
  __device__ __forceinline__ float operator()(long long r, long long c) const { return to_f(p[c * ld + r]); }
};

// ---- epilogue for plain row-major C ----
template <typename TO, typename TA>
struct EpiStore {
  TO* C; long long ldc;
  const float* bias;
  int mode;
  const TA* aux_in; long long ld_ai;
  TA* aux_out; long long ld_ao;
  int atomic;  // accumulate with atomicAdd (TO == float)
  __device__ __forceinline__ void operator()(long long m, int n, float acc, bool first_split) const {
    float v = acc;
    if (bias && first_split) v += bias[n];
    if (mode == VVAE_EPI_SILU) {
      if (aux_out) aux_out[m * ld_ao + n] = from_f<TA>(v);
      v = siluf_(round_to<TA>(v));
    } else if (mode == VVAE_EPI_RESIDUAL) {
      v += to_f(aux_in[m * ld_ai + n]);
    } else if (mode == VVAE_EPI_DSILU) {
      v *= dsiluf_(to_f(aux_in[m * ld_ai + n]));
    }
    if (atomic) {
      if constexpr (sizeof(TO) == 4) atomicAdd(reinterpret_cast<float*>(C) + m * ldc + n, v);
    } else {
      C[m * ldc + n] = from_f<TO>(v);
    }
  }
};

template <class AL, class BL, class EP>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(AL al, BL bl, EP ep, long long M, int N, long long K, long long k_per_split) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const long long kbeg = (long long)blockIdx.z * k_per_split;
  const long long kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = kbeg; k0 < kend; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm = (tid >> 4) + 16 * i, kk = tid & 15;
      long long m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < kend) ? al(m, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk = (tid >> 6) + 4 * i, nn = tid & 63;
      long long k = k0 + kk;
      int n = n0 + nn;
      Bs[kk][nn] = (k < kend && n < N) ? bl(k, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool first = (blockIdx.z == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) ep(m, n, acc[i][j], first);
    }
  }
}

template <class AL, class BL, class EP>
inline int launch_gemm_simt(AL al, BL bl, EP ep, long long M, int N, long long K, int splits, cudaStream_t s) {
  if (M <= 0 || N <= 0) return VVAE_OK;
  if (splits < 1) splits = 1;
  long long kps = cdiv(K, splits);
  kps = cdiv(kps, SG_BK) * SG_BK;
  if (kps < SG_BK) kps = SG_BK;
  splits = (int)cdiv(K > 0 ? K : 1, kps);
  dim3 grid((unsigned)cdiv(M, SG_BM), (unsigned)cdiv(N, SG_BN), (unsigned)splits);
  gemm_simt_kernel<<<grid, 256, 0, s>>>(al, bl, ep, M, N, K, kps);
  return check_launch("gemm_simt");
}

}  // namespace vvae
