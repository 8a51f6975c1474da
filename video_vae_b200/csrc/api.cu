// Error plumbing, device probe and the small elementwise plumbing kernels (cast, fill, column sums).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace vvae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  // queried once per process (C++11 magic static: thread-safe); one process drives one GPU
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        v <= 0) {
      cudaGetLastError();
      v = 148;   // B200; only reached without a device, where every launch fails loudly anyway
    }
    return v;
  }();
  return n;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VVAE_ERR_CUDA;
  }
  return VVAE_OK;
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long long n) {
  long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    if (i + 3 < n) {
      float v0 = to_f(src[i]), v1 = to_f(src[i + 1]), v2 = to_f(src[i + 2]), v3 = to_f(src[i + 3]);
      dst[i] = from_f<D>(v0);
      dst[i + 1] = from_f<D>(v1);
      dst[i + 2] = from_f<D>(v2);
      dst[i + 3] = from_f<D>(v3);
    } else {
      for (long long j = i; j < n; ++j) dst[j] = from_f<D>(to_f(src[j]));
    }
  }
}

// fp32 -> bf16 with 16-byte accesses on both sides (the per-step parameter shadow copy).
__global__ void cast_f32_bf16_vec_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, long long n8) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n8; i += stride) {
    float4 a = src[2 * i], b = src[2 * i + 1];
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0);
    o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2);
    o.w = *reinterpret_cast<uint32_t*>(&p3);
    dst[i] = o;
  }
}

// bf16 -> fp32 with 16-byte accesses (the reduced bf16 gradient back into the fp32 flat buffer).
__global__ void cast_bf16_f32_vec_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, long long n8) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n8; i += stride) {
    const uint4 v = src[i];
    dst[2 * i] = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                             __uint_as_float(v.y & 0xffff0000u));
    dst[2 * i + 1] = make_float4(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u), __uint_as_float(v.w << 16),
                                 __uint_as_float(v.w & 0xffff0000u));
  }
}

__global__ void fill_kernel(float* dst, float v, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}

// out[n] += sum_rows x[row, n].  Block = 32 x 8: each warp owns 32 consecutive columns (coalesced),
// the 8 warps of a block stride over a slab of rows; partials meet in smem, one atomic per column per block.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long ld, long long rows, int n, float* __restrict__ out,
                              int rows_per_block) {
  __shared__ float part[8][33];
  int col = blockIdx.x * 32 + threadIdx.x;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  if (col < n)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f(x[r * ld + col]);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < n) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][threadIdx.x];
    atomicAdd(out + col, s);
  }
}

// bf16, n % 8 == 0: each thread owns 8 consecutive columns (one 16-byte load per row), a warp 256 columns, the 8 warps
// of a block stride over a slab of rows with 4 independent loads in flight per thread.
__global__ void __launch_bounds__(256)
colsum_bf16_vec_kernel(const bf16* __restrict__ x, long long ld, long long rows, int n, float* __restrict__ out,
                       int rows_per_block) {
  __shared__ float part[8][32][9];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + lane) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  if (col < n) {
    long long r = r0 + w;
    for (; r + 56 < r1; r += 64) {              // 8 independent 16-byte loads in flight per thread
      Vec16<bf16> v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u].load(x + (r + 8 * u) * ld + col);
#pragma unroll
      for (int t = 0; t < 8; ++t)
        acc[t] += ((v[0].get(t) + v[1].get(t)) + (v[2].get(t) + v[3].get(t))) +
                  ((v[4].get(t) + v[5].get(t)) + (v[6].get(t) + v[7].get(t)));
    }
    for (; r < r1; r += 8) {
      Vec16<bf16> v0;
      v0.load(x + r * ld + col);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] += v0.get(t);
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) part[w][lane][t] = acc[t];
  __syncthreads();
  // 256 threads finish the 256 column sums of the block: thread (lane l, t = w) adds the 8 warps' partials
  if (col < n) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += part[j][lane][w];
    atomicAdd(out + col + w, sum);
  }
}


// Narrow bf16 matrices (bias gradients of the U-Net convolutions and of the per-pixel Linears: 3..64 columns over
// millions of rows).  The wide kernel above would use 2 lanes of 32 for 16 columns.
// (a) n % 4 == 0: thread (row lane, 4-column chunk); consecutive threads read consecutive chunks, then the next row.
__global__ void __launch_bounds__(256)
colsum_bf16_chunk4_kernel(const bf16* __restrict__ x, long long ld, long long rows, int n, float* __restrict__ out,
                          long long rows_per_block) {
  __shared__ float red[128];
  const int cpr = n >> 2, rpi = 256 / cpr;           // chunks per row, rows per block iteration
  const int c = threadIdx.x % cpr, rl = threadIdx.x / cpr;
  for (int i = threadIdx.x; i < n; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (rl < rpi) {
    const bf16* p = x + 4 * c;
    long long r = r0 + rl;
    for (; r + 3LL * rpi < r1; r += 4LL * rpi) {     // 4 independent 8-byte loads in flight
      uint2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint2*>(p + (r + (long long)u * rpi) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0 += __uint_as_float(v[u].x << 16); a1 += __uint_as_float(v[u].x & 0xffff0000u);
        a2 += __uint_as_float(v[u].y << 16); a3 += __uint_as_float(v[u].y & 0xffff0000u);
      }
    }
    for (; r < r1; r += rpi) {
      const uint2 v = *reinterpret_cast<const uint2*>(p + r * ld);
      a0 += __uint_as_float(v.x << 16); a1 += __uint_as_float(v.x & 0xffff0000u);
      a2 += __uint_as_float(v.y << 16); a3 += __uint_as_float(v.y & 0xffff0000u);
    }
    atomicAdd(&red[4 * c], a0); atomicAdd(&red[4 * c + 1], a1);
    atomicAdd(&red[4 * c + 2], a2); atomicAdd(&red[4 * c + 3], a3);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(out + i, red[i]);
}
// (b) packed rows (ld == n) of N <= 16 columns: the matrix is a flat array; a thread takes lcm(N, 8) consecutive elements
// (whole rows, 16-byte loads), so the column of each of its elements is a compile-time constant.
template <int N>
__global__ void __launch_bounds__(256)
colsum_bf16_packed_kernel(const bf16* __restrict__ x, long long rows, float* __restrict__ out) {
  constexpr int G = (N % 8 == 0) ? 8 : (N % 4 == 0) ? 4 : (N % 2 == 0) ? 2 : 1;   // gcd(N, 8)
  constexpr int E = N * 8 / G;                       // elements per thread step = lcm(N, 8)
  constexpr int RP = E / N;                          // rows per thread step
  __shared__ float red[N];
  if (threadIdx.x < N) red[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[N];
#pragma unroll
  for (int j = 0; j < N; ++j) acc[j] = 0.f;
  const long long steps = rows / RP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < steps; i += (long long)gridDim.x * blockDim.x) {
    const uint4* p = reinterpret_cast<const uint4*>(x + i * E);
    uint4 v[E / 8];
#pragma unroll
    for (int u = 0; u < E / 8; ++u) v[u] = p[u];
#pragma unroll
    for (int u = 0; u < E / 8; ++u) {
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int t = 0; t < 8; ++t)
        acc[(u * 8 + t) % N] += (t & 1) ? __uint_as_float(w[t >> 1] & 0xffff0000u) : __uint_as_float(w[t >> 1] << 16);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)           // the rows % RP leftover rows
    for (long long r = steps * RP; r < rows; ++r)
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] += __bfloat162float(x[r * N + j]);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float sres = warp_sum(acc[j]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[j], sres);
  }
  __syncthreads();
  if (threadIdx.x < N) atomicAdd(out + threadIdx.x, red[threadIdx.x]);
}

}  // namespace vvae

using namespace vvae;

extern "C" {

const char* vvae_last_error(void) { return g_err; }
int vvae_version(void) { return 100; }

int vvae_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? 1 : 0;
}

int vvae_cast(const void* src, int sd, void* dst, int dd, long long n, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  cudaStream_t s = as_stream(stream);
  int threads = 256;
  if (sd == VVAE_F32 && dd == VVAE_BF16 && n % 8 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) {
    long long n8 = n / 8;
    int blocks = (int)std::min<long long>(cdiv(n8, threads), num_sms() * 16);
    cast_f32_bf16_vec_kernel<<<blocks, threads, 0, s>>>((const float4*)src, (uint4*)dst, n8);
    return check_launch("cast");
  }
  if (sd == VVAE_BF16 && dd == VVAE_F32 && n % 8 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) {
    long long n8 = n / 8;
    int blocks = (int)std::min<long long>(cdiv(n8, threads), num_sms() * 16);
    cast_bf16_f32_vec_kernel<<<blocks, threads, 0, s>>>((const uint4*)src, (float4*)dst, n8);
    return check_launch("cast");
  }
  int blocks = (int)std::min<long long>(cdiv(n, threads * 4), num_sms() * 16);
  if (sd == VVAE_F32 && dd == VVAE_BF16) cast_kernel<<<blocks, threads, 0, s>>>((const float*)src, (bf16*)dst, n);
  else if (sd == VVAE_BF16 && dd == VVAE_F32) cast_kernel<<<blocks, threads, 0, s>>>((const bf16*)src, (float*)dst, n);
  else if (sd == VVAE_F32 && dd == VVAE_F32) cast_kernel<<<blocks, threads, 0, s>>>((const float*)src, (float*)dst, n);
  else if (sd == VVAE_BF16 && dd == VVAE_BF16) cast_kernel<<<blocks, threads, 0, s>>>((const bf16*)src, (bf16*)dst, n);
  else VVAE_REQUIRE(false, "vvae_cast: bad dtypes %d -> %d", sd, dd);
  return check_launch("cast");
}

int vvae_fill_f32(float* dst, float value, long long n, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  int blocks = (int)std::min<long long>(cdiv(n, 256), num_sms() * 16);
  fill_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dst, value, n);
  return check_launch("fill");
}

int vvae_colsum(const void* x, long long ld, long long rows, int n, float* out, int dtype, vvae_stream_t stream) {
  if (rows <= 0 || n <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && out, "vvae_colsum: null pointer");
  if (dtype == VVAE_BF16 && ld == n && n <= 16 && rows >= 4096 && ((uintptr_t)x % 16 == 0)) {
    const int blocks = (int)std::min<long long>(cdiv(rows, 256 * 8), (long long)num_sms() * 8);
    cudaStream_t s = as_stream(stream);
#define CS_PACKED(NN) case NN: colsum_bf16_packed_kernel<NN><<<blocks, 256, 0, s>>>((const bf16*)x, rows, out); break;
    switch (n) {
      CS_PACKED(1) CS_PACKED(2) CS_PACKED(3) CS_PACKED(4) CS_PACKED(5) CS_PACKED(6) CS_PACKED(7) CS_PACKED(8)
      CS_PACKED(9) CS_PACKED(10) CS_PACKED(11) CS_PACKED(12) CS_PACKED(13) CS_PACKED(14) CS_PACKED(15) CS_PACKED(16)
    }
#undef CS_PACKED
    return check_launch("colsum");
  }
  if (dtype == VVAE_BF16 && n % 4 == 0 && n <= 128 && ld % 4 == 0 && rows >= 4096 && ((uintptr_t)x % 8 == 0)) {
    const int rpi = 256 / (n / 4);
    const long long want = (long long)num_sms() * 8;
    const long long rpb = std::max<long long>(8LL * rpi, (cdiv(rows, want) + 4 * rpi - 1) / (4 * rpi) * (4 * rpi));
    colsum_bf16_chunk4_kernel<<<(unsigned)cdiv(rows, rpb), 256, 0, as_stream(stream)>>>((const bf16*)x, ld, rows, n, out, rpb);
    return check_launch("colsum");
  }
  if (dtype == VVAE_BF16 && n % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x % 16 == 0)) {
    const int cb = (int)cdiv(n, 256);
    const long long want = cdiv(num_sms() * 3, cb);
    const long long rpb = std::max<long long>(128, (cdiv(rows, want) + 63) / 64 * 64);   // whole 64-row unrolled steps
    dim3 grid(cb, (unsigned)cdiv(rows, rpb));
    colsum_bf16_vec_kernel<<<grid, 256, 0, as_stream(stream)>>>((const bf16*)x, ld, rows, n, out, (int)rpb);
    return check_launch("colsum");
  }
  int col_blocks = (int)cdiv(n, 32);
  // enough row slabs to fill the machine a few times over, at least 64 rows each
  long long want = cdiv(num_sms() * 8, col_blocks);
  long long rpb = std::max<long long>(64, cdiv(rows, want));
  int row_blocks = (int)cdiv(rows, rpb);
  dim3 grid(col_blocks, row_blocks), block(32, 8);
  VVAE_DISPATCH_DTYPE(dtype, T,
                      (colsum_kernel<T><<<grid, block, 0, as_stream(stream)>>>((const T*)x, ld, rows, n, out, (int)rpb)));
  return check_launch("colsum");
}

}  // extern "C"
