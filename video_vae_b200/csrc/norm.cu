// HBM-bound normalisation kernels: LayerNorm fwd/bwd, QK-LayerNorm + RoPE fwd/bwd, GroupNorm + SiLU fwd/bwd.
// Statistics follow Flax: fp32, "fast variance" max(0, E[x^2] - E[x]^2), eps inside the rsqrt.
#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "sm100.cuh"

namespace vvae {

// =====================================================================================================
// LayerNorm: one warp per row, the row lives in registers (NCHUNK 16-byte vectors per lane): 1 read, 1 write.
// =====================================================================================================
template <typename T, int NCHUNK>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long rows, int D, float eps) {
  constexpr int V = Vec16<T>::N;
  constexpr int UR = 2;                     // rows in flight per warp
  pdl_launch_dependents();                  // (grid = resident blocks)
  pdl_wait();
  extern __shared__ float ln_sm[];          // gamma[D] | beta[D]: read back as 16-byte broadcasts-free vectors
  float* sg = ln_sm;
  float* sb = ln_sm + D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    sg[i] = gamma ? gamma[i] : 1.f;
    sb[i] = beta ? beta[i] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = D / V;
  for (long long row0 = warp0; row0 < rows; row0 += UR * nwarps) {
    Vec16<T> v[UR][NCHUNK];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long row = row0 + u * nwarps;
      if (row < rows) {
#pragma unroll
        for (int i = 0; i < NCHUNK; ++i) {
          const int j = lane + 32 * i;
          if (j < nvec) v[u][i].load(x + row * D + j * V);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long row = row0 + u * nwarps;
      if (row >= rows) break;
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NCHUNK; ++i) {
        const int j = lane + 32 * i;
        if (j < nvec) {
#pragma unroll
          for (int t = 0; t < V; ++t) {
            const float f = v[u][i].get(t);
            s += f;
            s2 = fmaf(f, f, s2);
          }
        }
      }
      s = warp_sum(s);
      s2 = warp_sum(s2);
      const float mu = s / D;
      const float var = fmaxf(s2 / D - mu * mu, 0.f);
      const float r = rsqrtf(var + eps);
      if (lane == 0) {
        if (mean_out) mean_out[row] = mu;
        if (rstd_out) rstd_out[row] = r;
      }
      T* yr = y + row * D;
#pragma unroll
      for (int i = 0; i < NCHUNK; ++i) {
        const int j = lane + 32 * i;
        if (j < nvec) {
          Vec16<T> o;
#pragma unroll
          for (int t = 0; t < V; t += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(sg + j * V + t);
            const float4 b4 = *reinterpret_cast<const float4*>(sb + j * V + t);
            o.set(t, fmaf((v[u][i].get(t) - mu) * r, g4.x, b4.x));
            o.set(t + 1, fmaf((v[u][i].get(t + 1) - mu) * r, g4.y, b4.y));
            o.set(t + 2, fmaf((v[u][i].get(t + 2) - mu) * r, g4.z, b4.z));
            o.set(t + 3, fmaf((v[u][i].get(t + 3) - mu) * r, g4.w, b4.w));
          }
          o.store(yr + j * V);
        }
      }
    }
  }
}

// dx = r*(g - mean(g) - xhat*mean(g*xhat)) (+ dres), g = dy*gamma.  dgamma/dbeta: per-lane register partials over the
// rows a warp visits, combined across the block in smem, then one atomicAdd per column per block.
template <typename T, int NCHUNK>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, const T* __restrict__ dres,
                     T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int D) {
  constexpr int V = Vec16<T>::N;
  pdl_launch_dependents();        // (grid = resident blocks)
  pdl_wait();
  extern __shared__ float red[];  // [2][D]
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = D / V;
  float pg[NCHUNK][V], pb[NCHUNK][V];
  float* sgam = red + 2 * D;                 // gamma staged in shared memory (keeps 8*NCHUNK registers free)
  for (int i = threadIdx.x; i < D; i += blockDim.x) sgam[i] = gamma ? gamma[i] : 1.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i)
#pragma unroll
    for (int t = 0; t < V; ++t) pg[i][t] = pb[i][t] = 0.f;

  for (long long row = warp0; row < rows; row += nwarps) {
    const float mu = mean[row], r = rstd[row];
    Vec16<T> vx[NCHUNK], vd[NCHUNK], vr[NCHUNK];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {          // every load of the row is issued before the first use
      int j = lane + 32 * i;
      if (j < nvec) {
        vx[i].load(x + row * D + j * V);
        vd[i].load(dy + row * D + j * V);
        if (dres) vr[i].load(dres + row * D + j * V);
      }
    }
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      int j = lane + 32 * i;
      if (j < nvec) {
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float xh = (vx[i].get(t) - mu) * r;
          float d = vd[i].get(t);
          float g = d * sgam[j * V + t];
          sg += g;
          sgx += g * xh;
          pg[i][t] += d * xh;
          pb[i][t] += d;
        }
      }
    }
    sg = warp_sum(sg) / D;
    sgx = warp_sum(sgx) / D;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      int j = lane + 32 * i;
      if (j < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float xh = (vx[i].get(t) - mu) * r;
          float d = vd[i].get(t);
          float g = d * sgam[j * V + t];
          float f = r * (g - sg - xh * sgx);
          if (dres) f += vr[i].get(t);
          o.set(t, f);
        }
        o.store(dx + row * D + j * V);
      }
    }
  }
  if (dgamma || dbeta) {
    for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) red[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      int j = lane + 32 * i;
      if (j < nvec) {
#pragma unroll
        for (int t = 0; t < V; ++t) {
          atomicAdd(&red[j * V + t], pg[i][t]);
          atomicAdd(&red[D + j * V + t], pb[i][t]);
        }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      if (dgamma) atomicAdd(dgamma + c, red[c]);
      if (dbeta) atomicAdd(dbeta + c, red[D + c]);
    }
  }
}


// =====================================================================================================
// LayerNorm streaming path (bf16, D = NV * 256): rows are staged in shared memory by 1-D bulk async copies
// (cp.async.bulk + mbarrier), each warp owns a private ring of STAGES row slots and keeps STAGES rows of every input
// in flight while it reduces and writes the current one.  The register-staged kernels above hold 9 x 16 bytes per lane
// per row in registers, so occupancy (23-46 % of the warp slots, profiles/r02w_norm_ncu.json) caps the bytes in flight and
// they run at 3.7-4.0 TB/s; here the loads in flight cost no registers.
// =====================================================================================================
__device__ __forceinline__ float lns_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float lns_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t lns_pack(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ uint4 lns_lds(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <int NV, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1)
layernorm_bwd_stream_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean,
                            const float* __restrict__ rstd, const float* __restrict__ gamma, const bf16* __restrict__ dres,
                            bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows) {
  constexpr int D = NV * 256;
  constexpr uint32_t ROWB = D * 2;
  pdl_launch_dependents();
  extern __shared__ uint8_t lns_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lns_raw) + 127) & ~uintptr_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = dres ? 3 : 2;                                   // tensors per row slot: x, dy (, dres)
  uint8_t* ring = base + (size_t)warp * STAGES * 3 * ROWB;        // [STAGES][3][ROWB]
  const uint32_t ring_u32 = sm100::smem_u32(ring);
  float* red = reinterpret_cast<float*>(base + (size_t)WARPS * STAGES * 3 * ROWB);   // [2][D]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * D) + warp * STAGES;
  if (lane == 0) {
    for (int i = 0; i < STAGES; ++i) sm100::mbar_init(&bars[i], 1);
    sm100::fence_barrier_init();
  }
  pdl_wait();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) red[c] = 0.f;
  float gm[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int t = 0; t < 8; ++t) gm[i][t] = gamma ? gamma[(i * 32 + lane) * 8 + t] : 1.f;
  __syncthreads();

  const long long warp0 = (long long)blockIdx.x * WARPS + warp;
  const long long nwarps = (long long)gridDim.x * WARPS;
  auto issue = [&](long long row, int slot) {                    // lane 0 only
    sm100::mbar_expect_tx(&bars[slot], (uint32_t)NT * ROWB);
    uint8_t* dst = ring + (size_t)slot * 3 * ROWB;
    sm100::bulk_load_1d(dst, x + row * D, ROWB, &bars[slot]);
    sm100::bulk_load_1d(dst + ROWB, dy + row * D, ROWB, &bars[slot]);
    if (dres) sm100::bulk_load_1d(dst + 2 * ROWB, dres + row * D, ROWB, &bars[slot]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      const long long row = warp0 + i * nwarps;
      if (row < rows) issue(row, i);
    }
  }
  float pg[NV][8], pb[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int t = 0; t < 8; ++t) pg[i][t] = pb[i][t] = 0.f;

  int slot = 0;
  uint32_t phase = 0;
  float mu_n = 0.f, r_n = 0.f;
  if (warp0 < rows) { mu_n = mean[warp0]; r_n = rstd[warp0]; }
  for (long long row = warp0; row < rows; row += nwarps) {
    const float r = r_n, nmr = -mu_n * r_n;
    if (row + nwarps < rows) { mu_n = mean[row + nwarps]; r_n = rstd[row + nwarps]; }
    sm100::mbar_wait(&bars[slot], phase);
    const uint32_t sx = ring_u32 + (uint32_t)slot * 3u * ROWB + (uint32_t)lane * 16u;
    float sg = 0.f, sgx = 0.f;
    uint4 vx[NV], vd[NV], vr[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      vx[i] = lns_lds(sx + i * 512);
      vd[i] = lns_lds(sx + ROWB + i * 512);
      vr[i] = dres ? lns_lds(sx + 2 * ROWB + i * 512) : make_uint4(0, 0, 0, 0);
    }
    __syncwarp();                                                // every lane has read this slot: refill it
    {
      const long long nxt = row + (long long)STAGES * nwarps;
      if (lane == 0 && nxt < rows) issue(nxt, slot);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t wx[4] = {vx[i].x, vx[i].y, vx[i].z, vx[i].w}, wd[4] = {vd[i].x, vd[i].y, vd[i].z, vd[i].w};
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float xv = (t & 1) ? lns_hi(wx[t >> 1]) : lns_lo(wx[t >> 1]);
        const float d = (t & 1) ? lns_hi(wd[t >> 1]) : lns_lo(wd[t >> 1]);
        const float xh = fmaf(xv, r, nmr);
        const float g = d * gm[i][t];
        sg += g;
        sgx = fmaf(g, xh, sgx);
        pg[i][t] = fmaf(d, xh, pg[i][t]);
        pb[i][t] += d;
      }
    }
    sg = warp_sum(sg) * (1.f / D);
    sgx = warp_sum(sgx) * (1.f / D);
    // dx = (r*gamma)*dy - r*sg - (r*sgx)*xhat (+ dres), xhat = r*x - r*mu  =>  dx = (r*gamma)*dy + cx*x + cc
    const float cx = -r * sgx * r, cc = -r * sg - r * sgx * nmr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t wx[4] = {vx[i].x, vx[i].y, vx[i].z, vx[i].w}, wd[4] = {vd[i].x, vd[i].y, vd[i].z, vd[i].w};
      const uint32_t wr[4] = {vr[i].x, vr[i].y, vr[i].z, vr[i].w};
      uint32_t o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float f0 = fmaf(cx, lns_lo(wx[t]), fmaf(r * gm[i][2 * t], lns_lo(wd[t]), cc)) + lns_lo(wr[t]);
        const float f1 = fmaf(cx, lns_hi(wx[t]), fmaf(r * gm[i][2 * t + 1], lns_hi(wd[t]), cc)) + lns_hi(wr[t]);
        o[t] = lns_pack(f0, f1);
      }
      reinterpret_cast<uint4*>(dx + row * D)[i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (++slot == STAGES) { slot = 0; phase ^= 1; }
  }
  if (dgamma || dbeta) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int c = (i * 32 + lane) * 8 + t;
        atomicAdd(&red[c], pg[i][t]);
        atomicAdd(&red[D + c], pb[i][t]);
      }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      if (dgamma) atomicAdd(dgamma + c, red[c]);
      if (dbeta) atomicAdd(dbeta + c, red[D + c]);
    }
  }
}


template <int NV, int WARPS, int STAGES>
static int launch_ln_bwd_stream(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                                const void* dres, void* dx, float* dgamma, float* dbeta, long long rows, cudaStream_t s) {
  auto kern = layernorm_bwd_stream_kernel<NV, WARPS, STAGES>;
  const size_t sm = 128 + (size_t)WARPS * STAGES * 3 * (NV * 512) + 2 * (size_t)(NV * 256) * sizeof(float) + WARPS * STAGES * 8;
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
      set_error("layernorm_bwd: cudaFuncSetAttribute(%d)", (int)sm);
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  const int blocks = (int)std::min<long long>(num_sms(), cdiv(rows, WARPS));
  launch_pdl(kern, dim3(blocks), dim3(WARPS * 32), sm, s, (const bf16*)dy, (const bf16*)x, mean, rstd, gamma,
             (const bf16*)dres, (bf16*)dx, dgamma, dbeta, rows);
  return VVAE_OK;
}


template <int NV, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1)
layernorm_fwd_stream_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                            long long rows, float eps) {
  constexpr int D = NV * 256;
  constexpr uint32_t ROWB = D * 2;
  pdl_launch_dependents();
  extern __shared__ uint8_t lns_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lns_raw) + 127) & ~uintptr_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = base + (size_t)warp * STAGES * ROWB;
  const uint32_t ring_u32 = sm100::smem_u32(ring);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)WARPS * STAGES * ROWB) + warp * STAGES;
  if (lane == 0) {
    for (int i = 0; i < STAGES; ++i) sm100::mbar_init(&bars[i], 1);
    sm100::fence_barrier_init();
  }
  __syncwarp();
  float gm[NV][8], bt[NV][8];                                     // parameters: not produced by the previous kernel
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      gm[i][t] = gamma ? gamma[(i * 32 + lane) * 8 + t] : 1.f;
      bt[i][t] = beta ? beta[(i * 32 + lane) * 8 + t] : 0.f;
    }
  pdl_wait();
  const long long warp0 = (long long)blockIdx.x * WARPS + warp;
  const long long nwarps = (long long)gridDim.x * WARPS;
  auto issue = [&](long long row, int slot) {                    // lane 0 only
    sm100::mbar_expect_tx(&bars[slot], ROWB);
    sm100::bulk_load_1d(ring + (size_t)slot * ROWB, x + row * D, ROWB, &bars[slot]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      const long long row = warp0 + i * nwarps;
      if (row < rows) issue(row, i);
    }
  }
  int slot = 0;
  uint32_t phase = 0;
  for (long long row = warp0; row < rows; row += nwarps) {
    sm100::mbar_wait(&bars[slot], phase);
    const uint32_t sx = ring_u32 + (uint32_t)slot * ROWB + (uint32_t)lane * 16u;
    uint4 vx[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) vx[i] = lns_lds(sx + i * 512);
    __syncwarp();
    {
      const long long nxt = row + (long long)STAGES * nwarps;
      if (lane == 0 && nxt < rows) issue(nxt, slot);
    }
    float f[NV][8];
    float sm = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t wx[4] = {vx[i].x, vx[i].y, vx[i].z, vx[i].w};
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        f[i][t] = (t & 1) ? lns_hi(wx[t >> 1]) : lns_lo(wx[t >> 1]);
        sm += f[i][t];
        s2 = fmaf(f[i][t], f[i][t], s2);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                           // the two reductions interleaved
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mu = sm / D;
    const float var = fmaxf(s2 / D - mu * mu, 0.f);
    const float r = rsqrtf(var + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = r;
    }
    const float nmr = -mu * r;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint32_t o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float a = fmaf(fmaf(f[i][2 * t], r, nmr), gm[i][2 * t], bt[i][2 * t]);
        const float b = fmaf(fmaf(f[i][2 * t + 1], r, nmr), gm[i][2 * t + 1], bt[i][2 * t + 1]);
        o[t] = lns_pack(a, b);
      }
      reinterpret_cast<uint4*>(y + row * D)[i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (++slot == STAGES) { slot = 0; phase ^= 1; }
  }
}

template <int NV, int WARPS, int STAGES>
static int launch_ln_fwd_stream(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                                long long rows, float eps, cudaStream_t s) {
  auto kern = layernorm_fwd_stream_kernel<NV, WARPS, STAGES>;
  const size_t sm = 128 + (size_t)WARPS * STAGES * (NV * 512) + WARPS * STAGES * 8;
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
      set_error("layernorm_fwd: cudaFuncSetAttribute(%d)", (int)sm);
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  const int blocks = (int)std::min<long long>(num_sms(), cdiv(rows, WARPS));
  launch_pdl(kern, dim3(blocks), dim3(WARPS * 32), sm, s, (const bf16*)x, (bf16*)y, gamma, beta, mean, rstd, rows, eps);
  return VVAE_OK;
}

// =====================================================================================================
// QK-LayerNorm (scale only) + RoPE on the q|k part of a fused [rows, 3*H*hd] projection. One warp per head vector.
// =====================================================================================================
constexpr int QK_MAXP = 4;  // pairs per lane -> hd <= 256

template <typename T>
__global__ void __launch_bounds__(256)
qknorm_rope_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, const float* __restrict__ q_scale,
                       const float* __restrict__ k_scale, const T* __restrict__ cos_tab,
                       const T* __restrict__ sin_tab, long long rows, int H, int hd, long long pos_div, int pos_mod,
                       float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long nvecs = rows * 2 * H;
  const int half = hd >> 1;
  const long long in_ld = 3LL * H * hd, out_ld = 2LL * H * hd;
  for (long long vid = warp0; vid < nvecs; vid += nwarps) {
    const int head = (int)(vid % H);
    const int which = (int)((vid / H) & 1);
    const long long row = vid / (2 * H);
    const int pos = (int)((row / pos_div) % pos_mod);
    const T* src = qkv + row * in_ld + (long long)which * H * hd + (long long)head * hd;
    T* dst = out + row * out_ld + (long long)which * H * hd + (long long)head * hd;
    const float* sc = which ? k_scale : q_scale;
    float a[QK_MAXP], b[QK_MAXP];
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int p = 0; p < QK_MAXP; ++p) {
      int i = lane + 32 * p;
      if (i < half) {
        a[p] = to_f(src[i]);
        b[p] = to_f(src[i + half]);
        s += a[p] + b[p];
        s2 += a[p] * a[p] + b[p] * b[p];
      }
    }
    s = warp_sum(s);
    s2 = warp_sum(s2);
    const float mu = s / hd;
    const float r = rsqrtf(fmaxf(s2 / hd - mu * mu, 0.f) + eps);
#pragma unroll
    for (int p = 0; p < QK_MAXP; ++p) {
      int i = lane + 32 * p;
      if (i < half) {
        float x1 = round_to<T>((a[p] - mu) * r * sc[i]);
        float x2 = round_to<T>((b[p] - mu) * r * sc[i + half]);
        float c1 = to_f(cos_tab[(long long)pos * hd + i]), c2 = to_f(cos_tab[(long long)pos * hd + i + half]);
        float s1 = to_f(sin_tab[(long long)pos * hd + i]), sn2 = to_f(sin_tab[(long long)pos * hd + i + half]);
        // y = x*cos + rotate_half(x)*sin, rotate_half(x) = [-x2, x1]
        float y1 = round_to<T>(x1 * c1) + round_to<T>(-x2 * s1);
        float y2 = round_to<T>(x2 * c2) + round_to<T>(x1 * sn2);
        dst[i] = from_f<T>(y1);
        dst[i + half] = from_f<T>(y2);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
qknorm_rope_bwd_kernel(T* __restrict__ dqkv, const T* __restrict__ qkv, const float* __restrict__ q_scale,
                       const float* __restrict__ k_scale, const T* __restrict__ cos_tab,
                       const T* __restrict__ sin_tab, float* __restrict__ dq_scale, float* __restrict__ dk_scale,
                       long long rows, int H, int hd, long long pos_div, int pos_mod, float eps) {
  __shared__ float red[2][2 * 32 * QK_MAXP];  // [which][hd]
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long nvecs = rows * 2 * H;
  const int half = hd >> 1;
  const long long ld = 3LL * H * hd;
  float ps[2][QK_MAXP][2];
#pragma unroll
  for (int w = 0; w < 2; ++w)
#pragma unroll
    for (int p = 0; p < QK_MAXP; ++p) ps[w][p][0] = ps[w][p][1] = 0.f;

  for (long long vid = warp0; vid < nvecs; vid += nwarps) {
    const int head = (int)(vid % H);
    const int which = (int)((vid / H) & 1);
    const long long row = vid / (2 * H);
    const int pos = (int)((row / pos_div) % pos_mod);
    const long long off = row * ld + (long long)which * H * hd + (long long)head * hd;
    const T* src = qkv + off;
    T* g = dqkv + off;
    const float* sc = which ? k_scale : q_scale;
    float a[QK_MAXP], b[QK_MAXP], da[QK_MAXP], db[QK_MAXP];
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int p = 0; p < QK_MAXP; ++p) {
      int i = lane + 32 * p;
      if (i < half) {
        a[p] = to_f(src[i]);
        b[p] = to_f(src[i + half]);
        s += a[p] + b[p];
        s2 += a[p] * a[p] + b[p] * b[p];
        float dy1 = to_f(g[i]), dy2 = to_f(g[i + half]);
        float c1 = to_f(cos_tab[(long long)pos * hd + i]), c2 = to_f(cos_tab[(long long)pos * hd + i + half]);
        float s1 = to_f(sin_tab[(long long)pos * hd + i]), sn2 = to_f(sin_tab[(long long)pos * hd + i + half]);
        // y1 = x1*c1 - x2*s1 ; y2 = x2*c2 + x1*sn2
        da[p] = dy1 * c1 + dy2 * sn2;
        db[p] = dy2 * c2 - dy1 * s1;
      }
    }
    s = warp_sum(s);
    s2 = warp_sum(s2);
    const float mu = s / hd;
    const float r = rsqrtf(fmaxf(s2 / hd - mu * mu, 0.f) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int p = 0; p < QK_MAXP; ++p) {
      int i = lane + 32 * p;
      if (i < half) {
        float xh1 = (a[p] - mu) * r, xh2 = (b[p] - mu) * r;
        float g1 = da[p] * sc[i], g2 = db[p] * sc[i + half];
        sg += g1 + g2;
        sgx += g1 * xh1 + g2 * xh2;
        if (which) { ps[1][p][0] += da[p] * xh1; ps[1][p][1] += db[p] * xh2; }
        else       { ps[0][p][0] += da[p] * xh1; ps[0][p][1] += db[p] * xh2; }
      }
    }
    sg = warp_sum(sg) / hd;
    sgx = warp_sum(sgx) / hd;
#pragma unroll
    for (int p = 0; p < QK_MAXP; ++p) {
      int i = lane + 32 * p;
      if (i < half) {
        float xh1 = (a[p] - mu) * r, xh2 = (b[p] - mu) * r;
        float g1 = da[p] * sc[i], g2 = db[p] * sc[i + half];
        g[i] = from_f<T>(r * (g1 - sg - xh1 * sgx));
        g[i + half] = from_f<T>(r * (g2 - sg - xh2 * sgx));
      }
    }
  }
  for (int c = threadIdx.x; c < 2 * 2 * 32 * QK_MAXP; c += blockDim.x) (&red[0][0])[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 2; ++w)
#pragma unroll
    for (int p = 0; p < QK_MAXP; ++p) {
      int i = lane + 32 * p;
      if (i < half) {
        atomicAdd(&red[w][i], ps[w][p][0]);
        atomicAdd(&red[w][i + half], ps[w][p][1]);
      }
    }
  __syncthreads();
  for (int c = threadIdx.x; c < hd; c += blockDim.x) {
    if (dq_scale) atomicAdd(dq_scale + c, red[0][c]);
    if (dk_scale) atomicAdd(dk_scale + c, red[1][c]);
  }
}

// -----------------------------------------------------------------------------------------------------
// Fast path (bf16, head_dim 64): each thread owns 8 consecutive elements (one 16-byte access) of one head vector, 8
// lanes form a head vector (LayerNorm statistics: 3 shuffle steps), the rotate_half partner is lane ^ 4 (4 packed
// shuffles).  A thread keeps the same (q|k, head, slice) for every row it visits, so its scale slice lives in registers
// and its d(scale) partials need one smem atomic per element at the very end.
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf_pack(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

__global__ void __launch_bounds__(256)
qknorm_rope_fwd_hd64_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const float* __restrict__ q_scale,
                            const float* __restrict__ k_scale, const bf16* __restrict__ cos_tab,
                            const bf16* __restrict__ sin_tab, long long rows, int H, long long pos_div, int pos_mod,
                            float eps) {
  pdl_launch_dependents();                      // (grid = resident blocks)
  pdl_wait();
  const int cpr = 16 * H;                       // 16-byte chunks of q|k per row
  const int c = threadIdx.x % cpr, rpi = blockDim.x / cpr;
  const int part = c & 7;
  const bool second = part >= 4;
  const float* scp = (c >= 8 * H ? k_scale : q_scale) + part * 8;
  float sc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) sc[t] = __ldg(scp + t);
  const long long in_ld = 192LL * H, out_ld = 128LL * H;
  constexpr int UR = 2;                          // rows in flight per thread (independent 16-byte loads)
  const long long rstride = (long long)gridDim.x * rpi;
  for (long long row0 = (long long)blockIdx.x * rpi + threadIdx.x / cpr; row0 < rows; row0 += UR * rstride) {
    uint4 xv[UR], cv[UR], sv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long row = min(row0 + u * rstride, rows - 1);   // tail rows are recomputed, never stored twice
      const int pos = (int)(((unsigned)row / (unsigned)pos_div) % (unsigned)pos_mod);   // 32-bit: rows < 2^31 (host check)
      xv[u] = *reinterpret_cast<const uint4*>(qkv + row * in_ld + c * 8);
      cv[u] = __ldg(reinterpret_cast<const uint4*>(cos_tab + (long long)pos * 64 + part * 8));
      sv[u] = __ldg(reinterpret_cast<const uint4*>(sin_tab + (long long)pos * 64 + part * 8));
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {               // straight-line: no divergence around the shuffles
      const long long row = row0 + u * rstride;
      float f[8];
      const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) { f[2 * t] = bf_lo(xw[t]); f[2 * t + 1] = bf_hi(xw[t]); }
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) { s += f[t]; s2 = fmaf(f[t], f[t], s2); }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const float mu = s * (1.f / 64.f);
      const float r = rsqrtf(fmaxf(s2 * (1.f / 64.f) - mu * mu, 0.f) + eps);
      uint32_t xn[4], xp[4];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        xn[t] = bf_pack((f[2 * t] - mu) * (r * sc[2 * t]), (f[2 * t + 1] - mu) * (r * sc[2 * t + 1]));
#pragma unroll
      for (int t = 0; t < 4; ++t) xp[t] = __shfl_xor_sync(0xffffffffu, xn[t], 4);
      const uint32_t cw[4] = {cv[u].x, cv[u].y, cv[u].z, cv[u].w}, sw_[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
      uint32_t yo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        // y = x*cos + rotate_half(x)*sin, rotate_half(x) = [-x2, x1].  The reference does this in bf16 (each product and
        // the sum rounded to bf16): exactly what the packed bf16 multiply / add instructions compute.
        const uint32_t rot = second ? xp[t] : (xp[t] ^ 0x80008000u);
        const __nv_bfloat162 a2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&xn[t]),
                                          *reinterpret_cast<const __nv_bfloat162*>(&cw[t]));
        const __nv_bfloat162 b2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&rot),
                                          *reinterpret_cast<const __nv_bfloat162*>(&sw_[t]));
        const __nv_bfloat162 y2 = __hadd2(a2, b2);
        yo[t] = *reinterpret_cast<const uint32_t*>(&y2);
      }
      if (row < rows) *reinterpret_cast<uint4*>(out + row * out_ld + c * 8) = make_uint4(yo[0], yo[1], yo[2], yo[3]);
    }
  }
}

__global__ void __launch_bounds__(256)
qknorm_rope_bwd_hd64_kernel(bf16* __restrict__ dqkv, const bf16* __restrict__ qkv, const float* __restrict__ q_scale,
                            const float* __restrict__ k_scale, const bf16* __restrict__ cos_tab,
                            const bf16* __restrict__ sin_tab, float* __restrict__ dq_scale, float* __restrict__ dk_scale,
                            float* __restrict__ dbias, long long rows, int H, long long pos_div, int pos_mod, float eps) {
  __shared__ float red[2][64];
  __shared__ float redb[2048];                   // column sums of the produced d(q|k) (the QKV bias gradient), [16*H*8]
  pdl_launch_dependents();                       // (grid = resident blocks)
  pdl_wait();
  const int cpr = 16 * H;
  const int c = threadIdx.x % cpr, rpi = blockDim.x / cpr;
  const int part = c & 7;
  const bool second = part >= 4;
  const int which = c >= 8 * H ? 1 : 0;
  const float* scp = (which ? k_scale : q_scale) + part * 8;
  float sc[8], ps[8], pb[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { sc[t] = __ldg(scp + t); ps[t] = 0.f; pb[t] = 0.f; }
  const long long ld = 192LL * H;
  constexpr int UR = 2;                          // rows in flight per thread
  const long long rstride = (long long)gridDim.x * rpi;
  for (long long row0 = (long long)blockIdx.x * rpi + threadIdx.x / cpr; row0 < rows; row0 += UR * rstride) {
    uint4 xv_[UR], gv_[UR], cv_[UR], sv_[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long row = min(row0 + u * rstride, rows - 1);
      const int pos = (int)(((unsigned)row / (unsigned)pos_div) % (unsigned)pos_mod);
      const long long off = row * ld + c * 8;
      xv_[u] = *reinterpret_cast<const uint4*>(qkv + off);
      gv_[u] = *reinterpret_cast<const uint4*>(dqkv + off);
      cv_[u] = __ldg(reinterpret_cast<const uint4*>(cos_tab + (long long)pos * 64 + part * 8));
      sv_[u] = __ldg(reinterpret_cast<const uint4*>(sin_tab + (long long)pos * 64 + part * 8));
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {               // straight-line: tail rows are computed but contribute / store nothing
      const long long row = row0 + u * rstride;
      const bool live = row < rows;
      const long long off = min(row, rows - 1) * ld + c * 8;
      const uint4 xv = xv_[u], gv = gv_[u], cv = cv_[u], sv = sv_[u];
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
      const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, sw_[4] = {sv.x, sv.y, sv.z, sv.w};
      float f[8], d[8];
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        f[2 * t] = bf_lo(xw[t]); f[2 * t + 1] = bf_hi(xw[t]);
        const uint32_t gp = __shfl_xor_sync(0xffffffffu, gw[t], 4);      // partner's upstream gradient
        const uint32_t sp = __shfl_xor_sync(0xffffffffu, sw_[t], 4);     // partner's sin slice
        // y1 = x1*c1 - x2*s1 ; y2 = x2*c2 + x1*s2  =>  dx1 = dy1*c1 + dy2*s2 ; dx2 = dy2*c2 - dy1*s1
        const float q0 = bf_lo(gp) * bf_lo(sp), q1 = bf_hi(gp) * bf_hi(sp);
        d[2 * t] = fmaf(bf_lo(gw[t]), bf_lo(cw[t]), second ? -q0 : q0);
        d[2 * t + 1] = fmaf(bf_hi(gw[t]), bf_hi(cw[t]), second ? -q1 : q1);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) { s += f[t]; s2 = fmaf(f[t], f[t], s2); }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const float mu = s * (1.f / 64.f);
      const float r = rsqrtf(fmaxf(s2 * (1.f / 64.f) - mu * mu, 0.f) + eps);
      float sg = 0.f, sgx = 0.f, xh[8], g[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        xh[t] = (f[t] - mu) * r;
        g[t] = d[t] * sc[t];
        sg += g[t];
        sgx = fmaf(g[t], xh[t], sgx);
        if (live) ps[t] = fmaf(d[t], xh[t], ps[t]);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
      }
      sg *= (1.f / 64.f);
      sgx *= (1.f / 64.f);
      uint32_t o4[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        o4[t] = bf_pack(r * (g[2 * t] - sg - xh[2 * t] * sgx), r * (g[2 * t + 1] - sg - xh[2 * t + 1] * sgx));
        if (dbias && live) {                     // sum of the ROUNDED outputs, as a column-sum pass sees them
          pb[2 * t] += bf_lo(o4[t]);
          pb[2 * t + 1] += bf_hi(o4[t]);
        }
      }
      if (live) *reinterpret_cast<uint4*>(dqkv + off) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    }
  }
  if (dbias) {
    for (int i = threadIdx.x; i < cpr * 8; i += blockDim.x) redb[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 8; ++t) atomicAdd(&redb[c * 8 + t], pb[t]);
    __syncthreads();
    for (int i = threadIdx.x; i < cpr * 8; i += blockDim.x) atomicAdd(dbias + i, redb[i]);
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 8; ++t) atomicAdd(&red[which][part * 8 + t], ps[t]);
  __syncthreads();
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    if (dq_scale) atomicAdd(dq_scale + i, red[0][i]);
    if (dk_scale) atomicAdd(dk_scale + i, red[1][i]);
  }
}


// Streaming variant for 8 heads x 64 (one warp per row: 2 x 2 KB of q|k and d(q|k) staged by bulk async copies into a
// per-warp ring, as the LayerNorm kernels above; the in-place result goes straight to global memory).  Lane l owns the
// 16-byte chunks l, l+32, l+64, l+96 of the row: the same 8 columns `part` of heads l/8, l/8+4 of q and of k.  The
// rotate-half partner chunk (8 lanes away) is read from the staged row instead of shuffled.
template <int WARPS, int STAGES, bool DBIAS>
__global__ void __launch_bounds__(WARPS * 32, 1)
qknorm_rope_bwd_stream_kernel(bf16* __restrict__ dqkv, const bf16* __restrict__ qkv, const float* __restrict__ q_scale,
                              const float* __restrict__ k_scale, const bf16* __restrict__ cos_tab,
                              const bf16* __restrict__ sin_tab, float* __restrict__ dq_scale, float* __restrict__ dk_scale,
                              float* __restrict__ dbias, long long rows, long long pos_div, int pos_mod, float eps) {
  constexpr int QK = 1024;                       // q|k columns of a row (8 heads x 64 x 2)
  constexpr long long LD = 1536;
  constexpr uint32_t HALF = QK * 2;              // bytes of one staged tensor row
  pdl_launch_dependents();
  extern __shared__ uint8_t lns_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lns_raw) + 127) & ~uintptr_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = base + (size_t)warp * STAGES * 2 * HALF;
  const uint32_t ring_u32 = sm100::smem_u32(ring);
  float* redb = reinterpret_cast<float*>(base + (size_t)WARPS * STAGES * 2 * HALF);   // [1024] bias-gradient partials
  float* red = redb + QK;                                                             // [2][64] scale-gradient partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 128) + warp * STAGES;
  if (lane == 0) {
    for (int i = 0; i < STAGES; ++i) sm100::mbar_init(&bars[i], 1);
    sm100::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < QK + 128; i += blockDim.x) redb[i] = 0.f;
  const int part = lane & 7;
  const bool second = part >= 4;
  float sc[2][8], ps[2][8], pb[DBIAS ? 4 : 1][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    sc[0][t] = __ldg(q_scale + part * 8 + t);
    sc[1][t] = __ldg(k_scale + part * 8 + t);
    ps[0][t] = ps[1][t] = 0.f;
#pragma unroll
    for (int i = 0; i < (DBIAS ? 4 : 1); ++i) pb[i][t] = 0.f;
  }
  __syncthreads();
  pdl_wait();
  const long long warp0 = (long long)blockIdx.x * WARPS + warp;
  const long long nwarps = (long long)gridDim.x * WARPS;
  auto issue = [&](long long row, int slot) {                    // lane 0 only
    sm100::mbar_expect_tx(&bars[slot], 2 * HALF);
    uint8_t* dst = ring + (size_t)slot * 2 * HALF;
    sm100::bulk_load_1d(dst, qkv + row * LD, HALF, &bars[slot]);
    sm100::bulk_load_1d(dst + HALF, dqkv + row * LD, HALF, &bars[slot]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      const long long row = warp0 + i * nwarps;
      if (row < rows) issue(row, i);
    }
  }
  int slot = 0;
  uint32_t phase = 0;
  for (long long row = warp0; row < rows; row += nwarps) {
    const int pos = (int)(((unsigned long long)row / (unsigned long long)pos_div) % (unsigned)pos_mod);
    const uint4 cv = __ldg(reinterpret_cast<const uint4*>(cos_tab + (long long)pos * 64 + part * 8));
    const uint4 sv = __ldg(reinterpret_cast<const uint4*>(sin_tab + (long long)pos * 64 + part * 8));
    const uint4 spv = __ldg(reinterpret_cast<const uint4*>(sin_tab + (long long)pos * 64 + (part ^ 4) * 8));
    const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, sw_[4] = {sv.x, sv.y, sv.z, sv.w};
    const uint32_t spw[4] = {spv.x, spv.y, spv.z, spv.w};
    (void)sw_;
    sm100::mbar_wait(&bars[slot], phase);
    const uint32_t sx = ring_u32 + (uint32_t)slot * 2u * HALF;
    uint4 xv[4], gv[4], gpv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xv[i] = lns_lds(sx + (uint32_t)(lane + 32 * i) * 16u);
      gv[i] = lns_lds(sx + HALF + (uint32_t)(lane + 32 * i) * 16u);
      gpv[i] = lns_lds(sx + HALF + (uint32_t)((lane ^ 4) + 32 * i) * 16u);     // rotate-half partner's upstream gradient
    }
    __syncwarp();                                                // every lane has read this slot: refill it
    {
      const long long nxt = row + (long long)STAGES * nwarps;
      if (lane == 0 && nxt < rows) issue(nxt, slot);
    }
    bf16* orow = dqkv + row * LD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int which = i >> 1;                                  // chunks 0..63 are q, 64..127 k
      const uint32_t xw[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w}, gw[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
      const uint32_t gpw[4] = {gpv[i].x, gpv[i].y, gpv[i].z, gpv[i].w};
      float f[8], d[8];
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        f[2 * t] = bf_lo(xw[t]); f[2 * t + 1] = bf_hi(xw[t]);
        // y1 = x1*c1 - x2*s1 ; y2 = x2*c2 + x1*s2  =>  dx1 = dy1*c1 + dy2*s2 ; dx2 = dy2*c2 - dy1*s1
        const float q0 = bf_lo(gpw[t]) * bf_lo(spw[t]), q1 = bf_hi(gpw[t]) * bf_hi(spw[t]);
        d[2 * t] = fmaf(bf_lo(gw[t]), bf_lo(cw[t]), second ? -q0 : q0);
        d[2 * t + 1] = fmaf(bf_hi(gw[t]), bf_hi(cw[t]), second ? -q1 : q1);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) { s += f[t]; s2 = fmaf(f[t], f[t], s2); }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const float mu = s * (1.f / 64.f);
      const float r = rsqrtf(fmaxf(s2 * (1.f / 64.f) - mu * mu, 0.f) + eps);
      float sg = 0.f, sgx = 0.f, xh[8], g[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        xh[t] = (f[t] - mu) * r;
        g[t] = d[t] * sc[which][t];
        sg += g[t];
        sgx = fmaf(g[t], xh[t], sgx);
        ps[which][t] = fmaf(d[t], xh[t], ps[which][t]);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
      }
      sg *= (1.f / 64.f);
      sgx *= (1.f / 64.f);
      uint32_t o4[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        o4[t] = bf_pack(r * (g[2 * t] - sg - xh[2 * t] * sgx), r * (g[2 * t + 1] - sg - xh[2 * t + 1] * sgx));
        if constexpr (DBIAS) {                                   // sum of the ROUNDED outputs, as a column-sum pass sees them
          pb[i][2 * t] += bf_lo(o4[t]);
          pb[i][2 * t + 1] += bf_hi(o4[t]);
        }
      }
      *reinterpret_cast<uint4*>(orow + (lane + 32 * i) * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    }
    if (++slot == STAGES) { slot = 0; phase ^= 1; }
  }
  if constexpr (DBIAS) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int t = 0; t < 8; ++t) atomicAdd(&redb[(lane + 32 * i) * 8 + t], pb[i][t]);
  }
#pragma unroll
  for (int w = 0; w < 2; ++w)
#pragma unroll
    for (int t = 0; t < 8; ++t) atomicAdd(&red[w * 64 + part * 8 + t], ps[w][t]);
  __syncthreads();
  if constexpr (DBIAS)
    for (int i = threadIdx.x; i < QK; i += blockDim.x) atomicAdd(dbias + i, redb[i]);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    if (dq_scale) atomicAdd(dq_scale + i, red[i]);
    if (dk_scale) atomicAdd(dk_scale + i, red[64 + i]);
  }
}


template <int WARPS, int STAGES, bool DBIAS>
static int launch_qk_bwd_stream(void* dqkv, const void* qkv, const float* q_scale, const float* k_scale, const void* cos_tab,
                                const void* sin_tab, float* dq_scale, float* dk_scale, float* dbias_qk, long long rows,
                                long long pos_div, int pos_mod, float eps, cudaStream_t s) {
  auto kern = qknorm_rope_bwd_stream_kernel<WARPS, STAGES, DBIAS>;
  const size_t sm = 128 + (size_t)WARPS * STAGES * 4096 + (1024 + 128) * sizeof(float) + WARPS * STAGES * 8;
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
      set_error("qknorm_rope_bwd: cudaFuncSetAttribute(%d)", (int)sm);
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  const int blocks = (int)std::min<long long>(num_sms(), cdiv(rows, WARPS));
  launch_pdl(kern, dim3(blocks), dim3(WARPS * 32), sm, s, (bf16*)dqkv, (const bf16*)qkv, q_scale, k_scale, (const bf16*)cos_tab,
             (const bf16*)sin_tab, dq_scale, dk_scale, dbias_qk, rows, pos_div, pos_mod, eps);
  return VVAE_OK;
}

// =====================================================================================================
// GroupNorm + SiLU on [B, S, C] (channels last).  blockDim.x is a multiple of C, so a thread always sees the same
// channel: its gamma/beta/mean/rstd live in registers and group partials need one smem atomic per thread.
// =====================================================================================================
template <typename T>
__global__ void groupnorm_stats_kernel(const T* __restrict__ x, float* __restrict__ stats, long long S, int C, int G,
                                       long long rows_per_block) {
  __shared__ float sm[2 * 64];
  const int b = blockIdx.y;
  const int c = threadIdx.x % C, cg = C / G, g = c / cg;
  const int rpi = blockDim.x / C;  // rows per iteration
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const T* xb = x + (long long)b * S * C;
  float s = 0.f, s2 = 0.f;
  for (long long r = r0 + threadIdx.x / C; r < r1; r += rpi) {
    float f = to_f(xb[r * C + c]);
    s += f;
    s2 += f * f;
  }
  atomicAdd(&sm[2 * g], s);
  atomicAdd(&sm[2 * g + 1], s2);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(stats + (long long)b * 2 * G + i, sm[i]);
}

__global__ void groupnorm_finalize_kernel(const float* __restrict__ stats, float* __restrict__ mean,
                                          float* __restrict__ rstd, int BG, float inv_n, float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < BG) {
    float mu = stats[2 * i] * inv_n;
    float var = fmaxf(stats[2 * i + 1] * inv_n - mu * mu, 0.f);
    mean[i] = mu;
    rstd[i] = rsqrtf(var + eps);
  }
}

template <typename T>
__global__ void groupnorm_silu_apply_kernel(const T* __restrict__ x, T* __restrict__ y, long long y_ld,
                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const float* __restrict__ mean, const float* __restrict__ rstd, long long S,
                                            int C, int G, long long rows_per_block) {
  const int b = blockIdx.y;
  const int c = threadIdx.x % C, g = c / (C / G);
  const int rpi = blockDim.x / C;
  const float mu = mean[b * G + g], r = rstd[b * G + g], ga = gamma[c], be = beta[c];
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const T* xb = x + (long long)b * S * C;
  T* yb = y + (long long)b * S * y_ld;
  for (long long rr = r0 + threadIdx.x / C; rr < r1; rr += rpi) {
    float z = round_to<T>((to_f(xb[rr * C + c]) - mu) * r * ga + be);
    yb[rr * y_ld + c] = from_f<T>(siluf_(z));
  }
}

// pass 1 of backward: per-(b,g) sums of g and g*xhat (into stats[b,g,0:2]) and dgamma/dbeta.
template <typename T>
__global__ void groupnorm_silu_bwd_stats_kernel(const T* __restrict__ dy, long long dy_ld, const T* __restrict__ x,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                float* __restrict__ stats, float* __restrict__ dgamma,
                                                float* __restrict__ dbeta, long long S, int C, int G,
                                                long long rows_per_block) {
  __shared__ float sm[2 * 64];
  __shared__ float smc[2 * 256];
  const int b = blockIdx.y;
  const int c = threadIdx.x % C, g = c / (C / G);
  const int rpi = blockDim.x / C;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sm[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) smc[i] = 0.f;
  __syncthreads();
  const float mu = mean[b * G + g], r = rstd[b * G + g], ga = gamma[c], be = beta[c];
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const T* xb = x + (long long)b * S * C;
  const T* db = dy + (long long)b * S * dy_ld;
  float s1 = 0.f, s2 = 0.f, dg = 0.f, dbt = 0.f;
  for (long long rr = r0 + threadIdx.x / C; rr < r1; rr += rpi) {
    float xh = (to_f(xb[rr * C + c]) - mu) * r;
    float z = round_to<T>(xh * ga + be);
    float dz = to_f(db[rr * dy_ld + c]) * dsiluf_(z);
    float gg = dz * ga;
    s1 += gg;
    s2 += gg * xh;
    dg += dz * xh;
    dbt += dz;
  }
  atomicAdd(&sm[2 * g], s1);
  atomicAdd(&sm[2 * g + 1], s2);
  atomicAdd(&smc[c], dg);
  atomicAdd(&smc[C + c], dbt);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(stats + (long long)b * 2 * G + i, sm[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, smc[i]);
    if (dbeta) atomicAdd(dbeta + i, smc[C + i]);
  }
}

template <typename T>
__global__ void groupnorm_silu_bwd_apply_kernel(const T* __restrict__ dy, long long dy_ld, const T* __restrict__ x,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                const float* __restrict__ stats, T* __restrict__ dx, long long S, int C,
                                                int G, float inv_n, long long rows_per_block) {
  const int b = blockIdx.y;
  const int c = threadIdx.x % C, g = c / (C / G);
  const int rpi = blockDim.x / C;
  const float mu = mean[b * G + g], r = rstd[b * G + g], ga = gamma[c], be = beta[c];
  const float m1 = stats[((long long)b * G + g) * 2] * inv_n, m2 = stats[((long long)b * G + g) * 2 + 1] * inv_n;
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const T* xb = x + (long long)b * S * C;
  const T* db = dy + (long long)b * S * dy_ld;
  T* dxb = dx + (long long)b * S * C;
  for (long long rr = r0 + threadIdx.x / C; rr < r1; rr += rpi) {
    float xh = (to_f(xb[rr * C + c]) - mu) * r;
    float z = round_to<T>(xh * ga + be);
    float gg = to_f(db[rr * dy_ld + c]) * dsiluf_(z) * ga;
    dxb[rr * C + c] = from_f<T>(r * (gg - m1 - xh * m2));
  }
}

// -----------------------------------------------------------------------------------------------------
// bf16 fast path: a thread owns 8 consecutive channels (one 16-byte access) of a voxel and keeps them for every voxel
// it visits (blockDim is a multiple of C/8), so the per-channel constants sit in registers; per-channel partial sums
// are folded into per-group / per-channel shared-memory accumulators once, at the end.
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gn_sigmoid(float x) {   // 0.5*tanh(x/2)+0.5 (one MUFU); |err| < 3e-4, bf16 outputs
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ void gn_unpack(const uint4& u, float (&f)[8]) {
  f[0] = bf_lo(u.x); f[1] = bf_hi(u.x); f[2] = bf_lo(u.y); f[3] = bf_hi(u.y);
  f[4] = bf_lo(u.z); f[5] = bf_hi(u.z); f[6] = bf_lo(u.w); f[7] = bf_hi(u.w);
}

// V (4 or 8) consecutive channels of one voxel: 8- or 16-byte access
template <int V> struct GnVec;
template <> struct GnVec<8> {
  uint4 u;
  __device__ __forceinline__ void load(const bf16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const { gn_unpack(u, f); }
  static __device__ __forceinline__ void store(bf16* p, const float (&o)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(bf_pack(o[0], o[1]), bf_pack(o[2], o[3]), bf_pack(o[4], o[5]), bf_pack(o[6], o[7]));
  }
};
template <> struct GnVec<4> {
  uint2 u;
  __device__ __forceinline__ void load(const bf16* p) { u = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void unpack(float (&f)[4]) const {
    f[0] = bf_lo(u.x); f[1] = bf_hi(u.x); f[2] = bf_lo(u.y); f[3] = bf_hi(u.y);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&o)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(bf_pack(o[0], o[1]), bf_pack(o[2], o[3]));
  }
};

template <int V>
__global__ void __launch_bounds__(256)
gn_stats_vec_kernel(const bf16* __restrict__ x, float* __restrict__ stats, long long S, int C, int G, long long rows_per_block) {
  __shared__ float sm[2 * 64];
  const int b = blockIdx.y, cpr = C / V, ch = threadIdx.x % cpr, rpi = blockDim.x / cpr, cg = C / G;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const bf16* xb = x + (long long)b * S * C + ch * V;
  float s[V], s2[V];
#pragma unroll
  for (int t = 0; t < V; ++t) s[t] = s2[t] = 0.f;
  constexpr int UR = V == 4 ? 4 : 2;               // independent loads in flight
  long long r = r0 + threadIdx.x / cpr;
  for (; r + (UR - 1) * rpi < r1; r += UR * rpi) {
    GnVec<V> u[UR];
#pragma unroll
    for (int k = 0; k < UR; ++k) u[k].load(xb + (r + k * rpi) * C);
#pragma unroll
    for (int k = 0; k < UR; ++k) {
      float f[V];
      u[k].unpack(f);
#pragma unroll
      for (int t = 0; t < V; ++t) { s[t] += f[t]; s2[t] = fmaf(f[t], f[t], s2[t]); }
    }
  }
  for (; r < r1; r += rpi) {
    GnVec<V> u;
    u.load(xb + r * C);
    float f[V];
    u.unpack(f);
#pragma unroll
    for (int t = 0; t < V; ++t) { s[t] += f[t]; s2[t] = fmaf(f[t], f[t], s2[t]); }
  }
  // lanes l, l+cpr, l+2cpr, ... of a warp hold the same channels: fold them before touching shared memory
  for (int o = cpr; o < 32; o <<= 1) {
#pragma unroll
    for (int t = 0; t < V; ++t) {
      s[t] += __shfl_xor_sync(0xffffffffu, s[t], o);
      s2[t] += __shfl_xor_sync(0xffffffffu, s2[t], o);
    }
  }
  if ((threadIdx.x & 31) < cpr) {
#pragma unroll
    for (int t = 0; t < V; ++t) {
      const int g = (ch * V + t) / cg;
      atomicAdd(&sm[2 * g], s[t]);
      atomicAdd(&sm[2 * g + 1], s2[t]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(stats + (long long)b * 2 * G + i, sm[i]);
}

template <int V>
__global__ void __launch_bounds__(256)
gn_apply_vec_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long y_ld, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                    long long S, int C, int G, long long rows_per_block) {
  const int b = blockIdx.y, cpr = C / V, ch = threadIdx.x % cpr, rpi = blockDim.x / cpr, cg = C / G;
  float sc[V], sh[V];     // y = silu(x * sc + sh)
#pragma unroll
  for (int t = 0; t < V; ++t) {
    const int c = ch * V + t, g = c / cg;
    const float r = rstd[b * G + g];
    sc[t] = r * gamma[c];
    sh[t] = fmaf(-mean[b * G + g], sc[t], beta[c]);
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const bf16* xb = x + (long long)b * S * C + ch * V;
  bf16* yb = y + (long long)b * S * y_ld + ch * V;
  constexpr int UR = V == 4 ? 8 : 4;             // voxels in flight per thread
  for (long long rb = r0 + threadIdx.x / cpr; rb < r1; rb += UR * rpi) {
    GnVec<V> xv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u)
      if (rb + u * rpi < r1) xv[u].load(xb + (rb + u * rpi) * C);
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long r = rb + u * rpi;
      if (r >= r1) break;
      float f[V], o[V];
      xv[u].unpack(f);
#pragma unroll
      for (int t = 0; t < V; ++t) {
        const float z = bf_round(fmaf(f[t], sc[t], sh[t]));
        o[t] = z * gn_sigmoid(z);
      }
      GnVec<V>::store(yb + r * y_ld, o);
    }
  }
}

// backward pass 1: dgamma_c = sum dz*xhat, dbeta_c = sum dz per channel; the per-(b,g) sums the second pass needs follow
// from them (sum of g = dz*gamma over a group is gamma_c*dbeta_c summed over its channels, sum of g*xhat likewise from
// dgamma_c), so only two accumulators per channel are carried.  V channels per thread: V = 4 for narrow maps (16 / 32
// channels: the register-heavier V = 8 ran at 24 % of the warp slots and 3.2 TB/s, profiles/r02w_norm_ncu.json).
template <int V>
__global__ void __launch_bounds__(256)
gn_bwd_stats_vec_kernel(const bf16* __restrict__ dy, long long dy_ld, const bf16* __restrict__ x,
                        const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                        const float* __restrict__ rstd, float* __restrict__ stats, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, long long S, int C, int G, long long rows_per_block) {
  __shared__ float sm[2 * 64];
  __shared__ float smc[2 * 256];
  const int b = blockIdx.y, cpr = C / V, ch = threadIdx.x % cpr, rpi = blockDim.x / cpr, cg = C / G;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sm[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) smc[i] = 0.f;
  __syncthreads();
  float rs[V], nmr[V], ga[V], be[V], dg[V], db[V];
#pragma unroll
  for (int t = 0; t < V; ++t) {
    const int c = ch * V + t, g = c / cg;
    rs[t] = rstd[b * G + g]; nmr[t] = -mean[b * G + g] * rs[t]; ga[t] = gamma[c]; be[t] = beta[c];
    dg[t] = db[t] = 0.f;
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const bf16* xb = x + (long long)b * S * C + ch * V;
  const bf16* dyb = dy + (long long)b * S * dy_ld + ch * V;
  constexpr int UR = V == 4 ? 4 : 2;
  for (long long rb = r0 + threadIdx.x / cpr; rb < r1; rb += UR * rpi) {
    GnVec<V> xv[UR], dv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u)
      if (rb + u * rpi < r1) {
        xv[u].load(xb + (rb + u * rpi) * C);
        dv[u].load(dyb + (rb + u * rpi) * dy_ld);
      }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      if (rb + u * rpi >= r1) break;
      float f[V], d[V];
      xv[u].unpack(f);
      dv[u].unpack(d);
#pragma unroll
      for (int t = 0; t < V; ++t) {
        const float xh = fmaf(f[t], rs[t], nmr[t]);
        const float z = bf_round(fmaf(xh, ga[t], be[t]));
        const float sg = gn_sigmoid(z);
        const float dz = d[t] * sg * fmaf(z, 1.f - sg, 1.f);
        dg[t] = fmaf(dz, xh, dg[t]);
        db[t] += dz;
      }
    }
  }
  for (int o = cpr; o < 32; o <<= 1) {
#pragma unroll
    for (int t = 0; t < V; ++t) {
      dg[t] += __shfl_xor_sync(0xffffffffu, dg[t], o);
      db[t] += __shfl_xor_sync(0xffffffffu, db[t], o);
    }
  }
  if ((threadIdx.x & 31) < cpr) {
#pragma unroll
    for (int t = 0; t < V; ++t) {
      const int c = ch * V + t, g = c / cg;
      atomicAdd(&sm[2 * g], ga[t] * db[t]);
      atomicAdd(&sm[2 * g + 1], ga[t] * dg[t]);
      atomicAdd(&smc[c], dg[t]);
      atomicAdd(&smc[C + c], db[t]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(stats + (long long)b * 2 * G + i, sm[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, smc[i]);
    if (dbeta) atomicAdd(dbeta + i, smc[C + i]);
  }
}

template <int V>
__global__ void __launch_bounds__(256)
gn_bwd_apply_vec_kernel(const bf16* __restrict__ dy, long long dy_ld, const bf16* __restrict__ x,
                        const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ stats, bf16* __restrict__ dx,
                        float* __restrict__ dxsum, long long S, int C, int G, float inv_n, long long rows_per_block) {
  __shared__ float smc[256];
  const int b = blockIdx.y, cpr = C / V, ch = threadIdx.x % cpr, rpi = blockDim.x / cpr, cg = C / G;
  // dx = rs*(dz*ga - m1 - xh*m2) = dz*k1 + xh*c2 + c1
  float rs[V], nmr[V], ga[V], be[V], k1[V], c1[V], c2[V], cs[V];
#pragma unroll
  for (int t = 0; t < V; ++t) {
    const int c = ch * V + t, g = c / cg;
    rs[t] = rstd[b * G + g]; nmr[t] = -mean[b * G + g] * rs[t]; ga[t] = gamma[c]; be[t] = beta[c];
    k1[t] = rs[t] * ga[t];
    c1[t] = -rs[t] * (stats[((long long)b * G + g) * 2] * inv_n);
    c2[t] = -rs[t] * (stats[((long long)b * G + g) * 2 + 1] * inv_n);
    cs[t] = 0.f;
  }
  if (dxsum) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) smc[i] = 0.f;
    __syncthreads();
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < S ? r0 + rows_per_block : S;
  const bf16* xb = x + (long long)b * S * C + ch * V;
  const bf16* dyb = dy + (long long)b * S * dy_ld + ch * V;
  bf16* dxb = dx + (long long)b * S * C + ch * V;
  constexpr int UR = V == 4 ? 4 : 2;
  for (long long rb = r0 + threadIdx.x / cpr; rb < r1; rb += UR * rpi) {
    GnVec<V> xv[UR], dv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u)
      if (rb + u * rpi < r1) {
        xv[u].load(xb + (rb + u * rpi) * C);
        dv[u].load(dyb + (rb + u * rpi) * dy_ld);
      }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const long long r = rb + u * rpi;
      if (r >= r1) break;
      float f[V], d[V], o[V];
      xv[u].unpack(f);
      dv[u].unpack(d);
#pragma unroll
      for (int t = 0; t < V; ++t) {
        const float xh = fmaf(f[t], rs[t], nmr[t]);
        const float z = bf_round(fmaf(xh, ga[t], be[t]));
        const float sg = gn_sigmoid(z);
        const float dz = d[t] * sg * fmaf(z, 1.f - sg, 1.f);
        o[t] = fmaf(dz, k1[t], fmaf(xh, c2[t], c1[t]));
        cs[t] += bf_round(o[t]);
      }
      GnVec<V>::store(dxb + r * C, o);
    }
  }
  // per-channel sums of the produced gradient: the bias gradient of the convolution in front of this norm
  if (dxsum) {
    for (int o = cpr; o < 32; o <<= 1) {
#pragma unroll
      for (int t = 0; t < V; ++t) cs[t] += __shfl_xor_sync(0xffffffffu, cs[t], o);
    }
    if ((threadIdx.x & 31) < cpr) {
#pragma unroll
      for (int t = 0; t < V; ++t) atomicAdd(&smc[ch * V + t], cs[t]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dxsum + i, smc[i]);
  }
}

static inline bool gn_vec_ok(int dtype, int C, const void* a, long long a_ld, const void* b_, long long b_ld) {
  return dtype == VVAE_BF16 && C % 8 == 0 && C <= 256 && 32 % (C / 8) == 0 && a_ld % 8 == 0 && b_ld % 8 == 0 &&
         ((uintptr_t)a % 16 == 0) && ((uintptr_t)b_ % 16 == 0);
}

static inline int gn_threads(int C) { return C * (256 / C > 0 ? 256 / C : 1); }

}  // namespace vvae

using namespace vvae;

// grid = exactly the blocks that are resident at once (persistent grid-stride kernels: no partial second wave)
template <typename K>
static int resident_grid(K kern, int threads, size_t smem, long long max_useful) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 2;
  }
  return (int)std::max<long long>(1, std::min<long long>(max_useful, (long long)num_sms() * occ));
}


template <typename T>
static int ln_fwd_dispatch(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                           long long rows, int D, float eps, cudaStream_t s) {
  constexpr int V = Vec16<T>::N;
  if constexpr (std::is_same<T, bf16>::value) {
    const int mode = (int)(g_dbg[7] >> 4) & 15;      // vvae_debug_set(7, 16): register-staged kernel; (7, 32): 16 warps x 4 stages
    if (D == 768 && mode != 1) {
      // 8 warps x 8 stages: 23.4 us at [32768, 768] against 24.6 (16 x 4) and 26.7 (register-staged); a 100 MB copy
      // under the same protocol takes 23.6 us (scripts/norm_ab.py)
      const int rc = mode == 2 ? launch_ln_fwd_stream<3, 16, 4>(x, y, gamma, beta, mean, rstd, rows, eps, s)
                                   : launch_ln_fwd_stream<3, 8, 8>(x, y, gamma, beta, mean, rstd, rows, eps, s);
      if (rc != VVAE_OK) return rc;
      return check_launch("layernorm_fwd");
    }
  }
  const int need = (int)cdiv(D / V, 32);
  const size_t smem = 2 * (size_t)D * sizeof(float);
#define LN_FWD(NC)                                                                                               \
  do {                                                                                                           \
    const int blocks = resident_grid(layernorm_fwd_kernel<T, NC>, 256, smem, cdiv(rows, 16));                    \
    launch_pdl(layernorm_fwd_kernel<T, NC>, dim3(blocks), dim3(256), smem, s, (const T*)x, (T*)y, gamma, beta, mean, rstd, \
               rows, D, eps);                                                                                    \
  } while (0)
  if (need <= 1) LN_FWD(1);
  else if (need <= 2) LN_FWD(2);
  else if (need <= 3) LN_FWD(3);
  else if (need <= 4) LN_FWD(4);
  else if (need <= 6) LN_FWD(6);
  else if (need <= 8) LN_FWD(8);
  else {
    set_error("layernorm: D=%d too large", D);
    return VVAE_ERR_UNSUPPORTED;
  }
#undef LN_FWD
  return check_launch("layernorm_fwd");
}

template <typename T>
static int ln_bwd_dispatch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                           const void* dres, void* dx, float* dgamma, float* dbeta, long long rows, int D,
                           cudaStream_t s) {
  constexpr int V = Vec16<T>::N;
  if constexpr (std::is_same<T, bf16>::value) {
    const int mode = (int)g_dbg[7] & 15;
    if (D == 768 && mode != 1 && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dx % 16 == 0) &&
        (!dres || (uintptr_t)dres % 16 == 0)) {      // vvae_debug_set(7, 1): register-staged kernel
      // vvae_debug_set(7, 2): 8 warps x 4 stages instead of 12 x 3 (same bytes in flight per SM): 50.1 vs 49.2 us
      // (16 warps x 2 stages is capped at 128 registers, spills and runs at 64 us)
      const int rc = mode == 2 ? launch_ln_bwd_stream<3, 8, 4>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, s)
                                   : launch_ln_bwd_stream<3, 12, 3>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, s);
      if (rc != VVAE_OK) return rc;
      return check_launch("layernorm_bwd");
    }
  }
  const int need = (int)cdiv(D / V, 32);
  const size_t smem = 3 * (size_t)D * sizeof(float);
#define LN_BWD(NC)                                                                                              \
  do {                                                                                                          \
    const int blocks = resident_grid(layernorm_bwd_kernel<T, NC>, 256, smem, cdiv(rows, 8));                    \
    launch_pdl(layernorm_bwd_kernel<T, NC>, dim3(blocks), dim3(256), smem, s, (const T*)dy, (const T*)x, mean, rstd,    \
               gamma, (const T*)dres, (T*)dx, dgamma, dbeta, rows, D);                                          \
  } while (0)
  if (need <= 1) LN_BWD(1);
  else if (need <= 2) LN_BWD(2);
  else if (need <= 3) LN_BWD(3);
  else if (need <= 4) LN_BWD(4);
  else if (need <= 6) LN_BWD(6);
  else {
    set_error("layernorm_bwd: D=%d too large", D);
    return VVAE_ERR_UNSUPPORTED;
  }
#undef LN_BWD
  return check_launch("layernorm_bwd");
}

extern "C" {

int vvae_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                       long long rows, int D, float eps, int dtype, vvae_stream_t stream) {
  if (rows <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && y && D > 0, "layernorm_fwd: bad arguments");
  VVAE_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), "layernorm_fwd: pointers must be 16-byte aligned");
  if (dtype == VVAE_F32) {
    VVAE_REQUIRE(D % 4 == 0, "layernorm_fwd: D=%d must be a multiple of 4 (fp32)", D);
    return ln_fwd_dispatch<float>(x, y, gamma, beta, mean, rstd, rows, D, eps, as_stream(stream));
  }
  VVAE_REQUIRE(dtype == VVAE_BF16, "layernorm_fwd: bad dtype");
  VVAE_REQUIRE(D % 8 == 0, "layernorm_fwd: D=%d must be a multiple of 8 (bf16)", D);
  return ln_fwd_dispatch<bf16>(x, y, gamma, beta, mean, rstd, rows, D, eps, as_stream(stream));
}

int vvae_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                       const void* dres, void* dx, float* dgamma, float* dbeta, long long rows, int D, int dtype,
                       vvae_stream_t stream) {
  if (rows <= 0) return VVAE_OK;
  VVAE_REQUIRE(dy && x && mean && rstd && dx && D > 0, "layernorm_bwd: bad arguments");
  if (dtype == VVAE_F32) {
    VVAE_REQUIRE(D % 4 == 0, "layernorm_bwd: D=%d must be a multiple of 4 (fp32)", D);
    return ln_bwd_dispatch<float>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, D, as_stream(stream));
  }
  VVAE_REQUIRE(dtype == VVAE_BF16, "layernorm_bwd: bad dtype");
  VVAE_REQUIRE(D % 8 == 0, "layernorm_bwd: D=%d must be a multiple of 8 (bf16)", D);
  return ln_bwd_dispatch<bf16>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, D, as_stream(stream));
}

static bool qk_fast_ok(int dtype, int heads, int hd, const void* a, const void* b, const void* c, const void* d) {
  if (dtype != VVAE_BF16 || hd != 64 || heads < 2 || (heads & 1) || 16 * heads > 256 || 256 % (16 * heads) != 0) return false;   // a warp must stay inside one row
  return ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)c % 16 == 0) && ((uintptr_t)d % 16 == 0);
}

int vvae_qknorm_rope_fwd(const void* qkv, void* qk_out, const float* q_scale, const float* k_scale,
                         const void* cos_tab, const void* sin_tab, long long rows, int heads, int hd,
                         long long pos_div, int pos_mod, float eps, int dtype, vvae_stream_t stream) {
  if (rows <= 0) return VVAE_OK;
  VVAE_REQUIRE(qkv && qk_out && q_scale && k_scale && cos_tab && sin_tab, "qknorm_rope_fwd: null pointer");
  VVAE_REQUIRE(hd % 2 == 0 && hd <= 64 * QK_MAXP && pos_div > 0 && pos_mod > 0, "qknorm_rope_fwd: bad hd=%d", hd);
  if (qk_fast_ok(dtype, heads, hd, qkv, qk_out, cos_tab, sin_tab) && rows < (1LL << 31) && pos_div < (1LL << 31)) {
    const int rpi = 256 / (16 * heads);
    static int occ_grid = 0;
    if (!occ_grid) occ_grid = resident_grid(qknorm_rope_fwd_hd64_kernel, 256, 0, 1 << 30);
    const int blocks = (int)std::min<long long>(cdiv(rows, rpi), occ_grid);
    launch_pdl(qknorm_rope_fwd_hd64_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), (const bf16*)qkv, (bf16*)qk_out,
               q_scale, k_scale, (const bf16*)cos_tab, (const bf16*)sin_tab, rows, heads, pos_div, pos_mod, eps);
    return check_launch("qknorm_rope_fwd");
  }
  const long long nvec = rows * 2 * heads;
  const int blocks = (int)std::min<long long>(cdiv(nvec, 8), (long long)num_sms() * 16);
  VVAE_DISPATCH_DTYPE(dtype, T, (qknorm_rope_fwd_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(
                                    (const T*)qkv, (T*)qk_out, q_scale, k_scale, (const T*)cos_tab, (const T*)sin_tab, rows,
                                    heads, hd, pos_div, pos_mod, eps)));
  return check_launch("qknorm_rope_fwd");
}

int vvae_qknorm_rope_bwd(void* dqkv, const void* qkv, const float* q_scale, const float* k_scale,
                         const void* cos_tab, const void* sin_tab, float* dq_scale, float* dk_scale, float* dbias_qk,
                         long long rows, int heads, int hd, long long pos_div, int pos_mod, float eps, int dtype,
                         vvae_stream_t stream) {
  if (rows <= 0) return VVAE_OK;
  VVAE_REQUIRE(dqkv && qkv && q_scale && k_scale && cos_tab && sin_tab, "qknorm_rope_bwd: null pointer");
  VVAE_REQUIRE(hd % 2 == 0 && hd <= 64 * QK_MAXP && pos_div > 0 && pos_mod > 0, "qknorm_rope_bwd: bad hd=%d", hd);
  if (qk_fast_ok(dtype, heads, hd, qkv, dqkv, cos_tab, sin_tab) && heads == 8 && !(g_dbg[7] & 0x100)) {
    // vvae_debug_set(7, 0x100): the register-staged kernel below
    // without the bias-gradient sums (the product takes them from the weight-gradient GEMM) the kernel fits 128 registers:
    // 16 warps x 2 stages; with them 8 warps x 4 stages
    int rc;
    if (dbias_qk) rc = launch_qk_bwd_stream<8, 4, true>(dqkv, qkv, q_scale, k_scale, cos_tab, sin_tab, dq_scale, dk_scale, dbias_qk,
                                                        rows, pos_div, pos_mod, eps, as_stream(stream));
    else if (g_dbg[7] & 0x200)
      rc = launch_qk_bwd_stream<8, 4, false>(dqkv, qkv, q_scale, k_scale, cos_tab, sin_tab, dq_scale, dk_scale, dbias_qk, rows,
                                             pos_div, pos_mod, eps, as_stream(stream));
    else rc = launch_qk_bwd_stream<16, 2, false>(dqkv, qkv, q_scale, k_scale, cos_tab, sin_tab, dq_scale, dk_scale, dbias_qk, rows,
                                                 pos_div, pos_mod, eps, as_stream(stream));
    if (rc) return rc;
    return check_launch("qknorm_rope_bwd");
  }
  if (qk_fast_ok(dtype, heads, hd, qkv, dqkv, cos_tab, sin_tab) && rows < (1LL << 31) && pos_div < (1LL << 31)) {
    const int rpi = 256 / (16 * heads);
    static int occ_grid = 0;
    if (!occ_grid) occ_grid = resident_grid(qknorm_rope_bwd_hd64_kernel, 256, 0, 1 << 30);
    const int blocks = (int)std::min<long long>(cdiv(rows, rpi), occ_grid);
    launch_pdl(qknorm_rope_bwd_hd64_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), (bf16*)dqkv, (const bf16*)qkv,
               q_scale, k_scale, (const bf16*)cos_tab, (const bf16*)sin_tab, dq_scale, dk_scale, dbias_qk, rows, heads,
               pos_div, pos_mod, eps);
    return check_launch("qknorm_rope_bwd");
  }
  const long long nvec = rows * 2 * heads;
  const int blocks = (int)std::min<long long>(cdiv(nvec, 8), (long long)num_sms() * 8);
  VVAE_DISPATCH_DTYPE(dtype, T, (qknorm_rope_bwd_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(
                                    (T*)dqkv, (const T*)qkv, q_scale, k_scale, (const T*)cos_tab, (const T*)sin_tab, dq_scale,
                                    dk_scale, rows, heads, hd, pos_div, pos_mod, eps)));
  int rc = check_launch("qknorm_rope_bwd");
  if (rc || !dbias_qk) return rc;
  // generic path: the bias gradient of the q|k columns is a separate column-sum pass over the produced gradient
  return vvae_colsum(dqkv, 3LL * heads * hd, rows, 2 * heads * hd, dbias_qk, dtype, stream);
}

int vvae_groupnorm_silu_fwd(const void* x, void* y, long long y_ld, const float* gamma, const float* beta, float* mean,
                            float* rstd, float* stats, int B, long long S, int C, int G, float eps, int dtype,
                            vvae_stream_t stream) {
  if (B <= 0 || S <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && y && gamma && beta && mean && rstd && stats, "groupnorm_silu_fwd: null pointer");
  VVAE_REQUIRE(C > 0 && C <= 256 && G > 0 && G <= 64 && C % G == 0, "groupnorm_silu_fwd: bad C=%d G=%d", C, G);
  cudaStream_t s = as_stream(stream);
  const int threads = gn_threads(C);
  const long long rpb = std::max<long long>(threads / C, cdiv(S, std::max<long long>(1, ((long long)num_sms() * 8) / B)));
  dim3 grid((unsigned)cdiv(S, rpb), (unsigned)B);
  int rc = vvae_fill_f32(stats, 0.f, (long long)B * G * 2, stream);
  if (rc) return rc;
  if (gn_vec_ok(dtype, C, x, C, y, y_ld)) {
    const int rpi = 256 / (C / 8);
    const long long vrpb = std::max<long long>(2 * rpi, cdiv(S, std::max<long long>(1, ((long long)num_sms() * 8) / B)));
    dim3 vgrid((unsigned)cdiv(S, vrpb), (unsigned)B);
    const bool v4 = C <= 128 && !(g_dbg[7] & 0x400);   // 4 channels per thread (vvae_debug_set(7, 0x400): 8)
    if (v4) gn_stats_vec_kernel<4><<<vgrid, 256, 0, s>>>((const bf16*)x, stats, S, C, G, vrpb);
    else gn_stats_vec_kernel<8><<<vgrid, 256, 0, s>>>((const bf16*)x, stats, S, C, G, vrpb);
    groupnorm_finalize_kernel<<<(int)cdiv(B * G, 128), 128, 0, s>>>(stats, mean, rstd, B * G,
                                                                   1.f / (float)((double)S * (C / G)), eps);
    if (v4) gn_apply_vec_kernel<4><<<vgrid, 256, 0, s>>>((const bf16*)x, (bf16*)y, y_ld, gamma, beta, mean, rstd, S, C, G, vrpb);
    else gn_apply_vec_kernel<8><<<vgrid, 256, 0, s>>>((const bf16*)x, (bf16*)y, y_ld, gamma, beta, mean, rstd, S, C, G, vrpb);
    return check_launch("groupnorm_silu_fwd");
  }
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_stats_kernel<T><<<grid, threads, 0, s>>>((const T*)x, stats, S, C, G, rpb)));
  groupnorm_finalize_kernel<<<(int)cdiv(B * G, 128), 128, 0, s>>>(stats, mean, rstd, B * G,
                                                                 1.f / (float)((double)S * (C / G)), eps);
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_silu_apply_kernel<T><<<grid, threads, 0, s>>>(
                                    (const T*)x, (T*)y, y_ld, gamma, beta, mean, rstd, S, C, G, rpb)));
  return check_launch("groupnorm_silu_fwd");
}

int vvae_groupnorm_silu_bwd(const void* dy, long long dy_ld, const void* x, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, float* stats,
                            float* dx_colsum_accum, int B, long long S, int C, int G, int dtype, vvae_stream_t stream) {
  if (B <= 0 || S <= 0) return VVAE_OK;
  VVAE_REQUIRE(dy && x && gamma && beta && mean && rstd && dx && stats, "groupnorm_silu_bwd: null pointer");
  VVAE_REQUIRE(C > 0 && C <= 256 && G > 0 && G <= 64 && C % G == 0, "groupnorm_silu_bwd: bad C=%d G=%d", C, G);
  cudaStream_t s = as_stream(stream);
  const int threads = gn_threads(C);
  const long long rpb = std::max<long long>(threads / C, cdiv(S, std::max<long long>(1, ((long long)num_sms() * 8) / B)));
  dim3 grid((unsigned)cdiv(S, rpb), (unsigned)B);
  int rc = vvae_fill_f32(stats, 0.f, (long long)B * G * 2, stream);
  if (rc) return rc;
  if (gn_vec_ok(dtype, C, dy, dy_ld, x, C) && ((uintptr_t)dx % 16 == 0)) {
    const int rpi = 256 / (C / 8);
    const long long vrpb = std::max<long long>(2 * rpi, cdiv(S, std::max<long long>(1, ((long long)num_sms() * 8) / B)));
    dim3 vgrid((unsigned)cdiv(S, vrpb), (unsigned)B);
    // narrow maps: 4 channels per thread (vvae_debug_set(7, 0x400) keeps 8)
    if (C <= 128 && !(g_dbg[7] & 0x400)) {
      gn_bwd_stats_vec_kernel<4><<<vgrid, 256, 0, s>>>((const bf16*)dy, dy_ld, (const bf16*)x, gamma, beta, mean, rstd, stats,
                                                       dgamma, dbeta, S, C, G, vrpb);
      gn_bwd_apply_vec_kernel<4><<<vgrid, 256, 0, s>>>((const bf16*)dy, dy_ld, (const bf16*)x, gamma, beta, mean, rstd,
                                                       stats, (bf16*)dx, dx_colsum_accum, S, C, G,
                                                       1.f / (float)((double)S * (C / G)), vrpb);
    } else {
      gn_bwd_stats_vec_kernel<8><<<vgrid, 256, 0, s>>>((const bf16*)dy, dy_ld, (const bf16*)x, gamma, beta, mean, rstd, stats,
                                                       dgamma, dbeta, S, C, G, vrpb);
      gn_bwd_apply_vec_kernel<8><<<vgrid, 256, 0, s>>>((const bf16*)dy, dy_ld, (const bf16*)x, gamma, beta, mean, rstd,
                                                       stats, (bf16*)dx, dx_colsum_accum, S, C, G,
                                                       1.f / (float)((double)S * (C / G)), vrpb);
    }
    return check_launch("groupnorm_silu_bwd");
  }
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_silu_bwd_stats_kernel<T><<<grid, threads, 0, s>>>(
                                    (const T*)dy, dy_ld, (const T*)x, gamma, beta, mean, rstd, stats, dgamma, dbeta, S, C,
                                    G, rpb)));
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_silu_bwd_apply_kernel<T><<<grid, threads, 0, s>>>(
                                    (const T*)dy, dy_ld, (const T*)x, gamma, beta, mean, rstd, stats, (T*)dx, S, C, G,
                                    1.f / (float)((double)S * (C / G)), rpb)));
  rc = check_launch("groupnorm_silu_bwd");
  if (rc || !dx_colsum_accum) return rc;
  return vvae_colsum(dx, C, (long long)B * S, C, dx_colsum_accum, dtype, stream);   // generic path: separate pass
}

}  // extern "C"
