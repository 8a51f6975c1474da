// HBM-bound normalisation kernels: LayerNorm fwd/bwd, QK-LayerNorm + RoPE fwd/bwd, GroupNorm + SiLU fwd/bwd.
// Statistics follow Flax: fp32, "fast variance" max(0, E[x^2] - E[x]^2), eps inside the rsqrt.
#include "common.cuh"

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
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = D / V;
  for (long long row = warp0; row < rows; row += nwarps) {
    const T* xr = x + row * D;
    Vec16<T> v[NCHUNK];
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      int j = lane + 32 * i;
      if (j < nvec) {
        v[i].load(xr + j * V);
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float f = v[i].get(t);
          s += f;
          s2 += f * f;
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
      int j = lane + 32 * i;
      if (j < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int t = 0; t < V; ++t) {
          int c = j * V + t;
          float f = (v[i].get(t) - mu) * r;
          if (gamma) f *= gamma[c];
          if (beta) f += beta[c];
          o.set(t, f);
        }
        o.store(yr + j * V);
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
  extern __shared__ float red[];  // [2][D]
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = D / V;
  float pg[NCHUNK][V], pb[NCHUNK][V];
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i)
#pragma unroll
    for (int t = 0; t < V; ++t) pg[i][t] = pb[i][t] = 0.f;

  for (long long row = warp0; row < rows; row += nwarps) {
    const float mu = mean[row], r = rstd[row];
    Vec16<T> vx[NCHUNK], vd[NCHUNK];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      int j = lane + 32 * i;
      if (j < nvec) {
        vx[i].load(x + row * D + j * V);
        vd[i].load(dy + row * D + j * V);
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float xh = (vx[i].get(t) - mu) * r;
          float d = vd[i].get(t);
          float g = gamma ? d * gamma[j * V + t] : d;
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
        Vec16<T> o, rs;
        if (dres) rs.load(dres + row * D + j * V);
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float xh = (vx[i].get(t) - mu) * r;
          float d = vd[i].get(t);
          float g = gamma ? d * gamma[j * V + t] : d;
          float f = r * (g - sg - xh * sgx);
          if (dres) f += rs.get(t);
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
// QK-LayerNorm (scale only) + RoPE on the q|k part of a fused [rows, 3*H*hd] projection. One warp per head vector.
// =====================================================================================================
constexpr int QK_MAXP = 4;  // pairs per lane -> hd <= 256

template <typename T>
__global__ void __launch_bounds__(256)
qknorm_rope_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, const float* __restrict__ q_scale,
                       const float* __restrict__ k_scale, const float* __restrict__ cos_tab,
                       const float* __restrict__ sin_tab, long long rows, int H, int hd, long long pos_div, int pos_mod,
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
        float c1 = round_to<T>(cos_tab[(long long)pos * hd + i]), c2 = round_to<T>(cos_tab[(long long)pos * hd + i + half]);
        float s1 = round_to<T>(sin_tab[(long long)pos * hd + i]), sn2 = round_to<T>(sin_tab[(long long)pos * hd + i + half]);
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
                       const float* __restrict__ k_scale, const float* __restrict__ cos_tab,
                       const float* __restrict__ sin_tab, float* __restrict__ dq_scale, float* __restrict__ dk_scale,
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
        float c1 = round_to<T>(cos_tab[(long long)pos * hd + i]), c2 = round_to<T>(cos_tab[(long long)pos * hd + i + half]);
        float s1 = round_to<T>(sin_tab[(long long)pos * hd + i]), sn2 = round_to<T>(sin_tab[(long long)pos * hd + i + half]);
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

static inline int gn_threads(int C) { return C * (256 / C > 0 ? 256 / C : 1); }

}  // namespace vvae

using namespace vvae;

template <typename T>
static int ln_fwd_dispatch(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                           long long rows, int D, float eps, cudaStream_t s) {
  constexpr int V = Vec16<T>::N;
  const int need = (int)cdiv(D / V, 32);
  const int blocks = (int)std::min<long long>(cdiv(rows, 8), 148LL * 8);
#define LN_FWD(NC)                                                                                               \
  layernorm_fwd_kernel<T, NC><<<blocks, 256, 0, s>>>((const T*)x, (T*)y, gamma, beta, mean, rstd, rows, D, eps)
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
  const int need = (int)cdiv(D / V, 32);
  const int blocks = (int)std::min<long long>(cdiv(rows, 8), 148LL * 4);
  const size_t smem = 2 * (size_t)D * sizeof(float);
#define LN_BWD(NC)                                                                                              \
  layernorm_bwd_kernel<T, NC><<<blocks, 256, smem, s>>>((const T*)dy, (const T*)x, mean, rstd, gamma, (const T*)dres, \
                                                        (T*)dx, dgamma, dbeta, rows, D)
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

int vvae_qknorm_rope_fwd(const void* qkv, void* qk_out, const float* q_scale, const float* k_scale,
                         const float* cos_tab, const float* sin_tab, long long rows, int heads, int hd,
                         long long pos_div, int pos_mod, float eps, int dtype, vvae_stream_t stream) {
  if (rows <= 0) return VVAE_OK;
  VVAE_REQUIRE(qkv && qk_out && q_scale && k_scale && cos_tab && sin_tab, "qknorm_rope_fwd: null pointer");
  VVAE_REQUIRE(hd % 2 == 0 && hd <= 64 * QK_MAXP && pos_div > 0 && pos_mod > 0, "qknorm_rope_fwd: bad hd=%d", hd);
  const long long nvec = rows * 2 * heads;
  const int blocks = (int)std::min<long long>(cdiv(nvec, 8), 148LL * 16);
  VVAE_DISPATCH_DTYPE(dtype, T, (qknorm_rope_fwd_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(
                                    (const T*)qkv, (T*)qk_out, q_scale, k_scale, cos_tab, sin_tab, rows, heads, hd,
                                    pos_div, pos_mod, eps)));
  return check_launch("qknorm_rope_fwd");
}

int vvae_qknorm_rope_bwd(void* dqkv, const void* qkv, const float* q_scale, const float* k_scale,
                         const float* cos_tab, const float* sin_tab, float* dq_scale, float* dk_scale, long long rows,
                         int heads, int hd, long long pos_div, int pos_mod, float eps, int dtype,
                         vvae_stream_t stream) {
  if (rows <= 0) return VVAE_OK;
  VVAE_REQUIRE(dqkv && qkv && q_scale && k_scale && cos_tab && sin_tab, "qknorm_rope_bwd: null pointer");
  VVAE_REQUIRE(hd % 2 == 0 && hd <= 64 * QK_MAXP && pos_div > 0 && pos_mod > 0, "qknorm_rope_bwd: bad hd=%d", hd);
  const long long nvec = rows * 2 * heads;
  const int blocks = (int)std::min<long long>(cdiv(nvec, 8), 148LL * 8);
  VVAE_DISPATCH_DTYPE(dtype, T, (qknorm_rope_bwd_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(
                                    (T*)dqkv, (const T*)qkv, q_scale, k_scale, cos_tab, sin_tab, dq_scale, dk_scale,
                                    rows, heads, hd, pos_div, pos_mod, eps)));
  return check_launch("qknorm_rope_bwd");
}

int vvae_groupnorm_silu_fwd(const void* x, void* y, long long y_ld, const float* gamma, const float* beta, float* mean,
                            float* rstd, float* stats, int B, long long S, int C, int G, float eps, int dtype,
                            vvae_stream_t stream) {
  if (B <= 0 || S <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && y && gamma && beta && mean && rstd && stats, "groupnorm_silu_fwd: null pointer");
  VVAE_REQUIRE(C > 0 && C <= 256 && G > 0 && G <= 64 && C % G == 0, "groupnorm_silu_fwd: bad C=%d G=%d", C, G);
  cudaStream_t s = as_stream(stream);
  const int threads = gn_threads(C);
  const long long rpb = std::max<long long>(threads / C, cdiv(S, std::max<long long>(1, (148LL * 8) / B)));
  dim3 grid((unsigned)cdiv(S, rpb), (unsigned)B);
  int rc = vvae_fill_f32(stats, 0.f, (long long)B * G * 2, stream);
  if (rc) return rc;
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_stats_kernel<T><<<grid, threads, 0, s>>>((const T*)x, stats, S, C, G, rpb)));
  groupnorm_finalize_kernel<<<(int)cdiv(B * G, 128), 128, 0, s>>>(stats, mean, rstd, B * G,
                                                                 1.f / (float)((double)S * (C / G)), eps);
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_silu_apply_kernel<T><<<grid, threads, 0, s>>>(
                                    (const T*)x, (T*)y, y_ld, gamma, beta, mean, rstd, S, C, G, rpb)));
  return check_launch("groupnorm_silu_fwd");
}

int vvae_groupnorm_silu_bwd(const void* dy, long long dy_ld, const void* x, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, float* stats,
                            int B, long long S, int C, int G, int dtype, vvae_stream_t stream) {
  if (B <= 0 || S <= 0) return VVAE_OK;
  VVAE_REQUIRE(dy && x && gamma && beta && mean && rstd && dx && stats, "groupnorm_silu_bwd: null pointer");
  VVAE_REQUIRE(C > 0 && C <= 256 && G > 0 && G <= 64 && C % G == 0, "groupnorm_silu_bwd: bad C=%d G=%d", C, G);
  cudaStream_t s = as_stream(stream);
  const int threads = gn_threads(C);
  const long long rpb = std::max<long long>(threads / C, cdiv(S, std::max<long long>(1, (148LL * 8) / B)));
  dim3 grid((unsigned)cdiv(S, rpb), (unsigned)B);
  int rc = vvae_fill_f32(stats, 0.f, (long long)B * G * 2, stream);
  if (rc) return rc;
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_silu_bwd_stats_kernel<T><<<grid, threads, 0, s>>>(
                                    (const T*)dy, dy_ld, (const T*)x, gamma, beta, mean, rstd, stats, dgamma, dbeta, S, C,
                                    G, rpb)));
  VVAE_DISPATCH_DTYPE(dtype, T, (groupnorm_silu_bwd_apply_kernel<T><<<grid, threads, 0, s>>>(
                                    (const T*)dy, dy_ld, (const T*)x, gamma, beta, mean, rstd, stats, (T*)dx, S, C, G,
                                    1.f / (float)((double)S * (C / G)), rpb)));
  return check_launch("groupnorm_silu_bwd");
}

}  // extern "C"
