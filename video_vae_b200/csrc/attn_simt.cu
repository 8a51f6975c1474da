// Generic attention (any head_dim <= 128, any L, arbitrary broadcastable mask, fp32 or bf16 storage), fp32 math.
// Flash-style streaming softmax on CUDA cores.  This is the any-shape / fp32 path and the HBM-bound temporal
// path (L = 16..64); semantics follow jax.nn.dot_product_attention (see include/vvae.h).
#include <cfloat>

#include "common.cuh"

namespace vvae {

constexpr int AT_QB = 16;   // queries per block (4 per warp)
constexpr int AT_KB = 32;   // keys per tile (one per lane)
constexpr int AT_MAXC = 4;  // head-dim columns per lane -> hd <= 128
#define AT_BIG_NEG (-0.7f * FLT_MAX)

struct AttnGeom {
  int n_inner, L, H, hd;
  long long ts_o, ts_i, ts_l;
  const unsigned char* mask; long long mask_seq_div, ms_seq, ms_head, ms_q, ms_k;
  float scale;
  __device__ __forceinline__ long long tok(int seq, int l) const {
    return (long long)(seq / n_inner) * ts_o + (long long)(seq % n_inner) * ts_i + (long long)l * ts_l;
  }
  __device__ __forceinline__ bool attend(int seq, int h, int qi, int kj) const {
    if (!mask) return true;
    return mask[(seq / mask_seq_div) * ms_seq + h * ms_head + qi * ms_q + kj * ms_k] != 0;
  }
};

// stage `n` rows (positions l0..l0+n) of head h into smem as fp32 [rows][hd+1]; rows beyond L are zero.
template <typename T>
__device__ __forceinline__ void stage_rows(float* dst, int ld, const T* src, long long rs, const AttnGeom& g, int seq, int h,
                                           int l0, int n) {
  for (int e = threadIdx.x; e < n * g.hd; e += blockDim.x) {
    int r = e / g.hd, d = e % g.hd;
    int l = l0 + r;
    dst[r * ld + d] = (l < g.L) ? to_f(src[g.tok(seq, l) * rs + (long long)h * g.hd + d]) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(AttnGeom g, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                long long q_rs, long long k_rs, long long v_rs, long long o_rs, float* __restrict__ lse) {
  extern __shared__ float sm[];
  const int hd = g.hd, ldk = hd + 1;
  float* sK = sm;                       // [KB][hd+1]
  float* sV = sK + AT_KB * ldk;         // [KB][hd+1]
  float* sQ = sV + AT_KB * ldk;         // [QB][hd]
  const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(sQ, hd, q, q_rs, g, seq, h, q0, AT_QB);
  float m[4], l[4], acc[4][AT_MAXC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) acc[i][c] = 0.f;
  }
  for (int k0 = 0; k0 < g.L; k0 += AT_KB) {
    __syncthreads();
    stage_rows(sK, ldk, k, k_rs, g, seq, h, k0, AT_KB);
    stage_rows(sV, ldk, v, v_rs, g, seq, h, k0, AT_KB);
    __syncthreads();
    const int kj = k0 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + warp * 4 + i;
      if (qi >= g.L) break;  // warp-uniform
      const float* qv = sQ + (warp * 4 + i) * hd;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(qv[d], sK[lane * ldk + d], s);
      s *= g.scale;
      const bool valid = kj < g.L;
      if (valid && !g.attend(seq, h, qi, kj)) s = AT_BIG_NEG;
      float tmax = warp_max(valid ? s : -INFINITY);
      float mn = fmaxf(m[i], tmax);
      float p = valid ? __expf(s - mn) : 0.f;
      float corr = __expf(m[i] - mn);
      l[i] = l[i] * corr + warp_sum(p);
      m[i] = mn;
      p = round_to<T>(p);
#pragma unroll
      for (int c = 0; c < AT_MAXC; ++c) acc[i][c] *= corr;
      for (int jj = 0; jj < AT_KB; ++jj) {
        float pj = __shfl_sync(0xffffffffu, p, jj);
#pragma unroll
        for (int c = 0; c < AT_MAXC; ++c) {
          int d = lane + 32 * c;
          if (d < hd) acc[i][c] = fmaf(pj, sV[jj * ldk + d], acc[i][c]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + warp * 4 + i;
    if (qi >= g.L) break;
    const float inv = 1.f / l[i];
    T* orow = o + g.tok(seq, qi) * o_rs + (long long)h * hd;
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) {
      int d = lane + 32 * c;
      if (d < hd) orow[d] = from_f<T>(acc[i][c] * inv);
    }
    if (lane == 0 && lse) lse[((long long)seq * g.H + h) * g.L + qi] = m[i] + __logf(l[i]);
  }
}

// dQ (and delta = rowsum(dO * O), consumed by the dK/dV kernel).
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(AttnGeom g, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                   const T* __restrict__ o, const T* __restrict__ d_o, T* __restrict__ dq, long long q_rs,
                   long long k_rs, long long v_rs, long long o_rs, long long do_rs, long long dq_rs,
                   const float* __restrict__ lse, float* __restrict__ delta) {
  extern __shared__ float sm[];
  const int hd = g.hd, ldk = hd + 1;
  float* sK = sm;
  float* sV = sK + AT_KB * ldk;
  float* sQ = sV + AT_KB * ldk;   // [QB][hd]
  float* sdO = sQ + AT_QB * hd;   // [QB][hd]
  const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(sQ, hd, q, q_rs, g, seq, h, q0, AT_QB);
  stage_rows(sdO, hd, d_o, do_rs, g, seq, h, q0, AT_QB);
  __syncthreads();
  float acc[4][AT_MAXC], ls[4], dl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + warp * 4 + i;
    ls[i] = 0.f;
    dl[i] = 0.f;
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) acc[i][c] = 0.f;
    if (qi < g.L) {
      ls[i] = lse[((long long)seq * g.H + h) * g.L + qi];
      const T* orow = o + g.tok(seq, qi) * o_rs + (long long)h * hd;
      float part = 0.f;
      for (int d = lane; d < hd; d += 32) part += to_f(orow[d]) * sdO[(warp * 4 + i) * hd + d];
      dl[i] = warp_sum(part);
      if (lane == 0) delta[((long long)seq * g.H + h) * g.L + qi] = dl[i];
    }
  }
  for (int k0 = 0; k0 < g.L; k0 += AT_KB) {
    __syncthreads();
    stage_rows(sK, ldk, k, k_rs, g, seq, h, k0, AT_KB);
    stage_rows(sV, ldk, v, v_rs, g, seq, h, k0, AT_KB);
    __syncthreads();
    const int kj = k0 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + warp * 4 + i;
      if (qi >= g.L) break;
      const float* qv = sQ + (warp * 4 + i) * hd;
      const float* dv_ = sdO + (warp * 4 + i) * hd;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) {
        s = fmaf(qv[d], sK[lane * ldk + d], s);
        dp = fmaf(dv_[d], sV[lane * ldk + d], dp);
      }
      s *= g.scale;
      const bool valid = kj < g.L;
      const bool masked = valid && !g.attend(seq, h, qi, kj);
      if (masked) s = AT_BIG_NEG;
      float p = valid ? __expf(s - ls[i]) : 0.f;
      if (masked && ls[i] <= 0.5f * AT_BIG_NEG) p = 1.f / (float)g.L;   // fully masked row: uniform
      float ds = masked ? 0.f : p * (dp - dl[i]);                       // where(mask, logits, const): no gradient
      for (int jj = 0; jj < AT_KB; ++jj) {
        float dj = __shfl_sync(0xffffffffu, ds, jj);
#pragma unroll
        for (int c = 0; c < AT_MAXC; ++c) {
          int d = lane + 32 * c;
          if (d < hd) acc[i][c] = fmaf(dj, sK[jj * ldk + d], acc[i][c]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + warp * 4 + i;
    if (qi >= g.L) break;
    T* row = dq + g.tok(seq, qi) * dq_rs + (long long)h * hd;
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) {
      int d = lane + 32 * c;
      if (d < hd) row[d] = from_f<T>(acc[i][c] * g.scale);
    }
  }
}

// dK, dV: a block owns 16 keys (4 per warp) and streams over query tiles (one query per lane).
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(AttnGeom g, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const T* __restrict__ d_o, T* __restrict__ dk, T* __restrict__ dv, long long q_rs, long long k_rs,
                    long long v_rs, long long do_rs, long long dk_rs, long long dv_rs, const float* __restrict__ lse,
                    const float* __restrict__ delta) {
  extern __shared__ float sm[];
  const int hd = g.hd, ldk = hd + 1;
  float* sQ = sm;                    // [KB(queries)][hd+1]
  float* sdO = sQ + AT_KB * ldk;     // [KB][hd+1]
  float* sK = sdO + AT_KB * ldk;     // [QB(keys)][hd]
  float* sV = sK + AT_QB * hd;       // [QB][hd]
  const int seq = blockIdx.z, h = blockIdx.y, kb0 = blockIdx.x * AT_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(sK, hd, k, k_rs, g, seq, h, kb0, AT_QB);
  stage_rows(sV, hd, v, v_rs, g, seq, h, kb0, AT_QB);
  float adk[4][AT_MAXC], adv[4][AT_MAXC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) adk[i][c] = adv[i][c] = 0.f;
  for (int q0 = 0; q0 < g.L; q0 += AT_KB) {
    __syncthreads();
    stage_rows(sQ, ldk, q, q_rs, g, seq, h, q0, AT_KB);
    stage_rows(sdO, ldk, d_o, do_rs, g, seq, h, q0, AT_KB);
    __syncthreads();
    const int qi = q0 + lane;
    const bool qvalid = qi < g.L;
    const float ls = qvalid ? lse[((long long)seq * g.H + h) * g.L + qi] : 0.f;
    const float dl = qvalid ? delta[((long long)seq * g.H + h) * g.L + qi] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kj = kb0 + warp * 4 + i;
      if (kj >= g.L) break;
      const float* kv = sK + (warp * 4 + i) * hd;
      const float* vv = sV + (warp * 4 + i) * hd;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) {
        s = fmaf(sQ[lane * ldk + d], kv[d], s);
        dp = fmaf(sdO[lane * ldk + d], vv[d], dp);
      }
      s *= g.scale;
      const bool masked = qvalid && !g.attend(seq, h, qi, kj);
      if (masked) s = AT_BIG_NEG;
      float p = qvalid ? __expf(s - ls) : 0.f;
      if (masked && ls <= 0.5f * AT_BIG_NEG) p = 1.f / (float)g.L;
      float ds = masked ? 0.f : p * (dp - dl);
      float pr = round_to<T>(p);
      for (int ii = 0; ii < AT_KB; ++ii) {
        float pi = __shfl_sync(0xffffffffu, pr, ii);
        float di = __shfl_sync(0xffffffffu, ds, ii);
#pragma unroll
        for (int c = 0; c < AT_MAXC; ++c) {
          int d = lane + 32 * c;
          if (d < hd) {
            adv[i][c] = fmaf(pi, sdO[ii * ldk + d], adv[i][c]);
            adk[i][c] = fmaf(di, sQ[ii * ldk + d], adk[i][c]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kj = kb0 + warp * 4 + i;
    if (kj >= g.L) break;
    T* rk = dk + g.tok(seq, kj) * dk_rs + (long long)h * hd;
    T* rv = dv + g.tok(seq, kj) * dv_rs + (long long)h * hd;
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) {
      int d = lane + 32 * c;
      if (d < hd) {
        rk[d] = from_f<T>(adk[i][c] * g.scale);
        rv[d] = from_f<T>(adv[i][c]);
      }
    }
  }
}

static AttnGeom make_geom(const vvae_attn_args& a) {
  AttnGeom g;
  g.n_inner = a.n_inner; g.L = a.L; g.H = a.heads; g.hd = a.hd;
  g.ts_o = a.tok_stride_outer; g.ts_i = a.tok_stride_inner; g.ts_l = a.tok_stride_pos;
  g.mask = a.mask; g.mask_seq_div = a.mask_seq_div > 0 ? a.mask_seq_div : 1;
  g.ms_seq = a.ms_seq; g.ms_head = a.ms_head; g.ms_q = a.ms_q; g.ms_k = a.ms_k;
  g.scale = a.scale;
  return g;
}

template <typename K>
static int ensure_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("attention: cannot get %zu bytes of shared memory: %s", bytes, cudaGetErrorString(e));
      return VVAE_ERR_CUDA;
    }
  }
  return VVAE_OK;
}

int attn_simt_fwd(const vvae_attn_args& a, cudaStream_t s) {
  AttnGeom g = make_geom(a);
  const int n_seq = a.n_outer * a.n_inner;
  dim3 grid((unsigned)cdiv(a.L, AT_QB), (unsigned)a.heads, (unsigned)n_seq);
  size_t smem = ((size_t)2 * AT_KB * (a.hd + 1) + (size_t)AT_QB * a.hd) * sizeof(float);
  VVAE_DISPATCH_DTYPE(a.dtype, T, {
    int rc = ensure_smem(attn_fwd_kernel<T>, smem);
    if (rc) return rc;
    attn_fwd_kernel<T><<<grid, 128, smem, s>>>(g, (const T*)a.q, (const T*)a.k, (const T*)a.v, (T*)a.o, a.q_rs, a.k_rs,
                                               a.v_rs, a.o_rs, a.lse);
  });
  return check_launch("attn_fwd");
}

int attn_simt_bwd(const vvae_attn_args& a, cudaStream_t s) {
  AttnGeom g = make_geom(a);
  const int n_seq = a.n_outer * a.n_inner;
  dim3 grid((unsigned)cdiv(a.L, AT_QB), (unsigned)a.heads, (unsigned)n_seq);
  size_t smem = ((size_t)2 * AT_KB * (a.hd + 1) + (size_t)2 * AT_QB * a.hd) * sizeof(float);
  VVAE_DISPATCH_DTYPE(a.dtype, T, {
    int rc = ensure_smem(attn_bwd_dq_kernel<T>, smem);
    if (rc) return rc;
    rc = ensure_smem(attn_bwd_dkv_kernel<T>, smem);
    if (rc) return rc;
    attn_bwd_dq_kernel<T><<<grid, 128, smem, s>>>(g, (const T*)a.q, (const T*)a.k, (const T*)a.v, (const T*)a.o,
                                                  (const T*)a.d_o, (T*)a.dq, a.q_rs, a.k_rs, a.v_rs, a.o_rs, a.do_rs,
                                                  a.dq_rs, a.lse, a.delta);
    attn_bwd_dkv_kernel<T><<<grid, 128, smem, s>>>(g, (const T*)a.q, (const T*)a.k, (const T*)a.v, (const T*)a.d_o,
                                                   (T*)a.dk, (T*)a.dv, a.q_rs, a.k_rs, a.v_rs, a.do_rs, a.dk_rs,
                                                   a.dv_rs, a.lse, a.delta);
  });
  return check_launch("attn_bwd");
}

}  // namespace vvae

namespace vvae {
int attn_tc_supported(const vvae_attn_args& a);
int attn_tc_fwd(const vvae_attn_args& a, cudaStream_t s);
int attn_tc_bwd_supported(const vvae_attn_args& a);
int attn_tc_bwd(const vvae_attn_args& a, cudaStream_t s);
// attn_warp.cu: one warp per (sequence, head) for L <= 16 (mma.sync, registers only)
int attn_warp_supported(const vvae_attn_args& a, bool bwd);
int attn_warp_fwd(const vvae_attn_args& a, cudaStream_t s);
int attn_warp_bwd(const vvae_attn_args& a, cudaStream_t s);
extern long long g_dbg[32];      // vvae_debug_set: key 9 != 0 keeps short sequences on the tcgen05 packed-tile kernels
}

using namespace vvae;

static int attn_validate(const vvae_attn_args* a, bool bwd) {
  VVAE_REQUIRE(a, "attention: null args");
  VVAE_REQUIRE(a->n_outer >= 0 && a->n_inner > 0 && a->L > 0 && a->heads > 0, "attention: bad extents");
  VVAE_REQUIRE(a->hd > 0 && a->hd <= 32 * AT_MAXC, "attention: head_dim %d unsupported (max %d)", a->hd, 32 * AT_MAXC);
  VVAE_REQUIRE(a->q && a->k && a->v && a->o && a->lse, "attention: null tensor");
  if (bwd) VVAE_REQUIRE(a->d_o && a->dq && a->dk && a->dv && a->delta, "attention bwd: null tensor");
  return VVAE_OK;
}

extern "C" {

int vvae_attn_fwd(const vvae_attn_args* args, vvae_stream_t stream) {
  int rc = attn_validate(args, false);
  if (rc) return rc;
  if (args->n_outer == 0) return VVAE_OK;
  // L <= 16: one warp per (sequence, head).  L = 32 / 64 (and 128 ..): tcgen05 tiles.  Any other L <= 64 (the frame-count
  // curriculum of train/rl_nonadversarial.py:287-295 produces them): the multi-block warp kernels.
  if (args->backend == VVAE_BACKEND_AUTO && !g_dbg[9] && attn_warp_supported(*args, false) &&
      (args->L <= 16 || !attn_tc_supported(*args)))
    return attn_warp_fwd(*args, as_stream(stream));
  if (args->backend != VVAE_BACKEND_SIMT && attn_tc_supported(*args)) return attn_tc_fwd(*args, as_stream(stream));
  if (args->backend == VVAE_BACKEND_TCGEN05) {
    set_error("attention: shape not supported by the tensor-core path");
    return VVAE_ERR_UNSUPPORTED;
  }
  return attn_simt_fwd(*args, as_stream(stream));
}

int vvae_attn_bwd(const vvae_attn_args* args, vvae_stream_t stream) {
  int rc = attn_validate(args, true);
  if (rc) return rc;
  if (args->n_outer == 0) return VVAE_OK;
  if (args->backend == VVAE_BACKEND_AUTO && !g_dbg[9] && attn_warp_supported(*args, true) &&
      (args->L <= 16 || !attn_tc_bwd_supported(*args)))
    return attn_warp_bwd(*args, as_stream(stream));
  if (args->backend != VVAE_BACKEND_SIMT && attn_tc_bwd_supported(*args)) return attn_tc_bwd(*args, as_stream(stream));
  if (args->backend == VVAE_BACKEND_TCGEN05) {
    set_error("attention bwd: shape not supported by the tensor-core path");
    return VVAE_ERR_UNSUPPORTED;
  }
  return attn_simt_bwd(*args, as_stream(stream));
}

}  // extern "C"
