// Short-sequence attention (L <= 64, head_dim 64, bf16; the single-warp form below is L <= 16): temporal attention over 16 frames (train/layers.py:204-214 at
// the production shape: 2048 sequences x 8 heads of 16 tokens).  At this length the work is HBM-bound (10 KB in, 6 KB out
// per (sequence, head) in the backward), and a 128-row tcgen05 tile holding 8 block-diagonal sequences spends its time
// in fixed per-tile latency.  Here ONE WARP owns one (sequence, head): every operand lives in registers, the 16x16 score
// tile is two m16n8k16 mma.sync accumulators, no shared memory, no barriers; latency is hidden by the ~12 resident
// warps per SM, each with >= 12 independent 16-byte loads in flight.
//
// Register layouts (g = lane >> 2, tig = lane & 3):
//   R layout of X[16 rows][64]: the lane holds rows g and g+8, columns {8 tig .. 8 tig+7} and {32+8 tig .. 32+8 tig+7}
//     (two 16-byte loads per row).  For a product contracted over the 64 columns (Q.K^T, dO.V^T and their transposes)
//     the contraction index may be permuted as long as both operands agree, so register 2 kk (+1) of a row serves as the
//     low (high) k-pair of k-step kk of BOTH the A fragment (rows) and the B fragment (columns) -- no shuffles.
//   P layout of X[16 rows][64]: the lane holds rows 2 tig, 2 tig+1, 2 tig+8, 2 tig+9, columns 8 g .. 8 g+7 (one 16-byte
//     load per row).  For a product contracted over the 16 rows (P.V, dS.K, P^T.dO, dS^T.Q) the B fragment of n-tile
//     nt is element nt of the four vectors (byte permutes); the OUTPUT column order is permuted instead: accumulator
//     (nt, j) of row g is column 16 tig + 8 j + nt, i.e. 16 contiguous columns per lane -> two 16-byte stores per row.
//   The accumulator layout of a 16x16 product is already the A-fragment layout of the next one (P -> P.V), and the
//   transposed probabilities needed by dV / dK come from 8 extra MMAs (K.Q^T, V.dO^T) instead of a register transpose.
#include <float.h>

#include "common.cuh"

namespace vvae {

struct AttnWarpParams {
  const bf16 *q, *k, *v, *o, *d_o;
  bf16 *out_o, *dq, *dk, *dv;
  long long q_rs, k_rs, v_rs, o_rs, do_rs, dq_rs, dk_rs, dv_rs;
  long long ts_outer, ts_inner, ts_pos;
  float* lse;
  const unsigned char* mask;
  long long mask_seq_div, ms_seq, ms_k;
  int n_inner, L, heads;
  long long n_tasks;       // sequences * heads
  float scale;
};

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float aw_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

struct RowPair {           // R layout: rows g (lo) and g+8 (hi), 8 registers each
  uint32_t lo[8], hi[8];
};
struct RowQuad {           // P layout: rows 2tig, 2tig+1, 2tig+8, 2tig+9, 4 registers each
  uint32_t r[4][4];
};

__device__ __forceinline__ void load_row16(uint32_t (&dst)[8], const bf16* row, int tig, bool ok) {
  uint4 a = make_uint4(0, 0, 0, 0), b = a;
  if (ok) {
    a = __ldg(reinterpret_cast<const uint4*>(row + 8 * tig));
    b = __ldg(reinterpret_cast<const uint4*>(row + 32 + 8 * tig));
  }
  dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
  dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
}
__device__ __forceinline__ void load_R(RowPair& x, const bf16* base, long long rs, long long ts_pos, int g, int tig, int L) {
  load_row16(x.lo, base + (long long)g * ts_pos * rs, tig, g < L);
  load_row16(x.hi, base + (long long)(g + 8) * ts_pos * rs, tig, g + 8 < L);
}
__device__ __forceinline__ void load_P(RowQuad& x, const bf16* base, long long rs, long long ts_pos, int g, int tig, int L) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = 2 * tig + (i & 1) + 8 * (i >> 1);
    uint4 a = make_uint4(0, 0, 0, 0);
    if (row < L) a = __ldg(reinterpret_cast<const uint4*>(base + (long long)row * ts_pos * rs + 8 * g));
    x.r[i][0] = a.x; x.r[i][1] = a.y; x.r[i][2] = a.z; x.r[i][3] = a.w;
  }
}

// C[16 x 16] = X . Y^T, both in R layout (contraction over the 64 columns): c[nt] covers Y rows 8 nt .. 8 nt+7
__device__ __forceinline__ void mma_RRt(float (&c)[2][4], const RowPair& x, const RowPair& y) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[nt][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    mma16816(c[0], x.lo[2 * kk], x.hi[2 * kk], x.lo[2 * kk + 1], x.hi[2 * kk + 1], y.lo[2 * kk], y.lo[2 * kk + 1]);
    mma16816(c[1], x.lo[2 * kk], x.hi[2 * kk], x.lo[2 * kk + 1], x.hi[2 * kk + 1], y.hi[2 * kk], y.hi[2 * kk + 1]);
  }
}

// out[16 x 64] (+)= A[16 x 16] . X, A given as an A fragment, X in P layout; acc[nt][j]: see the header comment
template <bool ACCUM = false>
__device__ __forceinline__ void mma_AP(float (&acc)[8][4], const uint32_t (&a)[4], const RowQuad& x) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const uint32_t sel = (nt & 1) ? 0x7632u : 0x5410u;
    const uint32_t b0 = __byte_perm(x.r[0][nt >> 1], x.r[1][nt >> 1], sel);
    const uint32_t b1 = __byte_perm(x.r[2][nt >> 1], x.r[3][nt >> 1], sel);
    if (!ACCUM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
    }
    mma16816(acc[nt], a[0], a[1], a[2], a[3], b0, b1);
  }
}

// rows g and g+8 of a [16 x 64] accumulator -> bf16, 16 contiguous columns (16 tig ..) per lane and row
__device__ __forceinline__ void store_acc(bf16* base, long long rs, long long ts_pos, const float (&acc)[8][4], int g, int tig,
                                          int L) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = g + 8 * h;
    if (row < L) {
      bf16* p = base + (long long)row * ts_pos * rs + 16 * tig;
      uint4 v0, v1;
      v0.x = pack_bf16(acc[0][2 * h], acc[1][2 * h]); v0.y = pack_bf16(acc[2][2 * h], acc[3][2 * h]);
      v0.z = pack_bf16(acc[4][2 * h], acc[5][2 * h]); v0.w = pack_bf16(acc[6][2 * h], acc[7][2 * h]);
      v1.x = pack_bf16(acc[0][2 * h + 1], acc[1][2 * h + 1]); v1.y = pack_bf16(acc[2][2 * h + 1], acc[3][2 * h + 1]);
      v1.z = pack_bf16(acc[4][2 * h + 1], acc[5][2 * h + 1]); v1.w = pack_bf16(acc[6][2 * h + 1], acc[7][2 * h + 1]);
      *reinterpret_cast<uint4*>(p) = v0;
      *reinterpret_cast<uint4*>(p + 8) = v1;
    }
  }
}

struct TaskCtx {
  long long tok0;          // first token of the sequence
  int head;
  uint32_t valid;          // bit k: key k < L and not masked
  uint32_t inlen;          // bit k: key k < L
};

__device__ __forceinline__ TaskCtx task_ctx(const AttnWarpParams& p, long long task, int lane) {
  TaskCtx t;
  const long long seq = task / p.heads;
  t.head = (int)(task - seq * p.heads);
  const long long o = seq / p.n_inner, i = seq - o * p.n_inner;
  t.tok0 = o * p.ts_outer + i * p.ts_inner;
  t.inlen = (p.L >= 16) ? 0xffffu : ((1u << p.L) - 1u);
  bool ok = lane < p.L;
  if (ok && p.mask) ok = p.mask[(seq / p.mask_seq_div) * p.ms_seq + (long long)lane * p.ms_k] != 0;
  t.valid = __ballot_sync(0xffffffffu, ok) & 0xffffu;
  return t;
}

constexpr float AW_LOG2E = 1.4426950408889634f;
constexpr float AW_MASKED = -0.7f * FLT_MAX;

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128) attn_warp_fwd_kernel(const AttnWarpParams p) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const long long task = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (task >= p.n_tasks) return;
  const TaskCtx t = task_ctx(p, task, lane);
  const int L = p.L;
  RowPair q, k;
  RowQuad v;
  load_R(q, p.q + t.tok0 * p.q_rs + t.head * 64, p.q_rs, p.ts_pos, g, tig, L);
  load_R(k, p.k + t.tok0 * p.k_rs + t.head * 64, p.k_rs, p.ts_pos, g, tig, L);
  load_P(v, p.v + t.tok0 * p.v_rs + t.head * 64, p.v_rs, p.ts_pos, g, tig, L);
  float s[2][4];
  mma_RRt(s, q, k);
  // softmax over the 16 keys of rows g (j = 0,1) and g+8 (j = 2,3); key of s[nt][j] = 8 nt + 2 tig + (j & 1)
  float mx[2] = {-FLT_MAX, -FLT_MAX};
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int key = 8 * nt + 2 * tig + (j & 1);
      float x = s[nt][j] * p.scale;
      if (!((t.valid >> key) & 1u)) x = AW_MASKED;        // JAX: where(mask, logits, -0.7 * finfo.max)
      if (!((t.inlen >> key) & 1u)) x = -FLT_MAX;         // beyond the sequence: not a key at all
      s[nt][j] = x;
      mx[j >> 1] = fmaxf(mx[j >> 1], x);
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int key = 8 * nt + 2 * tig + (j & 1);
      float e = aw_exp2((s[nt][j] - mx[j >> 1]) * AW_LOG2E);
      if (!((t.inlen >> key) & 1u)) e = 0.f;
      s[nt][j] = e;
      sum[j >> 1] += e;
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
  }
  const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
  uint32_t pa[4];
  pa[0] = pack_bf16(s[0][0] * inv0, s[0][1] * inv0);
  pa[1] = pack_bf16(s[0][2] * inv1, s[0][3] * inv1);
  pa[2] = pack_bf16(s[1][0] * inv0, s[1][1] * inv0);
  pa[3] = pack_bf16(s[1][2] * inv1, s[1][3] * inv1);
  float acc[8][4];
  mma_AP(acc, pa, v);
  store_acc(p.out_o + t.tok0 * p.o_rs + t.head * 64, p.o_rs, p.ts_pos, acc, g, tig, L);
  if (tig == 0) {
    float* lse = p.lse + task * L;
    if (g < L) lse[g] = mx[0] + __logf(sum[0]);
    if (g + 8 < L) lse[g + 8] = mx[1] + __logf(sum[1]);
  }
}

// ------------------------------------------------------------------ backward
// probabilities and score gradients of one accumulator tile; rows of the tile are "row-side" indices (queries for the
// plain tile, keys for the transposed one), columns the other side.
//   plain      : P[q][k]   = exp(s - lse[q]),   dS = P (dP - delta[q]) scale      (lse/delta per ROW)
//   transposed : P^T[k][q] = exp(s - lse[q]),   dS^T likewise                      (lse/delta per COLUMN)
__device__ __forceinline__ float aw_prob(float s_scaled_l2, float lse_l2, bool key_valid, bool key_in, bool q_in,
                                         bool any_valid, float inv_L) {
  if (!key_in || !q_in) return 0.f;
  if (!key_valid) return any_valid ? 0.f : inv_L;         // fully masked sequence: uniform attention, no gradient
  return aw_exp2(s_scaled_l2 - lse_l2);
}

__global__ void __launch_bounds__(128) attn_warp_bwd_kernel(const AttnWarpParams p) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const long long task = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (task >= p.n_tasks) return;
  const TaskCtx t = task_ctx(p, task, lane);
  const int L = p.L;
  const bool any_valid = t.valid != 0;
  const float inv_L = 1.f / (float)L;
  const long long hoff = t.head * 64;
  RowPair q, k, v, d_o;
  load_R(q, p.q + t.tok0 * p.q_rs + hoff, p.q_rs, p.ts_pos, g, tig, L);
  load_R(k, p.k + t.tok0 * p.k_rs + hoff, p.k_rs, p.ts_pos, g, tig, L);
  load_R(v, p.v + t.tok0 * p.v_rs + hoff, p.v_rs, p.ts_pos, g, tig, L);
  load_R(d_o, p.d_o + t.tok0 * p.do_rs + hoff, p.do_rs, p.ts_pos, g, tig, L);
  float delta[2];
  {
    RowPair o;
    load_R(o, p.o + t.tok0 * p.o_rs + hoff, p.o_rs, p.ts_pos, g, tig, L);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      d0 += bf_lo(o.lo[i]) * bf_lo(d_o.lo[i]) + bf_hi(o.lo[i]) * bf_hi(d_o.lo[i]);
      d1 += bf_lo(o.hi[i]) * bf_lo(d_o.hi[i]) + bf_hi(o.hi[i]) * bf_hi(d_o.hi[i]);
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    delta[0] = d0; delta[1] = d1;
  }
  const float* lse = p.lse + task * L;
  float lse_r[2];                                   // rows g, g+8 (log2 units)
  lse_r[0] = (g < L) ? lse[g] * AW_LOG2E : 0.f;
  lse_r[1] = (g + 8 < L) ? lse[g + 8] * AW_LOG2E : 0.f;
  // per-COLUMN statistics of the transposed tile: columns 8 nt + 2 tig + (j & 1)
  float lse_c[2][2], delta_c[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int src = (2 * tig + e) * 4;
      lse_c[nt][e] = __shfl_sync(0xffffffffu, lse_r[nt], src);
      delta_c[nt][e] = __shfl_sync(0xffffffffu, delta[nt], src);
    }
  const float sl2 = p.scale * AW_LOG2E;

  float s[2][4], dp[2][4];
  uint32_t dsa[4], pta[4], dsta[4];
  // ---- plain tile: rows = queries, columns = keys
  mma_RRt(s, q, k);
  mma_RRt(dp, d_o, v);
  {
    float ds[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = 8 * nt + 2 * tig + (j & 1), h = j >> 1, qrow = g + 8 * h;
        const bool kv = (t.valid >> key) & 1u, kin = (t.inlen >> key) & 1u;
        const float pv = aw_prob(s[nt][j] * sl2, lse_r[h], kv, kin, qrow < L, any_valid, inv_L);
        ds[nt][j] = kv ? pv * (dp[nt][j] - delta[h]) * p.scale : 0.f;
      }
    dsa[0] = pack_bf16(ds[0][0], ds[0][1]); dsa[1] = pack_bf16(ds[0][2], ds[0][3]);
    dsa[2] = pack_bf16(ds[1][0], ds[1][1]); dsa[3] = pack_bf16(ds[1][2], ds[1][3]);
  }
  // ---- transposed tile: rows = keys, columns = queries
  mma_RRt(s, k, q);
  mma_RRt(dp, v, d_o);
  {
    float pr[2][4], ds[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int qcol = 8 * nt + 2 * tig + (j & 1), h = j >> 1, key = g + 8 * h;
        const bool kv = (t.valid >> key) & 1u, kin = (t.inlen >> key) & 1u;
        const float pv = aw_prob(s[nt][j] * sl2, lse_c[nt][j & 1], kv, kin, qcol < L, any_valid, inv_L);
        pr[nt][j] = pv;
        ds[nt][j] = kv ? pv * (dp[nt][j] - delta_c[nt][j & 1]) * p.scale : 0.f;
      }
    pta[0] = pack_bf16(pr[0][0], pr[0][1]); pta[1] = pack_bf16(pr[0][2], pr[0][3]);
    pta[2] = pack_bf16(pr[1][0], pr[1][1]); pta[3] = pack_bf16(pr[1][2], pr[1][3]);
    dsta[0] = pack_bf16(ds[0][0], ds[0][1]); dsta[1] = pack_bf16(ds[0][2], ds[0][3]);
    dsta[2] = pack_bf16(ds[1][0], ds[1][1]); dsta[3] = pack_bf16(ds[1][2], ds[1][3]);
  }
  float acc[8][4];
  RowQuad x;
  // dQ = dS . K
  load_P(x, p.k + t.tok0 * p.k_rs + hoff, p.k_rs, p.ts_pos, g, tig, L);
  mma_AP(acc, dsa, x);
  store_acc(p.dq + t.tok0 * p.dq_rs + hoff, p.dq_rs, p.ts_pos, acc, g, tig, L);
  // dV = P^T . dO
  load_P(x, p.d_o + t.tok0 * p.do_rs + hoff, p.do_rs, p.ts_pos, g, tig, L);
  mma_AP(acc, pta, x);
  store_acc(p.dv + t.tok0 * p.dv_rs + hoff, p.dv_rs, p.ts_pos, acc, g, tig, L);
  // dK = dS^T . Q
  load_P(x, p.q + t.tok0 * p.q_rs + hoff, p.q_rs, p.ts_pos, g, tig, L);
  mma_AP(acc, dsta, x);
  store_acc(p.dk + t.tok0 * p.dk_rs + hoff, p.dk_rs, p.ts_pos, acc, g, tig, L);
}

// ------------------------------------------------------------------ 16 < L <= 64: NB = ceil(L/16) blocks of 16 rows
// Forward: one CTA per (sequence, head), warp w owns query block w (scores of all NB key blocks in registers, keys and
// values streamed block by block).  Backward: 3 NB independent warps per (sequence, head): dQ of a query block (streams
// the key blocks), dV of a key block and dK of a key block (stream the query blocks).  S and dP are recomputed by each
// role (tiny MMAs) instead of being reduced across warps; the re-read operands hit L1 (the warps of one sequence are
// neighbours in the same CTA or the next one).
struct TaskCtx64 {
  long long tok0;
  int head;
  unsigned long long valid, inlen;
};

__device__ __forceinline__ TaskCtx64 task_ctx64(const AttnWarpParams& p, long long task, int lane) {
  TaskCtx64 t;
  const long long seq = task / p.heads;
  t.head = (int)(task - seq * p.heads);
  const long long o = seq / p.n_inner, i = seq - o * p.n_inner;
  t.tok0 = o * p.ts_outer + i * p.ts_inner;
  t.inlen = (p.L >= 64) ? ~0ull : ((1ull << p.L) - 1ull);
  bool ok0 = lane < p.L, ok1 = lane + 32 < p.L;
  if (p.mask) {
    const unsigned char* m = p.mask + (seq / p.mask_seq_div) * p.ms_seq;
    if (ok0) ok0 = m[(long long)lane * p.ms_k] != 0;
    if (ok1) ok1 = m[(long long)(lane + 32) * p.ms_k] != 0;
  }
  const unsigned lo = __ballot_sync(0xffffffffu, ok0), hi = __ballot_sync(0xffffffffu, ok1);
  t.valid = (unsigned long long)lo | ((unsigned long long)hi << 32);
  return t;
}

template <int NB>
__global__ void __launch_bounds__(32 * NB) attn_warp_fwd_mb_kernel(const AttnWarpParams p) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
  const long long task = blockIdx.x;
  const TaskCtx64 t = task_ctx64(p, task, lane);
  const int L = p.L, q0 = 16 * w;
  if (q0 >= L) return;
  const long long hoff = t.head * 64;
  const bf16* kb = p.k + t.tok0 * p.k_rs + hoff;
  const bf16* vb = p.v + t.tok0 * p.v_rs + hoff;
  RowPair q;
  load_R(q, p.q + (t.tok0 + (long long)q0 * p.ts_pos) * p.q_rs + hoff, p.q_rs, p.ts_pos, g, tig, L - q0);
  float s[NB][2][4];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    RowPair k;
    load_R(k, kb + (long long)(16 * j) * p.ts_pos * p.k_rs, p.k_rs, p.ts_pos, g, tig, L - 16 * j);
    mma_RRt(s[j], q, k);
  }
  float mx[2] = {-FLT_MAX, -FLT_MAX};
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = 16 * j + 8 * nt + 2 * tig + (e & 1);
        float x = s[j][nt][e] * p.scale;
        if (!((t.valid >> key) & 1ull)) x = AW_MASKED;
        if (!((t.inlen >> key) & 1ull)) x = -FLT_MAX;
        s[j][nt][e] = x;
        mx[e >> 1] = fmaxf(mx[e >> 1], x);
      }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = 16 * j + 8 * nt + 2 * tig + (e & 1);
        float x = aw_exp2((s[j][nt][e] - mx[e >> 1]) * AW_LOG2E);
        if (!((t.inlen >> key) & 1ull)) x = 0.f;
        s[j][nt][e] = x;
        sum[e >> 1] += x;
      }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
  }
  const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    uint32_t pa[4];
    pa[0] = pack_bf16(s[j][0][0] * inv0, s[j][0][1] * inv0);
    pa[1] = pack_bf16(s[j][0][2] * inv1, s[j][0][3] * inv1);
    pa[2] = pack_bf16(s[j][1][0] * inv0, s[j][1][1] * inv0);
    pa[3] = pack_bf16(s[j][1][2] * inv1, s[j][1][3] * inv1);
    RowQuad v;
    load_P(v, vb + (long long)(16 * j) * p.ts_pos * p.v_rs, p.v_rs, p.ts_pos, g, tig, L - 16 * j);
    mma_AP<true>(acc, pa, v);
  }
  store_acc(p.out_o + (t.tok0 + (long long)q0 * p.ts_pos) * p.o_rs + hoff, p.o_rs, p.ts_pos, acc, g, tig, L - q0);
  if (tig == 0) {
    float* lse = p.lse + task * L + q0;
    if (q0 + g < L) lse[g] = mx[0] + __logf(sum[0]);
    if (q0 + g + 8 < L) lse[g + 8] = mx[1] + __logf(sum[1]);
  }
}

// delta = rowsum(dO o O) and lse (log2 units) of the 16 query rows starting at q0: rows g, g+8 per lane
__device__ __forceinline__ void aw_row_stats(const AttnWarpParams& p, long long tok0, long long hoff, const float* lse_seq,
                                             int q0, int L, const RowPair& d_o, int g, int tig, float (&delta)[2],
                                             float (&lse_r)[2]) {
  RowPair o;
  load_R(o, p.o + (tok0 + (long long)q0 * p.ts_pos) * p.o_rs + hoff, p.o_rs, p.ts_pos, g, tig, L - q0);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    d0 += bf_lo(o.lo[i]) * bf_lo(d_o.lo[i]) + bf_hi(o.lo[i]) * bf_hi(d_o.lo[i]);
    d1 += bf_lo(o.hi[i]) * bf_lo(d_o.hi[i]) + bf_hi(o.hi[i]) * bf_hi(d_o.hi[i]);
  }
  d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
  delta[0] = d0; delta[1] = d1;
  lse_r[0] = (q0 + g < L) ? lse_seq[q0 + g] * AW_LOG2E : 0.f;
  lse_r[1] = (q0 + g + 8 < L) ? lse_seq[q0 + g + 8] * AW_LOG2E : 0.f;
}

template <int NB>
__global__ void __launch_bounds__(128, 3) attn_warp_bwd_mb_kernel(const AttnWarpParams p) {
  // independent warps: (task, role) with role < NB: dQ of query block role; < 2 NB: dV of key block; else dK of key block
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const long long gw = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const long long task = gw / (3 * NB);
  if (task >= p.n_tasks) return;
  const int role = (int)(gw - task * (3 * NB));
  const TaskCtx64 t = task_ctx64(p, task, lane);
  const int L = p.L;
  const bool any_valid = t.valid != 0ull;
  const float inv_L = 1.f / (float)L;
  const long long hoff = t.head * 64;
  const float sl2 = p.scale * AW_LOG2E;
  const float* lse_seq = p.lse + task * L;
  const bf16* qb = p.q + t.tok0 * p.q_rs + hoff;
  const bf16* kb = p.k + t.tok0 * p.k_rs + hoff;
  const bf16* vb = p.v + t.tok0 * p.v_rs + hoff;
  const bf16* gb = p.d_o + t.tok0 * p.do_rs + hoff;
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  if (role < NB) {
    // ---------------- dQ of query block `role`: streams the key blocks
    const int q0 = 16 * role;
    if (q0 >= L) return;
    RowPair q, d_o;
    load_R(q, qb + (long long)q0 * p.ts_pos * p.q_rs, p.q_rs, p.ts_pos, g, tig, L - q0);
    load_R(d_o, gb + (long long)q0 * p.ts_pos * p.do_rs, p.do_rs, p.ts_pos, g, tig, L - q0);
    float delta[2], lse_r[2];
    aw_row_stats(p, t.tok0, hoff, lse_seq, q0, L, d_o, g, tig, delta, lse_r);
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
      const int k0 = 16 * j;
      if (k0 >= L) break;
      RowPair k, v;
      load_R(k, kb + (long long)k0 * p.ts_pos * p.k_rs, p.k_rs, p.ts_pos, g, tig, L - k0);
      load_R(v, vb + (long long)k0 * p.ts_pos * p.v_rs, p.v_rs, p.ts_pos, g, tig, L - k0);
      float s[2][4], dp[2][4], ds[2][4];
      mma_RRt(s, q, k);
      mma_RRt(dp, d_o, v);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = k0 + 8 * nt + 2 * tig + (e & 1), h = e >> 1, qrow = q0 + g + 8 * h;
          const bool kv = (t.valid >> key) & 1ull, kin = (t.inlen >> key) & 1ull;
          const float pv = aw_prob(s[nt][e] * sl2, lse_r[h], kv, kin, qrow < L, any_valid, inv_L);
          ds[nt][e] = kv ? pv * (dp[nt][e] - delta[h]) * p.scale : 0.f;
        }
      uint32_t dsa[4];
      dsa[0] = pack_bf16(ds[0][0], ds[0][1]); dsa[1] = pack_bf16(ds[0][2], ds[0][3]);
      dsa[2] = pack_bf16(ds[1][0], ds[1][1]); dsa[3] = pack_bf16(ds[1][2], ds[1][3]);
      RowQuad x;
      load_P(x, kb + (long long)k0 * p.ts_pos * p.k_rs, p.k_rs, p.ts_pos, g, tig, L - k0);
      mma_AP<true>(acc, dsa, x);
    }
    store_acc(p.dq + (t.tok0 + (long long)q0 * p.ts_pos) * p.dq_rs + hoff, p.dq_rs, p.ts_pos, acc, g, tig, L - q0);
    return;
  }
  // ---------------- dV (want_dk = false) or dK (want_dk = true) of one key block: streams the query blocks
  const bool want_dk = role >= 2 * NB;
  const int k0 = 16 * (role - (want_dk ? 2 * NB : NB));
  if (k0 >= L) return;
  RowPair k, v;
  load_R(k, kb + (long long)k0 * p.ts_pos * p.k_rs, p.k_rs, p.ts_pos, g, tig, L - k0);
  if (want_dk) load_R(v, vb + (long long)k0 * p.ts_pos * p.v_rs, p.v_rs, p.ts_pos, g, tig, L - k0);
#pragma unroll 1
  for (int i = 0; i < NB; ++i) {
    const int q0 = 16 * i;
    if (q0 >= L) break;
    // per-COLUMN (query) statistics of the transposed tile: columns q0 + 8 nt + 2 tig + e
    float lse_c[2][2], delta_c[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int qc = q0 + 8 * nt + 2 * tig + e;
        lse_c[nt][e] = qc < L ? lse_seq[qc] * AW_LOG2E : 0.f;
        delta_c[nt][e] = 0.f;
      }
    RowPair q;
    load_R(q, qb + (long long)q0 * p.ts_pos * p.q_rs, p.q_rs, p.ts_pos, g, tig, L - q0);
    float s[2][4], dp[2][4];
    mma_RRt(s, k, q);
    if (want_dk) {
      RowPair d_o;
      load_R(d_o, gb + (long long)q0 * p.ts_pos * p.do_rs, p.do_rs, p.ts_pos, g, tig, L - q0);
      float delta[2], lse_r[2];
      aw_row_stats(p, t.tok0, hoff, lse_seq, q0, L, d_o, g, tig, delta, lse_r);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) delta_c[nt][e] = __shfl_sync(0xffffffffu, delta[nt], (2 * tig + e) * 4);
      mma_RRt(dp, v, d_o);
    }
    float o[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qcol = q0 + 8 * nt + 2 * tig + (e & 1), h = e >> 1, key = k0 + g + 8 * h;
        const bool kv = (t.valid >> key) & 1ull, kin = (t.inlen >> key) & 1ull;
        const float pv = aw_prob(s[nt][e] * sl2, lse_c[nt][e & 1], kv, kin, qcol < L, any_valid, inv_L);
        o[nt][e] = want_dk ? (kv ? pv * (dp[nt][e] - delta_c[nt][e & 1]) * p.scale : 0.f) : pv;
      }
    uint32_t fa[4];
    fa[0] = pack_bf16(o[0][0], o[0][1]); fa[1] = pack_bf16(o[0][2], o[0][3]);
    fa[2] = pack_bf16(o[1][0], o[1][1]); fa[3] = pack_bf16(o[1][2], o[1][3]);
    RowQuad x;
    if (want_dk) load_P(x, qb + (long long)q0 * p.ts_pos * p.q_rs, p.q_rs, p.ts_pos, g, tig, L - q0);   // dK += dS^T . Q_i
    else load_P(x, gb + (long long)q0 * p.ts_pos * p.do_rs, p.do_rs, p.ts_pos, g, tig, L - q0);         // dV += P^T . dO_i
    mma_AP<true>(acc, fa, x);
  }
  if (want_dk) store_acc(p.dk + (t.tok0 + (long long)k0 * p.ts_pos) * p.dk_rs + hoff, p.dk_rs, p.ts_pos, acc, g, tig, L - k0);
  else store_acc(p.dv + (t.tok0 + (long long)k0 * p.ts_pos) * p.dv_rs + hoff, p.dv_rs, p.ts_pos, acc, g, tig, L - k0);
}

// ------------------------------------------------------------------ host
static bool aw_fill(const vvae_attn_args& a, AttnWarpParams& p, bool bwd) {
  if (a.dtype != VVAE_BF16 || a.hd != 64 || a.L < 1 || a.L > 64) return false;
  if (a.mask && (a.ms_head != 0 || a.ms_q != 0)) return false;       // key-padding masks only
  auto al16 = [](const void* x) { return ((uintptr_t)x % 16) == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(a.o)) return false;
  if ((a.q_rs % 8) || (a.k_rs % 8) || (a.v_rs % 8) || (a.o_rs % 8)) return false;
  if (bwd) {
    if (!al16(a.d_o) || !al16(a.dq) || !al16(a.dk) || !al16(a.dv)) return false;
    if ((a.do_rs % 8) || (a.dq_rs % 8) || (a.dk_rs % 8) || (a.dv_rs % 8)) return false;
  }
  const long long n_tasks = (long long)a.n_outer * a.n_inner * a.heads;
  if (n_tasks <= 0 || n_tasks * 3 > 0x7fffffffLL) return false;
  p.q = (const bf16*)a.q; p.k = (const bf16*)a.k; p.v = (const bf16*)a.v;
  p.q_rs = a.q_rs; p.k_rs = a.k_rs; p.v_rs = a.v_rs; p.o_rs = a.o_rs;
  p.o = (const bf16*)a.o; p.out_o = (bf16*)a.o;
  p.d_o = (const bf16*)a.d_o; p.do_rs = a.do_rs;
  p.dq = (bf16*)a.dq; p.dk = (bf16*)a.dk; p.dv = (bf16*)a.dv;
  p.dq_rs = a.dq_rs; p.dk_rs = a.dk_rs; p.dv_rs = a.dv_rs;
  p.ts_outer = a.tok_stride_outer; p.ts_inner = a.tok_stride_inner; p.ts_pos = a.tok_stride_pos;
  p.lse = a.lse;
  p.mask = a.mask; p.mask_seq_div = a.mask_seq_div > 0 ? a.mask_seq_div : 1; p.ms_seq = a.ms_seq; p.ms_k = a.ms_k;
  p.n_inner = a.n_inner; p.L = a.L; p.heads = a.heads; p.n_tasks = n_tasks;
  p.scale = a.scale;
  return true;
}

int attn_warp_supported(const vvae_attn_args& a, bool bwd) {
  AttnWarpParams p;
  return aw_fill(a, p, bwd) ? 1 : 0;
}

int attn_warp_fwd(const vvae_attn_args& a, cudaStream_t s) {
  AttnWarpParams p;
  if (!aw_fill(a, p, false)) {
    set_error("attention: shape not supported by the short-sequence kernel");
    return VVAE_ERR_UNSUPPORTED;
  }
  const int nb = (p.L + 15) / 16;
  if (nb == 1) attn_warp_fwd_kernel<<<(unsigned)cdiv(p.n_tasks, 4), 128, 0, s>>>(p);
  else if (nb == 2) attn_warp_fwd_mb_kernel<2><<<(unsigned)p.n_tasks, 64, 0, s>>>(p);
  else if (nb == 3) attn_warp_fwd_mb_kernel<3><<<(unsigned)p.n_tasks, 96, 0, s>>>(p);
  else attn_warp_fwd_mb_kernel<4><<<(unsigned)p.n_tasks, 128, 0, s>>>(p);
  return check_launch("attn_warp_fwd");
}

int attn_warp_bwd(const vvae_attn_args& a, cudaStream_t s) {
  AttnWarpParams p;
  if (!aw_fill(a, p, true)) {
    set_error("attention bwd: shape not supported by the short-sequence kernel");
    return VVAE_ERR_UNSUPPORTED;
  }
  const int nb = (p.L + 15) / 16;
  if (nb == 1) attn_warp_bwd_kernel<<<(unsigned)cdiv(p.n_tasks, 4), 128, 0, s>>>(p);
  else if (nb == 2) attn_warp_bwd_mb_kernel<2><<<(unsigned)cdiv(p.n_tasks * 6, 4), 128, 0, s>>>(p);
  else if (nb == 3) attn_warp_bwd_mb_kernel<3><<<(unsigned)cdiv(p.n_tasks * 9, 4), 128, 0, s>>>(p);
  else attn_warp_bwd_mb_kernel<4><<<(unsigned)cdiv(p.n_tasks * 12, 4), 128, 0, s>>>(p);
  return check_launch("attn_warp_bwd");
}

}  // namespace vvae
