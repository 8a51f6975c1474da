// Attention on tcgen05 (bf16, head_dim 64): forward.
//
// One CTA = one tile of 128 query rows for one head.  A tile is either
//   * 128 consecutive positions of ONE sequence of length L = 128 or 256 (spatial attention, NK = L keys), or
//   * G = 128/L whole sequences of length L in {8,16,32,64} packed block-diagonally (temporal attention and small
//     spatial grids): the TMA box [64 dims][L positions][G sequences] lands the G sequences one after the other, the
//     tensor core computes the full 128x128 score tile and the softmax only looks at the L columns of its own sequence
//     (the other columns get P = 0).  Strided temporal sequences (tokens hw apart) are gathered by the TMA tensor map,
//     nothing is transposed in memory.
// Pipeline per CTA: TMA(Q,K | V) -> S = Q.K^T (tcgen05, accumulators in TMEM) -> softmax by 128 threads, one query row
// per thread straight out of TMEM (key-padding mask and the block-diagonal rule applied in-tile, masked logits =
// -0.7*FLT_MAX as in jax.nn.dot_product_attention) -> P (bf16) written into shared memory in the UMMA K-major
// 128B-swizzled layout, over the dead Q/K tiles -> O = P.V (tcgen05; V is consumed MN-major as it lies) ->
// O / rowsum and the log-sum-exp are stored.  Two CTAs share an SM (<= 97 KB smem, <= 256 TMEM columns each) so one
// CTA's softmax overlaps the other's MMAs and loads.
#include <atomic>
#include <cuda.h>
#include <cfloat>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

namespace vvae {

#define ATC_BIG_NEG (-0.7f * FLT_MAX)

struct AttnTcPlan {
  int L, heads, G, NK, nqb, pack_inner, n_outer, n_inner;
  long long ts_o, ts_i, ts_l;
  long long tiles;
  int tiles_per_outer;
};

struct AttnTcParams {
  AttnTcPlan pl;
  bf16* o; long long o_rs;
  float* lse;
  const unsigned char* mask; long long mask_seq_div, ms_seq, ms_k;
  float scale;
  int dbg, dbg_cta;          // vvae_debug_set(10, 16): CTA vvae_debug_set(0, n) records a clock64 timeline (vvae_debug_get(1, .))
};
// clock64 stamps of one CTA of the forward / backward tcgen05 kernels (see the stamp sites); [31] = SM id
__device__ unsigned long long g_attn_dbg[32];

__host__ __device__ inline bool atc_len_long(int L) { return L > 256 && L % 128 == 0 && L <= 8192; }
__host__ __device__ inline bool atc_len_ok(int L) {
  return L == 8 || L == 16 || L == 32 || L == 64 || L == 128 || L == 256 || atc_len_long(L);
}

static bool atc_make_plan(const vvae_attn_args& a, AttnTcPlan& p) {
  if (a.dtype != VVAE_BF16 || a.hd != 64 || !atc_len_ok(a.L)) return false;
  if (a.mask && (a.ms_head != 0 || a.ms_q != 0)) return false;       // key-padding masks only
  p.L = a.L; p.heads = a.heads; p.n_outer = a.n_outer; p.n_inner = a.n_inner;
  p.ts_o = a.tok_stride_outer; p.ts_i = a.tok_stride_inner; p.ts_l = a.tok_stride_pos;
  if (a.L >= 128) {
    p.G = 1; p.NK = a.L; p.nqb = a.L / 128; p.pack_inner = 0; p.tiles_per_outer = 0;
    p.tiles = (long long)a.n_outer * a.n_inner * p.nqb;
  } else {
    p.G = 128 / a.L; p.NK = 128; p.nqb = 1;
    p.pack_inner = a.n_inner > 1 ? 1 : 0;
    if (p.pack_inner) {
      p.tiles_per_outer = (a.n_inner + p.G - 1) / p.G;
      p.tiles = (long long)a.n_outer * p.tiles_per_outer;
    } else {
      p.tiles_per_outer = 0;
      p.tiles = (a.n_outer + p.G - 1) / p.G;
    }
  }
  return p.tiles > 0 && p.tiles < 0x7fffffffLL;
}

// 4-D map (dims, position, inner sequence index, outer sequence index) over a token-major tensor X[token*rs + col]
static int atc_make_map(CUtensorMap* m, const void* base, long long rs, const AttnTcPlan& p, int box_l, int box_i, int box_o) {
  const uint64_t rb = (uint64_t)rs * 2;
  uint64_t dims[4] = {(uint64_t)p.heads * 64, (uint64_t)p.L, (uint64_t)p.n_inner, (uint64_t)p.n_outer};
  uint64_t str[3] = {(uint64_t)p.ts_l * rb, (uint64_t)p.ts_i * rb, (uint64_t)p.ts_o * rb};
  if (p.n_inner == 1 || str[1] == 0) str[1] = str[0] * (uint64_t)p.L;   // extent-1 dimension: any legal stride
  if (p.n_outer == 1 || str[2] == 0) str[2] = str[1] * (uint64_t)p.n_inner;
  uint32_t box[4] = {64, (uint32_t)box_l, (uint32_t)box_i, (uint32_t)box_o};
  return encode_tmap_nd_bf16(m, base, 4, dims, str, box, 128);
}

__device__ __forceinline__ float atc_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(sm100::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(sm100::smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// row r of a tile -> (sequence index, position, token index); returns false for rows beyond the tensor
struct AtcRow { long long seq, tok; int l; bool ok; };
template <bool PACKED>
__device__ __forceinline__ AtcRow atc_row(const AttnTcPlan& p, long long tile, int r) {
  AtcRow o;
  if (PACKED) {
    const int si = r / p.L;
    o.l = r - si * p.L;
    long long outer, inner;
    if (p.pack_inner) {
      outer = tile / p.tiles_per_outer;
      inner = (tile % p.tiles_per_outer) * p.G + si;
      o.ok = inner < p.n_inner;
    } else {
      outer = tile * p.G + si;
      inner = 0;
      o.ok = outer < p.n_outer;
    }
    o.seq = outer * p.n_inner + inner;
    o.tok = outer * p.ts_o + inner * p.ts_i + (long long)o.l * p.ts_l;
  } else {
    const long long seq = tile / p.nqb;
    const int qb = (int)(tile % p.nqb);
    o.l = qb * 128 + r;
    o.seq = seq;
    o.ok = true;
    o.tok = (seq / p.n_inner) * p.ts_o + (seq % p.n_inner) * p.ts_i + (long long)o.l * p.ts_l;
  }
  return o;
}

template <int NK, bool PACKED, bool MASKED>
__global__ void __launch_bounds__(192, 2)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, const AttnTcParams q) {
  const AttnTcPlan& p = q.pl;
  constexpr int P_BYTES = (NK / 64) * 16384;                 // P: NK/64 K-blocks of [128 rows x 128 B]
  constexpr int QK_BYTES = 16384 + NK * 128;
  constexpr int R0_BYTES = P_BYTES > QK_BYTES ? P_BYTES : QK_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // [128][64] K-major SW128
  uint8_t* sK = smem + 16384;         // [NK][64]  K-major SW128
  uint8_t* sP = smem;                 // overlays Q|K once S is complete
  uint8_t* sV = smem + R0_BYTES;      // [NK][64]  (MN-major B operand of P.V)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NK * 128);
  uint64_t* qk_full = bars;
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* kpen = reinterpret_cast<float*>(bars + 6);   // [NK] 0 = attend, 1 = masked key

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tile = blockIdx.x;
  const int h = blockIdx.y;
  // forward timeline slots: 0 start, 1 after griddepcontrol.wait, 2 Q|K landed (issuer), 3 P ready (issuer), 4 PV issued,
  // 5 softmax warp 2: S ready, 6 row max done, 7 P written, 8 O ready, 9 O stored
  const bool fdbg = q.dbg && (int)(blockIdx.y * gridDim.x + blockIdx.x) == q.dbg_cta && lane == 0;
  if (fdbg && warp == 0) g_attn_dbg[0] = clock64();

  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_q);
    sm100::tma_prefetch_desc(&tma_k);
    sm100::tma_prefetch_desc(&tma_v);
    sm100::mbar_init(qk_full, 1);
    sm100::mbar_init(v_full, 1);
    sm100::mbar_init(s_full, 1);
    sm100::mbar_init(p_full, 4);
    sm100::mbar_init(o_full, 1);
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<NK>(tmem_slot);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // PDL (common.cuh): barrier init / TMEM allocation above overlap the previous kernel's tail
  if (fdbg && warp == 0) g_attn_dbg[1] = clock64();

  if (warp == 0) {
    if (lane == 0) {
      int c1q, c1k, c2, c3;
      if (PACKED) {
        c1q = c1k = 0;
        if (p.pack_inner) { c3 = (int)(tile / p.tiles_per_outer); c2 = (int)(tile % p.tiles_per_outer) * p.G; }
        else { c3 = (int)tile * p.G; c2 = 0; }
      } else {
        const long long seq = tile / p.nqb;
        c1q = (int)(tile % p.nqb) * 128; c1k = 0;
        c3 = (int)(seq / p.n_inner); c2 = (int)(seq % p.n_inner);
      }
      sm100::mbar_expect_tx(qk_full, 16384 + NK * 128);
      tma_load_4d(sQ, &tma_q, qk_full, h * 64, c1q, c2, c3);
      tma_load_4d(sK, &tma_k, qk_full, h * 64, c1k, c2, c3);
      sm100::mbar_expect_tx(v_full, NK * 128);
      tma_load_4d(sV, &tma_v, v_full, h * 64, c1k, c2, c3);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // S[128 x NK] = Q . K^T : both operands K-major (head dim contiguous), 4 K-steps of 16
      constexpr uint32_t idesc_s = sm100::make_idesc_bf16(128, NK, false, false);
      sm100::mbar_wait(qk_full, 0);
      sm100::tc_fence_after();
      if (fdbg) g_attn_dbg[2] = clock64();
      const uint32_t qa = sm100::smem_u32(sQ), ka = sm100::smem_u32(sK);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sm100::umma_f16(tmem_base, sm100::make_smem_desc_sw128(qa + k * 32, 16, 1024),
                        sm100::make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
      sm100::umma_commit(s_full);
      // O[128 x 64] = P . V : P K-major (keys contiguous, 64-key blocks 16 KB apart), V MN-major (head dim contiguous)
      constexpr uint32_t idesc_o = sm100::make_idesc_bf16(128, 64, false, true);
      sm100::mbar_wait(v_full, 0);
      sm100::mbar_wait(p_full, 0);
      sm100::tc_fence_after();
      if (fdbg) g_attn_dbg[3] = clock64();
      const uint32_t pa = sm100::smem_u32(sP), va = sm100::smem_u32(sV);
#pragma unroll
      for (int k = 0; k < NK / 16; ++k)
        sm100::umma_f16(tmem_base, sm100::make_smem_desc_sw128(pa + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                        sm100::make_smem_desc_sw128(va + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
      sm100::umma_commit(o_full);
      if (fdbg) g_attn_dbg[4] = clock64();
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // query row == TMEM lane
    const int tid = threadIdx.x - 64;                  // 0..127
    const bool sdbg = fdbg && warp == 2;
    // key penalties for the tile's NK key rows
    for (int c = tid; c < NK; c += 128) {
      float pen = 0.f;
      if (q.mask) {
        AtcRow kr = atc_row<PACKED>(p, PACKED ? tile : (tile / p.nqb) * p.nqb, c);   // unpacked: key c is position c
        if (kr.ok && q.mask[(kr.seq / q.mask_seq_div) * q.ms_seq + (long long)kr.l * q.ms_k] == 0) pen = 1.f;
      }
      kpen[c] = pen;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const AtcRow row = atc_row<PACKED>(p, tile, r);
    const int seq_c0 = PACKED ? (r / p.L) * p.L : 0;             // first key column of this row's sequence
    int cbeg = 0, cend = NK;
    if (PACKED) {
      if (p.L >= 32) { cbeg = seq_c0; cend = seq_c0 + p.L; }
      else { cbeg = quarter * 32; cend = cbeg + 32; }
    }
    const float k2 = q.scale * 1.4426950408889634f;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);

    sm100::mbar_wait(s_full, 0);
    sm100::tc_fence_after();
    if (sdbg) g_attn_dbg[5] = clock64();
    // pass 1: row maximum of the (masked) logits
    float mx = -INFINITY;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t sr[32];
      sm100::tmem_ld_32x32(trow + c0, sr);
      sm100::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int c = c0 + i;
        const bool inseq = !PACKED || (unsigned)(c - seq_c0) < (unsigned)p.L;
        float s = __uint_as_float(sr[i]) * q.scale;
        if constexpr (MASKED) { if (kpen[c] != 0.f) s = ATC_BIG_NEG; }
        if (inseq) mx = fmaxf(mx, s);
      }
    }
    if (sdbg) g_attn_dbg[6] = clock64();
    // pass 2: p = exp(s - max), row sum, bf16 P into the swizzled K-major operand tile
    const float mx2 = mx * 1.4426950408889634f;
    float sum = 0.f;
    uint8_t* prow = sP + r * 128;
    const uint32_t sw = (uint32_t)(r & 7);
    for (int c0 = 0; c0 < NK; c0 += 32) {
      uint8_t* pblk = prow + (c0 >> 6) * 16384;
      const uint32_t ch0 = (uint32_t)(c0 & 63) >> 3;     // first 16-byte chunk of this 32-key group inside the 128 B row
      if (c0 < cbeg || c0 >= cend) {                      // warp-uniform
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(pblk + (((ch0 + j) ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
        continue;
      }
      uint32_t sr[32];
      sm100::tmem_ld_32x32(trow + c0, sr);
      sm100::tmem_ld_wait();
      float pv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int c = c0 + i;
        const bool inseq = !PACKED || (unsigned)(c - seq_c0) < (unsigned)p.L;
        float e = atc_exp2(__uint_as_float(sr[i]) * k2 - mx2);
        if constexpr (MASKED) { if (kpen[c] != 0.f) e = (mx == ATC_BIG_NEG ? 1.f : 0.f); }
        pv[i] = inseq ? e : 0.f;
        sum += pv[i];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(pv[8 * j + 0], pv[8 * j + 1]);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(pv[8 * j + 2], pv[8 * j + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(pv[8 * j + 4], pv[8 * j + 5]);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(pv[8 * j + 6], pv[8 * j + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(pblk + (((ch0 + j) ^ sw) << 4)) = pk;
      }
    }
    sm100::fence_proxy_async();      // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
    sm100::tc_fence_before();
    __syncwarp();
    if (lane == 0) sm100::mbar_arrive(p_full);
    if (sdbg) g_attn_dbg[7] = clock64();

    // epilogue: O / rowsum -> bf16, log-sum-exp
    sm100::mbar_wait(o_full, 0);
    sm100::tc_fence_after();
    if (sdbg) g_attn_dbg[8] = clock64();
    const float inv = 1.f / sum;
    bf16* orow = q.o + row.tok * q.o_rs + (long long)h * 64;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t orr[32];
      sm100::tmem_ld_32x32(trow + c0, orr);
      sm100::tmem_ld_wait();
      if (row.ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          Vec16<bf16> v;
#pragma unroll
          for (int t = 0; t < 8; ++t) v.set(t, __uint_as_float(orr[8 * j + t]) * inv);
          v.store(orow + c0 + 8 * j);
        }
      }
    }
    if (row.ok && q.lse) q.lse[(row.seq * p.heads + h) * p.L + row.l] = mx + __logf(sum);
    if (sdbg) g_attn_dbg[9] = clock64();
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (fdbg && warp == 0) {
    g_attn_dbg[10] = clock64();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_attn_dbg[31] = smid;
  }
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<NK>(tmem_base);
  }
}

template <int NK, bool PACKED, bool MASKED>
static int atc_launch_fwd(const vvae_attn_args& a, const AttnTcPlan& p, cudaStream_t s) {
  constexpr int P_BYTES = (NK / 64) * 16384, QK_BYTES = 16384 + NK * 128;
  constexpr int R0 = P_BYTES > QK_BYTES ? P_BYTES : QK_BYTES;
  constexpr int SMEM = R0 + NK * 128 + 64 + NK * 4 + 1024;
  CUtensorMap mq, mk, mv;
  int rc;
  if (PACKED) {
    const int bi = p.pack_inner ? p.G : 1, bo = p.pack_inner ? 1 : p.G;
    if ((rc = atc_make_map(&mq, a.q, a.q_rs, p, p.L, bi, bo))) return rc;
    if ((rc = atc_make_map(&mk, a.k, a.k_rs, p, p.L, bi, bo))) return rc;
    if ((rc = atc_make_map(&mv, a.v, a.v_rs, p, p.L, bi, bo))) return rc;
  } else {
    if ((rc = atc_make_map(&mq, a.q, a.q_rs, p, 128, 1, 1))) return rc;
    if ((rc = atc_make_map(&mk, a.k, a.k_rs, p, NK, 1, 1))) return rc;
    if ((rc = atc_make_map(&mv, a.v, a.v_rs, p, NK, 1, 1))) return rc;
  }
  AttnTcParams q;
  q.pl = p;
  q.o = (bf16*)a.o; q.o_rs = a.o_rs; q.lse = a.lse;
  q.mask = a.mask; q.mask_seq_div = a.mask_seq_div > 0 ? a.mask_seq_div : 1; q.ms_seq = a.ms_seq; q.ms_k = a.ms_k;
  q.scale = a.scale;
  q.dbg = (g_dbg[10] & 16) ? 1 : 0;
  q.dbg_cta = (int)g_dbg[0];
  auto kern = attn_fwd_sm100_kernel<NK, PACKED, MASKED>;
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {   // idempotent: a racing second call sets the same value
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  dim3 grid((unsigned)p.tiles, (unsigned)p.heads);
  launch_pdl(kern, grid, dim3(192), SMEM, s, mq, mk, mv, q);
  return check_launch("attn_fwd_sm100");
}


// ====================================================================================================== backward
// One CTA = one tile set of one head: NB x NB tiles of 128 queries x 128 keys (NB = 1: one 128-row tile, i.e. a
// sequence of 128 or a block-diagonal pack of short sequences; NB = 2: a sequence of 256).  Everything a tile needs
// (Q, K, V, dO: NB x 16 KB each) is TMA-staged once; the five contractions of the attention backward run on tcgen05
// with all accumulators resident in TMEM (512 columns):
//     S  = Q_i K_j^T          [  0,128)        dP = dO_i V_j^T       [128,256)
//     dQ_i += dS K_j          [256+64i, +64)   dK_j += dS^T Q_i      [256+64NB, +64)     dV_j += P^T dO_i   [.. +64, +64)
// 8 softmax warps (warp group g owns key columns [64g, 64g+64) of the tile) read S and dP out of TMEM, rebuild
// P = exp(S*scale - LSE), form dS = P o (dP - D) * scale with D = rowsum(dO o O), and write bf16 P and dS into shared
// memory in ONE 128B-swizzled layout that the tensor core reads both K-major (dQ = dS.K) and MN-major (dV = P^T.dO,
// dK = dS^T.Q), so no transposed copy exists anywhere.  The MMA warp issues S/dP of tile t+1 before the three output
// contractions of tile t, so the softmax of t+1 overlaps them.  Mask semantics follow the forward: masked logits are
// the constant -0.7*FLT_MAX, so they receive no gradient; a fully masked row has uniform P (dV only).
struct AttnTcBwdParams {
  AttnTcPlan pl;
  const bf16* o; long long o_rs;
  const float* lse;
  const unsigned char* mask; long long mask_seq_div, ms_seq, ms_k;
  bf16 *dq, *dk, *dv; long long dq_rs, dk_rs, dv_rs;
  float scale;
  int dbg;     // vvae_debug_set(10, 16): the CTA given by vvae_debug_set(0, n) records a clock64 timeline (vvae_debug_get(1, .))
  int dbg_cta;
};

#define ATB_STAMP(slot) do { if (dbg_on) g_attn_dbg[slot] = clock64(); } while (0)

// row r of 128-row block `blk` of the CTA's tile set
template <bool PACKED>
__device__ __forceinline__ AtcRow atc_bwd_row(const AttnTcPlan& p, long long tile, int blk, int r) {
  return atc_row<PACKED>(p, PACKED ? tile : tile * p.nqb + blk, r);
}

__device__ __forceinline__ uint32_t atc_pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int NB, bool PACKED, bool MASKED>
__global__ void __launch_bounds__(320, 1)
attn_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                      const AttnTcBwdParams q) {
  const AttnTcPlan& p = q.pl;
  constexpr int BLK = 16384;                       // one [128 rows x 64 bf16] swizzled tile
  constexpr int NT = NB * NB;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 256 + 64 * NB, COL_DV = COL_DK + 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NB * BLK;
  uint8_t* sV = sK + NB * BLK;
  uint8_t* sdO = sV + NB * BLK;
  uint8_t* sP = sdO + NB * BLK;                    // [2 key halves][128 q][64 keys] bf16, swizzled
  uint8_t* sdS = sP + 2 * BLK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * BLK);
  uint64_t* ld_bar = bars;                         // [NB] block b of Q, K, V, dO has landed
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* mma3_done = bars + 4;
  uint64_t* kv_drained = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* kpen = reinterpret_cast<float*>(bars + 7);   // [NB*128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tile = blockIdx.x;
  const int h = blockIdx.y;
  const bool dbg_on = q.dbg && (int)(blockIdx.y * gridDim.x + blockIdx.x) == q.dbg_cta && lane == 0;
  if (warp == 0) ATB_STAMP(0);

  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_q);
    sm100::tma_prefetch_desc(&tma_k);
    sm100::tma_prefetch_desc(&tma_v);
    sm100::tma_prefetch_desc(&tma_do);
    for (int b = 0; b < NB; ++b) sm100::mbar_init(&ld_bar[b], 1);
    sm100::mbar_init(s_full, 1);
    sm100::mbar_init(p_full, 8);
    sm100::mbar_init(mma3_done, 1);
    sm100::mbar_init(kv_drained, 8);
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<512>(tmem_slot);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) ATB_STAMP(1);
  pdl_wait();   // PDL (common.cuh)
  if (warp == 0) ATB_STAMP(2);

  if (warp == 0) {
    if (lane == 0) {
      int c2, c3;
      if (PACKED) {
        if (p.pack_inner) { c3 = (int)(tile / p.tiles_per_outer); c2 = (int)(tile % p.tiles_per_outer) * p.G; }
        else { c3 = (int)tile * p.G; c2 = 0; }
      } else {
        c3 = (int)(tile / p.n_inner); c2 = (int)(tile % p.n_inner);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        sm100::mbar_expect_tx(&ld_bar[b], 4 * BLK);
        tma_load_4d(sQ + b * BLK, &tma_q, &ld_bar[b], h * 64, b * 128, c2, c3);
        tma_load_4d(sK + b * BLK, &tma_k, &ld_bar[b], h * 64, b * 128, c2, c3);
        tma_load_4d(sV + b * BLK, &tma_v, &ld_bar[b], h * 64, b * 128, c2, c3);
        tma_load_4d(sdO + b * BLK, &tma_do, &ld_bar[b], h * 64, b * 128, c2, c3);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = sm100::make_idesc_bf16(128, 128, false, false);   // S, dP: both K-major
      constexpr uint32_t idesc_kv = sm100::make_idesc_bf16(128, 64, true, true);     // dV, dK: A^T (MN-major), B MN-major
      constexpr uint32_t idesc_q = sm100::make_idesc_bf16(128, 64, false, true);     // dQ: A K-major, B MN-major
      const uint32_t aQ = sm100::smem_u32(sQ), aK = sm100::smem_u32(sK), aV = sm100::smem_u32(sV);
      const uint32_t aO = sm100::smem_u32(sdO), aP = sm100::smem_u32(sP), aS = sm100::smem_u32(sdS);
      auto issue_sdp = [&](int t) {
        const int j = t / NB, i = t % NB;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + COL_S, sm100::make_smem_desc_sw128(aQ + i * BLK + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(aK + j * BLK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + COL_DP, sm100::make_smem_desc_sw128(aO + i * BLK + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(aV + j * BLK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        sm100::umma_commit(s_full);
      };
      sm100::mbar_wait(&ld_bar[0], 0);
      sm100::tc_fence_after();
      ATB_STAMP(3);
      issue_sdp(0);
      int loaded = 1;
#pragma unroll 1
      for (int t = 0; t < NT; ++t) {
        const int j = t / NB, i = t % NB;
        sm100::mbar_wait(p_full, t & 1);
        sm100::tc_fence_after();
        ATB_STAMP(4 + (t & 3));
        if (t + 1 < NT) {
          const int need = max((t + 1) / NB, (t + 1) % NB) + 1;
          while (loaded < need) { sm100::mbar_wait(&ld_bar[loaded], 0); ++loaded; }
          sm100::tc_fence_after();
          issue_sdp(t + 1);
        }
        if (i == 0 && j > 0) {            // dK/dV of the previous key block must have been read out
          sm100::mbar_wait(kv_drained, (j - 1) & 1);
          sm100::tc_fence_after();
        }
        // dV_j (+)= P^T dO_i ; dK_j (+)= dS^T Q_i : contraction over the 128 queries, 16 per step
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sm100::umma_f16(tmem_base + COL_DV, sm100::make_smem_desc_sw128(aP + k * 2048, BLK, 1024),
                          sm100::make_smem_desc_sw128(aO + i * BLK + k * 2048, 8192, 1024), idesc_kv,
                          (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sm100::umma_f16(tmem_base + COL_DK, sm100::make_smem_desc_sw128(aS + k * 2048, BLK, 1024),
                          sm100::make_smem_desc_sw128(aQ + i * BLK + k * 2048, 8192, 1024), idesc_kv,
                          (i > 0 || k > 0) ? 1u : 0u);
        // dQ_i (+)= dS K_j : contraction over the 128 keys
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sm100::umma_f16(tmem_base + COL_DQ + 64 * i,
                          sm100::make_smem_desc_sw128(aS + (k >> 2) * BLK + (k & 3) * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(aK + j * BLK + k * 2048, 8192, 1024), idesc_q,
                          (j > 0 || k > 0) ? 1u : 0u);
        sm100::umma_commit(mma3_done);
        ATB_STAMP(8 + (t & 3));
      }
    }
  } else {
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
    const int g = (warp - 2) >> 2;                     // key-column half of the tile
    const bool dbg_on2 = dbg_on && warp == 2;
    const int r = quarter * 32 + lane;                 // row of the 128-row block == TMEM lane
    const int tid = threadIdx.x - 64;                  // 0..255
    for (int c = tid; c < NB * 128; c += 256) {
      float pen = 0.f;
      if (q.mask) {
        AtcRow kr = atc_bwd_row<PACKED>(p, tile, c >> 7, c & 127);
        if (kr.ok && q.mask[(kr.seq / q.mask_seq_div) * q.ms_seq + (long long)kr.l * q.ms_k] == 0) pen = 1.f;
      }
      kpen[c] = pen;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t sw = (uint32_t)(r & 7);
    const int seq_c0 = PACKED ? (r / p.L) * p.L : 0;
    const float k2 = q.scale * 1.4426950408889634f;
    const float inv_l = 1.f / (float)p.L;

    // per query block: row validity, LSE (in log2 units), D = rowsum(dO o O).  The O rows come straight from global
    // memory: their loads are issued before the wait for the TMA tiles so both latencies overlap.
    float lse2[NB], dlt[NB];
    bool rok[NB], allm[NB];
    uint4 ov[NB][8];
    float lraw[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const AtcRow row = atc_bwd_row<PACKED>(p, tile, i, r);
      rok[i] = row.ok;
      lraw[i] = 0.f;
      if (row.ok) {
        lraw[i] = q.lse[(row.seq * p.heads + h) * p.L + row.l];
        const uint4* orow = reinterpret_cast<const uint4*>(q.o + row.tok * q.o_rs + (long long)h * 64);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) ov[i][ch] = orow[ch];
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      float d = 0.f;
      sm100::mbar_wait(&ld_bar[i], 0);
      if (rok[i]) {
        const uint8_t* drow = sdO + i * BLK + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          Vec16<bf16> a, b;
          a.raw = ov[i][ch];
          b.raw = *reinterpret_cast<const uint4*>(drow + (((uint32_t)ch ^ sw) << 4));
#pragma unroll
          for (int e = 0; e < 8; ++e) d = fmaf(a.get(e), b.get(e), d);
        }
      }
      allm[i] = lraw[i] <= 0.5f * ATC_BIG_NEG;
      lse2[i] = lraw[i] * 1.4426950408889634f;
      dlt[i] = d;
    }
    if (dbg_on2) g_attn_dbg[12] = clock64();

#pragma unroll 1
    for (int t = 0; t < NT; ++t) {
      const int j = t / NB, i = t % NB;
      float my_lse2 = lse2[0], my_d = dlt[0];
      bool my_ok = rok[0], my_allm = allm[0];
#pragma unroll
      for (int ii = 1; ii < NB; ++ii)
        if (i == ii) { my_lse2 = lse2[ii]; my_d = dlt[ii]; my_ok = rok[ii]; my_allm = allm[ii]; }
      sm100::mbar_wait(s_full, t & 1);
      sm100::tc_fence_after();
      if (dbg_on2) g_attn_dbg[13 + (t & 3)] = clock64();
      uint32_t pkP[32], pkS[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = g * 64 + hf * 32;
        uint32_t sr[32], dr[32];
        sm100::tmem_ld_32x32(trow + COL_S + c0, sr);
        sm100::tmem_ld_32x32(trow + COL_DP + c0, dr);
        sm100::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = c0 + e + u;
            float pr = atc_exp2(__uint_as_float(sr[e + u]) * k2 - my_lse2);
            bool masked = false;
            if constexpr (MASKED) {
              masked = kpen[j * 128 + c] != 0.f;
              if (masked) pr = my_allm ? inv_l : 0.f;
            }
            if constexpr (PACKED) {
              if (!(my_ok && (unsigned)(c - seq_c0) < (unsigned)p.L)) pr = 0.f;
            } else {
              if (!my_ok) pr = 0.f;
            }
            pv[u] = pr;
            dv[u] = masked ? 0.f : pr * (__uint_as_float(dr[e + u]) - my_d) * q.scale;
          }
          pkP[hf * 16 + (e >> 1)] = atc_pack2(pv[0], pv[1]);
          pkS[hf * 16 + (e >> 1)] = atc_pack2(dv[0], dv[1]);
        }
      }
      sm100::tc_fence_before();
      if (dbg_on2) g_attn_dbg[17 + (t & 3)] = clock64();
      if (t > 0) sm100::mbar_wait(mma3_done, (t - 1) & 1);      // P / dS tiles of the previous tile are consumed
      {
        uint8_t* prow = sP + g * BLK + r * 128;
        uint8_t* srow = sdS + g * BLK + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t off = (((uint32_t)ch ^ sw) << 4);
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(pkP[4 * ch], pkP[4 * ch + 1], pkP[4 * ch + 2], pkP[4 * ch + 3]);
          *reinterpret_cast<uint4*>(srow + off) = make_uint4(pkS[4 * ch], pkS[4 * ch + 1], pkS[4 * ch + 2], pkS[4 * ch + 3]);
        }
      }
      sm100::fence_proxy_async();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(p_full);
      if (dbg_on2) g_attn_dbg[21 + (t & 3)] = clock64();

      if (i == NB - 1) {                    // key block j is complete: warp group 0 drains dK_j, warp group 1 dV_j
        sm100::mbar_wait(mma3_done, t & 1);
        sm100::tc_fence_after();
        const AtcRow kr = atc_bwd_row<PACKED>(p, tile, j, r);
        bf16* dst = (g == 0 ? q.dk + kr.tok * q.dk_rs : q.dv + kr.tok * q.dv_rs) + (long long)h * 64;
        const uint32_t col = g == 0 ? COL_DK : COL_DV;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t rr[32];
          sm100::tmem_ld_32x32(trow + col + c0, rr);
          sm100::tmem_ld_wait();
          if (kr.ok) {
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              uint4 o4 = make_uint4(atc_pack2(__uint_as_float(rr[8 * v4 + 0]), __uint_as_float(rr[8 * v4 + 1])),
                                    atc_pack2(__uint_as_float(rr[8 * v4 + 2]), __uint_as_float(rr[8 * v4 + 3])),
                                    atc_pack2(__uint_as_float(rr[8 * v4 + 4]), __uint_as_float(rr[8 * v4 + 5])),
                                    atc_pack2(__uint_as_float(rr[8 * v4 + 6]), __uint_as_float(rr[8 * v4 + 7])));
              *reinterpret_cast<uint4*>(dst + c0 + 8 * v4) = o4;
            }
          }
        }
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(kv_drained);
        if (dbg_on2) g_attn_dbg[25 + (j & 1)] = clock64();
      }
      if (j == NB - 1) {                    // query block i is complete: each warp group drains 32 of dQ_i's 64 columns
        sm100::mbar_wait(mma3_done, t & 1);
        sm100::tc_fence_after();
        const AtcRow qr = atc_bwd_row<PACKED>(p, tile, i, r);
        bf16* dst = q.dq + qr.tok * q.dq_rs + (long long)h * 64 + g * 32;
        uint32_t rr[32];
        sm100::tmem_ld_32x32(trow + COL_DQ + 64 * i + g * 32, rr);
        sm100::tmem_ld_wait();
        if (qr.ok) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 o4 = make_uint4(atc_pack2(__uint_as_float(rr[8 * v4 + 0]), __uint_as_float(rr[8 * v4 + 1])),
                                  atc_pack2(__uint_as_float(rr[8 * v4 + 2]), __uint_as_float(rr[8 * v4 + 3])),
                                  atc_pack2(__uint_as_float(rr[8 * v4 + 4]), __uint_as_float(rr[8 * v4 + 5])),
                                  atc_pack2(__uint_as_float(rr[8 * v4 + 6]), __uint_as_float(rr[8 * v4 + 7])));
            *reinterpret_cast<uint4*>(dst + 8 * v4) = o4;
          }
        }
      }
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ATB_STAMP(27);
    if (dbg_on) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_attn_dbg[31] = smid;
    }
  }
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<512>(tmem_base);
  }
}

int attn_debug_read(unsigned long long* out32) {
  return cudaMemcpyFromSymbol(out32, g_attn_dbg, 32 * sizeof(unsigned long long)) == cudaSuccess ? VVAE_OK : VVAE_ERR_CUDA;
}

template <int NB, bool PACKED, bool MASKED>
static int atc_launch_bwd(const vvae_attn_args& a, const AttnTcPlan& p, cudaStream_t s) {
  constexpr int SMEM = (4 * NB + 4) * 16384 + 64 + NB * 128 * 4 + 1024;
  CUtensorMap mq, mk, mv, md;
  int rc;
  const int bl = PACKED ? p.L : 128;
  const int bi = PACKED ? (p.pack_inner ? p.G : 1) : 1, bo = PACKED ? (p.pack_inner ? 1 : p.G) : 1;
  if ((rc = atc_make_map(&mq, a.q, a.q_rs, p, bl, bi, bo))) return rc;
  if ((rc = atc_make_map(&mk, a.k, a.k_rs, p, bl, bi, bo))) return rc;
  if ((rc = atc_make_map(&mv, a.v, a.v_rs, p, bl, bi, bo))) return rc;
  if ((rc = atc_make_map(&md, a.d_o, a.do_rs, p, bl, bi, bo))) return rc;
  AttnTcBwdParams q;
  q.pl = p;
  q.o = (const bf16*)a.o; q.o_rs = a.o_rs; q.lse = a.lse;
  q.mask = a.mask; q.mask_seq_div = a.mask_seq_div > 0 ? a.mask_seq_div : 1; q.ms_seq = a.ms_seq; q.ms_k = a.ms_k;
  q.dq = (bf16*)a.dq; q.dk = (bf16*)a.dk; q.dv = (bf16*)a.dv;
  q.dq_rs = a.dq_rs; q.dk_rs = a.dk_rs; q.dv_rs = a.dv_rs;
  q.scale = a.scale;
  q.dbg = (g_dbg[10] & 16) ? 1 : 0;
  q.dbg_cta = (int)g_dbg[0];
  auto kern = attn_bwd_sm100_kernel<NB, PACKED, MASKED>;
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {   // idempotent: a racing second call sets the same value
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("attention bwd: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  const long long tiles = PACKED ? p.tiles : (long long)p.n_outer * p.n_inner;
  dim3 grid((unsigned)tiles, (unsigned)p.heads);
  launch_pdl(kern, grid, dim3(320), SMEM, s, mq, mk, mv, md, q);
  return check_launch("attn_bwd_sm100");
}


// ====================================================================================================== long sequences
// L = 128*NB > 256 (spatial attention at 512x512: L = 1024).  The whole sequence no longer fits one CTA's TMEM, so:
//   forward : one CTA per (sequence, 128-query block, head) streams the key blocks twice -- pass 1 finds the row maxima
//             (S = Q.K_j^T only), pass 2 recomputes S, forms P = exp(S - max) and accumulates O += P.V_j in TMEM without
//             any rescaling; S and P are double buffered so the tensor core runs one block ahead of the softmax warps.
//   backward: two streaming kernels built from one template.  dK/dV: a CTA owns a key block and streams the query
//             blocks (dK_j, dV_j stay in TMEM); dQ: a CTA owns a query block and streams the key blocks (dQ_i stays in
//             TMEM).  S and dP are recomputed in both (2 extra small MMAs per tile) instead of reducing dQ across CTAs
//             with atomics.  D = rowsum(dO o O) comes from a tiny pre-pass (attn_delta_kernel).
struct AttnLongParams {
  AttnTcPlan pl;
  bf16* o; long long o_rs;
  float* lse;
  const float* delta;
  const unsigned char* mask; long long mask_seq_div, ms_seq, ms_k;
  bf16 *dq, *dk, *dv; long long dq_rs, dk_rs, dv_rs;
  float scale;
};

__device__ __forceinline__ long long atl_tok(const AttnTcPlan& p, long long seq, int l) {
  return (seq / p.n_inner) * p.ts_o + (seq % p.n_inner) * p.ts_i + (long long)l * p.ts_l;
}

template <bool MASKED>
__global__ void __launch_bounds__(192, 1)
attn_fwd_long_sm100_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                           const __grid_constant__ CUtensorMap tma_v, const AttnLongParams q) {
  const AttnTcPlan& p = q.pl;
  constexpr int BLK = 16384, KS = 3, VS = 2;
  const int NKB = p.nqb;                               // key blocks of 128 (= query blocks: L = 128*nqb)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + BLK;               // [KS]
  uint8_t* sV = sK + KS * BLK;          // [VS]
  uint8_t* sP = sV + VS * BLK;          // [2][2 key halves][128 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * BLK);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;          // [3]
  uint64_t* k_empty = bars + 4;         // [3]
  uint64_t* v_full = bars + 7;          // [2]
  uint64_t* v_empty = bars + 9;         // [2]
  uint64_t* s_full = bars + 11;         // [2]
  uint64_t* s_empty = bars + 13;        // [2]
  uint64_t* p_full = bars + 15;         // [2]
  uint64_t* p_empty = bars + 17;        // [2]
  uint64_t* o_full = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* kpen = reinterpret_cast<float*>(bars + 21);   // [L]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long seq = blockIdx.x / p.nqb;
  const int qb = (int)(blockIdx.x % p.nqb);
  const int h = blockIdx.y;
  const int c2 = (int)(seq % p.n_inner), c3 = (int)(seq / p.n_inner);

  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_q);
    sm100::tma_prefetch_desc(&tma_k);
    sm100::tma_prefetch_desc(&tma_v);
    sm100::mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) { sm100::mbar_init(&k_full[i], 1); sm100::mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&v_full[i], 1); sm100::mbar_init(&v_empty[i], 1);
      sm100::mbar_init(&s_full[i], 1); sm100::mbar_init(&s_empty[i], 4);
      sm100::mbar_init(&p_full[i], 4); sm100::mbar_init(&p_empty[i], 1);
    }
    sm100::mbar_init(o_full, 1);
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<512>(tmem_slot);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_O = 256;

  if (warp == 0) {
    if (lane == 0) {
      sm100::mbar_expect_tx(q_full, BLK);
      tma_load_4d(sQ, &tma_q, q_full, h * 64, qb * 128, c2, c3);
      for (int s = 0; s < 2 * NKB; ++s) {
        const int j = s % NKB, slot = s % KS;
        sm100::mbar_wait(&k_empty[slot], ((s / KS) & 1) ^ 1);
        sm100::mbar_expect_tx(&k_full[slot], BLK);
        tma_load_4d(sK + slot * BLK, &tma_k, &k_full[slot], h * 64, j * 128, c2, c3);
        if (s >= NKB) {
          const int vs = j & 1;
          sm100::mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
          sm100::mbar_expect_tx(&v_full[vs], BLK);
          tma_load_4d(sV + vs * BLK, &tma_v, &v_full[vs], h * 64, j * 128, c2, c3);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = sm100::make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_o = sm100::make_idesc_bf16(128, 64, false, true);
      const uint32_t qa = sm100::smem_u32(sQ);
      auto issue_s = [&](int s) {
        const int slot = s % KS, b = s & 1;
        sm100::mbar_wait(&k_full[slot], (s / KS) & 1);
        if (s >= 2) sm100::mbar_wait(&s_empty[b], ((s >> 1) - 1) & 1);
        sm100::tc_fence_after();
        const uint32_t ka = sm100::smem_u32(sK + slot * BLK);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + b * 128, sm100::make_smem_desc_sw128(qa + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        sm100::umma_commit(&k_empty[slot]);
        sm100::umma_commit(&s_full[b]);
      };
      sm100::mbar_wait(q_full, 0);
      issue_s(0);
      for (int s = 0; s < 2 * NKB; ++s) {
        if (s + 1 < 2 * NKB) issue_s(s + 1);
        if (s >= NKB) {
          const int j = s - NKB, pb = j & 1;
          sm100::mbar_wait(&v_full[pb], (j >> 1) & 1);
          sm100::mbar_wait(&p_full[pb], (j >> 1) & 1);
          sm100::tc_fence_after();
          const uint32_t pa = sm100::smem_u32(sP + pb * 2 * BLK), va = sm100::smem_u32(sV + pb * BLK);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            sm100::umma_f16(tmem_base + COL_O, sm100::make_smem_desc_sw128(pa + (k >> 2) * BLK + (k & 3) * 32, 16, 1024),
                            sm100::make_smem_desc_sw128(va + k * 2048, 8192, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          sm100::umma_commit(&v_empty[pb]);
          sm100::umma_commit(&p_empty[pb]);
        }
      }
      sm100::umma_commit(o_full);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    if (MASKED) {
      for (int c = tid; c < p.L; c += 128)
        kpen[c] = q.mask[(seq / q.mask_seq_div) * q.ms_seq + (long long)c * q.ms_k] == 0 ? 1.f : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const float k2 = q.scale * 1.4426950408889634f;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t sw = (uint32_t)(r & 7);
    // pass 1: row maximum
    float mx = -INFINITY;
    for (int s = 0; s < NKB; ++s) {
      const int b = s & 1;
      sm100::mbar_wait(&s_full[b], (s >> 1) & 1);
      sm100::tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t sr[32];
        sm100::tmem_ld_32x32(trow + b * 128 + c0, sr);
        sm100::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = __uint_as_float(sr[i]) * q.scale;
          if constexpr (MASKED) { if (kpen[s * 128 + c0 + i] != 0.f) v = ATC_BIG_NEG; }
          mx = fmaxf(mx, v);
        }
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&s_empty[b]);
    }
    // pass 2: P = exp(S - max) (bf16, swizzled K-major operand tile), row sums
    const float mx2 = mx * 1.4426950408889634f;
    float sum = 0.f;
    for (int j = 0; j < NKB; ++j) {
      const int s = NKB + j, b = s & 1, pb = j & 1;
      sm100::mbar_wait(&s_full[b], (s >> 1) & 1);
      if (j >= 2) sm100::mbar_wait(&p_empty[pb], ((j >> 1) - 1) & 1);
      sm100::tc_fence_after();
      uint8_t* prow = sP + pb * 2 * BLK + r * 128;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t sr[32];
        sm100::tmem_ld_32x32(trow + b * 128 + c0, sr);
        sm100::tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = atc_exp2(__uint_as_float(sr[i]) * k2 - mx2);
          if constexpr (MASKED) { if (kpen[j * 128 + c0 + i] != 0.f) e = (mx == ATC_BIG_NEG ? 1.f : 0.f); }
          pv[i] = e;
          sum += e;
        }
        uint8_t* pblk = prow + (c0 >> 6) * BLK;
        const uint32_t ch0 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4)
          *reinterpret_cast<uint4*>(pblk + (((ch0 + v4) ^ sw) << 4)) =
              make_uint4(atc_pack2(pv[8 * v4], pv[8 * v4 + 1]), atc_pack2(pv[8 * v4 + 2], pv[8 * v4 + 3]),
                         atc_pack2(pv[8 * v4 + 4], pv[8 * v4 + 5]), atc_pack2(pv[8 * v4 + 6], pv[8 * v4 + 7]));
      }
      sm100::fence_proxy_async();
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) { sm100::mbar_arrive(&s_empty[b]); sm100::mbar_arrive(&p_full[pb]); }
    }
    sm100::mbar_wait(o_full, 0);
    sm100::tc_fence_after();
    const float inv = 1.f / sum;
    const int l = qb * 128 + r;
    bf16* orow = q.o + atl_tok(p, seq, l) * q.o_rs + (long long)h * 64;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t orr[32];
      sm100::tmem_ld_32x32(trow + COL_O + c0, orr);
      sm100::tmem_ld_wait();
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4)
        *reinterpret_cast<uint4*>(orow + c0 + 8 * v4) =
            make_uint4(atc_pack2(__uint_as_float(orr[8 * v4]) * inv, __uint_as_float(orr[8 * v4 + 1]) * inv),
                       atc_pack2(__uint_as_float(orr[8 * v4 + 2]) * inv, __uint_as_float(orr[8 * v4 + 3]) * inv),
                       atc_pack2(__uint_as_float(orr[8 * v4 + 4]) * inv, __uint_as_float(orr[8 * v4 + 5]) * inv),
                       atc_pack2(__uint_as_float(orr[8 * v4 + 6]) * inv, __uint_as_float(orr[8 * v4 + 7]) * inv));
    }
    if (q.lse) q.lse[(seq * p.heads + h) * p.L + l] = mx + __logf(sum);
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<512>(tmem_base);
  }
}

// delta[seq, h, l] = sum_d O[tok, h, d] * dO[tok, h, d]; 8 lanes per (token, head), 16 bytes per lane
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, long long o_rs, const bf16* __restrict__ d_o, long long do_rs,
                  float* __restrict__ delta, AttnTcPlan p, long long total) {
  const int cph = 8 * p.heads;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();                                            // dO is the previous kernel's output
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx - lane < total;
       idx += (long long)gridDim.x * blockDim.x) {       // whole warps iterate together (shuffles below)
    const bool act = idx < total;
    const long long ii = act ? idx : 0;
    const int chunk = (int)(ii % cph);
    const long long rowid = ii / cph;
    const long long seq = rowid / p.L;
    const int l = (int)(rowid % p.L);
    const long long tok = atl_tok(p, seq, l);
    float d = 0.f;
    if (act) {
      Vec16<bf16> a, b;
      a.load(o + tok * o_rs + chunk * 8);
      b.load(d_o + tok * do_rs + chunk * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(a.get(e), b.get(e), d);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (act && (chunk & 7) == 0) delta[(seq * p.heads + (chunk >> 3)) * p.L + l] = d;
  }
}

// DQ_MODE = false: CTA owns key block `blk`, streams query blocks, produces dK, dV.
// DQ_MODE = true : CTA owns query block `blk`, streams key blocks, produces dQ.
template <bool DQ_MODE, bool MASKED>
__global__ void __launch_bounds__(320, 1)
attn_bwd_long_sm100_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                           const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                           const AttnLongParams q) {
  const AttnTcPlan& p = q.pl;
  constexpr int BLK = 16384;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_A0 = 256, COL_A1 = 320;
  const int NB = p.nqb;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sR0 = smem;                  // resident: K_j | Q_i
  uint8_t* sR1 = sR0 + BLK;             // resident: V_j | dO_i
  uint8_t* sS0 = sR1 + BLK;             // streamed [2]: Q_x | K_x
  uint8_t* sS1 = sS0 + 2 * BLK;         // streamed [2]: dO_x | V_x
  uint8_t* sP = sS1 + 2 * BLK;          // [2 key halves][128 q][64 keys]  (dK/dV mode only)
  uint8_t* sdS = sP + 2 * BLK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * BLK);
  uint64_t* res_full = bars;
  uint64_t* st_full = bars + 1;         // [2]
  uint64_t* st_empty = bars + 3;        // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* mma3_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* kpen = reinterpret_cast<float*>(bars + 9);    // [L]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long seq = blockIdx.x / NB;
  const int blk = (int)(blockIdx.x % NB);
  const int h = blockIdx.y;
  const int c2 = (int)(seq % p.n_inner), c3 = (int)(seq / p.n_inner);

  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_q);
    sm100::tma_prefetch_desc(&tma_k);
    sm100::tma_prefetch_desc(&tma_v);
    sm100::tma_prefetch_desc(&tma_do);
    sm100::mbar_init(res_full, 1);
    for (int i = 0; i < 2; ++i) { sm100::mbar_init(&st_full[i], 1); sm100::mbar_init(&st_empty[i], 1); }
    sm100::mbar_init(s_full, 1);
    sm100::mbar_init(p_full, 8);
    sm100::mbar_init(mma3_done, 1);
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<512>(tmem_slot);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      sm100::mbar_expect_tx(res_full, 2 * BLK);
      if (DQ_MODE) {
        tma_load_4d(sR0, &tma_q, res_full, h * 64, blk * 128, c2, c3);
        tma_load_4d(sR1, &tma_do, res_full, h * 64, blk * 128, c2, c3);
      } else {
        tma_load_4d(sR0, &tma_k, res_full, h * 64, blk * 128, c2, c3);
        tma_load_4d(sR1, &tma_v, res_full, h * 64, blk * 128, c2, c3);
      }
      for (int x = 0; x < NB; ++x) {
        const int slot = x & 1;
        sm100::mbar_wait(&st_empty[slot], ((x >> 1) & 1) ^ 1);
        sm100::mbar_expect_tx(&st_full[slot], 2 * BLK);
        if (DQ_MODE) {
          tma_load_4d(sS0 + slot * BLK, &tma_k, &st_full[slot], h * 64, x * 128, c2, c3);
          tma_load_4d(sS1 + slot * BLK, &tma_v, &st_full[slot], h * 64, x * 128, c2, c3);
        } else {
          tma_load_4d(sS0 + slot * BLK, &tma_q, &st_full[slot], h * 64, x * 128, c2, c3);
          tma_load_4d(sS1 + slot * BLK, &tma_do, &st_full[slot], h * 64, x * 128, c2, c3);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = sm100::make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_kv = sm100::make_idesc_bf16(128, 64, true, true);
      constexpr uint32_t idesc_q = sm100::make_idesc_bf16(128, 64, false, true);
      const uint32_t r0 = sm100::smem_u32(sR0), r1 = sm100::smem_u32(sR1);
      const uint32_t aP = sm100::smem_u32(sP), aS = sm100::smem_u32(sdS);
      auto issue_sdp = [&](int x) {
        const int slot = x & 1;
        sm100::mbar_wait(&st_full[slot], (x >> 1) & 1);
        sm100::tc_fence_after();
        const uint32_t s0 = sm100::smem_u32(sS0 + slot * BLK), s1 = sm100::smem_u32(sS1 + slot * BLK);
        const uint32_t aq = DQ_MODE ? r0 : s0, ak = DQ_MODE ? s0 : r0, ad = DQ_MODE ? r1 : s1, av = DQ_MODE ? s1 : r1;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + COL_S, sm100::make_smem_desc_sw128(aq + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(ak + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + COL_DP, sm100::make_smem_desc_sw128(ad + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(av + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        sm100::umma_commit(s_full);
      };
      sm100::mbar_wait(res_full, 0);
      issue_sdp(0);
      for (int x = 0; x < NB; ++x) {
        const int slot = x & 1;
        sm100::mbar_wait(p_full, x & 1);
        sm100::tc_fence_after();
        if (x + 1 < NB) issue_sdp(x + 1);
        const uint32_t s0 = sm100::smem_u32(sS0 + slot * BLK), s1 = sm100::smem_u32(sS1 + slot * BLK);
        if (DQ_MODE) {        // dQ_i (+)= dS . K_x
#pragma unroll
          for (int k = 0; k < 8; ++k)
            sm100::umma_f16(tmem_base + COL_A0, sm100::make_smem_desc_sw128(aS + (k >> 2) * BLK + (k & 3) * 32, 16, 1024),
                            sm100::make_smem_desc_sw128(s0 + k * 2048, 8192, 1024), idesc_q, (x > 0 || k > 0) ? 1u : 0u);
        } else {              // dV_j (+)= P^T . dO_x ; dK_j (+)= dS^T . Q_x
#pragma unroll
          for (int k = 0; k < 8; ++k)
            sm100::umma_f16(tmem_base + COL_A1, sm100::make_smem_desc_sw128(aP + k * 2048, BLK, 1024),
                            sm100::make_smem_desc_sw128(s1 + k * 2048, 8192, 1024), idesc_kv, (x > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            sm100::umma_f16(tmem_base + COL_A0, sm100::make_smem_desc_sw128(aS + k * 2048, BLK, 1024),
                            sm100::make_smem_desc_sw128(s0 + k * 2048, 8192, 1024), idesc_kv, (x > 0 || k > 0) ? 1u : 0u);
        }
        sm100::umma_commit(&st_empty[slot]);
        sm100::umma_commit(mma3_done);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int g = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    if (MASKED) {
      for (int c = tid; c < p.L; c += 256)
        kpen[c] = q.mask[(seq / q.mask_seq_div) * q.ms_seq + (long long)c * q.ms_k] == 0 ? 1.f : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t sw = (uint32_t)(r & 7);
    const float k2 = q.scale * 1.4426950408889634f;
    const float inv_l = 1.f / (float)p.L;
    const long long stat_base = (seq * p.heads + h) * p.L;
    float lse_c = 0.f, dlt_c = 0.f;
    if (DQ_MODE) { lse_c = q.lse[stat_base + blk * 128 + r]; dlt_c = q.delta[stat_base + blk * 128 + r]; }
    else { lse_c = q.lse[stat_base + r]; dlt_c = q.delta[stat_base + r]; }

#pragma unroll 1
    for (int x = 0; x < NB; ++x) {
      const float lraw = lse_c, my_d = dlt_c;
      if (!DQ_MODE && x + 1 < NB) {            // prefetch the next query block's statistics
        lse_c = q.lse[stat_base + (x + 1) * 128 + r];
        dlt_c = q.delta[stat_base + (x + 1) * 128 + r];
      }
      const float my_lse2 = lraw * 1.4426950408889634f;
      const bool my_allm = lraw <= 0.5f * ATC_BIG_NEG;
      const int kb = DQ_MODE ? x : blk;        // key block of this tile
      sm100::mbar_wait(s_full, x & 1);
      sm100::tc_fence_after();
      uint32_t pkP[32], pkS[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = g * 64 + hf * 32;
        uint32_t sr[32], dr[32];
        sm100::tmem_ld_32x32(trow + COL_S + c0, sr);
        sm100::tmem_ld_32x32(trow + COL_DP + c0, dr);
        sm100::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float pr = atc_exp2(__uint_as_float(sr[e + u]) * k2 - my_lse2);
            bool masked = false;
            if constexpr (MASKED) {
              masked = kpen[kb * 128 + c0 + e + u] != 0.f;
              if (masked) pr = my_allm ? inv_l : 0.f;
            }
            pv[u] = pr;
            dv[u] = masked ? 0.f : pr * (__uint_as_float(dr[e + u]) - my_d) * q.scale;
          }
          pkP[hf * 16 + (e >> 1)] = atc_pack2(pv[0], pv[1]);
          pkS[hf * 16 + (e >> 1)] = atc_pack2(dv[0], dv[1]);
        }
      }
      sm100::tc_fence_before();
      if (x > 0) sm100::mbar_wait(mma3_done, (x - 1) & 1);
      {
        uint8_t* prow = sP + g * BLK + r * 128;
        uint8_t* srow = sdS + g * BLK + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t off = (((uint32_t)ch ^ sw) << 4);
          if (!DQ_MODE)
            *reinterpret_cast<uint4*>(prow + off) = make_uint4(pkP[4 * ch], pkP[4 * ch + 1], pkP[4 * ch + 2], pkP[4 * ch + 3]);
          *reinterpret_cast<uint4*>(srow + off) = make_uint4(pkS[4 * ch], pkS[4 * ch + 1], pkS[4 * ch + 2], pkS[4 * ch + 3]);
        }
      }
      sm100::fence_proxy_async();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(p_full);
    }
    sm100::mbar_wait(mma3_done, (NB - 1) & 1);
    sm100::tc_fence_after();
    const long long tok = atl_tok(p, seq, blk * 128 + r);
    if (DQ_MODE) {            // each warp group drains 32 of dQ's 64 columns
      bf16* dst = q.dq + tok * q.dq_rs + (long long)h * 64 + g * 32;
      uint32_t rr[32];
      sm100::tmem_ld_32x32(trow + COL_A0 + g * 32, rr);
      sm100::tmem_ld_wait();
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4)
        *reinterpret_cast<uint4*>(dst + 8 * v4) =
            make_uint4(atc_pack2(__uint_as_float(rr[8 * v4]), __uint_as_float(rr[8 * v4 + 1])),
                       atc_pack2(__uint_as_float(rr[8 * v4 + 2]), __uint_as_float(rr[8 * v4 + 3])),
                       atc_pack2(__uint_as_float(rr[8 * v4 + 4]), __uint_as_float(rr[8 * v4 + 5])),
                       atc_pack2(__uint_as_float(rr[8 * v4 + 6]), __uint_as_float(rr[8 * v4 + 7])));
    } else {                  // warp group 0 drains dK, warp group 1 dV
      bf16* dst = (g == 0 ? q.dk + tok * q.dk_rs : q.dv + tok * q.dv_rs) + (long long)h * 64;
      const uint32_t col = g == 0 ? COL_A0 : COL_A1;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t rr[32];
        sm100::tmem_ld_32x32(trow + col + c0, rr);
        sm100::tmem_ld_wait();
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4)
          *reinterpret_cast<uint4*>(dst + c0 + 8 * v4) =
              make_uint4(atc_pack2(__uint_as_float(rr[8 * v4]), __uint_as_float(rr[8 * v4 + 1])),
                         atc_pack2(__uint_as_float(rr[8 * v4 + 2]), __uint_as_float(rr[8 * v4 + 3])),
                         atc_pack2(__uint_as_float(rr[8 * v4 + 4]), __uint_as_float(rr[8 * v4 + 5])),
                         atc_pack2(__uint_as_float(rr[8 * v4 + 6]), __uint_as_float(rr[8 * v4 + 7])));
      }
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<512>(tmem_base);
  }
}

template <typename K>
static int atl_set_smem(K kern, int bytes, std::atomic<bool>* done) {
  if (done->load(std::memory_order_acquire)) return VVAE_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("attention (long): cudaFuncSetAttribute(%d): %s", bytes, cudaGetErrorString(e));
    return VVAE_ERR_CUDA;
  }
  done->store(true, std::memory_order_release);
  return VVAE_OK;
}

static void atl_fill_params(const vvae_attn_args& a, const AttnTcPlan& p, AttnLongParams& q) {
  q.pl = p;
  q.o = (bf16*)a.o; q.o_rs = a.o_rs; q.lse = a.lse; q.delta = a.delta;
  q.mask = a.mask; q.mask_seq_div = a.mask_seq_div > 0 ? a.mask_seq_div : 1; q.ms_seq = a.ms_seq; q.ms_k = a.ms_k;
  q.dq = (bf16*)a.dq; q.dk = (bf16*)a.dk; q.dv = (bf16*)a.dv;
  q.dq_rs = a.dq_rs; q.dk_rs = a.dk_rs; q.dv_rs = a.dv_rs;
  q.scale = a.scale;
}

static int atl_launch_fwd(const vvae_attn_args& a, const AttnTcPlan& p, cudaStream_t s) {
  const int SMEM = 10 * 16384 + 256 + p.L * 4 + 1024;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = atc_make_map(&mq, a.q, a.q_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mk, a.k, a.k_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mv, a.v, a.v_rs, p, 128, 1, 1))) return rc;
  AttnLongParams q;
  atl_fill_params(a, p, q);
  static std::atomic<bool> set_m{false}, set_u{false};
  dim3 grid((unsigned)((long long)p.n_outer * p.n_inner * p.nqb), (unsigned)p.heads);
  if (a.mask) {
    if ((rc = atl_set_smem(attn_fwd_long_sm100_kernel<true>, 227 * 1024, &set_m))) return rc;
    attn_fwd_long_sm100_kernel<true><<<grid, 192, SMEM, s>>>(mq, mk, mv, q);
  } else {
    if ((rc = atl_set_smem(attn_fwd_long_sm100_kernel<false>, 227 * 1024, &set_u))) return rc;
    attn_fwd_long_sm100_kernel<false><<<grid, 192, SMEM, s>>>(mq, mk, mv, q);
  }
  return check_launch("attn_fwd_long_sm100");
}

static int atl_launch_bwd(const vvae_attn_args& a, const AttnTcPlan& p, cudaStream_t s) {
  const int SMEM = 10 * 16384 + 256 + p.L * 4 + 1024;
  CUtensorMap mq, mk, mv, md;
  int rc;
  if ((rc = atc_make_map(&mq, a.q, a.q_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mk, a.k, a.k_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mv, a.v, a.v_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&md, a.d_o, a.do_rs, p, 128, 1, 1))) return rc;
  AttnLongParams q;
  atl_fill_params(a, p, q);
  const long long n_seq = (long long)p.n_outer * p.n_inner;
  {
    const long long total = n_seq * p.L * 8 * p.heads;
    const int blocks = (int)std::min<long long>(cdiv(total, 256), (long long)num_sms() * 16);
    attn_delta_kernel<<<blocks, 256, 0, s>>>((const bf16*)a.o, a.o_rs, (const bf16*)a.d_o, a.do_rs, a.delta, p, total);
    if ((rc = check_launch("attn_delta"))) return rc;
  }
  static std::atomic<bool> set[4] = {{false}, {false}, {false}, {false}};
  dim3 grid((unsigned)(n_seq * p.nqb), (unsigned)p.heads);
  if (a.mask) {
    if ((rc = atl_set_smem(attn_bwd_long_sm100_kernel<false, true>, 227 * 1024, &set[0]))) return rc;
    if ((rc = atl_set_smem(attn_bwd_long_sm100_kernel<true, true>, 227 * 1024, &set[1]))) return rc;
    attn_bwd_long_sm100_kernel<false, true><<<grid, 320, SMEM, s>>>(mq, mk, mv, md, q);
    attn_bwd_long_sm100_kernel<true, true><<<grid, 320, SMEM, s>>>(mq, mk, mv, md, q);
  } else {
    if ((rc = atl_set_smem(attn_bwd_long_sm100_kernel<false, false>, 227 * 1024, &set[2]))) return rc;
    if ((rc = atl_set_smem(attn_bwd_long_sm100_kernel<true, false>, 227 * 1024, &set[3]))) return rc;
    attn_bwd_long_sm100_kernel<false, false><<<grid, 320, SMEM, s>>>(mq, mk, mv, md, q);
    attn_bwd_long_sm100_kernel<true, false><<<grid, 320, SMEM, s>>>(mq, mk, mv, md, q);
  }
  return check_launch("attn_bwd_long_sm100");
}


// ====================================================================================================== backward, L = 256
// Persistent variant of the backward for sequences of 256 without a mask (spatial attention at 256x256 pixels: 21 of the
// 42 attention backward launches of a step).  The one-unit-per-CTA kernel above spends 41 k cycles per (sequence, head)
// of which the tensor core is busy 12 k (scripts/attn_timeline.py): operand loads, D = rowsum(dO o O), the accumulator
// drains and the softmax all sit on one serial chain, and 512 TMEM columns + 192 KB of shared memory leave no room for a
// second CTA to hide it.  Here one CTA per SM walks over its (sequence, head) units and the chain is cut into roles:
//   warp 0      producer: TMA loads of the NEXT unit's Q/K/V/dO blocks go into each 32 KB group of the operand buffers
//               as soon as the tile that last reads the group has finished (groups: {K0,V0} {Q0,dO0} {Q1,dO1} {K1,V1});
//   warp 1      MMA issuer: S/dP of tile T+1 are issued as soon as the softmax warps have READ S/dP of tile T out of
//               TMEM (s_free), ahead of the three output contractions of tile T, so the tensor pipe never idles behind
//               the softmax;
//   warps 4-11  softmax: P = exp2(S*k - LSE), dS = P o (dP - D) * scale, written as bf16 into the one swizzled layout the
//               tensor core reads K-major and MN-major (as above); D and LSE come from global memory (D from
//               attn_delta_kernel, a 67 MB pre-pass), prefetched one unit ahead;
//   warps 12-15 drain: dK_j, dV_j (after each key block) and dQ_i (after each query block's last tile) leave TMEM while
//               the other roles carry on with the next tile / unit.
// Register budget by setmaxnreg: softmax warps 184, drain 96, producer / issuer 48.
// Measured (scripts/attn_timeline.py, 128 sequences x 8 heads): 101 us against 184 us, 17 k cycles per unit = 3.5-3.9 k
// per tile (contraction 2.5 k + P/dS store 0.7 k: P/dS are single-buffered, so the contraction of tile T cannot overlap
// the store of tile T+1) plus ~3 k at the start of a unit, where S/dP of the second tile wait for Q1 / dO1, which the
// last MMAs of the previous unit were still reading.  A second buffer for that group (instead of the staging tiles, with
// the gradient tiles staged in dead operand buffers) moves the same wait to K1 / V1: tried, no gain, not kept.
// Every mbarrier completes exactly once per unit, so a wait's parity is the unit counter's low bit.
struct AttnBwd256Params {
  AttnTcPlan pl;
  const float* lse; const float* delta;
  bf16 *dq, *dk, *dv; long long dq_rs, dk_rs, dv_rs;
  float scale;
  long long units;
  int dbg, dbg_cta;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(sm100::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void atb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void atb_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void atb_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(512, 1)
attn_bwd256_sm100_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                         const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                         const __grid_constant__ CUtensorMap tma_dq, const __grid_constant__ CUtensorMap tma_dk,
                         const __grid_constant__ CUtensorMap tma_dv, const AttnBwd256Params q) {
  const AttnTcPlan& p = q.pl;
  constexpr int BLK = 16384;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 384, COL_DV = 448;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * BLK;
  uint8_t* sV = sK + 2 * BLK;
  uint8_t* sdO = sV + 2 * BLK;
  uint8_t* sP = sdO + 2 * BLK;                     // [2 key halves][128 q][64 keys] bf16, swizzled
  uint8_t* sdS = sP + 2 * BLK;
  uint8_t* stage = sdS + 2 * BLK;                  // [2] gradient tiles on their way out (TMA store)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * BLK);
  uint64_t* ld_bar = bars;                         // [4] operand groups A..D
  uint64_t* s_full = bars + 4;                     // [4] S/dP of tile t are in TMEM
  uint64_t* s_free = bars + 8;                     // [4] ... and have been read out
  uint64_t* p_full = bars + 12;                    // [4] P/dS of tile t are in shared memory
  uint64_t* mma_done = bars + 16;                  // [4] output contractions of tile t are complete
  uint64_t* kv_drained = bars + 20;                // [2] dK_j/dV_j have left TMEM
  uint64_t* dq_drained = bars + 22;                // [2] dQ_i has left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg_on = (q.dbg & 1) && (int)blockIdx.x == q.dbg_cta && lane == 0;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_q);
    sm100::tma_prefetch_desc(&tma_k);
    sm100::tma_prefetch_desc(&tma_v);
    sm100::tma_prefetch_desc(&tma_do);
    sm100::tma_prefetch_desc(&tma_dq);
    sm100::tma_prefetch_desc(&tma_dk);
    sm100::tma_prefetch_desc(&tma_dv);
    for (int b = 0; b < 4; ++b) {
      sm100::mbar_init(&ld_bar[b], 1);
      sm100::mbar_init(&s_full[b], 1);
      sm100::mbar_init(&s_free[b], 8);
      sm100::mbar_init(&p_full[b], 8);
      sm100::mbar_init(&mma_done[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      sm100::mbar_init(&kv_drained[b], 4);
      sm100::mbar_init(&dq_drained[b], 4);
    }
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<512>(tmem_slot);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const long long first = blockIdx.x, stride = gridDim.x;
  const int n_it = first < q.units ? (int)((q.units - first + stride - 1) / stride) : 0;

  if (warp < 4) {
    reg_dealloc<48>();
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------------------------------------------- producer
      for (int it = 0; it < n_it; ++it) {
        const long long u = first + (long long)it * stride;
        const long long seq = u / p.heads;
        const int h = (int)(u % p.heads);
        const int c3 = (int)(seq / p.n_inner), c2 = (int)(seq % p.n_inner);
        const uint32_t pp = (uint32_t)(it - 1) & 1u;
        if (it > 0) sm100::mbar_wait(&mma_done[1], pp);          // K0, V0: last read by tile 1
        sm100::mbar_expect_tx(&ld_bar[0], 2 * BLK);
        tma_load_4d(sK, &tma_k, &ld_bar[0], h * 64, 0, c2, c3);
        tma_load_4d(sV, &tma_v, &ld_bar[0], h * 64, 0, c2, c3);
        if (it > 0) sm100::mbar_wait(&mma_done[2], pp);          // Q0, dO0: last read by tile 2
        sm100::mbar_expect_tx(&ld_bar[1], 2 * BLK);
        tma_load_4d(sQ, &tma_q, &ld_bar[1], h * 64, 0, c2, c3);
        tma_load_4d(sdO, &tma_do, &ld_bar[1], h * 64, 0, c2, c3);
        if (it > 0) sm100::mbar_wait(&mma_done[3], pp);          // Q1, dO1, K1, V1: last read by tile 3
        sm100::mbar_expect_tx(&ld_bar[2], 2 * BLK);
        tma_load_4d(sQ + BLK, &tma_q, &ld_bar[2], h * 64, 128, c2, c3);
        tma_load_4d(sdO + BLK, &tma_do, &ld_bar[2], h * 64, 128, c2, c3);
        sm100::mbar_expect_tx(&ld_bar[3], 2 * BLK);
        tma_load_4d(sK + BLK, &tma_k, &ld_bar[3], h * 64, 128, c2, c3);
        tma_load_4d(sV + BLK, &tma_v, &ld_bar[3], h * 64, 128, c2, c3);
        // the operand buffers are released one group at a time, only ~1-3 k cycles before the next unit needs them:
        // pull the next unit's tiles into L2 now, a whole unit ahead, so those loads are L2 hits
        if (it + 1 < n_it && !(q.dbg & 4)) {
          const long long u1 = u + stride;
          const long long seq1 = u1 / p.heads;
          const int h1 = (int)(u1 % p.heads);
          const int d3 = (int)(seq1 / p.n_inner), d2 = (int)(seq1 % p.n_inner);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            tma_prefetch_l2_4d(&tma_k, h1 * 64, blk * 128, d2, d3);
            tma_prefetch_l2_4d(&tma_v, h1 * 64, blk * 128, d2, d3);
            tma_prefetch_l2_4d(&tma_q, h1 * 64, blk * 128, d2, d3);
            tma_prefetch_l2_4d(&tma_do, h1 * 64, blk * 128, d2, d3);
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ---------------------------------------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = sm100::make_idesc_bf16(128, 128, false, false);   // S, dP: both K-major
      constexpr uint32_t idesc_kv = sm100::make_idesc_bf16(128, 64, true, true);     // dV, dK: A^T (MN-major), B MN-major
      constexpr uint32_t idesc_q = sm100::make_idesc_bf16(128, 64, false, true);     // dQ: A K-major, B MN-major
      const uint32_t aQ = sm100::smem_u32(sQ), aK = sm100::smem_u32(sK), aV = sm100::smem_u32(sV);
      const uint32_t aO = sm100::smem_u32(sdO), aP = sm100::smem_u32(sP), aS = sm100::smem_u32(sdS);
      const int total = 4 * n_it;
      auto issue_sdp = [&](int T) {                // tile T of the flat sequence: loads, then S and dP
        const int t = T & 3, j = t >> 1, i = t & 1;
        const uint32_t ph = (uint32_t)(T >> 2) & 1u;
        if (t == 0) { sm100::mbar_wait(&ld_bar[0], ph); sm100::mbar_wait(&ld_bar[1], ph); }
        else if (t == 1) sm100::mbar_wait(&ld_bar[2], ph);
        else if (t == 2) sm100::mbar_wait(&ld_bar[3], ph);
        sm100::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + COL_S, sm100::make_smem_desc_sw128(aQ + i * BLK + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(aK + j * BLK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + COL_DP, sm100::make_smem_desc_sw128(aO + i * BLK + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(aV + j * BLK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        sm100::umma_commit(&s_full[t]);
      };
      if (total > 0) issue_sdp(0);
#pragma unroll 1
      for (int T = 0; T < total; ++T) {
        const int it = T >> 2, t = T & 3, j = t >> 1, i = t & 1;
        const uint32_t ph = (uint32_t)it & 1u;
        if (dbg_on && it == 1) g_attn_dbg[t] = clock64();
        if (T + 1 < total) {
          sm100::mbar_wait(&s_free[t], ph);        // S/dP of tile T are in registers: the columns can be overwritten
          sm100::tc_fence_after();
          issue_sdp(T + 1);
        }
        sm100::mbar_wait(&p_full[t], ph);
        if (dbg_on && it == 1) g_attn_dbg[4 + t] = clock64();
        if (it > 0 && j == 0) sm100::mbar_wait(&dq_drained[i], ph ^ 1u);   // dQ_i of the previous unit has left TMEM
        sm100::tc_fence_after();
        // dQ_i (+)= dS K_j : contraction over the 128 keys.  First: it does not wait for the dK/dV drain.
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sm100::umma_f16(tmem_base + COL_DQ + 64 * i,
                          sm100::make_smem_desc_sw128(aS + (k >> 2) * BLK + (k & 3) * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(aK + j * BLK + k * 2048, 8192, 1024), idesc_q,
                          (j > 0 || k > 0) ? 1u : 0u);
        if (i == 0) {                              // dK/dV are about to be overwritten: the previous key block has been read out
          if (j == 1) sm100::mbar_wait(&kv_drained[0], ph);
          else if (it > 0) sm100::mbar_wait(&kv_drained[1], ph ^ 1u);
          sm100::tc_fence_after();
        }
        // dV_j (+)= P^T dO_i ; dK_j (+)= dS^T Q_i : contraction over the 128 queries, 16 per step
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sm100::umma_f16(tmem_base + COL_DV, sm100::make_smem_desc_sw128(aP + k * 2048, BLK, 1024),
                          sm100::make_smem_desc_sw128(aO + i * BLK + k * 2048, 8192, 1024), idesc_kv,
                          (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sm100::umma_f16(tmem_base + COL_DK, sm100::make_smem_desc_sw128(aS + k * 2048, BLK, 1024),
                          sm100::make_smem_desc_sw128(aQ + i * BLK + k * 2048, 8192, 1024), idesc_kv,
                          (i > 0 || k > 0) ? 1u : 0u);
        sm100::umma_commit(&mma_done[t]);
        if (dbg_on && it == 1) g_attn_dbg[8 + t] = clock64();
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------------------------------------ softmax
    reg_alloc<184>();
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
    const int g = (warp - 4) >> 2;                     // key-column half of the tile
    const int r = quarter * 32 + lane;                 // row of the 128-row block == TMEM lane
    const bool dbg2 = dbg_on && warp == 4;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t sw = (uint32_t)(r & 7);
    const float k2 = q.scale * 1.4426950408889634f;
    float n_lse[2] = {0.f, 0.f}, n_dlt[2] = {0.f, 0.f};
    if (n_it > 0) {
      const long long base = first * 256 + r;          // (seq*heads + h) == unit index
      n_lse[0] = q.lse[base]; n_lse[1] = q.lse[base + 128];
      n_dlt[0] = q.delta[base]; n_dlt[1] = q.delta[base + 128];
    }
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const float lse2[2] = {n_lse[0] * 1.4426950408889634f, n_lse[1] * 1.4426950408889634f};
      const float dlt[2] = {n_dlt[0], n_dlt[1]};
      if (it + 1 < n_it) {
        const long long base = (first + (long long)(it + 1) * stride) * 256 + r;
        n_lse[0] = q.lse[base]; n_lse[1] = q.lse[base + 128];
        n_dlt[0] = q.delta[base]; n_dlt[1] = q.delta[base + 128];
      }
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const int i = t & 1;
        const float my_lse2 = i ? lse2[1] : lse2[0], my_d = i ? dlt[1] : dlt[0];
        sm100::mbar_wait(&s_full[t], ph);
        sm100::tc_fence_after();
        if (dbg2 && it == 1) g_attn_dbg[12 + t] = clock64();
        uint32_t pkP[32], pkS[32];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c0 = g * 64 + hf * 32;
          uint32_t sr[32], dr[32];
          sm100::tmem_ld_32x32(trow + COL_S + c0, sr);
          sm100::tmem_ld_32x32(trow + COL_DP + c0, dr);
          sm100::tmem_ld_wait();
          if (hf == 1) {                               // this warp's part of S/dP is in registers
            sm100::tc_fence_before();
            __syncwarp();
            if (lane == 0) sm100::mbar_arrive(&s_free[t]);
          }
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = atc_exp2(__uint_as_float(sr[e]) * k2 - my_lse2);
            const float p1 = atc_exp2(__uint_as_float(sr[e + 1]) * k2 - my_lse2);
            const float d0 = p0 * (__uint_as_float(dr[e]) - my_d) * q.scale;
            const float d1 = p1 * (__uint_as_float(dr[e + 1]) - my_d) * q.scale;
            pkP[hf * 16 + (e >> 1)] = atc_pack2(p0, p1);
            pkS[hf * 16 + (e >> 1)] = atc_pack2(d0, d1);
          }
        }
        if (dbg2 && it == 1) g_attn_dbg[16 + t] = clock64();
        // P / dS of the previous tile must have been consumed by its three contractions
        if (t > 0) sm100::mbar_wait(&mma_done[t - 1], ph);
        else if (it > 0) sm100::mbar_wait(&mma_done[3], ph ^ 1u);
        {
          uint8_t* prow = sP + g * BLK + r * 128;
          uint8_t* srow = sdS + g * BLK + r * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint32_t off = (((uint32_t)ch ^ sw) << 4);
            *reinterpret_cast<uint4*>(prow + off) = make_uint4(pkP[4 * ch], pkP[4 * ch + 1], pkP[4 * ch + 2], pkP[4 * ch + 3]);
            *reinterpret_cast<uint4*>(srow + off) = make_uint4(pkS[4 * ch], pkS[4 * ch + 1], pkS[4 * ch + 2], pkS[4 * ch + 3]);
          }
        }
        sm100::fence_proxy_async();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&p_full[t]);
        if (dbg2 && it == 1) g_attn_dbg[20 + t] = clock64();
      }
    }
  } else {
    // -------------------------------------------------------------------------------------------------- drain
    reg_dealloc<96>();
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool dbg3 = dbg_on && warp == 12;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    auto release = [&](uint64_t* bar) {
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(bar);
    };
    const uint32_t sw = (uint32_t)(r & 7);
    // 64 fp32 columns of this thread's TMEM lane -> one 128-byte bf16 row of a swizzled [128 x 64] staging tile (the
    // layout of the operand tiles, so the TMA store uses the same kind of tensor map as the loads).  `bar` is released as
    // soon as the values are in registers.  Row-per-thread global stores of these tiles (16 bytes x 32 different lines per
    // instruction) slowed the concurrent MMAs down by 25 %: scripts/attn_timeline.py.
    auto drain64 = [&](uint32_t col, uint8_t* stage, uint64_t* bar) {
      uint32_t ra[32], rb[32];
      sm100::tmem_ld_32x32(trow + col, ra);
      sm100::tmem_ld_32x32(trow + col + 32, rb);
      sm100::tmem_ld_wait();
      if (bar) release(bar);
      if (q.dbg & 2) return;                                   // TIMING ablation: nothing leaves the registers
      uint8_t* row = stage + r * 128;
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        *reinterpret_cast<uint4*>(row + (((uint32_t)v4 ^ sw) << 4)) =
            make_uint4(atc_pack2(__uint_as_float(ra[8 * v4 + 0]), __uint_as_float(ra[8 * v4 + 1])),
                       atc_pack2(__uint_as_float(ra[8 * v4 + 2]), __uint_as_float(ra[8 * v4 + 3])),
                       atc_pack2(__uint_as_float(ra[8 * v4 + 4]), __uint_as_float(ra[8 * v4 + 5])),
                       atc_pack2(__uint_as_float(ra[8 * v4 + 6]), __uint_as_float(ra[8 * v4 + 7])));
        *reinterpret_cast<uint4*>(row + (((uint32_t)(v4 + 4) ^ sw) << 4)) =
            make_uint4(atc_pack2(__uint_as_float(rb[8 * v4 + 0]), __uint_as_float(rb[8 * v4 + 1])),
                       atc_pack2(__uint_as_float(rb[8 * v4 + 2]), __uint_as_float(rb[8 * v4 + 3])),
                       atc_pack2(__uint_as_float(rb[8 * v4 + 4]), __uint_as_float(rb[8 * v4 + 5])),
                       atc_pack2(__uint_as_float(rb[8 * v4 + 6]), __uint_as_float(rb[8 * v4 + 7])));
      }
    };
    // begin(): the previous store has finished reading the staging tiles (its issuer waited before arriving here);
    // staged(): all four drain warps have written their rows and made them visible to the TMA engine
    auto begin = [&]() { asm volatile("bar.sync 2, 128;" ::: "memory"); };
    auto staged = [&]() {
      sm100::fence_proxy_async();
      asm volatile("bar.sync 3, 128;" ::: "memory");
    };
    const bool boss = warp == 12 && lane == 0;
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const long long u = first + (long long)it * stride;
      const long long seq = u / p.heads;
      const int h = (int)(u % p.heads);
      const int c3 = (int)(seq / p.n_inner), c2 = (int)(seq % p.n_inner);
      auto wait_done = [&](uint64_t* bar) {
        sm100::mbar_wait(bar, ph);
        sm100::tc_fence_after();
      };
      wait_done(&mma_done[1]);                         // key block 0 complete
      begin();
      drain64(COL_DK, stage, nullptr);
      drain64(COL_DV, stage + BLK, &kv_drained[0]);
      staged();
      if (boss) {
        tma_store_4d(&tma_dk, stage, h * 64, 0, c2, c3);
        tma_store_4d(&tma_dv, stage + BLK, h * 64, 0, c2, c3);
        atb_bulk_commit();
        atb_bulk_wait_read();
      }
      if (dbg3 && it == 1) g_attn_dbg[24] = clock64();
      wait_done(&mma_done[2]);                         // query block 0 complete
      begin();
      drain64(COL_DQ, stage, &dq_drained[0]);
      staged();
      if (boss) {
        tma_store_4d(&tma_dq, stage, h * 64, 0, c2, c3);
        atb_bulk_commit();
        atb_bulk_wait_read();
      }
      if (dbg3 && it == 1) g_attn_dbg[25] = clock64();
      wait_done(&mma_done[3]);                         // key block 1 and query block 1 complete
      begin();
      drain64(COL_DK, stage, nullptr);
      drain64(COL_DV, stage + BLK, &kv_drained[1]);
      staged();
      if (boss) {
        tma_store_4d(&tma_dk, stage, h * 64, 128, c2, c3);
        tma_store_4d(&tma_dv, stage + BLK, h * 64, 128, c2, c3);
        atb_bulk_commit();
        atb_bulk_wait_read();
      }
      begin();
      drain64(COL_DQ + 64, stage, &dq_drained[1]);
      staged();
      if (boss) {
        tma_store_4d(&tma_dq, stage, h * 64, 128, c2, c3);
        atb_bulk_commit();
        atb_bulk_wait_read();
      }
      if (dbg3 && it == 1) g_attn_dbg[26] = clock64();
    }
    if (boss) atb_bulk_wait_all();                     // the stores have been written before the CTA exits
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<512>(tmem_base);
  }
}

static int atc_launch_bwd256(const vvae_attn_args& a, const AttnTcPlan& p, cudaStream_t s) {
  constexpr int SMEM = 14 * 16384 + 256 + 1024;
  CUtensorMap mq, mk, mv, md;
  int rc;
  if ((rc = atc_make_map(&mq, a.q, a.q_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mk, a.k, a.k_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mv, a.v, a.v_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&md, a.d_o, a.do_rs, p, 128, 1, 1))) return rc;
  CUtensorMap mdq, mdk, mdv;
  if ((rc = atc_make_map(&mdq, a.dq, a.dq_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mdk, a.dk, a.dk_rs, p, 128, 1, 1))) return rc;
  if ((rc = atc_make_map(&mdv, a.dv, a.dv_rs, p, 128, 1, 1))) return rc;
  const long long n_seq = (long long)p.n_outer * p.n_inner;
  {
    const long long total = n_seq * p.L * 8 * p.heads;
    const int blocks = (int)std::min<long long>(cdiv(total, 256), (long long)num_sms() * 16);
    launch_pdl(attn_delta_kernel, dim3(blocks), dim3(256), 0, s, (const bf16*)a.o, a.o_rs, (const bf16*)a.d_o, a.do_rs,
               a.delta, p, total);
    if ((rc = check_launch("attn_delta"))) return rc;
  }
  AttnBwd256Params q;
  q.pl = p;
  q.lse = a.lse; q.delta = a.delta;
  q.dq = (bf16*)a.dq; q.dk = (bf16*)a.dk; q.dv = (bf16*)a.dv;
  q.dq_rs = a.dq_rs; q.dk_rs = a.dk_rs; q.dv_rs = a.dv_rs;
  q.scale = a.scale;
  q.units = n_seq * p.heads;
  q.dbg = ((g_dbg[10] & 16) ? 1 : 0) | ((g_dbg[10] & 1) ? 2 : 0) | ((g_dbg[10] & 2) ? 4 : 0) | ((g_dbg[10] & 4) ? 8 : 0);
  q.dbg_cta = (int)g_dbg[0];
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd256_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("attention bwd (L=256): cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  const int grid = (int)std::min<long long>(q.units, num_sms());
  launch_pdl(attn_bwd256_sm100_kernel, dim3(grid), dim3(512), SMEM, s, mq, mk, mv, md, mdq, mdk, mdv, q);
  return check_launch("attn_bwd256_sm100");
}

// ============================================================================== forward, unmasked L = 256, persistent
// One CTA per SM, one UNIT = one (sequence, head): both 128-row query blocks against the 256 keys, K and V staged once.
//   warp 0     : TMA producer (Q as one [256 x 64] box, K, V); the next unit's Q|K are requested as soon as both score
//                tiles of the current unit are complete, V as soon as both P.V contractions are
//   warp 1     : MMA issuer: S_A = Q0.K^T, S_B = Q1.K^T (TMEM columns [0,256) and [256,512)), then O_X = P_X.V into the
//                first 64 columns of S_X once softmax group X has turned S_X into P_X
//   warps 4-7  : softmax group A (query block 0), warps 8-11: group B (query block 1): one query row per thread out of
//                TMEM (row maximum, exp2, bf16 P into the 128B-swizzled K-major operand tile), then O / rowsum staged in
//                the dead P tile and stored by TMA, log-sum-exp stored directly
// While one group runs its softmax the tensor core works for the other one, and no per-tile prologue, operand-load wait
// or row-per-thread global store is left on the critical path (one-tile-per-CTA kernel above: 10.5 k cycles per tile of
// which 2.6 k prologue + load and 1.6 k output store; two co-resident CTAs reached 8.1 k per tile per SM).
// Measured (scripts/attn_fwd256_ab.py, cold L2): 128 sequences x 8 heads 64.5 -> 46.1 us, 512 sequences 216 -> 139 us
// (318 -> 493 TFLOP/s), outputs and log-sum-exp bit-identical to the one-tile-per-CTA kernel.  Timeline
// (scripts/attn_fwd256_timeline.py): 8.0 k cycles per unit = 4.0 k per tile; per group: row maxima 1.2 k, exp + P store
// 3.9 k (MUFU: 65 536 exp2 per unit at 16 per clock = 4.1 k per unit is the floor of this phase when both groups are
// in it), wait for P.V 1.5 k (16 MMAs of N = 64), O drain + store 1.0 k, next S 0.4 k.  Issuing both S at the start
// of a unit (lock-step, vvae_debug_set(10, 32)) is 12 % slower: both groups then contend for the MUFU at once.
// Every mbarrier completes exactly once per unit: a wait's parity is the unit counter's low bit.
struct AttnFwd256Params {
  AttnTcPlan pl;
  float* lse;
  float scale;
  long long units;
  int dbg, dbg_cta;
};

__global__ void __launch_bounds__(384, 1)
attn_fwd256_sm100_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                         const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o,
                         const AttnFwd256Params q) {
  const AttnTcPlan& p = q.pl;
  constexpr int BLK = 16384;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [256 q][64] K-major SW128 (query block i at + i * BLK)
  uint8_t* sK = sQ + 2 * BLK;                      // [256 keys][64] K-major SW128
  uint8_t* sV = sK + 2 * BLK;                      // [256 keys][64] (MN-major B operand of P.V)
  uint8_t* sP = sV + 2 * BLK;                      // [2 groups][4 key blocks][128 q][64 keys] bf16, swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 8 * BLK);
  uint64_t* qk_full = bars;                        // Q|K of the unit have landed
  uint64_t* qk_free = bars + 1;                    // ... and both score tiles have been computed from them
  uint64_t* v_full = bars + 2;
  uint64_t* v_free = bars + 3;
  uint64_t* s_full = bars + 4;                     // [2] S_X is in TMEM
  uint64_t* s_free = bars + 6;                     // [2] O_X has been read out: the columns of S_X can be overwritten
  uint64_t* p_full = bars + 8;                     // [2] P_X is in shared memory
  uint64_t* o_full = bars + 10;                    // [2] O_X is in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // vvae_debug_set(10, 16): CTA vvae_debug_set(0, n) stamps clock64 in its 2nd and 3rd unit (slot = 14 * (unit - 1) + k):
  // issuer k = 0 unit start, 1 P.V_A and the next S_A issued, 2 / 3 P_A / P_B seen; softmax group A k = 4 S seen, 5 row maxima, 6 P stored,
  // 7 O seen, 8 O store issued; group B k = 9..13 likewise
  const bool dbg_cta = (q.dbg & 1) && (int)blockIdx.x == q.dbg_cta && lane == 0;
#define AF_STAMP(k) do { if (dbg_cta && (it == 1 || it == 2)) g_attn_dbg[14 * (it - 1) + (k)] = clock64(); } while (0)
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_q);
    sm100::tma_prefetch_desc(&tma_k);
    sm100::tma_prefetch_desc(&tma_v);
    sm100::tma_prefetch_desc(&tma_o);
    sm100::mbar_init(qk_full, 1);
    sm100::mbar_init(qk_free, 1);
    sm100::mbar_init(v_full, 1);
    sm100::mbar_init(v_free, 1);
    for (int x = 0; x < 2; ++x) {
      sm100::mbar_init(&s_full[x], 1);
      sm100::mbar_init(&s_free[x], 4);
      sm100::mbar_init(&p_full[x], 4);
      sm100::mbar_init(&o_full[x], 1);
    }
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<512>(tmem_slot);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const long long first = blockIdx.x, stride = gridDim.x;
  const int n_it = first < q.units ? (int)((q.units - first + stride - 1) / stride) : 0;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------------------------------------ producer
      for (int it = 0; it < n_it; ++it) {
        const long long u = first + (long long)it * stride;
        const long long seq = u / p.heads;
        const int h = (int)(u % p.heads);
        const int c3 = (int)(seq / p.n_inner), c2 = (int)(seq % p.n_inner);
        const uint32_t pp = (uint32_t)(it - 1) & 1u;
        if (it > 0) sm100::mbar_wait(qk_free, pp);
        sm100::mbar_expect_tx(qk_full, 4 * BLK);
        tma_load_4d(sQ, &tma_q, qk_full, h * 64, 0, c2, c3);
        tma_load_4d(sK, &tma_k, qk_full, h * 64, 0, c2, c3);
        if (it > 0) sm100::mbar_wait(v_free, pp);
        sm100::mbar_expect_tx(v_full, 2 * BLK);
        tma_load_4d(sV, &tma_v, v_full, h * 64, 0, c2, c3);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = sm100::make_idesc_bf16(128, 256, false, false);   // Q, K both K-major
      constexpr uint32_t idesc_o = sm100::make_idesc_bf16(128, 64, false, true);     // P K-major, V MN-major
      const uint32_t qa = sm100::smem_u32(sQ), ka = sm100::smem_u32(sK), va = sm100::smem_u32(sV), pa = sm100::smem_u32(sP);
      // Software pipeline: ... P.V_A(u), S_A(u+1), P.V_B(u), S_B(u+1) ...  Each group's chain (S -> softmax -> P.V ->
      // drain -> next S) is independent of the other's, and the alternation staggers the groups by half a period, so
      // that one group's exp phase (MUFU-bound: 8 cycles per warp instruction) runs while the other group waits for the
      // tensor core instead of both contending for the MUFU at the same time (vvae_debug_set(10, 32): lock-step order).
      auto issue_s = [&](int x) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sm100::umma_f16(tmem_base + 256 * x, sm100::make_smem_desc_sw128(qa + x * BLK + k * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        sm100::umma_commit(&s_full[x]);
      };
      auto issue_pv = [&](int x) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
          sm100::umma_f16(tmem_base + 256 * x,
                          sm100::make_smem_desc_sw128(pa + x * 4 * BLK + (k >> 2) * BLK + (k & 3) * 32, 16, 1024),
                          sm100::make_smem_desc_sw128(va + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
        sm100::umma_commit(&o_full[x]);
      };
      const bool lockstep = (q.dbg & 2) != 0;
      if (n_it > 0) {
        sm100::mbar_wait(qk_full, 0);
        sm100::tc_fence_after();
        issue_s(0);
        issue_s(1);
        sm100::umma_commit(qk_free);
      }
#pragma unroll 1
      for (int it = 0; it < n_it; ++it) {
        const uint32_t ph = (uint32_t)it & 1u;
        const bool more = it + 1 < n_it;
        AF_STAMP(0);
        sm100::mbar_wait(v_full, ph);
        sm100::mbar_wait(&p_full[0], ph);
        sm100::tc_fence_after();
        AF_STAMP(2);
        issue_pv(0);
        if (more && !lockstep) {
          sm100::mbar_wait(qk_full, ph ^ 1u);
          sm100::mbar_wait(&s_free[0], ph);
          sm100::tc_fence_after();
          issue_s(0);
        }
        AF_STAMP(1);
        sm100::mbar_wait(&p_full[1], ph);
        sm100::tc_fence_after();
        AF_STAMP(3);
        issue_pv(1);
        sm100::umma_commit(v_free);
        if (more) {
          if (lockstep) {
            sm100::mbar_wait(qk_full, ph ^ 1u);
            sm100::mbar_wait(&s_free[0], ph);
            sm100::tc_fence_after();
            issue_s(0);
          }
          sm100::mbar_wait(&s_free[1], ph);
          sm100::tc_fence_after();
          issue_s(1);
          sm100::umma_commit(qk_free);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------------------------ softmax groups A and B
    const int x = (warp - 4) >> 2;                     // group = query block
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // query row of the block == TMEM lane
    const int gtid = threadIdx.x - 128 - 128 * x;      // 0..127 inside the group
    const int sk = (quarter == 0) ? 4 + 5 * x : -100;   // stamp slots of this group's first warp
    const float k2 = q.scale * 1.4426950408889634f;
    const uint32_t trow = tmem_base + 256 * x + ((uint32_t)(quarter * 32) << 16);
    uint8_t* sPx = sP + x * 4 * BLK;
    uint8_t* prow = sPx + r * 128;
    const uint32_t sw = (uint32_t)(r & 7);
    const int bar_id = 1 + x;
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const long long u = first + (long long)it * stride;
      const long long seq = u / p.heads;
      const int h = (int)(u % p.heads);
      sm100::mbar_wait(&s_full[x], ph);
      sm100::tc_fence_after();
      if (sk > 0) AF_STAMP(sk);
      // pass 1: row maximum of the logits
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t sr[32];
        sm100::tmem_ld_32x32(trow + c0, sr);
        sm100::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
      }
      // the row maximum of s * scale: scale > 0 is not assumed
      mx = q.scale >= 0.f ? mx * q.scale : -INFINITY;
      if (q.scale < 0.f) {
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t sr[32];
          sm100::tmem_ld_32x32(trow + c0, sr);
          sm100::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]) * q.scale);
        }
      }
      if (sk > 0) AF_STAMP(sk + 1);
      // the previous unit's O tile (staged in this group's P tile) has been read by its TMA store
      if (gtid == 0) atb_bulk_wait_read();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      // pass 2: p = exp(s - max), row sum, bf16 P into the swizzled K-major operand tile
      const float mx2 = mx * 1.4426950408889634f;
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint8_t* pblk = prow + (c0 >> 6) * BLK;
        const uint32_t ch0 = (uint32_t)(c0 & 63) >> 3;
        uint32_t sr[32];
        sm100::tmem_ld_32x32(trow + c0, sr);
        sm100::tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pv[i] = atc_exp2(__uint_as_float(sr[i]) * k2 - mx2);
          sum += pv[i];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(pv[8 * j + 0], pv[8 * j + 1]);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(pv[8 * j + 2], pv[8 * j + 3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(pv[8 * j + 4], pv[8 * j + 5]);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(pv[8 * j + 6], pv[8 * j + 7]);
          pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(pblk + (((ch0 + j) ^ sw) << 4)) = pk;
        }
      }
      sm100::fence_proxy_async();      // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&p_full[x]);
      if (sk > 0) AF_STAMP(sk + 2);

      // epilogue: O / rowsum -> bf16 rows staged in the (now dead) first block of P_X, stored by TMA; log-sum-exp
      sm100::mbar_wait(&o_full[x], ph);
      sm100::tc_fence_after();
      if (sk > 0) AF_STAMP(sk + 3);
      uint32_t o0[32], o1[32];
      sm100::tmem_ld_32x32(trow, o0);
      sm100::tmem_ld_32x32(trow + 32, o1);
      sm100::tmem_ld_wait();
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&s_free[x]);
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o0[8 * j + 0]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(prow + (((uint32_t)j ^ sw) << 4)) = pk;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o1[8 * j + 0]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(prow + (((uint32_t)(j + 4) ^ sw) << 4)) = pk;
      }
      if (q.lse) q.lse[(seq * p.heads + h) * 256 + x * 128 + r] = mx + __logf(sum);
      sm100::fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (gtid == 0) {
        tma_store_4d(&tma_o, sPx, h * 64, x * 128, (int)(seq % p.n_inner), (int)(seq / p.n_inner));
        atb_bulk_commit();
      }
      if (sk > 0) AF_STAMP(sk + 4);
    }
    if (gtid == 0) atb_bulk_wait_all();                // the stores have been written before the CTA exits
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<512>(tmem_base);
  }
}

static int atc_launch_fwd256(const vvae_attn_args& a, const AttnTcPlan& p, cudaStream_t s) {
  constexpr int SMEM = 14 * 16384 + 256 + 1024;
  CUtensorMap mq, mk, mv, mo;
  int rc;
  if ((rc = atc_make_map(&mq, a.q, a.q_rs, p, 256, 1, 1))) return rc;
  if ((rc = atc_make_map(&mk, a.k, a.k_rs, p, 256, 1, 1))) return rc;
  if ((rc = atc_make_map(&mv, a.v, a.v_rs, p, 256, 1, 1))) return rc;
  if ((rc = atc_make_map(&mo, a.o, a.o_rs, p, 128, 1, 1))) return rc;
  AttnFwd256Params q;
  q.pl = p;
  q.lse = a.lse;
  q.scale = a.scale;
  q.units = (long long)p.n_outer * p.n_inner * p.heads;
  q.dbg = ((g_dbg[10] & 16) ? 1 : 0) | ((g_dbg[10] & 32) ? 2 : 0);
  q.dbg_cta = (int)g_dbg[0];
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd256_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("attention fwd (L=256): cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  const int grid = (int)std::min<long long>(q.units, num_sms());
  launch_pdl(attn_fwd256_sm100_kernel, dim3(grid), dim3(384), SMEM, s, mq, mk, mv, mo, q);
  return check_launch("attn_fwd256_sm100");
}

int attn_tc_supported(const vvae_attn_args& a) {
  AttnTcPlan p;
  if (!atc_make_plan(a, p)) return 0;
  if (((uintptr_t)a.q % 16) || ((uintptr_t)a.k % 16) || ((uintptr_t)a.v % 16) || ((uintptr_t)a.o % 16)) return 0;
  if ((a.q_rs % 8) || (a.k_rs % 8) || (a.v_rs % 8) || (a.o_rs % 8)) return 0;
  if (a.heads > 65535) return 0;
  return 1;
}

int attn_tc_bwd_supported(const vvae_attn_args& a) {
  if (!attn_tc_supported(a)) return 0;
  if (((uintptr_t)a.d_o % 16) || ((uintptr_t)a.dq % 16) || ((uintptr_t)a.dk % 16) || ((uintptr_t)a.dv % 16)) return 0;
  if ((a.do_rs % 8) || (a.dq_rs % 8) || (a.dk_rs % 8) || (a.dv_rs % 8)) return 0;
  return 1;
}

int attn_tc_fwd(const vvae_attn_args& a, cudaStream_t s) {
  AttnTcPlan p;
  if (!atc_make_plan(a, p)) {
    set_error("attention: shape not supported by the tensor-core path");
    return VVAE_ERR_UNSUPPORTED;
  }
  if (atc_len_long(a.L)) return atl_launch_fwd(a, p, s);
  const bool m = a.mask != nullptr;
  if (p.G > 1) return m ? atc_launch_fwd<128, true, true>(a, p, s) : atc_launch_fwd<128, true, false>(a, p, s);
  if (p.NK == 128) return m ? atc_launch_fwd<128, false, true>(a, p, s) : atc_launch_fwd<128, false, false>(a, p, s);
  // vvae_debug_set(18, 1): the one-tile-per-CTA kernel for unmasked L = 256 too
  if (!m && !g_dbg[18]) return atc_launch_fwd256(a, p, s);
  return m ? atc_launch_fwd<256, false, true>(a, p, s) : atc_launch_fwd<256, false, false>(a, p, s);
}

int attn_tc_bwd(const vvae_attn_args& a, cudaStream_t s) {
  AttnTcPlan p;
  if (!atc_make_plan(a, p)) {
    set_error("attention bwd: shape not supported by the tensor-core path");
    return VVAE_ERR_UNSUPPORTED;
  }
  if (atc_len_long(a.L)) return atl_launch_bwd(a, p, s);
  const bool m = a.mask != nullptr;
  if (p.G > 1) return m ? atc_launch_bwd<1, true, true>(a, p, s) : atc_launch_bwd<1, true, false>(a, p, s);
  if (p.NK == 128) return m ? atc_launch_bwd<1, false, true>(a, p, s) : atc_launch_bwd<1, false, false>(a, p, s);
  // vvae_debug_set(16, 1): the one-unit-per-CTA kernel for unmasked L = 256 too
  if (!m && a.delta && !g_dbg[16] && p.ts_l * 0 == 0) return atc_launch_bwd256(a, p, s);
  return m ? atc_launch_bwd<2, false, true>(a, p, s) : atc_launch_bwd<2, false, false>(a, p, s);
}

}  // namespace vvae
