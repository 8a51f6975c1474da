// bf16 GEMM on the 5th-generation tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (accumulators in
// TMEM, double buffered) -> tcgen05.ld epilogue.  Persistent, warp specialised:
//   warp 0   : TMA producer (one lane)
//   warp 1   : TMEM allocator + MMA issuer (one lane)
//   warps 2-9: epilogue: two warps per TMEM lane quarter (warp_id % 4), each draining one 32-column half of every
//              64-column chunk.  The epilogue is a latency chain (tcgen05.ld -> convert -> staging -> TMA store), and with
//              four warps it took longer per tile than the tile's MMAs: with NO operand loads at all the kernel still ran
//              at only 1.2 PFLOP/s (profiles/r02p_gemm_ablate.jsonl) -- the accumulator drain, not L2, was the limiter.
//   warps 10-11: (weight-gradient mode) column sums of B
// Both operands may be K-major (contraction index contiguous in memory) or MN-major (row/column index contiguous),
// so forward (X . W, W is Flax (in,out) = MN-major B), dgrad (dY . W^T, K-major B) and wgrad (X^T . dY, both MN-major,
// split-K with fp32 atomics) all read the tensors where they lie -- no transposed copies are ever materialised.
#include <atomic>
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

namespace vvae {

// ---- debug / tuning knobs (vvae_debug_set) ----
long long g_dbg[32] = {0};   // (declared in common.cuh)

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

int encode_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return VVAE_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu stride %llu box %u,%u base %p", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0, base);
    return VVAE_ERR_CUDA;
  }
  return VVAE_OK;
}

int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                        uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  uint64_t dims[2] = {inner, outer}, str[1] = {outer_stride_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return encode_tmap_nd_bf16(out, base, 2, dims, str, box, swizzle_bytes);
}

// debug counters written by CTA 0's MMA-issuing thread when vvae_debug_set(10, ...) has bit 16 set:
// [0] clock64 ticks between its first and last MMA issue, [1] globaltimer ns for the same span, [2] MMAs issued
__device__ unsigned long long g_gemm_dbg[4];

// ------------------------------------------------------------------ device kernel
constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB per CTA
constexpr int STG_BYTES = 128 * 128;        // one epilogue staging buffer: 128 rows x 64 bf16, 128B-swizzled
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;
constexpr int MODE_QKN = VVAE_EPI_QKNORM_ROPE;   // 4: QKV projection with per-head LayerNorm + RoPE of q|k in the epilogue
constexpr int MODE_BSUM = 100;                // internal: weight-gradient GEMM that also sums the columns of its B operand // shared::cluster address of the same offset in the even CTA of a pair

struct Sm100Params {
  int M, N;                // output extent
  int m_tiles, n_tiles, splits;   // m_tiles counts (128*CG)-row tiles
  int kb_total, kb_per_split;     // K blocks of 64
  void* C; long long ldc; int out_f32;
  const float* bias;
  int mode;
  const bf16* aux_in; long long ld_ai;
  bf16* aux_out; long long ld_ao;
  int atomic;
  int tma_epi;             // bf16 output: epilogue goes TMEM -> registers -> swizzled smem -> TMA store
  float* bsum;             // MODE_BSUM: += column sums of op(B) over K (a Linear's bias gradient, fused into its wgrad)
  // descriptor encodings (bytes); overridable through vvae_debug_set for bring-up
  uint32_t a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv;
  // Partial last wave: the last `tiles_mn % clusters` output tiles are cut into tail_s column slices of BN / tail_s, so
  // that the clusters that would idle through the last wave share its work (tail_s = 1: off).  Tiles [0, full_tiles) are
  // whole, tile full_tiles + s*tail_s + q is slice q of whole-tile index full_tiles + s.
  int full_tiles, tail_s, total_tiles;
  uint32_t idesc_tail;     // instruction descriptor with N = BN / tail_s
  int dbg;                 // vvae_debug_set(10): TIMING ablations, wrong results (1: skip the A-tile TMA loads, 2: skip B)
  // MODE_QKN
  const float* qk_qs; const float* qk_ks; const bf16* rope_cos; const bf16* rope_sin;
  long long pos_div; int pos_mod; int qk_cols; float qk_eps;
};

// HEAVY epilogues (residual / dSiLU read an aux tile, SiLU writes two tiles) get 4 extra staging buffers: a 4-deep ring of
// TMA-prefetched aux-input chunks (one whole tile ahead: TMA latency under load is several microseconds), or the second
// output's double buffer.
template <int BN, int CG, bool HEAVY> struct StageCfg {
  static constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;          // per CTA
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // per CTA
  static constexpr int EPI_BYTES = (HEAVY ? 6 : 2) * STG_BYTES;
  static constexpr int BAR_BYTES = 320 + 256 * 4 + 128 * 4;   // barriers + the current tile's bias slice + q|k LayerNorm scales
  static constexpr int BUDGET = 232448 - 1024 - BAR_BYTES - EPI_BYTES;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024 /*align slack*/;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the barrier at the same offset in the pair's even CTA
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  if constexpr (CG == 1) {
    sm100::tma_load_2d(smem_dst, m, bar, c0, c1);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(sm100::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(sm100::smem_u32(bar) & PEER_MASK),
        "r"(c0), "r"(c1)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_f16_cg(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1) {
    sm100::umma_f16(tmem_d, desc_a, desc_b, idesc, acc);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
  }
}
// arrive (once the issuing thread's earlier MMAs have completed) on the barrier at this offset in every CTA of the pair
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
  if constexpr (CG == 1) {
    sm100::umma_commit(bar);
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(sm100::smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  if constexpr (CG == 1) {
    sm100::mbar_arrive(bar);
  } else {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(sm100::smem_u32(bar) & PEER_MASK) : "memory");
  }
}
template <int CG, int NCOLS>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* slot) {
  if constexpr (CG == 1) {
    sm100::tmem_alloc<NCOLS>(slot);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm100::smem_u32(slot)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG, int NCOLS>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr) {
  if constexpr (CG == 1) sm100::tmem_dealloc<NCOLS>(taddr);
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(sm100::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp32 / atomic outputs (weight gradients): one thread owns row m, 32 consecutive columns starting at n0
__device__ __forceinline__ void epilogue_store_f32(const Sm100Params& p, long long m, int n0, const uint32_t (&r)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const bool full = (n0 + 32 <= p.N);
  if (p.bias) {
    for (int j = 0; j < 32; ++j)
      if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
  }
  float* c = reinterpret_cast<float*>(p.C) + m * p.ldc + n0;
  if (p.atomic) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                     "f"(v[j + 3])
                     : "memory");
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) atomicAdd(c + j, v[j]);
    }
  } else if (full) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
    for (int j = 0; j < 32; ++j)
      if (n0 + j < p.N) c[j] = v[j];
  }
}

__device__ __forceinline__ float fast_sigmoid(float x) {   // 0.5*tanh(x/2)+0.5, one MUFU; |err| < 3e-4 (bf16 outputs)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float fast_silu(float x) { return x * fast_sigmoid(x); }
__device__ __forceinline__ float fast_dsilu(float x) {
  const float s = fast_sigmoid(x);
  return s * fmaf(x, 1.f - s, 1.f);
}

// tile index -> (row block, first column, columns) of the output tile (see Sm100Params::tail_s)
struct TileGeo { int mblk, n0, ncols; };
template <int BN>
__device__ __forceinline__ TileGeo tile_geo(const Sm100Params& p, int tile, int tiles_mn) {
  TileGeo g;
  if (p.tail_s == 1 || tile < p.full_tiles) {
    const int mn = tile % tiles_mn;
    g.mblk = mn / p.n_tiles;
    g.n0 = (mn % p.n_tiles) * BN;
    g.ncols = BN;
  } else {
    const int t2 = tile - p.full_tiles;
    const int mn = p.full_tiles + t2 / p.tail_s, q = t2 % p.tail_s;
    g.ncols = BN / p.tail_s;
    g.mblk = mn / p.n_tiles;
    g.n0 = (mn % p.n_tiles) * BN + q * g.ncols;
  }
  return g;
}

template <int BN, int CG, bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__(MODE == MODE_QKN ? 192 : 384, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_ai,
                  const __grid_constant__ CUtensorMap tma_ao, Sm100Params p) {
  constexpr bool HEAVY = MODE != VVAE_EPI_NONE && MODE != MODE_BSUM;   // fused epilogues are compiled in per mode
  constexpr bool BSUM = MODE == MODE_BSUM;        // two extra warps sum the B tiles' columns after the MMAs consumed them
  using Cfg = StageCfg<BN, CG, HEAVY>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NCH = BN / 64;                 // 64-column epilogue chunks per tile
  constexpr int AUXR = 4;                      // aux-input ring depth (HEAVY only)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* stg_out = smem + STAGES * Cfg::STAGE_BYTES;        // [2][128 x 128 B]
  uint8_t* stg_aux = stg_out + 2 * STG_BYTES;                 // HEAVY: [4][128 x 128 B] aux ring | second output [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* aux_full = bars + 2 * STAGES + 4;   // [4]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);
  uint64_t* mma_done = bars + 2 * STAGES + 9;   // [STAGES] (BSUM): the MMAs reading stage s have completed
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 320);   // [256]
  float* s_qks = s_bias + 256;                                                         // [2][64] (MODE_QKN)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  pdl_launch_dependents();   // every CTA of this grid is resident: the next kernel may move in as SMs free up

  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_a);
    sm100::tma_prefetch_desc(&tma_b);
    if (p.tma_epi) sm100::tma_prefetch_desc(&tma_c);
    for (int i = 0; i < STAGES; ++i) {
      sm100::mbar_init(&full_bar[i], 1);
      sm100::mbar_init(&empty_bar[i], BSUM ? 2 : 1);   // BSUM: the two column-sum warps release the slot
      if (BSUM) sm100::mbar_init(&mma_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&tmem_full[i], 1);
      sm100::mbar_init(&tmem_empty[i], (MODE == MODE_QKN ? 4 : 8) * CG);
    }
    for (int i = 0; i < AUXR; ++i) sm100::mbar_init(&aux_full[i], 1);
    sm100::fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_cg<CG, Cfg::TMEM_COLS>(tmem_base_slot);
  sm100::tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();                // prologue done; from here on global memory of the previous kernel is read / written

  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_tiles = p.total_tiles;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair load their own halves) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int split = p.tail_s == 1 ? tile / tiles_mn : 0;
        const TileGeo geo = tile_geo<BN>(p, tile, tiles_mn);
        const int m0 = geo.mblk * (BM * CG) + (int)cta_rank * BM;
        // (a column slice loads the whole B box from its first row: the MMA reads the first ncols / CG rows of it)
        const int n0 = geo.n0 + (int)cta_rank * (geo.ncols / CG);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (p.dbg & 4) continue;                 // ablation: the issuer does not wait for operands at all
          sm100::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (cta_rank == 0)
            sm100::mbar_expect_tx(&full_bar[stage], (((p.dbg & 1) ? 0 : A_STAGE_BYTES) + ((p.dbg & 2) ? 0 : Cfg::B_STAGE_BYTES)) * CG);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
          const int k0 = kb * BK;
          if (p.dbg & 1) {
          } else if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg<CG>(sa + j * 8192, &tma_a, &full_bar[stage], m0 + 64 * j, k0);
          } else {
            tma_load_2d_cg<CG>(sa, &tma_a, &full_bar[stage], k0, m0);
          }
          if (p.dbg & 2) {
          } else if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / CG / 64; ++j)
              tma_load_2d_cg<CG>(sb + j * 8192, &tma_b, &full_bar[stage], n0 + 64 * j, k0);
          } else {
            tma_load_2d_cg<CG>(sb, &tma_b, &full_bar[stage], k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the pair's even CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = sm100::make_idesc_bf16(BM * CG, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      unsigned long long dbg_c0 = 0, dbg_t0 = 0, dbg_n = 0;
      if (p.dbg & 16) {
        dbg_c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
      }
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int split = p.tail_s == 1 ? tile / tiles_mn : 0;
        const uint32_t idesc_t = (p.tail_s == 1 || tile < p.full_tiles) ? idesc : p.idesc_tail;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        if (!(p.dbg & 8)) sm100::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        sm100::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        // The issuing thread sits on the tensor pipe's critical path (the MMA queue is shallow: 179 cycles per 128-cycle
        // MMA in the full kernel against 165 when the thread never waits, profiles/r02z_gemm_mma_rate.jsonl), so the
        // poll of the NEXT stage's full barrier (~90 cycles even when the phase has completed) is started before the
        // last MMA of the current stage is issued and only re-checked afterwards.
        bool ready = (p.dbg & 4) ? true : sm100::mbar_test_wait(&full_bar[stage], phase);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!ready) sm100::mbar_wait(&full_bar[stage], phase);
          sm100::tc_fence_after();
          dbg_n += BK / UMMA_K;
          const uint32_t a_addr = sm100::smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = sm100::smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES);
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = sm100::make_smem_desc_sw128(a_addr + k * p.a_kadv, p.a_lbo, p.a_sbo);
            const uint64_t db = sm100::make_smem_desc_sw128(b_addr + k * p.b_kadv, p.b_lbo, p.b_sbo);
            if (k == BK / UMMA_K - 1)   // (also valid across tiles: the stage ring does not care about tile borders)
              ready = (p.dbg & 4) ? true : sm100::mbar_test_wait(&full_bar[nstage], nphase);
            umma_f16_cg<CG>(d_tmem, da, db, idesc_t, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_cg<CG>(BSUM ? &mma_done[stage] : &empty_bar[stage]);  // the MMAs have read the slot (both CTAs)
          stage = nstage;
          phase = nphase;
        }
        umma_commit_cg<CG>(&tmem_full[acc]);      // accumulator complete (signals both CTAs' epilogues)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if ((p.dbg & 16) && blockIdx.x == 0) {
        // wait for the last MMA to finish: one more commit on a scratch use of tmem_full is not available, so time the
        // issue span only (the queue is a few MMAs deep: the error is a few hundred cycles over ~10^5)
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        g_gemm_dbg[0] = clock64() - dbg_c0;
        g_gemm_dbg[1] = t1 - dbg_t0;
        g_gemm_dbg[2] = dbg_n;
      }
    }
  } else if (warp >= 10) {
    // ===================== column sums of B (bias gradient), after the tensor core is done with each stage ==========
    if constexpr (BSUM) {
      constexpr int NCH_B = BN / CG / 64;          // 64-column chunks of this CTA's B slice
      const int cw = warp - 10;                    // chunks cw, cw+2, ...
      const int cc = lane & 7, rg = lane >> 3;     // 16-byte column chunk, row group
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int split = tile / tiles_mn, mn = tile % tiles_mn;
        const bool mine = (mn / p.n_tiles) == 0;   // only the first row of output tiles contributes (B is shared by all)
        const int n0 = (mn % p.n_tiles) * BN + (int)cta_rank * (BN / CG);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        float acc[(NCH_B + 1) / 2][8];
#pragma unroll
        for (int j = 0; j < (NCH_B + 1) / 2; ++j)
#pragma unroll
          for (int t = 0; t < 8; ++t) acc[j][t] = 0.f;
        for (int kb = kb0; kb < kb1; ++kb) {
          sm100::mbar_wait(&mma_done[stage], phase);
          if (mine) {
            const uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < (NCH_B + 1) / 2; ++j) {
              const int ch = cw + 2 * j;
              if (ch < NCH_B) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int r = rg + 4 * i;
                  const uint4 v = *reinterpret_cast<const uint4*>(sb + ch * 8192 + r * 128 + (((uint32_t)cc ^ (uint32_t)(r & 7)) << 4));
                  acc[j][0] += __uint_as_float(v.x << 16); acc[j][1] += __uint_as_float(v.x & 0xffff0000u);
                  acc[j][2] += __uint_as_float(v.y << 16); acc[j][3] += __uint_as_float(v.y & 0xffff0000u);
                  acc[j][4] += __uint_as_float(v.z << 16); acc[j][5] += __uint_as_float(v.z & 0xffff0000u);
                  acc[j][6] += __uint_as_float(v.w << 16); acc[j][7] += __uint_as_float(v.w & 0xffff0000u);
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (mine) {
#pragma unroll
          for (int j = 0; j < (NCH_B + 1) / 2; ++j) {
            const int ch = cw + 2 * j;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              float v = acc[j][t];
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              const int n = n0 + ch * 64 + cc * 8 + t;
              if (rg == 0 && ch < NCH_B && n < p.N) atomicAdd(p.bsum + n, v);
            }
          }
        }
      }
    }
  } else if constexpr (MODE == MODE_QKN) {
    // ===================== epilogue of the QKV projection (train/layers.py:160-166) =====================
    // Four warps, one thread per accumulator row; a 64-column chunk is ONE head (hd = 64), so the per-head LayerNorm
    // (fp32 statistics of the bf16-rounded projection, fast variance, scale only) and RoPE (packed bf16 arithmetic, the
    // rotate_half partner 32 columns away) are thread-local.  Output 1 (C): q|k|v as projected (saved for backward,
    // v feeds attention); output 2 (aux_out): rope(LN(q)) | rope(LN(k)).  Same arithmetic, instruction for instruction,
    // as qknorm_rope_fwd_hd64_kernel -- which this makes unnecessary: one read of q|k (67 MB per attention) less.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int etid = threadIdx.x - 64;           // 0..127
    const uint32_t sw = (uint32_t)(r & 7);
    auto qkn_bar = []() { asm volatile("bar.sync 2, 128;" ::: "memory"); };
    for (int j = etid; j < 128; j += 128) s_qks[j] = j < 64 ? __ldg(p.qk_qs + j) : __ldg(p.qk_ks + j - 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t g = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int mn = tile % tiles_mn;
      const int m0 = (mn / p.n_tiles) * (BM * CG) + (int)cta_rank * BM, n0 = (mn % p.n_tiles) * BN;
      for (int j = etid; j < BN; j += 128) s_bias[j] = (p.bias && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
      qkn_bar();
      const long long grow = min((long long)m0 + r, (long long)p.M - 1);
      const int pos = (int)(((unsigned long long)grow / (unsigned long long)p.pos_div) % (unsigned)p.pos_mod);
      const uint4* cosp = reinterpret_cast<const uint4*>(p.rope_cos + (long long)pos * 64);
      const uint4* sinp = reinterpret_cast<const uint4*>(p.rope_sin + (long long)pos * 64);
      sm100::mbar_wait(&tmem_full[acc], acc_phase);
      sm100::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c, ++g) {
        const int nc = n0 + 64 * c;
        const uint32_t b = g & 1;
        const bool is_qk = nc < p.qk_cols;                      // warp-uniform
        uint32_t pk[32], y[32];
        {
          uint32_t rr0[32], rr1[32];
          sm100::tmem_ld_32x32(taddr + c * 64, rr0);
          sm100::tmem_ld_32x32(taddr + c * 64 + 32, rr1);
          sm100::tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 64);
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 bb = b4[j >> 2];
            const uint32_t* src = j < 32 ? rr0 + j : rr1 + (j - 32);
            pk[j >> 1] = pack_bf16x2(__uint_as_float(src[0]) + bb.x, __uint_as_float(src[1]) + bb.y);
            pk[(j >> 1) + 1] = pack_bf16x2(__uint_as_float(src[2]) + bb.z, __uint_as_float(src[3]) + bb.w);
          }
        }
        if (is_qk) {
          float s = 0.f, s2 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f0 = __uint_as_float(pk[j] << 16), f1 = __uint_as_float(pk[j] & 0xffff0000u);
            s += f0 + f1;
            s2 = fmaf(f0, f0, fmaf(f1, f1, s2));
          }
          const float mu = s * (1.f / 64.f);
          const float rs = rsqrtf(fmaxf(s2 * (1.f / 64.f) - mu * mu, 0.f) + p.qk_eps);
          const float* sc = s_qks + (nc < (p.qk_cols >> 1) ? 0 : 64);
          uint32_t xn[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f0 = __uint_as_float(pk[j] << 16), f1 = __uint_as_float(pk[j] & 0xffff0000u);
            xn[j] = pack_bf16x2((f0 - mu) * (rs * sc[2 * j]), (f1 - mu) * (rs * sc[2 * j + 1]));
          }
          // y = x*cos + rotate_half(x)*sin in bf16 (each product and the sum rounded, as the reference's bf16 ops do)
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4) {
            const uint4 cv = __ldg(cosp + v4), sv = __ldg(sinp + v4);
            const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, sw4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int j = v4 * 4 + t;                                   // bf16 pair index: elements 2j, 2j+1
              const uint32_t rot = j < 16 ? (xn[j + 16] ^ 0x80008000u) : xn[j - 16];
              const __nv_bfloat162 a2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&xn[j]),
                                                *reinterpret_cast<const __nv_bfloat162*>(&cw[t]));
              const __nv_bfloat162 b2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&rot),
                                                *reinterpret_cast<const __nv_bfloat162*>(&sw4[t]));
              const __nv_bfloat162 y2 = __hadd2(a2, b2);
              y[j] = *reinterpret_cast<const uint32_t*>(&y2);
            }
          }
        }
        if (etid == 0) bulk_wait_read<1>();
        qkn_bar();
        uint8_t* orow = stg_out + b * STG_BYTES + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<uint4*>(orow + (((uint32_t)ch ^ sw) << 4)) =
              make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        if (is_qk) {
          uint8_t* prow = stg_aux + b * STG_BYTES + r * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(prow + (((uint32_t)ch ^ sw) << 4)) =
                make_uint4(y[4 * ch], y[4 * ch + 1], y[4 * ch + 2], y[4 * ch + 3]);
        }
        sm100::fence_proxy_async();
        qkn_bar();
        if (etid == 0) {
          if (nc < p.N && m0 < p.M) {
            tma_store_2d(&tma_c, stg_out + b * STG_BYTES, nc, m0);
            if (is_qk) tma_store_2d(&tma_ao, stg_aux + b * STG_BYTES, nc, m0);
          }
          bulk_commit();
        }
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader<CG>(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (etid == 0) bulk_wait_all();
  } else {
    // ===================== epilogue (each CTA drains its own 128 accumulator rows) =====================
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;              // which 32-column half of each 64-column chunk this warp drains
    const int r = quarter * 32 + lane;           // row inside the CTA's 128-row tile == TMEM lane
    const int etid = threadIdx.x - 64;           // 0..255
    const uint32_t sw = (uint32_t)(r & 7);
    constexpr bool use_aux = MODE == VVAE_EPI_RESIDUAL || MODE == VVAE_EPI_DSILU;
    const bool two_out = MODE == VVAE_EPI_SILU && p.aux_out != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t g = 0;                               // running chunk counter (staging-buffer / aux-barrier parity)
    // aux-input prefetch: chunk sequence of this CTA is (tile, c) in order; keep two chunks in flight
    int pf_tile = cluster_id, pf_c = 0;
    uint32_t pf_g = 0;
    auto prefetch_aux = [&]() {
      if (pf_tile >= total_tiles) return;
      const TileGeo pg = tile_geo<BN>(p, pf_tile, tiles_mn);
      const int m0 = pg.mblk * (BM * CG) + (int)cta_rank * BM;
      const uint32_t b = pf_g & (AUXR - 1);
      sm100::mbar_expect_tx(&aux_full[b], STG_BYTES);
      sm100::tma_load_2d(stg_aux + b * STG_BYTES, &tma_ai, &aux_full[b], pg.n0 + 64 * pf_c, m0);
      ++pf_g;
      if (++pf_c == pg.ncols / 64) { pf_c = 0; pf_tile += num_clusters; }
    };
    if (use_aux && etid == 0) {
      for (int i = 0; i < AUXR; ++i) prefetch_aux();
    }

    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const TileGeo geo = tile_geo<BN>(p, tile, tiles_mn);
      const int m0 = geo.mblk * (BM * CG) + (int)cta_rank * BM, n0 = geo.n0;
      const int nch = p.tma_epi ? geo.ncols / 64 : NCH;   // 64-column chunks of this tile (column slices: fewer)
      float* bias_t = s_bias;   // single buffer: every reader of the previous tile's slice has passed that tile's last barrier
      if (p.tma_epi && p.bias) {                 // this tile's bias slice -> smem (read back as broadcasts)
        for (int j = etid; j < BN; j += 256) bias_t[j] = (n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
        epi_bar();
      }
      if (p.dbg & 8) {                           // ablation: the accumulators are never drained (nor waited for)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      sm100::mbar_wait(&tmem_full[acc], acc_phase);
      sm100::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
      if (!p.tma_epi) {
        const long long m = (long long)m0 + r;
#pragma unroll 1
        for (int c = hf; c < BN / 32; c += 2) {
          if (n0 + c * 32 >= p.N) break;  // warp-uniform
          uint32_t rr[32];
          sm100::tmem_ld_32x32(taddr + c * 32, rr);
          sm100::tmem_ld_wait();
          if (m < p.M) epilogue_store_f32(p, m, n0 + c * 32, rr);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < nch; ++c, ++g) {
          const int nc = n0 + 64 * c;
          const uint32_t b = g & 1;
          uint4 ax[4];
          if constexpr (use_aux) {
            const uint32_t ab = g & (AUXR - 1);
            sm100::mbar_wait(&aux_full[ab], (g >> 2) & 1);
            const uint8_t* arow = stg_aux + ab * STG_BYTES + r * 128;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              ax[ch] = *reinterpret_cast<const uint4*>(arow + (((uint32_t)(hf * 4 + ch) ^ sw) << 4));
          }
          uint32_t pk[16], pk2[16];
          uint32_t rr[32];
          sm100::tmem_ld_32x32(taddr + c * 64 + hf * 32, rr);
          sm100::tmem_ld_wait();
          {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
            if (p.bias) {
              const float4* b4 = reinterpret_cast<const float4*>(bias_t + c * 64 + hf * 32);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = b4[j >> 2];
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
              }
            }
            if constexpr (MODE == VVAE_EPI_SILU) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const uint32_t pre = pack_bf16x2(v[j], v[j + 1]);
                pk2[j >> 1] = pre;
                const float x0 = __uint_as_float(pre << 16), x1 = __uint_as_float(pre & 0xffff0000u);
                pk[j >> 1] = pack_bf16x2(fast_silu(x0), fast_silu(x1));
              }
            } else {
              if constexpr (use_aux) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const uint32_t w = (&ax[j >> 3].x)[(j >> 1) & 3];
                  const float a0 = __uint_as_float(w << 16), a1 = __uint_as_float(w & 0xffff0000u);
                  if constexpr (MODE == VVAE_EPI_RESIDUAL) { v[j] += a0; v[j + 1] += a1; }
                  else { v[j] *= fast_dsilu(a0); v[j + 1] *= fast_dsilu(a1); }
                }
              }
#pragma unroll
              for (int j = 0; j < 32; j += 2) pk[j >> 1] = pack_bf16x2(v[j], v[j + 1]);
            }
          }
          // staging buffer(s) must have been read out by the TMA store that last used them
          if (etid == 0) bulk_wait_read<1>();                 // the group that last used staging buffer b has been read out
          epi_bar();
          if (use_aux && etid == 0) prefetch_aux();          // every thread has consumed this chunk's aux buffer
          uint8_t* orow = stg_out + b * STG_BYTES + r * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<uint4*>(orow + (((uint32_t)(hf * 4 + ch) ^ sw) << 4)) =
                make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
          if (two_out) {
            uint8_t* prow = stg_aux + b * STG_BYTES + r * 128;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              *reinterpret_cast<uint4*>(prow + (((uint32_t)(hf * 4 + ch) ^ sw) << 4)) =
                  make_uint4(pk2[4 * ch], pk2[4 * ch + 1], pk2[4 * ch + 2], pk2[4 * ch + 3]);
          }
          sm100::fence_proxy_async();
          epi_bar();
          if (etid == 0) {
            if (nc < p.N && m0 < p.M) {
              tma_store_2d(&tma_c, stg_out + b * STG_BYTES, nc, m0);
              if (two_out) tma_store_2d(&tma_ao, stg_aux + b * STG_BYTES, nc, m0);
            }
            bulk_commit();   // one (possibly empty) group per chunk keeps the wait_group accounting uniform
          }
        }
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader<CG>(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_epi && etid == 0) bulk_wait_all();
  }

  sm100::tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    tmem_dealloc_cg<CG, Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ host launch
bool sm100_gemm_supported(const vvae_gemm_args& a) {
  if (a.dtype != VVAE_BF16) return false;
  // (weight gradients of narrow Linears: M = 96 rows of a 128-row tile; the boxes beyond M are zero-filled by TMA)
  if (a.M < (a.transA && a.out_dtype == VVAE_F32 ? 64 : 128) || a.N < 64 || a.K < 64) return false;
  if (a.N % 8 != 0) return false;
  if (a.lda % 8 || a.ldb % 8) return false;
  if (((uintptr_t)a.A % 16) || ((uintptr_t)a.B % 16) || ((uintptr_t)a.C % 16)) return false;
  if (a.out_dtype == VVAE_BF16 && (a.ldc % 8)) return false;
  if (a.out_dtype == VVAE_F32 && (a.ldc % 4)) return false;
  if (a.aux_in && ((a.ld_aux_in % 8) || ((uintptr_t)a.aux_in % 16))) return false;
  if (a.aux_out && ((a.ld_aux_out % 8) || ((uintptr_t)a.aux_out % 16))) return false;
  if (a.bias && ((uintptr_t)a.bias % 16)) return false;
  if (a.accumulate && a.out_dtype != VVAE_F32) return false;
  if (a.out_dtype == VVAE_F32 && a.epilogue != VVAE_EPI_NONE) return false;   // fused epilogues are bf16-out only
  if (a.transA && a.epilogue != VVAE_EPI_NONE) return false;
  if (a.epilogue == VVAE_EPI_QKNORM_ROPE) {   // one 64-column epilogue chunk must be one head; tables in bf16, 16-byte rows
    if (a.qk_hd != 64 || a.transB || !a.aux_out || a.N != 3 * a.qk_heads * 64 || a.M < 256) return false;
    if (((uintptr_t)a.rope_cos % 16) || ((uintptr_t)a.rope_sin % 16) || a.rope_pos_div <= 0 || a.rope_pos_mod <= 0) return false;
  }
  // MN-major operands are fetched in 64-wide boxes along M / N
  if (a.transA && (a.M % 8)) return false;
  return true;
}

template <int BN, int CG, bool A_MN, bool B_MN, int MODE>
static int launch_sm100(const vvae_gemm_args& a, cudaStream_t s) {
  using Cfg = StageCfg<BN, CG, MODE != VVAE_EPI_NONE && MODE != MODE_BSUM>;   // must match the kernel's HEAVY
  static_assert(Cfg::STAGES >= 3, "pipeline too shallow");
  CUtensorMap ta, tb, tc, tai, tao;
  int rc;
  // A: op(A)[m,k].  K-major: memory [M rows][K cols];  MN-major (transA): memory [K rows][M cols].
  if (A_MN) rc = encode_tmap_2d_bf16(&ta, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda * 2, 64, BK, 128);
  else      rc = encode_tmap_2d_bf16(&ta, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda * 2, BK, BM, 128);
  if (rc) return rc;
  // B: op(B)[k,n].  K-major (transB): memory [N rows][K cols];  MN-major: memory [K rows][N cols].
  if (B_MN) rc = encode_tmap_2d_bf16(&tb, a.B, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb * 2, 64, BK, 128);
  else      rc = encode_tmap_2d_bf16(&tb, a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb * 2, BK, BN / CG, 128);
  if (rc) return rc;

  Sm100Params p;
  p.M = a.M; p.N = a.N;
  p.m_tiles = (int)cdiv(a.M, BM * CG);
  p.n_tiles = (int)cdiv(a.N, BN);
  p.kb_total = (int)cdiv(a.K, BK);
  const int n_clusters = num_sms() / CG;
  int splits = 1;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  if (a.accumulate && tiles_mn < n_clusters) {  // weight gradients: few output tiles, very long K -> split K over the SMs
    splits = std::max(1, std::min(n_clusters / tiles_mn, p.kb_total / 8));
  }
  p.kb_per_split = (int)cdiv(p.kb_total, splits);
  p.splits = (int)cdiv(p.kb_total, p.kb_per_split);
  p.C = a.C; p.ldc = a.ldc; p.out_f32 = (a.out_dtype == VVAE_F32);
  p.bias = a.bias; p.mode = a.epilogue;
  p.aux_in = (const bf16*)a.aux_in; p.ld_ai = a.ld_aux_in;
  p.aux_out = (bf16*)a.aux_out; p.ld_ao = a.ld_aux_out;
  p.atomic = (a.accumulate || p.splits > 1) ? 1 : 0;
  p.tma_epi = p.out_f32 ? 0 : 1;
  p.bsum = a.bsum_accum;
  p.dbg = (int)g_dbg[10];
  p.full_tiles = tiles_mn; p.tail_s = 1; p.total_tiles = tiles_mn * p.splits; p.idesc_tail = 0;
  if (CG == 2 && BN == 256 && p.splits == 1 && p.tma_epi && MODE != MODE_BSUM && MODE != MODE_QKN &&
      tiles_mn > n_clusters && !g_dbg[17]) {   // vvae_debug_set(17, 1): whole tiles only
    // relative cost of one 256 x (256/s) slice: its MMAs fetch the same A operand for fewer columns
    const double cost[5] = {0, 1.0, 0.6, 0, 0.4};
    const int R = tiles_mn % n_clusters;
    int best = 1;
    double best_t = 1.0;
    if (R) {
      for (int sl = 2; sl <= 4; sl *= 2) {
        if (sl == 4 && B_MN) continue;           // an MN-major B slice must be a whole 64-column box per CTA
        const double t = (double)cdiv((long long)R * sl, n_clusters) * cost[sl];
        if (t < best_t - 1e-9) { best_t = t; best = sl; }
      }
    }
    if (g_dbg[17] > 1 && R) best = (int)g_dbg[17] <= 4 && !((int)g_dbg[17] == 4 && B_MN) ? (int)g_dbg[17] : best;
    if (best > 1) {
      p.tail_s = best;
      p.full_tiles = tiles_mn - R;
      p.total_tiles = p.full_tiles + R * best;
      p.idesc_tail = sm100::make_idesc_bf16(BM * CG, BN / best, A_MN, B_MN);
    }
  }
  p.qk_qs = a.qk_q_scale; p.qk_ks = a.qk_k_scale;
  p.rope_cos = (const bf16*)a.rope_cos; p.rope_sin = (const bf16*)a.rope_sin;
  p.pos_div = a.rope_pos_div > 0 ? a.rope_pos_div : 1; p.pos_mod = a.rope_pos_mod > 0 ? a.rope_pos_mod : 1;
  p.qk_cols = 2 * a.qk_heads * a.qk_hd; p.qk_eps = a.qk_eps;
  tc = ta; tai = ta; tao = ta;   // placeholders for maps a mode does not use
  if (p.tma_epi) {
    if ((rc = encode_tmap_2d_bf16(&tc, a.C, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldc * 2, 64, 128, 128))) return rc;
    if (a.epilogue == VVAE_EPI_RESIDUAL || a.epilogue == VVAE_EPI_DSILU)
      if ((rc = encode_tmap_2d_bf16(&tai, a.aux_in, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ld_aux_in * 2, 64, 128, 128)))
        return rc;
    if (a.epilogue == VVAE_EPI_SILU && a.aux_out)
      if ((rc = encode_tmap_2d_bf16(&tao, a.aux_out, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ld_aux_out * 2, 64, 128, 128)))
        return rc;
    if (a.epilogue == VVAE_EPI_QKNORM_ROPE)
      if ((rc = encode_tmap_2d_bf16(&tao, a.aux_out, (uint64_t)(2 * a.qk_heads * a.qk_hd), (uint64_t)a.M,
                                    (uint64_t)a.ld_aux_out * 2, 64, 128, 128)))
        return rc;
  }
  // K-major SW128: 8-row groups 1024 B apart, K advance 32 B inside the swizzled row.
  // MN-major SW128: 64-wide MN chunks one TMA box (64 k-rows x 128 B = 8192 B) apart, 8-k groups 1024 B apart,
  // K advance = 16 rows x 128 B.
  p.a_lbo = A_MN ? 8192 : 16;  p.a_sbo = 1024;  p.a_kadv = A_MN ? 2048 : 32;
  p.b_lbo = B_MN ? 8192 : 16;  p.b_sbo = 1024;  p.b_kadv = B_MN ? 2048 : 32;
  if (g_dbg[1]) { if (A_MN) p.a_lbo = (uint32_t)g_dbg[1]; if (B_MN) p.b_lbo = (uint32_t)g_dbg[1]; }
  if (g_dbg[2]) { if (A_MN) p.a_sbo = (uint32_t)g_dbg[2]; if (B_MN) p.b_sbo = (uint32_t)g_dbg[2]; }
  if (g_dbg[3]) { if (A_MN) p.a_kadv = (uint32_t)g_dbg[3]; if (B_MN) p.b_kadv = (uint32_t)g_dbg[3]; }
  if (g_dbg[4]) { if (!A_MN) p.a_lbo = (uint32_t)g_dbg[4]; if (!B_MN) p.b_lbo = (uint32_t)g_dbg[4]; }
  if (g_dbg[5]) { if (!A_MN) p.a_sbo = (uint32_t)g_dbg[5]; if (!B_MN) p.b_sbo = (uint32_t)g_dbg[5]; }

  auto kern = gemm_sm100_kernel<BN, CG, A_MN, B_MN, MODE>;
  static std::atomic<bool> attr_set{false};  // per template instantiation
  if (!attr_set.load(std::memory_order_acquire)) {   // idempotent: a racing second call sets the same value
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return VVAE_ERR_CUDA;
    }
    attr_set.store(true, std::memory_order_release);
  }
  const int total = p.total_tiles;
  int clusters = std::min(total, g_dbg[0] ? (int)g_dbg[0] : n_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CG));
  cfg.blockDim = dim3(MODE == MODE_BSUM ? 384 : (MODE == MODE_QKN ? 192 : 320));
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dbg[11] ? 1 : 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tai, tao, p);
  if (e != cudaSuccess) {
    set_error("gemm_sm100 launch: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return VVAE_ERR_CUDA;
  }
  return check_launch("gemm_sm100");
}

template <int BN, int CG>
static int dispatch_major(const vvae_gemm_args& a, cudaStream_t s) {
  const bool a_mn = a.transA != 0;   // op(A)[m,k] = A[k*lda+m]  -> M contiguous
  const bool b_mn = a.transB == 0;   // op(B)[k,n] = B[k*ldb+n]  -> N contiguous
  if (a_mn) {
    if (a.epilogue != VVAE_EPI_NONE) { set_error("gemm_sm100: fused epilogues need a K-major A operand"); return VVAE_ERR_UNSUPPORTED; }
    if (b_mn && a.bsum_accum) return launch_sm100<BN, CG, true, true, MODE_BSUM>(a, s);
    return b_mn ? launch_sm100<BN, CG, true, true, 0>(a, s) : launch_sm100<BN, CG, true, false, 0>(a, s);
  }
  switch (a.epilogue) {
    case VVAE_EPI_SILU:
      return b_mn ? launch_sm100<BN, CG, false, true, 1>(a, s) : launch_sm100<BN, CG, false, false, 1>(a, s);
    case VVAE_EPI_RESIDUAL:
      return b_mn ? launch_sm100<BN, CG, false, true, 2>(a, s) : launch_sm100<BN, CG, false, false, 2>(a, s);
    case VVAE_EPI_DSILU:
      return b_mn ? launch_sm100<BN, CG, false, true, 3>(a, s) : launch_sm100<BN, CG, false, false, 3>(a, s);
    case VVAE_EPI_QKNORM_ROPE:
      if (!b_mn) { set_error("gemm_sm100: VVAE_EPI_QKNORM_ROPE expects the Flax (in,out) weight layout"); return VVAE_ERR_UNSUPPORTED; }
      return launch_sm100<BN, CG, false, true, MODE_QKN>(a, s);
    default:
      return b_mn ? launch_sm100<BN, CG, false, true, 0>(a, s) : launch_sm100<BN, CG, false, false, 0>(a, s);
  }
}

int sm100_gemm(const vvae_gemm_args& a, cudaStream_t s) {
  int bn = (int)g_dbg[6];
  if (!bn) bn = (a.N % 256 == 0 || a.N > 512) ? 256 : (a.N % 128 == 0 || a.N > 192 ? 128 : 64);
  // CTA pairs (cta_group::2, 256-row tiles) halve the L2 -> smem operand traffic per FLOP; used whenever there are at
  // least two 128-row blocks.  vvae_debug_set(8, 1) forces single-CTA tiles.
  const bool pair = a.M >= 256 && !g_dbg[8];
  if (bn == 256 && !pair) bn = 128;   // single-CTA 128x256 tiles leave too little smem for a deep pipeline + staging
  // N = 768 dgrads (K-major B): 192-column tiles give 4 x 128 = 512 tiles = 6.9 waves of 0.75 instead of 384 tiles =
  // 5.2 -> 6 waves of 1.  Measured: bit-identical, but 72.8 us against 65.7 us (M = 32768, K = 1536): the smaller tile's
  // extra operand traffic costs more than the partial wave.  Kept behind vvae_debug_set(12, 1).
  if (g_dbg[12] && bn == 256 && a.N % 192 == 0 && !a.transA && a.transB && a.epilogue == VVAE_EPI_NONE && !a.bsum_accum) {
    const long long mt = cdiv(a.M, 256), pairs = num_sms() / 2;
    const double cost256 = (double)cdiv(mt * cdiv(a.N, 256), pairs), cost192 = 0.75 * (double)cdiv(mt * (a.N / 192), pairs);
    if (cost192 < 0.9 * cost256) return launch_sm100<192, 2, false, false, 0>(a, s);
  }
  if (bn == 256) return dispatch_major<256, 2>(a, s);
  if (bn == 128) return pair ? dispatch_major<128, 2>(a, s) : dispatch_major<128, 1>(a, s);
  return dispatch_major<64, 1>(a, s);
}

}  // namespace vvae

namespace vvae { int attn_debug_read(unsigned long long* out32); }
extern "C" int vvae_debug_get(int what, unsigned long long* out4) {
  if (!out4) return VVAE_ERR_INVALID;
  if (what == 1) return vvae::attn_debug_read(out4);
  return cudaMemcpyFromSymbol(out4, vvae::g_gemm_dbg, 4 * sizeof(unsigned long long)) == cudaSuccess ? VVAE_OK : VVAE_ERR_CUDA;
}

extern "C" int vvae_debug_set(int key, long long value) {
  if (key < 0 || key >= 32) return VVAE_ERR_INVALID;
  vvae::g_dbg[key] = value;
  return VVAE_OK;
}
