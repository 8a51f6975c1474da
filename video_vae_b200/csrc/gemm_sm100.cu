// bf16 GEMM on the 5th-generation tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (accumulators in
// TMEM, double buffered) -> tcgen05.ld epilogue.  Persistent, warp specialised:
//   warp 0   : TMA producer (one lane)
//   warp 1   : TMEM allocator + MMA issuer (one lane)
//   warps 2-5: epilogue (each owns the TMEM lane quarter warp_id % 4)
// Both operands may be K-major (contraction index contiguous in memory) or MN-major (row/column index contiguous),
// so forward (X . W, W is Flax (in,out) = MN-major B), dgrad (dY . W^T, K-major B) and wgrad (X^T . dY, both MN-major,
// split-K with fp32 atomics) all read the tensors where they lie -- no transposed copies are ever materialised.
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

namespace vvae {

// ---- debug / tuning knobs (vvae_debug_set) ----
long long g_dbg[16] = {0};

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

// ------------------------------------------------------------------ device kernel
constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

struct Sm100Params {
  int M, N;                // output extent
  int m_tiles, n_tiles, splits;
  int kb_total, kb_per_split;  // K blocks of 64
  void* C; long long ldc; int out_f32;
  const float* bias;
  int mode;
  const bf16* aux_in; long long ld_ai;
  bf16* aux_out; long long ld_ao;
  int atomic;
  // descriptor encodings (bytes); overridable through vvae_debug_set for bring-up
  uint32_t a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv;
  uint32_t dbg_a_shift;  // bring-up experiment: byte offset added to the A start address (row-shifted operand views)
};

template <int BN> struct StageCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void epilogue_store_chunk(const Sm100Params& p, long long m, int n0, const uint32_t (&r)[32]) {
  // one thread owns row m, 32 consecutive columns starting at n0
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const bool full = (n0 + 32 <= p.N);
  if (p.bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
    }
  }
  if (p.mode == VVAE_EPI_SILU) {
    if (p.aux_out) {
      bf16* ao = p.aux_out + m * p.ld_ao + n0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          Vec16<bf16> o;
#pragma unroll
          for (int t = 0; t < 8; ++t) o.set(t, v[j + t]);
          o.store(ao + j);
        }
      } else {
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) ao[j] = __float2bfloat16_rn(v[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = siluf_(round_to<bf16>(v[j]));
  } else if (p.mode == VVAE_EPI_RESIDUAL || p.mode == VVAE_EPI_DSILU) {
    const bf16* ai = p.aux_in + m * p.ld_ai + n0;
    float a[32];
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        Vec16<bf16> x;
        x.load(ai + j);
#pragma unroll
        for (int t = 0; t < 8; ++t) a[j + t] = x.get(t);
      }
    } else {
      for (int j = 0; j < 32; ++j) a[j] = (n0 + j < p.N) ? __bfloat162float(ai[j]) : 0.f;
    }
    if (p.mode == VVAE_EPI_RESIDUAL) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += a[j];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= dsiluf_(a[j]);
    }
  }
  if (p.out_f32) {
    float* c = reinterpret_cast<float*>(p.C) + m * p.ldc + n0;
    if (p.atomic) {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) atomicAdd(c + j, v[j]);
    } else if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) c[j] = v[j];
    }
  } else {
    bf16* c = reinterpret_cast<bf16*>(p.C) + m * p.ldc + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        Vec16<bf16> o;
#pragma unroll
        for (int t = 0; t < 8; ++t) o.set(t, v[j + t]);
        o.store(c + j);
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) c[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, Sm100Params p) {
  using Cfg = StageCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_a);
    sm100::tma_prefetch_desc(&tma_b);
    for (int i = 0; i < STAGES; ++i) {
      sm100::mbar_init(&full_bar[i], 1);
      sm100::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&tmem_full[i], 1);
      sm100::mbar_init(&tmem_empty[i], 4);
    }
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<Cfg::TMEM_COLS>(tmem_base_slot);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_mn * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn, mn = tile % tiles_mn;
        const int m0 = (mn / p.n_tiles) * BM, n0 = (mn % p.n_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          sm100::mbar_wait(&empty_bar[stage], phase ^ 1);
          sm100::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
          const int k0 = kb * BK;
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) sm100::tma_load_2d(sa + j * 8192, &tma_a, &full_bar[stage], m0 + 64 * j, k0);
          } else {
            sm100::tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) sm100::tma_load_2d(sb + j * 8192, &tma_b, &full_bar[stage], n0 + 64 * j, k0);
          } else {
            sm100::tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = sm100::make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        sm100::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        sm100::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          sm100::mbar_wait(&full_bar[stage], phase);
          sm100::tc_fence_after();
          const uint32_t a_addr = sm100::smem_u32(smem_a + stage * A_STAGE_BYTES) + p.dbg_a_shift;
          const uint32_t b_addr = sm100::smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = sm100::make_smem_desc_sw128(a_addr + k * p.a_kadv, p.a_lbo, p.a_sbo);
            const uint64_t db = sm100::make_smem_desc_sw128(b_addr + k * p.b_kadv, p.b_lbo, p.b_sbo);
            sm100::umma_f16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          sm100::umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        sm100::umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mn = tile % tiles_mn;
      const int m0 = (mn / p.n_tiles) * BM, n0 = (mn % p.n_tiles) * BN;
      sm100::mbar_wait(&tmem_full[acc], acc_phase);
      sm100::tc_fence_after();
      const long long m = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (n0 + c * 32 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        sm100::tmem_ld_32x32(taddr + c * 32, r);
        sm100::tmem_ld_wait();
        if (m < p.M) epilogue_store_chunk(p, m, n0 + c * 32, r);
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ host launch
bool sm100_gemm_supported(const vvae_gemm_args& a) {
  if (a.dtype != VVAE_BF16) return false;
  if (a.M < 128 || a.N < 64 || a.K < 64) return false;
  if (a.N % 8 != 0) return false;
  if (a.lda % 8 || a.ldb % 8) return false;
  if (((uintptr_t)a.A % 16) || ((uintptr_t)a.B % 16) || ((uintptr_t)a.C % 16)) return false;
  if (a.out_dtype == VVAE_BF16 && (a.ldc % 8)) return false;
  if (a.out_dtype == VVAE_F32 && (a.ldc % 4)) return false;
  if (a.aux_in && ((a.ld_aux_in % 8) || ((uintptr_t)a.aux_in % 16))) return false;
  if (a.aux_out && ((a.ld_aux_out % 8) || ((uintptr_t)a.aux_out % 16))) return false;
  if (a.bias && ((uintptr_t)a.bias % 16)) return false;
  if (a.accumulate && a.out_dtype != VVAE_F32) return false;
  // MN-major operands are fetched in 64-wide boxes along M / N
  if (a.transA && (a.M % 8)) return false;
  return true;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_sm100(const vvae_gemm_args& a, cudaStream_t s) {
  using Cfg = StageCfg<BN>;
  CUtensorMap ta, tb;
  int rc;
  // A: op(A)[m,k].  K-major: memory [M rows][K cols];  MN-major (transA): memory [K rows][M cols].
  if (A_MN) rc = encode_tmap_2d_bf16(&ta, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda * 2, 64, BK, 128);
  else      rc = encode_tmap_2d_bf16(&ta, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda * 2, BK, BM, 128);
  if (rc) return rc;
  // B: op(B)[k,n].  K-major (transB): memory [N rows][K cols];  MN-major: memory [K rows][N cols].
  if (B_MN) rc = encode_tmap_2d_bf16(&tb, a.B, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb * 2, 64, BK, 128);
  else      rc = encode_tmap_2d_bf16(&tb, a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb * 2, BK, BN, 128);
  if (rc) return rc;

  Sm100Params p;
  p.M = a.M; p.N = a.N;
  p.m_tiles = (int)cdiv(a.M, BM);
  p.n_tiles = (int)cdiv(a.N, BN);
  p.kb_total = (int)cdiv(a.K, BK);
  int splits = 1;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  if (a.accumulate && tiles_mn < 148) {  // weight gradients: few output tiles, very long K -> split K over the SMs
    splits = std::max(1, std::min(148 / tiles_mn, p.kb_total / 8));
  }
  p.kb_per_split = (int)cdiv(p.kb_total, splits);
  p.splits = (int)cdiv(p.kb_total, p.kb_per_split);
  p.C = a.C; p.ldc = a.ldc; p.out_f32 = (a.out_dtype == VVAE_F32);
  p.bias = a.bias; p.mode = a.epilogue;
  p.aux_in = (const bf16*)a.aux_in; p.ld_ai = a.ld_aux_in;
  p.aux_out = (bf16*)a.aux_out; p.ld_ao = a.ld_aux_out;
  p.atomic = (a.accumulate || p.splits > 1) ? 1 : 0;
  // K-major SW128: 8-row groups 1024 B apart, K advance 32 B inside the swizzled row.
  // MN-major SW128: 64-wide MN chunks one TMA box (64 k-rows x 128 B = 8192 B) apart, 8-k groups 1024 B apart,
  // K advance = 16 rows x 128 B.
  p.a_lbo = A_MN ? 8192 : 16;  p.a_sbo = 1024;  p.a_kadv = A_MN ? 2048 : 32;
  p.b_lbo = B_MN ? 8192 : 16;  p.b_sbo = 1024;  p.b_kadv = B_MN ? 2048 : 32;
  p.dbg_a_shift = (uint32_t)g_dbg[7];
  if (g_dbg[1]) { if (A_MN) p.a_lbo = (uint32_t)g_dbg[1]; if (B_MN) p.b_lbo = (uint32_t)g_dbg[1]; }
  if (g_dbg[2]) { if (A_MN) p.a_sbo = (uint32_t)g_dbg[2]; if (B_MN) p.b_sbo = (uint32_t)g_dbg[2]; }
  if (g_dbg[3]) { if (A_MN) p.a_kadv = (uint32_t)g_dbg[3]; if (B_MN) p.b_kadv = (uint32_t)g_dbg[3]; }
  if (g_dbg[4]) { if (!A_MN) p.a_lbo = (uint32_t)g_dbg[4]; if (!B_MN) p.b_lbo = (uint32_t)g_dbg[4]; }
  if (g_dbg[5]) { if (!A_MN) p.a_sbo = (uint32_t)g_dbg[5]; if (!B_MN) p.b_sbo = (uint32_t)g_dbg[5]; }

  auto kern = gemm_sm100_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return VVAE_ERR_CUDA;
    }
    attr_set = true;
  }
  const int total = tiles_mn * p.splits;
  int grid = std::min(total, g_dbg[0] ? (int)g_dbg[0] : 148);
  kern<<<grid, 192, Cfg::SMEM_BYTES, s>>>(ta, tb, p);
  return check_launch("gemm_sm100");
}

template <int BN>
static int dispatch_major(const vvae_gemm_args& a, cudaStream_t s) {
  const bool a_mn = a.transA != 0;   // op(A)[m,k] = A[k*lda+m]  -> M contiguous
  const bool b_mn = a.transB == 0;   // op(B)[k,n] = B[k*ldb+n]  -> N contiguous
  if (a_mn) return b_mn ? launch_sm100<BN, true, true>(a, s) : launch_sm100<BN, true, false>(a, s);
  return b_mn ? launch_sm100<BN, false, true>(a, s) : launch_sm100<BN, false, false>(a, s);
}

int sm100_gemm(const vvae_gemm_args& a, cudaStream_t s) {
  int bn = (int)g_dbg[6];
  if (!bn) bn = (a.N % 256 == 0 || a.N > 512) ? 256 : (a.N % 128 == 0 || a.N > 192 ? 128 : 64);
  if (bn == 256) return dispatch_major<256>(a, s);
  if (bn == 128) return dispatch_major<128>(a, s);
  return dispatch_major<64>(a, s);
}

}  // namespace vvae

extern "C" int vvae_debug_set(int key, long long value) {
  if (key < 0 || key >= 16) return VVAE_ERR_INVALID;
  vvae::g_dbg[key] = value;
  return VVAE_OK;
}
