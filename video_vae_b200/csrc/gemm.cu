// vvae_gemm: dispatch between the tcgen05 kernel (bf16, aligned shapes) and the generic SIMT kernel.
#include "gemm_simt.cuh"

namespace vvae {
bool sm100_gemm_supported(const vvae_gemm_args& a);
int sm100_gemm(const vvae_gemm_args& a, cudaStream_t s);

struct SmallLinArgs {
  const bf16* x; long long x_ld;
  const bf16* w; long long wk, wn;
  const float* bias;
  bf16* y; long long y_ld;
  const bf16* aux; long long aux_ld;
  long long M; int K, N;
};
bool small_linear_ok(int K, int N);
bool small_linear_wgrad_ok(int K, int N);
int small_linear_fwd(const SmallLinArgs& a, cudaStream_t s);
int small_linear_wgrad(const bf16* x, long long x_ld, const bf16* dy, long long dy_ld, float* dw, long long dk, long long dn,
                       long long M, int K, int N, cudaStream_t s);
bool rank1_ok(int K);
int rank1_fwd(const bf16* x, long long x_ld, const bf16* w, long long w_st, const float* bias, bf16* y, long long y_ld,
              long long M, int K, cudaStream_t s);
int rank1_dgrad(const bf16* dy, long long dy_ld, const bf16* w, long long w_st, const bf16* aux, long long aux_ld, bf16* dx,
                long long dx_ld, long long M, int K, cudaStream_t s);
int rank1_wgrad(const bf16* x, long long x_ld, const bf16* dy, long long dy_ld, float* dw, long long dw_st, long long M, int K,
                cudaStream_t s);

// per-row linears with a handful of features (HBM-bound streams, see small_linear.cu)
static bool small_route(const vvae_gemm_args& a, cudaStream_t s, int* rc) {
  if (a.dtype != VVAE_BF16 || a.backend != VVAE_BACKEND_AUTO) return false;
  auto al16 = [](const void* p_) { return ((uintptr_t)p_ % 16) == 0; };
  // rank-1 shapes of the frame-selection head (Linear K -> 1 over every token and its two gradients)
  if (!a.transA && a.N == 1 && rank1_ok(a.K) && a.M >= 1024 && a.out_dtype == VVAE_BF16 && !a.accumulate &&
      a.epilogue == VVAE_EPI_NONE && !a.bsum_accum && al16(a.A) && a.lda % 8 == 0) {
    *rc = rank1_fwd((const bf16*)a.A, a.lda, (const bf16*)a.B, a.transB ? 1 : a.ldb, a.bias, (bf16*)a.C, a.ldc, a.M, a.K, s);
    return true;
  }
  if (!a.transA && a.K == 1 && rank1_ok(a.N) && a.M >= 1024 && a.out_dtype == VVAE_BF16 && !a.accumulate && !a.bias &&
      (a.epilogue == VVAE_EPI_NONE || a.epilogue == VVAE_EPI_RESIDUAL) && !a.bsum_accum && al16(a.C) && a.ldc % 8 == 0 &&
      (a.epilogue == VVAE_EPI_NONE || (al16(a.aux_in) && a.ld_aux_in % 8 == 0))) {
    // C[m,n] = A[m,0] * op(B)[0,n]: op(B)[0,n] = B[n*ldb] (transB, B is [N,1]) or B[n] (B is [1,N])
    *rc = rank1_dgrad((const bf16*)a.A, a.lda, (const bf16*)a.B, a.transB ? a.ldb : 1,
                      a.epilogue == VVAE_EPI_RESIDUAL ? (const bf16*)a.aux_in : nullptr, a.ld_aux_in, (bf16*)a.C, a.ldc,
                      a.M, a.N, s);
    return true;
  }
  if (a.transA && !a.transB && a.N == 1 && rank1_ok(a.M) && a.K >= 1024 && a.accumulate && a.out_dtype == VVAE_F32 &&
      a.epilogue == VVAE_EPI_NONE && !a.bias && !a.bsum_accum && al16(a.A) && a.lda % 8 == 0) {
    // C[k,0] += sum_rows A[row,k] * B[row,0]   (A is [rows, M'] with M' = a.M the feature count)
    *rc = rank1_wgrad((const bf16*)a.A, a.lda, (const bf16*)a.B, a.ldb, (float*)a.C, a.ldc, a.K, a.M, s);
    return true;
  }
  if (!a.transA && a.out_dtype == VVAE_BF16 && !a.accumulate && small_linear_ok(a.K, a.N) && a.M >= 4096 &&
      (a.epilogue == VVAE_EPI_NONE || a.epilogue == VVAE_EPI_RESIDUAL)) {
    SmallLinArgs q{(const bf16*)a.A, a.lda, (const bf16*)a.B, a.transB ? 1 : a.ldb, a.transB ? a.ldb : 1, a.bias,
                   (bf16*)a.C, a.ldc, a.epilogue == VVAE_EPI_RESIDUAL ? (const bf16*)a.aux_in : nullptr, a.ld_aux_in,
                   a.M, a.K, a.N};
    *rc = small_linear_fwd(q, s);
    return true;
  }
  if (a.transA && !a.transB && a.accumulate && a.out_dtype == VVAE_F32 && a.epilogue == VVAE_EPI_NONE && !a.bias &&
      small_linear_wgrad_ok(a.M, a.N) && a.K >= 4096) {
    *rc = small_linear_wgrad((const bf16*)a.A, a.lda, (const bf16*)a.B, a.ldb, (float*)a.C, a.ldc, 1, a.K, a.M, a.N, s);
    return true;
  }
  return false;
}

template <typename T, typename TO>
static int gemm_simt_typed(const vvae_gemm_args& a, cudaStream_t s) {
  EpiStore<TO, T> ep{(TO*)a.C, a.ldc, a.bias, a.epilogue, (const T*)a.aux_in, a.ld_aux_in, (T*)a.aux_out, a.ld_aux_out, 0};
  int splits = 1;
  if (a.accumulate) {
    ep.atomic = 1;
    long long tiles = cdiv(a.M, SG_BM) * cdiv(a.N, SG_BN);
    if (tiles < num_sms() * 2) splits = (int)std::max<long long>(1, std::min<long long>((num_sms() * 4) / tiles, a.K / 256));
  }
  const T* A = (const T*)a.A;
  const T* B = (const T*)a.B;
  if (!a.transA && !a.transB)
    return launch_gemm_simt(RowMajorLoader<T>{A, a.lda}, RowMajorLoader<T>{B, a.ldb}, ep, a.M, a.N, a.K, splits, s);
  if (!a.transA && a.transB)
    return launch_gemm_simt(RowMajorLoader<T>{A, a.lda}, ColMajorLoader<T>{B, a.ldb}, ep, a.M, a.N, a.K, splits, s);
  if (a.transA && !a.transB)
    return launch_gemm_simt(ColMajorLoader<T>{A, a.lda}, RowMajorLoader<T>{B, a.ldb}, ep, a.M, a.N, a.K, splits, s);
  return launch_gemm_simt(ColMajorLoader<T>{A, a.lda}, ColMajorLoader<T>{B, a.ldb}, ep, a.M, a.N, a.K, splits, s);
}
}  // namespace vvae

using namespace vvae;

extern "C" int vvae_gemm_uses_tcgen05(const vvae_gemm_args* args) {
  if (!args) return 0;
  if (args->backend == VVAE_BACKEND_SIMT) return 0;
  return sm100_gemm_supported(*args) ? 1 : 0;
}

extern "C" int vvae_gemm(const vvae_gemm_args* args, vvae_stream_t stream) {
  VVAE_REQUIRE(args, "vvae_gemm: null args");
  const vvae_gemm_args& a = *args;
  VVAE_REQUIRE(a.M >= 0 && a.N >= 0 && a.K >= 0, "vvae_gemm: negative extent");
  if (a.M == 0 || a.N == 0) return VVAE_OK;
  VVAE_REQUIRE(a.A && a.B && a.C, "vvae_gemm: null operand");
  VVAE_REQUIRE(a.dtype == VVAE_F32 || a.dtype == VVAE_BF16, "vvae_gemm: bad dtype %d", a.dtype);
  VVAE_REQUIRE(a.out_dtype == a.dtype || a.out_dtype == VVAE_F32, "vvae_gemm: out_dtype must be dtype or f32");
  VVAE_REQUIRE(!a.accumulate || a.out_dtype == VVAE_F32, "vvae_gemm: accumulate needs fp32 output");
  VVAE_REQUIRE(!a.accumulate || a.epilogue == VVAE_EPI_NONE, "vvae_gemm: accumulate excludes fused epilogues");
  VVAE_REQUIRE((a.epilogue != VVAE_EPI_RESIDUAL && a.epilogue != VVAE_EPI_DSILU) || a.aux_in,
               "vvae_gemm: epilogue %d needs aux_in", a.epilogue);
  cudaStream_t s = as_stream(stream);
  if (a.epilogue == VVAE_EPI_QKNORM_ROPE) {
    VVAE_REQUIRE(a.aux_out && a.qk_q_scale && a.qk_k_scale && a.rope_cos && a.rope_sin && a.qk_heads > 0 && a.qk_hd > 0 &&
                     a.N == 3 * a.qk_heads * a.qk_hd && a.ld_aux_out >= 2LL * a.qk_heads * a.qk_hd && !a.accumulate &&
                     a.out_dtype == a.dtype && a.ldc == a.N,
                 "vvae_gemm: VVAE_EPI_QKNORM_ROPE needs N = 3*heads*hd, a dense C, aux_out [M, 2*heads*hd], scales and tables");
    if (!(sm100_gemm_supported(a) && a.backend != VVAE_BACKEND_SIMT)) {
      // generic route: the projection, then the stand-alone QK-LayerNorm + RoPE kernel over its q|k columns
      VVAE_REQUIRE(a.ld_aux_out == 2LL * a.qk_heads * a.qk_hd, "vvae_gemm: VVAE_EPI_QKNORM_ROPE generic route needs a dense aux_out");
      vvae_gemm_args b = a;
      b.epilogue = VVAE_EPI_NONE;
      b.aux_out = nullptr;
      int rc = vvae_gemm(&b, stream);
      if (rc) return rc;
      return vvae_qknorm_rope_fwd(a.C, a.aux_out, a.qk_q_scale, a.qk_k_scale, a.rope_cos, a.rope_sin, a.M, a.qk_heads,
                                  a.qk_hd, a.rope_pos_div, a.rope_pos_mod, a.qk_eps, a.dtype, stream);
    }
    return sm100_gemm(a, s);
  }
  const bool tc_ok = sm100_gemm_supported(a);
  if (a.bsum_accum) {
    VVAE_REQUIRE(!a.transB, "vvae_gemm: bsum_accum needs op(B) = B (rows = contraction index)");
    const bool fused = tc_ok && a.backend != VVAE_BACKEND_SIMT && a.transA && a.dtype == VVAE_BF16;
    if (!fused) {   // generic kernels: the column sums are a separate pass over B
      int rc = vvae_colsum(a.B, a.ldb, a.K, a.N, a.bsum_accum, a.dtype, stream);
      if (rc) return rc;
      vvae_gemm_args b = a;
      b.bsum_accum = nullptr;
      return vvae_gemm(&b, stream);
    }
  }
  if (a.backend == VVAE_BACKEND_TCGEN05) {
    if (!tc_ok) {
      set_error("vvae_gemm: shape/alignment not supported by the tcgen05 path (M=%d N=%d K=%d)", a.M, a.N, a.K);
      return VVAE_ERR_UNSUPPORTED;
    }
    return sm100_gemm(a, s);
  }
  if (a.backend == VVAE_BACKEND_AUTO && tc_ok) return sm100_gemm(a, s);
  {
    int rc = 0;
    if (small_route(a, s, &rc)) return rc;
  }
  if (a.dtype == VVAE_F32) return gemm_simt_typed<float, float>(a, s);
  if (a.out_dtype == VVAE_F32) return gemm_simt_typed<bf16, float>(a, s);
  return gemm_simt_typed<bf16, bf16>(a, s);
}
