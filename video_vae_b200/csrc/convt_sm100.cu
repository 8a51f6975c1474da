// nnx.ConvTranspose kernel (1,2,2), strides (1,2,2) (train/unet.py:61-69) on the tcgen05 GEMM.
//
// A non-overlapping 2x2 up-sampling conv is one GEMM over the low-res voxels:  Yv[V, (a,c,co)] = X[V, Cin] . Wp + b  with
// Wp[ci, (a,c,co)] = W[1-a, 1-c, ci, co] (Flax's tap order is flipped), followed by the pixel shuffle
// Y[i', 2i+a, 2j+c, co] = Yv[(i',i,j), (a,c,co)].  Forward = GEMM + shuffle (16-byte vectors, channel stride y_ld so
// the result lands directly in the skip-concat buffer); backward = un-shuffle of dY, then the dgrad GEMM
// dX = dYv . Wp^T and the split-K wgrad GEMM dWp = X^T . dYv on the same persistent tensor-core kernel.
#include <algorithm>

#include "common.cuh"

namespace vvae {

int sm100_gemm(const vvae_gemm_args& a, cudaStream_t s);

// Wp[ci, (a, c, co)] = w[(1-a, 1-c), ci, co];  bias4[(a, c, co)] = bias[co]
__global__ void convt_wprep_kernel(const bf16* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ wp,
                                   float* __restrict__ bias4, int Cin, int Cout) {
  const int n4 = 4 * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Cin * n4; i += gridDim.x * blockDim.x) {
    const int ci = i / n4, n = i % n4, tap = n / Cout, co = n % Cout;
    wp[i] = w[((long long)(3 - tap) * Cin + ci) * Cout + co];
  }
  if (bias4)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) bias4[i] = bias ? bias[i % Cout] : 0.f;
}
// dw[(1-a, 1-c), ci, co] += dWp[ci, (a, c, co)]
__global__ void convt_dw_scatter_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cin, int Cout) {
  const int n4 = 4 * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Cin * n4; i += gridDim.x * blockDim.x) {
    const int ci = i / n4, n = i % n4, tap = n / Cout, co = n % Cout;
    dw[((long long)(3 - tap) * Cin + ci) * Cout + co] += dwp[i];
  }
}

// dense Yv[V, 4*Cout] <-> Y[bt, 2H, 2W, ld] (first Cout channels); one 16-byte vector (8 channels) per thread per step.
// dir 0: dense -> Y (forward scatter), dir 1: Y -> dense (backward gather).
__global__ void __launch_bounds__(256)
convt_shuffle_kernel(bf16* __restrict__ dense, bf16* __restrict__ y, long long y_ld, long long n_vec, int H, int W, int Cout,
                     int dir) {
  const unsigned cv = (unsigned)Cout >> 3;      // vectors per (voxel, tap)
  const unsigned W2 = 2u * (unsigned)W, H2 = 2u * (unsigned)H;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)n_vec; i += stride) {   // n_vec < 2^31 (host check)
    // enumerate in FULL-RES order (coalesced on the strided side): i = ((bt, y2, x2), v)
    const unsigned v = i % cv, p1 = i / cv, x2 = p1 % W2, p2 = p1 / W2, y2 = p2 % H2, bt = p2 / H2;
    const long long vox = ((long long)bt * H + (y2 >> 1)) * W + (x2 >> 1);
    const unsigned tap = ((y2 & 1) << 1) | (x2 & 1);
    bf16* d = dense + (vox * 4 + tap) * Cout + v * 8;
    bf16* f = y + (((long long)bt * H2 + y2) * W2 + x2) * y_ld + v * 8;
    if (dir == 0) *reinterpret_cast<uint4*>(f) = *reinterpret_cast<const uint4*>(d);
    else          *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(f);
  }
}

static inline long long up256(long long v) { return (v + 255) / 256 * 256; }

long long convt_tc_workspace_bytes(int b_t, int H, int W, int Cin, int Cout) {
  const long long V = (long long)b_t * H * W;
  return up256((long long)Cin * 4 * Cout * 2) + up256(4LL * Cout * 4) + up256((long long)Cin * 4 * Cout * 4) +
         up256(V * 4 * Cout * 2);
}

bool convt_tc_supported(int dtype, int b_t, int H, int W, int Cin, int Cout, const void* x, const void* y, long long y_ld,
                        const void* ws, long long ws_bytes) {
  if (dtype != VVAE_BF16 || !ws || ws_bytes < convt_tc_workspace_bytes(b_t, H, W, Cin, Cout)) return false;
  if (((uintptr_t)ws % 256) || ((uintptr_t)x % 16) || ((uintptr_t)y % 16)) return false;
  if (Cin % 8 || Cout % 8 || y_ld % 8) return false;
  const long long V = (long long)b_t * H * W;
  return V >= 128 && V * 4 * (Cout / 8) < (1LL << 31) && Cin >= 16 && Cout >= 16;
}

static void split_ws(void* ws, int Cin, int Cout, bf16** wp, float** bias4, float** dwp, bf16** dense) {
  uint8_t* p = (uint8_t*)ws;
  *wp = (bf16*)p;
  p += up256((long long)Cin * 4 * Cout * 2);
  *bias4 = (float*)p;
  p += up256(4LL * Cout * 4);
  *dwp = (float*)p;
  p += up256((long long)Cin * 4 * Cout * 4);
  *dense = (bf16*)p;
}

static int shuffle(bf16* dense, void* y, long long y_ld, int b_t, int H, int W, int Cout, int dir, cudaStream_t s) {
  const long long n_vec = (long long)b_t * H * W * 4 * (Cout / 8);
  const int blocks = (int)std::min<long long>(cdiv(n_vec, 256), (long long)num_sms() * 16);
  convt_shuffle_kernel<<<blocks, 256, 0, s>>>(dense, (bf16*)y, y_ld, n_vec, H, W, Cout, dir);
  return check_launch("convt_shuffle");
}

int convt_tc_fwd(const void* x, const void* w, const float* bias, void* y, long long y_ld, int b_t, int H, int W, int Cin,
                 int Cout, void* ws, cudaStream_t s) {
  bf16 *wp, *dense; float *bias4, *dwp;
  split_ws(ws, Cin, Cout, &wp, &bias4, &dwp, &dense);
  convt_wprep_kernel<<<std::min(num_sms(), (Cin * 4 * Cout + 255) / 256), 256, 0, s>>>((const bf16*)w, bias, wp, bias4, Cin, Cout);
  int rc = check_launch("convt_wprep");
  if (rc) return rc;
  vvae_gemm_args g = {};
  g.M = (int)((long long)b_t * H * W); g.N = 4 * Cout; g.K = Cin;
  g.A = x; g.lda = Cin; g.transA = 0;
  g.B = wp; g.ldb = 4 * Cout; g.transB = 0;
  g.C = dense; g.ldc = 4 * Cout;
  g.dtype = VVAE_BF16; g.out_dtype = VVAE_BF16; g.bias = bias4; g.epilogue = VVAE_EPI_NONE;
  if ((rc = sm100_gemm(g, s))) return rc;
  return shuffle(dense, y, y_ld, b_t, H, W, Cout, 0, s);
}

int convt_tc_bwd(const void* dy, long long dy_ld, const void* x, const void* w, void* dx, float* dw_accum, int b_t, int H,
                 int W, int Cin, int Cout, void* ws, cudaStream_t s) {
  bf16 *wp, *dense; float *bias4, *dwp;
  split_ws(ws, Cin, Cout, &wp, &bias4, &dwp, &dense);
  convt_wprep_kernel<<<std::min(num_sms(), (Cin * 4 * Cout + 255) / 256), 256, 0, s>>>((const bf16*)w, nullptr, wp, nullptr, Cin, Cout);
  int rc = check_launch("convt_wprep");
  if (rc) return rc;
  if ((rc = shuffle(dense, const_cast<void*>(dy), dy_ld, b_t, H, W, Cout, 1, s))) return rc;
  const int V = (int)((long long)b_t * H * W);
  if (dx) {   // dX[V, Cin] = dYv[V, 4Cout] . Wp^T
    vvae_gemm_args g = {};
    g.M = V; g.N = Cin; g.K = 4 * Cout;
    g.A = dense; g.lda = 4 * Cout; g.transA = 0;
    g.B = wp; g.ldb = 4 * Cout; g.transB = 1;
    g.C = dx; g.ldc = Cin;
    g.dtype = VVAE_BF16; g.out_dtype = VVAE_BF16; g.epilogue = VVAE_EPI_NONE;
    if ((rc = sm100_gemm(g, s))) return rc;
  }
  if (dw_accum) {   // dWp[Cin, 4Cout] = X^T . dYv, then scattered (+=) into the Flax layout
    if ((rc = vvae_fill_f32(dwp, 0.f, (long long)Cin * 4 * Cout, s))) return rc;
    vvae_gemm_args g = {};
    g.M = Cin; g.N = 4 * Cout; g.K = V;
    g.A = x; g.lda = Cin; g.transA = 1;
    g.B = dense; g.ldb = 4 * Cout; g.transB = 0;
    g.C = dwp; g.ldc = 4 * Cout;
    g.dtype = VVAE_BF16; g.out_dtype = VVAE_F32; g.accumulate = 1; g.epilogue = VVAE_EPI_NONE;
    if ((rc = sm100_gemm(g, s))) return rc;
    convt_dw_scatter_kernel<<<std::min(num_sms(), (Cin * 4 * Cout + 255) / 256), 256, 0, s>>>(dwp, dw_accum, Cin, Cout);
    if ((rc = check_launch("convt_dw_scatter"))) return rc;
  }
  return VVAE_OK;
}

}  // namespace vvae
