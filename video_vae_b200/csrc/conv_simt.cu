// Convolutions as implicit GEMMs on the generic functor GEMM (any dtype / channel count):
// conv3d NDHWC 'SAME' forward / dgrad / wgrad, and ConvTranspose k = s = (1,2,2) forward / backward.
// The bf16 production shapes are routed to the tensor-core path in conv_sm100.cu when it supports them.
#include "gemm_simt.cuh"

namespace vvae {

int conv_tc_supported(const vvae_conv_args& a, int which);
int conv_tc_launch(const vvae_conv_args& a, int which, cudaStream_t s);
int conv_wgrad_tc_supported(const vvae_conv_args& a);
int conv_wgrad_tc_launch(const vvae_conv_args& a, cudaStream_t s);

struct ConvGeom {
  int T, H, W, C;       // C = channels of the gathered tensor
  int kt, kh, kw;
  long long ld;
  int sign;             // +1: in = out + (tap - k/2) (forward / wgrad);  -1: dgrad
};

// A(m = voxel, k = (tap, c)) gathered from a channels-last tensor with zero padding.
template <typename T>
struct Im2colLoader {
  const T* x; ConvGeom g;
  __device__ __forceinline__ float operator()(long long m, long long k) const {
    int c = (int)(k % g.C);
    int tap = (int)(k / g.C);
    int dw = tap % g.kw, r = tap / g.kw;
    int dh = r % g.kh, dt = r / g.kh;
    int w = (int)(m % g.W);
    long long r2 = m / g.W;
    int h = (int)(r2 % g.H);
    long long r3 = r2 / g.H;
    int t = (int)(r3 % g.T);
    long long b = r3 / g.T;
    int tt = t + g.sign * (dt - g.kt / 2), hh = h + g.sign * (dh - g.kh / 2), ww = w + g.sign * (dw - g.kw / 2);
    if (tt < 0 || tt >= g.T || hh < 0 || hh >= g.H || ww < 0 || ww >= g.W) return 0.f;
    return to_f(x[(((b * g.T + tt) * g.H + hh) * g.W + ww) * g.ld + c]);
  }
};
template <typename T>
struct Im2colTLoader {  // A'(m' = (tap, c), k = voxel) for wgrad
  Im2colLoader<T> l;
  __device__ __forceinline__ float operator()(long long mp, long long k) const { return l(k, mp); }
};
// dgrad weights: B(k = (tap, co), n = ci) = w[tap][ci][co]
template <typename T>
struct WDgradLoader {
  const T* w; int Cin, Cout;
  __device__ __forceinline__ float operator()(long long k, long long n) const {
    long long tap = k / Cout, co = k % Cout;
    return to_f(w[(tap * Cin + n) * Cout + co]);
  }
};

// ---- ConvTranspose (1,2,2)/s2: out[2i+a, 2j+c] = x[i,j] . w[1-a, 1-c] ----
struct CTGeom { int H, W, Cin, Cout; };
template <typename T>
struct CTWeightLoader {  // B(k = ci, n = (a, c, co))
  const T* w; CTGeom g;
  __device__ __forceinline__ float operator()(long long k, long long n) const {
    int co = (int)(n % g.Cout), ac = (int)(n / g.Cout);
    int a = ac >> 1, c = ac & 1;
    return to_f(w[((long long)((1 - a) * 2 + (1 - c)) * g.Cin + k) * g.Cout + co]);
  }
};
template <typename T>
struct CTWeightTLoader {  // B'(k = (a,c,co), n = ci)
  CTWeightLoader<T> l;
  __device__ __forceinline__ float operator()(long long k, long long n) const { return l(n, k); }
};
__device__ __forceinline__ long long ct_out_index(const CTGeom& g, long long m, int n, long long ld) {
  int co = n % g.Cout, ac = n / g.Cout;
  int a = ac >> 1, c = ac & 1;
  int j = (int)(m % g.W);
  long long r = m / g.W;
  int i = (int)(r % g.H);
  long long bt = r / g.H;
  return ((bt * (2 * g.H) + 2 * i + a) * (2LL * g.W) + 2 * j + c) * ld + co;
}
template <typename T>
struct CTScatterEpi {  // forward: y[(bt, 2i+a, 2j+c), co] = acc + bias[co]
  T* y; long long ld; const float* bias; CTGeom g;
  __device__ __forceinline__ void operator()(long long m, int n, float acc, bool) const {
    float v = acc + (bias ? bias[n % g.Cout] : 0.f);
    y[ct_out_index(g, m, n, ld)] = from_f<T>(v);
  }
};
template <typename T>
struct CTGatherLoader {  // A(m = input voxel, k = (a,c,co)) = dy[(bt,2i+a,2j+c), co]
  const T* dy; long long ld; CTGeom g;
  __device__ __forceinline__ float operator()(long long m, long long k) const { return to_f(dy[ct_out_index(g, m, (int)k, ld)]); }
};
template <typename T>
struct CTGatherBLoader {  // B(k = input voxel, n = (a,c,co))
  CTGatherLoader<T> l;
  __device__ __forceinline__ float operator()(long long k, long long n) const { return l(k, n); }
};
struct CTWgradEpi {  // dw[(1-a,1-c), ci, co] += acc
  float* dw; CTGeom g;
  __device__ __forceinline__ void operator()(long long m /*ci*/, int n, float acc, bool) const {
    int co = n % g.Cout, ac = n / g.Cout;
    int a = ac >> 1, c = ac & 1;
    atomicAdd(dw + ((long long)((1 - a) * 2 + (1 - c)) * g.Cin + m) * g.Cout + co, acc);
  }
};

template <typename T>
static int conv_simt(const vvae_conv_args& a, int which, cudaStream_t s) {
  const long long V = (long long)a.B * a.T * a.H * a.W;
  const int taps = a.kt * a.kh * a.kw;
  if (which == 0) {  // forward
    ConvGeom g{a.T, a.H, a.W, a.Cin, a.kt, a.kh, a.kw, a.x_ld, +1};
    EpiStore<T, T> ep{(T*)a.y, a.y_ld, a.bias, a.epilogue, (const T*)a.aux_in, a.ld_aux, nullptr, 0, 0};
    return launch_gemm_simt(Im2colLoader<T>{(const T*)a.x, g}, RowMajorLoader<T>{(const T*)a.w, a.Cout}, ep, V, a.Cout,
                            (long long)taps * a.Cin, 1, s);
  }
  if (which == 1) {  // dgrad: dx[V,Cin] = im2col_flipped(dy)[V, taps*Cout] . W'[taps*Cout, Cin]
    ConvGeom g{a.T, a.H, a.W, a.Cout, a.kt, a.kh, a.kw, a.y_ld, -1};
    EpiStore<T, T> ep{(T*)const_cast<void*>(a.x), a.x_ld, nullptr, VVAE_EPI_NONE, nullptr, 0, nullptr, 0, 0};
    return launch_gemm_simt(Im2colLoader<T>{(const T*)a.y, g}, WDgradLoader<T>{(const T*)a.w, a.Cin, a.Cout}, ep, V, a.Cin,
                            (long long)taps * a.Cout, 1, s);
  }
  // wgrad: dw[taps*Cin, Cout] += im2col(x)^T . dy ; reduction over V voxels, split across the grid
  ConvGeom g{a.T, a.H, a.W, a.Cin, a.kt, a.kh, a.kw, a.x_ld, +1};
  EpiStore<float, T> ep{a.dw_accum, a.Cout, nullptr, VVAE_EPI_NONE, nullptr, 0, nullptr, 0, 1};
  const long long Mp = (long long)taps * a.Cin;
  const long long tiles = cdiv(Mp, SG_BM) * cdiv(a.Cout, SG_BN);
  int splits = (int)std::max<long long>(1, std::min<long long>(((long long)num_sms() * 8) / tiles, V / 512));
  return launch_gemm_simt(Im2colTLoader<T>{Im2colLoader<T>{(const T*)a.x, g}}, RowMajorLoader<T>{(const T*)a.y, a.y_ld}, ep,
                          Mp, a.Cout, V, splits, s);
}

}  // namespace vvae

using namespace vvae;

static int conv_validate(const vvae_conv_args* a, int which) {
  VVAE_REQUIRE(a, "conv3d: null args");
  VVAE_REQUIRE(a->B >= 0 && a->T > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "conv3d: bad extents");
  VVAE_REQUIRE((a->kt & 1) && (a->kh & 1) && (a->kw & 1), "conv3d: 'SAME' path needs odd kernel sizes");
  VVAE_REQUIRE(a->x && a->y, "conv3d: null tensor");
  VVAE_REQUIRE(which == 2 ? a->dw_accum != nullptr : a->w != nullptr, "conv3d: null weights");
  VVAE_REQUIRE(a->x_ld >= a->Cin && a->y_ld >= a->Cout, "conv3d: channel stride smaller than channel count");
  VVAE_REQUIRE(which != 0 || a->epilogue == VVAE_EPI_NONE || (a->epilogue == VVAE_EPI_RESIDUAL && a->aux_in),
               "conv3d: unsupported epilogue %d", a->epilogue);
  return VVAE_OK;
}

namespace vvae {
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
}  // namespace vvae

namespace vvae {
long long convt_tc_workspace_bytes(int b_t, int H, int W, int Cin, int Cout);
bool convt_tc_supported(int dtype, int b_t, int H, int W, int Cin, int Cout, const void* x, const void* y, long long y_ld,
                        const void* ws, long long ws_bytes);
int convt_tc_fwd(const void* x, const void* w, const float* bias, void* y, long long y_ld, int b_t, int H, int W, int Cin,
                 int Cout, void* ws, cudaStream_t s);
int convt_tc_bwd(const void* dy, long long dy_ld, const void* x, const void* w, void* dx, float* dw_accum, int b_t, int H,
                 int W, int Cin, int Cout, void* ws, cudaStream_t s);
}  // namespace vvae

static int conv_run(const vvae_conv_args* a, int which, vvae_stream_t stream) {
  int rc = conv_validate(a, which);
  if (rc) return rc;
  if (a->B == 0) return VVAE_OK;
  cudaStream_t s = as_stream(stream);
  // 1x1x1 convs with a handful of channels (the U-Net's output conv) are per-voxel streams: small_linear.cu
  if (a->backend == VVAE_BACKEND_AUTO && a->dtype == VVAE_BF16 && a->kt == 1 && a->kh == 1 && a->kw == 1 &&
      small_linear_ok(a->Cin, a->Cout)) {
    const long long V = (long long)a->B * a->T * a->H * a->W;
    if (which == 0 && !a->pad_out) {   // y[v, co] = sum_ci x[v, ci] * w[ci, co] + bias (+ residual)
      SmallLinArgs q{(const bf16*)a->x, a->x_ld, (const bf16*)a->w, a->Cout, 1, a->bias, (bf16*)a->y, a->y_ld,
                     a->epilogue == VVAE_EPI_RESIDUAL ? (const bf16*)a->aux_in : nullptr, a->ld_aux, V, a->Cin, a->Cout};
      return small_linear_fwd(q, s);
    }
    if (which == 1) {   // dx[v, ci] = sum_co dy[v, co] * w[ci, co]
      SmallLinArgs q{(const bf16*)a->y, a->y_ld, (const bf16*)a->w, 1, a->Cout, nullptr, (bf16*)const_cast<void*>(a->x),
                     a->x_ld, nullptr, 0, V, a->Cout, a->Cin};
      return small_linear_fwd(q, s);
    }
    if (which == 2 && small_linear_wgrad_ok(a->Cin, a->Cout))
      return small_linear_wgrad((const bf16*)a->x, a->x_ld, (const bf16*)a->y, a->y_ld, a->dw_accum, a->Cout, 1, V, a->Cin,
                                a->Cout, s);
  }
  if (a->backend != VVAE_BACKEND_SIMT && which == 2 && conv_wgrad_tc_supported(*a)) return conv_wgrad_tc_launch(*a, s);
  if (a->backend != VVAE_BACKEND_SIMT && which < 2 && conv_tc_supported(*a, which)) return conv_tc_launch(*a, which, s);
  if (a->backend == VVAE_BACKEND_TCGEN05) {
    set_error("conv3d: shape not supported by the tensor-core path");
    return VVAE_ERR_UNSUPPORTED;
  }
  VVAE_DISPATCH_DTYPE(a->dtype, T, return conv_simt<T>(*a, which, s));
  return VVAE_OK;
}

extern "C" {

int vvae_conv3d_fwd(const vvae_conv_args* args, vvae_stream_t stream) { return conv_run(args, 0, stream); }
int vvae_conv3d_dgrad(const vvae_conv_args* args, vvae_stream_t stream) { return conv_run(args, 1, stream); }
int vvae_conv3d_wgrad(const vvae_conv_args* args, vvae_stream_t stream) { return conv_run(args, 2, stream); }

long long vvae_convT122_workspace_bytes(int b_t, int H, int W, int Cin, int Cout) {
  return convt_tc_workspace_bytes(b_t, H, W, Cin, Cout);
}

int vvae_convT122_fwd(const void* x, const void* w, const float* bias, void* y, long long y_ld, int b_t, int H, int W,
                      int Cin, int Cout, int dtype, void* workspace, long long workspace_bytes, vvae_stream_t stream) {
  if (b_t <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && w && y && y_ld >= Cout, "convT122_fwd: bad arguments");
  if (convt_tc_supported(dtype, b_t, H, W, Cin, Cout, x, y, y_ld, workspace, workspace_bytes))
    return convt_tc_fwd(x, w, bias, y, y_ld, b_t, H, W, Cin, Cout, workspace, as_stream(stream));
  const long long V = (long long)b_t * H * W;
  CTGeom g{H, W, Cin, Cout};
  VVAE_DISPATCH_DTYPE(dtype, T,
                      return launch_gemm_simt(RowMajorLoader<T>{(const T*)x, Cin}, CTWeightLoader<T>{(const T*)w, g},
                                              CTScatterEpi<T>{(T*)y, y_ld, bias, g}, V, 4 * Cout, Cin, 1, as_stream(stream)));
  return VVAE_OK;
}

int vvae_convT122_bwd(const void* dy, long long dy_ld, const void* x, const void* w, void* dx, float* dw_accum, int b_t,
                      int H, int W, int Cin, int Cout, int dtype, void* workspace, long long workspace_bytes,
                      vvae_stream_t stream) {
  if (b_t <= 0) return VVAE_OK;
  VVAE_REQUIRE(dy && x && w && dy_ld >= Cout, "convT122_bwd: bad arguments");
  if (convt_tc_supported(dtype, b_t, H, W, Cin, Cout, x, dy, dy_ld, workspace, workspace_bytes) &&
      (!dx || ((uintptr_t)dx % 16 == 0)))
    return convt_tc_bwd(dy, dy_ld, x, w, dx, dw_accum, b_t, H, W, Cin, Cout, workspace, as_stream(stream));
  const long long V = (long long)b_t * H * W;
  CTGeom g{H, W, Cin, Cout};
  cudaStream_t s = as_stream(stream);
  VVAE_DISPATCH_DTYPE(dtype, T, {
    CTGatherLoader<T> gl{(const T*)dy, dy_ld, g};
    if (dx) {
      EpiStore<T, T> ep{(T*)dx, Cin, nullptr, VVAE_EPI_NONE, nullptr, 0, nullptr, 0, 0};
      int rc = launch_gemm_simt(gl, CTWeightTLoader<T>{CTWeightLoader<T>{(const T*)w, g}}, ep, V, Cin, 4LL * Cout, 1, s);
      if (rc) return rc;
    }
    if (dw_accum) {
      const long long tiles = cdiv(Cin, SG_BM) * cdiv(4 * Cout, SG_BN);
      int splits = (int)std::max<long long>(1, std::min<long long>(((long long)num_sms() * 8) / tiles, V / 512));
      int rc = launch_gemm_simt(ColMajorLoader<T>{(const T*)x, Cin}, CTGatherBLoader<T>{gl}, CTWgradEpi{dw_accum, g}, Cin,
                                4 * Cout, V, splits, s);
      if (rc) return rc;
    }
  });
  return VVAE_OK;
}

}  // extern "C"
