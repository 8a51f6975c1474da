// HBM-bound elementwise / reduction kernels of the hot path: patchify & pixel shuffle, max-pool, channel-slice copy,
// latent head (softplus-log, selection gate, reparameterisation + gate, fused backward with KL gradient),
// masked reconstruction / KL losses, and the fused Adam step.
#include "common.cuh"

namespace vvae {

// ---------------- Philox4x32-10 (counter-based RNG; one 128-bit block per (offset, index)) ----------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(unsigned long long seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(unsigned long long ctr_lo, unsigned long long ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// 23 random bits: (k + 0.5) / 2^23 is exact in fp32 and strictly inside (0,1) (24 bits would round k = 2^24 - 1 up to 1.0)
__device__ __forceinline__ float u01(uint32_t x) { return ((x >> 9) + 0.5f) * (1.0f / 8388608.0f); }
__device__ __forceinline__ float normal_from(uint32_t a, uint32_t b) {
  float u1 = u01(a), u2 = u01(b);
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

// The same draws the fused kernels make, materialised (kind 0: N(0,1) of the reparameterisation; kind 1: U(0,1) of the
// Gumbel gate), so a captured CUDA graph can read per-step noise from a buffer instead of baking (seed, offset) in.
__global__ void philox_fill_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset,
                                   int kind) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  Philox rng(seed);
  for (; i < n; i += stride) {
    if (kind == 0) {
      const uint4 r = rng(offset + (unsigned long long)i, 0x9a55ull);
      out[i] = normal_from(r.x, r.y);
    } else {
      out[i] = u01(rng(offset + (unsigned long long)i, 0x5e1ec7ull).x);
    }
  }
}

// ---------------- tokens <-> voxels rearrangement ----------------
// tokens [BT, (h w), (p1 p2 c)]  <->  voxels [BT, (h p1), (w p2), c];  runs of P*C elements are contiguous on both sides.
template <typename TI, typename TO>
__global__ void rearrange_kernel(const TI* __restrict__ src, TO* __restrict__ dst, long long total, int H, int W, int C,
                                 int P, int to_tokens) {
  const int hp = H / P, wp = W / P, run = P * C;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    // decode the token-side index
    int r = (int)(i % run);            // (p2, c)
    long long t1 = i / run;
    int p1 = (int)(t1 % P);
    long long t2 = t1 / P;
    int w = (int)(t2 % wp);
    long long t3 = t2 / wp;
    int h = (int)(t3 % hp);
    long long bt = t3 / hp;
    long long vox = ((bt * H + (long long)h * P + p1) * W + (long long)w * P) * C + r;
    if (to_tokens) dst[i] = from_f<TO>(to_f(src[vox]));
    else           dst[vox] = from_f<TO>(to_f(src[i]));
  }
}
// same mapping, 16 bytes per thread (same dtype both sides; run*sizeof(T) % 16 == 0)
template <typename T>
__global__ void rearrange_vec_kernel(const T* __restrict__ src, T* __restrict__ dst, long long total_vec, int H, int W,
                                     int C, int P, int to_tokens) {
  constexpr int V = Vec16<T>::N;
  const int hp = H / P, wp = W / P, runv = P * C / V;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total_vec; i += stride) {
    int r = (int)(i % runv);
    long long t1 = i / runv;
    int p1 = (int)(t1 % P);
    long long t2 = t1 / P;
    int w = (int)(t2 % wp);
    long long t3 = t2 / wp;
    int h = (int)(t3 % hp);
    long long bt = t3 / hp;
    long long vox = ((bt * H + (long long)h * P + p1) * W + (long long)w * P) * C + (long long)r * V;
    Vec16<T> v;
    if (to_tokens) { v.load(src + vox); v.store(dst + i * V); }
    else           { v.load(src + i * V); v.store(dst + vox); }
  }
}

// same mapping with a voxel-side pixel pitch vox_ld >= C (bf16, C % 4 == 0, vox_ld % 8 == 0): one thread per pixel; the
// pad channels [C, vox_ld) are written as zeros (tokens -> voxels) / ignored (voxels -> tokens).  The U-Net's tensor-core
// convolutions read 16-channel blocks, so its 12-channel input lives at a pitch of 16 from the start.
template <int CP>
__global__ void rearrange_pitched_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long pixels, int H, int W,
                                         int C, int P, int vox_ld, int to_tokens) {
  const int hp = H / P, wp = W / P;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < pixels; i += stride) {                 // i = voxel-side pixel index (bt, y, x)
    const int x = (int)(i % W);
    const long long t1 = i / W;
    const int y = (int)(t1 % H);
    const long long bt = t1 / H;
    const int h = y / P, p1 = y % P, w = x / P, p2 = x % P;
    const long long tok = ((((bt * hp + h) * wp + w) * P + p1) * P + p2) * C;
    uint2 v[CP / 4];
    if (to_tokens) {
      const bf16* sp = src + i * vox_ld;
#pragma unroll
      for (int j = 0; j < CP / 4; ++j)
        if (4 * j < C) v[j] = *reinterpret_cast<const uint2*>(sp + 4 * j);
#pragma unroll
      for (int j = 0; j < CP / 4; ++j)
        if (4 * j < C) *reinterpret_cast<uint2*>(dst + tok + 4 * j) = v[j];
    } else {
#pragma unroll
      for (int j = 0; j < CP / 4; ++j) v[j] = 4 * j < C ? *reinterpret_cast<const uint2*>(src + tok + 4 * j) : make_uint2(0u, 0u);
      bf16* dp = dst + i * vox_ld;
#pragma unroll
      for (int j = 0; j < CP / 8; ++j)
        if (8 * j < vox_ld) *reinterpret_cast<uint4*>(dp + 8 * j) = make_uint4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y);
    }
  }
}

// ---------------- max-pool (1,2,2) ----------------
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, long long x_ld, T* __restrict__ y, long long total, int H,
                                   int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c = (int)(i % C);
    long long t1 = i / C;
    int j = (int)(t1 % Wo);
    long long t2 = t1 / Wo;
    int ii = (int)(t2 % Ho);
    long long bt = t2 / Ho;
    const T* p = x + ((bt * H + 2 * ii) * W + 2 * j) * x_ld + c;
    float m = to_f(p[0]);
    m = fmaxf(m, to_f(p[x_ld]));
    m = fmaxf(m, to_f(p[(long long)W * x_ld]));
    m = fmaxf(m, to_f(p[(long long)W * x_ld + x_ld]));
    y[i] = from_f<T>(m);
  }
}
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, long long x_ld, const T* __restrict__ dy,
                                   const T* __restrict__ dskip, long long dskip_ld, T* __restrict__ dx, long long total,
                                   int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c = (int)(i % C);
    long long t1 = i / C;
    int j = (int)(t1 % Wo);
    long long t2 = t1 / Wo;
    int ii = (int)(t2 % Ho);
    long long bt = t2 / Ho;
    const long long pix = (bt * H + 2 * ii) * W + 2 * j;
    const long long offs[4] = {0, 1, (long long)W, (long long)W + 1};
    float vals[4];
    int arg = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      vals[k] = to_f(x[(pix + offs[k]) * x_ld + c]);
      if (vals[k] > vals[arg]) arg = k;  // strict: first maximum wins
    }
    const float g = to_f(dy[i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = (k == arg) ? g : 0.f;
      if (dskip) v += to_f(dskip[(pix + offs[k]) * dskip_ld + c]);
      dx[(pix + offs[k]) * C + c] = from_f<T>(v);
    }
  }
}

template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ src, long long src_ld, long long src_off, T* __restrict__ dst,
                                     long long dst_ld, long long dst_off, long long total, int C) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    long long r = i / C;
    int c = (int)(i % C);
    dst[r * dst_ld + dst_off + c] = src[r * src_ld + src_off + c];
  }
}


// ---- bf16 fast paths: 8 channels (one 16-byte vector) per thread, 32-bit index arithmetic ----
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__global__ void __launch_bounds__(256)
maxpool_fwd_vec_kernel(const bf16* __restrict__ x, long long x_ld, bf16* __restrict__ y, unsigned total_vec, int H, int W,
                       int C) {
  const unsigned cv = (unsigned)C >> 3, Wo = (unsigned)W >> 1, Ho = (unsigned)H >> 1;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += gridDim.x * blockDim.x) {
    const unsigned v = i % cv, t1 = i / cv, j = t1 % Wo, t2 = t1 / Wo, ii = t2 % Ho, bt = t2 / Ho;
    const bf16* p = x + (((long long)bt * H + 2 * ii) * W + 2 * j) * x_ld + v * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + x_ld);
    const uint4 c = *reinterpret_cast<const uint4*>(p + (long long)W * x_ld);
    const uint4 d = *reinterpret_cast<const uint4*>(p + (long long)W * x_ld + x_ld);
    uint4 o;
    o.x = bf2_max(bf2_max(a.x, b.x), bf2_max(c.x, d.x));
    o.y = bf2_max(bf2_max(a.y, b.y), bf2_max(c.y, d.y));
    o.z = bf2_max(bf2_max(a.z, b.z), bf2_max(c.z, d.z));
    o.w = bf2_max(bf2_max(a.w, b.w), bf2_max(c.w, d.w));
    *reinterpret_cast<uint4*>(y + (long long)i * 8) = o;
  }
}
__global__ void __launch_bounds__(256)
maxpool_bwd_vec_kernel(const bf16* __restrict__ x, long long x_ld, const bf16* __restrict__ dy, const bf16* __restrict__ dskip,
                       long long dskip_ld, bf16* __restrict__ dx, unsigned total_vec, int H, int W, int C) {
  const unsigned cv = (unsigned)C >> 3, Wo = (unsigned)W >> 1, Ho = (unsigned)H >> 1;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += gridDim.x * blockDim.x) {
    const unsigned v = i % cv, t1 = i / cv, j = t1 % Wo, t2 = t1 / Wo, ii = t2 % Ho, bt = t2 / Ho;
    const long long pix = ((long long)bt * H + 2 * ii) * W + 2 * j;
    const long long offs[4] = {0, 1, (long long)W, (long long)W + 1};
    Vec16<bf16> xv[4], g, sk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xv[k].load(x + (pix + offs[k]) * x_ld + v * 8);
      if (dskip) sk[k].load(dskip + (pix + offs[k]) * dskip_ld + v * 8);
    }
    g.load(dy + (long long)i * 8);
    Vec16<bf16> o[4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      int arg = 0;
      float best = xv[0].get(t);
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const float f = xv[k].get(t);
        if (f > best) { best = f; arg = k; }      // strict: the first maximum wins
      }
      const float gg = g.get(t);
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k].set(t, (k == arg ? gg : 0.f) + (dskip ? sk[k].get(t) : 0.f));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k].store(dx + (pix + offs[k]) * C + v * 8);
  }
}
__global__ void __launch_bounds__(256)
copy_channels_vec_kernel(const bf16* __restrict__ src, long long src_ld, bf16* __restrict__ dst, long long dst_ld,
                         unsigned total_vec, int C) {
  const unsigned cv = (unsigned)C >> 3;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += gridDim.x * blockDim.x) {
    const unsigned v = i % cv, r = i / cv;
    *reinterpret_cast<uint4*>(dst + (long long)r * dst_ld + v * 8) =
        *reinterpret_cast<const uint4*>(src + (long long)r * src_ld + v * 8);
  }
}

// ---------------- latent head ----------------
__device__ __forceinline__ float softplusf_(float a) { return a > 20.f ? a : log1pf(__expf(a)); }

template <typename T>
__global__ void softplus_log_fwd_kernel(const T* __restrict__ a, T* __restrict__ lv, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) lv[i] = from_f<T>(__logf(round_to<T>(softplusf_(to_f(a[i])))));
}
template <typename T>
__global__ void softplus_log_bwd_kernel(const T* __restrict__ dlv, const T* __restrict__ a, T* __restrict__ da,
                                        long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float x = to_f(a[i]);
    da[i] = from_f<T>(to_f(dlv[i]) * sigmoidf_(x) / softplusf_(x));
  }
}

// one block per frame: logit = s1[frame,:] . w2 + b2 + 1 ; p = sigmoid((logit + logistic noise)/temp) ; sel = rint(p)
template <typename T>
__global__ void selection_fwd_kernel(const T* __restrict__ s1, const float* __restrict__ w2, const float* __restrict__ b2,
                                     const float* __restrict__ u, unsigned long long seed, unsigned long long offset,
                                     int train, float temperature, float* __restrict__ logit, float* __restrict__ p_out,
                                     float* __restrict__ sel, int hw) {
  __shared__ float scratch[33];
  const int f = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) acc += to_f(s1[(long long)f * hw + i]) * round_to<T>(w2[i]);
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    float lg = round_to<T>(round_to<T>(round_to<T>(acc) + round_to<T>(b2[0])) + 1.f);
    float z = lg;
    if (train) {
      float uu = u ? u[f] : u01(Philox(seed)(offset + (unsigned long long)f, 0x5e1ec7ull).x);
      uu = fminf(fmaxf(uu, 1e-20f), 1.0f - 5.9604645e-8f);   // largest fp32 below 1 (1 - 1e-20 rounds to 1.0f)
      z += __logf(uu / (1.f - uu));
    }
    float p = 1.f / (1.f + expf(-z / temperature));
    logit[f] = lg;
    p_out[f] = p;
    sel[f] = rintf(p);  // round half to even, as jnp.round
  }
}

template <typename T>
__global__ void reparam_gate_fwd_kernel(const T* __restrict__ mean, const T* __restrict__ logvar,
                                        const float* __restrict__ eps, unsigned long long seed,
                                        unsigned long long offset, float* __restrict__ eps_out,
                                        const float* __restrict__ sel, const float* __restrict__ fill,
                                        float* __restrict__ c32, T* __restrict__ cT, long long n, int tok_per_frame,
                                        int Dl, int train) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  Philox rng(seed);
  for (; i < n; i += stride) {
    const long long tok = i / Dl;
    const int d = (int)(i % Dl);
    const float s = sel[tok / tok_per_frame];
    float z = to_f(mean[i]);
    if (train) {
      float e;
      if (eps) e = eps[i];
      else {
        uint4 r = rng(offset + (unsigned long long)i, 0x9a55ull);
        e = normal_from(r.x, r.y);
      }
      if (eps_out) eps_out[i] = e;
      float sd = round_to<T>(__expf(round_to<T>(to_f(logvar[i]) * 0.5f)));
      z += e * sd;
    }
    float c = fill[d] * (1.f - s) + z * s;
    if (c32) c32[i] = c;
    if (cT) cT[i] = from_f<T>(c);
  }
}

// blockDim.x is a multiple of Dl: a thread keeps one latent channel d, walks the tokens of one frame slab.
template <typename T>
__global__ void reparam_gate_bwd_kernel(const T* __restrict__ dc, const T* __restrict__ mean,
                                        const T* __restrict__ logvar, const float* __restrict__ eps,
                                        const float* __restrict__ sel, const float* __restrict__ fill,
                                        const T* dmean_in, const T* dlogvar_in, T* dmean, T* dlogvar,
                                        float* __restrict__ dfill, float* __restrict__ dsel, int tok_per_frame, int Dl,
                                        int train, int slabs) {
  __shared__ float scratch[33];
  const int frame = blockIdx.x / slabs, slab = blockIdx.x % slabs;
  const int d = threadIdx.x % Dl, r = threadIdx.x / Dl, rpi = blockDim.x / Dl;
  const int per_slab = (tok_per_frame + slabs - 1) / slabs;
  const int t0 = slab * per_slab, t1 = min(t0 + per_slab, tok_per_frame);
  const float s = sel[frame], f = fill[d];
  float afill = 0.f, asel = 0.f;
  for (int t = t0 + r; t < t1; t += rpi) {
    const long long i = ((long long)frame * tok_per_frame + t) * Dl + d;
    const float g = to_f(dc[i]);
    const float mu = to_f(mean[i]), lv = to_f(logvar[i]);
    float z = mu, dlv = dlogvar_in ? to_f(dlogvar_in[i]) : 0.f;
    if (train) {
      const float sd = round_to<T>(__expf(round_to<T>(lv * 0.5f)));
      const float e = eps[i];
      z += e * sd;
      dlv += g * s * e * 0.5f * sd;
    }
    dmean[i] = from_f<T>(g * s + (dmean_in ? to_f(dmean_in[i]) : 0.f));
    dlogvar[i] = from_f<T>(dlv);
    afill += g * (1.f - s);
    asel += g * (z - f);
  }
  if (dfill) atomicAdd(dfill + d, afill);  // few blocks per channel; contention is negligible
  asel = block_sum(asel, scratch);
  if (threadIdx.x == 0 && dsel) atomicAdd(dsel + frame, asel);
}

// ---------------- losses ----------------
// one thread per (b, p) position: time-sum in fp32, rounded to T (XLA reduces bf16 with fp32 accumulation), / len.
// PS (per-sample, rl_nonadversarial.py:114-121): blockIdx.y is the sample, out2 is [B,2].
template <typename TV, typename T, bool PS>
__global__ void recon_loss_fwd_kernel(const TV* __restrict__ video, const T* __restrict__ recon,
                                      const float* __restrict__ fmask, const float* __restrict__ inv_len,
                                      float* __restrict__ out2, int B, int Tn, long long per_frame) {
  __shared__ float scratch[33];
  const long long total = PS ? per_frame : (long long)B * per_frame;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float a_sq = 0.f, a_ab = 0.f;
  for (; i < total; i += stride) {
    const int b = PS ? (int)blockIdx.y : (int)(i / per_frame);
    const long long p = PS ? i : i % per_frame;
    float ssq = 0.f, sab = 0.f;
    for (int t = 0; t < Tn; ++t) {
      const float m = fmask[b * Tn + t];
      if (m == 0.f) continue;
      const long long idx = ((long long)b * Tn + t) * per_frame + p;
      float e = round_to<T>(round_to<T>(to_f(video[idx])) - to_f(recon[idx]));
      ssq += round_to<T>(e * e);
      sab += fabsf(e);
    }
    a_sq += round_to<T>(ssq) * inv_len[b];
    a_ab += round_to<T>(sab) * inv_len[b];
  }
  a_sq = block_sum(a_sq, scratch);
  a_ab = block_sum(a_ab, scratch);
  if (threadIdx.x == 0) {
    float* o = PS ? out2 + 2 * blockIdx.y : out2;
    atomicAdd(o, a_sq);
    atomicAdd(o + 1, a_ab);
  }
}
template <typename TV, typename T>
__global__ void recon_loss_bwd_kernel(const TV* __restrict__ video, const T* __restrict__ recon,
                                      const float* __restrict__ fmask, const float* __restrict__ inv_len, float w_mse,
                                      float w_mae, float inv_count, const float* __restrict__ gscale,
                                      T* __restrict__ drecon, int Tn, long long per_frame, long long total) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (gscale) inv_count *= __ldg(gscale);   // upstream d(loss) as a device scalar: no host round trip
  for (; i < total; i += stride) {
    const long long frame = i / per_frame;
    const float m = fmask[frame];
    float g = 0.f;
    if (m != 0.f) {
      float e = round_to<T>(to_f(video[i])) - to_f(recon[i]);
      float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
      g = -(2.f * w_mse * e + w_mae * sg) * inv_len[frame / Tn] * inv_count;
    }
    drecon[i] = from_f<T>(g);
  }
}
// PS: blockIdx.y is the sample (n = elements per sample), out1 is [B].
template <typename T, bool PS>
__global__ void kl_fwd_kernel(const T* __restrict__ mean, const T* __restrict__ logvar, const float* __restrict__ frame_w,
                              float* __restrict__ out1, long long n, long long per_frame_el) {
  __shared__ float scratch[33];
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long base = PS ? blockIdx.y * n : 0;
  float acc = 0.f;
  for (; i < n; i += stride) {
    const float w = frame_w[(base + i) / per_frame_el];
    if (w == 0.f) continue;
    const float mu = to_f(mean[base + i]), lv = to_f(logvar[base + i]);
    acc += round_to<T>(0.5f * (__expf(lv) - 1.f - lv + mu * mu)) * w;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(PS ? out1 + blockIdx.y : out1, acc);
}

template <typename T>
__global__ void kl_bwd_kernel(const T* __restrict__ mean, const T* __restrict__ logvar, const float* __restrict__ frame_w,
                              float scale, const float* __restrict__ gscale, T* __restrict__ dmean,
                              T* __restrict__ dlogvar, long long n, long long per_frame_el) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (gscale) scale *= __ldg(gscale);
  for (; i < n; i += stride) {
    const float w = frame_w[i / per_frame_el] * scale;
    dmean[i] = from_f<T>(w * to_f(mean[i]));
    dlogvar[i] = from_f<T>(w * 0.5f * (__expf(to_f(logvar[i])) - 1.f));
  }
}

// ---------------- VGG-16 perceptual features (train/vgg_tests.py:8-68; flaxmodels 0.1.3 VGG16) ----------------
// ReLU after a conv (bias already added by the conv kernel) and its backward through the saved output.
template <typename T>
__global__ void relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n) {
  constexpr int V = Vec16<T>::N;
  const long long nv = n / V;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = i; j < nv; j += stride) {
    Vec16<T> v;
    v.load(x + j * V);
#pragma unroll
    for (int t = 0; t < V; ++t) v.set(t, fmaxf(v.get(t), 0.f));
    v.store(y + j * V);
  }
  for (long long j = nv * V + i; j < n; j += stride) y[j] = from_f<T>(fmaxf(to_f(x[j]), 0.f));
}
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long long n) {
  constexpr int V = Vec16<T>::N;
  const long long nv = n / V;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = i; j < nv; j += stride) {
    Vec16<T> a, b;
    a.load(dy + j * V);
    b.load(y + j * V);
#pragma unroll
    for (int t = 0; t < V; ++t) a.set(t, b.get(t) > 0.f ? a.get(t) : 0.f);
    a.store(dx + j * V);
  }
  for (long long j = nv * V + i; j < n; j += stride) dx[j] = to_f(y[j]) > 0.f ? dy[j] : from_f<T>(0.f);
}
// ImageNet normalisation of an RGB frame into a 16-channel (zero-padded) tensor the tensor-core conv can gather.
__constant__ float c_vgg_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_vgg_std[3] = {0.229f, 0.224f, 0.225f};
template <typename TI, typename T>
__global__ void vgg_pre_fwd_kernel(const TI* __restrict__ x, T* __restrict__ y, long long V, int ld) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < V * ld; i += stride) {
    const long long v = i / ld;
    const int c = (int)(i - v * ld);
    float o = 0.f;
    if (c < 3) o = (round_to<T>(to_f(x[v * 3 + c])) - c_vgg_mean[c]) / c_vgg_std[c];
    y[i] = from_f<T>(o);
  }
}
template <typename T>
__global__ void vgg_pre_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, long long V, int ld) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < V * 3; i += stride) {
    const long long v = i / 3;
    const int c = (int)(i - v * 3);
    dx[i] = from_f<T>(to_f(dy[v * ld + c]) / c_vgg_std[c]);
  }
}

// ---------------- optimizer ----------------
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float scratch[33];
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i < n; i += stride) acc += g[i] * g[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
// deterministic form: per-block partials, then ONE block adds them in a fixed order (same bits on every rank for the
// same input, so the clip scale -- and with it the replicas -- stay bit-identical)
__global__ void sumsq_partials_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials) {
  __shared__ float scratch[33];
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i < n; i += stride) acc += g[i] * g[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
__global__ void sumsq_final_kernel(const float* __restrict__ partials, int count, float* __restrict__ out) {
  __shared__ float scratch[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partials[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) *out += acc;
}
// One Adam element (optax.scale_by_adam + scale(-lr) after clip_by_global_norm); every kernel below goes through it so
// that the scalar and the vectorised paths are bit-identical.
__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float scale, float lr, float b1, float b2,
                                          float eps, float bc1, float bc2) {
  // explicit roundings: the compiler may not contract these differently in the scalar and the 16-byte kernel
  const float gi = __fmul_rn(g, scale);
  const float mi = __fmaf_rn(b1, m, __fmul_rn(1.f - b1, gi));
  const float vi = __fmaf_rn(b2, v, __fmul_rn(__fmul_rn(1.f - b2, gi), gi));
  m = mi;
  v = vi;
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr, __fdiv_rn(mi, bc1)), __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, bc2)), eps)));
  return p;
}
__device__ __forceinline__ float adam_clip_scale(const float* gnorm_sq, float clip, float grad_scale) {
  float scale = grad_scale;
  if (gnorm_sq) {
    float gn = sqrtf(*gnorm_sq) * grad_scale;
    if (gn > clip) scale *= clip / gn;  // optax.clip_by_global_norm
  }
  return scale;
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, bf16* __restrict__ shadow, long long n, float lr, float b1, float b2,
                            float eps, float bc1, float bc2, const float* __restrict__ gnorm_sq, float clip,
                            float grad_scale) {
  const float scale = adam_clip_scale(gnorm_sq, clip, grad_scale);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_one(pi, g[i], mi, vi, scale, lr, b1, b2, eps, bc1, bc2);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (shadow) shadow[i] = __float2bfloat16(pi);
  }
}
// 16-byte version: four parameters per thread and pass, two passes in flight; streams 7 x 4 B (+ 2 B of bf16 shadow) per
// parameter, so what matters is bytes in flight: 3 x 2 x 16 B of independent loads per thread before the first use.
// The bf16 shadow (the copy of the weights the bf16 kernels read) is written here instead of by a separate cast pass
// over the freshly written parameters.
__global__ void __launch_bounds__(256)
adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                 uint2* __restrict__ shadow, long long n4, float lr, float b1, float b2, float eps, float bc1, float bc2,
                 const float* __restrict__ gnorm_sq, float clip, float grad_scale) {
  const float scale = adam_clip_scale(gnorm_sq, clip, grad_scale);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  auto apply = [&](long long j, float4 pj, float4 gj, float4 mj, float4 vj) {
    adam_one(pj.x, gj.x, mj.x, vj.x, scale, lr, b1, b2, eps, bc1, bc2);
    adam_one(pj.y, gj.y, mj.y, vj.y, scale, lr, b1, b2, eps, bc1, bc2);
    adam_one(pj.z, gj.z, mj.z, vj.z, scale, lr, b1, b2, eps, bc1, bc2);
    adam_one(pj.w, gj.w, mj.w, vj.w, scale, lr, b1, b2, eps, bc1, bc2);
    p[j] = pj; m[j] = mj; v[j] = vj;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pj.x, pj.y), hi = __floats2bfloat162_rn(pj.z, pj.w);
      shadow[j] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  };
  for (; i + stride < n4; i += 2 * stride) {
    const long long j = i + stride;
    const float4 g0 = g[i], g1 = g[j], p0 = p[i], p1 = p[j], m0 = m[i], m1 = m[j], v0 = v[i], v1 = v[j];
    apply(i, p0, g0, m0, v0);
    apply(j, p1, g1, m1, v1);
  }
  if (i < n4) apply(i, p[i], g[i], m[i], v[i]);
}

static inline int ew_blocks(long long n, int threads = 256) { return (int)std::min<long long>(cdiv(n, threads), (long long)num_sms() * 16); }

}  // namespace vvae

using namespace vvae;

template <typename TI, typename TO>
static int rearrange_launch(const void* src, void* dst, long long total, int H, int W, int C, int P, int to_tokens,
                            cudaStream_t s) {
  rearrange_kernel<TI, TO><<<ew_blocks(total), 256, 0, s>>>((const TI*)src, (TO*)dst, total, H, W, C, P, to_tokens);
  return check_launch("rearrange");
}
template <typename T>
static int rearrange_same(const void* src, void* dst, long long total, int H, int W, int C, int P, int to_tokens,
                          cudaStream_t s) {
  constexpr int V = Vec16<T>::N;
  if ((P * C) % V == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) {
    rearrange_vec_kernel<T><<<ew_blocks(total / V), 256, 0, s>>>((const T*)src, (T*)dst, total / V, H, W, C, P, to_tokens);
    return check_launch("rearrange_vec");
  }
  return rearrange_launch<T, T>(src, dst, total, H, W, C, P, to_tokens, s);
}

extern "C" {

int vvae_patchify(const void* video, int in_dtype, void* tokens, int b_t, int H, int W, int C, int P, int dtype,
                  vvae_stream_t stream) {
  if (b_t <= 0) return VVAE_OK;
  VVAE_REQUIRE(video && tokens && P > 0 && H % P == 0 && W % P == 0, "patchify: bad arguments (H=%d W=%d P=%d)", H, W, P);
  const long long total = (long long)b_t * H * W * C;
  cudaStream_t s = as_stream(stream);
  if (in_dtype == dtype) {
    VVAE_DISPATCH_DTYPE(dtype, T, return rearrange_same<T>(video, tokens, total, H, W, C, P, 1, s));
  }
  if (in_dtype == VVAE_F32 && dtype == VVAE_BF16) return rearrange_launch<float, bf16>(video, tokens, total, H, W, C, P, 1, s);
  if (in_dtype == VVAE_BF16 && dtype == VVAE_F32) return rearrange_launch<bf16, float>(video, tokens, total, H, W, C, P, 1, s);
  VVAE_REQUIRE(false, "patchify: bad dtypes %d -> %d", in_dtype, dtype);
  return VVAE_ERR_INVALID;
}

int vvae_pixel_shuffle(const void* src, void* dst, int b_t, int H, int W, int CU, int P, int dir, int dtype,
                       vvae_stream_t stream) {
  if (b_t <= 0) return VVAE_OK;
  VVAE_REQUIRE(src && dst && P > 0 && H % P == 0 && W % P == 0, "pixel_shuffle: bad arguments");
  const long long total = (long long)b_t * H * W * CU;
  VVAE_DISPATCH_DTYPE(dtype, T, return rearrange_same<T>(src, dst, total, H, W, CU, P, dir, as_stream(stream)));
  return VVAE_OK;
}

int vvae_pixel_shuffle_pitched(const void* src, void* dst, int b_t, int H, int W, int CU, int P, int dir, long long vox_ld,
                               int dtype, vvae_stream_t stream) {
  if (b_t <= 0) return VVAE_OK;
  VVAE_REQUIRE(src && dst && P > 0 && H % P == 0 && W % P == 0 && vox_ld >= CU, "pixel_shuffle_pitched: bad arguments");
  if (vox_ld == CU) return vvae_pixel_shuffle(src, dst, b_t, H, W, CU, P, dir, dtype, stream);
  VVAE_REQUIRE(dtype == VVAE_BF16 && CU % 4 == 0 && vox_ld % 8 == 0 && vox_ld <= 32 && vox_ld - CU < 8 + (CU % 8) &&
                   ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0),
               "pixel_shuffle_pitched: bf16, channels %% 4 == 0 and a pitch of ceil8..ceil16(channels) <= 32 expected "
               "(CU=%d pitch=%lld)", CU, vox_ld);
  const long long pixels = (long long)b_t * H * W;
  const int blocks = ew_blocks(pixels);
  cudaStream_t s = as_stream(stream);
  if (vox_ld <= 16)
    rearrange_pitched_kernel<16><<<blocks, 256, 0, s>>>((const bf16*)src, (bf16*)dst, pixels, H, W, CU, P, (int)vox_ld, dir);
  else
    rearrange_pitched_kernel<32><<<blocks, 256, 0, s>>>((const bf16*)src, (bf16*)dst, pixels, H, W, CU, P, (int)vox_ld, dir);
  return check_launch("pixel_shuffle_pitched");
}

int vvae_maxpool122_fwd(const void* x, long long x_ld, void* y, int b_t, int H, int W, int C, int dtype,
                        vvae_stream_t stream) {
  if (b_t <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && y && H % 2 == 0 && W % 2 == 0 && x_ld >= C, "maxpool122_fwd: bad arguments");
  const long long total = (long long)b_t * (H / 2) * (W / 2) * C;
  if (dtype == VVAE_BF16 && C % 8 == 0 && x_ld % 8 == 0 && total / 8 < (1LL << 31) && ((uintptr_t)x % 16 == 0) &&
      ((uintptr_t)y % 16 == 0)) {
    maxpool_fwd_vec_kernel<<<ew_blocks(total / 8), 256, 0, as_stream(stream)>>>((const bf16*)x, x_ld, (bf16*)y,
                                                                               (unsigned)(total / 8), H, W, C);
    return check_launch("maxpool_fwd");
  }
  VVAE_DISPATCH_DTYPE(dtype, T, (maxpool_fwd_kernel<T><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
                                    (const T*)x, x_ld, (T*)y, total, H, W, C)));
  return check_launch("maxpool_fwd");
}

int vvae_maxpool122_bwd(const void* x, long long x_ld, const void* dy, const void* dskip, long long dskip_ld, void* dx,
                        int b_t, int H, int W, int C, int dtype, vvae_stream_t stream) {
  if (b_t <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && dy && dx && H % 2 == 0 && W % 2 == 0 && x_ld >= C, "maxpool122_bwd: bad arguments");
  const long long total = (long long)b_t * (H / 2) * (W / 2) * C;
  if (dtype == VVAE_BF16 && C % 8 == 0 && x_ld % 8 == 0 && total / 8 < (1LL << 31) && ((uintptr_t)x % 16 == 0) &&
      ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dx % 16 == 0) &&
      (!dskip || (dskip_ld % 8 == 0 && ((uintptr_t)dskip % 16 == 0)))) {
    maxpool_bwd_vec_kernel<<<ew_blocks(total / 8), 256, 0, as_stream(stream)>>>(
        (const bf16*)x, x_ld, (const bf16*)dy, (const bf16*)dskip, dskip_ld, (bf16*)dx, (unsigned)(total / 8), H, W, C);
    return check_launch("maxpool_bwd");
  }
  VVAE_DISPATCH_DTYPE(dtype, T, (maxpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
                                    (const T*)x, x_ld, (const T*)dy, (const T*)dskip, dskip_ld, (T*)dx, total, H, W, C)));
  return check_launch("maxpool_bwd");
}

int vvae_copy_channels(const void* src, long long src_ld, long long src_off, void* dst, long long dst_ld,
                       long long dst_off, long long rows, int C, int dtype, vvae_stream_t stream) {
  if (rows <= 0 || C <= 0) return VVAE_OK;
  VVAE_REQUIRE(src && dst, "copy_channels: null pointer");
  const long long total = rows * C;
  if (dtype == VVAE_BF16 && C % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0 && src_off % 8 == 0 && dst_off % 8 == 0 &&
      total / 8 < (1LL << 31) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) {
    copy_channels_vec_kernel<<<ew_blocks(total / 8), 256, 0, as_stream(stream)>>>(
        (const bf16*)src + src_off, src_ld, (bf16*)dst + dst_off, dst_ld, (unsigned)(total / 8), C);
    return check_launch("copy_channels");
  }
  VVAE_DISPATCH_DTYPE(dtype, T, (copy_channels_kernel<T><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
                                    (const T*)src, src_ld, src_off, (T*)dst, dst_ld, dst_off, total, C)));
  return check_launch("copy_channels");
}

int vvae_softplus_log_fwd(const void* a, void* lv, long long n, int dtype, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(a && lv, "softplus_log_fwd: null pointer");
  VVAE_DISPATCH_DTYPE(dtype, T, (softplus_log_fwd_kernel<T><<<ew_blocks(n), 256, 0, as_stream(stream)>>>((const T*)a, (T*)lv, n)));
  return check_launch("softplus_log_fwd");
}

int vvae_softplus_log_bwd(const void* dlv, const void* a, void* da, long long n, int dtype, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(dlv && a && da, "softplus_log_bwd: null pointer");
  VVAE_DISPATCH_DTYPE(dtype, T, (softplus_log_bwd_kernel<T><<<ew_blocks(n), 256, 0, as_stream(stream)>>>(
                                    (const T*)dlv, (const T*)a, (T*)da, n)));
  return check_launch("softplus_log_bwd");
}

int vvae_selection_fwd(const void* s1, const float* w2, const float* b2, const float* u, unsigned long long seed,
                       unsigned long long offset, int train, float temperature, float* logit, float* p, float* sel,
                       int bt, int hw, int dtype, vvae_stream_t stream) {
  if (bt <= 0) return VVAE_OK;
  VVAE_REQUIRE(s1 && w2 && b2 && logit && p && sel && hw > 0 && temperature > 0.f, "selection_fwd: bad arguments");
  VVAE_DISPATCH_DTYPE(dtype, T, (selection_fwd_kernel<T><<<bt, 128, 0, as_stream(stream)>>>(
                                    (const T*)s1, w2, b2, u, seed, offset, train, temperature, logit, p, sel, hw)));
  return check_launch("selection_fwd");
}

int vvae_reparam_gate_fwd(const void* mean, const void* logvar, const float* eps, unsigned long long seed,
                          unsigned long long offset, float* eps_out, const float* sel, const float* fill, float* c32,
                          void* cT, long long n_tok, int tok_per_frame, int Dl, int train, int dtype,
                          vvae_stream_t stream) {
  if (n_tok <= 0) return VVAE_OK;
  VVAE_REQUIRE(mean && logvar && sel && fill && (c32 || cT) && tok_per_frame > 0 && Dl > 0, "reparam_gate_fwd: bad arguments");
  const long long n = n_tok * Dl;
  VVAE_DISPATCH_DTYPE(dtype, T, (reparam_gate_fwd_kernel<T><<<ew_blocks(n), 256, 0, as_stream(stream)>>>(
                                    (const T*)mean, (const T*)logvar, eps, seed, offset, eps_out, sel, fill, c32, (T*)cT, n,
                                    tok_per_frame, Dl, train)));
  return check_launch("reparam_gate_fwd");
}

int vvae_reparam_gate_bwd(const void* dc, const void* mean, const void* logvar, const float* eps, const float* sel,
                          const float* fill, const void* dmean_in, const void* dlogvar_in, void* dmean, void* dlogvar,
                          float* dfill, float* dsel, long long n_tok, int tok_per_frame, int Dl, int train, int dtype,
                          vvae_stream_t stream) {
  if (n_tok <= 0) return VVAE_OK;
  VVAE_REQUIRE(dc && mean && logvar && sel && fill && dmean && dlogvar && (!train || eps), "reparam_gate_bwd: null pointer");
  VVAE_REQUIRE(Dl > 0 && Dl <= 1024 && n_tok % tok_per_frame == 0, "reparam_gate_bwd: bad extents");
  const int frames = (int)(n_tok / tok_per_frame);
  const int threads = Dl * std::max(1, 256 / Dl);
  const int slabs = (int)std::max<long long>(1, std::min<long long>(tok_per_frame / (threads / Dl), ((long long)num_sms() * 4) / frames));
  VVAE_DISPATCH_DTYPE(dtype, T, (reparam_gate_bwd_kernel<T><<<frames * slabs, threads, 0, as_stream(stream)>>>(
                                    (const T*)dc, (const T*)mean, (const T*)logvar, eps, sel, fill, (const T*)dmean_in,
                                    (const T*)dlogvar_in, (T*)dmean, (T*)dlogvar, dfill, dsel, tok_per_frame, Dl, train,
                                    slabs)));
  return check_launch("reparam_gate_bwd");
}

static int recon_loss_fwd_impl(const void* video, int video_dtype, const void* recon, const float* frame_mask,
                               const float* inv_len, float* out, int B, int T, long long per_frame, int dtype,
                               bool per_sample, vvae_stream_t stream) {
  if (B <= 0) return VVAE_OK;
  VVAE_REQUIRE(video && recon && frame_mask && inv_len && out, "recon_loss_fwd: null pointer");
  VVAE_REQUIRE(!per_sample || B <= 65535, "recon_loss_per_sample_fwd: B > 65535");
  cudaStream_t s = as_stream(stream);
  dim3 grid(per_sample ? (unsigned)std::max(1, ew_blocks(per_frame) / std::min(B, 8)) : (unsigned)ew_blocks((long long)B * per_frame),
            per_sample ? (unsigned)B : 1u);
#define RL_FWD(TV, TT)                                                                                                 \
  do {                                                                                                                 \
    if (per_sample)                                                                                                    \
      recon_loss_fwd_kernel<TV, TT, true><<<grid, 256, 0, s>>>((const TV*)video, (const TT*)recon, frame_mask, inv_len, \
                                                               out, B, T, per_frame);                                  \
    else                                                                                                               \
      recon_loss_fwd_kernel<TV, TT, false><<<grid, 256, 0, s>>>((const TV*)video, (const TT*)recon, frame_mask,        \
                                                                inv_len, out, B, T, per_frame);                        \
  } while (0)
  if (video_dtype == VVAE_F32 && dtype == VVAE_F32) RL_FWD(float, float);
  else if (video_dtype == VVAE_F32 && dtype == VVAE_BF16) RL_FWD(float, bf16);
  else if (video_dtype == VVAE_BF16 && dtype == VVAE_BF16) RL_FWD(bf16, bf16);
  else if (video_dtype == VVAE_BF16 && dtype == VVAE_F32) RL_FWD(bf16, float);
  else VVAE_REQUIRE(false, "recon_loss_fwd: bad dtypes");
#undef RL_FWD
  return check_launch("recon_loss_fwd");
}

int vvae_recon_loss_fwd(const void* video, int video_dtype, const void* recon, const float* frame_mask,
                        const float* inv_len, float* out2, int B, int T, long long per_frame, int dtype,
                        vvae_stream_t stream) {
  return recon_loss_fwd_impl(video, video_dtype, recon, frame_mask, inv_len, out2, B, T, per_frame, dtype, false, stream);
}

int vvae_recon_loss_per_sample_fwd(const void* video, int video_dtype, const void* recon, const float* frame_mask,
                                   const float* inv_len, float* out_b2, int B, int T, long long per_frame, int dtype,
                                   vvae_stream_t stream) {
  return recon_loss_fwd_impl(video, video_dtype, recon, frame_mask, inv_len, out_b2, B, T, per_frame, dtype, true, stream);
}

int vvae_recon_loss_bwd(const void* video, int video_dtype, const void* recon, const float* frame_mask,
                        const float* inv_len, float w_mse, float w_mae, float inv_count, const float* gscale, void* drecon, int B, int T,
                        long long per_frame, int dtype, vvae_stream_t stream) {
  if (B <= 0) return VVAE_OK;
  VVAE_REQUIRE(video && recon && frame_mask && inv_len && drecon, "recon_loss_bwd: null pointer");
  const long long total = (long long)B * T * per_frame;
  const int blocks = ew_blocks(total);
  cudaStream_t s = as_stream(stream);
#define RL_BWD(TV, TT) recon_loss_bwd_kernel<TV, TT><<<blocks, 256, 0, s>>>((const TV*)video, (const TT*)recon, frame_mask, inv_len, w_mse, w_mae, inv_count, gscale, (TT*)drecon, T, per_frame, total)
  if (video_dtype == VVAE_F32 && dtype == VVAE_F32) RL_BWD(float, float);
  else if (video_dtype == VVAE_F32 && dtype == VVAE_BF16) RL_BWD(float, bf16);
  else if (video_dtype == VVAE_BF16 && dtype == VVAE_BF16) RL_BWD(bf16, bf16);
  else if (video_dtype == VVAE_BF16 && dtype == VVAE_F32) RL_BWD(bf16, float);
  else VVAE_REQUIRE(false, "recon_loss_bwd: bad dtypes");
#undef RL_BWD
  return check_launch("recon_loss_bwd");
}

int vvae_kl_fwd(const void* mean, const void* logvar, const float* frame_w, float* out1, long long n_tok,
                int tok_per_frame, int Dl, int dtype, vvae_stream_t stream) {
  if (n_tok <= 0) return VVAE_OK;
  VVAE_REQUIRE(mean && logvar && frame_w && out1, "kl_fwd: null pointer");
  const long long n = n_tok * Dl;
  VVAE_DISPATCH_DTYPE(dtype, T, (kl_fwd_kernel<T, false><<<ew_blocks(n), 256, 0, as_stream(stream)>>>(
                                    (const T*)mean, (const T*)logvar, frame_w, out1, n, (long long)tok_per_frame * Dl)));
  return check_launch("kl_fwd");
}

int vvae_kl_per_sample_fwd(const void* mean, const void* logvar, const float* frame_w, float* out_b, int B,
                           long long tok_per_sample, int tok_per_frame, int Dl, int dtype, vvae_stream_t stream) {
  if (B <= 0 || tok_per_sample <= 0) return VVAE_OK;
  VVAE_REQUIRE(mean && logvar && frame_w && out_b, "kl_per_sample_fwd: null pointer");
  VVAE_REQUIRE(B <= 65535, "kl_per_sample_fwd: B > 65535");
  const long long n = tok_per_sample * Dl;
  dim3 grid((unsigned)std::max(1, ew_blocks(n) / std::min(B, 8)), (unsigned)B);
  VVAE_DISPATCH_DTYPE(dtype, T, (kl_fwd_kernel<T, true><<<grid, 256, 0, as_stream(stream)>>>(
                                    (const T*)mean, (const T*)logvar, frame_w, out_b, n, (long long)tok_per_frame * Dl)));
  return check_launch("kl_per_sample_fwd");
}

int vvae_kl_bwd(const void* mean, const void* logvar, const float* frame_w, float scale, const float* gscale, void* dmean,
                void* dlogvar,
                long long n_tok, int tok_per_frame, int Dl, int dtype, vvae_stream_t stream) {
  if (n_tok <= 0) return VVAE_OK;
  VVAE_REQUIRE(mean && logvar && frame_w && dmean && dlogvar, "kl_bwd: null pointer");
  const long long n = n_tok * Dl;
  VVAE_DISPATCH_DTYPE(dtype, T, (kl_bwd_kernel<T><<<ew_blocks(n), 256, 0, as_stream(stream)>>>(
                                    (const T*)mean, (const T*)logvar, frame_w, scale, gscale, (T*)dmean, (T*)dlogvar, n,
                                    (long long)tok_per_frame * Dl)));
  return check_launch("kl_bwd");
}

int vvae_relu_fwd(const void* x, void* y, long long n, int dtype, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && y, "relu_fwd: null pointer");
  VVAE_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), "relu_fwd: pointers must be 16-byte aligned");
  VVAE_DISPATCH_DTYPE(dtype, T, (relu_fwd_kernel<T><<<ew_blocks(n / 8 + 1), 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, n)));
  return check_launch("relu_fwd");
}

int vvae_relu_bwd(const void* dy, const void* y, void* dx, long long n, int dtype, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(dy && y && dx, "relu_bwd: null pointer");
  VVAE_REQUIRE(((uintptr_t)dy % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)dx % 16 == 0),
               "relu_bwd: pointers must be 16-byte aligned");
  VVAE_DISPATCH_DTYPE(dtype, T, (relu_bwd_kernel<T><<<ew_blocks(n / 8 + 1), 256, 0, as_stream(stream)>>>(
                                    (const T*)dy, (const T*)y, (T*)dx, n)));
  return check_launch("relu_bwd");
}

int vvae_vgg_preprocess_fwd(const void* x, int x_dtype, void* y, long long voxels, int y_ld, int dtype,
                            vvae_stream_t stream) {
  if (voxels <= 0) return VVAE_OK;
  VVAE_REQUIRE(x && y && y_ld >= 3, "vgg_preprocess_fwd: bad arguments");
  cudaStream_t s = as_stream(stream);
  const int blocks = ew_blocks(voxels * y_ld);
  if (x_dtype == VVAE_F32 && dtype == VVAE_F32) vgg_pre_fwd_kernel<float, float><<<blocks, 256, 0, s>>>((const float*)x, (float*)y, voxels, y_ld);
  else if (x_dtype == VVAE_F32 && dtype == VVAE_BF16) vgg_pre_fwd_kernel<float, bf16><<<blocks, 256, 0, s>>>((const float*)x, (bf16*)y, voxels, y_ld);
  else if (x_dtype == VVAE_BF16 && dtype == VVAE_BF16) vgg_pre_fwd_kernel<bf16, bf16><<<blocks, 256, 0, s>>>((const bf16*)x, (bf16*)y, voxels, y_ld);
  else if (x_dtype == VVAE_BF16 && dtype == VVAE_F32) vgg_pre_fwd_kernel<bf16, float><<<blocks, 256, 0, s>>>((const bf16*)x, (float*)y, voxels, y_ld);
  else VVAE_REQUIRE(false, "vgg_preprocess_fwd: bad dtypes");
  return check_launch("vgg_preprocess_fwd");
}

int vvae_vgg_preprocess_bwd(const void* dy, void* dx, long long voxels, int dy_ld, int dtype, vvae_stream_t stream) {
  if (voxels <= 0) return VVAE_OK;
  VVAE_REQUIRE(dy && dx && dy_ld >= 3, "vgg_preprocess_bwd: bad arguments");
  VVAE_DISPATCH_DTYPE(dtype, T, (vgg_pre_bwd_kernel<T><<<ew_blocks(voxels * 3), 256, 0, as_stream(stream)>>>(
                                    (const T*)dy, (T*)dx, voxels, dy_ld)));
  return check_launch("vgg_preprocess_bwd");
}

int vvae_philox_fill(float* out, long long n, unsigned long long seed, unsigned long long offset, int kind,
                     vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(out && (kind == 0 || kind == 1), "philox_fill: bad arguments");
  philox_fill_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(out, n, seed, offset, kind);
  return check_launch("philox_fill");
}

int vvae_sumsq_f32(const float* g, long long n, float* out1, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(g && out1, "sumsq: null pointer");
  sumsq_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(g, n, out1);
  return check_launch("sumsq");
}

int vvae_sumsq_partials(long long n) { return n > 0 ? ew_blocks(n) : 0; }

int vvae_sumsq_f32_det(const float* g, long long n, float* partials, float* out1, vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(g && partials && out1, "sumsq_det: null pointer");
  const int blocks = ew_blocks(n);
  sumsq_partials_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, n, partials);
  sumsq_final_kernel<<<1, 256, 0, as_stream(stream)>>>(partials, blocks, out1);
  return check_launch("sumsq_det");
}

int vvae_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float b1,
                   float b2, float eps, int step, const float* gnorm_sq, float clip, float grad_scale,
                   vvae_stream_t stream) {
  if (n <= 0) return VVAE_OK;
  VVAE_REQUIRE(p && g && m && v && step >= 1, "adam_step: bad arguments");
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  cudaStream_t s = as_stream(stream);
  const bool vec = n % 4 == 0 && (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0) &&
                   ((uintptr_t)shadow_bf16 % 8 == 0);
  if (vec) {
    const long long n4 = n / 4;
    const int blocks = (int)std::min<long long>(cdiv(n4, 512), (long long)num_sms() * 8);
    adam_vec4_kernel<<<blocks, 256, 0, s>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, (uint2*)shadow_bf16, n4, lr,
                                            b1, b2, eps, bc1, bc2, gnorm_sq, clip, grad_scale);
  } else {
    adam_kernel<<<ew_blocks(n), 256, 0, s>>>(p, g, m, v, (bf16*)shadow_bf16, n, lr, b1, b2, eps, bc1, bc2, gnorm_sq, clip,
                                             grad_scale);
  }
  return check_launch("adam_step");
}

}  // extern "C"
