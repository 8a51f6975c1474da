// conv3d weight gradient on tcgen05 (bf16 operands, fp32 accumulation in TMEM, fp32 atomics into dW).
//
//   dW[dt,dh,dw][ci][co] = sum over voxels v of  x[v + (dt,dh,dw) - pad][ci] * dy[v][co]
//
// is, per filter tap, a GEMM with M = Cout, N = Cin and K = voxels whose operands are both "MN-major" (channels are the
// contiguous index of the NDHWC tensors).  Channel counts of the U-Net are tiny (12..128) while the tensor core wants
// M = 128, so the M dimension is filled with SHIFTS: one image row of dy lies in shared memory as a dense
// [pixel][Cout] array; reading 128 consecutive elements from pixel k gives rows (j, co) = dy[k + j][co], j = 0..128/Cout-1,
// and one MMA against x[k + const][ci] therefore produces the gradients of 128/Cout horizontally adjacent taps at
// once (UMMA descriptor: MN-major, swizzle = channel row, leading-dimension stride = ONE pixel).  The (dt,dh) taps are
// separate MMAs that only move the start address of the x operand inside the TMA-staged, zero-padded input rows --
// except that the NDH VERTICAL taps of a tap group share one MMA whenever NDH*Cin <= 256 (PACKN): the x operand is
// MN-major too, so its N dimension can be built from NDH chunks of Cin channels one IMAGE ROW apart (leading-dimension
// stride = P pixels): N = NDH*Cin, D[(j,co)][(dh,ci)].  The A operand (4 KB per MMA, the shared-memory bound of this
// kernel: ncu l1tex tc wavefronts 82-87 % of peak, tensor pipe 18 %, profiles/r02a_conv_ncu.json) is then read once
// per temporal tap instead of once per (dt,dh): 2.4x (3x3x3) to 4.2x (3x7x7) fewer operand bytes.
// K runs over the pixels of a row (16 per MMA); accumulators for every tap of the CTA stay resident in TMEM for the
// whole kernel and are flushed once with atomics.  Nothing is gathered, transposed or im2col'ed.
//
// Work split: a tile is one image row segment (b, t, h, w-block of <= 128 pixels).  A CTA owns a tap group
// (NDT temporal x NDH vertical taps, all horizontal taps) so that its accumulators fit the 512 TMEM columns, and
// strides over the tiles.
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <set>

#include "common.cuh"
#include "sm100.cuh"

namespace vvae {

struct WgPlan {
  int B, T, H, W, kt, kh, kw, Cin, Cout;
  int cin_pad, cout_pad, ndt, ndh, ngroups, ctas_per_group;
  int Ct, P, ksteps, wblocks;
  long long tiles;
  uint32_t x_sub_bytes, x_sub_stride, a_data_bytes, a_slot_stride, stage_stride, tx_bytes;
  int stages, smem_bytes;
};

struct WgParams {
  WgPlan pl;
  float* dw;
};

static inline uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

constexpr uint32_t WG_PRE = 1024;  // zero bytes in front of every dy row (negative pixel shifts read zeros)

static bool wg_make_plan(const vvae_conv_args& a, WgPlan& p) {
  p.B = a.B; p.T = a.T; p.H = a.H; p.W = a.W; p.kt = a.kt; p.kh = a.kh; p.kw = a.kw; p.Cin = a.Cin; p.Cout = a.Cout;
  p.cin_pad = (a.Cin + 15) / 16 * 16;
  p.cout_pad = (a.Cout + 15) / 16 * 16;
  auto ok_c = [](int c) { return c == 16 || c == 32 || c == 64 || c == 128; };
  if (!ok_c(p.cin_pad) || !ok_c(p.cout_pad)) return false;
  if (p.cin_pad > a.x_ld || p.cout_pad > a.y_ld) return false;   // pad channels must exist in memory (and be zero)
  if ((a.x_ld * 2) % 16 || (a.y_ld * 2) % 16) return false;
  if (((uintptr_t)a.x % 16) || ((uintptr_t)a.y % 16)) return false;
  const int sh = p.cout_pad <= 64 ? 128 / p.cout_pad : 1;
  const int g = (a.kw + sh - 1) / sh;
  // largest tap group whose accumulators fit TMEM
  const int cand[3][2] = {{a.kt, a.kh}, {1, a.kh}, {1, 1}};
  p.ndt = 0;
  for (auto& c : cand) {
    if (c[0] * c[1] * g * p.cin_pad <= 512) { p.ndt = c[0]; p.ndh = c[1]; break; }
  }
  if (!p.ndt) return false;
  p.ngroups = (a.kt / p.ndt) * (a.kh / p.ndh);
  p.Ct = std::min(a.W, 128);
  p.P = p.Ct + a.kw - 1;
  if (p.P > 256) return false;
  p.ksteps = (p.P + 15) / 16;
  p.wblocks = (a.W + p.Ct - 1) / p.Ct;
  p.tiles = (long long)a.B * a.T * a.H * p.wblocks;
  const int rb_a = 2 * std::min(p.cout_pad, 64), nca = p.cout_pad / std::min(p.cout_pad, 64);
  const int rb_b = 2 * std::min(p.cin_pad, 64), ncb = p.cin_pad / std::min(p.cin_pad, 64);
  p.x_sub_bytes = (uint32_t)p.ndh * p.P * rb_b;
  p.x_sub_stride = up(p.x_sub_bytes + 16u * rb_b, 1024);          // + over-read of the K padding
  p.a_data_bytes = (uint32_t)p.Ct * rb_a;
  p.a_slot_stride = up(WG_PRE + (uint32_t)(p.ksteps * 16 + sh) * rb_a, 1024);
  p.stage_stride = (uint32_t)p.ndt * ncb * p.x_sub_stride + (uint32_t)nca * p.a_slot_stride;
  p.tx_bytes = (uint32_t)p.ndt * ncb * p.x_sub_bytes + (uint32_t)nca * p.a_data_bytes;
  const uint32_t slack = 1024 + 256;
  int s = (int)((225u * 1024u - slack) / p.stage_stride);
  if (s < 2) return false;
  p.stages = std::min(s, 4);
  p.smem_bytes = (int)std::max<uint32_t>(p.stages * p.stage_stride + slack, 120u * 1024u);
  p.ctas_per_group = (int)std::min<long long>(num_sms() / p.ngroups, p.tiles);
  return p.ctas_per_group >= 1;
}

__device__ __forceinline__ bool wg_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void wg_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <int KH, int KW, int NDT, int NDH, int CINP, int COUTP>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_sm100_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_dy,
                        const WgParams q) {
  constexpr int SH = COUTP <= 64 ? 128 / COUTP : 1;       // horizontal taps covered by one MMA
  constexpr int G = (KW + SH - 1) / SH;                   // MMAs per (dt,dh) tap row
  constexpr int CA = COUTP < 64 ? COUTP : 64, NCA = COUTP / CA, RBA = 2 * CA;
  constexpr int CBX = CINP < 64 ? CINP : 64, NCBX = CINP / CBX, RBB = 2 * CBX;
  constexpr int NACC = NDT * NDH * G, ACC_COLS = NACC * CINP;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  constexpr bool PACKN = NDH > 1 && CINP <= 64 && NDH * CINP <= 256;   // vertical taps packed into N
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;

  const WgPlan& p = q.pl;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.stage_stride);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 4;
  uint64_t* done_bar = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();      // every CTA of this grid is resident
  // zero all operand memory once: the pads around every dy row must read as zeros for the whole kernel, and the
  // K-padding over-reads of x must be finite
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (int)(((size_t)S * p.stage_stride) >> 4);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_x);
    sm100::tma_prefetch_desc(&tma_dy);
    for (int i = 0; i < S; ++i) {
      sm100::mbar_init(&full_bar[i], 1);
      sm100::mbar_init(&empty_bar[i], 1);
    }
    sm100::mbar_init(done_bar, 1);
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc<TMEM_COLS>(tmem_slot);
  sm100::fence_proxy_async();   // generic-proxy zero stores -> visible to the async proxy (UMMA operand reads)
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                   // PDL (common.cuh): the smem zero-fill / TMEM allocation above overlap the previous kernel

  const int group = blockIdx.x % p.ngroups, cta_in_group = blockIdx.x / p.ngroups;
  const int dh_groups = KH / NDH;
  const int dt0 = (group / dh_groups) * NDT, dh0 = (group % dh_groups) * NDH;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = cta_in_group; tile < p.tiles; tile += p.ctas_per_group) {
        long long rest = tile;
        const int wb = (int)(rest % p.wblocks); rest /= p.wblocks;
        const int h = (int)(rest % p.H); rest /= p.H;
        const int t = (int)(rest % p.T);
        const int b = (int)(rest / p.T);
        const int w0 = wb * p.Ct;
        sm100::mbar_wait(&empty_bar[stage], phase ^ 1);
        sm100::mbar_expect_tx(&full_bar[stage], p.tx_bytes);
        uint8_t* st = smem + (size_t)stage * p.stage_stride;
#pragma unroll
        for (int dtl = 0; dtl < NDT; ++dtl)
#pragma unroll
          for (int cb = 0; cb < NCBX; ++cb)
            sm100::tma_load_5d(st + (size_t)(dtl * NCBX + cb) * p.x_sub_stride, &tma_x, &full_bar[stage], cb * CBX,
                               w0 - KW / 2, h - KH / 2 + dh0, t + dt0 + dtl - p.kt / 2, b);
        uint8_t* sa = st + (size_t)NDT * NCBX * p.x_sub_stride;
#pragma unroll
        for (int ca = 0; ca < NCA; ++ca)
          sm100::tma_load_5d(sa + (size_t)ca * p.a_slot_stride + WG_PRE, &tma_dy, &full_bar[stage], ca * CA, w0, h, t, b);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = sm100::make_idesc_bf16(128, PACKN ? NDH * CINP : CINP, true, true);
    constexpr uint32_t LAY_A = RBA == 128 ? 2u : (RBA == 64 ? 4u : 6u);
    constexpr uint32_t LAY_B = RBB == 128 ? 2u : (RBB == 64 ? 4u : 6u);
    constexpr uint32_t hi_a = ((8u * RBA) >> 4) | (1u << 14) | (LAY_A << 29);   // SBO = 8 pixels, version 1, swizzle
    constexpr uint32_t hi_b = ((8u * RBB) >> 4) | (1u << 14) | (LAY_B << 29);
    // leading-dimension stride: next 16/32/64-channel chunk of M = next PIXEL (Cout <= 64) or the second channel block
    const uint32_t lbo_a = (COUTP <= 64) ? (uint32_t)RBA : p.a_slot_stride;
    // x: chunks of Cin channels one image row apart (PACKN), or the second 64-channel block of Cin = 128
    const uint32_t lbo_b = PACKN ? (uint32_t)p.P * RBB : p.x_sub_stride;
    const uint32_t lo_flags_a = ((lbo_a >> 4) & 0x3FFFu) << 16, lo_flags_b = ((lbo_b >> 4) & 0x3FFFu) << 16;
    const uint32_t smem16 = sm100::smem_u32(smem) >> 4;
    const uint32_t stage16 = p.stage_stride >> 4, xsub16 = p.x_sub_stride >> 4;
    const uint32_t xrow16 = ((uint32_t)p.P * RBB) >> 4;
    const int ksteps = p.ksteps;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum = 0;
    for (long long tile = cta_in_group; tile < p.tiles; tile += p.ctas_per_group) {
      sm100::mbar_wait(&full_bar[stage], phase);
      sm100::tc_fence_after();
      if (wg_elect_one()) {
        const uint32_t st16 = smem16 + stage * stage16;
        const uint32_t a16 = st16 + NDT * NCBX * xsub16 + (WG_PRE >> 4);
        for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            // rows (j, co) of this MMA are dy[k + j - (KW-1-g*SH)][co]  ->  tap dw = KW-1 - (g*SH + j)
            const uint32_t a_lo = ((a16 + (uint32_t)((ks * 16 - (KW - 1 - g * SH)) * (RBA >> 4))) & 0x3FFFu) | lo_flags_a;
            const uint64_t da = ((uint64_t)hi_a << 32) | a_lo;
#pragma unroll
            for (int dtl = 0; dtl < NDT; ++dtl) {
              if constexpr (PACKN) {
                const uint32_t b_lo = ((st16 + dtl * NCBX * xsub16 + (uint32_t)(ks * 16 * (RBB >> 4))) & 0x3FFFu) | lo_flags_b;
                const uint64_t db = ((uint64_t)hi_b << 32) | b_lo;
                sm100::umma_f16(tmem_base + (dtl * G + g) * (NDH * CINP), da, db, idesc, accum);
              } else {
#pragma unroll
                for (int dhl = 0; dhl < NDH; ++dhl) {
                  const uint32_t b_lo =
                      ((st16 + dtl * NCBX * xsub16 + dhl * xrow16 + (uint32_t)(ks * 16 * (RBB >> 4))) & 0x3FFFu) | lo_flags_b;
                  const uint64_t db = ((uint64_t)hi_b << 32) | b_lo;
                  sm100::umma_f16(tmem_base + ((dtl * NDH + dhl) * G + g) * CINP, da, db, idesc, accum);
                }
              }
            }
          }
          accum = 1;
        }
        sm100::umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
    if (wg_elect_one()) sm100::umma_commit(done_bar);
    __syncwarp();
  } else {
    // ===================== flush: TMEM -> fp32 atomics into dW[kt,kh,kw,Cin,Cout] =====================
    const bool any = cta_in_group < p.tiles;
    if (any) {
      sm100::mbar_wait(done_bar, 0);
      sm100::tc_fence_after();
      const int quarter = warp & 3;
      const int m = quarter * 32 + lane;
      const int j = COUTP <= 64 ? m / COUTP : 0;
      const int co = COUTP <= 64 ? m % COUTP : m;
#pragma unroll 1
      for (int acc = 0; acc < NACC; ++acc) {
        const int g = acc % G, dhl = (acc / G) % NDH, dtl = acc / (G * NDH);
        const int dwi = KW - 1 - (g * SH + j);
        const bool row_ok = dwi >= 0 && co < p.Cout;
        const long long tap = ((long long)(dt0 + dtl) * KH + (dh0 + dhl)) * KW + dwi;
        float* dst = q.dw + (tap * p.Cin) * p.Cout + co;
        const uint32_t acc_col = PACKN ? ((dtl * G + g) * NDH + dhl) * CINP : acc * CINP;
        const uint32_t taddr = tmem_base + acc_col + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int n0 = 0; n0 < CINP; n0 += 16) {
          uint32_t r[16];
          wg_tmem_ld16(taddr + n0, r);
          sm100::tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + i < p.Cin) atomicAdd(dst + (long long)(n0 + i) * p.Cout, __uint_as_float(r[i]));
          }
        }
      }
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

typedef void (*WgKernelFn)(const CUtensorMap, const CUtensorMap, const WgParams);

static WgKernelFn wg_pick(const WgPlan& p) {
#define WG_CASE(KH_, KW_, NDT_, NDH_, CI_, CO_)                                                                      \
  if (p.kh == KH_ && p.kw == KW_ && p.ndt == NDT_ && p.ndh == NDH_ && p.cin_pad == CI_ && p.cout_pad == CO_)         \
    return conv_wgrad_sm100_kernel<KH_, KW_, NDT_, NDH_, CI_, CO_>;
  WG_CASE(7, 7, 3, 7, 16, 16)
  WG_CASE(3, 3, 3, 3, 16, 16)
  WG_CASE(3, 3, 3, 3, 16, 32)
  WG_CASE(3, 3, 3, 3, 32, 16)
  WG_CASE(3, 3, 3, 3, 32, 32)
  WG_CASE(3, 3, 1, 3, 32, 64)
  WG_CASE(3, 3, 1, 3, 64, 32)
  WG_CASE(3, 3, 1, 3, 64, 64)
  WG_CASE(3, 3, 1, 1, 64, 128)
  WG_CASE(3, 3, 1, 1, 128, 64)
  WG_CASE(3, 3, 1, 1, 128, 128)
#undef WG_CASE
  return nullptr;
}

int conv_wgrad_tc_supported(const vvae_conv_args& a) {
  if (a.dtype != VVAE_BF16 || !a.x || !a.y || !a.dw_accum) return 0;
  WgPlan p;
  if (!wg_make_plan(a, p)) return 0;
  return wg_pick(p) != nullptr;
}

int conv_wgrad_tc_launch(const vvae_conv_args& a, cudaStream_t s) {
  WgParams q;
  if (!wg_make_plan(a, q.pl)) {
    set_error("conv3d wgrad: shape not supported by the tensor-core path");
    return VVAE_ERR_UNSUPPORTED;
  }
  const WgPlan& p = q.pl;
  WgKernelFn kern = wg_pick(p);
  if (!kern) {
    set_error("conv3d wgrad: no tensor-core kernel instance for this shape");
    return VVAE_ERR_UNSUPPORTED;
  }
  const int cbx = std::min(p.cin_pad, 64), ca = std::min(p.cout_pad, 64);
  CUtensorMap tx, ty;
  {
    uint64_t dims[5] = {(uint64_t)p.cin_pad, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.T, (uint64_t)a.B};
    uint64_t str[4] = {(uint64_t)a.x_ld * 2, (uint64_t)a.W * a.x_ld * 2, (uint64_t)a.H * a.W * a.x_ld * 2,
                       (uint64_t)a.T * a.H * a.W * a.x_ld * 2};
    uint32_t box[5] = {(uint32_t)cbx, (uint32_t)p.P, (uint32_t)p.ndh, 1, 1};
    int rc = encode_tmap_nd_bf16(&tx, a.x, 5, dims, str, box, cbx * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)p.cout_pad, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.T, (uint64_t)a.B};
    uint64_t str[4] = {(uint64_t)a.y_ld * 2, (uint64_t)a.W * a.y_ld * 2, (uint64_t)a.H * a.W * a.y_ld * 2,
                       (uint64_t)a.T * a.H * a.W * a.y_ld * 2};
    uint32_t box[5] = {(uint32_t)ca, (uint32_t)p.Ct, 1, 1, 1};
    int rc = encode_tmap_nd_bf16(&ty, a.y, 5, dims, str, box, ca * 2);
    if (rc) return rc;
  }
  q.dw = a.dw_accum;
  {
    static std::mutex mu;
    static std::set<WgKernelFn> configured;
    std::lock_guard<std::mutex> lk(mu);
    if (!configured.count(kern)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        set_error("conv3d wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return VVAE_ERR_CUDA;
      }
      configured.insert(kern);
    }
  }
  const int grid = p.ngroups * p.ctas_per_group;
  launch_pdl(kern, dim3(grid), dim3(192), (size_t)p.smem_bytes, s, tx, ty, q);
  return check_launch("conv_wgrad_sm100");
}

}  // namespace vvae
