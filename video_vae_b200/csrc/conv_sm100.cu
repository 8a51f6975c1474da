// conv3d (NDHWC, 'SAME', stride 1) as an implicit GEMM on tcgen05, bf16, forward and dgrad.
//
// Data flow per output tile (one frame t of one clip, R image rows x Ct columns, NT output channels):
//   * TMA (5-D tiled map over [B,T,H,W,C], out-of-bounds = zero fill = 'SAME' padding in t, h and w) loads, per
//     temporal tap dt and channel block cb, ONE zero-padded pixel tile of (R+kh-1) x (Ct+kw-1) pixels x CB channels into
//     shared memory, pixel-major, swizzled (32/64/128 B = CB*2 bytes per pixel row).
//   * The tile is addressed as a flat "padded plane" of pitch P = Ct+kw-1: output position m = r*P + c and filter tap
//     (dh,dw) read plane index m + dh*P + dw, so every tap is the SAME K-major UMMA operand at a different start address
//     (swizzled operand views may start at any row: measured).  No im2col is ever materialised and the
//     input tile is read from L2/HBM once per (dt, cb), not once per tap.
//   * Weights are pre-arranged once per step (conv_wprep_kernel) into the swizzled K-major image of every stage, and a
//     single cp.async.bulk brings a stage's taps in.
//   * tcgen05.mma (M=128 positions, K=16 channels) accumulates all taps x channel slices in TMEM; accumulators
//     are double buffered so the epilogue (bias / residual, bf16 store, pad positions dropped) overlaps the next tile.
//   * HORIZONTAL TAPS ARE PACKED INTO N (PACK = true, the default).  The U-Net's channel counts are tiny (12..128): with
//     N = 16 output channels a 128x16x16 MMA needs 8 tensor-pipe cycles but 32 cycles of shared-memory reads for its
//     4 KB A operand (ncu, profiles/r02a_conv_ncu.json: l1tex tc wavefronts 84 % of peak, tensor pipe 19 %).  So one MMA
//     per (dt, dh, k-slice) multiplies the A rows at plane index m + dh*P by ALL kw taps' weights at once,
//     N = kw*NT: D'[m][dw][co] = sum_ci X[m + dh*P][ci] W[dh][dw][ci][co], and the epilogue forms
//     out[m][co] = sum_dw D'[m + dw][dw][co] -- a shift by dw TMEM LANES, done with warp shuffles plus a small
//     shared-memory exchange of the first kw-1 rows of the next warp's quarter.  128-row blocks overlap by kw-1 rows
//     (block stride BS = 128-(kw-1)) so that no sum crosses a block.  A is read once per (dt,dh) instead of once per tap:
//     2.4x (3x3x3, 16 ch) to 4.2x (3x7x7) fewer shared-memory operand bytes per output.
// dgrad is the same kernel over dy with flipped, transposed weights.
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <set>

#include "common.cuh"
#include "sm100.cuh"

namespace vvae {

struct ConvPlan {
  int B, T, H, W, kt, kh, kw;
  int Cin_pad, Cout, CB, nCB, NT, nNT;
  int R, Ct, P, rows, nblk, Mtot;
  int rowbytes, layout_type, swizzle_bytes;
  uint32_t a_bytes, a_stride, w_bytes, w_stride;
  int stages, smem_bytes;
  int hblocks, wblocks, total_tiles;
  int pack, BS;     // pack: horizontal taps in N; BS: plane positions per 128-row MMA block that produce outputs
  int w_res;        // the whole weight image stays in shared memory (loaded once per CTA) instead of riding along with
                    // every stage: with ~37k tiles x 3 stages, all 148 SMs re-fetching the same 14 KB hammered a few L2
                    // slices -- the barrier/weight skeleton alone took 0.21 of 0.34 ms (profiles/r02g_conv_ablate.jsonl)
  uint32_t w_total; // bytes of the weight image (nNT * kt * nCB stage slots of w_stride)
  int sub;          // input slabs (dt, cb) per pipeline stage: 1, or all kt*nCB of a tile behind ONE barrier handshake
                    // (the single-thread producer / issuer loops cost ~450 cycles per handshake: ablation r02h)
  uint32_t stage_stride;   // bytes between stages of the input ring = sub * a_stride
  // slide: a CTA walks one spatial tile through Tc consecutive frames and keeps the zero-padded input slabs of frames
  // t-1, t, t+1 in the ring, so every frame slab is fetched ONCE per tile column instead of once per temporal tap
  // (3x less L2 -> smem traffic: the TMA-only ablation ran at 6.6 TB/s of L2 reads, profiles/r02i_conv_ablate.jsonl)
  int slide, Tc, tchunks, units;
};

struct ConvParams {
  ConvPlan pl;
  const bf16* wimg;
  bf16* y; long long y_ld;
  const float* bias;
  int mode; const bf16* aux; long long ld_aux;
  int store_c;  // channels written per voxel: Cout, or ceil16(Cout) when the caller asked for zeroed pad channels
  int dbg;      // vvae_debug_set(14): timing ablations (results are WRONG): 1 no global stores, 2 epilogue releases the
                // accumulator at once, 4 no input-tile TMA, 8 no MMAs
};

static inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

extern long long g_dbg[32];   // vvae_debug_set: key 13 != 0 selects the one-MMA-per-tap kernels (round-1 behaviour)
constexpr int CONV_THREADS = 320;        // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5 / 6-9: two epilogue groups
constexpr uint32_t CONV_MISC_BYTES = 1024;   // barriers (256 B) + the layer's bias as fp32 (<= 128 values) behind the stages
// epilogue row exchange (PACK): [group 2][buffer 2][quarter 4][block][kw-1 rows][kw-1 taps][16 channels] fp32
// (blocks per epilogue pass: both blocks of a tile for 3-wide filters, one for wider ones -- EB in the kernel; NBLK <= 2)
static inline uint32_t conv_xch_bytes(int kw, int pack) {
  return pack ? 2u * 2u * 4u * (kw <= 3 ? 2u : 1u) * (kw - 1) * (kw - 1) * 16u * 4u : 0u;
}

// which: 0 = forward (gathers x, Cin -> Cout), 1 = dgrad (gathers dy, Cout -> Cin)
static bool make_plan(const vvae_conv_args& a, int which, ConvPlan& p) {
  const int cin = which == 0 ? a.Cin : a.Cout;    // channels of the gathered tensor
  const int cout = which == 0 ? a.Cout : a.Cin;   // channels produced
  const long long in_ld = which == 0 ? a.x_ld : a.y_ld;
  p.B = a.B; p.T = a.T; p.H = a.H; p.W = a.W; p.kt = a.kt; p.kh = a.kh; p.kw = a.kw;
  p.Cin_pad = (cin + 15) / 16 * 16;
  p.Cout = cout;
  if (p.Cin_pad > in_ld) return false;            // padded channels must exist in memory (and be zero / finite)
  if ((in_ld * 2) % 16) return false;
  p.CB = p.Cin_pad >= 64 ? 32 : (p.Cin_pad >= 32 ? 32 : 16);
  if (p.Cin_pad % p.CB) return false;
  p.nCB = p.Cin_pad / p.CB;
  const int cout_pad = (cout + 15) / 16 * 16;
  p.NT = std::min(cout_pad, 64);
  if (cout_pad % p.NT) return false;
  p.nNT = cout_pad / p.NT;
  p.pack = (g_dbg[13] == 0 && a.kw > 1 && a.kw * p.NT <= 256) ? 1 : 0;
  p.BS = p.pack ? 128 - (a.kw - 1) : 128;
  const int acc_cols_per_blk = p.pack ? a.kw * p.NT : p.NT;
  // Tile shape.  A tile of R rows x Ct columns is fetched with its halo ((R+kh-1) x (Ct+kw-1) pixels per temporal tap)
  // and computed as nblk blocks of 128 positions of the flat padded plane, Mtot = R*Ct + (R-1)*(kw-1) <= 128*nblk.
  // Short, wide tiles (R = 1) re-read every input row kh times; choose the (nblk, R, Ct) that minimises
  // halo traffic + padded MMA work per useful output, including the ragged right / bottom edges of the image.
  {
    const int max_nblk = (p.pack && g_dbg[15] != 4) ? 4 : 2;   // 4 blocks: four independent accumulator chains (see kernel)
    double best = 1e30;
    int bR = 0, bC = 0;
    for (int nb = 1; nb <= max_nblk; ++nb) {
      if (nb == 3) continue;                         // kernel instances exist for 1, 2 and 4 blocks
      if (nb * acc_cols_per_blk > 256) break;
      if (nb == 4 && p.NT != 16) break;
      for (int R = 1; R <= std::min(a.H, 32); ++R) {
        int cmax = (p.BS * nb - (R - 1) * (a.kw - 1)) / R;
        cmax = std::min(cmax, std::min(a.W, 256 - (a.kw - 1)));
        if (cmax < 1) break;
        for (int C = cmax; C >= std::max(1, cmax - 24); --C) {
          if (R + a.kh - 1 > 256) continue;
          const double useful = (double)a.H * a.W;
          const double tiles = (double)((a.H + R - 1) / R) * ((a.W + C - 1) / C);
          const double traffic = tiles * (R + a.kh - 1) * (C + a.kw - 1) / useful;
          const double mma = tiles * 128.0 * nb / useful;
          const double cost = traffic + mma + 0.02 * nb;       // slight preference for the smaller accumulator
          if (cost < best) { best = cost; bR = R; bC = C; }
        }
      }
    }
    if (!bR) return false;
    p.R = bR;
    p.Ct = bC;
  }
  p.P = p.Ct + a.kw - 1;
  if (p.P > 256) return false;
  p.rows = p.R + a.kh - 1;
  if (p.rows > 256) return false;
  p.Mtot = (p.R - 1) * p.P + p.Ct;
  p.nblk = (p.Mtot + p.BS - 1) / p.BS;
  if (p.nblk * acc_cols_per_blk > 256) return false;          // TMEM: 2 x nblk x (kw x) NT columns <= 512
  p.rowbytes = p.CB * 2;
  p.swizzle_bytes = p.rowbytes;
  p.layout_type = p.rowbytes == 128 ? 2 : (p.rowbytes == 64 ? 4 : 6);
  p.a_bytes = (uint32_t)p.rows * p.P * p.rowbytes;
  p.a_stride = align_up(p.a_bytes, 1024);
  p.w_bytes = (uint32_t)a.kh * a.kw * p.NT * p.rowbytes;
  p.w_stride = align_up(p.w_bytes, 1024);
  // alignment + over-read of pad rows + barriers / bias + exchange
  const uint32_t slack = 1024 + 128u * 128u + CONV_MISC_BYTES + conv_xch_bytes(a.kw, p.pack);
  p.w_total = (uint32_t)(p.nNT * a.kt * p.nCB) * p.w_stride;
  p.w_res = (g_dbg[15] == 0 && p.nNT == 1 && p.w_total <= 112u * 1024u && p.w_total + 3 * p.a_stride + slack <= 225u * 1024u) ? 1 : 0;
  const uint32_t fixed = slack + (p.w_res ? p.w_total : 0u);
  p.sub = 1;
  if (p.w_res && g_dbg[15] != 2 && (225u * 1024u - fixed) / ((uint32_t)(a.kt * p.nCB) * p.a_stride) >= 3) p.sub = a.kt * p.nCB;
  p.slide = 0; p.Tc = a.T; p.tchunks = 1; p.units = 0;
  if (p.w_res && a.kt == 3 && g_dbg[15] != 2 && g_dbg[15] != 3 && (225u * 1024u - fixed) / ((uint32_t)p.nCB * p.a_stride) >= 4) {
    p.slide = 1;
    p.sub = p.nCB;                    // a ring slot = one frame slab = all channel blocks of that frame
    const long long spatial = (long long)a.B * ((a.H + p.R - 1) / p.R) * ((a.W + p.Ct - 1) / p.Ct);
    // time chunks: long enough to amortise the two extra slabs per chunk, short enough for >= ~3 units per SM
    p.Tc = a.T;
    while (p.Tc > 4 && spatial * ((a.T + p.Tc - 1) / p.Tc) < 3LL * num_sms()) p.Tc = (p.Tc + 1) / 2;
    p.tchunks = (a.T + p.Tc - 1) / p.Tc;
    if (spatial * p.tchunks > 0x7fffffffLL) return false;
    p.units = (int)(spatial * p.tchunks);
  }
  p.stage_stride = (uint32_t)p.sub * p.a_stride;
  const uint32_t per_stage = p.stage_stride + (p.w_res ? 0u : p.w_stride);
  int s = (int)((225u * 1024u - fixed) / per_stage);
  if (s < 2) return false;
  p.stages = std::min(s, p.slide ? 8 : (p.sub > 1 ? 6 : (p.w_res ? 8 : 6)));
  p.smem_bytes = (int)std::max<uint32_t>(p.stages * per_stage + fixed, 120u * 1024u);  // >= 120 KB: one CTA per SM (TMEM)
  if (p.nblk > 4 || p.nblk == 3) p.nblk = p.nblk == 3 ? 4 : 99;   // instances exist for 1, 2 and 4 blocks
  if (p.nblk > 4 || p.nblk * acc_cols_per_blk > 256) return false;
  p.hblocks = (a.H + p.R - 1) / p.R;
  p.wblocks = (a.W + p.Ct - 1) / p.Ct;
  const long long tiles = (long long)a.B * a.T * p.hblocks * p.wblocks * p.nNT;
  if (tiles > 0x7fffffffLL) return false;
  p.total_tiles = (int)tiles;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight image: [nNT][kt][nCB][kh*kw][NT rows][CB cols] bf16, each [NT][CB] tile K-major with the stage swizzle applied
// (Swizzle<B,4,3>: byte-offset bits [4,4+B) ^= bits [7,7+B), B = log2(rowbytes/16)); each stage starts 1024-aligned.
__global__ void conv_wprep_kernel(const bf16* __restrict__ w, bf16* __restrict__ img, ConvPlan p, int which, int Cin,
                                  int Cout) {
  const int taps_hw = p.kh * p.kw;
  const long long per_stage_el = (long long)p.w_stride / 2;
  const long long total = (long long)p.nNT * p.kt * p.nCB * per_stage_el;
  const uint32_t mask = (uint32_t)(p.rowbytes / 16 - 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long stage = i / per_stage_el;
    const uint32_t byte = (uint32_t)(i % per_stage_el) * 2;
    float val = 0.f;
    if (byte < p.w_bytes) {
      // undo the swizzle to find which logical (tap, n, k) lives at this physical position
      const uint32_t tile_bytes = (uint32_t)p.NT * p.rowbytes;
      const uint32_t tap = byte / tile_bytes;
      uint32_t off = byte % tile_bytes;
      off ^= ((off >> 7) & mask) << 4;
      const int n = off / p.rowbytes, k = (off % p.rowbytes) / 2;
      const int cb = (int)(stage % p.nCB);
      const int dt = (int)((stage / p.nCB) % p.kt);
      const int nt = (int)(stage / ((long long)p.nCB * p.kt));
      const int dh = tap / p.kw, dw = tap % p.kw;
      const int kin = cb * p.CB + k;      // channel of the gathered tensor
      const int nout = nt * p.NT + n;     // produced channel
      if (which == 0) {
        if (kin < Cin && nout < Cout) val = __bfloat162float(w[((((long long)dt * p.kh + dh) * p.kw + dw) * Cin + kin) * Cout + nout]);
      } else {  // dgrad: gathered = dy (Cout channels), produced = dx (Cin channels), taps mirrored
        if (kin < Cout && nout < Cin)
          val = __bfloat162float(w[((((long long)(p.kt - 1 - dt) * p.kh + (p.kh - 1 - dh)) * p.kw + (p.kw - 1 - dw)) * Cin + nout) * Cout + kin]);
      }
      (void)taps_hw;
    }
    img[i] = __float2bfloat16_rn(val);
  }
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;                              // LBO: unused for swizzled K-major
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sm100::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(sm100::smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// store 16 consecutive output channels [n0, n0+16) of one pixel (only those < Cout are written); sb = this chunk's 16 bias
// values in shared memory (zeros when the layer has no bias)
__device__ __forceinline__ void conv_store16(const ConvParams& q, long long pix, int n0, const uint32_t (&r)[16],
                                             const float* sb) {
  const int cout = q.pl.Cout;
  if (n0 >= q.store_c) return;
  float v[16];
  const float4* b4 = reinterpret_cast<const float4*>(sb);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 bb = b4[j];
    v[4 * j] = __uint_as_float(r[4 * j]) + bb.x;
    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z;
    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
  }
  if (q.mode == VVAE_EPI_RESIDUAL) {
    const bf16* ax = q.aux + pix * q.ld_aux + n0;
    if (n0 + 16 <= cout && ((reinterpret_cast<uintptr_t>(ax) & 15) == 0)) {
      Vec16<bf16> a0, a1;
      a0.load(ax);
      a1.load(ax + 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[j] += a0.get(j); v[8 + j] += a1.get(j); }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + j < cout) v[j] += __bfloat162float(ax[j]);
    }
  }
  bf16* dst = q.y + pix * q.y_ld + n0;
  const int nvalid = min(16, q.store_c - n0);  // accumulators of channels >= Cout are exact zeros (zero weights)
  if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      Vec16<bf16> o;
#pragma unroll
      for (int t = 0; t < 8; ++t) o.set(t, v[j + t]);
      o.store(dst + j);
    }
  } else if ((nvalid % 4) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      if (j < nvalid) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[j], v[j + 1]), b = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + j) = pk;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) dst[j] = __float2bfloat16_rn(v[j]);
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

// Walks the tile sequence blockIdx.x, blockIdx.x + gridDim.x, ... in mixed radix (nt, wb, hb, t, b) without a division
// per tile (five runtime div/mod pairs per tile were ~500 cycles of the single-thread producer loop).
struct ConvTileIter {
  int nt, wb, hb, t, b;
  int s_nt, s_wb, s_hb, s_t, s_b;
  __device__ __forceinline__ void init(const ConvPlan& p, int tile0, int step) {
    int r = tile0;
    nt = r % p.nNT; r /= p.nNT;  wb = r % p.wblocks; r /= p.wblocks;  hb = r % p.hblocks; r /= p.hblocks;  t = r % p.T;  b = r / p.T;
    r = step;
    s_nt = r % p.nNT; r /= p.nNT;  s_wb = r % p.wblocks; r /= p.wblocks;  s_hb = r % p.hblocks; r /= p.hblocks;  s_t = r % p.T;  s_b = r / p.T;
  }
  __device__ __forceinline__ void advance(const ConvPlan& p) {
    nt += s_nt;  int c = nt >= p.nNT;      nt -= c ? p.nNT : 0;
    wb += s_wb + c;  c = wb >= p.wblocks;  wb -= c ? p.wblocks : 0;
    hb += s_hb + c;  c = hb >= p.hblocks;  hb -= c ? p.hblocks : 0;
    t += s_t + c;    c = t >= p.T;         t -= c ? p.T : 0;
    b += s_b + c;
  }
};

// slide mode: unit u -> (wb, hb, time chunk, b); one division chain per unit (= per Tc tiles)
struct ConvUnit {
  int wb, hb, b, t0, nf;
  __device__ __forceinline__ void decode(const ConvPlan& p, int u) {
    wb = u % p.wblocks; u /= p.wblocks;
    hb = u % p.hblocks; u /= p.hblocks;
    const int tc = u % p.tchunks;
    b = u / p.tchunks;
    t0 = tc * p.Tc;
    nf = min(p.Tc, p.T - t0);
  }
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// KH x KW filter taps per frame, NKS = CB/16 K-slices per tap, NBLK 128-position blocks per tile, NT output channels per
// MMA.  Everything the single MMA-issuing thread touches is a compile-time constant (or a loop-invariant register), so
// the issue loop is one descriptor add + one tcgen05.mma per MMA: with N = 16..64 the tensor pipe needs a new
// instruction every 8..32 cycles and the generic (runtime-loop) version spent ~150 cycles of scalar work per MMA.
//
// PACK: the kw horizontal taps of a filter row share one MMA (N = KW*NT) and are combined in the epilogue (see the file
// header); PACK = false issues one N = NT MMA per tap at a shifted A address (kept for 1x1 filters and as a reference).
template <int KH, int KW, int NKS, int NBLK, int NT, bool PACK>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_sm100_kernel(const __grid_constant__ CUtensorMap tma_x, const ConvParams q) {
  const ConvPlan& p = q.pl;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + (size_t)S * p.stage_stride;   // per-stage weight slots, or the resident weight image
  // barriers live after the weights plus the over-read slack
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.stage_stride +
                                               (p.w_res ? (size_t)p.w_total : (size_t)S * p.w_stride) + 128 * 128);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* tmem_full = bars + 16;
  uint64_t* tmem_empty = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t* w_bar = bars + 22;                                         // resident weights have landed
  float* s_bias = reinterpret_cast<float*>(bars + 32);                 // [nNT * NT] fp32 (zeros without a bias)
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + CONV_MISC_BYTES);   // PACK epilogue only
  constexpr int ROWBYTES = NKS * 32;
  constexpr int NW = PACK ? KW * NT : NT;                              // accumulator columns per 128-row block
  constexpr int BS = PACK ? 128 - (KW - 1) : 128;                      // block stride in plane positions
  constexpr int ACC_COLS = NBLK * NW;
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(NW % 16 == 0 && NW <= 256, "UMMA N");
  constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128
                            : (2 * ACC_COLS <= 256) ? 256 : 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    sm100::tma_prefetch_desc(&tma_x);
    for (int i = 0; i < S; ++i) {
      sm100::mbar_init(&full_bar[i], 1);
      sm100::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&tmem_full[i], 1);
      sm100::mbar_init(&tmem_empty[i], NBLK >= 2 ? 8 : 4);   // warps that drain one accumulator (SPLIT: both groups)
    }
    sm100::mbar_init(w_bar, 1);
    sm100::fence_barrier_init();
  }
  pdl_launch_dependents();   // PDL (common.cuh): every CTA of this grid is resident
  if (warp == 1) sm100::tmem_alloc<TMEM_COLS>(tmem_slot);
  pdl_wait();                // the bias below may be the previous kernel's (optimizer) output
  for (int i = threadIdx.x; i < p.nNT * NT; i += blockDim.x) s_bias[i] = (q.bias && i < p.Cout) ? __ldg(q.bias + i) : 0.f;
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int slabs_per_tile = p.kt * p.nCB;
  const int stages_per_tile = slabs_per_tile / p.sub;

  if (warp == 0) {
    if (lane == 0) {
      if (p.w_res) {     // the whole weight image, once (bulk copies of <= 32 KB each)
        sm100::mbar_expect_tx(w_bar, p.w_total);
        for (uint32_t off = 0; off < p.w_total; off += 32768u)
          bulk_g2s(smem_w + off, reinterpret_cast<const uint8_t*>(q.wimg) + off, min(32768u, p.w_total - off), w_bar);
      }
      const uint32_t tx_bytes = ((q.dbg & 4) ? 0u : (uint32_t)p.sub * p.a_bytes) + (p.w_res ? 0u : p.w_bytes);
      if (p.slide) {
        uint32_t sq = 0;                                     // slab sequence number of this CTA: slot sq % S, phase (sq / S) & 1
        int slot = 0;
        uint32_t ph = 0;
        for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
          ConvUnit cu;
          cu.decode(p, u);
          const int w0 = cu.wb * p.Ct - KW / 2, h0 = cu.hb * p.R - KH / 2;
          for (int f = -1; f <= cu.nf; ++f, ++sq) {          // frames t0-1 .. t0+nf (out-of-range frames: TMA zero fill)
            sm100::mbar_wait(&empty_bar[slot], ph ^ 1);
            sm100::mbar_expect_tx(&full_bar[slot], tx_bytes);
            if (!(q.dbg & 4))
              for (int cb = 0; cb < p.nCB; ++cb)
                sm100::tma_load_5d(smem_a + (size_t)slot * p.stage_stride + (size_t)cb * p.a_stride, &tma_x, &full_bar[slot],
                                   cb * p.CB, w0, h0, cu.t0 + f, cu.b);
            if (++slot == S) { slot = 0; ph ^= 1; }
          }
        }
      } else {
      int stage = 0;
      uint32_t phase = 0;
      ConvTileIter it;
      it.init(p, blockIdx.x, gridDim.x);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it.advance(p)) {
        const int w0 = it.wb * p.Ct - KW / 2, h0 = it.hb * p.R - KH / 2;
        int dt = 0, cb = 0;
        for (int s = 0; s < stages_per_tile; ++s) {
          sm100::mbar_wait(&empty_bar[stage], phase ^ 1);
          sm100::mbar_expect_tx(&full_bar[stage], tx_bytes);
          for (int u = 0; u < p.sub; ++u) {
            if (!(q.dbg & 4))
              sm100::tma_load_5d(smem_a + (size_t)stage * p.stage_stride + (size_t)u * p.a_stride, &tma_x, &full_bar[stage],
                                 cb * p.CB, w0, h0, it.t + dt - p.kt / 2, it.b);
            if (!p.w_res)   // (sub == 1)
              bulk_g2s(smem_w + (size_t)stage * p.w_stride,
                       reinterpret_cast<const uint8_t*>(q.wimg) + ((size_t)(it.nt * p.kt + dt) * p.nCB + cb) * p.w_stride,
                       p.w_bytes, &full_bar[stage]);
            if (++cb == p.nCB) { cb = 0; ++dt; }
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop (warp-uniform control flow keeps the address arithmetic in uniform registers);
    // one elected lane issues the tensor-core instructions.
    constexpr uint32_t idesc = sm100::make_idesc_bf16(128, NW, false, false);
    constexpr uint32_t LAYOUT = ROWBYTES == 128 ? 2u : (ROWBYTES == 64 ? 4u : 6u);
    constexpr uint32_t RB16 = ROWBYTES >> 4;                             // row pitch in 16-byte units
    constexpr uint32_t desc_hi = ((8u * ROWBYTES) >> 4) | (1u << 14) | (LAYOUT << 29);
    constexpr uint32_t lo_flags = 1u << 16;                              // LBO field (unused for swizzled K-major)
    const uint32_t a_row_step = (uint32_t)p.P * RB16;
    const uint32_t a_stride16 = p.a_stride >> 4, w_stride16 = p.w_stride >> 4, stage_stride16 = p.stage_stride >> 4;
    const uint32_t a_base16 = sm100::smem_u32(smem_a) >> 4, w_base16 = sm100::smem_u32(smem_w) >> 4;
    // all MMAs of one input slab (one frame, one channel block): every filter row / k-slice / block, taps packed in N
    auto issue_slab = [&](uint32_t d_base, uint32_t a_lo0, uint32_t w_lo0, uint32_t accumulate_first) {
#pragma unroll
      for (int dh = 0; dh < KH; ++dh) {
        const uint32_t a_row = a_lo0 + dh * a_row_step;
        if constexpr (PACK) {
          // one MMA per (dh, k-slice, block): the B operand spans the KW taps of filter row dh (KW*NT rows of the
          // weight image, contiguous), the A operand is NOT shifted by dw -- the epilogue applies the shift
#pragma unroll
          for (int ks = 0; ks < NKS; ++ks) {
            const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(w_lo0 + (dh * KW) * NT * RB16 + 2u * ks);
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
              const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_row + 2u * ks + blk * (uint32_t)BS * RB16);
              if (dh == 0 && ks == 0)
                sm100::umma_f16(d_base + blk * NW, da, db, idesc, accumulate_first);
              else
                sm100::umma_f16_acc(d_base + blk * NW, da, db, idesc);
            }
          }
        } else {
#pragma unroll
          for (int dw = 0; dw < KW; ++dw) {
#pragma unroll
            for (int ks = 0; ks < NKS; ++ks) {
              const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(w_lo0 + (dh * KW + dw) * NT * RB16 + 2u * ks);
#pragma unroll
              for (int blk = 0; blk < NBLK; ++blk) {
                const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_row + dw * RB16 + 2u * ks + blk * 128u * RB16);
                if (dh == 0 && dw == 0 && ks == 0)
                  sm100::umma_f16(d_base + blk * NT, da, db, idesc, accumulate_first);
                else
                  sm100::umma_f16_acc(d_base + blk * NT, da, db, idesc);
              }
            }
          }
        }
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    if (p.w_res) sm100::mbar_wait(w_bar, 0);
    if (p.slide) {
      // output frame t0+ot reads the slabs of frames t0+ot-1, t0+ot, t0+ot+1 = sequence numbers sq+ot .. sq+ot+2
      uint32_t sq = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        ConvUnit cu;
        cu.decode(p, u);
        for (int ot = 0; ot < cu.nf; ++ot) {
          sm100::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          const uint32_t d_base = tmem_base + acc * ACC_COLS;
          uint32_t slots[3];
#pragma unroll
          for (int dt = 0; dt < 3; ++dt) {
            const uint32_t sn = sq + ot + dt;
            slots[dt] = sn % (uint32_t)S;
            if (ot == 0 || dt == 2) sm100::mbar_wait(&full_bar[slots[dt]], (sn / (uint32_t)S) & 1u);   // older slabs: waited before
          }
          sm100::tc_fence_after();
          if (elect_one()) {
            if (!(q.dbg & 8)) {
#pragma unroll
              for (int dt = 0; dt < 3; ++dt)
                for (int cb = 0; cb < p.nCB; ++cb) {
                  const uint32_t a_lo0 = ((a_base16 + slots[dt] * stage_stride16 + cb * a_stride16) & 0x3FFFu) | lo_flags;
                  const uint32_t w_lo0 = ((w_base16 + (dt * p.nCB + cb) * w_stride16) & 0x3FFFu) | lo_flags;
                  issue_slab(d_base, a_lo0, w_lo0, (dt > 0 || cb > 0) ? 1u : 0u);
                }
            }
            sm100::umma_commit(&empty_bar[slots[0]]);        // frame t0+ot-1 is not needed again
            if (ot == cu.nf - 1) {                            // end of the chunk: the last two slabs as well
              sm100::umma_commit(&empty_bar[slots[1]]);
              sm100::umma_commit(&empty_bar[slots[2]]);
            }
            sm100::umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        sq += (uint32_t)cu.nf + 2u;
      }
    } else {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      sm100::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      sm100::tc_fence_after();
      const uint32_t d_base = tmem_base + acc * ACC_COLS;
      for (int s = 0; s < stages_per_tile; ++s) {
        sm100::mbar_wait(&full_bar[stage], phase);
        sm100::tc_fence_after();
        if (elect_one()) {
          if (!(q.dbg & 8)) {
            for (int u = 0; u < p.sub; ++u) {
              const int slab = s * p.sub + u;                 // (dt, cb) index inside the tile
              const uint32_t a_lo0 = ((a_base16 + stage * stage_stride16 + u * a_stride16) & 0x3FFFu) | lo_flags;
              const uint32_t w_lo0 = ((w_base16 + (p.w_res ? slab : stage) * w_stride16) & 0x3FFFu) | lo_flags;
              issue_slab(d_base, a_lo0, w_lo0, slab > 0 ? 1u : 0u);
            }
          }
          sm100::umma_commit(&empty_bar[stage]);
          if (s == stages_per_tile - 1) sm100::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
  } else {
    // ===================== epilogue: two groups of four warps =====================
    // NBLK == 1: group g drains accumulator g (every other tile).  NBLK >= 2 (SPLIT): both groups work on EVERY tile,
    // group g drains blocks [g*NBLK/2, (g+1)*NBLK/2) -- the epilogue is a latency chain, two groups halve it per tile.
    constexpr bool SPLIT = NBLK >= 2;
    constexpr int GB = SPLIT ? NBLK / 2 : NBLK;              // blocks per group and tile
    const int eg = (warp - 2) >> 2;
    const int gb0 = SPLIT ? eg * GB : 0;
    const int quarter = warp & 3;
    const int mp = quarter * 32 + lane;                      // row of the 128-row MMA block = TMEM lane
    // row -> (image row, column) of the tile, per block; the same for every tile
    int row_r[NBLK], row_c[NBLK];                            // (SPLIT: entries [0, GB) describe this group's blocks)
    bool row_ok[NBLK];
#pragma unroll
    for (int blk = 0; blk < NBLK; ++blk) {
      const int m = ((SPLIT ? ((warp - 2) >> 2) * (NBLK / 2) : 0) + blk) * BS + mp;   // position in the padded plane
      row_r[blk] = m / p.P;
      row_c[blk] = m - row_r[blk] * p.P;
      row_ok[blk] = (mp < BS) && (m < p.Mtot) && (row_c[blk] < p.Ct);
    }
    constexpr int XROW = (KW - 1) * 16;                      // floats per exchanged row: taps 1..KW-1 x 16 channels
    constexpr int XQ = (KW - 1) * XROW;                      // floats per quarter
    // this group's two exchange buffers (shared-space byte address; sized for NBLK blocks per pass)
    const uint32_t xg_s = sm100::smem_u32(xch) + (uint32_t)(eg * (2 * 4 * ((KW <= 3) ? 2 : 1) * XQ)) * 4u;
    uint32_t acc_phase = 0;
    int acc = SPLIT ? 0 : eg;
    uint32_t xch_it = 0;
    (void)xch_it; (void)xg_s;
    auto drain = [&](int nt, int wb, int hb, int t, int b) {
      const long long pix0 = (((long long)b * p.T + t) * p.H + hb * p.R) * p.W + wb * p.Ct;
      const int a_cur = acc;
      sm100::mbar_wait(&tmem_full[a_cur], acc_phase);
      if (SPLIT) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } } else { acc_phase ^= 1; }
      sm100::tc_fence_after();
      if (q.dbg & 2) {
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&tmem_empty[a_cur]);
        return;
      }
      bool valid[NBLK];
      long long pix[NBLK];
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        valid[blk] = row_ok[blk] && (hb * p.R + row_r[blk] < p.H) && (wb * p.Ct + row_c[blk] < p.W) && !(q.dbg & 1);
        pix[blk] = pix0 + (long long)row_r[blk] * p.W + row_c[blk];
      }
      const uint32_t taddr0 = tmem_base + a_cur * ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      if constexpr (!PACK) {
#pragma unroll
        for (int j = 0; j < GB; ++j) {
          uint32_t rr[NT / 16][16];
#pragma unroll
          for (int cc = 0; cc < NT / 16; ++cc) tmem_ld_32x16(taddr0 + (gb0 + j) * NW + cc * 16, rr[cc]);
          sm100::tmem_ld_wait();
          if (valid[j]) {
#pragma unroll
            for (int cc = 0; cc < NT / 16; ++cc)
              conv_store16(q, pix[j], nt * NT + cc * 16, rr[cc], s_bias + nt * NT + cc * 16);
          }
        }
      } else {
        // out[m][co] = sum_dw D'[m + dw][dw*NT + co]: rows m + dw are the next dw TMEM lanes -> warp shuffles; the last
        // dw lanes of a quarter take the first rows of the next quarter from shared memory (xch).  Rows mp >= BS of a
        // block are incomplete by construction and belong to the next block.  EB blocks are drained per pass (all of the
        // tile's blocks for 3-wide filters): one TMEM-load batch, one exchange and ONE barrier per pass -- the epilogue is
        // a latency chain (ncu: 350 instructions but ~1500 cycles per block), so fewer, fatter passes are what counts.
        constexpr int EB = (KW <= 3) ? (GB < 2 ? GB : 2) : 1;
        constexpr int XB = (KW - 1) * XROW;                  // floats per (quarter, block): KW-1 rows
#pragma unroll
        for (int bp = 0; bp < GB / EB; ++bp) {               // compile-time block index: valid[] / pix[] stay in registers
#pragma unroll 1
        for (int cc = 0; cc < NT / 16; ++cc) {
          const int blk0 = gb0 + bp * EB;
          uint32_t v[EB][KW][16];
#pragma unroll
          for (int e = 0; e < EB; ++e)
#pragma unroll
            for (int dw = 0; dw < KW; ++dw) tmem_ld_32x16(taddr0 + (blk0 + e) * NW + dw * NT + cc * 16, v[e][dw]);
          sm100::tmem_ld_wait();
          const uint32_t mine = xg_s + (uint32_t)(((xch_it & 1) * 4 + quarter) * (EB * XB)) * 4u;
          if (lane < KW - 1) {
#pragma unroll
            for (int e = 0; e < EB; ++e)
#pragma unroll
              for (int dw = 1; dw < KW; ++dw)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(mine + (uint32_t)(e * XB + lane * XROW + (dw - 1) * 16 + 4 * j) * 4u, v[e][dw][4 * j],
                               v[e][dw][4 * j + 1], v[e][dw][4 * j + 2], v[e][dw][4 * j + 3]);
          }
          // the four warps of this group (named barrier 1 + group); xch is double buffered, one barrier per pass
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
          const uint32_t next = xg_s + (uint32_t)(((xch_it & 1) * 4 + ((quarter + 1) & 3)) * (EB * XB)) * 4u;
          ++xch_it;
#pragma unroll
          for (int e = 0; e < EB; ++e) {
#pragma unroll
            for (int dw = 1; dw < KW; ++dw) {
              // row m + dw: lane + dw of this warp (shuffle), or row lane + dw - 32 of the next quarter (xch).  No
              // divergent branches: every lane reads xch (clamped address, mostly a broadcast) and selects.
              const int jn = lane + dw - 32;
              const uint32_t nrow = next + (uint32_t)(e * XB + (jn < 0 ? 0 : jn) * XROW + (dw - 1) * 16) * 4u;
              uint32_t nv[16];
#pragma unroll
              for (int j = 0; j < 4; ++j) ld_shared_v4(nrow + 16u * j, nv[4 * j], nv[4 * j + 1], nv[4 * j + 2], nv[4 * j + 3]);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sv = __shfl_down_sync(0xffffffffu, __uint_as_float(v[e][dw][i]), dw);
                v[e][0][i] = __float_as_uint(__uint_as_float(v[e][0][i]) + (jn >= 0 ? __uint_as_float(nv[i]) : sv));
              }
            }
          }
#pragma unroll
          for (int e = 0; e < EB; ++e)
            if (valid[bp * EB + e]) conv_store16(q, pix[bp * EB + e], nt * NT + cc * 16, v[e][0], s_bias + nt * NT + cc * 16);
        }
        }
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&tmem_empty[a_cur]);
    };
    int local = 0;
    if (p.slide) {
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        ConvUnit cu;
        cu.decode(p, u);
        for (int ot = 0; ot < cu.nf; ++ot, ++local)
          if (SPLIT || (local & 1) == eg) drain(0, cu.wb, cu.hb, cu.t0 + ot, cu.b);
      }
    } else {
      ConvTileIter it;
      it.init(p, blockIdx.x, gridDim.x);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local, it.advance(p))
        if (SPLIT || (local & 1) == eg) drain(it.nt, it.wb, it.hb, it.t, it.b);
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

typedef void (*ConvKernelFn)(const CUtensorMap, const ConvParams);

template <int KH, int KW, int NKS, bool PACK>
static ConvKernelFn pick_conv_kernel2(int nblk, int NT) {
  constexpr int F = PACK ? KW : 1;         // accumulator columns per block = F * NT; 2 * nblk * F * NT <= 512
  if (nblk == 1) {
    if constexpr (F * 16 <= 256) if (NT == 16) return conv_sm100_kernel<KH, KW, NKS, 1, 16, PACK>;
    if constexpr (F * 32 <= 256) if (NT == 32) return conv_sm100_kernel<KH, KW, NKS, 1, 32, PACK>;
    if constexpr (F * 64 <= 256) if (NT == 64) return conv_sm100_kernel<KH, KW, NKS, 1, 64, PACK>;
  } else if (nblk == 2) {
    if constexpr (2 * F * 16 <= 256) if (NT == 16) return conv_sm100_kernel<KH, KW, NKS, 2, 16, PACK>;
    if constexpr (2 * F * 32 <= 256) if (NT == 32) return conv_sm100_kernel<KH, KW, NKS, 2, 32, PACK>;
    if constexpr (2 * F * 64 <= 256) if (NT == 64) return conv_sm100_kernel<KH, KW, NKS, 2, 64, PACK>;
  } else if (nblk == 4) {
    if constexpr (PACK && 4 * F * 16 <= 256) if (NT == 16) return conv_sm100_kernel<KH, KW, NKS, 4, 16, PACK>;
  }
  return nullptr;
}

static ConvKernelFn pick_conv_kernel(const ConvPlan& p) {
  const int nks = p.CB / 16;
  if (p.kh == 3 && p.kw == 3) {
    if (p.pack) return nks == 1 ? pick_conv_kernel2<3, 3, 1, true>(p.nblk, p.NT) : nks == 2 ? pick_conv_kernel2<3, 3, 2, true>(p.nblk, p.NT) : nullptr;
    return nks == 1 ? pick_conv_kernel2<3, 3, 1, false>(p.nblk, p.NT) : nks == 2 ? pick_conv_kernel2<3, 3, 2, false>(p.nblk, p.NT) : nullptr;
  }
  if (p.kh == 1 && p.kw == 1 && !p.pack)
    return nks == 1 ? pick_conv_kernel2<1, 1, 1, false>(p.nblk, p.NT) : nks == 2 ? pick_conv_kernel2<1, 1, 2, false>(p.nblk, p.NT) : nullptr;
  if (p.kh == 7 && p.kw == 7 && nks == 1 && p.NT == 16) {
    if (p.pack) return p.nblk == 1 ? conv_sm100_kernel<7, 7, 1, 1, 16, true> : (p.nblk == 2 ? conv_sm100_kernel<7, 7, 1, 2, 16, true> : nullptr);
    return p.nblk == 1 ? conv_sm100_kernel<7, 7, 1, 1, 16, false> : (p.nblk == 2 ? conv_sm100_kernel<7, 7, 1, 2, 16, false> : nullptr);
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
int conv_tc_supported(const vvae_conv_args& a, int which) {
  if (a.dtype != VVAE_BF16 || which > 1) return 0;
  if (!a.wprep) return 0;
  ConvPlan p;
  if (!make_plan(a, which, p) || !pick_conv_kernel(p)) return 0;
  const void* in = which == 0 ? a.x : a.y;
  if ((uintptr_t)in % 16) return 0;
  return 1;
}

int conv_tc_launch(const vvae_conv_args& a, int which, cudaStream_t s) {
  ConvParams q;
  if (!make_plan(a, which, q.pl)) {
    set_error("conv3d: shape not supported by the tensor-core path");
    return VVAE_ERR_UNSUPPORTED;
  }
  const ConvPlan& p = q.pl;
  const void* in = which == 0 ? a.x : a.y;
  const long long in_ld = which == 0 ? a.x_ld : a.y_ld;
  CUtensorMap tm;
  uint64_t dims[5] = {(uint64_t)p.Cin_pad, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.T, (uint64_t)a.B};
  uint64_t str[4] = {(uint64_t)in_ld * 2, (uint64_t)a.W * in_ld * 2, (uint64_t)a.H * a.W * in_ld * 2,
                     (uint64_t)a.T * a.H * a.W * in_ld * 2};
  uint32_t box[5] = {(uint32_t)p.CB, (uint32_t)p.P, (uint32_t)p.rows, 1, 1};
  int rc = encode_tmap_nd_bf16(&tm, in, 5, dims, str, box, p.swizzle_bytes);
  if (rc) return rc;
  q.wimg = (const bf16*)a.wprep;
  if (which == 0) {
    q.y = (bf16*)a.y; q.y_ld = a.y_ld; q.bias = a.bias; q.mode = a.epilogue; q.aux = (const bf16*)a.aux_in; q.ld_aux = a.ld_aux;
  } else {
    q.y = (bf16*)const_cast<void*>(a.x); q.y_ld = a.x_ld; q.bias = nullptr; q.mode = VVAE_EPI_NONE; q.aux = nullptr; q.ld_aux = 0;
  }
  q.store_c = p.Cout;
  q.dbg = (int)g_dbg[14];
  if (a.pad_out) {
    const int padded = (p.Cout + 15) / 16 * 16;
    if (q.y_ld < padded) {
      set_error("conv3d: pad_out needs %d channels of storage per voxel, row stride is %lld", padded, q.y_ld);
      return VVAE_ERR_INVALID;
    }
    q.store_c = padded;
  }
  ConvKernelFn kern = pick_conv_kernel(p);
  if (!kern) {
    set_error("conv3d: no tensor-core kernel instance for this shape");
    return VVAE_ERR_UNSUPPORTED;
  }
  {
    static std::mutex mu;
    static std::set<ConvKernelFn> configured;
    std::lock_guard<std::mutex> lk(mu);
    if (!configured.count(kern)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        set_error("conv3d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return VVAE_ERR_CUDA;
      }
      configured.insert(kern);
    }
  }
  const int grid = std::min(p.slide ? p.units : p.total_tiles, num_sms());
  launch_pdl(kern, dim3(grid), dim3(CONV_THREADS), (size_t)p.smem_bytes, s, tm, q);
  return check_launch("conv_sm100");
}

}  // namespace vvae

using namespace vvae;

extern "C" {

long long vvae_conv3d_wprep_bytes(const vvae_conv_args* a, int which) {
  if (!a || a->dtype != VVAE_BF16 || which < 0 || which > 1) return 0;
  ConvPlan p;
  if (!make_plan(*a, which, p) || !pick_conv_kernel(p)) return 0;
  return (long long)p.nNT * p.kt * p.nCB * p.w_stride;
}

int vvae_conv3d_wprep(const vvae_conv_args* a, int which, void* out, vvae_stream_t stream) {
  VVAE_REQUIRE(a && out && a->w, "conv3d_wprep: null pointer");
  ConvPlan p;
  VVAE_REQUIRE(a->dtype == VVAE_BF16 && make_plan(*a, which, p), "conv3d_wprep: shape not supported by the tensor-core path");
  const long long total = (long long)p.nNT * p.kt * p.nCB * (p.w_stride / 2);
  const int blocks = (int)std::min<long long>(cdiv(total, 256), num_sms() * 8);
  conv_wprep_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const bf16*)a->w, (bf16*)out, p, which, a->Cin, a->Cout);
  return check_launch("conv_wprep");
}

}  // extern "C"
