/*
 * libvvae -- C ABI of the B200-native video-VAE hot path (sm_100a).
 *
 * The reference (floatingtrees/video-VAE) is pure Python on JAX/Flax and has no
 * plugin / FFI interface; its boundary is the module API of train/layers.py,
 * train/model.py and train/unet.py.  These entry points are what a binding for
 * that path would call, one per fused operator of the hot path; each comment
 * cites the reference lines the operator replaces.  INTEGRATION.md shows the
 * reference-side stub (a ctypes / jax.ffi call) a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (VVAE_OK) or a vvae_status; the message is
 *     available from vvae_last_error() (thread local).  Nothing throws.
 *   - plain device pointers, sizes and element strides; no library types.
 *   - nothing allocates, frees or synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*).
 *   - `dtype` is the activation/compute dtype T (VVAE_F32 or VVAE_BF16).
 *     Parameters, parameter gradients, statistics and loss scalars are fp32.
 *     Gradient outputs named "accumulate" are added into (+=), so parameter
 *     gradients live in one flat fp32 buffer that is zeroed once per step.
 *   - activations are channels-last: tokens [b,t,hw,c], voxels [b,t,h,w,c].
 *   - there is NO CPU path: without a CUDA device every compute entry point
 *     fails with VVAE_ERR_CUDA.
 */
#ifndef VVAE_H_
#define VVAE_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vvae_stream_t;

enum vvae_status { VVAE_OK = 0, VVAE_ERR_INVALID = 1, VVAE_ERR_CUDA = 2, VVAE_ERR_UNSUPPORTED = 3 };
enum vvae_dtype { VVAE_F32 = 0, VVAE_BF16 = 1 };

/* GEMM epilogues */
enum vvae_epilogue {
  VVAE_EPI_NONE = 0,     /* C = acc (+bias)                                         */
  VVAE_EPI_SILU = 1,     /* aux_out = acc+bias (pre-activation); C = silu(acc+bias) */
  VVAE_EPI_RESIDUAL = 2, /* C = acc + bias + aux_in            (aux_in may alias C) */
  VVAE_EPI_DSILU = 3,    /* C = acc * silu'(aux_in)   (dgrad through MLP SiLU)      */
  VVAE_EPI_QKNORM_ROPE = 4 /* the QKV projection of train/layers.py:160-166 in one call: C = acc + bias = q|k|v
                            * [M, 3*heads*hd] and aux_out [M, 2*heads*hd] = rope(LN(q)) | rope(LN(k)) (per-head LayerNorm
                            * without bias, eps qk_eps, then RoPE at position (row / rope_pos_div) % rope_pos_mod); the
                            * tcgen05 kernel does it in its epilogue (one 64-column chunk = one head in one thread's
                            * registers), every other case runs the GEMM and then vvae_qknorm_rope_fwd */
};
enum vvae_backend { VVAE_BACKEND_AUTO = 0, VVAE_BACKEND_SIMT = 1, VVAE_BACKEND_TCGEN05 = 2 };

const char* vvae_last_error(void);
int vvae_version(void);
/* 1 if a CUDA device with compute capability 10.x is visible, else 0. */
int vvae_device_ok(void);
/* Debug/tuning knobs, all 0 by default (bring-up scripts only; never set by the product path).  Keys 0-6: grid size,
 * UMMA descriptor fields and the N-tile of the tcgen05 GEMM (csrc/gemm_sm100.cu); 8: force single-CTA GEMM tiles;
 * 7: LayerNorm D = 768 bf16 kernel choice, low nibble = backward, next nibble = forward (0: bulk-copy staged streaming
 *    kernel, 1: register-staged kernel, 2: the other warps x stages split of the streaming kernel);
 * 9: keep short sequences (L <= 16) on the packed tcgen05 attention tiles instead of the one-warp kernels;
 * 10: tcgen05 GEMM TIMING ablations, results are wrong (bit 1: no A-tile TMA loads, 2: no B-tile loads, 4: no loads and
 *     the issuer never waits for operands, 8: accumulators never drained; 16: record counters for vvae_debug_get);
 * 11: launch without programmatic dependent launch (every kernel fully serialized behind its predecessor);
 * 12: 192-column GEMM tiles for N = 768 dgrads (measured slower than 256: kept for the record);
 * 13: conv3d fwd/dgrad: one MMA per filter tap (round-1 kernels) instead of the kw taps packed into N;
 * 14: conv3d fwd/dgrad TIMING ablations, results are wrong (bit 1: no global stores, 2: epilogue releases the accumulator
 *     at once, 4: no input-tile TMA, 8: no MMAs);
 * 15: conv3d fwd/dgrad: stream the filter taps with every stage (round-1 behaviour) instead of keeping the whole
 *     weight image resident in shared memory;
 * 16: attention backward, unmasked L = 256: the one-(sequence, head)-per-CTA kernel instead of the persistent one;
 * 17: tcgen05 GEMM: 1 = whole 256-column tiles in a partial last wave (no column slicing), 2 / 4 = force that many slices;
 * 18: attention forward, unmasked L = 256: the one-tile-per-CTA kernel instead of the persistent one (key 10 bit 32 with
 *     the persistent kernel: issue both score tiles at the start of a unit instead of the staggered pipeline). */
int vvae_debug_set(int key, long long value);
/* what = 0: counters of the last tcgen05 GEMM launched with vvae_debug_set(10, ... | 16): {clock64 ticks, globaltimer ns,
 * MMAs issued} of CTA 0's issuing thread (synchronises the device).
 * what = 1: `out4` must hold 32 values: the clock64 timeline of one CTA of the last tcgen05 attention backward launched
 * with vvae_debug_set(10, 16) (CTA index = vvae_debug_set(0, n)); slots are listed in scripts/attn_timeline.py. */
int vvae_debug_get(int what, unsigned long long* out4);

/* ---- elementwise plumbing ------------------------------------------------- */
/* dst[i] = (dst_dtype) src[i]; the per-step fp32 -> bf16 parameter shadow copy. */
int vvae_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, vvae_stream_t stream);
int vvae_fill_f32(float* dst, float value, long long n, vvae_stream_t stream);
/* out[n] += sum_rows x[row*ld + n]  (bias gradients of every Linear / Conv). */
int vvae_colsum(const void* x, long long ld, long long rows, int n, float* out_accum, int dtype, vvae_stream_t stream);

/* ---- GEMM: every nnx.Linear of the path (train/layers.py:24,46-47,52,160,170,193,195;
 *      train/model.py:53-58,91), forward, dgrad and wgrad ------------------------------
 * C[M,N] = epilogue( op(A)[M,K] * op(B)[K,N] ), fp32 accumulation.
 *   op(A)[m,k] = transA ? A[k*lda + m] : A[m*lda + k]
 *   op(B)[k,n] = transB ? B[n*ldb + k] : B[k*ldb + n]     (Flax kernels are (in,out): transB = 0)
 * A and B have dtype `dtype`; C has `out_dtype` (dtype or VVAE_F32); aux tensors have `dtype`.
 * accumulate != 0 (fp32 C only): C += result, used for weight gradients (split-K atomics).  */
typedef struct vvae_gemm_args {
  int M, N, K;
  const void* A; long long lda; int transA;
  const void* B; long long ldb; int transB;
  void* C; long long ldc;
  int dtype, out_dtype;
  const float* bias;
  int epilogue;
  const void* aux_in; long long ld_aux_in;
  void* aux_out; long long ld_aux_out;
  int accumulate;
  int backend;
  float* bsum_accum; /* optional, fp32 [N]: += sum_k op(B)[k, n].  For a weight gradient X^T . dY this is the Linear's bias
                      * gradient, summed from the dY tiles while they sit in shared memory (no second pass over dY). */
  /* VVAE_EPI_QKNORM_ROPE only (arguments as in vvae_qknorm_rope_fwd) */
  const float* qk_q_scale; const float* qk_k_scale;
  const void* rope_cos; const void* rope_sin;
  long long rope_pos_div; int rope_pos_mod;
  int qk_heads, qk_hd;
  float qk_eps;
} vvae_gemm_args;
int vvae_gemm(const vvae_gemm_args* args, vvae_stream_t stream);
/* 1 if vvae_gemm would run these arguments on the tcgen05 kernel, 0 if on the generic SIMT kernel. */
int vvae_gemm_uses_tcgen05(const vvae_gemm_args* args);

/* ---- LayerNorm (nnx.LayerNorm eps 1e-6: train/layers.py:17,153,178) ------- */
int vvae_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                       long long rows, int D, float eps, int dtype, vvae_stream_t stream);
/* dx = LN'(dy) (+ dres);  dgamma/dbeta accumulate (may be NULL). */
int vvae_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                       const void* dres, void* dx, float* dgamma_accum, float* dbeta_accum,
                       long long rows, int D, int dtype, vvae_stream_t stream);

/* ---- QK-LayerNorm + RoPE (train/layers.py:104-129,164-166) ----------------
 * qkv: [rows, 3*H*hd] (q|k|v).  qk_out: [rows, 2*H*hd] = rope(LN(q)) | rope(LN(k)).
 * position of row r = (r / pos_div) % pos_mod; cos/sin: [>=pos_mod, hd] in the compute dtype T (the reference casts
 * the tables to q's dtype before use, train/layers.py:124-127). */
int vvae_qknorm_rope_fwd(const void* qkv, void* qk_out, const float* q_scale, const float* k_scale,
                         const void* cos_tab, const void* sin_tab, long long rows, int heads, int hd,
                         long long pos_div, int pos_mod, float eps, int dtype, vvae_stream_t stream);
/* dqkv: [rows, 3*H*hd]; on entry its q|k part holds d(rotated q|k), on exit d(raw q|k); v part untouched.
 * dbias_qk_accum (optional, fp32 [2*H*hd]) += column sums of the produced d(raw q|k): the q|k part of the QKV
 * projection's bias gradient, accumulated while the gradient is in registers. */
int vvae_qknorm_rope_bwd(void* dqkv, const void* qkv, const float* q_scale, const float* k_scale,
                         const void* cos_tab, const void* sin_tab, float* dq_scale_accum, float* dk_scale_accum,
                         float* dbias_qk_accum,
                         long long rows, int heads, int hd, long long pos_div, int pos_mod, float eps, int dtype,
                         vvae_stream_t stream);

/* ---- attention (jax.nn.dot_product_attention: train/layers.py:168; mask semantics of
 *      train/attention_mask_tests.py and training_loop_adversarial.py:127-130) ----------
 * Sequences are views into token-major buffers: sequence s = (o, i), o in [0,n_outer), i in [0,n_inner);
 * token(s, l) = o*tok_stride_outer + i*tok_stride_inner + l*tok_stride_pos; element (token, head h, d) of
 * tensor X is X[token*X_rs + h*hd + d].  Temporal attention over [b,t,hw,*]: n_outer=b, n_inner=hw,
 * strides (t*hw, 1, hw); spatial: n_outer=b*t, n_inner=1, strides (hw, 0, 1).
 * mask (uint8, nonzero = attend, NULL = none) is indexed
 *   mask[(s / mask_seq_div)*ms_seq + h*ms_head + q*ms_q + k*ms_k];
 * masked logits are replaced by -0.7*FLT_MAX (a fully masked row attends uniformly, as in JAX). */
typedef struct vvae_attn_args {
  int n_outer, n_inner, L, heads, hd;
  long long tok_stride_outer, tok_stride_inner, tok_stride_pos;
  const void *q, *k, *v; void* o;
  long long q_rs, k_rs, v_rs, o_rs;
  float* lse;                 /* [n_seq, heads, L] */
  const unsigned char* mask; long long mask_seq_div, ms_seq, ms_head, ms_q, ms_k;
  float scale;
  int dtype;
  /* backward only */
  const void* d_o; long long do_rs;
  void *dq, *dk, *dv; long long dq_rs, dk_rs, dv_rs;
  float* delta;               /* workspace [n_seq, heads, L] */
  int backend;
} vvae_attn_args;
int vvae_attn_fwd(const vvae_attn_args* args, vvae_stream_t stream);
int vvae_attn_bwd(const vvae_attn_args* args, vvae_stream_t stream);

/* ---- patchify / unpatchify (train/layers.py:21-23,48-51) -------------------
 * video [b,t,H,W,C] (dtype in_dtype) -> tokens [b,t,(h w),(p1 p2 c)] (dtype). */
int vvae_patchify(const void* video, int in_dtype, void* tokens, int b_t, int H, int W, int C, int P, int dtype,
                  vvae_stream_t stream);
/* tokens [b_t,(h w),(p1 p2 cu)] <-> voxels [b_t,(h p1),(w p2),cu]; dir 0 = tokens->voxels, 1 = voxels->tokens. */
int vvae_pixel_shuffle(const void* src, void* dst, int b_t, int H, int W, int CU, int P, int dir, int dtype,
                       vvae_stream_t stream);
/* The same with `vox_ld` >= cu elements between consecutive voxels (bf16): dir 0 also writes the pad channels
 * [cu, vox_ld) as zeros, dir 1 ignores them.  The U-Net input (train/unet.py:155-160, 12 channels) is produced at a
 * pitch of 16, the channel block the tensor-core convolutions gather. */
int vvae_pixel_shuffle_pitched(const void* src, void* dst, int b_t, int H, int W, int CU, int P, int dir, long long vox_ld,
                               int dtype, vvae_stream_t stream);

/* ---- conv3d, NDHWC, 'SAME', stride 1 (nnx.Conv: train/unet.py:13-21,111-113,144-153) ----
 * x [B,T,H,W,Cin] with channel stride x_ld (>= Cin, lets x be a slice of a concat buffer),
 * w [kt,kh,kw,Cin,Cout] (dtype), bias fp32 [Cout] or NULL, y [B,T,H,W,Cout] with stride y_ld.
 * epilogue: VVAE_EPI_NONE or VVAE_EPI_RESIDUAL (y = conv + bias + aux_in, aux stride ld_aux). */
typedef struct vvae_conv_args {
  int B, T, H, W, Cin, Cout, kt, kh, kw;
  const void* x; long long x_ld;
  const void* w;
  const float* bias;
  void* y; long long y_ld;
  int epilogue; const void* aux_in; long long ld_aux;
  float* dw_accum; /* wgrad: fp32 [kt,kh,kw,Cin,Cout], += */
  int dtype;
  int backend;
  const void* wprep; /* tensor-core path: weight image from vvae_conv3d_wprep (NULL -> generic kernel) */
  int pad_out;       /* != 0: the produced tensor has storage for ceil16(channels) channels per voxel and the kernel also
                      * writes zeros to the pad channels (so a following tensor-core conv may gather them) */
} vvae_conv_args;
int vvae_conv3d_fwd(const vvae_conv_args* args, vvae_stream_t stream);
/* dgrad: reads dy from args->y (stride y_ld), writes dx to args->x (stride x_ld; cast away const). */
int vvae_conv3d_dgrad(const vvae_conv_args* args, vvae_stream_t stream);
/* wgrad: dw_accum += x^T (*) dy with x = args->x, dy = args->y. */
int vvae_conv3d_wgrad(const vvae_conv_args* args, vvae_stream_t stream);
/* Tensor-core (tcgen05) path: the weights are re-laid out once per optimizer step into the swizzled per-stage image
 * the kernel streams with bulk copies.  which = 0 forward, 1 dgrad.  _bytes returns 0 when the shape is not supported
 * by the tensor-core path (the generic kernel is used then). */
long long vvae_conv3d_wprep_bytes(const vvae_conv_args* args, int which);
int vvae_conv3d_wprep(const vvae_conv_args* args, int which, void* out, vvae_stream_t stream);

/* ---- ConvTranspose k=s=(1,2,2) (nnx.ConvTranspose: train/unet.py:61-69) ----
 * x [V,Cin] voxels of a [B_T,H,W] grid; y [B_T,2H,2W,Cout] with channel stride y_ld;
 * w [1,2,2,Cin,Cout]; out[2i+a,2j+c] = x[i,j] . w[1-a,1-c] + bias.
 * workspace: caller-provided scratch of vvae_convT122_workspace_bytes(...) bytes (256-byte aligned) for the
 * tensor-core path (re-laid-out weights, the dense [V, 4*Cout] GEMM-side image of y / dy, fp32 weight-gradient
 * staging); NULL / too small -> generic kernel. */
long long vvae_convT122_workspace_bytes(int b_t, int H, int W, int Cin, int Cout);
int vvae_convT122_fwd(const void* x, const void* w, const float* bias, void* y, long long y_ld,
                      int b_t, int H, int W, int Cin, int Cout, int dtype, void* workspace, long long workspace_bytes,
                      vvae_stream_t stream);
/* dx [V,Cin] from dy (stride dy_ld); dw_accum fp32 [1,2,2,Cin,Cout] += ; (bias grad: vvae_colsum on dy). */
int vvae_convT122_bwd(const void* dy, long long dy_ld, const void* x, const void* w, void* dx, float* dw_accum,
                      int b_t, int H, int W, int Cin, int Cout, int dtype, void* workspace, long long workspace_bytes,
                      vvae_stream_t stream);

/* ---- GroupNorm + SiLU (nnx.GroupNorm eps 1e-6 + nnx.silu: train/unet.py:22-29) ----
 * x [B, S, C] (S = t*h*w voxels per sample), groups G, statistics per (sample, group).
 * fwd writes mean/rstd fp32 [B,G] (workspace `stats` fp32 [B,G,2] is scratch) and
 * y = silu((x-mean)*rstd*gamma+beta) with channel stride y_ld. */
int vvae_groupnorm_silu_fwd(const void* x, void* y, long long y_ld, const float* gamma, const float* beta,
                            float* mean, float* rstd, float* stats, int B, long long S, int C, int G, float eps,
                            int dtype, vvae_stream_t stream);
/* dy has channel stride dy_ld; dx contiguous [B,S,C]; dgamma/dbeta accumulate.  dx_colsum_accum (optional, fp32 [C])
 * += per-channel sums of the produced dx: the bias gradient of the convolution that feeds this norm
 * (ConvBlock3D, train/unet.py:13-29), taken while dx is in registers instead of a second pass over it. */
int vvae_groupnorm_silu_bwd(const void* dy, long long dy_ld, const void* x, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, void* dx, float* dgamma_accum, float* dbeta_accum,
                            float* stats, float* dx_colsum_accum, int B, long long S, int C, int G, int dtype,
                            vvae_stream_t stream);

/* ---- max_pool (1,2,2) (train/unet.py:50) and channel concat (train/unet.py:80) ---- */
/* x [b_t,H,W,C] with channel stride x_ld -> y [b_t,H/2,W/2,C] contiguous. */
int vvae_maxpool122_fwd(const void* x, long long x_ld, void* y, int b_t, int H, int W, int C, int dtype,
                        vvae_stream_t stream);
/* dx[b_t,H,W,C] (contiguous) = route(dy) to the first maximum of each window (+ dskip, stride dskip_ld, if given). */
int vvae_maxpool122_bwd(const void* x, long long x_ld, const void* dy, const void* dskip, long long dskip_ld,
                        void* dx, int b_t, int H, int W, int C, int dtype, vvae_stream_t stream);
/* dst[r*dst_ld + dst_off + c] = src[r*src_ld + src_off + c], c < C (channel-slice copy for concat / split). */
int vvae_copy_channels(const void* src, long long src_ld, long long src_off, void* dst, long long dst_ld,
                       long long dst_off, long long rows, int C, int dtype, vvae_stream_t stream);

/* ---- latent head (train/model.py:53-59,119-133; train/layers.py:238-252) ---- */
/* lv = log(softplus(a)) */
int vvae_softplus_log_fwd(const void* a, void* lv, long long n, int dtype, vvae_stream_t stream);
/* da = dlv * sigmoid(a) / softplus(a) */
int vvae_softplus_log_bwd(const void* dlv, const void* a, void* da, long long n, int dtype, vvae_stream_t stream);
/* Selection gate: logit[bt] = s1[bt,:] . w2 + b2 + 1;  p = sigmoid((logit + g)/temp), g = log(u/(1-u)) (train)
 * or 0 (eval); sel = round_half_even(p).  u: explicit fp32 [bt] or NULL -> Philox(seed, offset). */
int vvae_selection_fwd(const void* s1, const float* w2, const float* b2, const float* u, unsigned long long seed,
                       unsigned long long offset, int train, float temperature, float* logit, float* p, float* sel,
                       int bt, int hw, int dtype, vvae_stream_t stream);
/* z = mean + eps*exp(lv/2) (train) | mean (eval);  c = fill*(1-sel) + z*sel.
 * eps: explicit fp32 [n_tok, Dl] or NULL -> Philox(seed, offset) (written to eps_out if non-NULL).
 * c32: fp32 output (the reference returns fp32); cT: compute-dtype copy feeding the decoder (may be NULL). */
int vvae_reparam_gate_fwd(const void* mean, const void* logvar, const float* eps, unsigned long long seed,
                          unsigned long long offset, float* eps_out, const float* sel, const float* fill,
                          float* c32, void* cT, long long n_tok, int tok_per_frame, int Dl, int train, int dtype,
                          vvae_stream_t stream);
/* Backward of reparam+gate:
 *   dmean = dc*sel (+ dmean_in) ; dlogvar = dc*sel*eps*0.5*exp(lv/2) (+ dlogvar_in)
 *   dfill_accum[d] += sum dc*(1-sel) ;  dsel_accum[frame] += sum_{hw,d} dc*(z - fill)
 * dc has dtype T; dmean_in / dlogvar_in (dtype T, may be NULL, may alias the outputs) are the gradients arriving
 * from other consumers of mean / logvar (the KL term, the selection head). */
int vvae_reparam_gate_bwd(const void* dc, const void* mean, const void* logvar, const float* eps, const float* sel,
                          const float* fill, const void* dmean_in, const void* dlogvar_in, void* dmean, void* dlogvar,
                          float* dfill_accum, float* dsel_accum, long long n_tok, int tok_per_frame, int Dl, int train,
                          int dtype, vvae_stream_t stream);

/* ---- losses (train/legacy/training_loop_adversarial.py:94-124; MAE train/rl_nonadversarial.py:114-117) ----
 * video [B,T,per_frame] (dtype video_dtype), recon same shape (dtype); frame_mask fp32 [B*T] (0/1); inv_len fp32 [B].
 * out2[0] += sum_{b,p} round_T(sum_t ((video-recon)*m)^2) * inv_len[b];  out2[1] += same with |.|
 * (the caller divides by B*per_frame to get the means MSE / MAE). */
int vvae_recon_loss_fwd(const void* video, int video_dtype, const void* recon, const float* frame_mask,
                        const float* inv_len, float* out2, int B, int T, long long per_frame, int dtype,
                        vvae_stream_t stream);
/* Per-sample form of the above for the RL loss (train/rl_nonadversarial.py:114-121,161): out_b2 fp32 [B,2],
 * out_b2[b,0] += sum_p round_T(sum_t e^2) * inv_len[b], out_b2[b,1] the same with |e|.  (A per-sample upstream gradient
 * g[b] is applied in vvae_recon_loss_bwd by passing inv_len[b]*g[b].) */
int vvae_recon_loss_per_sample_fwd(const void* video, int video_dtype, const void* recon, const float* frame_mask,
                                   const float* inv_len, float* out_b2, int B, int T, long long per_frame, int dtype,
                                   vvae_stream_t stream);
/* drecon[b,t,p] = -(2*w_mse*e + w_mae*sign(e)) * m[b,t] * inv_len[b] * inv_count * (*gscale), e = (video-recon)*m.
 * gscale: optional DEVICE fp32 scalar (the upstream d(loss); NULL = 1) so the step never synchronises with the host. */
int vvae_recon_loss_bwd(const void* video, int video_dtype, const void* recon, const float* frame_mask,
                        const float* inv_len, float w_mse, float w_mae, float inv_count, const float* gscale,
                        void* drecon, int B, int T,
                        long long per_frame, int dtype, vvae_stream_t stream);
/* out1[0] += sum 0.5*(exp(lv)-1-lv+mean^2) * frame_w[frame], frame_w = m/len   (caller divides by numel). */
int vvae_kl_fwd(const void* mean, const void* logvar, const float* frame_w, float* out1, long long n_tok,
                int tok_per_frame, int Dl, int dtype, vvae_stream_t stream);
/* Per-sample KL (train/rl_nonadversarial.py:145-146): out_b fp32 [B], out_b[b] += sum over the sample's tok_per_sample
 * tokens.  (Per-sample upstream gradients go through frame_w in vvae_kl_bwd.) */
int vvae_kl_per_sample_fwd(const void* mean, const void* logvar, const float* frame_w, float* out_b, int B,
                           long long tok_per_sample, int tok_per_frame, int Dl, int dtype, vvae_stream_t stream);
/* dmean = s*frame_w[frame]*mean ; dlogvar = s*frame_w[frame]*0.5*(exp(lv)-1), s = scale * (*gscale)  (gscale: optional
 * device scalar, NULL = 1; scale = gamma2 / numel). */
int vvae_kl_bwd(const void* mean, const void* logvar, const float* frame_w, float scale, const float* gscale, void* dmean,
                void* dlogvar,
                long long n_tok, int tok_per_frame, int Dl, int dtype, vvae_stream_t stream);

/* Materialise the Philox draws the fused kernels would make from (seed, offset): kind 0 = the N(0,1) noise of
 * vvae_reparam_gate_fwd, kind 1 = the U(0,1) draws of vvae_selection_fwd.  Lets a captured CUDA graph take per-step
 * randomness from a buffer (the explicit `eps` / `u` arguments) instead of kernel-argument constants. */
int vvae_philox_fill(float* out, long long n, unsigned long long seed, unsigned long long offset, int kind,
                     vvae_stream_t stream);

/* ---- VGG-16 perceptual features ("next" row f4: train/vgg_tests.py:8-68 over flaxmodels 0.1.3 VGG16) ----
 * The 3x3 convolutions run on vvae_conv3d_* with kt = 1; these are the pointwise pieces around them. */
/* y = max(x, 0);  dx = y > 0 ? dy : 0.  n elements, 16-byte aligned pointers. */
int vvae_relu_fwd(const void* x, void* y, long long n, int dtype, vvae_stream_t stream);
int vvae_relu_bwd(const void* dy, const void* y, void* dx, long long n, int dtype, vvae_stream_t stream);
/* x [voxels,3] RGB in [0,1] (x_dtype) -> y [voxels,y_ld] (dtype): (round_T(x) - mean)/std per ImageNet channel
 * statistics in channels 0..2, zeros in the pad channels 3..y_ld-1 (the tensor-core conv gathers 16 channels). */
int vvae_vgg_preprocess_fwd(const void* x, int x_dtype, void* y, long long voxels, int y_ld, int dtype,
                            vvae_stream_t stream);
/* dx [voxels,3] = dy[voxels,:3] / std */
int vvae_vgg_preprocess_bwd(const void* dy, void* dx, long long voxels, int dy_ld, int dtype, vvae_stream_t stream);

/* ---- optimizer ("next" row f1: optax.chain(clip_by_global_norm, adam), train/rl_nonadversarial.py:241-253) ---- */
/* out[0] += sum g^2 */
int vvae_sumsq_f32(const float* g, long long n, float* out1, vvae_stream_t stream);
/* Same sum with a fixed summation order (bit-reproducible: data-parallel replicas compute the identical clip scale).
 * partials: caller scratch of vvae_sumsq_partials(n) floats. */
int vvae_sumsq_partials(long long n);
int vvae_sumsq_f32_det(const float* g, long long n, float* partials, float* out1, vvae_stream_t stream);
/* Adam with bias correction on flat fp32 buffers; grad scaled by min(1, clip / sqrt(*gnorm_sq)) if gnorm_sq.
 * shadow_bf16 (may be NULL): n bf16 values that receive the updated parameters rounded to nearest even -- the copy of the
 * weights the bf16 kernels read, written in the same pass instead of by a vvae_cast over the whole buffer. */
int vvae_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float b1,
                   float b2, float eps, int step, const float* gnorm_sq, float clip, float grad_scale,
                   vvae_stream_t stream);

/* ---- scratch sizes in one place (vvae_workspace_bytes(op, shape...)) ----
 * Nothing in this library allocates: the three entry points that need scratch take it from the caller.  Bytes of scratch
 * (0 if none is needed for that shape, -1 for an unknown op or a malformed shape):
 *   VVAE_WS_CONVT122   dims = {b_t, H, W, Cin, Cout}    workspace of vvae_convT122_{fwd,bwd}  (= vvae_convT122_workspace_bytes)
 *   VVAE_WS_SUMSQ_DET  dims = {n}                       partials of vvae_sumsq_f32_det        (= 4 * vvae_sumsq_partials(n))
 *   VVAE_WS_ATTN_BWD   dims = {n_seq, heads, L}         vvae_attn_args.delta of vvae_attn_bwd (fp32 [n_seq, heads, L]) */
enum vvae_workspace_op { VVAE_WS_CONVT122 = 0, VVAE_WS_SUMSQ_DET = 1, VVAE_WS_ATTN_BWD = 2 };
long long vvae_workspace_bytes(int op, const long long* dims, int ndims);

/* ---- gradient exchange of the data-parallel step ----
 * The reference's all-reduce is implicit in its jitted SPMD step (claude_distributed/distributed_train.py:107-109,
 * 378-380,412) and its start-up replication is broadcast_one_to_all (:339).  These calls give a host without
 * torch.distributed the same two collectives over the flat gradient / parameter buffers.  They wrap NCCL, resolved with
 * dlopen at the first call (VVAE_ERR_UNSUPPORTED if libnccl.so.2 cannot be loaded); one communicator per process / GPU.
 * vvae_comm_init is a collective over all `world` ranks and binds the calling thread's current device; it is the one
 * entry point of this library that blocks.  The Python package itself uses torch.distributed (ddp.py); ddp.NativeComm
 * wraps these calls. */
typedef struct vvae_comm* vvae_comm_t;
/* rank 0: fill the 128-byte rendezvous token; the host carries it to every other rank */
int vvae_comm_unique_id(void* id128);
int vvae_comm_init(vvae_comm_t* comm, const void* id128, int rank, int world);
int vvae_comm_rank(vvae_comm_t comm, int* rank, int* world);
/* in place over `count` elements of `dtype` (VVAE_F32 / VVAE_BF16): sum, or mean over the ranks if average != 0
 * (the reference's loss is a mean over the GLOBAL batch => mean of the per-rank gradients) */
int vvae_comm_allreduce(vvae_comm_t comm, void* buf, long long count, int dtype, int average, vvae_stream_t stream);
/* in place: every rank's buf becomes root's */
int vvae_comm_broadcast(vvae_comm_t comm, void* buf, long long count, int dtype, int root, vvae_stream_t stream);
int vvae_comm_destroy(vvae_comm_t comm);

#ifdef __cplusplus
}
#endif
#endif /* VVAE_H_ */
