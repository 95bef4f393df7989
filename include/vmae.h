/*
 * vmae.h — C ABI of libvmae_sm100a.so: hand-written sm_100a (B200) kernels for the video-MAE training step.
 *
 * The reference (lzc452/SSL-VIT-VIDEO-ANALYTICS) is pure PyTorch and has no FFI; each entry point below replaces the
 * PyTorch op(s) cited next to it (paths relative to the reference root). INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (the library never allocates, frees or retains memory);
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous and stream-ordered, no host sync inside;
 *  - return value: 0 on success, negative on error (VMAE_ERR_*); vmae_last_error() gives the message (thread-local);
 *  - there is no CPU or other-backend fallback: unsupported shapes fail with VMAE_ERR_UNSUPPORTED;
 *  - bf16 tensors are raw uint16 storage (__nv_bfloat16); "ld*" are leading dimensions in ELEMENTS.
 */
#ifndef VMAE_H_
#define VMAE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define VMAE_ABI_VERSION 3

#define VMAE_ERR_BAD_ARG (-1)
#define VMAE_ERR_UNSUPPORTED (-2)
#define VMAE_ERR_CUDA (-3)
#define VMAE_ERR_DRIVER (-4)

#define VMAE_DTYPE_BF16 0
#define VMAE_DTYPE_F32 1
#define VMAE_DTYPE_U8 2 /* only as the 8-bit gelu' side tensor of the GEMM (VMAE_ACT_GELU_Q8 / VMAE_ACT_AUX_Q8) */

#define VMAE_MAJOR_K 0  /* contraction dimension contiguous: A is [M,K], B is [N,K] row-major */
#define VMAE_MAJOR_MN 1 /* M/N dimension contiguous:        A is [K,M], B is [K,N] row-major */

#define VMAE_ACT_NONE 0
#define VMAE_ACT_GELU 1 /* exact erf GELU, as nn.GELU() / activation="gelu" in the reference */
/* The saved gelu'(pre-activation) as ONE BYTE per element: q = round((g + 0.13) * 255 / 1.26), g in [-0.129, 1.129]; step 0.0049, i.e.
 * |error| <= 0.0025 — the size of bf16 rounding for the typical g ~ 0.5..1. The dual-output GELU GEMM is bound by HBM WRITE bandwidth
 * (3.9 TB/s write-only on this part): 462 MB instead of 616 MB per fc1 launch at batch 64.
 *   forward: act = VMAE_ACT_GELU_Q8, aux_out = uint8 [M, ld_aux_out];   dgrad: act = VMAE_ACT_AUX_Q8, aux_in = that uint8 tensor. */
#define VMAE_ACT_GELU_Q8 2
#define VMAE_ACT_AUX_Q8 3

/* ---- library ------------------------------------------------------------------------------------------------- */
const char* vmae_last_error(void);
int vmae_abi_version(void);
int vmae_device_sm_count(void);

/* ---- K1: GEMM on tcgen05/TMEM, TMA-fed, fused epilogue ---------------------------------------------------------
 * out[M,N] = epi( A[M,K] * B[N,K]^T ), bf16 operands, fp32 accumulation.
 *   epi(v) = ( act(v + bias[n]) * aux_in[m,n] ) + residual[m,n]            (each term optional)
 *   aux_out (with act=GELU) receives gelu'(v + bias) as bf16 — the factor backward needs; the dgrad of the next
 *   layer passes it back as aux_in, so its epilogue is a plain multiply.
 *   accumulate=1 (fp32 out): out += epi(...) with red.add — used for split-K weight gradients into .grad.
 * Replaces nn.Linear fwd/bwd: tiny_vit.py:79-81,93-94,105 ; mae_vit_adapter.py:24,40-53 ; tinyvit_mae.py:15,26.
 * forward  : A = x (MAJOR_K),  B = W [N,K] (MAJOR_K)
 * dgrad    : A = dy (MAJOR_K), B = W stored [N'=Kc, K'=N] (MAJOR_MN)
 * wgrad    : A = dy stored [tokens=Kc, M] (MAJOR_MN), B = x stored [tokens=Kc, N] (MAJOR_MN), accumulate=1, k_splits>=1
 */
int vmae_gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major, int M, int N,
                   int K, const float* bias, const float* residual, long long ld_res, const void* aux_in,
                   long long ld_aux_in, void* aux_out, long long ld_aux_out, void* out, long long ld_out,
                   int out_dtype, int accumulate, int act, int k_splits, void* stream);

/* Forward GEMM with the feature perturbation z' = z + N(0, sigma^2 I) of privacy/feature_noise.py:4-7 (call site run_privacy.py:299)
 * fused into its epilogue: out[M,N] (bf16) = A[M,K] B[N,K]^T + bias + sigma * eps. eps comes from Philox4x32-10 keyed by
 * (seed, offset + (m*N + n) / 4), lane (m*N + n) % 4 — the same stream vmae_add_gaussian_noise draws for the flattened [M,N]
 * tensor. eps_out (optional, fp32 dense [M,N]) receives the normals, so parity is exact given eps. */
int vmae_gemm_bf16_noise(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const float* bias,
                         void* out, long long ld_out, float sigma, unsigned long long seed, unsigned long long offset,
                         float* eps_out, void* stream);

/* Weight + bias gradient of y = x W^T + b in one launch (split-K, fp32 accumulation into existing buffers):
 *   dw[n_out, k_in] += dy[tokens, n_out]^T x[tokens, k_in];   dbias[n_out] += column sums of dy (optional, may be NULL)
 * The bias gradient costs one extra N=16 UMMA per K step against an all-ones tile (no second pass over dy). */
int vmae_linear_wgrad(const void* dy, long long ld_dy, const void* x, long long ld_x, long long tokens, int n_out,
                      int k_in, float* dw, long long ld_dw, float* dbias, int k_splits, void* stream);

/* ---- K2: flash attention on tcgen05/TMEM (non-causal, no mask; optional attention-probability dropout) -----------
 * qkv: bf16 [B, L, 3, H, hd] (the packed projection output: q|k|v blocks, then head, then dim) — the layout of
 * Attention.qkv (tiny_vit.py:93,100) and of nn.MultiheadAttention.in_proj (mae_vit_adapter.py:40-48).
 * out: bf16 [B, L, H*hd]; lse: fp32 [B, H, L] (natural-log sum-exp of the scaled scores, saved for backward).
 * Replaces F.scaled_dot_product_attention (tiny_vit.py:103) and the MHA core of nn.TransformerEncoderLayer.
 * p_drop > 0 drops attention probabilities in the kernel (nn.MultiheadAttention(dropout=p) inside the decoder layers,
 * mae_vit_adapter.py:40-47): element (b, h, q, k) is kept with probability 1 - round(p*65536)/65536 according to a
 * stateless hash of (seed, b, h, q, k); kept probabilities are scaled by 1/(1-p), the softmax normaliser is not. The
 * backward call must receive the same (p_drop, seed). Statistical — not bit-wise — equivalence with torch's Philox stream.
 */
int vmae_attn_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int hd, float scale, float p_drop,
                  unsigned long long seed, void* stream);
/* dqkv: bf16 [B, L, 3, H, hd]; delta: fp32 workspace [B, H, L] (rowsum(dO*O), written by the call).
 * dq_accum: optional fp32 workspace [B, L, H*hd]. When given (and hd == 64) the single-pass backward runs: one
 * kernel computes dK, dV and the per-block partial dQ (5 GEMM units, one exponential per score) and reduces dQ into this
 * buffer through L2 (TMA reduce-add; the summation order of dQ is then not deterministic); NULL selects the two-kernel form. */
int vmae_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, float* dq_accum,
                  void* dqkv, int B, int L, int H, int hd, float scale, float p_drop, unsigned long long seed, void* stream);
/* The keep mask the two calls above use, materialised as uint8 [B, H, L, L] (1 = kept). Test/debug helper. */
int vmae_attn_dropout_mask(void* mask, int B, int L, int H, float p_drop, unsigned long long seed, void* stream);

/* ---- input pipeline tail (datasets/mae_loader.py:63-77) ------------------------------------------------------
 * frames_u8: device uint8 [B, T, H, W, 3] (decoded RGB frames); clip: device fp32 [B, 3, T, H, W] =
 * ((u / 255) - mean[c]) / std[c] per channel (ToTensor + Normalize), channels reversed when flip_channels != 0 (the loader's
 * `img[[2, 1, 0]]`), time moved behind the channel axis. mean3 / std3 are HOST arrays of 3 floats. */
int vmae_frames_to_clip(const void* frames_u8, float* clip, int B, int T, int H, int W, const float* mean3, const float* std3,
                        int flip_channels, void* stream);

/* ---- stack-2 latent losses on gathered rows (mae/losses.py:19-63, called from tinyvit_mae.py / mae/engine.py) ----------
 * pred, target: [rows, cols] bf16 or fp32. mode 0 = L2 (normalize != 0: target rows normalised with the biased variance,
 * eps inside the sqrt — :36-38), 1 = L1, 2 = cosine (eps = F.normalize's). loss: device scalar; partials: fp32 workspace
 * of vmae_row_loss_grid(rows) elements (per-block sums, added in a fixed order: deterministic). The target is detached in
 * the reference, so backward produces dpred only (dtype of pred); gout: device scalar (upstream gradient of the loss). */
int vmae_row_loss_grid(long long rows);
int vmae_row_loss_fwd(const void* pred, int pred_dtype, const void* target, int target_dtype, long long rows, int cols, int mode,
                      int normalize, float eps, float* partials, float* loss, void* stream);
int vmae_row_loss_bwd(const void* pred, int pred_dtype, const void* target, int target_dtype, long long rows, int cols, int mode,
                      int normalize, float eps, const float* gout, void* dpred, void* stream);

/* ---- element dropout (nn.Dropout inside nn.TransformerEncoderLayer: dropout / dropout1 / dropout2, mae_vit_adapter.py:40-47)
 * dst[i] = (add ? add[i] : 0) + keep(seed, i) * src[i] / (1 - p).  src: bf16 or fp32 (src_dtype); add: optional fp32 (the
 * residual stream); dst: bf16 or fp32 (dst_dtype), may alias src. The same (p_drop, seed) reproduces the same mask, which is
 * how backward re-applies it. vmae_dropout_mask writes the keep mask as uint8 [n] (test/debug helper). */
int vmae_dropout_apply(const void* src, int src_dtype, const float* add, void* dst, int dst_dtype, long long n, float p_drop,
                       unsigned long long seed, void* stream);
int vmae_dropout_mask(void* mask, long long n, float p_drop, unsigned long long seed, void* stream);

/* ---- K3: LayerNorm (nn.LayerNorm eps=1e-5: tiny_vit.py:112,115; decoder norms mae_vit_adapter.py:40-50) ---------
 * x: [rows, cols] fp32 or bf16 (x_dtype); y: bf16 or fp32 (y_dtype); mean/rstd: fp32 [rows] saved for backward. */
int vmae_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype,
                       float* mean, float* rstd, long long rows, int cols, float eps, void* stream);
/* dx_out[r,:] = LN'(dy)[r,:] + (dx_resid ? dx_resid[r,:] : 0)   (fp32); optional bf16 copy of dx_out;
 * dgamma/dbeta (fp32 [cols]) are ACCUMULATED into (red.add), matching autograd's .grad accumulation. */
int vmae_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* gamma, const float* mean,
                       const float* rstd, const float* dx_resid, float* dx_out, void* dx_out_bf16, float* dgamma,
                       float* dbeta, long long rows, int cols, void* stream);

/* ---- K4: fused patchify + per-patch normalise + masked MSE (train_ssl_mae.py:26-31,72-84; mae/losses.py:19-41) --
 * clip: fp32 [B,3,T,H,W]; tokens are (pt x ps x ps) tubelets, token order (t',h,w), element order (dt,py,px,c).
 * pred: bf16 or fp32 [B, N, D], D = pt*ps*ps*3; mask: uint8 [B,N] (1 = masked).
 * acc: fp32 [2] zero-initialised by the caller: acc[0] += sum over masked tokens of mean_e (pred-target)^2,
 *      acc[1] += number of masked tokens.   loss = acc[0] / (acc[1] + 1e-6)  (vmae_masked_mse_finalize).
 * diff (optional, bf16 [B,N,D]): pred - target on masked tokens, 0 on visible tokens (saved for backward).
 * norm_pix: 0 = raw pixels, 1 = normalise with unbiased variance (torch.var default, train_ssl_mae.py:75),
 *           2 = biased variance (mae/losses.py:37).  eps = 1e-6 inside the sqrt. */
int vmae_masked_mse_fwd(const float* clip, const void* pred, int pred_dtype, const unsigned char* mask, float* acc,
                        void* diff, int B, int T, int H, int W, int pt, int ps, int norm_pix, void* stream);
/* loss[0] = acc[0]/(acc[1]+1e-6); scale[0] = 2/(D*(acc[1]+1e-6)) */
int vmae_masked_mse_finalize(const float* acc, float* loss, float* scale, int D, void* stream);
/* dpred = diff * scale[0] * gout[0]   (bf16 out, n elements) */
int vmae_masked_mse_bwd(const void* diff, const float* scale, const float* gout, void* dpred, long long n, void* stream);

/* ---- K5/K6: tube-mask index helpers, visible-tubelet gather, decoder-input scatter -------------------------------
 * mask_frame: uint8 [B, L] (one frame of the tube mask, mae_loader.py:80-90); vis_idx: int32 [B, n_vis] ascending. */
int vmae_mask_to_indices(const unsigned char* mask_frame, int* vis_idx, int B, int L, int n_vis, void* stream);
/* im2col of the VISIBLE tubelets only: patches bf16 [B*Tp*n_vis, 3*pt*ps*ps] in Conv3d weight order (c,dt,py,px),
 * row order (b, t', i); pos_rows (optional, fp32 [rows, D]) = temporal_pos[t'] + spatial_pos[vis_idx[b,i]]. */
int vmae_gather_tubelets(const float* clip, const int* vis_idx, void* patches, const float* temporal_pos,
                         const float* spatial_pos, float* pos_rows, int B, int T, int H, int W, int pt, int ps,
                         int n_vis, int D, void* stream);
/* decoder input (mae_vit_adapter.py:97-107): x_dec[b,t,l,:] = visible ? z[row(b,t,l),:] + tpos[t] + spos[l] : mask_token
 * z: bf16 [B*Tp*n_vis, D] (vis_idx given) or bf16 [B*Tp*L, D] (vis_idx NULL: blend form); x_dec fp32 [B,Tp,L,D]. */
int vmae_decoder_input_fwd(const void* z, const int* vis_idx, const unsigned char* mask, const float* temporal_pos,
                           const float* spatial_pos, const float* mask_token, float* x_dec, int B, int Tp, int L,
                           int n_vis, int D, void* stream);
/* backward: dz (bf16, same shape as z; masked rows 0 in blend form), and ACCUMULATED fp32 dtpos[Tp,D], dspos[L,D],
 * dmask_token[D]. */
int vmae_decoder_input_bwd(const float* dx_dec, const int* vis_idx, const unsigned char* mask, void* dz, float* dtpos,
                           float* dspos, float* dmask_token, int B, int Tp, int L, int n_vis, int D, void* stream);
/* d(pos) of the encoder-side embedding: dtpos[t'] += rows, dspos[vis_idx[b,i]] += rows; dx fp32 [B*Tp*n_vis, D] */
int vmae_pos_rows_bwd(const float* dx, const int* vis_idx, float* dtpos, float* dspos, int B, int Tp, int n_vis, int D,
                      void* stream);

/* ---- K6 as ONE kernel: tubelet embedding of the visible tubes = implicit GEMM over the clip (north_star bullet 1) ----------------
 * out[m,:] (fp32 [B*(T/pt)*n_vis, E]) = W patch(m) + bias + temporal_pos[tp(m),:] + spatial_pos[l(m),:], rows in gathered order
 * (clip, temporal slot, i-th visible patch; l = vis_idx[b,i]). weight: bf16 [E, 3*pt*ps*ps] = the Conv3d weight [E,3,pt,ps,ps]
 * flattened (K order c, dt, py, px). The A operand is gathered by 5-D TMA boxes straight from clip fp32 [B,3,T,H,W]; no patch
 * matrix is materialised. ps must be 16, E a multiple of 64 (<= 384); other geometries return VMAE_ERR_UNSUPPORTED.
 * wgrad: dw[E, 3*pt*ps*ps] (fp32, ACCUMULATED) += dy^T patches, dy bf16 [rows, E]; E and K multiples of 128.
 * (bias gradient = vmae_colsum_accum(dy), positional gradients = vmae_pos_rows_bwd(dy).) */
int vmae_tubelet_embed_fwd(const float* clip, const int* vis_idx, const void* weight_bf16, const float* bias,
                           const float* temporal_pos, const float* spatial_pos, float* out, int B, int T, int H, int W, int pt,
                           int ps, int n_vis, int E, void* stream);
int vmae_tubelet_embed_wgrad(const float* clip, const int* vis_idx, const void* dy_bf16, float* dw, int B, int T, int H, int W,
                             int pt, int ps, int n_vis, int E, void* stream);

/* ---- K7: TinyViT convolution kit on channels-last bf16 feature maps (rows = pixels), tiny_vit.py:12-72 ---------------
 * Train-mode BatchNorm2d (batch statistics over the M = N*H*W rows, biased variance for normalisation, running statistics
 * updated with `momentum` and the unbiased variance, as nn.BatchNorm2d) with optional fused exact GELU and residual add:
 *   y = act((x - mean) * rstd * gamma + beta) (+ residual).  x, y, residual: bf16 [M, C]; mean/rstd: fp32 [C] (saved);
 *   workspace: fp32 [2*C]; running_mean/var may be NULL.  act: VMAE_ACT_NONE | VMAE_ACT_GELU. */
int vmae_bn_train_fwd(const void* x, const float* gamma, const float* beta, const void* residual, void* y, float* mean,
                      float* rstd, float* running_mean, float* running_var, float* workspace, long long M, int C,
                      float eps, float momentum, int act, void* stream);
/* eval-mode / recompute form: statistics supplied (mean, rstd = 1/sqrt(var + eps)) */
int vmae_bn_apply(const void* x, const float* gamma, const float* beta, const float* mean, const float* rstd,
                  const void* residual, void* y, long long M, int C, int act, void* stream);
/* backward of the train-mode form: dx bf16 [M, C]; dgamma/dbeta fp32 [C] are ACCUMULATED; workspace fp32 [2*C]. */
int vmae_bn_train_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, void* dx, float* dgamma, float* dbeta, float* workspace, long long M, int C,
                      int act, void* stream);

/* depthwise 3x3 convolution, pad 1, stride 1|2, NHWC bf16 (tiny_vit.py:46): x [N,H,W,C] -> y [N,Ho,Wo,C];
 * w: fp32 [C, 9] (nn.Conv2d weight [C,1,3,3] as stored). bwd: dx (bf16, may be NULL) and dw (fp32 [C,9], ACCUMULATED, may be NULL). */
int vmae_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride, void* stream);
int vmae_dwconv3x3_bwd(const void* x, const void* dy, const float* w, void* dx, float* dw, int N, int H, int W, int C,
                       int stride, void* stream);
/* squeeze-excite (tiny_vit.py:20-34). pool: out[n,c] = scale * sum_p x[n,p,c] (dy == NULL) or scale * sum_p x*dy (gate gradient);
 * scale: y[n,p,c] = a[n,p,c] * gate[n,c] (+ add_scale * add[n,c]); a == NULL broadcasts add_scale * add only.
 * x/a/y/dy bf16 [N,HW,C]; gate/add/out fp32 [N,C]. */
int vmae_se_pool(const void* x, const void* dy, float* out, int N, int HW, int C, float scale, void* stream);
int vmae_se_scale(const void* a, const float* gate, const float* add, void* y, int N, int HW, int C, float add_scale, void* stream);
/* im2col / col2im of 3x3 (pad 1) windows for the stem convolutions (tiny_vit.py:66-70); the GEMM is vmae_gemm_bf16.
 * NHWC bf16 -> cols [N*Ho*Wo, 9*C] in (kh,kw,c) order; nchw_f32: fp32 NCHW image -> cols [N*Ho*Wo, KP] in (c,kh,kw) order, zero padded. */
int vmae_im2col3x3(const void* x, void* col, int N, int H, int W, int C, int stride, void* stream);
int vmae_col2im3x3(const void* dcol, void* dx, int N, int H, int W, int C, int stride, void* stream);
int vmae_im2col3x3_nchw_f32(const float* x, void* col, int N, int Cin, int H, int W, int stride, int KP, void* stream);

/* ---- small memory-bound helpers -------------------------------------------------------------------------------- */
int vmae_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream);
int vmae_cast_bf16_to_f32(const void* x, float* y, long long n, void* stream);
/* fp32 master weights -> bf16 operand copies for a whole parameter set in one launch. `table` is a DEVICE array of
 * `count` entries {const float* src; bf16* dst; long long n;} (24 bytes each; src and dst 16-byte aligned). */
int vmae_multi_cast_f32_to_bf16(const void* table, int count, void* stream);
/* a18: AdamW with decoupled weight decay over many fp32 tensors in one launch (torch.optim.AdamW semantics,
 * train_ssl_mae.py:163). table (device): `chunks` entries {float* p; const float* g; float* m; float* v; int64 n}, each at most
 * vmae_adamw_chunk() elements of one tensor. hp (device, 6 floats): lr, beta1, beta2, eps, weight_decay, step; the call first
 * increments step, then updates p, m, v in place with bias corrections 1 - beta^step. Nothing is read from the host at run
 * time, so a captured call replays correctly after the host has changed hp[0] (learning-rate schedule). */
int vmae_adamw_multi(const void* table, int chunks, float* hp, void* stream);
int vmae_adamw_chunk(void);
/* out = dy * dact (dact = the gelu' tensor saved by the GELU epilogue); bf16, n elements (n % 4 == 0) */
int vmae_gelu_bwd(const void* dy, const void* dact, void* out, long long n, void* stream);
/* db[n] += sum_m dy[m,n]   (bias gradient; dy bf16 or fp32 [M, ld]) */
int vmae_colsum_accum(const void* dy, int dy_dtype, long long ld, float* db, long long M, int N, void* stream);
/* K8: z' = z + sigma * N(0,1) (privacy/feature_noise.py:4-7). Philox4x32-10 keyed by (seed, offset + element/4).
 * eps_out (optional fp32) receives the unit normals that were drawn. In-place allowed (out == z). */
int vmae_add_gaussian_noise(const void* z, void* out, int dtype, float* eps_out, long long n, float sigma,
                            unsigned long long seed, unsigned long long offset, void* stream);
/* K9 helper (FedAvg pre-scale, federated/fed_loop.py:45-49): y = alpha * x (+ y if accumulate), fp32 */
int vmae_scale_f32(const float* x, float* y, float alpha, int accumulate, long long n, void* stream);

/* ---- stack 2: latent-space MAE mask / gather logic (tinyvit_mae.py:104-114) -------------------------------------------
 * vmae_mask_row_indices: mask uint8 [B, N] -> idx int32 [B * n_sel]: ascending GLOBAL row numbers b*N + n of the entries with
 *   (mask != 0) == want, n_sel per clip; this is the row order of torch's boolean gather `rec[token_mask]` (:113-114) when every
 *   clip holds n_sel hits (both modes of make_token_mask do). flag (optional device int, zeroed by the caller): bit 0 = some clip
 *   had fewer hits (padded with -1), bit 1 = some clip had more. No host sync.
 * vmae_mask_token_blend_fwd: out[r,:] = mask[r] ? mask_token[:] : tok[r,:]   (torch.where, :107-108); tok/out bf16 or fp32.
 * vmae_mask_token_blend_bwd: dtok[r,:] = mask[r] ? 0 : dout[r,:] (dtok may be NULL); dmask_token[:] += sum of masked rows (may be NULL).
 * vmae_gather_rows: out[i,:] = src[idx[i],:] (idx < 0 -> zero row).  vmae_scatter_rows: dst = 0; dst[idx[i],:] = src[i,:]
 *   (the backward of the gather; idx unique). */
int vmae_mask_row_indices(const unsigned char* mask, int* idx, int* flag, int B, int N, int n_sel, int want, void* stream);
int vmae_mask_token_blend_fwd(const void* tok, int dtype, const unsigned char* mask, const float* mask_token, void* out,
                              long long rows, int D, void* stream);
int vmae_mask_token_blend_bwd(const void* dout, int dtype, const unsigned char* mask, void* dtok, float* dmask_token,
                              long long rows, int D, void* stream);
int vmae_gather_rows(const void* src, int dtype, const int* idx, void* out, long long n_out, int D, void* stream);
int vmae_scatter_rows(const void* src, int dtype, const int* idx, void* dst, long long n_src, long long n_dst_rows, int D,
                      void* stream);

/* F.adaptive_avg_pool2d(f, (s, s)) of tinyvit_backbone.py:74-81 on a channels-last bf16 map x [N,H,W,C] -> out fp32 [N, s*s, C];
 * bin i covers [floor(i*H/s), ceil((i+1)*H/s)). bwd: dx bf16 [N,H,W,C] from dout fp32 [N, s*s, C]. */
int vmae_adaptive_avgpool_fwd(const void* x, float* out, int N, int H, int W, int C, int s, void* stream);
int vmae_adaptive_avgpool_bwd(const float* dout, void* dx, int N, int H, int W, int C, int s, void* stream);

/* Video embedding + feature noise (train_finetune.py:19-40 pooled feature; feature_noise.py:4-7 at run_privacy.py:299):
 * out[b,:] (fp32 [B,D]) = mean over the n token rows of x[b] (bf16 [B,n,D]) + sigma * eps; Philox stream of the flattened [B,D]
 * output (as vmae_add_gaussian_noise); eps_out optional (fp32 [B,D]); sigma == 0: plain mean. D % 64 == 0. */
int vmae_meanpool_noise(const void* x, float* out, float* eps_out, int B, int n, int D, float sigma, unsigned long long seed,
                        unsigned long long offset, void* stream);

/* ---- K10: fused softmax cross-entropy of the finetune head (train_finetune.py:86-107, nn.CrossEntropyLoss, mean reduction) ----
 * fwd: logits bf16|fp32 [rows, C], labels int64 [rows] -> lse fp32 [rows] (saved), row_loss fp32 [rows], loss fp32 scalar = mean.
 * bwd: dlogits (dtype of logits) = (softmax(logits) - onehot(label)) * gout[0] / rows. */
int vmae_softmax_ce_fwd(const void* logits, int dtype, const long long* labels, float* lse, float* row_loss, float* loss,
                        long long rows, int C, void* stream);
int vmae_softmax_ce_bwd(const void* logits, int dtype, const long long* labels, const float* lse, const float* gout,
                        void* dlogits, long long rows, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMAE_H_ */
