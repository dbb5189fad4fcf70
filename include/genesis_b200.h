/* genesis_b200.h -- C ABI of libgenesis_b200.so (sm_100a).
 *
 * The reference (applied-ai-lab/genesis) has no native layer: every operator of its hot path is a
 * PyTorch nn.Module call that ends in cuDNN / cuBLAS / ATen.  This library is what a replacement of
 * those calls binds to.  Each entry point names the reference call sites it replaces (paths relative
 * to the reference checkout).  Conventions:
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller, including
 *     workspaces and tensors saved for the backward pass; nothing is allocated or freed inside;
 *   - all work is enqueued on `stream`; no host synchronisation; no global mutable state, so every
 *     call is re-entrant per stream and CUDA-graph capturable;
 *   - returns 0 on success, a negative G2_ERR_* code for a rejected argument, or a positive
 *     cudaError_t from the launch;
 *   - image activations are NHWC fp32 ([N, H*W, C] row-major) unless stated otherwise; tensors that
 *     cross the model boundary (x, recon, x_r_k, log_m_k) are the reference's NCHW.
 */
#ifndef GENESIS_B200_H_
#define GENESIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* g2_stream_t; /* == cudaStream_t */

#define G2_OK 0
#define G2_ERR_ARG (-1)
#define G2_ERR_UNSUPPORTED (-2)

/* activation / epilogue codes */
#define G2_ACT_NONE 0
#define G2_ACT_RELU 1
#define G2_ACT_ELU 2
#define G2_ACT_MUL_RELU_GRAD 3 /* out = acc * relu'(pre), aux = saved post-activation */
#define G2_ACT_MUL_ELU_GRAD 4
#define G2_ACT_SIGMOID 5
/* normalisation modes and post-ops */
#define G2_NORM_NONE 0
#define G2_NORM_BATCH 1
#define G2_NORM_INSTANCE 2
#define G2_NORM_GROUP 3
#define G2_POST_GATE 0
#define G2_POST_RELU 1
#define G2_POST_NONE 2 /* normalise only (LayerNorm of models/genesisv2_config.py:83) */

int g2_abi_version(void);

/* ---- convolutions, exact fp32 SIMT implicit GEMM (igemm_simt.cu) -----------------------------------
 * Replaces F.conv2d / F.conv_transpose2d and their autograd backward for every conv of the path:
 * third_party/sylvester/layers.py:19-20,43,65-67,90 (gated 5x5 / 16x16 convs and conv-transposes),
 * modules/encoders.py:31-34 (3x3 s2), modules/decoders.py:27-31 (3x3 VALID, 1x1),
 * modules/blocks.py:151-165 (3x3 p1), models/genesisv2_config.py:89-99 (conv-transpose 5x5 s2).
 * mode 0: out[n,oh,ow,co] = b[co] + sum in[n,oh*S+r-P,ow*S+s-P,c] w[r,s,c,co]   (conv fwd, convT dgrad)
 * mode 1: out[n,h,w,co] = b[co] + sum_{(h+P-r)%S==0,(w+P-s)%S==0} in[n,(h+P-r)/S,(w+P-s)/S,c] w[r,s,c,co]
 *                                                                                (convT fwd, conv dgrad)
 * w packed [R,S,Ci,Co] (wT=0) or [R,S,Co,Ci] (wT=1); bias/aux may be NULL; act = G2_ACT_*. */
int g2_conv_igemm_f32(const float* in, const float* w, const float* bias, const float* aux, float* out,
                      int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride,
                      int pad, int mode, int wT, int act, g2_stream_t stream);
/* dW[r,s,a,b] = sum_{n,oh,ow} g[n,oh*S+r-P,ow*S+s-P,a] * t[n,oh,ow,b]; dw layout [R,S,Cg,Ct] (outT=0) or
 * [R,S,Ct,Cg] (outT=1); dw is overwritten. */
int g2_conv_wgrad_f32(const float* g, const float* t, float* dw, int N, int Hg, int Wg, int Cg, int Ht,
                      int Wt, int Ct, int R, int S, int stride, int pad, int outT, g2_stream_t stream);
/* C[M,N] (+)= op(A)[M,K] op(B)[K,N] + bias[N], row-major.  Replaces nn.Linear / nn.LSTM matmuls
 * (third_party/sylvester/VAE.py:100-104; modules/attention.py:81-82,94-97; modules/encoders.py:35-37;
 * modules/unet.py:60-65; models/genesis_config.py:131-138; models/genesisv2_config.py:82-86,104-105)
 * and the full-map gated convs VAE.py:23,29 viewed as GEMMs. */
int g2_gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda,
                int ldb, int ldc, int transA, int transB, int act, int accumulate, g2_stream_t stream);
/* out[c] (+)= sum_m x[m,c]   (bias gradients) */
int g2_colsum_f32(const float* x, float* out, long M, int C, int accumulate, g2_stream_t stream);

/* dw[4][Cin] = sum_pix d4[pix][:]^T h[pix][:] -- weight gradient of the 1x1 output heads (out1x1). */
int g2_head_wgrad_f32(const float* h, const float* d4, float* dw, long NP, int Cin, g2_stream_t stream);

/* ---- normalisation + gate / ReLU (norm.cu) ---------------------------------------------------------
 * Replaces nn.BatchNorm2d / nn.InstanceNorm2d / nn.GroupNorm + sigmoid gate or ReLU and their backward:
 * third_party/sylvester/layers.py:22-54,69-101; modules/blocks.py:151-165;
 * models/genesisv2_config.py:92-98.  y is [N,HW,Cy]; for POST_GATE Cy = 2C (h | g), out is [N,HW,C]. */
int g2_norm_stats_f32(const float* y, double* sums /*[N,Cy,2]*/, int N, int HW, int Cy, g2_stream_t stream);
int g2_norm_finalize_f32(const double* sums, const float* g0, const float* b0, const float* g1, const float* b1,
                         float* rm0, float* rv0, float* rm1, float* rv1, float* mean, float* rstd, float* scale,
                         float* shift, int N, int HW, int Cy, int half, int mode, int groups, int training,
                         float eps, float momentum, g2_stream_t stream);
int g2_norm_apply_f32(const float* y, const float* scale, const float* shift, float* out, int N, int HW, int C,
                      int sn, int post, g2_stream_t stream);
int g2_norm_bwd_stats_f32(const float* y, const float* dout, const float* scale, const float* shift,
                          const float* mean, const float* rstd, double* sums2 /*[N,Cy,2]*/, int N, int HW, int C,
                          int sn, int post, g2_stream_t stream);
int g2_norm_bwd_finalize_f32(const double* sums2, const float* g0, const float* g1, float* m1, float* m2,
                             float* dg0, float* db0, float* dg1, float* db1, int N, int HW, int Cy, int half,
                             int mode, int groups, g2_stream_t stream);
/* the same with `accumulate`: dg0 / db0 / dg1 / db1 are added to (direct-gradient mode: they point into param.grad) */
int g2_norm_bwd_finalize_acc_f32(const double* sums2, const float* g0, const float* g1, float* m1, float* m2,
                                 float* dg0, float* db0, float* dg1, float* db1, int N, int HW, int Cy, int half,
                                 int mode, int groups, int accumulate, g2_stream_t stream);
int g2_norm_bwd_apply_f32(const float* y, const float* dout, const float* scale, const float* shift,
                          const float* mean, const float* rstd, const float* m1, const float* m2, float* dy,
                          int N, int HW, int C, int sn, int post, g2_stream_t stream);
/* Same, and dbias[c] += column sums of dy over all N*HW rows (Cy channels): the bias gradient of the convolution that produced
 * y, fused so that no separate pass re-reads dy.  dbias is accumulated into (caller-initialised). */
int g2_norm_bwd_apply_bias_f32(const float* y, const float* dout, const float* scale, const float* shift,
                               const float* mean, const float* rstd, const float* m1, const float* m2, float* dy,
                               float* dbias, int N, int HW, int C, int sn, int post, g2_stream_t stream);

/* ---- layout, scans, packing, reductions (pointwise.cu) --------------------------------------------- */
/* x [N,C,P] <-> y [N,P,C] */
int g2_layout_f32(const float* x, float* y, long N, int C, int P, int to_nchw, g2_stream_t stream);
/* stick-breaking scan, modules/attention.py:40-50,114-130 + models/genesis_config.py:169-171 */
int g2_sbp_scan_fwd_f32(const float* logits, float* log_m, float* log_s, long BP, int K, int nl, g2_stream_t stream);
int g2_sbp_scan_bwd_f32(const float* logits, const float* dlog_m, float* dlogits, long BP, int K, int nl,
                        g2_stream_t stream);
/* component-VAE encoder input: repeat(x) (+) cat(log_m), modules/component_vae.py:58-63 */
int g2_comp_pack_f32(const float* x, const float* log_m, float* out, int K, int B, int P, int Cp, g2_stream_t stream);
/* x [N,C,P] (NCHW) -> y [N,P,Cp] (NHWC, zero channels C..Cp-1): lets the 3-channel image feed 32-channel k-blocks */
int g2_nhwc_pad_f32(const float* x, float* y, long N, int C, int P, int Cp, g2_stream_t stream);
/* out[n,p,c] = act(a[n,c] + m[p,c]): first broadcast-decoder layer without materialising the broadcast,
 * modules/blocks.py:104-130 + modules/decoders.py:25-26 */
int g2_bcast_add_act_f32(const float* a, const float* m, float* out, long N, int P, int C, int act,
                         g2_stream_t stream);
int g2_act_bwd_f32(const float* dout, const float* out, float* dpre, long total, int act, g2_stream_t stream);
/* The same fused with the bias gradient of a conv + bias + activation layer: dbias[c] += column sums of dpre viewed as
 * [M rows, C] (accumulated into a caller-initialised buffer). */
int g2_act_bwd_bias_f32(const float* dout, const float* out, float* dpre, float* dbias, long M, int C, int act,
                        g2_stream_t stream);
int g2_seg_colsum_f32(const float* x, float* out, int N, int P, int C, g2_stream_t stream);
int g2_sum_dim0_f32(const float* x, float* out, int N, long J, g2_stream_t stream);

/* ---- GENESIS-V2 / MONet specific (v2.cu) -----------------------------------------------------------
 * nearest x0.5 / x2 resampling of modules/unet.py:77-78,88-89 (modes 0,1) and their adjoints (modes 2,3);
 * Ho, Wo are the OUTPUT sizes. */
int g2_resample_f32(const float* x, float* y, long N, int Ho, int Wo, int C, int mode, g2_stream_t stream);
/* InstanceColouringSBP.forward, gaussian kernel, modules/attention.py:177-223: one CTA per image loops the K-1
 * steps (argmax seed -> distance -> exp -> clamp -> log scan).  colour [B,P,8] NHWC, u [B,P]. */
int g2_icsbp_fwd_f32(const float* colour, const float* u, const float* log_sigma, float* log_m, float* log_s,
                     int* seed_idx, int B, int P, int K, int colour_dim, g2_stream_t stream);
int g2_icsbp_bwd_f32(const float* colour, const float* log_sigma, const int* seed_idx, const float* dlog_m,
                     float* dcolour, float* dlog_sigma_b, int B, int P, int K, int colour_dim, g2_stream_t stream);
/* The same with the `kernel` option of InstanceColouringSBP (modules/attention.py:146-153, 195-203):
 * kernel_type 0 gaussian, 1 laplacian (euclidian distance, blocks.py:49-61), 2 epanechnikov (relu(1 - d2 / sigma)). */
int g2_icsbp_kernel_fwd_f32(const float* colour, const float* u, const float* log_sigma, float* log_m, float* log_s,
                            int* seed_idx, int B, int P, int K, int colour_dim, int kernel_type, g2_stream_t stream);
int g2_icsbp_kernel_bwd_f32(const float* colour, const float* log_sigma, const int* seed_idx, const float* dlog_m,
                            float* dcolour, float* dlog_sigma_b, int B, int P, int K, int colour_dim, int kernel_type,
                            g2_stream_t stream);
/* dynamic_K (models/genesisv2_config.py:118-137, modules/attention.py:168-169, 218-219): a step whose mask would hold fewer than
 * 20 pixels ends that image's loop and the current scope becomes its last mask.  n_masks [B] int32 = masks per image;
 * log_m slots k >= n_masks[b] are -1e10 (the reference's batch padding), seed_idx beyond the last step is -1. */
int g2_icsbp_dynamic_fwd_f32(const float* colour, const float* u, const float* log_sigma, float* log_m, float* log_s,
                             int* seed_idx, int* n_masks, int B, int P, int K, int colour_dim, int kernel_type,
                             g2_stream_t stream);
int g2_icsbp_dynamic_bwd_f32(const float* colour, const float* log_sigma, const int* seed_idx, const int* n_masks,
                             const float* dlog_m, float* dcolour, float* dlog_sigma_b, int B, int P, int K, int colour_dim,
                             int kernel_type, g2_stream_t stream);
/* masked feature pooling of models/genesisv2_config.py:147-152: num[k,b,c] = sum_p m_k f, msum[k,b] = sum_p m_k */
int g2_masked_pool_fwd_f32(const float* f, const float* log_m, float* num, float* msum, int B, int P, int C, int K,
                           g2_stream_t stream);
int g2_masked_pool_bwd_f32(const float* f, const float* log_m, const float* dnum, const float* dmsum, float* df,
                           float* dlog_m, int B, int P, int C, int K, g2_stream_t stream);

/* ---- decoder head + mixture likelihood (loss.cu) ---------------------------------------------------
 * out1x1: final 1x1 conv (+ sigmoid on the first nsig channels) writing NCHW planes:
 * modules/decoders.py:31, modules/component_vae.py:89-93, third_party/sylvester/VAE.py:121,
 * modules/unet.py:66, models/genesisv2_config.py:99. */
int g2_out1x1_fwd_f32(const float* h, const float* w, const float* bias, float* out, long N, int P, int Cin,
                      int nout, int nsig, g2_stream_t stream);
int g2_out1x1_bwd_f32(const float* dout, const float* out, const float* w, float* dh, float* dpre4, long N,
                      int P, int Cin, int nout, int nsig, g2_stream_t stream);
/* Genesis.x_loss + recon (+ log-softmax of mask logits): models/genesis_config.py:188-190,273-286,
 * models/monet_config.py:136-140, models/genesisv2_config.py:213-223. */
int g2_mixture_fwd_f32(const float* x, const float* xr, const float* lm, const float* stdv, float* err,
                       float* recon, float* lse, float* lm_out, int K, int B, int P, int softmax, int xr_cs,
                       int lm_cs, g2_stream_t stream);
int g2_mixture_bwd_f32(const float* x, const float* xr, const float* lm, const float* stdv, const float* lse,
                       const float* gerr, float* dxr, float* dlm, int K, int B, int P, int softmax, int xr_cs,
                       int lm_cs, int dlm_cs, g2_stream_t stream);
/* Backward of the pixel-wise form (Genesis.x_loss(pixel_wise=True), models/genesis_config.py:283-284): the loss is -lse
 * [B,3,P] and gpix [B,3,P] its upstream gradient; masks are given (no softmax). */
int g2_mixture_bwd_pix_f32(const float* x, const float* xr, const float* lm, const float* stdv, const float* lse,
                           const float* gpix, float* dxr, float* dlm, int K, int B, int P, int xr_cs, int lm_cs, int dlm_cs,
                           g2_stream_t stream);
/* xr_cs / lm_cs / dlm_cs: channels per (slot,image) in the xr / lm / dlm tensors -- 3,1,1 for separate tensors, 4,4,4
 * when x_r and the mask logit are the 4 planes of one decoder output [K,B,4,P] (lm = dec + 3P). */

/* MONet.kl_m_loss fused with the log-softmax of the reconstructed mask logits: models/monet_config.py:136-140,157-170.
 * lm [K,B,lm_cs,P] log masks (plane 0), logits [K,B,lg_cs,P] (plane 0 of the pointer given) -> lmr [K,B,P], kl [B].
 * bwd: dlm (+= when accumulate_dlm) and dlogits get d(sum_b gkl_b kl_b). */
int g2_mask_kl_fwd_f32(const float* lm, const float* logits, float* lmr, float* kl, int K, int B, int P, int lm_cs,
                       int lg_cs, g2_stream_t stream);
int g2_mask_kl_bwd_f32(const float* lm, const float* logits, const float* gkl, float* dlm, float* dlogits, int K, int B,
                       int P, int lm_cs, int lg_cs, int dlm_cs, int dlg_cs, int accumulate_dlm, g2_stream_t stream);

/* ---- TF32 tensor-core implicit GEMM: tcgen05.mma + TMEM + TMA (igemm_tc.cu) -------------------------
 * Same call sites as g2_conv_igemm_f32 / g2_gemm_f32, for the shapes that fit the 128 x {32,64,128} UMMA
 * tiles (Ci % 32 == 0, Co in {32,64,128} or a multiple of 64, power-of-two widths or VALID convs).
 * Weights packed [R*S][Co][Ci].  g2_conv_tf32_supported returns 1/0 (it is a query, not an error code). */
int g2_conv_tf32_supported(int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride,
                           int pad, int mode);
int g2_conv_igemm_tf32(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi,
                       int Ci, int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act,
                       g2_stream_t stream);
/* Weight gradient on the tensor cores (wgrad_tc.cu; MN-major operands): same contract as g2_conv_wgrad_f32 plus
 * a caller-owned workspace of g2_conv_wgrad_tf32_workspace(...) bytes (0 = shape not supported). */
long g2_conv_wgrad_tf32_workspace(int N, int Hg, int Wg, int Cg, int Ht, int Wt, int Ct, int R, int S, int stride);
/* Host-only query: plan of the experimental halo weight-gradient kernel (G2_WGRAD_HALO=1; csrc/wgrad_tc.cu, namespace wgh) for
 * the CPU replay test; out needs 16 + 7 * 12 ints.  Returns the number of ints written (out[0] = 0: shape not covered). */
int g2_conv_wgrad_halo_plan(int N, int Hg, int Wg, int Cg, int Ht, int Wt, int Ct, int R, int S, int stride, int* out);
int g2_conv_wgrad_tf32(const float* g, const float* t, float* dw, float* ws, int N, int Hg, int Wg, int Cg, int Ht,
                       int Wt, int Ct, int R, int S, int stride, int pad, int outT, g2_stream_t stream);
/* C[M,N] = A[M,K] W[N,K]^T + bias[N] */
/* As g2_conv_wgrad_tf32, but dw is addressed as dw[tap*s_tap + a*s_a + b*s_b] for a < a_lim (channel of g), b < b_lim
 * (channel of t) -- e.g. the torch-layout .grad of the parameter (Conv2d [Co,Ci,R,S], a = Ci, b = Co: strides
 * (1, R*S, Ci*R*S)) -- and is accumulated into when accumulate != 0.  Replaces the permute + AccumulateGrad add. */
int g2_conv_wgrad_tf32_to(const float* g, const float* t, float* dw, float* ws, int N, int Hg, int Wg, int Cg, int Ht,
                          int Wt, int Ct, int R, int S, int stride, int pad, long s_tap, long s_a, long s_b, int a_lim,
                          int b_lim, int accumulate, g2_stream_t stream);
/* Both tensor-core operand packs of a torch-layout conv weight in one pass: packA[tap][co][ci_pad] (forward) and
 * packB[tap][ci_pad][co] (data gradient; may be NULL); input channels >= Ci are zero.  transposed: 0 = Conv2d
 * [Co,Ci,R,S], 1 = ConvTranspose2d [Ci,Co,R,S]. */
int g2_pack_conv_weight_f32(const float* w, float* packA, float* packB, int Co, int Ci, int Ci_pad, int RS,
                            int transposed, g2_stream_t stream);
int g2_gemm_tf32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K,
                 g2_stream_t stream);
/* Same product with a caller-owned workspace of g2_gemm_tf32_workspace(M, N, K) bytes (0 = no split needed; ws may then be
 * NULL) and a fused activation (G2_ACT_NONE / RELU / ELU).  Long reductions with few output tiles are split along K: every
 * split writes its partial product to its own slab of `ws` and an ordered reduction adds the slabs, bias and activation --
 * no float atomics, bitwise reproducible.  g2_gemm_tf32 is the unsplit form. */
long g2_gemm_tf32_workspace(int M, int N, int K);
int g2_gemm_tf32_ws(const float* A, const float* W, const float* bias, float* C, float* ws, int M, int N, int K, int act,
                    g2_stream_t stream);

/* Halo variant of g2_conv_igemm_tf32 (igemm_halo.cu): stride-1 problems and the sub-pixel classes of stride-2
 * conv-transposes with the zero-padded activation window loaded once per CTA and all filter taps issued from it
 * through shifted UMMA descriptors.  g2_conv_igemm_tf32 routes to it whenever g2_conv_halo_supported() == 1;
 * g2_conv_halo_enable(0/1) switches the routing at run time (A/B measurements) and returns the previous setting;
 * g2_conv_halo_plan fills plan[96] with the launch plan of sub-pixel class `cls` (host-only query, no launch):
 * {nclasses, TH, TNB, RH, chunk_rows, chunks, m_tiles, a_bytes, tiles_h, Wp, dh_min, dw_min, os, ph, pw, Hv, Wv,
 *  ntaps, BN, smem_bytes}, plan[32+i] = flat row offset of tap i, plan[64+i] = its weight index. */
/* 3xTF32 form of g2_conv_halo_tf32 (same problems, fp32-level accuracy on the tensor cores): `w` is the pack [R*S][Co][2*Ci]
 * with channel halves w_hi (the weights as exact TF32 values) | w_lo = w - w_hi; `in` is the plain fp32 activation.  Per channel
 * block the kernel issues (x_hi, w_hi) and (x_hi, w_lo) from the raw window (the tensor core truncates fp32 operand bits to TF32),
 * rewrites the window in place to x_lo = x - trunc(x) and issues (x_lo, w_hi): no split pass over HBM, no extra shared memory. */
int g2_conv_halo_x3_tf32(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi, int Ci,
                         int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act, g2_stream_t stream);
int g2_conv_halo_enable(int on);
/* debug: per-CTA clock64() phase timestamps of subsequent halo launches into a device buffer [CTAs][64] (NULL = off) */
int g2_conv_halo_debug(int64_t* buf);
int g2_conv_halo_supported(int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride,
                           int pad, int mode);
int g2_conv_halo_tf32(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi,
                      int Ci, int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act,
                      g2_stream_t stream);
int g2_conv_halo_plan(int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride, int pad,
                      int mode, int cls, int* plan);

/* ---- latent path, fused elementwise kernels (latent.cu) ---------------------------------------------
 * LSTM cell after the two gate GEMMs (gates = gx + gh, [B,4H], order i,f,g,o; c_prev NULL = zero state): replaces the
 * sigmoid/tanh/mul/add chain of one torch.nn.LSTM step (modules/attention.py:94-97, models/genesis_config.py:301-305).
 * bwd: dgates [B,4H] is the gradient of both gx and gh; dh / dc / dc_prev may be NULL. */
int g2_lstm_cell_fwd_f32(const float* gx, const float* gh, const float* c_prev, float* h, float* c, int B, int H,
                         g2_stream_t stream);
int g2_lstm_cell_bwd_f32(const float* gx, const float* gh, const float* c_prev, const float* c, const float* dh,
                         const float* dc, float* dgates, float* dc_prev, int B, int H, g2_stream_t stream);
/* Gaussian head on the un-chunked Linear output lo [B,2D] = (mu | raw): sigma = softplus(raw + 0.5) + 1e-8
 * (modules/blocks.py:22-23), z = mu + sigma * eps (rsample with caller-supplied noise: component_vae.py:67-69,
 * attention.py:98-103); mu is also written contiguously.  bwd: dlo = (dz + dmu | (dz*eps + dsigma) * sigmoid(raw + 0.5));
 * dz / dmu / dsigma may be NULL. */
int g2_gauss_head_fwd_f32(const float* lo, const float* eps, float* z, float* mu, float* sigma, int B, int D,
                          g2_stream_t stream);
int g2_gauss_head_bwd_f32(const float* lo, const float* eps, const float* dz, const float* dmu, const float* dsigma,
                          float* dlo, int B, int D, g2_stream_t stream);
/* Prior head on lo [B,2D] = (a | b): pmu = tanh(a) (use_tanh = 0: pmu = a, Genesis.sample genesis_config.py:359),
 * psigma = sigmoid(b + 4) + 1e-4 (modules/blocks.py:28-34). */
int g2_prior_head_fwd_f32(const float* lo, float* pmu, float* psigma, int B, int D, int use_tanh, g2_stream_t stream);
int g2_prior_head_bwd_f32(const float* pmu, const float* psigma, const float* dpmu, const float* dpsigma, float* dlo,
                          int B, int D, int use_tanh, g2_stream_t stream);
/* Monte-Carlo KL per row: kl[b] = sum_d log N(z;mu,sigma) - log N(z;pmu,psigma) (models/genesis_config.py:328-336);
 * pmu = psigma = NULL: standard-normal prior.  bwd writes the gradients of all five inputs for a given dkl[B]. */
int g2_mc_kl_fwd_f32(const float* z, const float* mu, const float* sigma, const float* pmu, const float* psigma,
                     float* kl, int B, int D, g2_stream_t stream);
int g2_mc_kl_bwd_f32(const float* z, const float* mu, const float* sigma, const float* pmu, const float* psigma,
                     const float* dkl, float* dz, float* dmu, float* dsigma, float* dpmu, float* dpsigma, int B, int D,
                     g2_stream_t stream);

/* ---- evaluation metrics (metrics.cu) ------------------------------------------------------------------
 * Adjusted Rand index (all pixels / foreground only) and segmentation covering (mean / size-weighted, with / without the
 * background) per image from one confusion matrix per image; replaces utils/misc.py:101-114 (average_ari over
 * sklearn.metrics.adjusted_rand_score) and :173-235 (average_segcover) as called by train.py:536-546, plus the argmax
 * over slots (train.py:541).  Exactly one of log_m [K,B,P] (fp32, slot-major) / pred [B,P] (int64) is non-NULL;
 * inst [B,P] int64 ground-truth labels (0 = background; labels outside 0..31 are ignored); seg_out [B,P] int64 or NULL;
 * out [B][8] doubles = {ari, ari_fg, msc, msc_fg, msc_scaled, msc_fg_scaled, labels present, pixels counted}. */
int g2_seg_metrics(const float* log_m, const int64_t* pred, const int64_t* inst, int64_t* seg_out, double* out, int B,
                   int P, int K, g2_stream_t stream);

/* 3xTF32 operand split (pointwise.cu): hi = x rounded to TF32, lo = x - hi.  mode 0: out [rows,3C] = [hi|hi|lo] (activation
 * side of a tf32x3 contraction whose weight side is [w_hi|w_lo|w_hi]); 1: out [rows,2C] = [hi|lo]; 2: out = hi; 3: out = lo.
 * With it the ill-conditioned layers (MONet / GENESIS-V2 UNet: modules/unet.py:53-65, modules/blocks.py:151-165) run on the
 * tensor cores at fp32-level accuracy through the unchanged conv / GEMM entry points. */
int g2_split_tf32_f32(const float* x, float* out, long rows, int C, int mode, g2_stream_t stream);

/* ---- optimiser (pointwise.cu) --------------------------------------------------------------------
 * Fused Adam over a flat fp32 parameter / gradient arena; replaces torch.optim.Adam.step() in the
 * caller's loop (train.py:175,263).  `step` is a device-resident float counter (1-based). n % 4 == 0. */
int g2_adam_f32(float* p, float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps,
                const float* step, float grad_scale, int zero_grad, g2_stream_t stream);
/* The other two optimisers train.py:171-176 offers, same conventions: RMSprop(lr) = alpha 0.99, eps 1e-8, no momentum;
 * SGD(lr, 0.9) = momentum buffer initialised with the first gradient (step <= 1). */
int g2_rmsprop_f32(float* p, float* g, float* sq, long n, float lr, float alpha, float eps, float grad_scale,
                   int zero_grad, g2_stream_t stream);
int g2_sgd_f32(float* p, float* g, float* buf, long n, float lr, float momentum, const float* step, float grad_scale,
               int zero_grad, g2_stream_t stream);
/* GECO update (utils/geco.py:35-51) on device scalars, one launch, no host sync: state = {beta, err_ema, started};
 * err_kl = {sum over ranks of the batch-mean err, ditto kl} (the tail of the gradient arena after the all-reduce);
 * update = 0 leaves the GECO state alone (plain beta objective).  Also advances the optimiser step counter and writes
 * elbo = err + kl (either may be NULL). */
int g2_geco_step_f32(float* state, const float* err_kl, float* step_count, float* elbo, float inv_world, float goal,
                     float step_size, float alpha, float speedup, float beta_min, float beta_max, int update,
                     g2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GENESIS_B200_H_ */
