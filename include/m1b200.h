/*
 * m1b200.h — C-ABI of libm1b200.so: the sm_100a kernels underneath the M1 (Hierarchical
 * Probabilistic 3D U-Net) forward/backward path.
 *
 * The reference (DIAGNijmegen/prostateMR_3D-CAD-csPCa) has NO native / FFI boundary: every
 * op below is reached implicitly through TensorFlow 2.5 layers.  Each entry point therefore
 * cites the reference *call-site* whose arithmetic it replaces (R: = tf2.5/scripts/model/unets/,
 * L: = tf2.5/scripts/model/losses.py).  The Python host (m1b200.model.unets.networks.M1) binds
 * these through ctypes with DLPack-exported device pointers; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - all tensors are dense NDHWC ("channels last"), caller-owned DEVICE memory, 16-byte aligned;
 *   - activations are fp16, bf16 or fp32 (m1_dtype; gradients of fp16 activations are bf16); statistics,
 *     parameters, gradients of parameters and all reductions are fp32;
 *   - every call takes the cudaStream_t to launch on (as void*), never synchronises, and returns
 *     0 on success; on failure it returns non-zero and m1_last_error() (thread-local) says why;
 *   - there is no CPU fallback: a call on a machine without an sm_100 device fails.
 */
#ifndef M1B200_H_
#define M1B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M1_MAX_SRC 8
#define M1_MAX_OUT 8

/* element types. Activation VALUES are fp32, bf16 or fp16; the GRADIENT of an activation is stored as
 * m1_grad_dtype(value type): fp32 -> fp32, bf16 -> bf16, fp16 -> bf16 (fp16 has too little range for
 * gradients; bf16 gradient storage is harmless). Backward entry points take the
 * VALUE dtype and derive the gradient dtype by this rule. */
typedef enum { M1_F32 = 0, M1_BF16 = 1, M1_F16 = 2 } m1_dtype;
#define M1_GRAD_DTYPE(dt) ((dt) == M1_F16 ? M1_BF16 : (dt))
/* m1_softmax_focal only: the fp32 input already holds probabilities (the softmax is skipped) */
#define M1_PROBS 16

/* gather direction of a convolution launch (see m1_conv_desc) */
typedef enum {
  M1_CONV_FWD = 0,       /* out[o] += in[o*s + k - pad] * W[k]            (Conv3D, dgrad of ConvT) */
  M1_CONV_TRANSPOSED = 1 /* out[o] += in[(o + pad - k)/s] * W[k] if s | .. (ConvT,  dgrad of Conv3D) */
} m1_conv_mode;

typedef enum { M1_ENGINE_AUTO = 0, M1_ENGINE_SIMT = 1, M1_ENGINE_TCGEN05 = 2 } m1_engine;

/* One convolution-shaped launch.  "in" is the tensor that is gathered from (possibly a virtual
 * channel-concatenation of several tensors, R:networks.py:596,604,613,621,653,677,701,725),
 * "out" the tensor that is produced (possibly split over several tensors along channels: the fused
 * conv1||conv4 pair of an SE block, R:network_blocks.py:37,43, or - in a data-gradient launch - the
 * gradients of the concatenated tensors).
 * TF "SAME" padding (R:networks.py:259 'padding':'same'): pad[] is pad_before of the FORWARD
 * convolution this launch belongs to (for M1_CONV_TRANSPOSED it is the forward conv's pad). */
typedef struct {
  int32_t mode;                       /* m1_conv_mode */
  int32_t batch;
  int32_t in_dhw[3];                  /* grid of the gathered tensor(s) */
  int32_t out_dhw[3];                 /* grid of the produced tensor(s) */
  int32_t kernel[3], stride[3], pad[3];
  int32_t nsrc;
  int32_t src_c[M1_MAX_SRC];          /* channels of each gathered tensor */
  int32_t nout;
  int32_t out_c[M1_MAX_OUT];          /* channels of each produced tensor */
  /* fp32 master weights, element strides: W[tap*w_stride_tap + r*w_stride_red + n*w_stride_out]
   * r = reduced (gathered) channel over the concatenation, n = produced channel of output j
   * (one weight tensor per output). */
  int64_t w_stride_tap[M1_MAX_OUT], w_stride_red[M1_MAX_OUT], w_stride_out[M1_MAX_OUT];
  /* w_by_src = 1 (data gradient of several fused layers at once): one weight tensor per (produced j,
   * gathered s) pair, w[j * nsrc + s]; the strides above are then indexed by the GATHERED tensor s and
   * the reduced channel r restarts at 0 for every gathered tensor. */
  int32_t w_by_src;
  int32_t accumulate;                 /* bit j set: outs[j] += result (gradient accumulation) */
  int32_t act_dtype;                  /* m1_dtype of the gathered tensors */
  int32_t out_dtype;                  /* m1_dtype of the produced tensors (wgrad: of dout) */
  int32_t engine;                     /* m1_engine */
  /* tcgen05 engine: element type of the packed weight operand (m1_conv3d_pack_weights), M1_BF16 or M1_F16;
   * 0 = the type of the gathered tensors, which is the only combination the hardware executes: an instruction
   * descriptor with different A and B formats traps with "illegal instruction" (measured on B200, both ways).
   * fp16 mode therefore packs fp16 weights for the forward launches and bf16 weights for the data gradients. */
  int32_t w_dtype;
  /* tcgen05 tiling overrides found by the host's one-off autotuning (0 = heuristic default):
   * conv:  tune[0] = engine variant: 1 = one TMA box per filter tap (conv_tc.cu), 2 = halo tile shared by
   *        the in-plane taps through row-shifted UMMA descriptors (conv_tc_halo.cu; stride-1 gathers only),
   *        3 = multi-tile CTAs: the TMA ring streams across tile boundaries, double-buffered TMEM accumulators,
   *        dedicated epilogue warps (short-K launches: 1x1x1 convolutions, phases of transposed convolutions)
   * wgrad: tune[0] = max voxels per K brick (16..128), tune[1] = taps sharing one dY tile (1 or kw; 2 = SHIFT
   *        mode: the kw taps also share one activation box through row-shifted descriptors, stride 1 only),
   *        tune[2] = pipeline stage cap, tune[3] = 128-row M tiles per CTA (they share the dY tile) */
  int32_t tune[4];
} m1_conv_desc;

typedef struct m1_ctx m1_ctx;

/* ---- context / errors ------------------------------------------------------------------ */
const char* m1_last_error(void);
int  m1_version(void);
/* Creates the per-GPU context (TMA descriptor cache, workspaces).  One ctx per GPU, one host
 * thread per ctx. */
int  m1_ctx_create(int device, m1_ctx** out);
int  m1_ctx_destroy(m1_ctx* ctx);
/* number of kernels this ctx has launched since creation / since the last reset */
int64_t m1_ctx_launch_count(m1_ctx* ctx, int reset);
/* 1 if the tcgen05 engine can take this launch (shape/dtype constraints), else 0 */
int  m1_conv3d_tc_supported(const m1_conv_desc* d);
/* Tiling the tcgen05 engines would use for d (host-side planning only, no GPU needed): which = 0 per-tap
 * convolution {ck, n_tile, n_tiles, bd, bh, bw, stages, k-steps per stage, smem bytes, TMEM columns, CTAs},
 * 1 halo convolution {ck, n_tile, n_tiles, G, bh, bw, P, L, stages, smem bytes, TMEM columns, CTAs per SM,
 * stage bytes, activation-tile bytes}, 2 weight gradient {ck, cb, n_tile, taps per group, M tiles per CTA,
 * voxels per brick, bd, bh, bw, stages, smem bytes, TMEM columns, shift mode, taps-in-M, stage bytes}.
 * Returns the number of values written to out (>= 16 slots), 0 if the engine does not take the launch. */
int  m1_conv3d_plan_info(const m1_conv_desc* d, int which, int32_t* out);
/* 1 if m1_conv3d would run this launch on the halo variant of the tcgen05 engine (honours d->tune[0]) */
int  m1_conv3d_halo_engine(const m1_conv_desc* d);

/* ---- K1/K2: convolution, transposed convolution and their data gradients ----------------
 * replaces tf.keras.layers.Conv3D / Conv3DTranspose (+BiasAdd) at R:networks.py:472,496-553,
 * R:network_blocks.py:37-46,100-103,275 and Conv3DBackpropInputV2 of autodiff.
 * srcs[i]: gathered tensors; w[j]: fp32 master weights of output j; w_packed: bf16 operand
 * pack produced by m1_conv3d_pack_weights (tcgen05 engine only, may be NULL for SIMT);
 * bias[j]: fp32 or NULL; outs[j]: produced tensors. */
int m1_conv3d(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
              const float* const* w, const void* w_packed, const float* const* bias,
              void* const* outs, void* stream);
/* bytes of the bf16 operand pack for d (0 if the tcgen05 engine does not support d) */
int64_t m1_conv3d_packed_bytes(const m1_conv_desc* d);
/* fp32 master weights -> bf16 [tap][n_total][k_total] K-major pack used by the tcgen05 engine */
int m1_conv3d_pack_weights(m1_ctx* ctx, const m1_conv_desc* d, const float* const* w,
                           void* w_packed, void* stream);

/* All operand packs of a model in ONE launch (the re-pack that follows every optimizer step: ~350 packs). The plan
 * holds device-side copies of the job table (weight pointers, strides, pack geometry, output pointer): create it once
 * - outside any CUDA-graph capture - for fixed weight / pack buffers, run it every step, destroy it at the end. */
typedef struct m1_pack_plan m1_pack_plan;
int m1_pack_plan_create(m1_ctx* ctx, int njobs, const m1_conv_desc* const* descs, const float* const* const* ws,
                        void* const* packed, m1_pack_plan** out);
int m1_pack_plan_run(m1_ctx* ctx, const m1_pack_plan* plan, void* stream);
int m1_pack_plan_destroy(m1_pack_plan* plan);

/* ---- K3: weight gradient (Conv3DBackpropFilterV2 of autodiff) + BiasAddGrad ----------------
 * dW_j[tap, r, n] += sum_{batch,o} gathered(o,tap)[r] * dout_j[o, n]   (same strides as d->w_*)
 * dbias_j[n]     += sum dout_j[.., n]  (if dbias[j] != NULL).  Always accumulates (shared
 * weights receive gradients from several passes, R:networks.py:348-352).  d->engine AUTO takes the
 * tcgen05 engine for stride-1 bf16 launches with 16-aligned channel counts, else the CUDA cores. */
int m1_conv3d_wgrad(m1_ctx* ctx, const m1_conv_desc* d, const void* const* srcs,
                    const void* const* douts, float* const* dw, float* const* dbias, void* stream);

/* 1 if the tcgen05 engine takes the weight gradient of output 0 of this launch */
int m1_conv3d_wgrad_tc_supported0(const m1_conv_desc* d);
/* BiasAddGrad on its own: dbias[n] += sum_rows dout[row][n] (Conv3DTranspose layers, whose
 * weight gradient runs with the operand roles swapped) */
int m1_bias_grad(m1_ctx* ctx, const void* dout, int dtype, int64_t rows, int C, float* dbias,
                 void* stream);

/* ---- K4: tfa.layers.InstanceNormalization (eps 1e-3, biased variance) + LeakyReLU(0.1) ----
 * R:networks.py:473,576; R:network_blocks.py:38-44,55,58,104.  stats = [batch][C][2] = mean,rstd */
int m1_inorm_stats(m1_ctx* ctx, const void* x, int dtype, int batch, int64_t voxels, int C,
                   float eps, float* stats, void* stream);
/* y_bf16 (here and in m1_se_gate_fwd / m1_attn_fwd): optional second copy of the output rounded to bf16, or NULL.
 * fp16 mode only: tcgen05.mma.kind::f16 traps (illegal instruction, measured on B200) when its two operands
 * have different formats, so the tensor-core WEIGHT GRADIENT - activations x bf16 output gradients - reads this
 * bf16 twin of every activation that feeds a convolution, while the forward pass reads the fp16 tensor. */
int m1_inorm_act_fwd(m1_ctx* ctx, const void* x, const float* stats, const float* gamma,
                     const float* beta, int dtype, int batch, int64_t voxels, int C,
                     float slope /* 1 = no activation */, void* y, void* y_bf16, void* stream);
/* dy: gradient w.r.t. y; x: the raw (pre-norm) tensor; dx written (or accumulated);
 * dgamma/dbeta accumulate. */
int m1_inorm_act_bwd(m1_ctx* ctx, const void* dy, const void* x, const float* stats,
                     const float* gamma, const float* beta, int dtype, int batch,
                     int64_t voxels, int C, float slope, void* dx, int accumulate,
                     float* dgamma, float* dbeta, void* stream);

/* ---- K5: squeeze-excite gate + multiplicative residual + LeakyReLU + dropout ----------------
 * R:network_blocks.py:68-78 (GAP, conv6, conv7, sigmoid, x_*g, *residual, relu(alpha=.1)) and the
 * dropout that always follows an SE block (R:networks.py:579-582,597,607,616,624,652,676,700,728;
 * tf.nn.dropout semantics: keep iff u >= rate, scale 1/(1-rate)).
 * raw3/raw4 are the conv3/conv4 outputs BEFORE norm3/norm4; their normalisation is applied on
 * the fly.  pool = GAP(norm3(raw3)) [batch][C]. */
int m1_se_squeeze(m1_ctx* ctx, const void* raw3, const float* stats3, const float* gamma3,
                  const float* beta3, int dtype, int batch, int64_t voxels, int C,
                  float* pool, void* stream);
/* gate = sigmoid(W7 . lrelu(W6 . pool + b6) + b7);  W6 [C][Cr], W7 [Cr][C] (Keras 1x1x1 kernels);
 * hidden [batch][Cr] is the pre-activation of conv6 (kept for backward). */
/* stats3 / gamma3 / beta3 != NULL: the squeeze is folded in - pool is first computed (and written) from the
 * statistics of raw3 as m1_se_squeeze does (one launch fewer per block). */
int m1_se_excite_fwd(m1_ctx* ctx, float* pool, const float* w6, const float* b6,
                     const float* w7, const float* b7, int batch, int C, int Cr,
                     float* hidden, float* gate, const float* stats3, const float* gamma3, const float* beta3,
                     void* stream);
/* red5 != NULL (the reductions of m1_se_gate_bwd_reduce): also accumulates the parameter gradients of norm3 / norm4
 * (dgamma3 += A2, dbeta3 += A1 + dpool, dgamma4 += B2, dbeta4 += B1); m1_se_gate_bwd_apply is then called with NULL
 * parameter-gradient pointers. */
int m1_se_excite_bwd(m1_ctx* ctx, const float* dgate, const float* pool, const float* hidden,
                     const float* gate, const float* w6, const float* w7, int batch, int C,
                     int Cr, float* dpool, float* dw6, float* db6, float* dw7, float* db7,
                     const float* red5, float* dgamma3, float* dbeta3, float* dgamma4, float* dbeta4,
                     void* stream);
/* dropout source: u != NULL -> injected uniforms (same dtype fp32, one per element);
 * else Philox4x32-10(seed, stream_id, element index).  rate == 0 -> no dropout. */
/* Philox stream advance per training step when the launch is replayed from a captured CUDA graph:
 * effective stream = stream_id + (*step) * M1_PHILOX_STEP_STRIDE (step: device-resident counter, or NULL) */
#define M1_PHILOX_STEP_STRIDE 4096ull
typedef struct {
  const float* u;
  uint64_t seed;
  uint64_t stream_id;
  float rate;
  const uint64_t* step;
  /* optional keep-mask, 1 bit per element ((elements + 7) / 8 bytes, C % 8 == 0): m1_se_gate_fwd writes it,
   * the two backward kernels read it instead of regenerating the noise (NULL: regenerate) */
  uint8_t* mask;
} m1_dropout;
int m1_se_gate_fwd(m1_ctx* ctx, const void* raw3, const void* raw4, const float* stats3,
                   const float* stats4, const float* gamma3, const float* beta3,
                   const float* gamma4, const float* beta4, const float* gate,
                   const m1_dropout* drop, int dtype, int batch, int64_t voxels, int C,
                   void* out, void* out_bf16, void* stream);
/* Backward of the fused gate INCLUDING the two instance norms: produces draw3/draw4 (gradients
 * w.r.t. the raw conv outputs), dgate [batch][C] (to feed m1_se_excite_bwd) in phase 1, then,
 * after the excite backward produced dpool, phase 2 writes draw3/draw4 and accumulates
 * dgamma/dbeta of norm3/norm4. Scratch: red [batch][C][6] fp32. */
int m1_se_gate_bwd_reduce(m1_ctx* ctx, const void* dout, const void* raw3, const void* raw4,
                          const float* stats3, const float* stats4, const float* gamma3,
                          const float* beta3, const float* gamma4, const float* beta4,
                          const float* gate, const m1_dropout* drop, int dtype, int batch,
                          int64_t voxels, int C, float* red, float* dgate, void* stream);
int m1_se_gate_bwd_apply(m1_ctx* ctx, const void* dout, const void* raw3, const void* raw4,
                         const float* stats3, const float* stats4, const float* gamma3,
                         const float* beta3, const float* gamma4, const float* beta4,
                         const float* gate, const m1_dropout* drop, const float* red,
                         const float* dpool, int dtype, int batch, int64_t voxels, int C,
                         void* draw3, void* draw4, int accumulate /* draw3 / draw4 += (several gates share raw3 / raw4) */,
                         float* dgamma3, float* dbeta3, float* dgamma4, float* dbeta4, void* stream);

/* ---- K6: additive attention gate, R:network_blocks.py:106-130 -------------------------------
 * psi = sigmoid(w_psi . lrelu(theta + up(phi)) + b_psi) on theta's grid (up = nearest, integer
 * floor ratio), y = up(psi) * x on x's grid.  theta [batch][tg][F], phi [batch][gg][F],
 * psi [batch][tg] fp32, x/y [batch][xg][Cx]. */
int m1_attn_fwd(m1_ctx* ctx, const void* theta, const void* phi, const float* w_psi,
                const float* b_psi, const void* x, int dtype, int batch, const int32_t* tg,
                const int32_t* gg, const int32_t* xg, int F, int Cx, float* psi, void* y,
                void* y_bf16, void* stream);
/* dy -> dx (accumulated if acc_dx), dtheta (written), dphi (fp32, accumulated), dw_psi/db_psi acc */
int m1_attn_bwd(m1_ctx* ctx, const void* dy, const void* theta, const void* phi,
                const float* w_psi, const float* psi, const void* x, int dtype, int batch,
                const int32_t* tg, const int32_t* gg, const int32_t* xg, int F, int Cx,
                void* dx, int acc_dx, void* dtheta, float* dphi, float* dw_psi, float* db_psi,
                void* stream);

/* ---- K7: probabilistic latent heads, R:networks.py:637-649 (x4 levels) and KL :373-385 ------
 * ml [batch][voxels][2L] = [mu | logsigma] (output of the 1x1x1 mu_logsig conv, fp32);
 * mode 0: z = mu + exp(clip(logsigma,-0.1,0.1)) * eps, mode 1: z = mu.
 * z is written as activation dtype with zc >= L channels (channels L..zc-1 zero). */
int m1_latent_fwd(m1_ctx* ctx, const float* ml, const float* eps, int mode, int batch,
                  int64_t voxels, int L, int zdtype, int zc, void* z, void* stream);
/* dml += d(z)/d(ml) . dz   (dml fp32 [batch][voxels][2L], accumulated) */
int m1_latent_bwd(m1_ctx* ctx, const void* dz, const float* ml, const float* eps, int mode,
                  int batch, int64_t voxels, int L, int zdtype, int zc, float* dml, void* stream);
/* kl_out[0] += (1/batch) sum_{b,v,c} KL(N(mu_q,s_q) || N(mu_p,s_p))  (tfp MultivariateNormalDiag) */
int m1_kl_fwd(m1_ctx* ctx, const float* ml_q, const float* ml_p, int batch, int64_t voxels,
              int L, float* kl_out, void* stream);
/* dml_q += scale * dKL/dml_q ; dml_p += scale * dKL/dml_p  (scale = loss weight * upstream) */
int m1_kl_bwd(m1_ctx* ctx, const float* ml_q, const float* ml_p, int batch, int64_t voxels,
              int L, float scale, float* dml_q, float* dml_p, void* stream);

/* ---- K8: softmax + focal loss, R:networks.py:388-390,751-755 and L:32-49 --------------------
 * logits [batch][lg][nc] fp32 (ldtype M1_F32; ldtype M1_PROBS = the tensor already holds probabilities,
 * the softmax is skipped - Focal.FL called on predictions), nearest-upsampled by `up` to the label grid
 * xg = lg*up (deep-supervision heads: the 1x1x1 conv commutes with the nearest upsample);
 * softmax written to prob[..., head_off : head_off+nc] of a [batch][xg][prob_c] fp32 tensor;
 * loss_out[0] += head_weight * mean_b sum_{voxels,c} alpha_c y (1-p)^gamma (-log p), p clipped
 * to [1e-7, 1-1e-7] after renormalisation;  dlogits (if != NULL) = dloss/dlogits * grad_scale. */
int m1_softmax_focal(m1_ctx* ctx, const void* logits, int ldtype, const void* y_true, int ydtype,
                     const float* alpha, float gamma, int batch, const int32_t* lg,
                     const int32_t* up, int nc, float* prob, int prob_c, int head_off,
                     float head_weight, float* loss_out, void* dlogits, float grad_scale,
                     void* stream);

/* K8 fused with the final 1x1x1 logits convolution (StitchingProbDecoder / M1Core.logits,
 * R:network_blocks.py:275-278, R:networks.py:526,627) and BOTH of its gradients: one pass reads the
 * decoder features, writes softmax + d(features), accumulates loss, dW [C][nc] and db [nc].
 * y_true == NULL: softmax only. Returns 2 (and sets the error) if (C, nc) has no instantiation - the
 * caller then uses m1_conv3d + m1_softmax_focal. */
int m1_logits_softmax_focal(m1_ctx* ctx, const void* feat, int fdtype, const float* w, const float* bias,
                            const void* y_true, int ydtype, const float* alpha, float gamma, int batch,
                            int64_t voxels, int C, int nc, float* prob, int prob_c, int head_off,
                            float head_weight, float* loss_out, void* dfeat, int acc_dfeat, float* dw,
                            float* db, float grad_scale, void* stream);

/* ---- cascaded two-stage model, R:networks.py:109-193,209-223 ---------------------------------------
 * Backward of the final 1x1x1 logits convolution + softmax for an upstream gradient w.r.t. the PROBABILITIES (the
 * stage-1 softmax feeds the second stage's input and the decision fusion): dlogit = p * (dprob - sum_k p_k dprob_k);
 * dfeat (+)= W dlogit (stored as M1_GRAD_DTYPE(fdtype)); dw += feat x dlogit; db += dlogit.
 * prob [rows][prob_c] (channels head_off .. head_off+nc), dprob [rows][nc] fp32. Returns 2 if (C, nc) has no
 * instantiation. */
int m1_logits_prob_bwd(m1_ctx* ctx, const void* feat, int fdtype, const float* w, const float* prob, int prob_c,
                       int head_off, const float* dprob, int64_t rows, int C, int nc, void* dfeat, int acc_dfeat,
                       float* dw, float* db, void* stream);
/* decision_fusion (strategy as m1_decision_fusion) of the class-1 probabilities p1 = prob1[row][ch1] (row pitch pc1)
 * and p2 = prob2[row][ch2] fused with Focal.FL on the joint prediction: det1 [rows][2] = [1-p1, p1], det2 = [1-j, j]
 * (either may be NULL); if y_true != NULL: loss_out[0] += weight * mean_b sum FL(y, [1-j, j]) and dp1 / dp2 [rows]
 * = grad_scale * weight / batch * d FL / d p1, p2 (NULL: not wanted). Two classes only. */
int m1_fusion_focal(m1_ctx* ctx, const float* prob1, int pc1, int ch1, const float* prob2, int pc2, int ch2,
                    int strategy, const void* y_true, int ydtype, const float* alpha, float gamma, int batch,
                    int64_t voxels, float* det1, float* det2, float weight, float* loss_out, float* dp1, float* dp2,
                    float grad_scale, void* stream);

/* ---- K9: Keras Adam (amsgrad=True: train_model.py:113-120; amsgrad=0: the Keras default) + L2 regulariser
 * gradient.  g' = g*gscale + 2*l2*w ; m,v update; vhat = max(vhat, v) (amsgrad) or v;
 * w -= lr_t * m / (sqrt(vhat) + eps)
 * l2_sq_out[0] += l2 * sum w^2 (regularisation loss term, R:networks.py:259-263) if non-NULL. */
int m1_adam_amsgrad(m1_ctx* ctx, float* w, const float* g, float* m, float* v, float* vhat,
                    int64_t n, float lr_t, float beta1, float beta2, float eps, float l2,
                    float gscale, float* l2_sq_out, int amsgrad, void* stream);

/* Same update with the step size read from DEVICE memory (lr_t_dev[0]): the form that can be captured in a
 * CUDA graph and replayed while the learning-rate schedule advances on the host. */
int m1_adam_amsgrad_dev(m1_ctx* ctx, float* w, const float* g, float* m, float* v, float* vhat,
                        int64_t n, const float* lr_t_dev, float beta1, float beta2, float eps, float l2,
                        float gscale, float* l2_sq_out, int amsgrad, void* stream);

/* ---- K10: train-time augmentations on the device ------------------------------------------------
 * tf2.5/scripts/model/augmentations.py:36-378 (`augment_tensors`: zoom_4D_tensor, axial_4D_hflip, rotate_4D_tensor,
 * translate_4D_tensor, channel_shift_4D_tensor, gamma_shift_4D_tensor, sim_poor_scan_4D_tensor,
 * gaussian_noise_4D_tensor), which the reference maps over the tf.data pipeline on the host CPUs
 * (train_model.py:181). One launch per transform over a whole batch (B, D, H, W, C) fp32; every sample has its own
 * m1_aug_plan (device array of `batch` plans, drawn by the host: model/augmentations.py draw_plans). `in` and `out`
 * must not alias; samples whose transform is switched off are copied. D plays the batch role of TensorFlow's 4-D
 * image ops. M1_AUG_NOISE reads eps ~ N(0,1) of shape (B, D, H, W, 3); M1_AUG_POOR_SCAN needs H == W (the
 * reference resizes to (H, H)). */
typedef struct {
  int32_t zoom_on, zoom_scale;                       /* resize to scale x scale, keep the bottom-right H x W window */
  int32_t flip_on;
  int32_t rot_on, rot_pad, rot_crop_h, rot_crop_w;   /* SYMMETRIC pad, rotate about the centre, central crop offsets */
  float rot_cos, rot_sin, rot_xoff, rot_yoff;        /* tfa.image.rotate: in = (cos x - sin y + xoff, sin x + cos y + yoff) */
  int32_t tr_on, tr_top, tr_bottom, tr_right, tr_left;
  int32_t cs_on, cs_channel, cs_top, cs_bottom, cs_right, cs_left;
  int32_t gamma_on[3];
  float gamma;
  int32_t poor_on[3];
  int32_t noise_on;
  float noise_std;
} m1_aug_plan;
typedef enum {
  M1_AUG_ZOOM = 0, M1_AUG_FLIP = 1, M1_AUG_ROTATE = 2, M1_AUG_TRANSLATE = 3, M1_AUG_CHANNEL_SHIFT = 4,
  M1_AUG_GAMMA = 5, M1_AUG_POOR_SCAN = 6, M1_AUG_NOISE = 7
} m1_aug_op;
int m1_augment(m1_ctx* ctx, int op, const float* in, float* out, const float* eps, const m1_aug_plan* plans,
               int batch, int D, int H, int W, int C, void* stream);

/* ---- small utilities used by the host ------------------------------------------------------- */
int m1_cast(m1_ctx* ctx, const void* src, int sdtype, void* dst, int ddtype, int64_t n,
            void* stream);
/* dst[..., dst_off:dst_off+c] = src[..., src_off:src_off+c] over `rows` rows (channel slicing of
 * the model input, R:networks.py:300-301, Q4) */
int m1_copy_channels(m1_ctx* ctx, const void* src, int sdtype, int src_c, int src_off, void* dst,
                     int ddtype, int dst_c, int dst_off, int c, int64_t rows, void* stream);
int m1_axpy(m1_ctx* ctx, const void* x, int dtype, float a, void* y, int64_t n, void* stream);
/* out[i] ~ N(0,1) from Philox4x32-10(seed, stream_id, i/4) + Box-Muller: the latent noise of
 * tfp MultivariateNormalDiag.sample() (R:networks.py:647,671,695) when no eps is injected */
int m1_philox_normal(m1_ctx* ctx, uint64_t seed, uint64_t stream_id, float* out, int64_t n,
                     void* stream);
/* same, stream advanced by a device-resident step counter (graph replay): stream_id + *step_dev * 4096 */
int m1_philox_normal_step(m1_ctx* ctx, uint64_t seed, uint64_t stream_id, const uint64_t* step_dev, float* out,
                          int64_t n, void* stream);
/* decision fusion of the cascaded model, R:networks.py:209-223 (strategy 0 identity, 1 noisy-or,
 * 2 bayes): out [rows][2] = [1-j, j] */
int m1_decision_fusion(m1_ctx* ctx, const float* prior, const float* follow, int strategy,
                       int64_t rows, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* M1B200_H_ */
