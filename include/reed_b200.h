/* C-ABI of libreed_sm100.so - the B200 (sm_100a) drop-in for the REED image hot path.
 *
 * The reference (ChenyuWang-Monica/REED) has no native code on this path: its SiT training step is Python over
 * PyTorch library kernels.  Each entry point below therefore replaces the PyTorch call sequence at the cited
 * reference lines (paths relative to /root/reference/image); the Python-side binding is reed_b200/_cabi.py
 * (ctypes), wrapped as autograd functions in reed_b200/ops.py.  INTEGRATION.md shows the stub a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; reed_last_error() returns the thread-local message
 *   - plain pointers and sizes only; all buffers are device memory owned by the caller (PyTorch's allocator);
 *     the library allocates nothing persistent and never synchronises: calls only enqueue work on `stream`
 *   - `stream` is a cudaStream_t passed as void*; re-entrant, no global "current stream"
 *   - dtype codes: 0 = float32, 1 = bfloat16.  "act dtype" is the activation dtype of the precision mode
 *     (0: fp32 mode, parity bar 1e-5 relative; 1: bf16 tensor-core mode, parity bar 2e-2 relative)
 *   - backend codes (gemm/attention): 0 = auto, 1 = force the fp32-math SIMT kernel, 2 = require the tensor-core kernel
 *     (gemm only: 3 / 4 = require the tcgen05 kernel with cta_group::1 / cta_group::2 instead of the planner's choice)
 */
#ifndef REED_B200_H_
#define REED_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int reed_version(void);
const char* reed_last_error(void);
/* 0 iff the current CUDA device is an sm_100 part; copies its name into `name` when non-NULL. */
int reed_device_check(char* name, int name_len);
/* Number of SMs the persistent tensor-core GEMM grids leave free from now on (0 = none).  The data-parallel trainer
 * sets it around backward so the overlapped NCCL gradient all-reduce (train.py:151,293,401 - accelerate/DDP) finds
 * SMs without splitting a GEMM grid into two waves.  Host-side state, not a stream operation. */
int reed_gemm_reserve_sms(int n);
/* Running count of tcgen05 GEMM kernel launches issued through reed_gemm / reed_gemm_wgrad_bias since the library was
 * loaded (host-side counter; measurement bookkeeping: bench.py's roofline sample must cover every one of a step). */
int reed_gemm_tcgen05_launches(long long* out);

/* D[M,N] = epilogue(A[M,K] . B[N,K]^T)   -- every nn.Linear of the model and its dgrad/wgrad.
 * Replaces: timm Attention.qkv/.proj and Mlp.fc1/.fc2 (models/sit.py:114-124,134-135), adaLN linears (125-133,
 * 148-154), projector MLP (17-24, 292-301), t-embedder MLP (38-42, 69), patch-embed conv as a linear (198-200, 279),
 * final linear (147, 156), and autograd's dgrad/wgrad GEMMs for all of them.
 *   a_mn_major / b_mn_major: 0 = operand stored [rows, K] (K contiguous); 1 = stored [K, rows] (rows contiguous).
 *     forward y = x W^T : (0, 0);  dgrad dx = dy W : A = dy (0), B = W stored [N_out, K_in] (1);
 *     wgrad dW = dy^T x : A = dy stored [M, N_out] (1), B = x stored [M, K_in] (1).
 *   epilogue: 0 none (+bias; `accumulate` adds into fp32 D)          1 GELU(tanh)   2 SiLU  (pre-activation -> out2)
 *             3 D = aux + gate[row / rows_per_group] * (acc + bias)   (aux fp32 residual; y -> out2)
 *             4 D = acc * gelu_tanh'(aux)   5 D = acc * silu'(aux)     (aux = saved pre-activation, act dtype)
 *   bias: fp32 [N] or NULL.  gate: fp32 rows of pitch ld_gate.  out2: act dtype, pitch ld_out2, or NULL. */
int reed_gemm(int act_dtype, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
              void* D, int64_t ldd, int d_dtype, int M, int N, int K, int epilogue, const void* bias, const void* aux,
              int64_t ld_aux, const void* gate, int64_t ld_gate, int rows_per_group, void* out2, int64_t ld_out2,
              int accumulate, int backend, void* stream);

/* Weight gradient of a Linear together with its bias gradient, one tensor-core GEMM (bf16 operands, fp32 outputs):
 *   dW[n_out, k_in] (+)= dy^T x   and   db[n_out] += sum_tokens dy       (autograd of timm Attention.qkv / Mlp.fc1,
 *   models/sit.py:114-124).  dy is [tokens, n_out] (pitch ld_dy); x_ext is [tokens, k_in + 8] (pitch ld_x) with ones in
 *   column k_in (written by reed_ln_modulate_fwd), so db is one more output column of the GEMM instead of a separate
 *   reduction pass over dy.  db must be zeroed by the caller once per step; `accumulate` adds into dW. */
int reed_gemm_wgrad_bias(const void* dy, int64_t ld_dy, const void* x_ext, int64_t ld_x, void* dW, int64_t ldd, void* db,
                         int n_out, int k_in, int tokens, int accumulate, void* stream);

/* Weight and bias gradient of a Linear whose contraction is the batch (rows <= 64): autograd of `adaLN_modulation` and the
 * timestep-embedder linears (models/sit.py:125-128, 40-44) - an outer-product stream bounded by the write of dW.
 *   dW[n_out, k_in] (+)= dy^T x;  db[n_out] += column sums of dy (NULL to skip; the caller zeroes it once per step).
 *   dy [rows, n_out] (fp32 or bf16, pitch ld_dy), x [rows, k_in] (bf16, or fp32 with fp32 dy; pitch ld_x). */
int reed_outer_wgrad(const void* dy, int dy_dtype, int64_t ld_dy, const void* x, int x_dtype, int64_t ld_x, void* dW, int64_t ldd,
                     void* db, int n_out, int k_in, int rows, int accumulate, void* stream);

/* Several Linear layers that share their input, as ONE tensor-core launch: the adaLN-Zero modulation linears of all
 * transformer blocks, `adaLN_modulation(c)` of models/sit.py:125-133 (28 x [6D, D] weights in separate allocations, the same
 * silu(c) input) and the gradient of that shared input.  bf16 operands, fp32 D, M <= 1024 rows, groups <= 32.
 * A / B are HOST arrays of `groups` device pointers (mode 0 reads A[0] only); enqueue-only and graph-capturable.
 *   mode 0: D[M, groups * per_group] = A[0][M,K] . [B_0; B_1; ...]^T + bias[groups * per_group]   (B_g [per_group, K], pitch ldb;
 *           per_group % 256 == 0) - block g's modulation vectors are columns g * per_group ... of D.
 *   mode 1: D[M, N] (+)= sum_g A_g[M, per_group] . B_g[per_group, N]   (K = groups * per_group, per_group % 64 == 0; split over
 *           the reduction, fp32 atomics; `accumulate` 0 zeroes D first). */
int reed_gemm_grouped(int mode, const void* const* A, int64_t lda, const void* const* B, int64_t ldb, int groups, int per_group,
                      void* D, int64_t ldd, int M, int N, int K, const void* bias, int accumulate, void* stream);

/* Multi-head attention, non-causal, scale head_dim^-0.5, over the packed qkv GEMM output.
 * Replaces timm Attention.forward's reshape/permute + F.scaled_dot_product_attention (models/sit.py:13,114-118,134).
 *   qkv [B,T,3,H,hd] (act dtype) -> o [B,T,H,hd] (act dtype), lse [B,H,T] fp32 (saved for backward).
 *   backend: 0 auto - bf16 with T = 128 n and head_dim 64 / 72 runs the flash-style tcgen05 kernels (P in tensor memory;
 *   the backward is one kernel over clusters of T/128 CTAs with a shared-memory ring for dQ), other bf16 shapes the
 *   mma.sync or SIMT kernels, fp32 the SIMT kernels; 1 force SIMT; 2 require a tensor-core kernel; 3 require mma.sync;
 *   4 / 5 require the tcgen05 kernels. */
int reed_attn_fwd(int act_dtype, const void* qkv, void* o, void* lse, int B, int T, int H, int hd, int backend,
                  void* stream);
/* dqkv [B,T,3,H,hd] from d_o [B,T,H,hd]; delta [B,H,T] fp32 is workspace (rowsum(dO*O)). */
int reed_attn_bwd(int act_dtype, const void* qkv, const void* o, const void* d_o, const void* lse, void* dqkv,
                  void* delta, int B, int T, int H, int hd, int backend, void* stream);

/* Optional q/k LayerNorm of timm Attention(qk_norm=True) (models/sit.py:114-116; nn.LayerNorm(head_dim), eps 1e-5),
 * applied per (token, head) to the q and k thirds of the packed qkv [rows, 3, H, hd]; v is copied through.
 *   fwd: out (act dtype, same layout), stats fp32 [rows, H, 2, 2] = (mean, rstd) of q and k (saved for backward).
 *   bwd: dqkv (w.r.t. the raw qkv) from dout (w.r.t. the normalised one); dwq/dbq/dwk/dbk fp32 [hd] accumulated into. */
int reed_qk_norm_fwd(const void* qkv, int act_dtype, const void* wq, const void* bq, const void* wk, const void* bk,
                     void* out, void* stats, int64_t rows, int H, int hd, float eps, void* stream);
int reed_qk_norm_bwd(const void* dout, int act_dtype, const void* qkv, const void* stats, const void* wq, const void* wk,
                     void* dqkv, void* dwq, void* dbq, void* dwk, void* dbk, int64_t rows, int H, int hd, void* stream);

/* out = LayerNorm(x; no affine, eps) * (1 + scale[g]) + shift[g], g = row / rows_per_group.
 * Replaces norm1/norm2/norm_final + modulate (models/sit.py:26-27,113,119,134-135,146,155).
 *   x fp32 [M,D]; shift/scale fp32 rows of pitch ld_mod; out act dtype, rows of pitch ld_out >= D (multiple of 8);
 *   mean/rstd fp32 [M] (saved).  With bf16 output and ld_out >= D + 8 every row also gets [1,0,0,0,0,0,0,0] at columns
 *   D..D+7: the "ones column" reed_gemm_wgrad_bias contracts against to produce the bias gradient. */
int reed_ln_modulate_fwd(const void* x, const void* shift, const void* scale, int64_t ld_mod, int rows_per_group,
                         void* out, int64_t ld_out, int act_dtype, void* mean, void* rstd, int M, int D, float eps,
                         void* stream);
/* dx = dres + LN-backward(dout * (1+scale)); dshift[g] += sum dout; dscale[g] += sum dout * xhat (fp32 atomics).
 * dres may be NULL.  Replaces autograd of the above plus the residual-gradient add.  ld_mod is the row pitch of the
 * modulation vectors (scale, gate), ld_dmod that of their gradients (dshift, dscale, dgate): the two differ when the
 * vectors of all blocks come out of one grouped GEMM (reed_gemm_grouped) while every block keeps its own gradient buffer. */
int reed_ln_modulate_bwd(const void* dout, int act_dtype, const void* x, const void* mean, const void* rstd,
                         const void* scale, int64_t ld_mod, int64_t ld_dmod, int rows_per_group, const void* dres, void* dx,
                         void* dshift, void* dscale, int M, int D, void* stream);
/* The two above in one pass over the rows: dx as in reed_ln_modulate_bwd, then the gate backward applied to that dx
 * (dy = gate[g] * dx, dgate[g] += sum dx * y, dbias += column sums of dy).  Used for the MLP-branch LayerNorm backward
 * followed by the attention-branch gate backward of the same block (models/sit.py:134-135): the fp32 residual gradient
 * is written once and not re-read.  gate shares ld_mod with scale, dgate shares ld_dmod with dshift/dscale. */
int reed_ln_modulate_gate_bwd(const void* dout, int act_dtype, const void* x, const void* mean, const void* rstd,
                              const void* scale, int64_t ld_mod, int64_t ld_dmod, int rows_per_group, const void* dres, void* dx,
                              void* dshift, void* dscale, const void* y, const void* gate, void* dy, void* dgate,
                              void* dbias, int M, int D, void* stream);
/* Backward of x_new = x + gate[g] * y (models/sit.py:134-135): dy = gate * dxn (act dtype), dgate[g] += sum dxn * y,
 * dbias (optional, fp32 [D]) += column sums of dy. */
int reed_gate_bwd(const void* dxn, const void* y, int act_dtype, const void* gate, int64_t ld_mod, int64_t ld_dmod,
                  int rows_per_group, void* dy, void* dgate, void* dbias, int M, int D, void* stream);
/* out[n] += sum_m src[m,n]  (bias gradients). */
int reed_colsum(const void* src, int act_dtype, int64_t ld, void* out, int M, int N, void* stream);
/* op 0: dtype cast; op 1: SiLU then cast (the SiLU in front of every adaLN linear, models/sit.py:126,149); op 2: exact
 * (erf) GELU then cast (nn.GELU() of the DINOv2 target encoder's MLP, utils.py:92-105). */
int reed_unary(const void* in, int in_dtype, void* out, int out_dtype, int op, int64_t n, void* stream);
/* dx = dy * act'(h); act 1 = GELU(tanh), 2 = SiLU. */
int reed_act_bwd(const void* dy, int d_dtype, const void* h, int h_dtype, void* dx, int act, int64_t n, void* stream);
/* Token mean for the text projector (models/sit.py:292,301) and its backward. */
int reed_group_mean_fwd(const void* x, void* out, int out_dtype, int groups, int rows_per_group, int D, void* stream);
int reed_group_mean_bwd(const void* dy, void* dx, int groups, int rows_per_group, int D, int accumulate, void* stream);
int reed_add_f32(const void* a, const void* b, void* out, int64_t n, void* stream);

/* SILoss (loss.py).  path_type: 0 linear, 1 cosine.  t fp32 [batch].
 * interp : x_t = alpha_t x + sigma_t eps                                    (loss.py:49-64,175-176)
 * mse_fwd: denoise[b] = mean((pred - (dalpha x + dsigma eps))^2)             (loss.py:178-186)
 * mse_bwd: dpred = g[b] * 2 (pred - target) / per_sample
 * cos_fwd: align[b] += -(1/T) sum_t cos(z~[b,t], z[b,t]) (F.normalize eps 1e-12); stats [batch*T,3] saved (204-221)
 * cos_bwd: dz~ from g[b]. */
int reed_siloss_interp(const void* x, const void* eps, const void* t, void* xt, int batch, int per_sample,
                       int path_type, void* stream);
int reed_siloss_mse_fwd(const void* pred, const void* x, const void* eps, const void* t, void* denoise, int batch,
                        int per_sample, int path_type, void* stream);
int reed_siloss_mse_bwd(const void* pred, const void* x, const void* eps, const void* t, const void* g, void* dpred,
                        int batch, int per_sample, int path_type, void* stream);
int reed_siloss_cos_fwd(const void* zt, int zt_dtype, const void* z, int z_dtype, void* stats, void* align, int batch,
                        int T, int Z, void* stream);
int reed_siloss_cos_bwd(const void* zt, int zt_dtype, const void* z, int z_dtype, const void* stats, const void* g,
                        void* dzt, int batch, int T, int Z, void* stream);

/* One fused sampler update on the fp64 state (samplers.py:61-104 Euler/Heun, 124-187 Euler-Maruyama).
 *   v: model output in model_dtype, n elements, or 2n (conditional half first) when guided.
 *   sde=1: slope = v - 0.5*(2 t)*score(v, x, t) (samplers.py:15-43,150-151), else slope = v.
 *   guided: slope = s_uncond + cfg * (s_cond - s_uncond).   d_prev != NULL: Heun corrector x + dt*(d_prev+slope)/2.
 *   eps != NULL: + sqrt(2 t) * sqrt(|dt|) * eps.   d_out (optional) receives the guided slope.
 *   x_model (optional) receives x_next cast to model_dtype, written twice when dup_out (next evaluation guided). */
int reed_sampler_step(const void* x_cur, const void* v, int model_dtype, const void* eps, const void* d_prev,
                      void* d_out, void* x_next, void* x_model, int64_t n, int guided, int dup_out, int sde,
                      int path_type, double cfg, double t_cur, double dt, void* stream);
int reed_sampler_cast(const void* x, void* x_model, int model_dtype, int64_t n, int dup, void* stream);

/* Gradient exchange of the data-parallel step over NVSwitch multicast (NVLS), fused with the arithmetic on either side
 * (train.py:151,293,401 DDP all-reduce -> 402-412 clip / AdamW / EMA).  Buffers are symmetric-memory allocations with a
 * multicast mapping; the caller orders ranks with stream barriers (see reed_b200/image/nvls.py).
 *   reduce_scatter_sumsq: grad_local[lo, lo+n) = SUM over ranks of that slice, read through grad_multicast with
 *     multimem.ld_reduce (the switch adds in flight); *norm_sq (double, optional) += sum of squares of the result.
 *   adamw_ema_mc: reed_adamw_ema on an owned slice whose bf16 operands are stored through the multicast address
 *     (shadow_multicast, already offset to the slice): the all-gather of the new weights rides in the optimizer's stores. */
int reed_nvls_reduce_scatter_sumsq(const void* grad_multicast, void* grad_local, int64_t lo, int64_t n, void* norm_sq,
                                   int ctas, void* stream);
int reed_adamw_ema_mc(void* p, const void* g, void* m, void* v, void* ema, void* shadow_multicast, int64_t n,
                      const void* norm_sq, float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, float ema_decay, const void* step_dev, void* stream);

/* Raw-image preprocessing for the frozen target encoders (train.py:53-74 preprocess_raw_image): x/255, per-channel
 * (x - mean)/std, bicubic resize (ATen upsample_bicubic2d, align_corners=False, A=-0.75) to out_size when out_size != in_size.
 *   src [batch, channels, in_size, in_size] uint8 (src_dtype 2) or fp32 (0), values 0..255; dst fp32 (0) or bf16 (1);
 *   mean / stdv: HOST arrays of `channels` (<= 4) floats; resize_first = 1 for the clip order (resize, then normalise). */
int reed_preprocess_image(const void* src, int src_dtype, void* dst, int dst_dtype, int batch, int channels, int in_size,
                          int out_size, const float* mean, const float* stdv, int resize_first, void* stream);

/* VAE-posterior draw of the latent data path (train.py:84-91 sample_posterior with the per-channel latents_scale /
 * latents_bias of train.py:226-231): out[b,c,:] = ((moments[b,c,:] + moments[b,C+c,:] * noise[b,c,:]) * scale[c]) + bias[c],
 * every operation rounded separately (bit-identical to the reference's PyTorch kernel sequence for the same noise).
 * moments fp32 [batch, 2*channels, hw]; noise, out fp32 [batch, channels, hw]; scale / bias: device fp32 [channels], or
 * NULL to use the scalar arguments. */
int reed_sample_posterior(const void* moments, const void* noise, const void* scale, const void* bias,
                          float scale_scalar, float bias_scalar, void* out, int batch, int channels, int hw,
                          void* stream);

/* Optimizer tail over flat fp32 buffers (train.py:94-105 update_ema, 253-259 AdamW, 402-412 clip/step/EMA).
 * grad_sumsq: *out (double) += sum g^2.   adamw_ema: g *= grad_scale * min(1, max_norm/(grad_scale*sqrt(*norm_sq)+1e-6))
 * (norm_sq NULL = no clipping), torch.optim.AdamW update with 1-based `step`, ema = decay*ema + (1-decay)*p, and
 * (optional) bf16 shadow of the new weights for the GEMMs.  step_dev (optional): device int32 holding the step, read
 * instead of `step` so that a captured CUDA graph of the train step can be replayed.  ema_update: EMA only (frozen
 * pos_embed). */
int reed_grad_sumsq(const void* g, int64_t n, void* out, void* stream);
int reed_adamw_ema(void* p, const void* g, void* m, void* v, void* ema, void* shadow_bf16, int64_t n,
                   const void* norm_sq, float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, float ema_decay, const void* step_dev, void* stream);
int reed_ema_update(const void* p, void* ema, int64_t n, float decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REED_B200_H_ */
