/*
 * bsig.h -- C ABI of libbsig_b200.so: the B200 (sm_100a) kernels behind the
 * BayesSimIG inference hot path (trajectory summarizers -> MDNN / MDRFF
 * mixture-density training -> mixture-of-Gaussians posterior).
 *
 * The reference (NVlabs/bayes-sim-ig) is pure Python and has no FFI; its
 * boundary for this path is the Python API resolved by name at
 * bayes_sim_ig/bayes_sim.py:56,82.  Each entry point below therefore cites the
 * reference torch / numpy call site it replaces; the Python mirror of the
 * reference API (bayes_sim_ig_b200/) reaches these functions through ctypes.
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it is
 *     documented as host; tensors are dense row-major fp32 unless stated;
 *   - sizes are int64_t; `stream` is a cudaStream_t passed as void*;
 *   - functions only enqueue work on `stream` (they never synchronise and
 *     are CUDA-graph capturable) unless documented otherwise;
 *   - return value 0 = ok, non-zero = error; text via bsig_last_error();
 *   - no function allocates device memory: scratch comes in as `ws`.
 */
#ifndef BSIG_H_
#define BSIG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSIG_VERSION 100

/* activation / epilogue selectors for bsig_linear_* */
#define BSIG_ACT_NONE 0
#define BSIG_ACT_TANH 1

/* GEMM engines: SIMT fp32 FFMA tiles, or tcgen05 tensor cores (tf32 operands,
 * fp32 accumulate in TMEM; TF32X3 = error-compensated 3-pass split that keeps
 * fp32-level accuracy). */
#define BSIG_GEMM_AUTO (-1)   /* TC_TF32X3 for large K-contiguous problems, SIMT otherwise */
#define BSIG_GEMM_SIMT 0
#define BSIG_GEMM_TC_TF32 1
#define BSIG_GEMM_TC_TF32X3 2

const char* bsig_last_error(void);
int bsig_version(void);
/* number of kernel launches this library has enqueued in this process (host counter;
 * launches recorded into a CUDA graph are counted once, at capture time) */
int64_t bsig_launch_count(void);
/* Programmatic dependent launch of the training-step kernels (default on; env BSIG_PDL=0
 * or bsig_set_pdl(0) turns it off).  No reference counterpart (launch plumbing). */
int bsig_set_pdl(int enabled);
/* host query: SM count and compute capability of the current device */
int bsig_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ summarizers
 * states [n, t_states, d], actions [n, t_actions, a]; only the leading steps
 * are read. */

/* summary_start / summary_waypts (utils/summarizers.py:65-70, 73-87):
 * out [n, max_t*(d+a)], out[i, t*(d+a)+j] = j<d ? states[i,t,j] : actions[i,t,j-d].
 * Requires t_states >= max_t and t_actions >= max_t. */
int bsig_summary_start(const float* states, const float* actions, float* out,
                       int64_t n, int64_t t_states, int64_t t_actions,
                       int64_t d, int64_t a, int64_t max_t, void* stream);

/* cross_correlation / summary_corr / summary_corrdiff (summarizers.py:90-130):
 * w leading steps; sf = adjacent state-dimension differences (use_state_diff)
 * or the first d-1 dims, flattened [w*(d-1)]; af [w*a];
 * out [n, w*(d-1)*w*a + 2] = [ sf (x) af | mean(sf) | unbiased std(sf) ].
 * *nonfinite_flag (int32, device) is OR-ed with 1 if any output is not finite
 * (the reference asserts isfinite at summarizers.py:120). */
int bsig_summary_crosscorr(const float* states, const float* actions, float* out,
                           int64_t n, int64_t t_states, int64_t t_actions,
                           int64_t d, int64_t a, int64_t w, int use_state_diff,
                           int* nonfinite_flag, void* stream);

/* Time-major rollout ingestion: the same two summarizers reading
 * states [t_states, n, d] / actions [t_actions, n, a] exactly as a vectorised simulator
 * writes them step by step -- no per-episode stack / cat assembly
 * (utils/collect_trajectories.py:55-69,86-89) and no transpose.  Outputs are identical to
 * the trajectory-major entry points on the transposed buffers. */
int bsig_summary_start_tm(const float* states, const float* actions, float* out,
                          int64_t n, int64_t t_states, int64_t t_actions,
                          int64_t d, int64_t a, int64_t max_t, void* stream);
int bsig_summary_crosscorr_tm(const float* states, const float* actions, float* out,
                              int64_t n, int64_t t_states, int64_t t_actions,
                              int64_t d, int64_t a, int64_t w, int use_state_diff,
                              int* nonfinite_flag, void* stream);

/* ------------------------------------------------- fused cross-correlation -> first layer
 * utils/summarizers.py:106-119 (the outer product sf (x) af and its two
 * statistics) feeding models/mdnn.py:108 (first nn.Linear) through bayes_sim.py:108-113.
 * The summary row x[i, p*q_n + q] = sf[i,p] * af[i,q] is rank one, so only its factors are
 * ever stored:  fac [n, ldf] = [ sf (s = w*(d-1)) | af (q = w*a) | mean(sf) | std(sf) | 0.. ]
 * (ldf >= s+q+2).  mean / std / the non-finite flag are exactly those of
 * bsig_summary_crosscorr on the same rollouts. */
int bsig_corr_factors(const float* states, const float* actions, float* fac, int64_t ldf,
                      int64_t n, int64_t t_states, int64_t t_actions, int64_t d, int64_t a,
                      int64_t w, int use_state_diff, int time_major, int* nonfinite_flag,
                      void* stream);
/* 1 if the fused first-layer kernels take this shape (minibatch `batch` <= 128 rows for the
 * weight gradient, forward passes of up to rows_max rows, n_out <= 128,
 * s*q + 2 even and < 2^20, factor rows within shared memory), else 0: the caller then materialises the
 * summary and uses bsig_linear_*. */
int bsig_corr_linear_applicable(int64_t batch, int64_t rows_max, int64_t n_out, int64_t s,
                                int64_t q);
int64_t bsig_corr_linear_ws_bytes(int64_t m, int64_t n_out, int64_t s, int64_t q);
/* y [m, n_out] = act( x w^T + b ) with x [m, s*q+2] generated on the fly from
 * fac[rows[i]] (rows nullable), w [n_out, s*q+2] row-major (8-byte aligned base):
 * nn.Linear (+ Tanh) of mdnn.py:108 on the never-materialised summary.  tcgen05, TF32x3. */
int bsig_corr_linear_fwd(const float* fac, int64_t ldf, const int64_t* rows, int64_t s,
                         int64_t q, const float* w, const float* b, float* y, int64_t m,
                         int64_t n_out, int act, void* ws, int64_t ws_bytes, void* stream);
/* Random Fourier features of the never-materialised summary (models/rff.py:128-132 on the output
 * of summarizers.py:106-119, the MDRFF input of bayes_sim.py:108-113): out [m, 2*nf_half] =
 * scale * [cos(x coeff^T) | sin(x coeff^T)], coeff [nf_half, s*q+2] = freqs / sigma, same
 * kernels and workspace as bsig_corr_linear_fwd (nf_half <= 128). */
int bsig_corr_rff_features(const float* fac, int64_t ldf, const int64_t* rows, int64_t s, int64_t q,
                           const float* coeff, float* out, int64_t m, int64_t nf_half, float scale,
                           void* ws, int64_t ws_bytes, void* stream);
/* dw [n_out, s*q+2] = dy^T [n_out, m] x [m, s*q+2] (autograd of mdnn.py:108, weight part).
 * exp_avg != NULL: torch.optim.Adam (mdnn.py:203,234; same arithmetic as bsig_adam_step,
 * gradient scaled by grad_scale) is applied to w / exp_avg / exp_avg_sq in the epilogue and
 * the gradient is only stored if dw != NULL; exp_avg == NULL: dw is stored (data-parallel
 * exchange, external optimisers). */
int bsig_corr_linear_wgrad(const float* dy, const float* fac, int64_t ldf, const int64_t* rows,
                           int64_t s, int64_t q, int64_t m, int64_t n_out, float* dw, float* w,
                           float* exp_avg, float* exp_avg_sq, int64_t step, float lr, float beta1,
                           float beta2, float eps, float grad_scale, void* stream);

/* summary_signatory (summarizers.py:144-168; arithmetic = signatory.signature):
 * path_t = [t+1 | states[i,t,:] | actions[i,t,:]], t < len; depth in {1,2,3};
 * out [n, sum_{k<=depth} c^k], c = 1+d+a, levels concatenated, each C-order.
 * depth 3 requires c <= 24. */
int bsig_signature_fwd(const float* states, const float* actions, float* out,
                       int64_t n, int64_t len, int64_t t_states, int64_t t_actions,
                       int64_t d, int64_t a, int depth, void* stream);
/* gradient of a scalar wrt the path given grad_out [n, siglen] (makes the summarizer
 * differentiable, as signatory's is): d_states [n, len, d], d_actions [n, len, a]
 * (the time channel has no gradient).  depth 3 is implemented for c <= 8 channels. */
int bsig_signature_bwd(const float* states, const float* actions, const float* grad_out,
                       float* d_states, float* d_actions,
                       int64_t n, int64_t len, int64_t t_states, int64_t t_actions,
                       int64_t d, int64_t a, int depth, void* stream);

/* ---------------------------------------------------------------- dense layers
 * Replaces nn.Linear (+Tanh) at models/mdnn.py:68-87,108-119 and the RFF
 * projection at models/rff.py:128-132 (cuBLAS sgemm + pointwise in the
 * reference).  x [m, k] (row stride ldx), w [n, k], b [n], y [m, n].
 * x_rows (int64, device, nullable): gather, row i of the operand is
 * x[x_rows[i]] -- fuses the minibatch gather of mdnn.py:221-222.
 * ws: scratch for split-K partials, ws_bytes >= bsig_linear_ws_bytes(). */
int64_t bsig_linear_ws_bytes(int64_t m, int64_t n, int64_t k);
int bsig_linear_fwd(const float* x, int64_t ldx, const int64_t* x_rows,
                    const float* w, const float* b, float* y,
                    int64_t m, int64_t n, int64_t k, int act, int engine,
                    void* ws, int64_t ws_bytes, void* stream);
/* dx [m,k] = (dy [m,n]) @ w [n,k]; if act_prev == BSIG_ACT_TANH the result is
 * multiplied by (1 - h_prev^2) where h_prev [m,k] is the previous layer's
 * tanh output (fused dtanh). */
int bsig_linear_dgrad(const float* dy, const float* w, const float* h_prev, float* dx,
                      int64_t m, int64_t n, int64_t k, int act_prev, int engine,
                      void* ws, int64_t ws_bytes, void* stream);
/* The last kernel of a single-GPU backward pass: weight gradient of the FIRST layer (weight
 * [n,k] at param[0], bias [n] at param[n*k]; x optionally row-gathered) with torch.optim.Adam
 * (mdnn.py:203,234; arithmetic of bsig_adam_step) applied in its epilogue, and Adam of all other
 * parameters param[n*k+n : n_params] from grad[...] by the otherwise idle CTAs of the launch.
 * Only for layers the small engine takes (n*k <= 128 K, m <= 8192). */
int bsig_linear_wgrad_adam(const float* dy, const float* x, int64_t ldx, const int64_t* x_rows,
                           int64_t m, int64_t n, int64_t k, float* param, const float* grad,
                           float* exp_avg, float* exp_avg_sq, int64_t n_params, int64_t step,
                           float lr, float beta1, float beta2, float eps, void* stream);
/* db [n] = colsum(dy [m,n]): the bias gradient alone (autograd of mdnn.py:108-119), for layers
 * whose weight gradient is formed by bsig_corr_linear_wgrad. */
int bsig_linear_colsum(const float* dy, float* db, int64_t m, int64_t n, void* stream);
/* dw [n,k] = dy^T [n,m] @ x [m,k] (x optionally row-gathered); db [n] = colsum(dy). */
int bsig_linear_wgrad(const float* dy, const float* x, int64_t ldx, const int64_t* x_rows,
                      float* dw, float* db, int64_t m, int64_t n, int64_t k, int engine,
                      void* ws, int64_t ws_bytes, void* stream);
/* dpre = dy * (1 - y^2) elementwise (tanh backward, used by the autograd path) */
int bsig_tanh_bwd(const float* dy, const float* y, float* dpre, int64_t count, void* stream);

/* RFF._to_cos_sin_features (rff.py:128-132): coeff [nf_half, d] = freqs/sigma;
 * out [m, 2*nf_half] = scale * [cos(x coeff^T) | sin(x coeff^T)]. */
int bsig_rff_features(const float* x, int64_t ldx, const int64_t* x_rows,
                      const float* coeff, float* out,
                      int64_t m, int64_t d, int64_t nf_half, float scale, int engine,
                      void* ws, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------ mixture-density head + NLL
 * z [b, n_head] is the concatenated head output, columns
 *   [0,K) pi logits | [K, K+PK) mu (col p*K+k) | [K+PK, K+2PK) log-diag |
 *   [K+2PK, K+2PK+LK) strict-lower entries (col l*K+k, np.tril_indices order)
 * with L = P(P-1)/2 if full covariance else 0 (mdnn.py:109-119).
 * noise [b, P, K]: the uniforms of torch.rand_like at mdnn.py:116. */

/* MDNN.forward epilogue (mdnn.py:109-119): weights [b,K] = renorm(clamp(softmax)),
 * l_d [b,P,K] = exp(z_d) + noise * 1e-5 * mean(exp(z_d)).  mu and L are views of z.
 * ws >= bsig_mdn_ws_bytes(b). flag |= 1 if any of weights/mu/l_d/L is non-finite. */
int64_t bsig_mdn_ws_bytes(int64_t b);
int bsig_mdn_head_fwd(const float* z, const float* noise, float* weights, float* l_d,
                      int64_t b, int64_t p, int64_t k, int full_cov,
                      void* ws, int64_t ws_bytes, int* flag, void* stream);
/* backward of the epilogue: given d_weights [b,K], d_mu [b,P,K], d_ld [b,P,K],
 * d_low [b,L,K] (nullable) -> dz [b, n_head]; includes the non-detached eps term. */
int bsig_mdn_head_bwd(const float* z, const float* noise, const float* weights,
                      const float* d_weights, const float* d_mu, const float* d_ld,
                      const float* d_low, float* dz,
                      int64_t b, int64_t p, int64_t k, int full_cov,
                      void* ws, int64_t ws_bytes, void* stream);

/* MDNN.mdn_loss_fn (mdnn.py:127-178) on explicit (weights, mu, l_d, low):
 * loss[0] = -mean_b logsumexp_k( clamp(log N(y; mu_k, L_k L_k^T), +-1e5)
 *                                + log clamp(w_k, 1e-5, 1) ).
 * y_rows (nullable) gathers target rows.  mu/l_d/low are addressed with an
 * explicit row stride so that they may alias columns of z. */
int bsig_mog_nll_fwd(const float* weights, const float* mu, int64_t ld_mu,
                     const float* l_d, int64_t ld_ld, const float* low, int64_t ld_low,
                     const float* y, const int64_t* y_rows, float* loss,
                     int64_t b, int64_t p, int64_t k,
                     void* ws, int64_t ws_bytes, int* flag, void* stream);
/* gradients of grad_scale[0] * loss wrt weights, mu, l_d, low (dense outputs
 * [b,K], [b,P,K], [b,P,K], [b,L,K]); grad_scale is a device scalar. */
int bsig_mog_nll_bwd(const float* weights, const float* mu, int64_t ld_mu,
                     const float* l_d, int64_t ld_ld, const float* low, int64_t ld_low,
                     const float* y, const int64_t* y_rows, const float* grad_scale,
                     float* d_weights, float* d_mu, float* d_ld, float* d_low,
                     int64_t b, int64_t p, int64_t k,
                     void* ws, int64_t ws_bytes, void* stream);

/* Fused training form: head epilogue + NLL forward + full backward in one pass
 * over z: loss[0] and dz [b, n_head] = d loss / d z (dz nullable: loss only).
 * This is what MDNN.run_training's captured step uses. */
int bsig_mdn_nll_fused(const float* z, const float* noise, const float* y,
                       const int64_t* y_rows, float* loss, float* dz,
                       int64_t b, int64_t p, int64_t k, int full_cov,
                       void* ws, int64_t ws_bytes, int* flag, void* stream);

/* ------------------------------------------------------------------- optimiser
 * torch.optim.Adam defaults (mdnn.py:203,234) over one flat fp32 buffer. */
int bsig_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t count, int64_t step, float lr, float beta1, float beta2,
                   float eps, float grad_scale, void* stream);
/* Data-parallel exchange fused with the optimiser (no counterpart in the single-process
 * reference; the exchange step of SURVEY 8.e): one-shot all-reduce of the flat gradient
 * buffer over NVLink peer memory + Adam in ONE kernel.  peer_grads / peer_flags are HOST
 * arrays of `world` device pointers (this rank's own buffers at index `rank`, the others
 * opened from IPC handles); ctrl is this rank's 2-word control block.  All buffers come
 * from bsig_p2p_alloc (zero-initialised).  Graph-capturable and replayable: the epoch
 * that orders the ranks lives in device memory and is advanced by the kernel. */
int bsig_p2p_alloc(void** ptr, int64_t bytes, unsigned char* handle64);   /* host call */
int bsig_p2p_open(const unsigned char* handle64, void** ptr);             /* host call */
int bsig_p2p_close(void* ptr);
int bsig_p2p_read(const void* dev_ptr, void* host_ptr, int64_t bytes);   /* Watchdog of the fused exchange: a peer that never publishes its epoch (dead rank) makes
 * the waiting kernel give up after this many milliseconds (default 20000, env
 * BSIG_P2P_TIMEOUT_MS), store 1 + peer index in the sticky error word ctrl[2] (read it with
 * bsig_p2p_read(ctrl + 8 bytes, ...)) and let every later launch skip its wait. */
int bsig_p2p_set_timeout_ms(int64_t ms);
/* Load the exchange kernel's code (no launch): lets a CUDA-graph capture contain the kernel's
 * first launch, so that the engine's pre-capture warm-up never has to rendezvous with peers. */
int bsig_p2p_preload(void);
/* blocking D2H copy */
int bsig_p2p_free(void* ptr);
int bsig_adam_allreduce_step(float* param, const void* const* peer_grads,
                             void* const* peer_flags, void* ctrl, int rank, int world,
                             float* exp_avg, float* exp_avg_sq, int64_t count, int64_t step,
                             float lr, float beta1, float beta2, float eps, void* stream);
/* ---------------------------------------------------------------- persistent training
 * ALL Adam updates [step0, step1) of MDNN.run_training's loop (models/mdnn.py:217-234:
 * np.random.randint rows -> gather -> forward (mdnn.py:89-125) -> mdn_loss_fn (mdnn.py:127-178)
 * -> backward -> torch.optim.Adam.step) in ONE launch of one 16-CTA thread-block cluster that
 * keeps the model resident in shared memory, partitioned by output column (csrc/
 * train_persistent.cu).  Dense layers = hidden tanh layers followed by the concatenated heads
 * [pi | mu | Diag | Lower] (1..3 layers; MDRFF passes its precomputed random Fourier features
 * as x and n_layers = 1).  All pointers are device pointers.
 *   x [n_train][ldx] (ldx % 4 == 0, pad columns zero, 16-byte aligned), y [n_train][p]
 *   normalised targets, idx [n_updates][batch] minibatch rows (mdnn.py:221), noise
 *   [n_updates][batch][p][k] eps-noise uniforms (mdnn.py:115-116), params / exp_avg /
 *   exp_avg_sq flat fp32 buffers (w_off / b_off: float offsets of each layer's weight
 *   [out,in] and bias), scratch >= the size reported by the query, loss_buf + loss_slot
 *   [n_updates] (slot of update u's minibatch loss, -1 = not logged), adam_coef
 *   [n_updates][2] = {lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)} for t = u + 1 (torch's host
 *   doubles, rounded to fp32), flag: |= 1 on a non-finite value (mdnn.py:120-124,172-174).
 * bsig_train_persistent_query returns 0 and the scratch / shared-memory needs if the shape is
 * inside the kernel's envelope, non-zero (reason in bsig_last_error) otherwise. */
typedef struct bsig_tp_desc {
  int32_t n_layers;
  int32_t in_dim[3];
  int32_t out_dim[3];
  int32_t batch, p, k, full_cov;
  int64_t w_off[3];
  int64_t b_off[3];
  const float* x;
  int64_t ldx;
  const float* y;
  const int64_t* idx;
  const float* noise;
  float* params;
  float* exp_avg;
  float* exp_avg_sq;
  float* scratch;
  int64_t scratch_floats;
  float* loss_buf;
  const int32_t* loss_slot;
  const float* adam_coef;
  int32_t* flag;
  float beta1, beta2, eps;
  int64_t* prof;   /* nullable profiling aid: 32 per-phase cycle counters of CTA 0, accumulated */
} bsig_tp_desc;
int bsig_train_persistent_query(const bsig_tp_desc* desc, int64_t* scratch_floats,
                                int64_t* smem_bytes);
int bsig_train_persistent(const bsig_tp_desc* desc, int64_t step0, int64_t step1, void* stream);

/* Opt-in (BSIG_FUSED_WGRAD=1 in the Python engine; single GPU, minibatch <= 128): the three
 * weight-gradient GEMMs of a two-hidden-layer MDNN update (autograd of mdnn.py:108-119) and
 * torch.optim.Adam.step (mdnn.py:234) in ONE launch: one 32 x 32 tile of one dW_l = dY_l^T X_l
 * per CTA, Adam applied to those parameters in the epilogue (no gradient buffer).  Validated
 * on hardware (round 2); measured equal to the default three GEMMs + Adam (profiles/r2/).
 * Layer l: dy_l [b, n_l], x_l [*, k_l] (x_0 gathered with x0_rows, pitch ld_x0), weight /
 * bias at float offsets w_off_l / b_off_l of param. */
int bsig_wgrad3_adam_step(const float* dy0, const float* x0, int64_t ld_x0, const int64_t* x0_rows,
                          int64_t n0, int64_t k0, int64_t w_off0, int64_t b_off0,
                          const float* dy1, const float* x1, int64_t n1, int64_t k1,
                          int64_t w_off1, int64_t b_off1,
                          const float* dy2, const float* x2, int64_t n2, int64_t k2,
                          int64_t w_off2, int64_t b_off2,
                          float* param, float* exp_avg, float* exp_avg_sq, int64_t b,
                          int64_t step, float lr, float beta1, float beta2, float eps,
                          void* stream);
/* out[i,:] = src[rows[i],:]  (x_train[ids], mdnn.py:222) */
int bsig_gather_rows(const float* src, int64_t ld_src, const int64_t* rows, float* out,
                     int64_t n_rows, int64_t width, void* stream);
/* flag |= 1 if any of x[0..count) is not finite (torch.isfinite(..).all() asserts) */
int bsig_finite_flag(const float* x, int64_t count, int* flag, void* stream);
/* y = (x - lows) / (highs - lows) row-wise (MDNN.normalize_samples, mdnn.py:245-248) */
int bsig_normalize_rows(const float* x, const float* lows, const float* highs, float* y,
                        int64_t n_rows, int64_t width, void* stream);

/* --------------------------------------------------------------------- posterior
 * MDNN.predict_MoGs de-normalisation (mdnn.py:264-288, with L[pt,:,k]):
 * a [r,K] ; means [r,K,P] = mu*rng + lows ; packed [r,K,P+L] =
 * [ rng*l_d | rng[row(l)]*low ] -- the `Ls` argument of pdf.MoG. lows/highs nullable. */
int bsig_mog_denorm(const float* weights, const float* mu, int64_t ld_mu,
                    const float* l_d, int64_t ld_ld, const float* low, int64_t ld_low,
                    const float* lows, const float* highs,
                    float* a_out, float* means_out, float* packed_out,
                    int64_t r, int64_t p, int64_t k, void* stream);

/* pdf.discrete_sample + MoG.gen + Gaussian.gen (utils/pdf.py:61-76,465-472,296-300).
 * a [K] mixture weights (float32 if a_is_f32 else float64, the dtype the
 * reference compares in), u [n] float64 uniforms, z [n,P] float64 normals,
 * means [K,P] and cmats [K,P,P] float64 (C = L^T).  comp_idx [n] int32 =
 * #{j: u > cumsum(a[:-1])_j} (bit exact); counts [K] int32 (must be zeroed
 * by the caller); samples [n,P] float64 GROUPED BY COMPONENT: block k holds
 * z_rows(block k) @ C_k + m_k. */
int bsig_mog_sample(const void* a, int a_is_f32, const double* u, const double* z,
                    const double* means, const double* cmats,
                    int32_t* comp_idx, int32_t* counts, double* samples,
                    int64_t n, int64_t p, int64_t k, void* stream);
/* Device-RNG variant for throughput (no host draws): Philox4x32-10 keyed by
 * (seed, row); fp32 parameters and fp32 output in DRAW order. */
int bsig_mog_sample_philox(const float* a, const float* means, const float* cmats,
                           int32_t* comp_idx, float* samples, uint64_t seed,
                           int64_t n, int64_t p, int64_t k, void* stream);
/* One clipped posterior draw per environment in ONE launch -- the batched form of
 * ParamsGenerator.sample (sim/params_generator.py:115-118: distr.gen(n_samples=1)[0] then
 * np.clip(.., lows, highs)), which the reference calls once per environment reset
 * (sim/apply_randomizations.py:154-158).  Environment e uses uniform u[e] and normals
 * z[e,:]; samples [n,P] float64 in DRAW order, component choice bit-exact for identical
 * uniforms; lows/highs [P] nullable (no clip); comp_idx [n] nullable. */
int bsig_mog_sample_envs(const void* a, int a_is_f32, const double* u, const double* z,
                         const double* means, const double* cmats, const double* lows,
                         const double* highs, double* samples, int32_t* comp_idx,
                         int64_t n, int64_t p, int64_t k, void* stream);
/* Same with the device RNG of bsig_mog_sample_philox (fp32). */
int bsig_mog_sample_envs_philox(const float* a, const float* means, const float* cmats,
                                const float* lows, const float* highs, int32_t* comp_idx,
                                float* samples, uint64_t seed, int64_t n, int64_t p,
                                int64_t k, void* stream);

/* MoG.eval, joint density (utils/pdf.py:474-491, 328-332):
 * x [m,P] (float32 if x_is_f32 else float64); a [K], log_a [K] (np.log(a) taken in
 * a's own dtype, as the reference does), means [K,P], precs [K,P,P], logdet_p [K]
 * float64; out [m] float64 = logsumexp_k(lp_k + log_a_k) if log_space else
 * sum_k a_k exp(lp_k). */
int bsig_mog_logpdf(const void* x, int x_is_f32, const double* a, const double* log_a,
                    const double* means, const double* precs, const double* logdet_p, double* out,
                    int64_t m, int64_t p, int64_t k, int log_space, void* stream);

/* Every pairwise 2-D marginal density a posterior plot evaluates, in one launch: the batched
 * form of utils/plot.py:38-44 (get_2d_posterior_data: posterior.eval(np.mgrid grid, ii=dims,
 * log=False) -> MoG.eval / Gaussian.eval marginal branch, utils/pdf.py:334-339, 474-491),
 * which the reference runs once per parameter pair (P(P-1)/2 times 100 x 100 points).
 * a, log_a [K]; params [n_pairs][K][6] = mean (2), precision p00 p01 p11 and log det
 * precision of each component's jittered 2 x 2 marginal (host, float64); lims [n_pairs][4]
 * = xmin, xmax, ymin, ymax; out [n_pairs][nbins][nbins] float64 in np.mgrid order. */
int bsig_mog_marginal_grid(const double* a, const double* log_a, const double* params,
                           const double* lims, double* out, int64_t n_pairs, int64_t k,
                           int64_t nbins, int log_space, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* BSIG_H_ */
