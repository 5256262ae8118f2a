// bsig_linear_* / bsig_rff_features: shape the three layer GEMMs (forward,
// dgrad, wgrad) and the RFF projection and hand them to a GEMM engine.
#include "common.cuh"
#include "gemm.cuh"

using namespace bsig;

static int run_gemm(const GemmArgs& g, int engine, void* ws, int64_t ws_bytes, cudaStream_t st) {
  // tensor-core engines take the GEMMs they can express (K-contiguous operands,
  // 16-byte aligned rows, no gather); everything else runs on the SIMT engines
  // (forward, dgrad and wgrad alike: >= 2^26 multiply-adds, e.g. every layer of a
  // 4096-row minibatch at the Cartpole widths and the ShadowHand first layer at 100 rows)
  if (engine == BSIG_GEMM_AUTO)   // tensor cores once the problem is big enough to feed them
    engine = ((int64_t)g.M * g.N * g.K >= (1ll << 26)) ? BSIG_GEMM_TC_TF32X3 : BSIG_GEMM_SIMT;
  if ((engine == BSIG_GEMM_TC_TF32 || engine == BSIG_GEMM_TC_TF32X3) && gemm_tc_applicable(g) &&
      ws != nullptr && ws_bytes >= gemm_tc_ws_bytes(g.M, g.N, g.K))
    return gemm_tc(g, engine == BSIG_GEMM_TC_TF32X3, ws, ws_bytes, st);
  if (gemm_small_applicable(g)) return gemm_small(g, st);
  return gemm_simt(g, ws, ws_bytes, st);
}

extern "C" int64_t bsig_linear_ws_bytes(int64_t m, int64_t n, int64_t k) {
  // the same scratch serves forward (m,n,k), dgrad (m,k,n) and wgrad (n,k,m)
  int64_t a = gemm_simt_ws_bytes(m, n, k);
  int64_t b = gemm_simt_ws_bytes(m, k, n);
  int64_t c = gemm_simt_ws_bytes(n, k, m);
  int64_t r = a > b ? a : b;
  r = r > c ? r : c;
  // tensor-core engine: staging copies of operands TMA cannot address + split-K partials
  const int64_t ta = gemm_tc_ws_bytes(m, n, k), tb = gemm_tc_ws_bytes(m, k, n),
                tcc = gemm_tc_ws_bytes(n, k, m);
  r = r > ta ? r : ta;
  r = r > tb ? r : tb;
  r = r > tcc ? r : tcc;
  return r + 256;
}

extern "C" int bsig_linear_fwd(const float* x, int64_t ldx, const int64_t* x_rows, const float* w,
                               const float* b, float* y, int64_t m, int64_t n, int64_t k, int act,
                               int engine, void* ws, int64_t ws_bytes, void* stream) {
  BSIG_REQUIRE(m >= 1 && n >= 1 && k >= 1, "linear_fwd: empty problem");
  BSIG_REQUIRE(act == BSIG_ACT_NONE || act == BSIG_ACT_TANH, "linear_fwd: unknown activation");
  GemmArgs g = gemm_args_zero();
  g.A = x; g.a_si = ldx; g.a_sr = 1; g.a_rows = x_rows;
  g.B = w; g.b_sr = 1; g.b_sj = k;          // B(r,j) = w[j,r]
  g.C = y; g.ldc = n; g.M = (int)m; g.N = (int)n; g.K = (int)k;
  g.bias = b;
  g.epi = b == nullptr ? EPI_STORE : (act == BSIG_ACT_TANH ? EPI_BIAS_TANH : EPI_BIAS);
  BSIG_REQUIRE(!(b == nullptr && act == BSIG_ACT_TANH), "linear_fwd: tanh needs a bias");
  return run_gemm(g, engine, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int bsig_linear_dgrad(const float* dy, const float* w, const float* h_prev, float* dx,
                                 int64_t m, int64_t n, int64_t k, int act_prev, int engine,
                                 void* ws, int64_t ws_bytes, void* stream) {
  BSIG_REQUIRE(m >= 1 && n >= 1 && k >= 1, "linear_dgrad: empty problem");
  GemmArgs g = gemm_args_zero();
  g.A = dy; g.a_si = n; g.a_sr = 1;         // A(i,r) = dy[i,r], r over n
  g.B = w; g.b_sr = k; g.b_sj = 1;          // B(r,j) = w[r,j]
  g.C = dx; g.ldc = k; g.M = (int)m; g.N = (int)k; g.K = (int)n;
  if (act_prev == BSIG_ACT_TANH) {
    BSIG_REQUIRE(h_prev != nullptr, "linear_dgrad: h_prev required for tanh");
    g.epi = EPI_MUL_DTANH; g.aux = h_prev; g.ld_aux = k;
  }
  return run_gemm(g, engine, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int bsig_linear_wgrad(const float* dy, const float* x, int64_t ldx,
                                 const int64_t* x_rows, float* dw, float* db, int64_t m, int64_t n,
                                 int64_t k, int engine, void* ws, int64_t ws_bytes, void* stream) {
  BSIG_REQUIRE(m >= 1 && n >= 1 && k >= 1, "linear_wgrad: empty problem");
  GemmArgs g = gemm_args_zero();
  g.A = dy; g.a_si = 1; g.a_sr = n;         // A(i,r) = dy[r,i]
  g.B = x; g.b_sr = ldx; g.b_sj = 1; g.b_rows = x_rows;   // B(r,j) = x[rows[r], j]
  g.C = dw; g.ldc = k; g.M = (int)n; g.N = (int)k; g.K = (int)m;
  // the small engine forms the bias gradient on the side; the other engines use colsum
  const bool wants_tc = engine == BSIG_GEMM_TC_TF32 || engine == BSIG_GEMM_TC_TF32X3 ||
                        (engine == BSIG_GEMM_AUTO && (int64_t)g.M * g.N * g.K >= (1ll << 26));
  const bool fused_bias = db != nullptr && gemm_small_applicable(g) && !wants_tc;
  if (fused_bias) g.rowsum = db;            // db[i] = sum_r dy[r,i] rides along in the GEMM
  if (run_gemm(g, engine, ws, ws_bytes, (cudaStream_t)stream)) return 1;
  // (after the GEMM in stream order: its workspace is free again)
  if (db != nullptr && !fused_bias) return colsum(dy, db, m, n, ws, ws_bytes, (cudaStream_t)stream);
  return 0;
}

extern "C" int bsig_rff_features(const float* x, int64_t ldx, const int64_t* x_rows,
                                 const float* coeff, float* out, int64_t m, int64_t d,
                                 int64_t nf_half, float scale, int engine, void* ws,
                                 int64_t ws_bytes, void* stream) {
  BSIG_REQUIRE(m >= 1 && d >= 1 && nf_half >= 1, "rff_features: empty problem");
  GemmArgs g = gemm_args_zero();
  g.A = x; g.a_si = ldx; g.a_sr = 1; g.a_rows = x_rows;
  g.B = coeff; g.b_sr = 1; g.b_sj = d;
  g.C = out; g.ldc = 2 * nf_half; g.M = (int)m; g.N = (int)nf_half; g.K = (int)d;
  g.epi = EPI_SINCOS; g.scale = scale;
  return run_gemm(g, engine, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int bsig_linear_colsum(const float* dy, float* db, int64_t m, int64_t n, void* stream) {
  BSIG_REQUIRE(m >= 1 && n >= 1, "linear_colsum: empty problem");
  return colsum(dy, db, m, n, nullptr, 0, (cudaStream_t)stream);
}

// Last kernel of a single-GPU backward pass: dW = dy^T x of the FIRST layer (weight at the start
// of the flat parameter buffer, bias right behind it) with torch.optim.Adam applied in the
// epilogue -- to that weight and bias from the accumulators, and to every other parameter from
// the flat gradient buffer by the CTAs of the launch that have nothing else left to do.  One
// launch instead of weight gradient + Adam on the critical path of an update.
extern "C" int bsig_linear_wgrad_adam(const float* dy, const float* x, int64_t ldx,
                                      const int64_t* x_rows, int64_t m, int64_t n, int64_t k,
                                      float* param, const float* grad, float* exp_avg,
                                      float* exp_avg_sq, int64_t n_params, int64_t step, float lr,
                                      float beta1, float beta2, float eps, void* stream) {
  BSIG_REQUIRE(m >= 1 && n >= 1 && k >= 1 && step >= 1, "linear_wgrad_adam: bad sizes");
  BSIG_REQUIRE(n_params >= n * k + n, "linear_wgrad_adam: parameter buffer smaller than the layer");
  GemmArgs g = gemm_args_zero();
  g.A = dy; g.a_si = 1; g.a_sr = n;         // A(i,r) = dy[r,i]
  g.B = x; g.b_sr = ldx; g.b_sj = 1; g.b_rows = x_rows;
  g.C = param; g.ldc = k; g.M = (int)n; g.N = (int)k; g.K = (int)m;
  g.epi = EPI_ADAM;
  g.rowsum = param;                          // (non-null: request the row sums = bias gradient)
  BSIG_REQUIRE(gemm_small_applicable(g), "linear_wgrad_adam: layer too large for the fused form");
  g.ad_p = param; g.ad_m = exp_avg; g.ad_v = exp_avg_sq; g.ad_g = grad;
  g.ad_b_off = n * k;
  g.ad_tail_off = n * k + n;
  g.ad_tail_cnt = n_params - g.ad_tail_off;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  g.ad_ob1 = 1.0f - beta1; g.ad_b2 = beta2; g.ad_ob2 = 1.0f - beta2;
  g.ad_step = (float)((double)lr / bc1);
  g.ad_ibc2 = (float)(1.0 / sqrt(bc2));
  g.ad_eps = eps;
  return gemm_small(g, (cudaStream_t)stream);
}
