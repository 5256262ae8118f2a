// Internal GEMM plumbing shared by the SIMT and tcgen05 engines.
#pragma once
#include "common.cuh"

namespace bsig {

enum Epilogue {
  EPI_STORE = 0,      // C = acc
  EPI_BIAS = 1,       // C = acc + bias[j]
  EPI_BIAS_TANH = 2,  // C = tanh(acc + bias[j])
  EPI_MUL_DTANH = 3,  // C = acc * (1 - aux[i,j]^2)
  EPI_SINCOS = 4,     // C[i,j] = scale*cos(acc), C[i,N+j] = scale*sin(acc)
  EPI_ADAM = 5        // small engine, weight gradients: acc is the gradient of C[i,j] = a
                      // PARAMETER; Adam is applied to it (and to the rest of the model) in place
};

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int64_t a_si, a_sr;        // A(i,r) strides
  int64_t b_sr, b_sj;        // B(r,j) strides
  int64_t ldc;
  const int64_t* a_rows;     // optional gather of A's i index
  const int64_t* b_rows;     // optional gather of B's r index
  int epi;
  const float* bias;
  const float* aux;
  int64_t ld_aux;
  float scale;
  float* rowsum;             // optional: rowsum[i] = sum_r A(i,r) (small engine only)
  // EPI_ADAM: C = the layer's weight inside the flat parameter buffer ad_p (at offset 0), its
  // bias at ad_b_off; ad_m / ad_v the flat Adam moments, ad_g the flat gradient buffer of
  // every OTHER parameter [ad_tail_off, ad_tail_off + ad_tail_cnt)
  float* ad_p; float* ad_m; float* ad_v; const float* ad_g;
  int64_t ad_b_off, ad_tail_off, ad_tail_cnt;
  float ad_ob1, ad_b2, ad_ob2, ad_step, ad_ibc2, ad_eps;
  // filled by the launcher
  int k_per_split;
  float* partial;
};

inline GemmArgs gemm_args_zero() {
  GemmArgs g;
  g.A = g.B = nullptr; g.C = nullptr; g.M = g.N = g.K = 0;
  g.a_si = g.a_sr = g.b_sr = g.b_sj = g.ldc = 0;
  g.a_rows = g.b_rows = nullptr; g.epi = EPI_STORE; g.bias = g.aux = nullptr;
  g.ld_aux = 0; g.scale = 1.f; g.rowsum = nullptr; g.k_per_split = 0; g.partial = nullptr;
  g.ad_p = g.ad_m = g.ad_v = nullptr; g.ad_g = nullptr; g.ad_b_off = g.ad_tail_off = g.ad_tail_cnt = 0;
  g.ad_ob1 = g.ad_b2 = g.ad_ob2 = g.ad_step = g.ad_ibc2 = g.ad_eps = 0.f;
  return g;
}

int gemm_simt(GemmArgs g, void* ws, int64_t ws_bytes, cudaStream_t st);
int64_t gemm_simt_ws_bytes(int64_t M, int64_t N, int64_t K);
int colsum(const float* dy, float* db, int64_t M, int64_t N, void* ws, int64_t ws_bytes,
           cudaStream_t st);
bool gemm_small_applicable(const GemmArgs& g);
bool gemm_tc_applicable(const GemmArgs& g);
int gemm_tc(const GemmArgs& g, bool x3, void* ws, int64_t ws_bytes, cudaStream_t st);
int64_t gemm_tc_ws_bytes(int64_t M, int64_t N, int64_t K);
int gemm_splitk_reduce(const GemmArgs& g, int splits, cudaStream_t st);
int gemm_small(GemmArgs g, cudaStream_t st);

}  // namespace bsig
