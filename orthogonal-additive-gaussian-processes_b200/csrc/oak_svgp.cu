// Whitened SVGP with a diagonal q(u) and a Bernoulli likelihood around the Kuf tiles
// (SURVEY.md section 8(f) #4).  The reference builds gpflow.models.SVGP(kernel=OAK, likelihood=
// Bernoulli(invlink=inv_logit), whiten=True, q_diag=True) and trains it full-batch with BFGS
// (examples/uci/uci_classification_train.py:108-124).  gpflow 2.2.1 is not vendored under /root/reference;
// its published algorithm is restated here:
//   conditional (gpflow/conditionals/util.py base_conditional, white=True, diagonal q_sqrt):
//     A = L^-1 Kuf,  mean_i = A[:, i] . q_mu,  var_i = K_diag_i - sum_m A_mi^2 + sum_m (q_sqrt_m A_mi)^2
//   likelihood (gpflow/likelihoods/scalar_discrete.py Bernoulli + gpflow/quadrature NDiagGHQuadrature):
//     E_q[log p(y|f)] = sum_k w_k log p(y | mean + sqrt(var) x_k),  (x_k, w_k) = hermgauss nodes scaled by
//     sqrt(2) and 1/sqrt(pi) (passed in by the host, which takes them from numpy exactly as gpflow does),
//     log p(y|f) = log(y == 1 ? p : 1 - p),  p = invlink(f);
//     predict_log_density = logsumexp_k(log p(y | f_k) + log w_k).
// A (M x n, from the fused Kuf tiles and one triangular solve) is the only large operand; each kernel
// reads it once.
#include <cmath>

#include "oak_common.cuh"

namespace oak {

constexpr int kMaxGH = 64;

// p = invlink(f) and dp/df.  OAK_LINK_LOGIT: sigmoid(f) (1 - 2 j) + j (the reference's inv_logit,
// uci_classification_train.py:43-45); OAK_LINK_PROBIT: gpflow's inv_probit, 0.5 (1 + erf(f / sqrt 2)) (1 - 2 j) + j.
__device__ __forceinline__ double invlink(double f, int link, double jitter, double* dp) {
  const double span = 1.0 - 2.0 * jitter;
  if (link == OAK_LINK_LOGIT) {
    const double e = exp(-fabs(f));
    const double s = f >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
    *dp = s * (1.0 - s) * span;
    return s * span + jitter;
  }
  *dp = 0.3989422804014327 * exp(-0.5 * f * f) * span;
  return 0.5 * (1.0 + erf(f * 0.7071067811865476)) * span + jitter;
}

__global__ void __launch_bounds__(256) bernoulli_quadrature_kernel(
    const double* __restrict__ mean, const double* __restrict__ var, const double* __restrict__ y, int64_t n, int link,
    double jitter, const double* __restrict__ gh_x, const double* __restrict__ gh_w, int n_gh,
    double* __restrict__ varexp, double* __restrict__ gmean, double* __restrict__ gvar, double* __restrict__ logdens) {
  __shared__ double sx[kMaxGH], sw[kMaxGH], slw[kMaxGH];
  if ((int)threadIdx.x < n_gh) {
    sx[threadIdx.x] = gh_x[threadIdx.x];
    sw[threadIdx.x] = gh_w[threadIdx.x];
    slw[threadIdx.x] = log(gh_w[threadIdx.x]);
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double mu = mean[i], sd = sqrt(var[i]);
  const bool one = y[i] == 1.0;
  double ve = 0.0, gm = 0.0, gs = 0.0;  // gs = d/d sd
  double lp[kMaxGH];
  double top = -INFINITY;
  for (int k = 0; k < n_gh; ++k) {
    const double f = fma(sd, sx[k], mu);
    double dp;
    const double p = invlink(f, link, jitter, &dp);
    const double q = one ? p : 1.0 - p;
    const double l = log(q);
    const double dl = (one ? dp : -dp) / q;
    ve = fma(sw[k], l, ve);
    gm = fma(sw[k], dl, gm);
    gs = fma(sw[k] * sx[k], dl, gs);
    lp[k] = l + slw[k];
    top = fmax(top, lp[k]);
  }
  if (varexp) varexp[i] = ve;
  if (gmean) gmean[i] = gm;
  if (gvar) gvar[i] = gs / (2.0 * sd);  // d sd / d var
  if (logdens) {
    double acc = 0.0;
    for (int k = 0; k < n_gh; ++k) acc += exp(lp[k] - top);
    logdens[i] = top + log(acc);
  }
}

// mean_i = sum_m A_mi q_mu_m,  var_i = kdiag_i - sum_m A_mi^2 (1 - q_sqrt_m^2): one thread per column
__global__ void __launch_bounds__(256) svgp_moments_kernel(const double* __restrict__ A, int64_t lda, int m, int64_t n,
                                                           const double* __restrict__ q_mu,
                                                           const double* __restrict__ q_sqrt,
                                                           const double* __restrict__ kdiag, double* __restrict__ mean,
                                                           double* __restrict__ var) {
  extern __shared__ double sh[];  // [m] q_mu | [m] 1 - q_sqrt^2
  double* s_mu = sh;
  double* s_c = sh + m;
  for (int r = threadIdx.x; r < m; r += 256) {
    s_mu[r] = q_mu[r];
    s_c[r] = 1.0 - q_sqrt[r] * q_sqrt[r];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double mu = 0.0, acc = 0.0;
  for (int r = 0; r < m; ++r) {
    const double a = A[(int64_t)r * lda + i];
    mu = fma(a, s_mu[r], mu);
    acc = fma(a * a, s_c[r], acc);
  }
  mean[i] = mu;
  var[i] = kdiag[i] - acc;
}

// Abar_mi = q_mu_m gmean_i - 2 (1 - q_sqrt_m^2) A_mi gvar_i;  gq_sqrt_m += 2 q_sqrt_m sum_i gvar_i A_mi^2.
// One block per row: fixed-order block reduction, no atomics.
__global__ void __launch_bounds__(256) svgp_moments_backward_kernel(
    const double* __restrict__ A, int64_t lda, int64_t n, const double* __restrict__ q_mu,
    const double* __restrict__ q_sqrt, const double* __restrict__ gmean, const double* __restrict__ gvar,
    double* __restrict__ Abar, int64_t ldb, double* __restrict__ gq_sqrt) {
  __shared__ double red[8];
  const int r = blockIdx.x;
  const double qm = q_mu[r], qs = q_sqrt[r];
  const double c2 = 2.0 * (1.0 - qs * qs);
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const double a = A[(int64_t)r * lda + i];
    const double gv = gvar[i];
    Abar[(int64_t)r * ldb + i] = fma(qm, gmean[i], -c2 * a * gv);
    acc = fma(gv, a * a, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w];
    gq_sqrt[r] += 2.0 * qs * v;
  }
}

}  // namespace oak

using namespace oak;

extern "C" int oak_bernoulli_quadrature_f64(const double* d_mean, const double* d_var, const double* d_y, int64_t n,
                                            int32_t link, double jitter, const double* d_gh_x, const double* d_gh_w,
                                            int32_t n_gh, double* d_varexp, double* d_gmean, double* d_gvar,
                                            double* d_logdensity, void* stream_) {
  OAK_REQUIRE(d_mean && d_var && d_y && d_gh_x && d_gh_w, "oak_bernoulli_quadrature_f64: null argument");
  OAK_REQUIRE(n_gh >= 1 && n_gh <= kMaxGH, "oak_bernoulli_quadrature_f64: 1 <= n_gh <= 64");
  OAK_REQUIRE(link == OAK_LINK_LOGIT || link == OAK_LINK_PROBIT, "oak_bernoulli_quadrature_f64: unknown link");
  if (n <= 0) return 0;
  bernoulli_quadrature_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      d_mean, d_var, d_y, n, link, jitter, d_gh_x, d_gh_w, n_gh, d_varexp, d_gmean, d_gvar, d_logdensity);
  OAK_LAUNCHED();
  return 0;
}

extern "C" int oak_svgp_moments_f64(const double* d_A, int64_t lda, int32_t m, int64_t n, const double* d_q_mu,
                                    const double* d_q_sqrt, const double* d_kdiag, double* d_mean, double* d_var,
                                    void* stream_) {
  OAK_REQUIRE(d_A && d_q_mu && d_q_sqrt && d_kdiag && d_mean && d_var, "oak_svgp_moments_f64: null argument");
  OAK_REQUIRE(m >= 1 && m <= 2048 && lda >= n, "oak_svgp_moments_f64: 1 <= m <= 2048, lda >= n");
  if (n <= 0) return 0;
  svgp_moments_kernel<<<(unsigned)((n + 255) / 256), 256, 2 * (size_t)m * sizeof(double), (cudaStream_t)stream_>>>(
      d_A, lda, m, n, d_q_mu, d_q_sqrt, d_kdiag, d_mean, d_var);
  OAK_LAUNCHED();
  return 0;
}

extern "C" int oak_svgp_moments_backward_f64(const double* d_A, int64_t lda, int32_t m, int64_t n,
                                             const double* d_q_mu, const double* d_q_sqrt, const double* d_gmean,
                                             const double* d_gvar, double* d_Abar, int64_t ldb, double* d_gq_sqrt,
                                             void* stream_) {
  OAK_REQUIRE(d_A && d_q_mu && d_q_sqrt && d_gmean && d_gvar && d_Abar && d_gq_sqrt,
              "oak_svgp_moments_backward_f64: null argument");
  OAK_REQUIRE(m >= 1 && lda >= n && ldb >= n, "oak_svgp_moments_backward_f64: bad shape");
  if (n <= 0) return 0;
  svgp_moments_backward_kernel<<<(unsigned)m, 256, 0, (cudaStream_t)stream_>>>(d_A, lda, n, d_q_mu, d_q_sqrt, d_gmean,
                                                                              d_gvar, d_Abar, ldb, d_gq_sqrt);
  OAK_LAUNCHED();
  return 0;
}
