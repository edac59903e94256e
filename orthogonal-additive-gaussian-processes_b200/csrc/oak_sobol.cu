// Sobol-index building blocks (oak/utils.py:116-165, 221-435).
//
//   oak_sobol_L_f64          L_d[i,j] = int k~_d(x_i,s) k~_d(s,x_j) p_d(s) ds at the conditioning
//                            points, with unit order-variance, for every sub-kernel kind
//   oak_sobol_quadforms_f64  out[c] = scale[c] * alpha^T (prod_{d in S_c} L_d) alpha for all
//                            additive components in one launch (the reference recomputes L_d for
//                            every subset it appears in and loops in Python, utils.py:369-432)
#include <cublas_v2.h>

#include <cstring>

#include "oak_common.cuh"

namespace oak {

// ---- Gaussian-measure closed forms, eq. (44)-(47) of the paper (utils.py:116-165), sigma = 1
__device__ __forceinline__ double sobol_f1(double x, double y, double l, double delta, double mu) {
  const double l2 = l * l, d2 = delta * delta;
  const double a = x - y, b = mu - (x + y) * 0.5;
  return l / sqrt(l2 + 2.0 * d2) * exp(-(a * a) / (4.0 * l2)) * exp(-(b * b) / (2.0 * d2 + l2));
}

__device__ __forceinline__ double sobol_f2(double x, double y, double l, double delta, double mu) {
  const double l2 = l * l, d2 = delta * delta;
  const double M = 1.0 / l2 + 1.0 / (l2 + d2);
  const double m = 1.0 / M * (mu / (l2 + d2) + x / l2);
  const double Cc = x * x / l2 + mu * mu / (l2 + d2) - m * m * M;
  const double ym = y - mu, mm = m - mu;
  return l * sqrt((l2 + 2.0 * d2) / (d2 * M + 1.0)) * exp(-Cc * 0.5) / (l2 + d2) *
         exp(-(ym * ym) / (2.0 * (l2 + d2))) * exp(-(mm * mm) / (2.0 * (1.0 / M + d2)));
}

__device__ __forceinline__ double sobol_f4(double x, double y, double l, double delta, double mu) {
  const double l2 = l * l, d2 = delta * delta;
  const double xm = x - mu, ym = y - mu;
  return l2 * (l2 + 2.0 * d2) * sqrt((l2 + d2) / (l2 + 3.0 * d2)) / ((l2 + d2) * (l2 + d2)) *
         exp(-(xm * xm + ym * ym) / (2.0 * (l2 + d2)));
}

// the four terms on their own, elementwise over paired (x_i, y_i) with the caller's sigma^4 (utils.py:116-165)
__global__ void sobol_terms_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, double s4,
                                   double l, double delta, double mu, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = s4 * sobol_f1(x[i], y[i], l, delta, mu);
  out[n + i] = s4 * sobol_f2(x[i], y[i], l, delta, mu);
  out[2 * n + i] = s4 * sobol_f2(y[i], x[i], l, delta, mu);
  out[3 * n + i] = s4 * sobol_f4(x[i], y[i], l, delta, mu);
}

// compute_L (utils.py:221-240): L = f1 - f2 - f3 + f4
__global__ void sobol_L_gaussian_kernel(const double* __restrict__ X, int64_t m, int64_t ldx, int col,
                                        double l, double delta, double mu, double* __restrict__ L,
                                        int64_t ldl) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= m || j >= m) return;
  const double x = X[i * ldx + col], y = X[j * ldx + col];
  L[i * ldl + j] = sobol_f1(x, y, l, delta, mu) - sobol_f2(x, y, l, delta, mu) -
                   sobol_f2(y, x, l, delta, mu) + sobol_f4(x, y, l, delta, mu);
}

// compute_L_binary_kernel (utils.py:243-272), evaluated on the float inputs like the reference
__global__ void sobol_L_binary_kernel(const double* __restrict__ X, int64_t m, int64_t ldx, int col,
                                      double p0, double* __restrict__ L, int64_t ldl) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= m || j >= m) return;
  const double x = X[i * ldx + col], y = X[j * ldx + col];
  const double p1 = 1.0 - p0;
  L[i * ldl + j] = p0 * (p1 * p1 * (1.0 - x) - p0 * p1 * x) * (p1 * p1 * (1.0 - y) - p0 * p1 * y) +
                   p1 * (-p0 * p1 * (1.0 - x) + p0 * p0 * x) * (-p0 * p1 * (1.0 - y) + p0 * p0 * y);
}

// compute_L_categorical_kernel (utils.py:275-309): L[i,j] = G[x_i, x_j], G = B diag(p) B^T
__global__ void sobol_L_categorical_kernel(const double* __restrict__ X, int64_t m, int64_t ldx,
                                           int col, const double* __restrict__ G, int C,
                                           double* __restrict__ L, int64_t ldl) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= m || j >= m) return;
  int a = (int)X[i * ldx + col], b = (int)X[j * ldx + col];
  a = max(0, min(a, C - 1));
  b = max(0, min(b, C - 1));
  L[i * ldl + j] = G[a * C + b];
}

// ---- empirical measure (utils.py:312-335): L = (w o k~(z,x))^T k~(z,x) ------------------
// chat[i] = cov_X_s(v_i) / sqrt(var_s) for raw values v (strided), one thread per value
__global__ void empirical_chat_kernel(const double* __restrict__ vals, int64_t n, int64_t stride,
                                      const double* __restrict__ loc, const double* __restrict__ w,
                                      int nloc, double inv_sqrt2_l, double s2,
                                      const double* __restrict__ inv_sqrt_v, double* __restrict__ a_out,
                                      double* __restrict__ chat_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = vals[i * stride] * inv_sqrt2_l;
  double acc = 0.0;
  for (int q = 0; q < nloc; ++q) {
    const double t = a - loc[q] * inv_sqrt2_l;
    acc = fma(w[q], exp(-t * t), acc);
  }
  a_out[i] = a;
  chat_out[i] = s2 * acc * inv_sqrt_v[0];
}

// Kxu[l, i] = k~(z_l, x_i) and its row-weighted copy
__global__ void empirical_kxu_kernel(const double* __restrict__ az, const double* __restrict__ cz,
                                     const double* __restrict__ w, int64_t nloc,
                                     const double* __restrict__ ax, const double* __restrict__ cx,
                                     int64_t m, double s2, double* __restrict__ K,
                                     double* __restrict__ Kw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t l = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (l >= nloc || i >= m) return;
  const double t = az[l] - ax[i];
  const double k = fma(-cz[l], cx[i], s2 * exp(-t * t));
  K[l * m + i] = k;
  Kw[l * m + i] = w[l] * k;
}

// part[c][s] = sum over the rows of split s of alpha_i alpha_j prod_{d in S_c} L_d[i,j]: block (component c, row
// split s); rows outermost, columns strided over the threads (no index division), fixed-order block reduction
__global__ void __launch_bounds__(512) sobol_quadforms_kernel(
    const double* __restrict__ Lstack, int64_t m, const int32_t* __restrict__ subsets, int max_order,
    const double* __restrict__ alpha, int splits, double* __restrict__ part) {
  __shared__ double sh[512];
  const int c = blockIdx.x, sp = blockIdx.y;
  const double* Ls[OAK_MAX_DEPTH];
  int order = 0;
  for (int q = 0; q < max_order; ++q) {
    const int d = subsets[c * max_order + q];
    if (d < 0) break;
    Ls[order++] = Lstack + (int64_t)d * m * m;
  }
  const int64_t rows_per = (m + splits - 1) / splits;
  const int64_t i0 = sp * rows_per, i1 = i0 + rows_per < m ? i0 + rows_per : m;
  double acc = 0.0;
  for (int64_t i = i0; i < i1; ++i) {
    const double ai = alpha[i];
    const int64_t base = i * m;
    for (int64_t j = threadIdx.x; j < m; j += 512) {
      double v = ai * alpha[j];
      for (int q = 0; q < order; ++q) v *= Ls[q][base + j];
      acc += v;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 256; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[(int64_t)c * splits + sp] = sh[0];
}
// out[c] = scale[c] * sum_s part[c][s] (fixed order)
__global__ void sobol_quadforms_fold_kernel(const double* __restrict__ part, int splits, const double* __restrict__ scale,
                                            int n, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += part[(int64_t)c * splits + s];
  out[c] = scale[c] * acc;
}

static cublasHandle_t g_sobol_cublas[64] = {nullptr};

}  // namespace oak

using namespace oak;

extern "C" size_t oak_sobol_L_work_bytes(const oak_spec* spec, int32_t dim, int64_t m) {
  if (!spec || dim < 0 || dim >= spec->D || m < 0) return 0;
  const DimDev& dd = spec->h_dims[spec->pos_of_orig[dim]];
  if (dd.type == OAK_DIM_RBF && dd.measure == OAK_MEASURE_EMPIRICAL)
    return (size_t)(2 * (int64_t)dd.count * m + 2 * dd.count + 2 * m) * sizeof(double);
  return 0;
}

extern "C" int oak_sobol_gaussian_terms_f64(const double* d_x, const double* d_y, int64_t n, double sigma,
                                            double lengthscale, double delta, double mu, double* d_out,
                                            void* stream_) {
  OAK_REQUIRE(d_x && d_y && d_out, "oak_sobol_gaussian_terms_f64: null argument");
  OAK_REQUIRE(lengthscale > 0.0, "oak_sobol_gaussian_terms_f64: lengthscale must be positive");
  if (n <= 0) return 0;
  const double s2 = sigma * sigma;
  sobol_terms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(d_x, d_y, n, s2 * s2, lengthscale,
                                                                                    delta, mu, d_out);
  OAK_LAUNCHED();
  return 0;
}

extern "C" int oak_sobol_L_f64(const oak_spec* spec, int32_t dim, const double* d_Xcond, int64_t m,
                               int64_t ldx, double delta, double mu, double* d_L, int64_t ldl,
                               void* d_work, void* stream_) {
  OAK_REQUIRE(spec && d_Xcond && d_L, "oak_sobol_L_f64: null argument");
  OAK_REQUIRE(dim >= 0 && dim < spec->D, "oak_sobol_L_f64: dim out of range");
  OAK_REQUIRE(ldl >= m, "oak_sobol_L_f64: ldl smaller than m");
  if (m <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int pos = spec->pos_of_orig[dim];
  const DimDev& dd = spec->h_dims[pos];
  dim3 block(32, 8);
  dim3 grid((unsigned)((m + 31) / 32), (unsigned)((m + 7) / 8));
  if (dd.type == OAK_DIM_RBF) {
    if (dd.measure == OAK_MEASURE_MOG) {
      set_error("oak_sobol_L_f64: Sobol indices are not implemented for the MOG measure");
      return 5;  // NotImplementedError in the reference (utils.py:413-414)
    }
    if (dd.measure == OAK_MEASURE_EMPIRICAL) {
      OAK_REQUIRE(d_work, "oak_sobol_L_f64: empirical measure needs a workspace");
      const int64_t nl = dd.count;
      double* K = (double*)d_work;
      double* Kw = K + nl * m;
      double* az = Kw + nl * m;
      double* cz = az + nl;
      double* ax = cz + nl;
      double* cx = ax + m;
      empirical_chat_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, stream>>>(
          dd.v0, nl, 1, dd.v0, dd.v1, (int)nl, dd.inv_sqrt2_l, dd.s2, spec->d_inv_sqrt_v + pos, az, cz);
      OAK_LAUNCHED();
      empirical_chat_kernel<<<(unsigned)((m + 127) / 128), 128, 0, stream>>>(
          d_Xcond + dd.column, m, ldx, dd.v0, dd.v1, (int)nl, dd.inv_sqrt2_l, dd.s2,
          spec->d_inv_sqrt_v + pos, ax, cx);
      OAK_LAUNCHED();
      dim3 g2((unsigned)((m + 31) / 32), (unsigned)((nl + 7) / 8));
      empirical_kxu_kernel<<<g2, block, 0, stream>>>(az, cz, dd.v1, nl, ax, cx, m, dd.s2, K, Kw);
      OAK_LAUNCHED();
      // L = Kw^T K.  Row-major (nl x m) == column-major (m x nl): L = K'w K'^T
      int dev = 0;
      OAK_CUDA(cudaGetDevice(&dev));
      if (!g_sobol_cublas[dev]) {
        if (cublasCreate(&g_sobol_cublas[dev]) != CUBLAS_STATUS_SUCCESS) {
          set_error("oak_sobol_L_f64: cublasCreate failed");
          return 3;
        }
      }
      cublasSetStream(g_sobol_cublas[dev], stream);
      const double one = 1.0, zero = 0.0;
      if (cublasDgemm(g_sobol_cublas[dev], CUBLAS_OP_N, CUBLAS_OP_T, (int)m, (int)m, (int)nl, &one, Kw,
                      (int)m, K, (int)m, &zero, d_L, (int)ldl) != CUBLAS_STATUS_SUCCESS) {
        set_error("oak_sobol_L_f64: cublasDgemm failed");
        return 3;
      }
      g_launches.fetch_add(1);
      return 0;
    }
    // Gaussian closed form; the reference also routes Uniform-measure kernels here (utils.py:386-400)
    sobol_L_gaussian_kernel<<<grid, block, 0, stream>>>(d_Xcond, m, ldx, dd.column, dd.lengthscale,
                                                        delta, mu, d_L, ldl);
    OAK_LAUNCHED();
    return 0;
  }
  if (dd.type == OAK_DIM_BINARY) {
    sobol_L_binary_kernel<<<grid, block, 0, stream>>>(d_Xcond, m, ldx, dd.column, dd.c0, d_L, ldl);
    OAK_LAUNCHED();
    return 0;
  }
  sobol_L_categorical_kernel<<<grid, block, 0, stream>>>(d_Xcond, m, ldx, dd.column,
                                                         spec->d_sobolG + spec->sobol_off[pos],
                                                         dd.count, d_L, ldl);
  OAK_LAUNCHED();
  return 0;
}

static int quadform_splits(int32_t num_components, int64_t m) {
  // few components (config A: 255): split the rows of every component over several blocks so that the 592 block
  // slots of the device are filled
  constexpr int kSlots = 592, kMaxSplits = 16;
  int splits = num_components < kSlots ? (kSlots + num_components - 1) / num_components : 1;
  if (splits > kMaxSplits) splits = kMaxSplits;
  if ((int64_t)splits * 8 > m) splits = (int)((m + 7) / 8);
  return splits < 1 ? 1 : splits;
}

extern "C" size_t oak_sobol_quadforms_work_bytes(int32_t num_components, int64_t m) {
  if (num_components <= 0 || m <= 0) return 0;
  return (size_t)num_components * quadform_splits(num_components, m) * sizeof(double);
}

extern "C" int oak_sobol_quadforms_f64(const double* d_Lstack, int32_t num_dims, int64_t m,
                                       const int32_t* d_subsets, const double* d_scale,
                                       int32_t num_components, int32_t max_order,
                                       const double* d_alpha, double* d_out, void* d_work, void* stream_) {
  OAK_REQUIRE(d_Lstack && d_subsets && d_scale && d_alpha && d_out && d_work,
              "oak_sobol_quadforms_f64: null argument");
  OAK_REQUIRE(max_order >= 1 && max_order <= OAK_MAX_DEPTH, "oak_sobol_quadforms_f64: bad max_order");
  (void)num_dims;
  if (num_components <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int splits = quadform_splits(num_components, m);
  double* part = (double*)d_work;  // [num_components][splits]
  sobol_quadforms_kernel<<<dim3((unsigned)num_components, (unsigned)splits), 512, 0, stream>>>(
      d_Lstack, m, d_subsets, max_order, d_alpha, splits, part);
  OAK_LAUNCHED();
  sobol_quadforms_fold_kernel<<<(unsigned)((num_components + 255) / 256), 256, 0, stream>>>(part, splits, d_scale,
                                                                                           num_components, d_out);
  OAK_LAUNCHED();
  return 0;
}
