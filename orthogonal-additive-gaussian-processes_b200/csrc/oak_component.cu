// Single additive components sigma^2_|S| prod_{d in S} k_d:
//   KernelComponenent.K / K_diag       (oak/oak_kernel.py:300-335)
//   get_prediction_component           (oak/utils.py:491-530), fused with the alpha contraction
//   compute_additive_terms on explicit matrices (oak/oak_kernel.py:223-249)
// These are consumers of the same prepared points as the Gram tile; they are not the FP64-bound
// hot loop (|S| <= depth factors per entry instead of D), so they use simple 2-D launches.
#include <cstring>

#include "oak_common.cuh"

namespace oak {

constexpr int kMaxOrder = OAK_MAX_DEPTH;

struct SubsetParams {
  int order;
  int pos[kMaxOrder];    // kernel-order positions of the dims in the subset
  double aux[kMaxOrder]; // -ln s^2 (RBF) or bits(table offset) (discrete)
  int discrete[kMaxOrder];
  double scale;          // order variance
};

__device__ __forceinline__ double dim_value(bool discrete, double aux, const double* tables,
                                            double2 a, double2 b) {
  if (discrete) {
    const double* tbl = tables + (int)__double_as_longlong(aux);
    return tbl[__double2hiint(a.x) + __double2loint(b.x)];
  }
  const double t = a.x - b.x;  // prepared coordinates carry sqrt(256/ln2)
  return fma(-a.y, b.y, exp(-fma(t * t, kInvXScale2, aux)));
}

__global__ void component_gram_kernel(SubsetParams sp, const double* __restrict__ tables,
                                      const double2* __restrict__ prow, int64_t n, int64_t n_pad,
                                      const double2* __restrict__ pcol, int64_t n2, int64_t n2_pad,
                                      double* __restrict__ K, int64_t ldk) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n || j >= n2) return;
  double v = sp.scale;
  for (int q = 0; q < sp.order; ++q) {
    const double2 a = prow[(int64_t)sp.pos[q] * n_pad + i];
    const double2 b = pcol[(int64_t)sp.pos[q] * n2_pad + j];
    v *= dim_value(sp.discrete[q], sp.aux[q], tables, a, b);
  }
  K[i * ldk + j] = v;
}

__global__ void component_diag_kernel(SubsetParams sp, const DimDev* __restrict__ dims,
                                      const double2* __restrict__ pts, int64_t n, int64_t n_pad,
                                      double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = sp.scale;
  for (int q = 0; q < sp.order; ++q) {
    const double2 a = pts[(int64_t)sp.pos[q] * n_pad + i];
    v *= sp.discrete[q] ? a.y : (dims[sp.pos[q]].s2 - a.y * a.y);
  }
  out[i] = v;
}

// out[c*n + i] = sigma2_|S_c| * sum_j prod_{d in S_c} k_d(x_i, z_j) alpha_j
// grid: (ceil(n/128), num_components); the conditioning points of the component's dims are
// streamed through shared memory in tiles of 128.
__global__ void __launch_bounds__(128) component_predict_kernel(
    const int32_t* __restrict__ subsets, int max_order, const int32_t* __restrict__ pos_of_orig,
    const double* __restrict__ dim_aux, int Dc, const double* __restrict__ tables,
    const double* __restrict__ sigma2, const double2* __restrict__ px, int64_t n, int64_t n_pad,
    const double2* __restrict__ pz, int64_t m, int64_t m_pad, const double* __restrict__ alpha,
    double* __restrict__ out) {
  __shared__ double2 sz[kMaxOrder][128];
  __shared__ double sa[128];
  const int c = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  int pos[kMaxOrder];
  double aux[kMaxOrder];
  int order = 0;
  for (int q = 0; q < max_order; ++q) {
    const int d = subsets[c * max_order + q];
    if (d < 0) break;
    pos[order] = pos_of_orig[d];
    aux[order] = dim_aux[pos[order]];
    ++order;
  }
  double2 xi[kMaxOrder];
  for (int q = 0; q < order; ++q)
    xi[q] = px[(int64_t)pos[q] * n_pad + (i < n ? i : 0)];
  double acc = 0.0;
  for (int64_t base = 0; base < m; base += 128) {
    const int64_t j = base + threadIdx.x;
    for (int q = 0; q < order; ++q) sz[q][threadIdx.x] = pz[(int64_t)pos[q] * m_pad + (j < m ? j : 0)];
    sa[threadIdx.x] = (j < m) ? alpha[j] : 0.0;
    __syncthreads();
    const int lim = (int)min((int64_t)128, m - base);
    for (int jj = 0; jj < lim; ++jj) {
      double v = sa[jj];
      for (int q = 0; q < order; ++q) v *= dim_value(pos[q] >= Dc, aux[q], tables, xi[q], sz[q][jj]);
      acc += v;
    }
    __syncthreads();
  }
  if (i < n) out[(int64_t)c * n + i] = sigma2[order] * acc;
}

// e_0..e_P of D explicit arrays (oak_kernel.py:223-249), element-wise.
__global__ void additive_terms_kernel(const double* __restrict__ mats, int D, int P, int64_t len,
                                      double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s[OAK_MAX_DEPTH + 1], e[OAK_MAX_DEPTH + 1];
  for (int p = 0; p <= P; ++p) s[p] = 0.0;
  for (int d = 0; d < D; ++d) {
    const double k = mats[(int64_t)d * len + i];
    double pw = 1.0;
    for (int p = 1; p <= P; ++p) {
      pw *= k;
      s[p] += pw;
    }
  }
  e[0] = 1.0;
  out[i] = 1.0;
  for (int n = 1; n <= P; ++n) {
    double acc = 0.0;
    for (int q = 1; q <= n; ++q) acc += ((q & 1) ? 1.0 : -1.0) * e[n - q] * s[q];
    e[n] = acc / n;
    out[(int64_t)n * len + i] = e[n];
  }
}

static int fill_subset(const oak_spec* spec, const int32_t* h_subset, int order, SubsetParams& sp) {
  OAK_REQUIRE(order >= 0 && order <= kMaxOrder, "component: order outside [0, OAK_MAX_DEPTH]");
  OAK_REQUIRE(order == 0 || h_subset, "component: null subset");
  sp.order = order;
  sp.scale = spec->sigma2[order];
  for (int q = 0; q < order; ++q) {
    OAK_REQUIRE(h_subset[q] >= 0 && h_subset[q] < spec->D, "component: dim index out of range");
    const int pos = spec->pos_of_orig[h_subset[q]];
    sp.pos[q] = pos;
    sp.discrete[q] = pos >= spec->Dc;
    if (sp.discrete[q]) {
      const int64_t off = spec->h_dims[pos].table_off;
      memcpy(&sp.aux[q], &off, sizeof(double));
    } else {
      sp.aux[q] = spec->h_dims[pos].neg_log_s2;
    }
  }
  return 0;
}

}  // namespace oak

using namespace oak;

extern "C" int oak_component_gram_f64(const oak_spec* spec, const int32_t* h_subset, int32_t order,
                                      const void* d_points, int64_t n, const void* d_points2,
                                      int64_t n2, double* d_K, int64_t ldk, void* stream_) {
  OAK_REQUIRE(spec && d_points && d_K, "oak_component_gram_f64: null argument");
  if (!d_points2) n2 = n;
  if (n <= 0 || n2 <= 0) return 0;
  SubsetParams sp;
  if (int rc = fill_subset(spec, h_subset, order, sp)) return rc;
  dim3 block(32, 8);
  dim3 grid((unsigned)((n2 + 31) / 32), (unsigned)((n + 7) / 8));
  component_gram_kernel<<<grid, block, 0, (cudaStream_t)stream_>>>(
      sp, spec->d_tables, (const double2*)d_points, n, padded(n),
      (const double2*)(d_points2 ? d_points2 : d_points), n2, padded(n2), d_K, ldk);
  OAK_LAUNCHED();
  return 0;
}

extern "C" int oak_component_diag_f64(const oak_spec* spec, const int32_t* h_subset, int32_t order,
                                      const void* d_points, int64_t n, double* d_out,
                                      void* stream_) {
  OAK_REQUIRE(spec && d_points && d_out, "oak_component_diag_f64: null argument");
  if (n <= 0) return 0;
  SubsetParams sp;
  if (int rc = fill_subset(spec, h_subset, order, sp)) return rc;
  component_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      sp, spec->d_dims, (const double2*)d_points, n, padded(n), d_out);
  OAK_LAUNCHED();
  return 0;
}

extern "C" int oak_component_predict_f64(const oak_spec* spec, const int32_t* d_subsets,
                                         int32_t num_components, int32_t max_order,
                                         const void* d_points, int64_t n,
                                         const void* d_points_cond, int64_t m,
                                         const double* d_alpha, double* d_out, void* stream_) {
  OAK_REQUIRE(spec && d_subsets && d_points && d_points_cond && d_alpha && d_out,
              "oak_component_predict_f64: null argument");
  OAK_REQUIRE(max_order >= 1 && max_order <= kMaxOrder, "oak_component_predict_f64: bad max_order");
  if (num_components <= 0 || n <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  // small per-call device copies of the position map and order variances
  int32_t* d_pos = nullptr;
  double* d_sig = nullptr;
  OAK_CUDA(cudaMallocAsync(&d_pos, spec->D * sizeof(int32_t), stream));
  OAK_CUDA(cudaMallocAsync(&d_sig, (OAK_MAX_DEPTH + 1) * sizeof(double), stream));
  OAK_CUDA(cudaMemcpyAsync(d_pos, spec->pos_of_orig.data(), spec->D * sizeof(int32_t),
                           cudaMemcpyHostToDevice, stream));
  OAK_CUDA(cudaMemcpyAsync(d_sig, spec->sigma2, (OAK_MAX_DEPTH + 1) * sizeof(double),
                           cudaMemcpyHostToDevice, stream));
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)num_components);
  component_predict_kernel<<<grid, 128, 0, stream>>>(
      d_subsets, max_order, d_pos, spec->d_neg_log_s2, spec->Dc, spec->d_tables, d_sig,
      (const double2*)d_points, n, padded(n), (const double2*)d_points_cond, m, padded(m), d_alpha,
      d_out);
  OAK_LAUNCHED();
  cudaFreeAsync(d_pos, stream);
  cudaFreeAsync(d_sig, stream);
  return 0;
}

extern "C" int oak_additive_terms_f64(const double* d_mats, int32_t num_mats, int32_t depth,
                                      int64_t len, double* d_out, void* stream_) {
  OAK_REQUIRE(d_mats && d_out, "oak_additive_terms_f64: null argument");
  OAK_REQUIRE(depth >= 0 && depth <= OAK_MAX_DEPTH, "oak_additive_terms_f64: depth out of range");
  OAK_REQUIRE(num_mats >= 1, "oak_additive_terms_f64: need at least one matrix");
  if (len <= 0) return 0;
  additive_terms_kernel<<<(unsigned)((len + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      d_mats, num_mats, depth, len, d_out);
  OAK_LAUNCHED();
  return 0;
}
