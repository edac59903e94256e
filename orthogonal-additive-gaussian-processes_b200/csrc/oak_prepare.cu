// Per-point prologue and K_diag.
//
// Prepared-point block ("points"): double2 pts[D][n_pad], feature-major, kernel order (RBF dims
// first).  RBF dim:   .x = x sqrt(T/ln2) / (sqrt(2) l), T = kExpTab   .y = cov_X_s(x) / sqrt(var_s())
//          discrete:  .x = bits{lo: idx, hi: idx*C}         .y = B_diag[idx]
// Padding rows (n <= i < n_pad) hold zeros so that tiles can be loaded without bounds checks.
// The block ends with 2 D uint64 keys: per-dimension min and max of .x over the n real points
// (order_key encoding); the Gram kernel uses them to prove that the clamp-free exp is safe.
#include "oak_common.cuh"

namespace oak {

constexpr int kSmallEmpirical = 64;

// One (point, sub-kernel) pair: the prepared coordinate / correction (RBF) or the table offsets (discrete).
// lo / hi receive the prepared coordinate of an RBF dim (min / max keys), +-inf otherwise.
__device__ __forceinline__ double2 prepare_one(const DimDev& d, double x, double inv_sqrt_v_k,
                                               const double* __restrict__ tables, double& lo, double& hi) {
  double2 out;
  if (d.type == OAK_DIM_RBF) {
    out.x = x * d.xscale;
    lo = hi = out.x;
    double c = 0.0;
    switch (d.measure) {
      case OAK_MEASURE_GAUSSIAN: {  // ortho_rbf_kernel.py:82-92
        const double t = x - d.c1;
        c = d.c0 * exp(-(t * t) * d.c2);
        break;
      }
      case OAK_MEASURE_UNIFORM:  // ortho_rbf_kernel.py:49-63
        c = d.c0 * (erf((d.c2 - x) * d.inv_sqrt2_l) - erf((d.c1 - x) * d.inv_sqrt2_l));
        break;
      case OAK_MEASURE_MOG: {  // ortho_rbf_kernel.py:124-136
        const double l2 = d.lengthscale * d.lengthscale;
        double acc = 0.0;
        for (int q = 0; q < d.count; ++q) {
          const double sc = l2 + d.v1[q];
          const double t = x - d.v0[q];
          acc += exp(-0.5 * (t * t) / sc) / sqrt(sc) * d.v2[q];
        }
        c = d.c0 * acc;
        break;
      }
      case OAK_MEASURE_EMPIRICAL: {  // ortho_rbf_kernel.py:101-107
        double acc = 0.0;
        const int cnt = d.count <= kSmallEmpirical ? d.count : 0;  // large: tiled kernel below
        for (int q = 0; q < cnt; ++q) {
          const double t = (x - d.v0[q]) * d.inv_sqrt2_l;
          acc = fma(d.v1[q], exp(-t * t), acc);
        }
        c = d.c0 * acc;
        break;
      }
      default:
        break;
    }
    out.y = c * inv_sqrt_v_k;
  } else {
    // tf.cast(x, int32): truncation toward zero (ortho_binary_kernel.py:47-51); clamped
    // into the table (the host wrapper validates the range and raises).
    int idx = (int)x;
    idx = max(0, min(idx, d.count - 1));
    out.x = __hiloint2double(idx * d.count, idx);
    out.y = tables[d.table_off + d.count * d.count + idx];
  }
  return out;
}

__global__ void prepare_points_kernel(const DimDev* __restrict__ dims,
                                      const double* __restrict__ inv_sqrt_v,
                                      const double* __restrict__ tables,
                                      const double* __restrict__ X, int64_t n, int64_t n_pad,
                                      int64_t ldx, double2* __restrict__ pts,
                                      unsigned long long* __restrict__ minmax, int D) {
  const int k = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double2 out = make_double2(0.0, 0.0);
  double lo = INFINITY, hi = -INFINITY;
  if (i < n) {
    const DimDev d = dims[k];
    out = prepare_one(d, X[i * ldx + d.column], inv_sqrt_v[k], tables, lo, hi);
  }
  if (i < n_pad) pts[(int64_t)k * n_pad + i] = out;
  // block-wide min / max of the prepared coordinate -> one atomic pair per block
  // (NaNs compare false and are skipped; they poison the Gram entries on either exp path)
  __shared__ double s_lo[8], s_hi[8];
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      lo = fmin(lo, s_lo[w]);
      hi = fmax(hi, s_hi[w]);
    }
    if (lo <= hi) {
      atomicMin(minmax + k, order_key(lo));
      atomicMax(minmax + D + k, order_key(hi));
    }
  }
}

// The same prologue for narrow inputs (ldx <= kRowTileMaxLd): one block takes kRowTile ROWS of X -- a contiguous
// range of the row-major matrix, read once with coalesced loads into shared memory -- and walks all D sub-kernels.
// The column-per-block kernel above reads a column with a stride of ldx doubles: 8 x the sectors, D times over
// (0.41 ms for 10^6 x 20 inputs, against 0.48 GB of useful traffic).
constexpr int kRowTile = 128;
constexpr int kRowTileMaxLd = 96;  // 128 x 96 doubles = 96 KB of dynamic shared memory (opt-in above 48 KB; + 4 KB static)
constexpr int kRowTileMaxD = 64;

__global__ void __launch_bounds__(kRowTile) prepare_points_rows_kernel(
    const DimDev* __restrict__ dims, const double* __restrict__ inv_sqrt_v, const double* __restrict__ tables,
    const double* __restrict__ X, int64_t n, int64_t n_pad, int ldx, double2* __restrict__ pts,
    unsigned long long* __restrict__ minmax, int D) {
  extern __shared__ double sx[];  // [kRowTile][ldx]
  __shared__ double s_lo[kRowTile / 32][kRowTileMaxD], s_hi[kRowTile / 32][kRowTileMaxD];
  const int64_t i0 = (int64_t)blockIdx.x * kRowTile;
  const int rows = (int)min((int64_t)kRowTile, n - i0);  // may be <= 0 for pure padding blocks
  const int64_t base = i0 * ldx;
  for (int e = threadIdx.x; e < rows * ldx; e += kRowTile) sx[e] = X[base + e];
  __syncthreads();
  const int t = threadIdx.x;
  const int64_t i = i0 + t;
  for (int k = 0; k < D; ++k) {
    double2 out = make_double2(0.0, 0.0);
    double lo = INFINITY, hi = -INFINITY;
    if (t < rows) {
      const DimDev d = dims[k];
      out = prepare_one(d, sx[t * ldx + d.column], inv_sqrt_v[k], tables, lo, hi);
    }
    if (i < n_pad) pts[(int64_t)k * n_pad + i] = out;
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((t & 31) == 0) {
      s_lo[t >> 5][k] = lo;
      s_hi[t >> 5][k] = hi;
    }
  }
  __syncthreads();
  for (int k = t; k < D; k += kRowTile) {
    double lo = s_lo[0][k], hi = s_hi[0][k];
    for (int w = 1; w < kRowTile / 32; ++w) {
      lo = fmin(lo, s_lo[w][k]);
      hi = fmax(hi, s_hi[w][k]);
    }
    if (lo <= hi) {
      atomicMin(minmax + k, order_key(lo));
      atomicMax(minmax + D + k, order_key(hi));
    }
  }
}

// Tiled variant for empirical measures with many locations: the locations/weights of the dim
// are streamed through shared memory once per block instead of once per thread.
__global__ void prepare_empirical_kernel(const DimDev* __restrict__ dims, int k,
                                         const double* __restrict__ inv_sqrt_v,
                                         const double* __restrict__ X, int64_t n, int64_t n_pad,
                                         int64_t ldx, double2* __restrict__ pts) {
  extern __shared__ double sh[];
  const int tile = blockDim.x;
  double* sl = sh;
  double* sw = sh + tile;
  const DimDev d = dims[k];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double x = (i < n) ? X[i * ldx + d.column] : 0.0;
  const double a = x * d.inv_sqrt2_l;
  double acc = 0.0;
  for (int base = 0; base < d.count; base += tile) {
    const int j = base + threadIdx.x;
    sl[threadIdx.x] = (j < d.count) ? d.v0[j] * d.inv_sqrt2_l : 0.0;
    sw[threadIdx.x] = (j < d.count) ? d.v1[j] : 0.0;
    __syncthreads();
    const int lim = min(tile, d.count - base);
    for (int q = 0; q < lim; ++q) {
      const double t = a - sl[q];
      acc = fma(sw[q], exp(-t * t), acc);
    }
    __syncthreads();
  }
  if (i < n_pad) {
    double2 out = make_double2(0.0, 0.0);
    if (i < n) out = make_double2(x * d.xscale, d.c0 * acc * inv_sqrt_v[k]);
    pts[(int64_t)k * n_pad + i] = out;
  }
}

// K_diag (oak_kernel.py:267-278): per-dim diagonals -> power sums -> Newton-Girard -> sum.
// O(N D P): one thread per point, the recurrence runs in registers (depth <= OAK_MAX_DEPTH).
struct DiagParams {
  double sigma2[OAK_MAX_DEPTH + 1];
  int D, Dc, depth, algo;
};

__global__ void gram_diag_kernel(DiagParams prm, const DimDev* __restrict__ dims,
                                 const double2* __restrict__ pts, int64_t n, int64_t n_pad,
                                 double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int P = prm.depth;
  double acc[OAK_MAX_DEPTH + 1];
#pragma unroll
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) acc[p] = 0.0;
  if (prm.algo == OAK_ESP_DIRECT) acc[0] = 1.0;
  for (int k = 0; k < prm.D; ++k) {
    const double2 v = pts[(int64_t)k * n_pad + i];
    double kd;
    if (k < prm.Dc) {
      kd = dims[k].s2 - v.y * v.y;  // s^2 - c^2 / v   (ortho_rbf_kernel.py:174-177)
    } else {
      kd = v.y;  // output_variance gather (ortho_binary_kernel.py:55-59)
    }
    if (prm.algo == OAK_ESP_DIRECT) {
#pragma unroll
      for (int p = OAK_MAX_DEPTH; p >= 1; --p)
        if (p <= P) acc[p] = fma(kd, acc[p - 1], acc[p]);
    } else {
      double pw = 1.0;
#pragma unroll
      for (int p = 1; p <= OAK_MAX_DEPTH; ++p)
        if (p <= P) {
          pw *= kd;
          acc[p] += pw;
        }
    }
  }
  double e[OAK_MAX_DEPTH + 1];
  e[0] = 1.0;
  if (prm.algo == OAK_ESP_DIRECT) {
#pragma unroll
    for (int p = 1; p <= OAK_MAX_DEPTH; ++p) e[p] = acc[p];
  } else {
#pragma unroll
    for (int nn = 1; nn <= OAK_MAX_DEPTH; ++nn) {
      double s = 0.0;
      if (nn <= P) {
#pragma unroll
        for (int q = 1; q <= nn; ++q) {
          const double term = e[nn - q] * acc[q];
          s = (q & 1) ? s + term : s - term;
        }
        s *= 1.0 / nn;
      }
      e[nn] = s;
    }
  }
  double r = prm.sigma2[0];
#pragma unroll
  for (int p = 1; p <= OAK_MAX_DEPTH; ++p)
    if (p <= P) r = fma(prm.sigma2[p], e[p], r);
  out[i] = r;
}

}  // namespace oak

using namespace oak;

extern "C" size_t oak_points_bytes(const oak_spec* spec, int64_t n) {
  if (!spec || n < 0) return 0;
  return (size_t)spec->D * (size_t)padded(n) * sizeof(double2) +
         2 * (size_t)spec->D * sizeof(unsigned long long);
}

extern "C" int oak_prepare_points_f64(const oak_spec* spec, const double* d_X, int64_t n,
                                      int64_t ldx, void* d_points, void* stream_) {
  OAK_REQUIRE(spec && d_points, "oak_prepare_points_f64: null argument");
  OAK_REQUIRE(n >= 0, "oak_prepare_points_f64: negative n");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t n_pad = padded(n);
  const int D = spec->D;
  // min keys start at +max, max keys at 0 (an empty block never qualifies for the fast exp)
  unsigned long long* minmax = (unsigned long long*)((double2*)d_points + (int64_t)D * n_pad);
  OAK_CUDA(cudaMemsetAsync(minmax, 0xFF, (size_t)D * sizeof(unsigned long long), stream));
  OAK_CUDA(cudaMemsetAsync(minmax + D, 0x00, (size_t)D * sizeof(unsigned long long), stream));
  if (n == 0) return 0;
  OAK_REQUIRE(d_X, "oak_prepare_points_f64: null X");
  const int threads = 256;
  if (ldx <= kRowTileMaxLd && D <= kRowTileMaxD) {
    const size_t tile_bytes = (size_t)kRowTile * ldx * sizeof(double);
    if (tile_bytes > 40 * 1024)
      OAK_CUDA(cudaFuncSetAttribute(prepare_points_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kRowTile * kRowTileMaxLd * (int)sizeof(double)));
    prepare_points_rows_kernel<<<(unsigned)(n_pad / kRowTile), kRowTile, (size_t)kRowTile * ldx * sizeof(double), stream>>>(
        spec->d_dims, spec->d_inv_sqrt_v, spec->d_tables, d_X, n, n_pad, (int)ldx, (double2*)d_points, minmax, D);
  } else {
    dim3 grid((unsigned)((n_pad + threads - 1) / threads), (unsigned)D);
    prepare_points_kernel<<<grid, threads, 0, stream>>>(spec->d_dims, spec->d_inv_sqrt_v, spec->d_tables, d_X, n, n_pad,
                                                        ldx, (double2*)d_points, minmax, D);
  }
  OAK_LAUNCHED();
  // large empirical measures: overwrite with the tiled kernel
  for (int k = 0; k < spec->Dc; ++k) {
    const DimDev& dd = spec->h_dims[k];
    if (dd.measure == OAK_MEASURE_EMPIRICAL && dd.count > kSmallEmpirical) {
      prepare_empirical_kernel<<<(unsigned)((n_pad + threads - 1) / threads), threads,
                                 2 * threads * sizeof(double), stream>>>(
          spec->d_dims, k, spec->d_inv_sqrt_v, d_X, n, n_pad, ldx, (double2*)d_points);
      OAK_LAUNCHED();
    }
  }
  return 0;
}

int oak::gram_diag_launch(const oak_spec* spec, const double2* pts, int64_t n, int64_t n_pad,
                          double* out, cudaStream_t stream) {
  DiagParams prm;
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) prm.sigma2[p] = spec->sigma2[p];
  prm.D = spec->D;
  prm.Dc = spec->Dc;
  prm.depth = spec->depth;
  prm.algo = spec->algo;
  const int threads = 256;
  gram_diag_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(
      prm, spec->d_dims, pts, n, n_pad, out);
  OAK_LAUNCHED();
  return 0;
}

extern "C" int oak_gram_diag_f64(const oak_spec* spec, const void* d_points, int64_t n,
                                 double* d_out, void* stream_) {
  OAK_REQUIRE(spec && d_points && d_out, "oak_gram_diag_f64: null argument");
  if (n <= 0) return 0;
  return gram_diag_launch(spec, (const double2*)d_points, n, padded(n), d_out,
                          (cudaStream_t)stream_);
}
