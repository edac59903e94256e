// SGPR statistics and the dense M^3 / N^3 tails.
//
// Two generations of entry points live here.
//   * oak_sgpr_factor_f64 / oak_sgpr_stats2_f64 / oak_sgpr_factor_stats_f64 / oak_sgpr_finish2_f64 (round 2, what
//     models.SGPR calls): L = chol(Kuu) FIRST by the one-launch bordered Cholesky (oak_chol.cu), a condition estimate
//     and a device-side route flag; the local N points are streamed in chunks -- fused Gram tiles write an M x chunk
//     block of Kuf (with Kuf y folded into the tile epilogue), the hand-written stream-K DMMA kernel of oak_syrk.cu
//     contracts it into Phi = Kuf Kuf^T (route 0) or, after the DMMA whitening product of oak_pgemm.cu, into
//     Psi = sum (L^-1 Kuf)(L^-1 Kuf)^T (route 1: gpflow's operation order, oak/utils.py:186-190); K_diag and y^T y are
//     reduced by deterministic fixed-order kernels.  Phi | Kuf y | sum K_diag | y^T y is one contiguous vector: the
//     only thing that is all-reduced across ranks.  The tail is one more bordered Cholesky with (L^-1 Kuf y) / noise as
//     its border row.  The fused call hides the factorisation behind the first chunk's tiles (side stream).
//   * oak_sgpr_stats_f64 / oak_sgpr_stats_keep_f64 / oak_sgpr_finish_f64 (round 1, kept for the training path, which
//     needs the un-whitened Kuf blocks): the same chunk loop with the contraction selectable by OAK_SYRK_MODE
//     (30 = the DMMA kernel, default; 0 / 1 / 9 / 20 = cuBLAS DSYRK / recursive DGEMM / DGEMM / batched DGEMM, the
//     alternatives measured in profiles/r01_ab_sgpr_contraction.txt) and a cuSOLVER potrf + cuBLAS trsm / trsv tail.
// oak_gpr_finish_f64 is the GPR tail (cuSOLVER potrf / potrs), timed separately from the tile kernels.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cmath>
#include <functional>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "oak_common.cuh"

namespace oak {

#define OAK_CUBLAS(call)                                                                 \
  do {                                                                                   \
    cublasStatus_t _s = (call);                                                          \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                                   \
      ::oak::set_error(std::string(#call) + " failed with cuBLAS status " +              \
                       std::to_string((int)_s));                                         \
      return 3;                                                                          \
    }                                                                                    \
  } while (0)

#define OAK_CUSOLVER(call)                                                               \
  do {                                                                                   \
    cusolverStatus_t _s = (call);                                                        \
    if (_s != CUSOLVER_STATUS_SUCCESS) {                                                 \
      ::oak::set_error(std::string(#call) + " failed with cuSOLVER status " +            \
                       std::to_string((int)_s));                                         \
      return 4;                                                                          \
    }                                                                                    \
  } while (0)

static std::mutex g_handle_mu;
static cublasHandle_t g_cublas[64] = {nullptr};
static cusolverDnHandle_t g_cusolver[64] = {nullptr};

static int handles(cublasHandle_t* cb, cusolverDnHandle_t* cs, cudaStream_t stream) {
  int dev = 0;
  OAK_CUDA(cudaGetDevice(&dev));
  OAK_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  std::lock_guard<std::mutex> lock(g_handle_mu);
  if (!g_cublas[dev]) {
    OAK_CUBLAS(cublasCreate(&g_cublas[dev]));
    OAK_CUBLAS(cublasSetPointerMode(g_cublas[dev], CUBLAS_POINTER_MODE_HOST));
  }
  if (!g_cusolver[dev]) OAK_CUSOLVER(cusolverDnCreate(&g_cusolver[dev]));
  OAK_CUBLAS(cublasSetStream(g_cublas[dev], stream));
  OAK_CUSOLVER(cusolverDnSetStream(g_cusolver[dev], stream));
  if (cb) *cb = g_cublas[dev];
  if (cs) *cs = g_cusolver[dev];
  return 0;
}

// Side stream of the overlapped front (oak_sgpr_factor_stats_f64): its own cuBLAS handle (a handle's
// workspace belongs to one stream at a time), a non-blocking stream (no implicit ordering against the legacy
// default stream the caller may be using) and the two events of the fork / join.
struct SideLane {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cublasHandle_t cublas = nullptr;
};
static SideLane g_side[64];

static int side_lane(SideLane** out) {
  int dev = 0;
  OAK_CUDA(cudaGetDevice(&dev));
  OAK_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  std::lock_guard<std::mutex> lock(g_handle_mu);
  SideLane& l = g_side[dev];
  if (!l.stream) {
    OAK_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    OAK_CUDA(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
    OAK_CUDA(cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming));
    OAK_CUBLAS(cublasCreate(&l.cublas));
    OAK_CUBLAS(cublasSetPointerMode(l.cublas, CUBLAS_POINTER_MODE_HOST));
    OAK_CUBLAS(cublasSetStream(l.cublas, l.stream));
  }
  *out = &l;
  return 0;
}

// out[0] += sum_i a[i] * (b ? b[i] : 1), fixed summation order (one block).
__global__ void __launch_bounds__(1024) reduce_accumulate_kernel(const double* __restrict__ a,
                                                                 const double* __restrict__ b,
                                                                 int64_t n, double* out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += b ? a[i] * b[i] : a[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] += sh[0];
}

// Two-level variant for long vectors: block b reduces the contiguous segment [b*seg, (b+1)*seg)
// into partial[b] (fixed order), a single block then folds the partials -- deterministic for a
// given n, and HBM-rate instead of one block's latency (y^T y over 10^6 points: 220 us -> ~10 us).
__global__ void __launch_bounds__(256) reduce_segments_kernel(const double* __restrict__ a,
                                                              const double* __restrict__ b, int64_t n,
                                                              int64_t seg, double* __restrict__ partial) {
  __shared__ double sh[256];
  const int64_t lo = (int64_t)blockIdx.x * seg;
  const int64_t hi = lo + seg < n ? lo + seg : n;
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += b ? a[i] * b[i] : a[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void add_diagonal_kernel(double* A, int64_t n, int64_t ld, double v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i * ld + i] += v;
}

// Phi arrives with one triangle valid (cuBLAS "lower", column-major == row-major upper).
__global__ void symmetrize_kernel(double* A, int64_t n) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // col-major row index
  const int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;  // col-major col index
  // column-major element (r, c) lives at A[c * n + r]; the lower triangle r >= c is valid
  if (i < n && j < n && j > i) A[j * n + i] = A[i * n + j];
}

// B = AAT / noise + I in place; scalars[1] = trace(AAT / noise)
__global__ void scale_add_identity_kernel(double* A, int64_t n, double inv_noise) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int64_t r = idx / n, c = idx - r * n;
  double v = A[idx] * inv_noise;
  if (r == c) v += 1.0;
  A[idx] = v;
}

// scalars[1] = trace(A) * inv_noise (fixed summation order, one block)
__global__ void __launch_bounds__(1024) scaled_trace_kernel(const double* A, int64_t n,
                                                            double inv_noise, double* scalars) {
  __shared__ double sh[1024];
  double tr = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) tr += A[i * n + i] * inv_noise;
  sh[threadIdx.x] = tr;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) scalars[1] = sh[0];
}

// scalars[slot] = sum_i log A[i,i]
__global__ void __launch_bounds__(1024) log_diag_sum_kernel(const double* A, int64_t n, int64_t ld,
                                                            double* scalars, int slot) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += log(A[i * ld + i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) scalars[slot] = sh[0];
}

__global__ void scale_vector_kernel(double* v, int64_t n, double s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= s;
}

// gpflow 2.2.1 SGPR.elbo:
//   -N/2 log 2pi - sum log diag LB - N/2 log s2 - yty / (2 s2) + c^T c / 2
//   - sum K_diag / (2 s2) + tr(AAT) / 2
__global__ void sgpr_bound_kernel(const double* scalars, const double* c, int64_t m,
                                  const double* tail /* sum_kdiag, yty */, double n_total,
                                  double noise, double* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ctc = 0.0;
  for (int64_t i = 0; i < m; ++i) ctc += c[i] * c[i];
  const double logdet = scalars[0], trace = scalars[1];
  double bound = -0.5 * n_total * log(2.0 * M_PI);
  bound += -logdet;
  bound -= 0.5 * n_total * log(noise);
  bound += -0.5 * tail[1] / noise;
  bound += 0.5 * ctc;
  bound += -0.5 * tail[0] / noise;
  bound += 0.5 * trace;
  out[0] = bound;
  out[1] = logdet;
  out[2] = trace;
  out[3] = ctc;
}

// gpflow GPR.log_marginal_likelihood: -1/2 y^T alpha - sum log diag L - N/2 log 2pi
__global__ void gpr_lml_kernel(const double* scalars, const double* y, const double* alpha,
                               int64_t n, double* out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += y[i] * alpha[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = -0.5 * sh[0] - scalars[0] - 0.5 * (double)n * log(2.0 * M_PI);
}

// work layout shared by the finish functions: [int devInfo | pad][8 scalars][potrf workspace]
constexpr size_t kFinishHeader = 16 + 8 * sizeof(double);

// C(lower, column-major M x M) += A'^T A' with A' column-major (k x M, lda).  cuBLAS DSYRK runs
// at full-GEMM cost on this shape (measured on B200), so the triangle is cut recursively into
// off-diagonal rectangles (plain DGEMM, the efficient path) and small diagonal blocks:
//   OAK_SYRK_MODE=0 one DSYRK | 1,2,3 recursion depth | 9 one full DGEMM | 20 batched blocks
//                 | 30 own stream-K DMMA kernel (oak_syrk.cu, default)
static int syrk_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OAK_SYRK_MODE");
    v = e ? atoi(e) : 30;  // measured on B200 (scripts/quick_sgpr.py, profiles/): 30 = 56.7, 20 = 58.7, 1 = 76 ms / 10^6
  }
  return v;
}

static int syrk_rec(cublasHandle_t cb, int off, int m, int k, const double* A, int lda, double* C,
                    int ldc, int depth) {
  const double one = 1.0;
  if (depth == 0 || m < 128 || (m & 1)) {
    OAK_CUBLAS(cublasDsyrk(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, m, k, &one, A + (size_t)off * lda, lda,
                           &one, C + (size_t)off * ldc + off, ldc));
    g_launches.fetch_add(1);
    return 0;
  }
  const int h = m / 2;
  // C[off+h : off+m, off : off+h] += A'[:, off+h:off+m]^T A'[:, off:off+h]
  OAK_CUBLAS(cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, m - h, h, k, &one, A + (size_t)(off + h) * lda, lda,
                         A + (size_t)off * lda, lda, &one, C + (size_t)off * ldc + off + h, ldc));
  g_launches.fetch_add(1);
  if (int rc = syrk_rec(cb, off, h, k, A, lda, C, ldc, depth - 1)) return rc;
  return syrk_rec(cb, off + h, m - h, k, A, lda, C, ldc, depth - 1);
}

// cuBLAS alternative (OAK_SYRK_MODE=20): the lower triangle in 128 x 128 blocks as ONE batched DGEMM (diagonal blocks computed in
// full: 590 k instead of 525 k block entries at M = 1024, against 786 k for DGEMM + 2 DSYRK).
// Measured on B200 (profiles/): 2.4 ms per 1024 x 65536 chunk against 3.4 ms.  A ragged last
// block row (M % 128 rows) goes through one DGEMM + one small DSYRK.
struct BatchPtrs {
  const double** dev = nullptr;  // [3][kMaxBatch]: A_i, A_j, C_ij
  const double* A = nullptr;
  double* C = nullptr;
  int m = 0, lda = 0, count = 0;
};
constexpr int kMaxBatch = 8192;
constexpr int kSyrkBlock = 128;
static BatchPtrs g_batch[64];

static int syrk_batched(cublasHandle_t cb, int m, int k, const double* A, int lda, double* C,
                        cudaStream_t stream) {
  const double one = 1.0;
  const int bs = kSyrkBlock;
  const int nb = m / bs, rem = m - nb * bs;
  const int cnt = nb * (nb + 1) / 2;
  if (cnt > kMaxBatch) return -1;  // caller falls back
  int dev = 0;
  OAK_CUDA(cudaGetDevice(&dev));
  if (cnt > 0) {
    BatchPtrs& bp = g_batch[dev];
    if (!bp.dev) OAK_CUDA(cudaMalloc(&bp.dev, 3 * kMaxBatch * sizeof(double*)));
    if (bp.A != A || bp.C != C || bp.m != m || bp.lda != lda) {
      std::vector<const double*> h(3 * (size_t)cnt);
      int c = 0;
      for (int j = 0; j < nb; ++j)
        for (int i = j; i < nb; ++i, ++c) {
          h[c] = A + (size_t)i * bs * lda;                            // rows of the block (op T)
          h[cnt + c] = A + (size_t)j * bs * lda;                      // columns of the block
          h[2 * (size_t)cnt + c] = C + (size_t)j * bs * m + (size_t)i * bs;  // column-major, lower
        }
      // pageable source: the copy is staged before the call returns, `h` may die afterwards
      OAK_CUDA(cudaMemcpyAsync(bp.dev, h.data(), h.size() * sizeof(double*), cudaMemcpyHostToDevice, stream));
      bp.A = A; bp.C = C; bp.m = m; bp.lda = lda; bp.count = cnt;
    }
    OAK_CUBLAS(cublasDgemmBatched(cb, CUBLAS_OP_T, CUBLAS_OP_N, bs, bs, k, &one, bp.dev, lda, bp.dev + cnt, lda,
                                  &one, (double**)(bp.dev + 2 * (size_t)cnt), m, cnt));
    g_launches.fetch_add(1);
  }
  if (rem > 0) {
    const int off = nb * bs;
    if (off > 0) {
      OAK_CUBLAS(cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, rem, off, k, &one, A + (size_t)off * lda, lda, A, lda,
                             &one, C + off, m));
      g_launches.fetch_add(1);
    }
    OAK_CUBLAS(cublasDsyrk(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, rem, k, &one, A + (size_t)off * lda, lda, &one,
                           C + (size_t)off * m + off, m));
    g_launches.fetch_add(1);
  }
  return 0;
}

static int syrk_lower_accumulate(cublasHandle_t cb, int m, int k, double* A, int lda, double* C,
                                 double* partials, size_t partial_bytes, int device, cudaStream_t stream) {
  const int mode = syrk_mode();
  if (mode == 30) return syrk_lower_dmma(m, k, A, lda, C, partials, partial_bytes, device, stream);
  if (mode == 20) {
    const int rc = syrk_batched(cb, m, k, A, lda, C, stream);
    if (rc >= 0) return rc;
    return syrk_rec(cb, 0, m, k, A, lda, C, m, 1);
  }
  if (mode == 9) {
    const double one = 1.0;
    OAK_CUBLAS(cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, m, m, k, &one, A, lda, A, lda, &one, C, m));
    g_launches.fetch_add(1);
    return 0;
  }
  return syrk_rec(cb, 0, m, k, A, lda, C, m, mode);
}

static int potrf_lwork(cusolverDnHandle_t cs, int n, int* lwork) {
  OAK_CUSOLVER(cusolverDnDpotrf_bufferSize(cs, CUBLAS_FILL_MODE_LOWER, n, nullptr, n, lwork));
  return 0;
}

}  // namespace oak

using namespace oak;

extern "C" size_t oak_sgpr_stats_count(int64_t m) { return m < 0 ? 0 : (size_t)(m * m + m + 2); }

extern "C" size_t oak_sgpr_stats_work_bytes(int64_t m, int64_t chunk) {
  if (m < 0 || chunk < 0) return 0;
  return (size_t)(m * chunk + chunk) * sizeof(double) + syrk_dmma_work_bytes((int)m);
}

static int sgpr_stats_impl(const oak_spec* spec, const void* d_pointsZ, int64_t m, const void* d_pointsX,
                           const double* d_y, int64_t n_local, int64_t chunk, double* d_stats, void* d_work,
                           double* d_kuf_store, void* stream_) {
  OAK_REQUIRE(spec && d_pointsZ && d_stats && d_work, "oak_sgpr_stats_f64: null argument");
  OAK_REQUIRE(m >= 1, "oak_sgpr_stats_f64: need at least one inducing point");
  OAK_REQUIRE(n_local >= 0, "oak_sgpr_stats_f64: negative n");
  if (n_local == 0) return 0;
  OAK_REQUIRE(d_pointsX && d_y, "oak_sgpr_stats_f64: null data");
  const int T = tile_rows_for_depth(spec->depth);
  OAK_REQUIRE(chunk >= T && chunk % T == 0, "oak_sgpr_stats_f64: chunk must be a multiple of 64");
  OAK_REQUIRE(m <= INT32_MAX && chunk <= INT32_MAX, "oak_sgpr_stats_f64: size exceeds cuBLAS int");
  cudaStream_t stream = (cudaStream_t)stream_;
  cublasHandle_t cb;
  if (int rc = handles(&cb, nullptr, stream)) return rc;

  double* kuf_scratch = (double*)d_work;   // M x chunk, row-major, ld = chunk
  double* kdiag = kuf_scratch + m * chunk; // chunk
  double* partials = kdiag + chunk;        // stream-K partial tiles of the contraction
  const size_t partial_bytes = syrk_dmma_work_bytes((int)m);
  double* phi = d_stats;                   // M x M
  double* kufy = d_stats + m * m;          // M
  double* tail = kufy + m;                 // sum_kdiag, yty
  const double2* pz = (const double2*)d_pointsZ;
  const double2* px = (const double2*)d_pointsX;
  const int64_t m_pad = padded(m), n_pad = padded(n_local);
  const double one = 1.0;

  for (int64_t c0 = 0; c0 < n_local; c0 += chunk) {
    const int64_t nc = (n_local - c0 < chunk) ? (n_local - c0) : chunk;
    // the chunk's Kuf block: scratch, or its own slot when the caller keeps Kuf for the backward pass
    double* kuf = d_kuf_store ? d_kuf_store + (c0 / chunk) * m * chunk : kuf_scratch;
    // Kuf chunk = K(Z, X[c0:c0+nc])    (gpflow Kuf, oak/utils.py:184)
    if (int rc = gram_launch(spec, pz, m_pad, 0, m, px, n_pad, c0, c0 + nc, 0, kuf, chunk, stream))
      return rc;
    // Phi += Kuf Kuf^T.  Row-major (M x nc, ld=chunk) == column-major (nc x M, lda=chunk) A';
    // Phi = A'^T A'  ->  DSYRK(trans = T).  Only one triangle is updated.
    if (int rc = syrk_lower_accumulate(cb, (int)m, (int)nc, kuf, (int)chunk, phi, partials, partial_bytes, spec->device, stream)) return rc;
    // Kuf_y += Kuf y_chunk = A'^T y
    OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_T, (int)nc, (int)m, &one, kuf, (int)chunk, d_y + c0, 1, &one,
                           kufy, 1));
    g_launches.fetch_add(1);
    // sum K_diag(X)   (kernel(X, full_cov=False) in SGPR.elbo)
    if (int rc = gram_diag_launch(spec, px + c0, nc, n_pad, kdiag, stream)) return rc;
    reduce_accumulate_kernel<<<1, 1024, 0, stream>>>(kdiag, nullptr, nc, tail + 0);
    OAK_LAUNCHED();
  }
  // y^T y: two-level reduction through the (now free) K_diag scratch
  {
    const int64_t seg = 4096;
    int64_t blocks = (n_local + seg - 1) / seg;
    if (blocks > chunk) blocks = chunk;  // scratch capacity; segments grow instead
    const int64_t seg_len = (n_local + blocks - 1) / blocks;
    reduce_segments_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_y, d_y, n_local, seg_len, kdiag);
    OAK_LAUNCHED();
    reduce_accumulate_kernel<<<1, 1024, 0, stream>>>(kdiag, nullptr, blocks, tail + 1);
    OAK_LAUNCHED();
  }
  return 0;
}

extern "C" int oak_sgpr_stats_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m,
                                  const void* d_pointsX, const double* d_y, int64_t n_local,
                                  int64_t chunk, double* d_stats, void* d_work, void* stream_) {
  return sgpr_stats_impl(spec, d_pointsZ, m, d_pointsX, d_y, n_local, chunk, d_stats, d_work, nullptr, stream_);
}

// Same, keeping every chunk's Kuf block for the backward pass: chunk c (points [c*chunk, (c+1)*chunk))
// is written to d_kuf_store + c * m * chunk as an m x chunk row-major block (ld = chunk);
// d_kuf_store holds ceil(n_local / chunk) such blocks.
extern "C" int oak_sgpr_stats_keep_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m,
                                       const void* d_pointsX, const double* d_y, int64_t n_local,
                                       int64_t chunk, double* d_stats, void* d_work, double* d_kuf_store,
                                       void* stream_) {
  OAK_REQUIRE(d_kuf_store, "oak_sgpr_stats_keep_f64: null Kuf store");
  return sgpr_stats_impl(spec, d_pointsZ, m, d_pointsX, d_y, n_local, chunk, d_stats, d_work, d_kuf_store, stream_);
}

extern "C" size_t oak_sgpr_finish_work_bytes(int64_t m) {
  if (m < 1 || m > INT32_MAX) return 0;
  cusolverDnHandle_t cs;
  if (handles(nullptr, &cs, nullptr)) return 0;
  int lwork = 0;
  if (potrf_lwork(cs, (int)m, &lwork)) return 0;
  return kFinishHeader + (size_t)lwork * sizeof(double) + (size_t)m * sizeof(double);
}

extern "C" int oak_sgpr_finish_f64(double* d_Kuu, double* d_stats, int64_t m, int64_t n_total,
                                   double noise, double jitter, double* d_out, double* d_alpha,
                                   void* d_work, void* stream_) {
  OAK_REQUIRE(d_Kuu && d_stats && d_out && d_work, "oak_sgpr_finish_f64: null argument");
  OAK_REQUIRE(m >= 1 && m <= INT32_MAX, "oak_sgpr_finish_f64: bad M");
  OAK_REQUIRE(noise > 0.0, "oak_sgpr_finish_f64: likelihood variance must be positive");
  cudaStream_t stream = (cudaStream_t)stream_;
  cublasHandle_t cb;
  cusolverDnHandle_t cs;
  if (int rc = handles(&cb, &cs, stream)) return rc;
  int lwork = 0;
  if (int rc = potrf_lwork(cs, (int)m, &lwork)) return rc;

  int* info = (int*)d_work;
  double* scalars = (double*)((char*)d_work + 16);
  double* potrf_ws = scalars + 8;
  double* cvec = potrf_ws + lwork;  // M
  double* phi = d_stats;
  double* kufy = d_stats + m * m;
  double* tail = kufy + m;
  const int M = (int)m;
  const double one = 1.0;
  const double sigma = std::sqrt(noise);
  const int t256 = 256;
  const unsigned gM = (unsigned)((m + t256 - 1) / t256);

  // Kuu + jitter I ; L = chol(Kuu)        (oak/utils.py:185,188)
  add_diagonal_kernel<<<gM, t256, 0, stream>>>(d_Kuu, m, m, jitter);
  OAK_LAUNCHED();
  OAK_CUSOLVER(cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, M, d_Kuu, M, potrf_ws, lwork, info));
  int h_info = 0;
  OAK_CUDA(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, stream));
  OAK_CUDA(cudaStreamSynchronize(stream));
  OAK_REQUIRE(h_info == 0, "oak_sgpr_finish_f64: Cholesky of Kuu failed (not positive definite)");

  // AAT = L^-1 Phi L^-T / noise            (utils.py:189-190; A = L^-1 Kuf / sigma)
  dim3 b2(32, 8), g2((unsigned)((m + 31) / 32), (unsigned)((m + 7) / 8));
  symmetrize_kernel<<<g2, b2, 0, stream>>>(phi, m);
  OAK_LAUNCHED();
  OAK_CUBLAS(cublasDtrsm(cb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N,
                         CUBLAS_DIAG_NON_UNIT, M, M, &one, d_Kuu, M, phi, M));
  OAK_CUBLAS(cublasDtrsm(cb, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T,
                         CUBLAS_DIAG_NON_UNIT, M, M, &one, d_Kuu, M, phi, M));
  // B = AAT + I, trace(AAT); LB = chol(B)   (utils.py:190-193)
  scaled_trace_kernel<<<1, 1024, 0, stream>>>(phi, m, 1.0 / noise, scalars);
  OAK_LAUNCHED();
  scale_add_identity_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, stream>>>(phi, m, 1.0 / noise);
  OAK_LAUNCHED();
  OAK_CUSOLVER(cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, M, phi, M, potrf_ws, lwork, info));
  log_diag_sum_kernel<<<1, 1024, 0, stream>>>(phi, m, m, scalars, 0);
  OAK_LAUNCHED();
  // Aerr = L^-1 Kuf y / sigma ; c = LB^-1 Aerr / sigma     (utils.py:194-195)
  OAK_CUDA(cudaMemcpyAsync(cvec, kufy, m * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  OAK_CUBLAS(cublasDtrsv(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, d_Kuu, M,
                         cvec, 1));
  scale_vector_kernel<<<gM, t256, 0, stream>>>(cvec, m, 1.0 / sigma);
  OAK_LAUNCHED();
  OAK_CUBLAS(cublasDtrsv(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, phi, M,
                         cvec, 1));
  scale_vector_kernel<<<gM, t256, 0, stream>>>(cvec, m, 1.0 / sigma);
  OAK_LAUNCHED();
  sgpr_bound_kernel<<<1, 32, 0, stream>>>(scalars, cvec, m, tail, (double)n_total, noise, d_out);
  OAK_LAUNCHED();
  if (d_alpha) {
    // alpha = L^-T LB^-T c                  (utils.py:197-198)
    OAK_CUDA(cudaMemcpyAsync(d_alpha, cvec, m * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    OAK_CUBLAS(cublasDtrsv(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, phi, M,
                           d_alpha, 1));
    OAK_CUBLAS(cublasDtrsv(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, d_Kuu, M,
                           d_alpha, 1));
  }
  OAK_CUDA(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, stream));
  OAK_CUDA(cudaStreamSynchronize(stream));
  OAK_REQUIRE(h_info == 0, "oak_sgpr_finish_f64: Cholesky of B = A A^T + I failed");
  g_launches.fetch_add(8);  // cuSOLVER / cuBLAS launches of this tail (library kernels)
  return 0;
}

// =====================================================================================================
// Round-2 SGPR path: L = chol(Kuu) FIRST, then the statistics on a route chosen on the device.
//
//   oak_sgpr_factor_f64   Kuu tiles -> [L ; L^-T] by one bordered Cholesky (oak_chol.cu), a condition
//                         estimate ||Kuu||_1 * lambda_max(Kuu^-1) (three power iterations on L^-T L^-1) and the
//                         route flag: 0 = accumulate Phi = Kuf Kuf^T and whiten once in the tail (fast),
//                         1 = gpflow's operation order, A_r = L^-1 Kuf_r per chunk and Psi = sum A_r A_r^T
//                         (oak/utils.py:186-190) -- needed when Kuu is ill conditioned: the rounding of Phi is
//                         amplified by cond(Kuu) in the tail (2e-6 relative ELBO error at cond 2.6e8, 1e-13 at
//                         1e4), the whitened sum only by sqrt(cond).
//   oak_sgpr_stats2_f64   the chunk loop; the whitening product and the operand of the contraction are gated
//                         by the device flag, so no host decision (and no synchronisation) is needed.
//   oak_sgpr_finish2_f64  B = I + AAT assembled with the border row (L^-1 Kuf y) / noise, factored by the same
//                         bordered Cholesky, which leaves c = LB^-1 Aerr / sigma in the border; bound; alpha.
//                         Both `info` codes travel in out[4], out[5]: one read-back per evaluation.
//
// d_fac layout (doubles): column-major matrix with leading dimension LD = 2 Mp, Mp = roundup8(M), M columns:
// rows [0, M) = L (lower triangle), rows [Mp, Mp + M) = L^-T (upper triangle; equivalently the row-major view
// fac[r * LD + Mp + c] is L^-1); then a 16-double header and 4 Mp doubles of scratch.
namespace oak {

static inline int64_t fac_mp(int64_t m) { return (m + 7) / 8 * 8; }
static inline int64_t fac_ld(int64_t m) { return 2 * fac_mp(m); }
static inline int64_t lb_ld(int64_t m) { return fac_mp(m) + 8; }
constexpr int kHdrDoubles = 16;
// header: [0] cond estimate, [1] ||Kuu||_1, [2] lambda_max(Kuu^-1) estimate, [3] route, [4] info of chol(Kuu),
// [5] sum log diag L, [6] threshold; ints at (int*)(h + 8): [0] route, [1] info_L, [2] tile counter, [3] info_B
__host__ __device__ static inline int* hdr_ints(double* h) { return reinterpret_cast<int*>(h + 8); }

// one CTA per column j of the factor buffer: jitter, zero padding, unit border row, column 1-norm, start vector
__global__ void __launch_bounds__(256) fac_init_kernel(double* fac, int m, int mp, int64_t ld, double jitter,
                                                       double* colsum, double* w0) {
  __shared__ double sh[256];
  const int j = blockIdx.x;
  double* col = fac + (int64_t)j * ld;
  double acc = 0.0;
  for (int i = threadIdx.x; i < (int)ld; i += 256) {
    if (i < m) {
      double v = col[i];
      if (i == j) {
        v += jitter;
        col[i] = v;
      }
      acc += fabs(v);
    } else {
      col[i] = (i == mp + j) ? 1.0 : 0.0;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    colsum[j] = sh[0];
    // fixed +-1 pattern (Knuth multiplicative hash bit), normalised
    w0[j] = ((((unsigned)j * 2654435761u) >> 7) & 1u) ? rsqrt((double)m) : -rsqrt((double)m);
  }
}

__global__ void __launch_bounds__(256) route_kernel(double* hdr, const double* colsum, const double* w_lo,
                                                    const double* w_hi, int m, int force_route, double threshold) {
  __shared__ double s0[256], s1[256], s2[256];
  double mx = 0.0, a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < m; i += 256) {
    mx = fmax(mx, colsum[i]);
    a = fma(w_lo[i], w_lo[i], a);
    b = fma(w_hi[i], w_hi[i], b);
  }
  s0[threadIdx.x] = mx;
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      s0[threadIdx.x] = fmax(s0[threadIdx.x], s0[threadIdx.x + s]);
      s1[threadIdx.x] += s1[threadIdx.x + s];
      s2[threadIdx.x] += s2[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int* hi = hdr_ints(hdr);
    const double lam = sqrt(s2[0] / s1[0]);
    const double cond = s0[0] * lam;
    int route = force_route >= 0 ? (force_route != 0) : (cond >= threshold ? 1 : 0);  // NaN -> 0
    if (hi[1] != 0) route = force_route > 0 ? 1 : 0;  // failed factorisation: the tail reports it
    hdr[0] = cond;
    hdr[1] = s0[0];
    hdr[2] = lam;
    hdr[3] = (double)route;
    hdr[4] = (double)hi[1];
    hdr[6] = threshold;
    hi[0] = route;
  }
}

// B (lower, column-major ld = ldb) = src / noise + I, border row = v / noise, dvec = diag(src) / noise
__global__ void __launch_bounds__(256) assemble_B_kernel(const double* __restrict__ psi, const double* __restrict__ s2,
                                                         const int* __restrict__ route, const double* __restrict__ v_white,
                                                         const double* __restrict__ v_raw, int m, int border_row,
                                                         int64_t ldb, double inv_noise, double* __restrict__ LB,
                                                         double* __restrict__ dvec) {
  const int j = blockIdx.x;
  const bool whitened = *route != 0;
  const double* src = (whitened ? psi : s2) + (int64_t)j * m;
  const double* v = v_white;  // Kuf y is accumulated un-whitened on both routes; v_white = L^-1 Kuf y
  (void)v_raw;
  double* col = LB + (int64_t)j * ldb;
  for (int i = j + threadIdx.x; i < m; i += 256) {
    const double val = src[i] * inv_noise;
    if (i == j) {
      dvec[j] = val;
      col[i] = val + 1.0;
    } else {
      col[i] = val;
    }
  }
  if (threadIdx.x == 0) col[border_row] = v[j] * inv_noise;
}

// gpflow 2.2.1 SGPR.elbo from the factored pieces; out[0..7]
__global__ void __launch_bounds__(256) sgpr_bound2_kernel(const double* __restrict__ dvec, const double* __restrict__ LB,
                                                          int64_t ldb, int border_row, int m, const double* scal,
                                                          const double* tail, double n_total, double noise,
                                                          const double* hdr, double* out, double* cvec) {
  __shared__ double s0[256], s1[256];
  double tr = 0.0, cc = 0.0;
  for (int i = threadIdx.x; i < m; i += 256) {
    tr += dvec[i];
    const double c = LB[(int64_t)i * ldb + border_row];
    if (cvec) cvec[i] = c;
    cc = fma(c, c, cc);
  }
  s0[threadIdx.x] = tr;
  s1[threadIdx.x] = cc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      s0[threadIdx.x] += s0[threadIdx.x + s];
      s1[threadIdx.x] += s1[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double logdet = scal[0], trace = s0[0], ctc = s1[0];
    double bound = -0.5 * n_total * log(2.0 * M_PI);
    bound += -logdet;
    bound -= 0.5 * n_total * log(noise);
    bound += -0.5 * tail[1] / noise;
    bound += 0.5 * ctc;
    bound += -0.5 * tail[0] / noise;
    bound += 0.5 * trace;
    const int* hi = reinterpret_cast<const int*>(hdr + 8);
    out[0] = bound;
    out[1] = logdet;
    out[2] = trace;
    out[3] = ctc;
    out[4] = (double)hi[1];
    out[5] = (double)hi[3];
    out[6] = (double)hi[0];
    out[7] = hdr[0];
  }
}

}  // namespace oak

extern "C" size_t oak_sgpr_factor_count(int64_t m) {
  if (m < 1) return 0;
  return (size_t)(fac_ld(m) * m + kHdrDoubles + 4 * fac_mp(m));
}
extern "C" int64_t oak_sgpr_factor_ld(int64_t m) { return m < 1 ? 0 : fac_ld(m); }
extern "C" int64_t oak_sgpr_lb_ld(int64_t m) { return m < 1 ? 0 : lb_ld(m); }

// Kuu(iv, kernel) (oak/utils.py:185) + jitter I with the identity border, column 1-norms, start vector
static int factor_pre(const oak_spec* spec, const void* d_pointsZ, int64_t m, double jitter, double* d_fac,
                      cudaStream_t stream, int max_ctas = 0) {
  const int64_t mp = fac_mp(m), ld = fac_ld(m);
  double* hdr = d_fac + ld * m;
  double* scratch = hdr + kHdrDoubles;  // colsum | wa | wb | wc
  const double2* pz = (const double2*)d_pointsZ;
  // row-major with pitch LD == column-major (symmetric)
  if (int rc = gram_launch(spec, pz, padded(m), 0, m, pz, padded(m), 0, m, 1, d_fac, ld, stream, nullptr, 0, nullptr,
                           nullptr, max_ctas > 0 ? -max_ctas : 0))
    return rc;
  fac_init_kernel<<<(unsigned)m, 256, 0, stream>>>(d_fac, (int)m, (int)mp, ld, jitter, scratch, scratch + mp);
  OAK_LAUNCHED();
  return 0;
}

// [L ; L^-T], the condition estimate and the route flag; `cb` must be bound to `stream`.  max_ctas > 0: the
// factorisation runs on that many CTAs (next to the first chunk's Gram tiles, see oak_sgpr_factor_stats_f64).
static int factor_chol(const oak_spec* spec, int64_t m, int route, double cond_threshold, double* d_fac,
                       cublasHandle_t cb, cudaStream_t stream, int max_ctas) {
  const int64_t mp = fac_mp(m), ld = fac_ld(m);
  const int M = (int)m;
  double* hdr = d_fac + ld * m;
  double* scratch = hdr + kHdrDoubles;
  double *colsum = scratch, *wa = scratch + mp, *wb = scratch + 2 * mp, *wc = scratch + 3 * mp;
  // L = chol(Kuu + jitter I) (utils.py:188) with the identity as border rows
  if (int rc = chol_bordered(d_fac, ld, M, 2 * M, (int)(mp - m), 1, hdr_ints(hdr) + 1, hdr + 5, spec->device, stream,
                             max_ctas))
    return rc;
  // lambda_max(Kuu^-1) = ||L^-T||_2^2: three power iterations on U U^T, U = L^-T
  const double* U = d_fac + mp;
  const double one = 1.0, zero = 0.0;
  const int LDi = (int)ld;
  OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_T, M, M, &one, U, LDi, wa, 1, &zero, wb, 1));
  OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_N, M, M, &one, U, LDi, wb, 1, &zero, wc, 1));
  OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_T, M, M, &one, U, LDi, wc, 1, &zero, wb, 1));
  OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_N, M, M, &one, U, LDi, wb, 1, &zero, wa, 1));
  OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_T, M, M, &one, U, LDi, wa, 1, &zero, wb, 1));
  OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_N, M, M, &one, U, LDi, wb, 1, &zero, wc, 1));
  g_launches.fetch_add(6);
  route_kernel<<<1, 256, 0, stream>>>(hdr, colsum, wa, wc, M, route, cond_threshold > 0.0 ? cond_threshold : 3.0e5);
  OAK_LAUNCHED();
  return 0;
}

extern "C" int oak_sgpr_factor_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m, double jitter, int route,
                                   double cond_threshold, double* d_fac, void* stream_) {
  OAK_REQUIRE(spec && d_pointsZ && d_fac, "oak_sgpr_factor_f64: null argument");
  OAK_REQUIRE(m >= 1 && m <= INT32_MAX / 8, "oak_sgpr_factor_f64: bad M");
  OAK_REQUIRE(reinterpret_cast<uintptr_t>(d_fac) % 16 == 0, "oak_sgpr_factor_f64: d_fac must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  cublasHandle_t cb;
  if (int rc = handles(&cb, nullptr, stream)) return rc;
  if (int rc = factor_pre(spec, d_pointsZ, m, jitter, d_fac, stream)) return rc;
  return factor_chol(spec, m, route, cond_threshold, d_fac, cb, stream, 0);
}

constexpr int kKySegments = 64;

// Kuf y from the per-tile row sums the Gram tiles leave (oak_gram.cu, YDOT): two fixed-order levels
__global__ void __launch_bounds__(256) ky_reduce_kernel(const double* __restrict__ part, int tiles, int m,
                                                        double* __restrict__ part2) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int per = (tiles + kKySegments - 1) / kKySegments;
  const int t0 = blockIdx.y * per, t1 = min(tiles, t0 + per);
  double acc = 0.0;
  for (int t = t0; t < t1; ++t) acc += part[(int64_t)t * m + i];
  part2[(int64_t)blockIdx.y * m + i] = acc;
}
__global__ void __launch_bounds__(256) ky_final_kernel(const double* __restrict__ part2, int m, double* __restrict__ kufy) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  double acc = 0.0;
  for (int s = 0; s < kKySegments; ++s) acc += part2[(int64_t)s * m + i];
  kufy[i] += acc;
}

static size_t stats2_piece_bytes(int64_t m) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  const size_t a = syrk2_work_bytes((int)m, dev), b = syrk_dmma_work_bytes((int)m);
  return a > b ? a : b;
}

extern "C" size_t oak_sgpr_stats2_work_bytes(int64_t m, int64_t chunk) {
  if (m < 0 || chunk < 0) return 0;
  const int64_t me = (m + 1) / 2 * 2;
  return (size_t)(2 * m * chunk + chunk + (chunk / 64 + kKySegments) * me) * sizeof(double) + stats2_piece_bytes(m);
}

// Phi (route 0) or Psi = sum (L^-1 Kuf)(L^-1 Kuf)^T (route 1) | Kuf y | sum K_diag | y^T y for the local points,
// accumulated into d_stats; the route is read on the device from the header of d_fac.  Kuf y comes out of the
// Gram-tile epilogue (warp-shuffle row sums per tile, summed in a fixed order) when the depth allows it.
// d_kuf_store (nullable): keeps every chunk's Kuf block as in oak_sgpr_stats_keep_f64.
// sm_reserve / route_ready: the overlapped front -- the first chunk's Gram tiles (which do not depend on the
// factorisation) leave sm_reserve SMs to it, and the stream waits for route_ready before the first kernel that
// reads the route flag.
static int sgpr_stats2_impl(const oak_spec* spec, const void* d_pointsZ, int64_t m, double* d_fac,
                            const void* d_pointsX, const double* d_y, int64_t n_local, int64_t chunk,
                            double* d_stats, void* d_work, double* d_kuf_store, void* stream_, int sm_reserve,
                            cudaEvent_t route_ready, const std::function<int()>& after_first_tiles) {
  OAK_REQUIRE(spec && d_pointsZ && d_fac && d_stats && d_work, "oak_sgpr_stats2_f64: null argument");
  OAK_REQUIRE(m >= 1, "oak_sgpr_stats2_f64: need at least one inducing point");
  OAK_REQUIRE(n_local >= 0, "oak_sgpr_stats2_f64: negative n");
  if (n_local == 0) return 0;
  OAK_REQUIRE(d_pointsX && d_y, "oak_sgpr_stats2_f64: null data");
  const int T = tile_rows_for_depth(spec->depth);
  OAK_REQUIRE(chunk >= T && chunk % T == 0, "oak_sgpr_stats2_f64: chunk must be a multiple of 64");
  OAK_REQUIRE(m <= INT32_MAX && chunk <= INT32_MAX, "oak_sgpr_stats2_f64: size exceeds cuBLAS int");
  cudaStream_t stream = (cudaStream_t)stream_;
  cublasHandle_t cb;
  if (int rc = handles(&cb, nullptr, stream)) return rc;
  const int64_t mp = fac_mp(m), ld = fac_ld(m);
  double* hdr = d_fac + ld * m;
  int* d_route = hdr_ints(hdr);
  int* d_counter = hdr_ints(hdr) + 2;
  double* kuf_scratch = (double*)d_work;
  double* abuf = kuf_scratch + m * chunk;
  double* kdiag = abuf + m * chunk;
  const int64_t me = (m + 1) / 2 * 2;                   // keeps the partial tiles behind 16-byte aligned
  double* ypart = kdiag + chunk;                       // [chunk / 64][m] per-tile row sums of Kuf y
  double* ypart2 = ypart + (chunk / 64) * me;          // [kKySegments][m]
  double* partials = ypart2 + kKySegments * me;
  const size_t partial_bytes = stats2_piece_bytes(m);
  const bool fold_ky = gram_can_fold_ky(spec) && !(getenv("OAK_NO_FOLD_KY") && atoi(getenv("OAK_NO_FOLD_KY")));
  double* phi = d_stats;
  double* kufy = d_stats + m * m;
  double* tail = kufy + m;
  const double2* pz = (const double2*)d_pointsZ;
  const double2* px = (const double2*)d_pointsX;
  const int64_t m_pad = padded(m), n_pad = padded(n_local);
  for (int64_t c0 = 0; c0 < n_local; c0 += chunk) {
    const int64_t nc = (n_local - c0 < chunk) ? (n_local - c0) : chunk;
    double* kuf = d_kuf_store ? d_kuf_store + (c0 / chunk) * m * chunk : kuf_scratch;
    // Kuf chunk = K(Z, X[c0:c0+nc]) (gpflow Kuf, oak/utils.py:184); with depth <= 4 the tile epilogue also leaves
    // the per-tile row sums of Kuf y (no second pass over the 2 GB block)
    const bool fold = fold_ky;
    if (int rc = gram_launch(spec, pz, m_pad, 0, m, px, n_pad, c0, c0 + nc, 0, kuf, chunk, stream, nullptr, 0,
                             fold ? d_y + c0 : nullptr, fold ? ypart : nullptr, c0 == 0 ? sm_reserve : 0))
      return rc;
    if (c0 == 0 && after_first_tiles) {
      // the heavy kernel is in flight: now enqueue the side stream's work (which records route_ready) ...
      if (int rc = after_first_tiles()) return rc;
    }
    // ... and wait for it before the first kernel that reads the route flag
    if (c0 == 0 && route_ready) OAK_CUDA(cudaStreamWaitEvent(stream, route_ready, 0));
    // route 1: A_r = L^-1 Kuf_r (utils.py:189), a no-op launch otherwise
    if (int rc = panel_gemm_dmma(d_fac + mp, ld, kuf, chunk, abuf, chunk, (int)m, (int)m, nc, 1, nullptr, nullptr,
                                 d_route, d_counter, spec->device, stream))
      return rc;
    // Phi += Kuf Kuf^T (route 0) / Psi += A A^T (route 1).  OAK_SYRK_GEN=2 selects the 128 x 128 variant
    // (oak_syrk2.cu), measured slower on B200 (DMMA pipe 78 % against 88 %: one 8-warp CTA per SM idles the pipe
    // at every stage barrier; profiles/r02i_*), kept for A/B runs.
    static const int syrk_gen = getenv("OAK_SYRK_GEN") ? atoi(getenv("OAK_SYRK_GEN")) : 1;
    if (syrk_gen == 2) {
      if (int rc = syrk2_lower_dmma((int)m, nc, kuf, chunk, phi, nullptr, nullptr, partials, partial_bytes,
                                    spec->device, stream, abuf, d_route))
        return rc;
    } else if (int rc = syrk_lower_dmma((int)m, nc, kuf, chunk, phi, partials, partial_bytes, spec->device, stream,
                                        abuf, d_route)) {
      return rc;
    }
    // Kuf_y += Kuf y: always of the UN-whitened block (the tail applies L^-1 once)
    if (fold) {
      const int tiles_c = (int)((nc + 63) / 64);
      ky_reduce_kernel<<<dim3((unsigned)((m + 255) / 256), kKySegments), 256, 0, stream>>>(ypart, tiles_c, (int)m, ypart2);
      OAK_LAUNCHED();
      ky_final_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(ypart2, (int)m, kufy);
      OAK_LAUNCHED();
    } else {
      const double one = 1.0;
      OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_T, (int)nc, (int)m, &one, kuf, (int)chunk, d_y + c0, 1, &one, kufy, 1));
      g_launches.fetch_add(1);
    }
    if (int rc = gram_diag_launch(spec, px + c0, nc, n_pad, kdiag, stream)) return rc;
    reduce_accumulate_kernel<<<1, 1024, 0, stream>>>(kdiag, nullptr, nc, tail + 0);
    OAK_LAUNCHED();
  }
  {
    const int64_t seg = 4096;
    int64_t blocks = (n_local + seg - 1) / seg;
    if (blocks > chunk) blocks = chunk;
    const int64_t seg_len = (n_local + blocks - 1) / blocks;
    reduce_segments_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_y, d_y, n_local, seg_len, kdiag);
    OAK_LAUNCHED();
    reduce_accumulate_kernel<<<1, 1024, 0, stream>>>(kdiag, nullptr, blocks, tail + 1);
    OAK_LAUNCHED();
  }
  return 0;
}

extern "C" int oak_sgpr_stats2_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m, double* d_fac,
                                   const void* d_pointsX, const double* d_y, int64_t n_local, int64_t chunk,
                                   double* d_stats, void* d_work, double* d_kuf_store, void* stream_) {
  return sgpr_stats2_impl(spec, d_pointsZ, m, d_fac, d_pointsX, d_y, n_local, chunk, d_stats, d_work, d_kuf_store,
                          stream_, 0, nullptr, nullptr);
}

// oak_sgpr_factor_f64 + oak_sgpr_stats2_f64 as ONE call that hides the factorisation: the first chunk's Kuf tiles --
// which need neither L nor the route flag -- are launched first and leave `overlap_ctas` SMs free; the Kuu tiles,
// [L ; L^-T], the condition estimate and the route flag run on an internal side stream capped at that many CTAs, and
// `stream` joins the side stream before the first kernel that reads the flag.  The bordered Cholesky is latency bound
// (1024 dependent pivots: 0.4 ms on a whole GPU it cannot fill), the tiles are throughput bound, so lending the
// factorisation 8 of 148 SMs costs the tiles 5 % of ONE chunk instead of 0.5 ms on the critical path of every
// evaluation -- which is replicated on every rank (it is what held the 8-GPU ELBO below linear scaling).
// overlap_ctas: 0 = serial (exactly the two calls), > 0 = that many CTAs, < 0 = automatic (4 or 8 when the first
// chunk is long enough to cover the slower factorisation, else serial).  On return all work is ordered on `stream`.
extern "C" int oak_sgpr_factor_stats_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m, double jitter,
                                         int route, double cond_threshold, double* d_fac, const void* d_pointsX,
                                         const double* d_y, int64_t n_local, int64_t chunk, double* d_stats,
                                         void* d_work, double* d_kuf_store, int overlap_ctas, void* stream_) {
  OAK_REQUIRE(spec && d_pointsZ && d_fac, "oak_sgpr_factor_stats_f64: null argument");
  OAK_REQUIRE(m >= 1 && m <= INT32_MAX / 8, "oak_sgpr_factor_stats_f64: bad M");
  OAK_REQUIRE(reinterpret_cast<uintptr_t>(d_fac) % 16 == 0, "oak_sgpr_factor_stats_f64: d_fac must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  static const int env_overlap = getenv("OAK_SGPR_OVERLAP") ? atoi(getenv("OAK_SGPR_OVERLAP")) : -1;
  if (overlap_ctas < 0) {
    // the factorisation on 8 CTAs takes ~3x its whole-GPU time; the tiles of the first chunk must outlast it:
    // m * nc Gram entries at ~5e10 entries/s against ~1.6 us * m (1.6 ms at M = 1024)
    const int64_t nc = n_local < chunk ? n_local : chunk;
    // measured at M = 1024 (profiles/r02r_*): the factorisation takes ~2.0 / 2.7 / 3.9 ms on 8 / 6 / 4 CTAs, the first
    // chunk's tiles 2.5 ms per 125 000 points -- 8 CTAs leave a margin on short chunks, 4 are enough on long ones
    // (depth <= 4: the one-CTA-per-SM tile kernel, which leaves whole SMs free when its grid is capped; the deeper
    // geometries are not measured and stay serial)
    overlap_ctas = env_overlap >= 0 ? env_overlap
                                    : (m < 256 || nc < 100000 || spec->depth > 4 ? 0 : (nc >= 250000 ? 4 : 8));
  }
  if (overlap_ctas == 0 || n_local <= 0) {
    if (int rc = oak_sgpr_factor_f64(spec, d_pointsZ, m, jitter, route, cond_threshold, d_fac, stream_)) return rc;
    return sgpr_stats2_impl(spec, d_pointsZ, m, d_fac, d_pointsX, d_y, n_local, chunk, d_stats, d_work, d_kuf_store,
                            stream_, 0, nullptr, nullptr);
  }
  SideLane* lane = nullptr;
  if (int rc = side_lane(&lane)) return rc;
  // the side stream starts from the prepared points (everything enqueued on `stream` so far) ...
  OAK_CUDA(cudaEventRecord(lane->fork, stream));
  OAK_CUDA(cudaStreamWaitEvent(lane->stream, lane->fork, 0));
  // ... but its kernels are enqueued AFTER the first chunk's tiles, so that the GPU starts on the heavy kernel at once
  // and the Kuu tiles, the factorisation and the estimate (all capped at overlap_ctas CTAs) fill the SMs it leaves
  auto side_work = [&]() -> int {
    if (int rc = factor_pre(spec, d_pointsZ, m, jitter, d_fac, lane->stream, overlap_ctas)) return rc;
    if (int rc = factor_chol(spec, m, route, cond_threshold, d_fac, lane->cublas, lane->stream, overlap_ctas)) return rc;
    OAK_CUDA(cudaEventRecord(lane->join, lane->stream));
    return 0;
  };
  return sgpr_stats2_impl(spec, d_pointsZ, m, d_fac, d_pointsX, d_y, n_local, chunk, d_stats, d_work, d_kuf_store,
                          stream_, overlap_ctas, lane->join, side_work);
}

extern "C" size_t oak_sgpr_finish2_work_bytes(int64_t m) {
  if (m < 1) return 0;
  return (size_t)(2 * m * m + 4 * fac_mp(m) + 16) * sizeof(double);
}

// d_LB: lb_ld(m) * m doubles, column-major: rows [0, M) = LB (lower), row Mp = c^T.
// d_out[8] = elbo, sum log diag LB, tr(AAT), c^T c, info of chol(Kuu), info of chol(B), route, cond estimate.
extern "C" int oak_sgpr_finish2_f64(double* d_fac, double* d_stats, int64_t m, int64_t n_total, double noise,
                                    double* d_out, double* d_alpha, double* d_LB, void* d_work, void* stream_) {
  OAK_REQUIRE(d_fac && d_stats && d_out && d_LB && d_work, "oak_sgpr_finish2_f64: null argument");
  OAK_REQUIRE(m >= 1 && m <= INT32_MAX / 8, "oak_sgpr_finish2_f64: bad M");
  OAK_REQUIRE(noise > 0.0, "oak_sgpr_finish2_f64: likelihood variance must be positive");
  cudaStream_t stream = (cudaStream_t)stream_;
  cublasHandle_t cb;
  if (int rc = handles(&cb, nullptr, stream)) return rc;
  int dev = 0;
  OAK_CUDA(cudaGetDevice(&dev));
  const int64_t mp = fac_mp(m), ld = fac_ld(m), ldb = lb_ld(m);
  const int M = (int)m;
  double* hdr = d_fac + ld * m;
  int* hi = hdr_ints(hdr);
  const double* U = d_fac + mp;  // L^-T, column-major upper
  double* W = (double*)d_work;
  double* S2 = W + m * m;
  double* vtmp = S2 + m * m;
  double* dvec = vtmp + mp;
  double* cvec = dvec + mp;
  double* scal = cvec + mp;  // [0] sum log diag LB
  double* phi = d_stats;
  double* kufy = d_stats + m * m;
  double* tail = kufy + m;
  const double one = 1.0, zero = 0.0;
  // route 0: sigma^2 AAT = L^-1 Phi L^-T = U^T (Phi U)   (utils.py:189-190 with A = L^-1 Kuf / sigma);
  // computed unconditionally (0.2 ms), the assemble kernel picks its source by the device flag
  OAK_CUBLAS(cublasDsymm(cb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, M, M, &one, phi, M, U, (int)ld, &zero, W, M));
  OAK_CUBLAS(cublasDtrmm(cb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, M, &one,
                         U, (int)ld, W, M, S2, M));
  // sigma Aerr = L^-1 Kuf y   (utils.py:194)
  OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_T, M, M, &one, U, (int)ld, kufy, 1, &zero, vtmp, 1));
  g_launches.fetch_add(3);
  // B = AAT + I (utils.py:190-191) with the border row Aerr / sigma; LB = chol(B) leaves c in the border
  assemble_B_kernel<<<(unsigned)m, 256, 0, stream>>>(phi, S2, hi, vtmp, kufy, M, (int)mp, ldb, 1.0 / noise, d_LB, dvec);
  OAK_LAUNCHED();
  if (int rc = chol_bordered(d_LB, ldb, M, M + 1, (int)(mp - m), 0, hi + 3, scal, dev, stream)) return rc;
  sgpr_bound2_kernel<<<1, 256, 0, stream>>>(dvec, d_LB, ldb, (int)mp, M, scal, tail, (double)n_total, noise, hdr,
                                            d_out, cvec);
  OAK_LAUNCHED();
  if (d_alpha) {
    // alpha = L^-T LB^-T c (utils.py:197-198)
    OAK_CUBLAS(cublasDtrsv(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, d_LB, (int)ldb, cvec, 1));
    OAK_CUBLAS(cublasDgemv(cb, CUBLAS_OP_N, M, M, &one, U, (int)ld, cvec, 1, &zero, d_alpha, 1));
    g_launches.fetch_add(2);
  }
  return 0;
}

extern "C" size_t oak_gpr_finish_work_bytes(int64_t n) {
  if (n < 1 || n > INT32_MAX) return 0;
  cusolverDnHandle_t cs;
  if (handles(nullptr, &cs, nullptr)) return 0;
  int lwork = 0;
  if (potrf_lwork(cs, (int)n, &lwork)) return 0;
  return kFinishHeader + (size_t)lwork * sizeof(double);
}

extern "C" int oak_gpr_finish_f64(double* d_K, const double* d_y, int64_t n, double noise,
                                  double* d_lml, double* d_alpha, void* d_work, void* stream_) {
  OAK_REQUIRE(d_K && d_y && d_lml && d_alpha && d_work, "oak_gpr_finish_f64: null argument");
  OAK_REQUIRE(n >= 1 && n <= INT32_MAX, "oak_gpr_finish_f64: bad n");
  cudaStream_t stream = (cudaStream_t)stream_;
  cusolverDnHandle_t cs;
  if (int rc = handles(nullptr, &cs, stream)) return rc;
  int lwork = 0;
  if (int rc = potrf_lwork(cs, (int)n, &lwork)) return rc;
  int* info = (int*)d_work;
  double* scalars = (double*)((char*)d_work + 16);
  double* potrf_ws = scalars + 8;
  const int N = (int)n;
  // K + noise I ; L = chol ; alpha = cholesky_solve(L, y)     (oak/utils.py:208-211)
  add_diagonal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_K, n, n, noise);
  OAK_LAUNCHED();
  OAK_CUSOLVER(cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, N, d_K, N, potrf_ws, lwork, info));
  OAK_CUDA(cudaMemcpyAsync(d_alpha, d_y, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  OAK_CUSOLVER(cusolverDnDpotrs(cs, CUBLAS_FILL_MODE_LOWER, N, 1, d_K, N, d_alpha, N, info + 1));
  log_diag_sum_kernel<<<1, 1024, 0, stream>>>(d_K, n, n, scalars, 0);
  OAK_LAUNCHED();
  gpr_lml_kernel<<<1, 1024, 0, stream>>>(scalars, d_y, d_alpha, n, d_lml);
  OAK_LAUNCHED();
  int h_info = 0;
  OAK_CUDA(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, stream));
  OAK_CUDA(cudaStreamSynchronize(stream));
  OAK_REQUIRE(h_info == 0, "oak_gpr_finish_f64: Cholesky of K + noise I failed (not positive definite)");
  g_launches.fetch_add(2);
  return 0;
}
