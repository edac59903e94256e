// SGPR statistics and the dense M^3 / N^3 tails.
//
// oak_sgpr_stats_f64 replaces the Kuf build and the A A^T / A err contractions of gpflow's
// SGPR.elbo as re-derived in oak/utils.py:180-191.  The local N points are streamed in chunks: the
// fused Gram tile kernel writes an M x chunk block of Kuf (sized to stay L2 resident), cuBLAS
// DSYRK folds it into Phi = Kuf Kuf^T and DGEMV into Kuf y; K_diag and y^T y are reduced by
// deterministic single-block kernels.  Phi | Kuf y | sum K_diag | y^T y is one contiguous vector:
// the only thing that has to be all-reduced across ranks.
//
// oak_sgpr_finish_f64 / oak_gpr_finish_f64 are the dense tails (cuSOLVER potrf, cuBLAS
// trsm/trsv), timed separately from the tile kernels.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cmath>
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
