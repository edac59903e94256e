// Host-buffer entry point: the call a NumPy user of OAKKernel.K makes (oak/oak_kernel.py:251-265
// takes and returns host arrays).  X (and X2) are copied to the device, prepared once, and K is
// produced in row blocks that are streamed back to the host on a second stream while the next
// block is being computed.  Host buffers should be pinned for the copies to overlap.
#include <mutex>
#include <thread>

#include "oak_common.cuh"

using namespace oak;

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

extern "C" size_t oak_gram_host_work_bytes(const oak_spec* spec, int64_t n, int64_t n2, int64_t ldx,
                                           int64_t block_rows) {
  if (!spec || n < 0 || n2 < 0 || block_rows < 1) return 0;
  size_t b = 0;
  b += align256((size_t)n * ldx * sizeof(double));
  b += align256(oak_points_bytes(spec, n));
  if (n2 > 0) {
    b += align256((size_t)n2 * ldx * sizeof(double));
    b += align256(oak_points_bytes(spec, n2));
  }
  const int64_t cols = n2 > 0 ? n2 : n;
  b += 2 * align256((size_t)block_rows * cols * sizeof(double));
  return b;
}

extern "C" int oak_gram_host_f64(const oak_spec* spec, const double* h_X, int64_t n,
                                 const double* h_X2, int64_t n2, int64_t ldx, double* h_K,
                                 int64_t ldk, int64_t block_rows, void* d_work, void* stream_) {
  OAK_REQUIRE(spec && h_X && h_K && d_work, "oak_gram_host_f64: null argument");
  OAK_REQUIRE(n >= 0 && ldx >= 1, "oak_gram_host_f64: bad shape");
  const int T = tile_rows_for_depth(spec->depth);
  OAK_REQUIRE(block_rows >= T && block_rows % T == 0,
              "oak_gram_host_f64: block_rows must be a multiple of 64");
  const bool same = (h_X2 == nullptr);
  if (same) n2 = 0;
  const int64_t cols = same ? n : n2;
  OAK_REQUIRE(ldk >= cols, "oak_gram_host_f64: ldk smaller than the number of columns");
  if (n == 0 || cols == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;

  char* w = (char*)d_work;
  double* dX = (double*)w;
  w += align256((size_t)n * ldx * sizeof(double));
  void* pX = w;
  w += align256(oak_points_bytes(spec, n));
  double* dX2 = nullptr;
  void* pX2 = nullptr;
  if (!same) {
    dX2 = (double*)w;
    w += align256((size_t)n2 * ldx * sizeof(double));
    pX2 = w;
    w += align256(oak_points_bytes(spec, n2));
  }
  double* blk[2];
  blk[0] = (double*)w;
  w += align256((size_t)block_rows * cols * sizeof(double));
  blk[1] = (double*)w;

  OAK_CUDA(cudaMemcpyAsync(dX, h_X, (size_t)n * ldx * sizeof(double), cudaMemcpyHostToDevice, stream));
  if (int rc = oak_prepare_points_f64(spec, dX, n, ldx, pX, stream)) return rc;
  if (!same) {
    OAK_CUDA(cudaMemcpyAsync(dX2, h_X2, (size_t)n2 * ldx * sizeof(double), cudaMemcpyHostToDevice,
                             stream));
    if (int rc = oak_prepare_points_f64(spec, dX2, n2, ldx, pX2, stream)) return rc;
  }

  cudaStream_t copy_stream;
  OAK_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  cudaEvent_t computed[2], copied[2];
  for (int i = 0; i < 2; ++i) {
    OAK_CUDA(cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming));
    OAK_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
  }
  int rc = 0;
  int b = 0;
  for (int64_t r0 = 0; r0 < n && rc == 0; r0 += block_rows, b ^= 1) {
    const int64_t r1 = (r0 + block_rows < n) ? r0 + block_rows : n;
    // the previous D2H out of this buffer must have finished before it is overwritten
    if (cudaStreamWaitEvent(stream, copied[b], 0) != cudaSuccess) rc = 1;
    if (rc == 0)
      rc = gram_launch(spec, (const double2*)pX, padded(n), r0, r1,
                       same ? (const double2*)pX : (const double2*)pX2, padded(cols), 0, cols, 0,
                       blk[b], cols, stream);
    if (rc) break;
    cudaEventRecord(computed[b], stream);
    cudaStreamWaitEvent(copy_stream, computed[b], 0);
    if (cudaMemcpy2DAsync(h_K + r0 * ldk, (size_t)ldk * sizeof(double), blk[b],
                          (size_t)cols * sizeof(double), (size_t)cols * sizeof(double),
                          (size_t)(r1 - r0), cudaMemcpyDeviceToHost, copy_stream) != cudaSuccess) {
      set_error(std::string("oak_gram_host_f64: D2H copy failed: ") +
                cudaGetErrorString(cudaGetLastError()));
      rc = 1;
      break;
    }
    cudaEventRecord(copied[b], copy_stream);
  }
  cudaError_t e1 = cudaStreamSynchronize(copy_stream);
  cudaError_t e2 = cudaStreamSynchronize(stream);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(computed[i]);
    cudaEventDestroy(copied[i]);
  }
  cudaStreamDestroy(copy_stream);
  if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess)) {
    set_error(std::string("oak_gram_host_f64: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    rc = 1;
  }
  return rc;
}

// ---- symmetric K(X, X) through host buffers: lower trapezoid only ---------------------------------------
// The full-matrix call above is PCIe-bound (8 N^2 bytes back to the host).  For the symmetric Gram only the
// lower triangle carries information: rows [row_begin, row_end) are produced in blocks of `block_rows`, each
// block evaluates the tiles left of / on the diagonal (mode 2 of the tile kernel) and only its columns
// [0, block end) travel back -- half the bytes for the whole matrix.  Entries right of the diagonal inside a
// row block are unspecified unless `mirror` is set (full row range only), in which case the strict upper
// triangle is filled on the HOST from the lower one after the copies (multi-threaded blocked transpose).
namespace {
struct HostPipe {
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t computed[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
};
HostPipe* host_pipe(int device) {
  static HostPipe pipes[64];
  static std::mutex mu;
  if (device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  HostPipe& p = pipes[device];
  if (!p.copy_stream) {
    if (cudaStreamCreateWithFlags(&p.copy_stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < 2; ++i)
      if (cudaEventCreateWithFlags(&p.computed[i], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&p.copied[i], cudaEventDisableTiming) != cudaSuccess)
        return nullptr;
  }
  return &p;
}

// K[i][j] = K[j][i] for i < j < n, blocked; thread t takes the block rows t, t + T, ...
void mirror_upper_on_host(double* K, int64_t n, int64_t ldk) {
  constexpr int64_t B = 64;
  const int64_t nb = (n + B - 1) / B;
  unsigned T = std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if (T > 64) T = 64;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < T; ++t)
    pool.emplace_back([=]() {
      for (int64_t bi = t; bi < nb; bi += T)
        for (int64_t bj = bi; bj < nb; ++bj) {
          const int64_t i1 = (bi + 1) * B < n ? (bi + 1) * B : n, j1 = (bj + 1) * B < n ? (bj + 1) * B : n;
          for (int64_t i = bi * B; i < i1; ++i)
            for (int64_t j = (bj == bi ? i + 1 : bj * B); j < j1; ++j) K[i * ldk + j] = K[j * ldk + i];
        }
    });
  for (auto& th : pool) th.join();
}
}  // namespace

extern "C" size_t oak_gram_host_lower_work_bytes(const oak_spec* spec, int64_t n, int64_t ldx, int64_t row_end,
                                                 int64_t block_rows) {
  if (!spec || n < 0 || block_rows < 1 || row_end < 0) return 0;
  return align256((size_t)n * ldx * sizeof(double)) + align256(oak_points_bytes(spec, n)) +
         2 * align256((size_t)block_rows * row_end * sizeof(double));
}

extern "C" int oak_gram_host_lower_f64(const oak_spec* spec, const double* h_X, int64_t n, int64_t ldx,
                                       int64_t row_begin, int64_t row_end, double* h_K, int64_t ldk,
                                       int64_t block_rows, int mirror, void* d_work, void* stream_) {
  OAK_REQUIRE(spec && h_X && h_K && d_work, "oak_gram_host_lower_f64: null argument");
  OAK_REQUIRE(n >= 0 && ldx >= 1, "oak_gram_host_lower_f64: bad shape");
  OAK_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n, "oak_gram_host_lower_f64: bad row range");
  const int T = tile_rows_for_depth(spec->depth);
  OAK_REQUIRE(block_rows >= T && block_rows % T == 0 && row_begin % T == 0,
              "oak_gram_host_lower_f64: block_rows and row_begin must be multiples of 64");
  OAK_REQUIRE(ldk >= row_end, "oak_gram_host_lower_f64: ldk smaller than row_end");
  OAK_REQUIRE(!mirror || (row_begin == 0 && row_end == n), "oak_gram_host_lower_f64: mirror needs the full row range");
  if (row_end == row_begin) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  HostPipe* pipe = host_pipe(spec->device);
  OAK_REQUIRE(pipe, "oak_gram_host_lower_f64: could not create the copy stream");

  char* w = (char*)d_work;
  double* dX = (double*)w;
  w += align256((size_t)n * ldx * sizeof(double));
  void* pX = w;
  w += align256(oak_points_bytes(spec, n));
  double* blk[2];
  blk[0] = (double*)w;
  w += align256((size_t)block_rows * row_end * sizeof(double));
  blk[1] = (double*)w;

  OAK_CUDA(cudaMemcpyAsync(dX, h_X, (size_t)n * ldx * sizeof(double), cudaMemcpyHostToDevice, stream));
  if (int rc = oak_prepare_points_f64(spec, dX, n, ldx, pX, stream)) return rc;
  // a previous call's copies out of the block buffers are complete (every call synchronises before returning)
  int rc = 0, b = 0;
  bool used[2] = {false, false};
  for (int64_t r0 = row_begin; r0 < row_end && rc == 0; r0 += block_rows, b ^= 1) {
    const int64_t r1 = (r0 + block_rows < row_end) ? r0 + block_rows : row_end;
    if (used[b] && cudaStreamWaitEvent(stream, pipe->copied[b], 0) != cudaSuccess) rc = 1;
    if (rc == 0)
      rc = gram_launch(spec, (const double2*)pX, padded(n), r0, r1, (const double2*)pX, padded(n), 0, r1, 2, blk[b],
                       r1, stream);
    if (rc) break;
    cudaEventRecord(pipe->computed[b], stream);
    cudaStreamWaitEvent(pipe->copy_stream, pipe->computed[b], 0);
    if (cudaMemcpy2DAsync(h_K + (r0 - row_begin) * ldk, (size_t)ldk * sizeof(double), blk[b], (size_t)r1 * sizeof(double),
                          (size_t)r1 * sizeof(double), (size_t)(r1 - r0), cudaMemcpyDeviceToHost,
                          pipe->copy_stream) != cudaSuccess) {
      set_error(std::string("oak_gram_host_lower_f64: D2H copy failed: ") + cudaGetErrorString(cudaGetLastError()));
      rc = 1;
      break;
    }
    cudaEventRecord(pipe->copied[b], pipe->copy_stream);
    used[b] = true;
  }
  cudaError_t e1 = cudaStreamSynchronize(pipe->copy_stream);
  cudaError_t e2 = cudaStreamSynchronize(stream);
  if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess)) {
    set_error(std::string("oak_gram_host_lower_f64: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    rc = 1;
  }
  if (rc == 0 && mirror) mirror_upper_on_host(h_K, n, ldk);
  return rc;
}
