// Host-buffer entry point: the call a NumPy user of OAKKernel.K makes (oak/oak_kernel.py:251-265
// takes and returns host arrays).  X (and X2) are copied to the device, prepared once, and K is
// produced in row blocks that are streamed back to the host on a second stream while the next
// block is being computed.  Host buffers should be pinned for the copies to overlap.
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
