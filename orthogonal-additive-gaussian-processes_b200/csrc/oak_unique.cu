// Input pipeline on the device: sorted distinct values of a column with their counts, and
// column means -- what oak_model.fit computes on the host with np.unique / ndarray.mean per column:
//   * empirical-measure locations and weights    oak/model_utils.py:334-344   (np.unique(X_scaled[:, ii], return_counts=True))
//   * p of a categorical feature                 oak/model_utils.py:736-739   (frequency of every distinct level)
//   * p0 of a binary feature                     oak/model_utils.py:731       (1 - X[:, j].mean())
// Sort + run-length, all integer / comparison work on order-preserving 64-bit keys, so the distinct values are
// the input bit patterns and the counts are exact: the weights count / n are bit-identical to NumPy's.
//   1. encode: key = order_key(x) (monotone uint64), padded with UINT64_MAX to a power of two
//   2. bitonic sort: compare-exchange passes; every stride below 2048 of a merge runs in shared memory
//      (one launch per merge instead of eleven), larger strides stream through HBM (45 passes at n = 10^6)
//   3. heads: flag key[i] != key[i-1], per-block counts, one-block scan of the block counts, scatter of the head
//      positions; count = distance to the next head
// NaNs sort last (positive NaN patterns are the largest keys below the padding), as in np.unique.
#include "oak_common.cuh"

namespace oak {

namespace uniq {
constexpr int kBlock = 1024;               // threads of the shared-memory sort kernel
constexpr int kSpan = 2 * kBlock;          // keys per block there
constexpr unsigned long long kPad = ~0ull;
}  // namespace uniq

__global__ void uniq_encode_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int64_t col,
                                   unsigned long long* __restrict__ keys, int64_t n_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  keys[i] = i < n ? order_key(X[i * ldx + col] + 0.0) : uniq::kPad;  // -0.0 -> +0.0: one value, as for np.unique
}

__device__ __forceinline__ void uniq_cmpx(unsigned long long& a, unsigned long long& b, bool ascending) {
  if ((a > b) == ascending) {
    const unsigned long long t = a;
    a = b;
    b = t;
  }
}

// all compare-exchange stages with stride < kSpan of the merges k = k_first .. k_last (powers of two)
__global__ void __launch_bounds__(uniq::kBlock) uniq_sort_shared_kernel(unsigned long long* __restrict__ keys,
                                                                         int64_t k_first, int64_t k_last) {
  using namespace uniq;
  __shared__ unsigned long long sh[kSpan];
  const int64_t base = (int64_t)blockIdx.x * kSpan;
  const int t = threadIdx.x;
  sh[t] = keys[base + t];
  sh[t + kBlock] = keys[base + t + kBlock];
  __syncthreads();
  for (int64_t k = k_first; k <= k_last; k <<= 1) {
    for (int j = (int)(k / 2 < kBlock ? k / 2 : kBlock); j > 0; j >>= 1) {
      const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // t-th pair of stride j
      const bool asc = (((base + lo) & k) == 0);
      uniq_cmpx(sh[lo], sh[lo + j], asc);
      __syncthreads();
    }
  }
  keys[base + t] = sh[t];
  keys[base + t + kBlock] = sh[t + kBlock];
}

// one compare-exchange stage with stride j >= kSpan of merge k
__global__ void uniq_sort_global_kernel(unsigned long long* __restrict__ keys, int64_t n_pad, int64_t j, int64_t k) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pad / 2) return;
  const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  unsigned long long a = keys[lo], b = keys[lo + j];
  const bool asc = ((lo & k) == 0);
  if ((a > b) == asc) {
    keys[lo] = b;
    keys[lo + j] = a;
  }
}

__global__ void __launch_bounds__(1024) uniq_heads_count_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                                 int* __restrict__ block_heads) {
  __shared__ int sh[32];
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  const int head = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
  int v = head;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = sh[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_heads[blockIdx.x] = v;
  }
}

// exclusive scan of the block counts (one block, fixed order); total -> *num
__global__ void __launch_bounds__(1024) uniq_scan_kernel(int* __restrict__ block_heads, int blocks, int* __restrict__ num) {
  __shared__ int sh[1024];
  int carry = 0;
  for (int b0 = 0; b0 < blocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = i < blocks ? block_heads[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int add = (int)threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < blocks) block_heads[i] = carry + sh[threadIdx.x] - v;
    carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *num = carry;
}

__global__ void __launch_bounds__(1024) uniq_scatter_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                             const int* __restrict__ block_offset,
                                                             double* __restrict__ vals, int64_t* __restrict__ starts) {
  __shared__ int sh[32];
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  const int head = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
  // rank of this head inside the block: warp ballot + scan of the warp totals
  const unsigned m = __ballot_sync(0xffffffffu, head);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int in_warp = __popc(m & ((1u << lane) - 1u));
  if (lane == 0) sh[warp] = __popc(m);
  __syncthreads();
  if (warp == 0) {
    int v = sh[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    sh[lane] = v - sh[lane];  // exclusive
  }
  __syncthreads();
  if (head) {
    const int r = block_offset[blockIdx.x] + sh[warp] + in_warp;
    vals[r] = order_key_decode(keys[i]);
    starts[r] = i;
  }
}

__global__ void uniq_counts_kernel(const int64_t* __restrict__ starts, const int* __restrict__ num, int64_t n,
                                   int64_t* __restrict__ counts) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int u = *num;
  if (r >= u) return;
  counts[r] = (r + 1 < u ? starts[r + 1] : n) - starts[r];
}

// column sum in a fixed order (two levels): out[0] = sum_i X[i, col]
__global__ void __launch_bounds__(256) colsum_partial_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int64_t col,
                                                             int64_t seg, double* __restrict__ partial) {
  __shared__ double sh[256];
  const int64_t lo = (int64_t)blockIdx.x * seg, hi = lo + seg < n ? lo + seg : n;
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += X[i * ldx + col];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const double* __restrict__ partial, int blocks, double n_rows,
                                                           double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 256) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0] / n_rows;  // a division, as ndarray.mean: bit-identical for 0/1 columns
}

static int64_t uniq_pad(int64_t n) {
  int64_t p = uniq::kSpan;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace oak

using namespace oak;

extern "C" size_t oak_column_unique_work_bytes(int64_t n) {
  if (n < 0) return 0;
  const int64_t n_pad = uniq_pad(n);
  const int64_t blocks = (n + 1023) / 1024 + 1;
  return (size_t)n_pad * sizeof(unsigned long long) + (size_t)n * sizeof(int64_t) + (size_t)blocks * sizeof(int) + 256;
}

// Sorted distinct values of column `col` of the row-major (n x ldx) device matrix and their multiplicities:
// np.unique(X[:, col], return_counts=True).  d_vals[n], d_counts[n] (the first *d_num entries are written).
extern "C" int oak_column_unique_f64(const double* d_X, int64_t n, int64_t ldx, int64_t col, double* d_vals,
                                     int64_t* d_counts, int32_t* d_num, void* d_work, void* stream_) {
  OAK_REQUIRE(d_X && d_vals && d_counts && d_num && d_work, "oak_column_unique_f64: null argument");
  OAK_REQUIRE(n >= 0 && ldx >= 1 && col >= 0 && col < ldx, "oak_column_unique_f64: bad shape");
  OAK_REQUIRE(n < (1ll << 31), "oak_column_unique_f64: more than 2^31 rows");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) {
    OAK_CUDA(cudaMemsetAsync(d_num, 0, sizeof(int32_t), stream));
    return 0;
  }
  using namespace uniq;
  const int64_t n_pad = uniq_pad(n);
  unsigned long long* keys = (unsigned long long*)d_work;
  int64_t* starts = (int64_t*)(keys + n_pad);
  int* block_heads = (int*)(starts + n);
  uniq_encode_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, stream>>>(d_X, n, ldx, col, keys, n_pad);
  OAK_LAUNCHED();
  // bitonic network: merges k = 2 .. n_pad; strides >= kSpan through HBM, the rest of each merge in shared memory
  uniq_sort_shared_kernel<<<(unsigned)(n_pad / kSpan), kBlock, 0, stream>>>(keys, 2, kSpan);
  OAK_LAUNCHED();
  for (int64_t k = 2 * kSpan; k <= n_pad; k <<= 1) {
    for (int64_t j = k / 2; j >= kSpan; j >>= 1) {
      uniq_sort_global_kernel<<<(unsigned)((n_pad / 2 + 255) / 256), 256, 0, stream>>>(keys, n_pad, j, k);
      OAK_LAUNCHED();
    }
    uniq_sort_shared_kernel<<<(unsigned)(n_pad / kSpan), kBlock, 0, stream>>>(keys, k, k);
    OAK_LAUNCHED();
  }
  const int blocks = (int)((n + 1023) / 1024);
  uniq_heads_count_kernel<<<blocks, 1024, 0, stream>>>(keys, n, block_heads);
  OAK_LAUNCHED();
  uniq_scan_kernel<<<1, 1024, 0, stream>>>(block_heads, blocks, d_num);
  OAK_LAUNCHED();
  uniq_scatter_kernel<<<blocks, 1024, 0, stream>>>(keys, n, block_heads, d_vals, starts);
  OAK_LAUNCHED();
  uniq_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(starts, d_num, n, d_counts);
  OAK_LAUNCHED();
  return 0;
}

// d_out[0] = mean of column `col` (fixed summation order; exact for 0/1 columns: p0 = 1 - mean, model_utils.py:731).
// d_work: ceil(n / 4096) + 1 doubles.
extern "C" int oak_column_mean_f64(const double* d_X, int64_t n, int64_t ldx, int64_t col, double* d_out, void* d_work,
                                   void* stream_) {
  OAK_REQUIRE(d_X && d_out && d_work, "oak_column_mean_f64: null argument");
  OAK_REQUIRE(n >= 1 && ldx >= 1 && col >= 0 && col < ldx, "oak_column_mean_f64: bad shape");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t seg = 4096;
  const int blocks = (int)((n + seg - 1) / seg);
  colsum_partial_kernel<<<blocks, 256, 0, stream>>>(d_X, n, ldx, col, seg, (double*)d_work);
  OAK_LAUNCHED();
  colsum_final_kernel<<<1, 256, 0, stream>>>((const double*)d_work, blocks, (double)n, d_out);
  OAK_LAUNCHED();
  return 0;
}
