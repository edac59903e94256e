// k-means on the device: the inducing-point initialisation that precedes the hot path
// (oak/model_utils.py:31-41, 376-391 and oak/utils.py:533-574 call scikit-learn's
// KMeans(n_clusters, random_state=0)).
//
// scikit-learn is a third-party dependency of the reference (unpinned "scikit-learn" in its setup.py; 1.9.0 in the
// build image); its published algorithm is restated here piece by piece so that the result follows the same
// trajectory: centring, k-means++ seeding with 2 + log(k) local trials (the random numbers themselves are drawn on
// the host from numpy's RandomState, exactly as sklearn/cluster/_kmeans.py:_kmeans_plusplus draws them), Lloyd
// iterations with first-index ties, centres = sums * (1 / count), the strict / tolerance stopping rules of
// _kmeans_single_lloyd.  The kernels below are the O(N k d) and O(N d) pieces; the few-element decisions stay in
// Python (kmeans.py).  Every reduction has a fixed order: the same input gives the same centres on every run and
// on every rank.
#include "oak_common.cuh"

namespace oak {

namespace km {
constexpr int kThreads = 256;
constexpr int kSeg = 4096;  // elements per block of the two-level reductions / scans

// deterministic block sum of one value per thread (256 threads)
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
  __syncthreads();
  return r;  // valid in thread 0
}
}  // namespace km

// ---- centring: column means, X - mean, mean of the column variances (the tolerance scale) -----------------------
// grid (blocks over rows, d): per (block, column) partial sums of X[:, c] (pass 0) or of (X[:, c] - mean_c)^2 while
// writing the centred copy (pass 1)
__global__ void __launch_bounds__(km::kThreads) kmeans_colsum_kernel(const double* __restrict__ X, int64_t n, int d,
                                                                     int64_t ldx, const double* __restrict__ mean,
                                                                     double* __restrict__ Xc,
                                                                     double* __restrict__ partial) {
  __shared__ double sh[8];
  const int c = blockIdx.y;
  const int64_t lo = (int64_t)blockIdx.x * km::kSeg;
  const int64_t hi = lo + km::kSeg < n ? lo + km::kSeg : n;
  const double mu = mean ? mean[c] : 0.0;
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += km::kThreads) {
    const double v = X[i * ldx + c] - mu;
    if (mean) {
      Xc[i * d + c] = v;
      acc = fma(v, v, acc);
    } else {
      acc += v;
    }
  }
  const double s = km::block_sum(acc, sh);
  if (threadIdx.x == 0) partial[(int64_t)c * gridDim.x + blockIdx.x] = s;
}

// out[c] = scale * sum_b partial[c][b] (fixed order); with `total` the d results are also averaged into total[0]
__global__ void kmeans_fold_kernel(const double* __restrict__ partial, int blocks, int d, double scale,
                                   double* __restrict__ out, double* __restrict__ total) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d) {
    double acc = 0.0;
    for (int b = 0; b < blocks; ++b) acc += partial[(int64_t)c * blocks + b];
    out[c] = acc * scale;
  }
  if (total) {
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      double t = 0.0;
      for (int q = 0; q < d; ++q) t += out[q];
      total[0] = t / d;
    }
  }
}

// ---- k-means++ ---------------------------------------------------------------------------------------------------
// dist[t][i] = min(closest[i], |x_i - x_cand[t]|^2) (closest == nullptr: no min), partial potentials per block.
// One thread per point; the candidates' coordinates sit in shared memory.
__global__ void __launch_bounds__(km::kThreads) kmeanspp_trial_kernel(const double* __restrict__ Xc, int64_t n, int d,
                                                                      const int64_t* __restrict__ cand, int trials,
                                                                      const double* __restrict__ closest,
                                                                      double* __restrict__ dist,
                                                                      double* __restrict__ partial) {
  extern __shared__ double sc[];  // [trials][d]
  __shared__ double sh[8];
  for (int e = threadIdx.x; e < trials * d; e += km::kThreads) sc[e] = Xc[cand[e / d] * d + e % d];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * km::kThreads + threadIdx.x;
  for (int t = 0; t < trials; ++t) {
    double v = 0.0;
    if (i < n) {
      const double* x = Xc + i * d;
      for (int q = 0; q < d; ++q) {
        const double df = x[q] - sc[t * d + q];
        v = fma(df, df, v);
      }
      if (closest) v = fmin(v, closest[i]);
      dist[(int64_t)t * n + i] = v;
    }
    const double s = km::block_sum(v, sh);
    if (threadIdx.x == 0) partial[(int64_t)t * gridDim.x + blockIdx.x] = s;
  }
}

// inclusive prefix sum in three fixed-order passes: per-block sequential-by-thread sums, scan of the block sums,
// per-block rescan with the offset
__global__ void __launch_bounds__(km::kThreads) scan_block_sums_kernel(const double* __restrict__ a, int64_t n,
                                                                       double* __restrict__ bsum) {
  // the same order as scan_apply_kernel (thread-contiguous runs, then the threads in order), so that the block
  // offsets and the in-block scans agree to the last bit and the result is monotone for non-negative input
  __shared__ double tsum[km::kThreads];
  constexpr int per = km::kSeg / km::kThreads;
  const int64_t lo = (int64_t)blockIdx.x * km::kSeg + (int64_t)threadIdx.x * per;
  double run = 0.0;
#pragma unroll
  for (int q = 0; q < per; ++q) run += (lo + q < n) ? a[lo + q] : 0.0;
  tsum[threadIdx.x] = run;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int t = 0; t < km::kThreads; ++t) r += tsum[t];
    bsum[blockIdx.x] = r;
  }
}
__global__ void scan_offsets_kernel(double* bsum, int blocks) {  // exclusive scan in place, one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double run = 0.0;
    for (int b = 0; b < blocks; ++b) {
      const double v = bsum[b];
      bsum[b] = run;
      run += v;
    }
  }
}
__global__ void __launch_bounds__(km::kThreads) scan_apply_kernel(const double* __restrict__ a, int64_t n,
                                                                  const double* __restrict__ boff,
                                                                  double* __restrict__ out) {
  // thread t owns the contiguous run [lo + t*per, lo + (t+1)*per): sequential inside, thread offsets by a shared scan
  __shared__ double tsum[km::kThreads];
  constexpr int per = km::kSeg / km::kThreads;
  const int64_t lo = (int64_t)blockIdx.x * km::kSeg + (int64_t)threadIdx.x * per;
  double loc[per];
  double run = 0.0;
#pragma unroll
  for (int q = 0; q < per; ++q) {
    const int64_t i = lo + q;
    run += (i < n) ? a[i] : 0.0;
    loc[q] = run;
  }
  tsum[threadIdx.x] = run;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = boff[blockIdx.x];
    for (int t = 0; t < km::kThreads; ++t) {
      const double v = tsum[t];
      tsum[t] = r;
      r += v;
    }
  }
  __syncthreads();
  const double off = tsum[threadIdx.x];
#pragma unroll
  for (int q = 0; q < per; ++q) {
    const int64_t i = lo + q;
    if (i < n) out[i] = off + loc[q];
  }
}

// np.searchsorted(cum, v) (side = "left": first i with cum[i] >= v), clipped to n - 1
__global__ void searchsorted_kernel(const double* __restrict__ cum, int64_t n, const double* __restrict__ v, int t,
                                    int64_t* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= t) return;
  const double x = v[q];
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (cum[mid] < x) lo = mid + 1; else hi = mid;
  }
  out[q] = lo < n - 1 ? lo : n - 1;
}

// ---- Lloyd: assignment -------------------------------------------------------------------------------------------
// labels[i] = argmin_k (|c_k|^2 - 2 x_i . c_k), first index on ties (sklearn/cluster/_k_means_lloyd.pyx); counts the
// labels that changed.  One point per thread with its coordinates in registers (DMAX >= d, zero padded), centres
// staged through shared memory in tiles of 64 (broadcast reads).
template <int DMAX>
__global__ void __launch_bounds__(128) kmeans_assign_kernel(const double* __restrict__ Xc, int64_t n, int d,
                                                            const double* __restrict__ centers, int k,
                                                            int32_t* __restrict__ labels,
                                                            unsigned long long* __restrict__ changed) {
  constexpr int kTile = 64;
  __shared__ double sc[kTile * DMAX];
  __shared__ double sn[kTile];
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  double x[DMAX];
#pragma unroll
  for (int q = 0; q < DMAX; ++q) x[q] = (i < n && q < d) ? Xc[i * d + q] : 0.0;
  double best = INFINITY;
  int arg = 0;
  for (int k0 = 0; k0 < k; k0 += kTile) {
    const int kt = min(kTile, k - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < kTile * DMAX; e += 128) {
      const int c = e / DMAX, q = e % DMAX;
      sc[e] = (c < kt && q < d) ? centers[(int64_t)(k0 + c) * d + q] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < kTile) {
      double nn = 0.0;
      for (int q = 0; q < DMAX; ++q) nn = fma(sc[threadIdx.x * DMAX + q], sc[threadIdx.x * DMAX + q], nn);
      sn[threadIdx.x] = nn;
    }
    __syncthreads();
    for (int c = 0; c < kt; ++c) {
      double dot = 0.0;
#pragma unroll
      for (int q = 0; q < DMAX; ++q) dot = fma(x[q], sc[c * DMAX + q], dot);
      const double s = fma(-2.0, dot, sn[c]);
      if (s < best) {
        best = s;
        arg = k0 + c;
      }
    }
  }
  bool diff = false;
  if (i < n) {
    diff = labels[i] != arg;
    labels[i] = arg;
  }
  const unsigned m = __ballot_sync(0xffffffffu, diff);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(changed, (unsigned long long)__popc(m));
}

// the same for inputs wider than 64 columns (rare: the continuous block of a data set): coordinates and centres are
// read through the caches instead of registers / shared memory -- correct for any d, not tuned
__global__ void __launch_bounds__(128) kmeans_assign_wide_kernel(const double* __restrict__ Xc, int64_t n, int d,
                                                                 const double* __restrict__ centers, int k,
                                                                 const double* __restrict__ cnorm,
                                                                 int32_t* __restrict__ labels,
                                                                 unsigned long long* __restrict__ changed) {
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  double best = INFINITY;
  int arg = 0;
  if (i < n) {
    const double* x = Xc + i * d;
    for (int c = 0; c < k; ++c) {
      const double* cc = centers + (int64_t)c * d;
      double dot = 0.0;
      for (int q = 0; q < d; ++q) dot = fma(x[q], __ldg(cc + q), dot);
      const double s = fma(-2.0, dot, cnorm[c]);
      if (s < best) {
        best = s;
        arg = c;
      }
    }
  }
  bool diff = false;
  if (i < n) {
    diff = labels[i] != arg;
    labels[i] = arg;
  }
  const unsigned m = __ballot_sync(0xffffffffu, diff);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(changed, (unsigned long long)__popc(m));
}
__global__ void kmeans_cnorm_kernel(const double* __restrict__ centers, int k, int d, double* __restrict__ cnorm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  double nn = 0.0;
  for (int q = 0; q < d; ++q) nn = fma(centers[(int64_t)c * d + q], centers[(int64_t)c * d + q], nn);
  cnorm[c] = nn;
}

// ---- Lloyd: update -------------------------------------------------------------------------------------------------
// partial[s][c][:] = sum of the points of segment s labelled c, cnt[s][c] their number.  CTA (segment s, cluster group
// g of kb clusters): each warp walks its own contiguous share of the segment in order, 32 labels at a time, and adds
// the matching rows into ITS shared-memory accumulators (lanes over coordinates), so the order of every sum is fixed;
// the warps' accumulators are then added in warp order.
__global__ void __launch_bounds__(km::kThreads) kmeans_update_kernel(const double* __restrict__ Xc, int64_t n, int d,
                                                                     const int32_t* __restrict__ labels, int k, int kb,
                                                                     int64_t seg_len, double* __restrict__ partial,
                                                                     double* __restrict__ cnt) {
  extern __shared__ double acc[];  // [8 warps][kb][d + 1]  (last slot: count)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dp = d + 1;
  double* mine = acc + (int64_t)warp * kb * dp;
  for (int e = lane; e < kb * dp; e += 32) mine[e] = 0.0;
  __syncwarp();
  const int k0 = blockIdx.y * kb;
  const int64_t s_lo = (int64_t)blockIdx.x * seg_len;
  const int64_t s_hi = s_lo + seg_len < n ? s_lo + seg_len : n;
  const int64_t share = (s_hi - s_lo + 7) / 8;
  const int64_t w_lo = s_lo + warp * share;
  const int64_t w_hi = w_lo + share < s_hi ? w_lo + share : s_hi;
  for (int64_t base = w_lo; base < w_hi; base += 32) {
    const int64_t i = base + lane;
    const int lab = (i < w_hi) ? labels[i] - k0 : -1;
    unsigned m = __ballot_sync(0xffffffffu, lab >= 0 && lab < kb);
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const int l = __shfl_sync(0xffffffffu, lab, j);
      const double* x = Xc + (base + j) * d;
      for (int q = lane; q < d; q += 32) mine[l * dp + q] += x[q];
      if (lane == 0) mine[l * dp + d] += 1.0;
    }
    __syncwarp();
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kb * dp; e += km::kThreads) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += acc[(int64_t)w * kb * dp + e];
    const int c = k0 + e / dp, q = e % dp;
    if (c < k) {
      if (q < d)
        partial[((int64_t)blockIdx.x * k + c) * d + q] = v;
      else
        cnt[(int64_t)blockIdx.x * k + c] = v;
    }
  }
}

// sums[c][:] = sum_s partial[s][c][:], counts[c] = sum_s cnt[s][c]; then, as sklearn's _average_centers and
// _center_shift: centres_new = sums * (1 / count) for non-empty clusters (empty ones are left to the host, which is
// told how many there are), shift_tot = sum_c (|new_c - old_c|)^2.
__global__ void __launch_bounds__(km::kThreads) kmeans_finish_kernel(const double* __restrict__ partial,
                                                                     const double* __restrict__ cnt, int segs, int k,
                                                                     int d, const double* __restrict__ centers_old,
                                                                     double* __restrict__ sums,
                                                                     double* __restrict__ counts,
                                                                     double* __restrict__ centers_new,
                                                                     double* __restrict__ shift,
                                                                     unsigned long long* __restrict__ n_empty) {
  const int c = blockIdx.x * km::kThreads + threadIdx.x;
  if (c >= k) return;
  double w = 0.0;
  for (int s = 0; s < segs; ++s) w += cnt[(int64_t)s * k + c];
  counts[c] = w;
  if (w == 0.0) atomicAdd(n_empty, 1ULL);
  const double alpha = w > 0.0 ? 1.0 / w : 0.0;
  double sh2 = 0.0;
  for (int q = 0; q < d; ++q) {
    double v = 0.0;
    for (int s = 0; s < segs; ++s) v += partial[((int64_t)s * k + c) * d + q];
    sums[(int64_t)c * d + q] = v;
    const double nc = w > 0.0 ? v * alpha : centers_old[(int64_t)c * d + q];
    centers_new[(int64_t)c * d + q] = nc;
    const double df = nc - centers_old[(int64_t)c * d + q];
    sh2 = fma(df, df, sh2);
  }
  const double sft = sqrt(sh2);
  shift[c] = sft * sft;  // (center_shift ** 2): the square of the rounded norm, as the reference library forms it
}

}  // namespace oak

using namespace oak;

extern "C" size_t oak_kmeans_work_bytes(int64_t n, int64_t d, int64_t k, int64_t trials) {
  if (n < 0 || d < 1 || k < 1 || trials < 1) return 0;
  const size_t blocks = (size_t)((n + km::kSeg - 1) / km::kSeg) + 1;
  const size_t blocks256 = (size_t)((n + km::kThreads - 1) / km::kThreads) + 1;
  const size_t a = (size_t)d * blocks + (size_t)d, b = (size_t)trials * blocks256,
               c = 64 * (size_t)k * (size_t)(d + 1) + (size_t)k;
  size_t m = a > b ? a : b;
  if (c > m) m = c;
  return (m + 64) * sizeof(double);
}

// d_stats[0..d) = column means, d_stats[d] = mean of the column variances; d_Xc (n x d, contiguous) = X - mean
extern "C" int oak_kmeans_center_f64(const double* d_X, int64_t n, int64_t d, int64_t ldx, double* d_Xc,
                                     double* d_stats, void* d_work, void* stream_) {
  OAK_REQUIRE(d_X && d_Xc && d_stats && d_work, "oak_kmeans_center_f64: null argument");
  OAK_REQUIRE(n >= 1 && d >= 1 && d <= 256 && ldx >= d, "oak_kmeans_center_f64: bad shape (at most 256 columns)");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int blocks = (int)((n + km::kSeg - 1) / km::kSeg);
  double* partial = (double*)d_work;
  dim3 grid((unsigned)blocks, (unsigned)d);
  kmeans_colsum_kernel<<<grid, km::kThreads, 0, stream>>>(d_X, n, (int)d, ldx, nullptr, nullptr, partial);
  OAK_LAUNCHED();
  kmeans_fold_kernel<<<(unsigned)((d + 255) / 256), 256, 0, stream>>>(partial, blocks, (int)d, 1.0 / (double)n, d_stats,
                                                                     nullptr);
  OAK_LAUNCHED();
  kmeans_colsum_kernel<<<grid, km::kThreads, 0, stream>>>(d_X, n, (int)d, ldx, d_stats, d_Xc, partial);
  OAK_LAUNCHED();
  kmeans_fold_kernel<<<1, 256, 0, stream>>>(partial, blocks, (int)d, 1.0 / (double)n, partial + (size_t)d * blocks,
                                            d_stats + d);
  OAK_LAUNCHED();
  return 0;
}

// One k-means++ round (sklearn/cluster/_kmeans.py:246-270): candidates = searchsorted(cumsum(closest), thresholds)
// (skipped when d_thresholds == NULL: d_cand is given, e.g. the first centre), d_dist[t][i] = min(closest_i,
// |x_i - x_cand_t|^2) (no min when d_closest == NULL), d_pots[t] = sum_i d_dist[t][i].
extern "C" int oak_kmeanspp_round_f64(const double* d_Xc, int64_t n, int64_t d, const double* d_closest,
                                      const double* d_thresholds, int64_t trials, int64_t* d_cand, double* d_cum,
                                      double* d_dist, double* d_pots, void* d_work, void* stream_) {
  OAK_REQUIRE(d_Xc && d_cand && d_dist && d_pots && d_work, "oak_kmeanspp_round_f64: null argument");
  OAK_REQUIRE(n >= 1 && d >= 1 && trials >= 1 && trials * d * sizeof(double) <= 40000,
              "oak_kmeanspp_round_f64: bad shape");
  cudaStream_t stream = (cudaStream_t)stream_;
  double* work = (double*)d_work;
  if (d_thresholds) {
    OAK_REQUIRE(d_closest && d_cum, "oak_kmeanspp_round_f64: the candidate search needs closest and cum");
    const int blocks = (int)((n + km::kSeg - 1) / km::kSeg);
    scan_block_sums_kernel<<<blocks, km::kThreads, 0, stream>>>(d_closest, n, work);
    OAK_LAUNCHED();
    scan_offsets_kernel<<<1, 32, 0, stream>>>(work, blocks);
    OAK_LAUNCHED();
    scan_apply_kernel<<<blocks, km::kThreads, 0, stream>>>(d_closest, n, work, d_cum);
    OAK_LAUNCHED();
    searchsorted_kernel<<<1, 64, 0, stream>>>(d_cum, n, d_thresholds, (int)trials, d_cand);
    OAK_LAUNCHED();
  }
  const int blocks = (int)((n + km::kThreads - 1) / km::kThreads);
  kmeanspp_trial_kernel<<<blocks, km::kThreads, (size_t)trials * d * sizeof(double), stream>>>(
      d_Xc, n, (int)d, d_cand, (int)trials, d_closest, d_dist, work);
  OAK_LAUNCHED();
  kmeans_fold_kernel<<<(unsigned)((trials + 255) / 256), 256, 0, stream>>>(work, blocks, (int)trials, 1.0, d_pots,
                                                                          nullptr);
  OAK_LAUNCHED();
  return 0;
}

template <int DMAX>
static void launch_assign(const double* Xc, int64_t n, int d, const double* centers, int k, int32_t* labels,
                          unsigned long long* changed, cudaStream_t stream) {
  kmeans_assign_kernel<DMAX><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(Xc, n, d, centers, k, labels, changed);
}

// One Lloyd iteration (sklearn/cluster/_k_means_lloyd.pyx lloyd_iter_chunked_dense + _k_means_common.pyx
// _average_centers / _center_shift).  d_labels: in = previous labels, out = new ones.
// d_out (4 x 8 bytes): [0] double: sum of squared centre shifts; [2] uint64: number of empty clusters (their rows of
// d_centers_new keep the old centre: the caller relocates them as sklearn does); [3] uint64: labels that changed.
extern "C" int oak_kmeans_lloyd_f64(const double* d_Xc, int64_t n, int64_t d, const double* d_centers, int64_t k,
                                    int32_t* d_labels, double* d_centers_new, double* d_sums, double* d_counts,
                                    double* d_out, int update_centers, void* d_work, void* stream_) {
  OAK_REQUIRE(d_Xc && d_centers && d_labels && d_out && d_work, "oak_kmeans_lloyd_f64: null argument");
  OAK_REQUIRE(n >= 1 && d >= 1 && d <= 1024 && k >= 1 && k <= (1 << 20), "oak_kmeans_lloyd_f64: bad shape (d <= 1024)");
  cudaStream_t stream = (cudaStream_t)stream_;
  OAK_CUDA(cudaMemsetAsync(d_out, 0, 4 * sizeof(double), stream));
  unsigned long long* changed = reinterpret_cast<unsigned long long*>(d_out + 3);
  if (d <= 8) launch_assign<8>(d_Xc, n, (int)d, d_centers, (int)k, d_labels, changed, stream);
  else if (d <= 16) launch_assign<16>(d_Xc, n, (int)d, d_centers, (int)k, d_labels, changed, stream);
  else if (d <= 32) launch_assign<32>(d_Xc, n, (int)d, d_centers, (int)k, d_labels, changed, stream);
  else if (d <= 64) launch_assign<64>(d_Xc, n, (int)d, d_centers, (int)k, d_labels, changed, stream);
  else {
    double* cnorm = (double*)d_work;  // free until the update kernel
    kmeans_cnorm_kernel<<<(unsigned)((k + 255) / 256), 256, 0, stream>>>(d_centers, (int)k, (int)d, cnorm);
    OAK_LAUNCHED();
    kmeans_assign_wide_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_Xc, n, (int)d, d_centers, (int)k, cnorm,
                                                                              d_labels, changed);
  }
  OAK_LAUNCHED();
  if (!update_centers) return 0;
  OAK_REQUIRE(d_centers_new && d_sums && d_counts, "oak_kmeans_lloyd_f64: null output");
  // clusters per CTA: 8 warps x kb x (d + 1) doubles of shared memory within 96 KB
  int kb = (int)(1536 / (d + 1));
  if (kb < 1) kb = 1;
  if (kb > 64) kb = 64;
  if (kb > k) kb = (int)k;
  const int groups = (int)((k + kb - 1) / kb);
  int segs = (2 * 148 + groups - 1) / groups;
  if (segs > 64) segs = 64;
  if ((int64_t)segs * 2048 > n) segs = (int)((n + 2047) / 2048);
  if (segs < 1) segs = 1;
  const int64_t seg_len = (n + segs - 1) / segs;
  double* partial = (double*)d_work;                 // [segs][k][d]
  double* cnt = partial + (size_t)segs * k * d;      // [segs][k]
  const size_t smem = (size_t)8 * kb * (d + 1) * sizeof(double);
  OAK_CUDA(cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  kmeans_update_kernel<<<dim3((unsigned)segs, (unsigned)groups), km::kThreads, smem, stream>>>(
      d_Xc, n, (int)d, d_labels, (int)k, kb, seg_len, partial, cnt);
  OAK_LAUNCHED();
  double* shift = cnt + (size_t)segs * k;            // [k]
  kmeans_finish_kernel<<<(unsigned)((k + km::kThreads - 1) / km::kThreads), km::kThreads, 0, stream>>>(
      partial, cnt, segs, (int)k, (int)d, d_centers, d_sums, d_counts, d_centers_new, shift,
      reinterpret_cast<unsigned long long*>(d_out + 2));
  OAK_LAUNCHED();
  // d_out[0] = sum of the squared shifts (fixed order)
  kmeans_fold_kernel<<<1, 32, 0, stream>>>(shift, (int)k, 1, 1.0, d_out, nullptr);
  OAK_LAUNCHED();
  return 0;
}

// ---- one-dimensional spherical Gaussian mixture (the MOG input measure) -----------------------------------------------
// oak/model_utils.py:753-770 fits sklearn's GaussianMixture(n_components=K, random_state=0,
// covariance_type="spherical") to one input column.  The E step and the sufficient statistics of the M step -- the O(N K)
// part of every EM iteration -- run here; the K-element M step and the stopping rule stay on the host (gmm.py).
// part[v][block], v = k (sum resp_k), K + k (sum resp_k x), 2K + k (sum resp_k x^2), 3K (sum log p(x)).
// labels != NULL: hard responsibilities from the k-means initialisation (mixture/_base.py:119-128).
namespace oak {
constexpr int kGmmMaxK = 16;

__global__ void __launch_bounds__(km::kThreads) gmm1d_estep_kernel(const double* __restrict__ x, int64_t n, int K,
                                                                   const double* __restrict__ par,  // [4][K]
                                                                   const int32_t* __restrict__ labels,
                                                                   double* __restrict__ part) {
  __shared__ double sh[8];
  __shared__ double sp[4 * kGmmMaxK];
  if (threadIdx.x < 4 * K) sp[threadIdx.x] = par[threadIdx.x];
  __syncthreads();
  const double* means = sp;               // mu_k
  const double* prec = sp + K;            // precisions_k = precisions_cholesky_k^2
  const double* logdet = sp + 2 * K;      // log precisions_cholesky_k
  const double* logw = sp + 3 * K;        // log weights_k
  double s0[kGmmMaxK], s1[kGmmMaxK], s2[kGmmMaxK], ll = 0.0;
#pragma unroll
  for (int k = 0; k < kGmmMaxK; ++k) s0[k] = s1[k] = s2[k] = 0.0;
  const int64_t lo = (int64_t)blockIdx.x * km::kSeg;
  const int64_t hi = lo + km::kSeg < n ? lo + km::kSeg : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += km::kThreads) {
    const double xi = x[i], x2 = xi * xi;
    double r[kGmmMaxK];
    if (labels) {
#pragma unroll
      for (int k = 0; k < kGmmMaxK; ++k) r[k] = (k < K && labels[i] == k) ? 1.0 : 0.0;
    } else {
      // _estimate_log_gaussian_prob (spherical) + log weights, then scipy's logsumexp
      double lp[kGmmMaxK], mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < kGmmMaxK; ++k) {
        if (k < K) {
          const double q = means[k] * means[k] * prec[k] - 2.0 * (xi * (means[k] * prec[k])) + x2 * prec[k];
          lp[k] = -0.5 * (1.8378770664093453 + q) + logdet[k] + logw[k];
          mx = fmax(mx, lp[k]);
        }
      }
      double se = 0.0;
#pragma unroll
      for (int k = 0; k < kGmmMaxK; ++k)
        if (k < K) se += exp(lp[k] - mx);
      const double lse = log(se) + mx;
      ll += lse;
#pragma unroll
      for (int k = 0; k < kGmmMaxK; ++k) r[k] = (k < K) ? exp(lp[k] - lse) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kGmmMaxK; ++k) {
      s0[k] += r[k];
      s1[k] = fma(r[k], xi, s1[k]);
      s2[k] = fma(r[k], x2, s2[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < kGmmMaxK; ++k) {
    if (k < K) {
      double v = km::block_sum(s0[k], sh);
      if (threadIdx.x == 0) part[(int64_t)k * gridDim.x + blockIdx.x] = v;
      v = km::block_sum(s1[k], sh);
      if (threadIdx.x == 0) part[(int64_t)(K + k) * gridDim.x + blockIdx.x] = v;
      v = km::block_sum(s2[k], sh);
      if (threadIdx.x == 0) part[(int64_t)(2 * K + k) * gridDim.x + blockIdx.x] = v;
    }
  }
  const double v = km::block_sum(ll, sh);
  if (threadIdx.x == 0) part[(int64_t)(3 * K) * gridDim.x + blockIdx.x] = v;
}
}  // namespace oak

extern "C" size_t oak_gmm1d_work_bytes(int64_t n, int64_t K) {
  if (n < 0 || K < 1) return 0;
  return (size_t)(3 * K + 1) * (size_t)((n + km::kSeg - 1) / km::kSeg + 1) * sizeof(double);
}

// d_par: [4][K] = means | precisions | log precisions_cholesky | log weights (ignored with d_labels);
// d_out[3K + 1] = sum resp_k | sum resp_k x | sum resp_k x^2 | sum_i log p(x_i)
extern "C" int oak_gmm1d_estep_f64(const double* d_x, int64_t n, int64_t K, const double* d_par,
                                   const int32_t* d_labels, double* d_out, void* d_work, void* stream_) {
  OAK_REQUIRE(d_x && d_par && d_out && d_work, "oak_gmm1d_estep_f64: null argument");
  OAK_REQUIRE(n >= 1 && K >= 1 && K <= kGmmMaxK, "oak_gmm1d_estep_f64: bad shape (at most 16 components)");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int blocks = (int)((n + km::kSeg - 1) / km::kSeg);
  gmm1d_estep_kernel<<<blocks, km::kThreads, 0, stream>>>(d_x, n, (int)K, d_par, d_labels, (double*)d_work);
  OAK_LAUNCHED();
  const int nv = (int)(3 * K + 1);
  kmeans_fold_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, stream>>>((const double*)d_work, blocks, nv, 1.0, d_out,
                                                                      nullptr);
  OAK_LAUNCHED();
  return 0;
}
