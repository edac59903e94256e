// Fused OAK Gram / cross-covariance tile kernel (sm_100a).
//
// Replaces OAKKernel.K (oak/oak_kernel.py:251-265): for every output entry the D per-dimension
// constrained kernels (ortho_rbf_kernel.py:157-172, ortho_binary_kernel.py:40-53,
// ortho_categorical_kernel.py:55-68) are evaluated and folded in registers into the power sums
// s_1..s_P (oak_kernel.py:236-239); the Newton-Girard recurrence (:241-248) and the
// variance-weighted sum (:256-260) run in the epilogue.  No per-dimension N x N2 intermediate is
// ever written: HBM traffic is 8 B per output entry plus the (tiny) prepared points.
//
// Structure
//   * persistent CTAs (one per SM), 256 threads as a 16x16 grid, each thread owns an RM x RN
//     register micro-tile (rows ty+16r, cols tx+16c) -> CTA tile (16 RM) x (16 RN).
//   * the prepared row/col point tiles (double2 per point and dim) are staged in shared memory
//     by cp.async, double buffered over (tile, dim-chunk) stages: one __syncthreads per stage.
//   * FP64 exp is table driven (oak_common.cuh): the 2^(j/256) table is replicated 16x in shared
//     memory so that the lookups of a half-warp never bank-conflict.
//   * symmetric mode (X2 = X, full row range): only tiles that intersect the lower triangle are
//     evaluated; each is also written transposed through a padded shared-memory tile so that the
//     mirrored stores are coalesced.
#include <cuda_pipeline.h>

#include <cstdlib>

#include "oak_common.cuh"

namespace oak {

constexpr int kDimChunk = 16;  // dims per pipeline stage

struct GramParams {
  double sigma2[OAK_MAX_DEPTH + 1];
  const double2* pts_row;
  const double2* pts_col;
  const double* dim_aux;  // [D] RBF: -ln s^2 ; discrete: bits(table offset)
  const double* tables;
  const double* exptab;
  double* K;
  int64_t n_row_pad, n_col_pad, ldk;
  int64_t row_begin, row_end, col_begin, n2;  // n2 = number of output columns
  int64_t tiles_n, num_tiles;
  int D, Dc;
  int symmetric;        // 0 general, 1 lower-triangle tiles + mirrored stores, 2 lower trapezoid only
  int64_t tile_row0;    // global row-block index of row_begin (modes 1, 2)
};

template <int TXD, int TYD, int RM, int RN>
struct SmemLayout {
  static constexpr int TM = TYD * RM, TN = TXD * RN;
  static constexpr int kTabDoubles = kExpTab * 16;
  static constexpr int kStageDouble2 = kDimChunk * (TM + TN);
  static constexpr size_t bytes(bool /*symmetric*/) {
    return sizeof(double) * kTabDoubles + 2 * sizeof(double2) * kStageDouble2 +
           2 * sizeof(double) * kDimChunk;
  }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n"); }

// power-sum / direct-recurrence accumulation of one per-dimension kernel value
template <int P, int ALGO>
__device__ __forceinline__ void accumulate(double (&a)[P], double k) {
  if constexpr (ALGO == OAK_ESP_DIRECT) {
#pragma unroll
    for (int n = P - 1; n >= 1; --n) a[n] = fma(k, a[n - 1], a[n]);  // a[n] holds e_{n+1}
    a[0] += k;
  } else {
    // a[p-1] holds s_p = sum_d k_d^p            (oak_kernel.py:236-239)
    a[0] += k;
    if constexpr (P == 2) a[1] = fma(k, k, a[1]);
    if constexpr (P >= 3) {
      const double k2 = k * k;
      a[1] += k2;
      a[2] = fma(k2, k, a[2]);
      if constexpr (P >= 4) a[3] = fma(k2, k2, a[3]);
      if constexpr (P >= 5) {
        const double k3 = k2 * k;
        a[4] = fma(k3, k2, a[4]);
        if constexpr (P >= 6) a[5] = fma(k3, k3, a[5]);
        if constexpr (P >= 7) {
          const double k4 = k2 * k2;
          a[6] = fma(k4, k3, a[6]);
          if constexpr (P >= 8) a[7] = fma(k4, k4, a[7]);
          if constexpr (P >= 9) {
            double pw = k4 * k4;  // k^8
#pragma unroll
            for (int p = 9; p <= P; ++p) {
              pw *= k;
              a[p - 1] += pw;
            }
          }
        }
      }
    }
  }
}

// Newton-Girard (oak_kernel.py:241-248) + variance-weighted sum (:256-260)
template <int P, int ALGO>
__device__ __forceinline__ double finish(const double (&a)[P], const double* __restrict__ sigma2) {
  double r = sigma2[0];
  if constexpr (ALGO == OAK_ESP_DIRECT) {
#pragma unroll
    for (int n = 1; n <= P; ++n) r = fma(sigma2[n], a[n - 1], r);
  } else {
    double e[P + 1];
    e[0] = 1.0;
#pragma unroll
    for (int n = 1; n <= P; ++n) {
      double s = e[n - 1] * a[0];
#pragma unroll
      for (int q = 2; q <= n; ++q) {
        if (q & 1)
          s = fma(e[n - q], a[q - 1], s);
        else
          s = fma(-e[n - q], a[q - 1], s);
      }
      e[n] = s * (1.0 / n);
      r = fma(sigma2[n], e[n], r);
    }
  }
  return r;
}

template <int P, int TXD, int TYD, int RM, int RN, int ALGO>
__global__ void __launch_bounds__(TXD * TYD, 1) gram_kernel(const GramParams prm) {
  using L = SmemLayout<TXD, TYD, RM, RN>;
  constexpr int TM = L::TM, TN = L::TN;
  constexpr int kThreads = TXD * TYD;
  static_assert(TXD % 16 == 0, "the exp-table replicas are indexed by lane % 16");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sTab = reinterpret_cast<double*>(smem_raw);
  double2* sStage = reinterpret_cast<double2*>(sTab + L::kTabDoubles);
  double* sAux = reinterpret_cast<double*>(sStage + 2 * L::kStageDouble2);

  const int tid = threadIdx.x;
  const int tx = tid % TXD, ty = tid / TXD;

  // replicate the exp table: entry j, replica r at sTab[j*16 + r]
  // (high words pre-compensated by -(j << 12), see exp_neg_tile)
  for (int i = tid; i < L::kTabDoubles; i += kThreads) {
    const int j = i >> 4;
    const double v = prm.exptab[j];
    sTab[i] = __hiloint2double(__double2hiint(v) - (j << 12), __double2loint(v));
  }
  const unsigned char* tab_bytes = smem_raw;
  const unsigned lane_bits = (unsigned)(tx & 15) * 8u;

  const int D = prm.D, Dc = prm.Dc;
  const int num_chunks = (D + kDimChunk - 1) / kDimChunk;

  auto tile_coords = [&](int64_t t, int64_t& bi, int64_t& bj) {
    if (prm.symmetric) {
      // tiles of the lower triangle, row-block major: t' = gbi (gbi+1)/2 + bj with the global
      // row-block index gbi = tile_row0 + bi (tile_row0 > 0 for a sharded lower trapezoid)
      const int64_t g0 = prm.tile_row0;
      t += g0 * (g0 + 1) / 2;
      int64_t b = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while (b * (b + 1) / 2 > t) --b;
      while ((b + 1) * (b + 2) / 2 <= t) ++b;
      bi = b - g0;
      bj = t - b * (b + 1) / 2;
    } else {
      bi = t / prm.tiles_n;
      bj = t - bi * prm.tiles_n;
    }
  };

  auto issue_stage = [&](int64_t t, int ch, int buf) {
    int64_t bi, bj;
    tile_coords(t, bi, bj);
    const int d0 = ch * kDimChunk;
    const int nd = min(kDimChunk, D - d0);
    double2* dst = sStage + buf * L::kStageDouble2;
    const int64_t row0 = prm.row_begin + bi * TM;  // global row index (rows are never padded
    const int64_t col0 = prm.col_begin + bj * TN;  // beyond n_pad because TM,TN divide 128)
    for (int i = tid; i < nd * (TM + TN); i += kThreads) {
      const int dl = i / (TM + TN);
      const int o = i - dl * (TM + TN);
      const double2* src = (o < TM)
                               ? prm.pts_row + (int64_t)(d0 + dl) * prm.n_row_pad + row0 + o
                               : prm.pts_col + (int64_t)(d0 + dl) * prm.n_col_pad + col0 + (o - TM);
      cp_async16(dst + dl * (TM + TN) + o, src);
    }
    if (tid < nd) cp_async8(sAux + buf * kDimChunk + tid, prm.dim_aux + d0 + tid);
    cp_async_commit();
  };

  double acc[RM][RN][P];

  int64_t t = blockIdx.x;
  int buf = 0;
  if (t < prm.num_tiles) issue_stage(t, 0, 0);

  for (; t < prm.num_tiles; t += gridDim.x) {
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < RN; ++c)
#pragma unroll
        for (int p = 0; p < P; ++p) acc[r][c][p] = 0.0;

    for (int ch = 0; ch < num_chunks; ++ch) {
      cp_async_wait_all();
      __syncthreads();  // stage landed; everyone is done with the other buffer
      // prefetch the next stage into the other buffer
      if (ch + 1 < num_chunks)
        issue_stage(t, ch + 1, buf ^ 1);
      else if (t + gridDim.x < prm.num_tiles)
        issue_stage(t + gridDim.x, 0, buf ^ 1);

      const double2* sRow = sStage + buf * L::kStageDouble2;
      const double* aux = sAux + buf * kDimChunk;
      const int d0 = ch * kDimChunk;
      const int nd = min(kDimChunk, D - d0);
      const int nc = max(0, min(nd, Dc - d0));  // continuous dims in this chunk

#pragma unroll 1
      for (int dl = 0; dl < nc; ++dl) {
        const double2* rowp = sRow + dl * (TM + TN);
        const double2* colp = rowp + TM;
        const double nls = aux[dl];
        double2 rv[RM], cv[RN];
#pragma unroll
        for (int r = 0; r < RM; ++r) rv[r] = rowp[ty * RM + r];
#pragma unroll
        for (int c = 0; c < RN; ++c) cv[c] = colp[tx + TXD * c];
#pragma unroll
        for (int r = 0; r < RM; ++r)
#pragma unroll
          for (int c = 0; c < RN; ++c) {
            const double d = rv[r].x - cv[c].x;
            const double z = fma(d, d, nls);            // (x-y)^2 / (2 l^2) - ln s^2
            const double e = exp_neg_tile(z, tab_bytes, lane_bits);  // s^2 exp(-(x-y)^2/(2 l^2))
            const double k = fma(-rv[r].y, cv[c].y, e); // - cov_X_s(x) cov_X_s(y) / var_s
            accumulate<P, ALGO>(acc[r][c], k);
          }
      }
#pragma unroll 1
      for (int dl = nc; dl < nd; ++dl) {  // discrete dims: table gather
        const double2* rowp = sRow + dl * (TM + TN);
        const double2* colp = rowp + TM;
        const double* tbl = prm.tables + (int)__double_as_longlong(aux[dl]);
        int ro[RM], co[RN];
#pragma unroll
        for (int r = 0; r < RM; ++r) ro[r] = __double2hiint(rowp[ty * RM + r].x);
#pragma unroll
        for (int c = 0; c < RN; ++c) co[c] = __double2loint(colp[tx + TXD * c].x);
#pragma unroll
        for (int r = 0; r < RM; ++r)
#pragma unroll
          for (int c = 0; c < RN; ++c) accumulate<P, ALGO>(acc[r][c], __ldg(tbl + ro[r] + co[c]));
      }
      buf ^= 1;
    }

    // ---- epilogue: Newton-Girard + variance-weighted sum, stores -------------------------
    int64_t bi, bj;
    tile_coords(t, bi, bj);
    const int64_t row0 = bi * TM;  // relative to row_begin
    const int64_t col0 = bj * TN;
    const int64_t nrows = prm.row_end - prm.row_begin;
    const bool mirror = prm.symmetric == 1 && (bi + prm.tile_row0 != bj);
    // A thread owns RM consecutive rows (ty*RM + r): in the mirrored (transposed) tile these
    // are RM consecutive columns, i.e. whole 32-byte sectors per thread, written straight from
    // registers with 16-byte streaming stores -- no shared-memory transpose, no extra barrier.
    const bool vec_ok = (RM % 2 == 0) && ((prm.ldk & 1) == 0) &&
                        ((reinterpret_cast<uintptr_t>(prm.K) & 15) == 0);
#pragma unroll
    for (int c = 0; c < RN; ++c) {
      const int64_t col = col0 + tx + TXD * c;
      double v[RM];
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        const int64_t row = row0 + ty * RM + r;
        v[r] = finish<P, ALGO>(acc[r][c], prm.sigma2);
        if (row < nrows && col < prm.n2) __stcs(prm.K + row * prm.ldk + col, v[r]);
      }
      if (mirror && col < nrows) {
        const int64_t ocol0 = row0 + ty * RM;  // columns of the mirrored row `col`
        double* dst = prm.K + col * prm.ldk + ocol0;
        if (vec_ok && ocol0 + RM <= prm.n2) {
#pragma unroll
          for (int r = 0; r + 1 < RM; r += 2)
            __stcs(reinterpret_cast<double2*>(dst + r), make_double2(v[r], v[r + 1]));
        } else {
#pragma unroll
          for (int r = 0; r < RM; ++r)
            if (ocol0 + r < prm.n2) __stcs(dst + r, v[r]);
        }
      }
    }
  }
  cp_async_wait_all();
}

// ---- launch plumbing -------------------------------------------------------------------
template <int P, int TXD, int TYD, int RM, int RN, int ALGO>
static int launch_gram(const GramParams& prm, int sms, cudaStream_t stream) {
  using L = SmemLayout<TXD, TYD, RM, RN>;
  constexpr int kThreads = TXD * TYD;
  const size_t smem = L::bytes(prm.symmetric == 1);
  auto kern = gram_kernel<P, TXD, TYD, RM, RN, ALGO>;
  OAK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)L::bytes(true)));
  int per_sm = 1;
  OAK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t resident = (int64_t)sms * per_sm;
  const int64_t grid = prm.num_tiles < resident ? prm.num_tiles : resident;
  kern<<<(unsigned)grid, kThreads, smem, stream>>>(prm);
  OAK_LAUNCHED();
  return 0;
}

template <int P, int TXD, int TYD, int RM, int RN>
static int launch_algo(const GramParams& prm, int algo, int sms, cudaStream_t stream) {
  if (algo == OAK_ESP_DIRECT)
    return launch_gram<P, TXD, TYD, RM, RN, OAK_ESP_DIRECT>(prm, sms, stream);
  return launch_gram<P, TXD, TYD, RM, RN, OAK_ESP_NEWTON_GIRARD>(prm, sms, stream);
}

// depth <= 4: 64x64 tiles; two geometries (selected by OAK_GRAM_VARIANT for experiments):
//   0: 256 threads (16x16), 4x4 micro-tile, ~226 registers, 8 warps / SM
//   1: 512 threads (32x16), 4x2 micro-tile, <= 128 registers, 16 warps / SM
static int gram_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OAK_GRAM_VARIANT");
    v = e ? atoi(e) : 0;  // measured: variant 0 is 2-3 % faster on B200 (profiles/)
  }
  return v;
}

template <int P>
static int launch_small_depth(const GramParams& prm, int algo, int sms, cudaStream_t stream) {
  if (gram_variant() == 0) return launch_algo<P, 16, 16, 4, 4>(prm, algo, sms, stream);
  return launch_algo<P, 32, 16, 4, 2>(prm, algo, sms, stream);
}

int tile_rows_for_depth(int depth) { return depth <= 4 ? 64 : 32; }

static int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}


int gram_launch(const oak_spec* spec, const double2* prow, int64_t n_row_pad, int64_t row_begin,
                int64_t row_end, const double2* pcol, int64_t n_col_pad, int64_t col_begin,
                int64_t col_end, int mode, double* K, int64_t ldk, cudaStream_t stream) {
  GramParams prm;
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) prm.sigma2[p] = spec->sigma2[p];
  prm.pts_row = prow;
  prm.pts_col = pcol;
  prm.n_row_pad = n_row_pad;
  prm.n_col_pad = n_col_pad;
  prm.dim_aux = spec->d_neg_log_s2;
  prm.tables = spec->d_tables;
  prm.exptab = spec->d_exptab;
  prm.K = K;
  prm.ldk = ldk;
  prm.row_begin = row_begin;
  prm.row_end = row_end;
  prm.col_begin = col_begin;
  prm.n2 = col_end - col_begin;
  prm.D = spec->D;
  prm.Dc = spec->Dc;
  const int depth = spec->depth;
  const int T = tile_rows_for_depth(depth);
  // tile origins must sit on tile boundaries so that tile loads stay inside the padded block
  OAK_REQUIRE(row_begin % T == 0 && col_begin % T == 0,
              "gram: row/column range must start on a multiple of the tile size (64)");
  const int64_t tiles_m = (row_end - row_begin + T - 1) / T;
  const int64_t tiles_n = (prm.n2 + T - 1) / T;
  prm.symmetric = mode;
  prm.tile_row0 = mode ? row_begin / T : 0;
  prm.tiles_n = tiles_n;
  prm.num_tiles = mode ? tiles_m * (prm.tile_row0 + 1) + tiles_m * (tiles_m - 1) / 2 : tiles_m * tiles_n;
  const int sms = sm_count(spec->device);
  const int algo = spec->algo;
  switch (depth) {
    case 0:
    case 1: return launch_small_depth<1>(prm, algo, sms, stream);
    case 2: return launch_small_depth<2>(prm, algo, sms, stream);
    case 3: return launch_small_depth<3>(prm, algo, sms, stream);
    case 4: return launch_small_depth<4>(prm, algo, sms, stream);
    case 5: return launch_algo<5, 16, 16, 2, 2>(prm, algo, sms, stream);
    case 6: return launch_algo<6, 16, 16, 2, 2>(prm, algo, sms, stream);
    case 7: return launch_algo<7, 16, 16, 2, 2>(prm, algo, sms, stream);
    case 8: return launch_algo<8, 16, 16, 2, 2>(prm, algo, sms, stream);
    default:
      if (depth <= 12) return launch_algo<12, 16, 16, 2, 2>(prm, algo, sms, stream);
      return launch_algo<16, 16, 16, 2, 2>(prm, algo, sms, stream);
  }
}

}  // namespace oak

using namespace oak;

extern "C" int oak_gram_f64(const oak_spec* spec, const void* d_points, int64_t n,
                            const void* d_points2, int64_t n2, int64_t row_begin,
                            int64_t row_end, double* d_K, int64_t ldk, void* stream_) {
  OAK_REQUIRE(spec && d_points, "oak_gram_f64: null argument");
  const bool same = (d_points2 == nullptr);
  if (same) n2 = n;
  OAK_REQUIRE(n >= 0 && n2 >= 0, "oak_gram_f64: negative size");
  OAK_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n,
              "oak_gram_f64: row range outside [0, n]");
  if (row_end == row_begin || n2 == 0) return 0;
  OAK_REQUIRE(d_K, "oak_gram_f64: null output");
  OAK_REQUIRE(ldk >= n2, "oak_gram_f64: ldk smaller than the number of columns");
  const bool symmetric = same && row_begin == 0 && row_end == n;
  return gram_launch(spec, (const double2*)d_points, padded(n), row_begin, row_end,
                     same ? (const double2*)d_points : (const double2*)d_points2, padded(n2), 0, n2,
                     symmetric ? 1 : 0, d_K, ldk, (cudaStream_t)stream_);
}

extern "C" int oak_gram_lower_f64(const oak_spec* spec, const void* d_points, int64_t n,
                                  int64_t row_begin, int64_t row_end, double* d_K, int64_t ldk,
                                  void* stream_) {
  OAK_REQUIRE(spec && d_points, "oak_gram_lower_f64: null argument");
  OAK_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n,
              "oak_gram_lower_f64: row range outside [0, n]");
  if (row_end == row_begin) return 0;
  OAK_REQUIRE(d_K, "oak_gram_lower_f64: null output");
  OAK_REQUIRE(ldk >= row_end, "oak_gram_lower_f64: ldk smaller than row_end");
  return gram_launch(spec, (const double2*)d_points, padded(n), row_begin, row_end,
                     (const double2*)d_points, padded(n), 0, row_end, 2, d_K, ldk,
                     (cudaStream_t)stream_);
}
