// Phi += A A^T (and A y) on the FP64 tensor cores, second generation: 128 x 128 CTA tiles, weighted stream-K.
//
// The first kernel (oak_syrk.cu: 64 x 64 tiles, 4 warps, 2 CTAs / SM) keeps the DMMA pipe 88.8 % busy on the
// SGPR shape and moves 16 B / clk / SM out of L2 -- close to what L2 delivers (ncu: 6.0 TB/s).  This one
//   * uses 128 x 128 tiles (8 warps, warp tile 32 x 64, 64 accumulators per lane): half the L2 traffic and
//     5.3 instead of 4 DMMA per 16-byte shared-memory load;
//   * cuts diagonal tiles at 32 x 32 granularity and hands the ten lower sub-blocks to the eight warps so that
//     the slowest warp does 1.25 sub-blocks (an off-diagonal tile costs 2 per warp): a diagonal tile costs
//     5/8 of a full one, and inside the diagonal sub-blocks the 8 x 8 blocks above the diagonal are skipped;
//   * cuts the k axis into S slices and hands the (slice, tile) units out through an atomic counter in
//     slice-major order, off-diagonal tiles first: the ~4 slices in flight are each worked on by all their
//     tiles at the same time and at the same k position, so every staged row block is shared through L2 by the
//     nine tiles that need it (a tile-major stream-K split, tried first, put the CTAs at unrelated k positions:
//     L2 hit rate 43 %, 7.7x the chunk read from HBM, DMMA pipe 77 %); S is sized for ~17 units per CTA, which
//     bounds the tail to a few per cent whatever the tile count; each unit's piece is written once and a second
//     kernel adds the S pieces of every tile in slice order -- deterministic, no floating-point atomics;
//   * folds Kuf y into the pass (north_star: "warp-shuffle reductions for ... the SGPR Kuf Kfu accumulation"
//     -- here the tiles of block column 0 dot their staged rows with the staged y stage), which removes the
//     separate HBM-bound GEMV that re-read all of Kuf (cuBLAS dot_kernel, 2.2 % of the statistics phase).
// Operand layout and the k permutation inside a 16-wide stage are those of oak_syrk.cu (row stride 18
// doubles, lane q takes k = 4q .. 4q+3: conflict-free 16-byte fragment loads for both operands).
#include "oak_common.cuh"

namespace oak {

namespace syrk2 {
constexpr int kTile = 128;
constexpr int kKT = 16;
#ifndef OAK_SYRK2_STAGES
#define OAK_SYRK2_STAGES 4
#endif
constexpr int kStages = OAK_SYRK2_STAGES;
constexpr int kThreads = 256;
constexpr int kRS = kKT + 2;                              // staged row stride (doubles)
constexpr int kStageDoubles = 2 * kTile * kRS + kKT;      // rows of block bi | rows of block bj | y stage
constexpr int kPieceDoubles = kTile * kTile + kTile;      // partial tile + partial A y of its rows
constexpr size_t kSmemBytes = ((size_t)kStages * kStageDoubles + 2 * kTile + 2) * sizeof(double);
}  // namespace syrk2

struct Syrk2Params {
  const double* A;       // row-major [m][lda], k contiguous
  const double* A_alt;   // contracted instead of A when *route != 0, or null
  const int* route;
  const double* y;       // [k] or null: pieces of block column 0 also carry A y of their rows
  double* pieces;        // [tiles][slices][kPieceDoubles]
  int64_t lda, k;        // k: valid columns (the y stage is zero-filled beyond)
  int m, nb, tiles, k_steps, slices;
  int* counter;          // unit scheduler, zeroed by the launcher
};

// unit -> (slice, tile): slice-major; inside a slice the off-diagonal tiles (cost 1) come before the diagonal
// ones (cost 5/8)
__device__ __forceinline__ void syrk2_unit(int unit, int tiles, int nb, int& slice, int& bi, int& bj) {
  slice = unit / tiles;
  const int idx = unit - slice * tiles;
  const int n_off = tiles - nb;
  if (idx < n_off) {
    int b = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
    while (b * (b + 1) / 2 > idx) --b;
    while ((b + 1) * (b + 2) / 2 <= idx) ++b;
    bi = b + 1;  // strictly lower triangle: row b + 1, column idx - b (b + 1) / 2
    bj = idx - b * (b + 1) / 2;
  } else {
    bi = bj = idx - n_off;
  }
}

__device__ __forceinline__ void s2_cp16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void s2_cp8(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void s2_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// one 32 x 32 sub-block per call: rows of sub-block rb against rows of sub-block cb of the staged block,
// accumulators acc[i][joff + j]; kDiagSub skips the 8 x 8 blocks above the diagonal
template <bool kDiagSub>
__device__ __forceinline__ void s2_subblock(const double* st, int rb, int cb, int g, int q, double (&acc)[4][8][2],
                                            const int joff) {
  using namespace syrk2;
  const double* Ar = st + (rb * 32 + g) * kRS + q * 4;
  const double* Br = st + (cb * 32 + g) * kRS + q * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double a[4][2], b[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 x = *reinterpret_cast<const double2*>(Ar + i * 8 * kRS + 2 * h);
      a[i][0] = x.x;
      a[i][1] = x.y;
      const double2 z = *reinterpret_cast<const double2*>(Br + i * 8 * kRS + 2 * h);
      b[i][0] = z.x;
      b[i][1] = z.y;
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (!kDiagSub || j <= i) s2_dmma(acc[i][joff + j], a[i][kk], b[j][kk]);
  }
}

__global__ void __launch_bounds__(syrk2::kThreads, 1) syrk2_lower_dmma_kernel(const Syrk2Params prm) {
  using namespace syrk2;
  extern __shared__ __align__(16) double smem[];
  double* const ysum = smem + kStages * kStageDoubles;  // [2][kTile]
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, q = lane & 3;
  const double* const Aop = (prm.route != nullptr && *prm.route != 0) ? prm.A_alt : prm.A;
  int* const s_unit = reinterpret_cast<int*>(ysum + 2 * kTile);
  const int units = prm.tiles * prm.slices;
  for (;;) {
    __syncthreads();  // the previous unit is done with the stage buffers, ysum and s_unit
    if (tid == 0) *s_unit = atomicAdd(prm.counter, 1);
    __syncthreads();
    const int unit = *s_unit;
    if (unit >= units) break;
    int slice, bi, bj;
    syrk2_unit(unit, prm.tiles, prm.nb, slice, bi, bj);
    const int s_lo = (int)((int64_t)slice * prm.k_steps / prm.slices);
    const int s_hi = (int)((int64_t)(slice + 1) * prm.k_steps / prm.slices);
    const int steps = s_hi - s_lo;
    const int t = bi * (bi + 1) / 2 + bj;
    const bool diag = (bi == bj);
    const bool ydot = (prm.y != nullptr) && (bj == 0);

    auto load_stage = [&](int step, int buf) {
      double* dst = smem + buf * kStageDoubles;
      const int64_t kbase = (int64_t)(s_lo + step) * kKT;
      const int nrows = diag ? kTile : 2 * kTile;
      for (int idx = tid; idx < nrows * (kKT / 2); idx += kThreads) {
        const int row = idx >> 3, ch = idx & 7;
        const int grow = (row < kTile) ? bi * kTile + row : bj * kTile + (row - kTile);
        const bool ok = grow < prm.m;
        s2_cp16(dst + row * kRS + ch * 2, Aop + (int64_t)(ok ? grow : 0) * prm.lda + kbase + ch * 2, ok);
      }
      if (ydot && tid < kKT) {
        const bool ok = kbase + tid < prm.k;
        s2_cp8(dst + 2 * kTile * kRS + tid, prm.y + (ok ? kbase + tid : 0), ok);
      }
      asm volatile("cp.async.commit_group;\n");
    };

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double yacc = 0.0;

#pragma unroll
    for (int p = 0; p < kStages - 1; ++p) {
      if (p < steps) load_stage(p, p);
      else asm volatile("cp.async.commit_group;\n");
    }
    for (int step = 0; step < steps; ++step) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(kStages - 2));
      __syncthreads();
      {
        const int nxt = step + kStages - 1;
        if (nxt < steps) load_stage(nxt, nxt % kStages);
        else asm volatile("cp.async.commit_group;\n");
      }
      const double* st = smem + (step % kStages) * kStageDoubles;
      if (!diag) {
        const double* Ar = st + (wm * 32 + g) * kRS + q * 4;
        const double* Br = st + (kTile + wn * 64 + g) * kRS + q * 4;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double a[4][2], b[8][2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double2 x = *reinterpret_cast<const double2*>(Ar + i * 8 * kRS + 2 * h);
            a[i][0] = x.x;
            a[i][1] = x.y;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double2 z = *reinterpret_cast<const double2*>(Br + j * 8 * kRS + 2 * h);
            b[j][0] = z.x;
            b[j][1] = z.y;
          }
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) s2_dmma(acc[i][j], a[i][kk], b[j][kk]);
        }
      } else {
        // diagonal tile: ten 32 x 32 sub-blocks (r >= c) over eight warps
        //   w0: (0,0) (1,1)   w1: (2,2) (3,3)   w2: (1,0)  w3: (2,0)  w4: (2,1)  w5: (3,0)  w6: (3,1)  w7: (3,2)
        if (warp < 2) {
          s2_subblock<true>(st, 2 * warp, 2 * warp, g, q, acc, 0);
          s2_subblock<true>(st, 2 * warp + 1, 2 * warp + 1, g, q, acc, 4);
        } else {
          const int rb = (warp == 2) ? 1 : (warp <= 4 ? 2 : 3);
          const int cb = (warp == 2 || warp == 3 || warp == 5) ? 0 : (warp == 7 ? 2 : 1);
          s2_subblock<false>(st, rb, cb, g, q, acc, 0);
        }
      }
      if (ydot) {
        // thread -> (row = tid & 127, k half = tid >> 7): 8 products per stage; the lane's k order is irrelevant
        const double* rowp = st + (tid & (kTile - 1)) * kRS + (tid >> 7) * 8;
        const double* yp = st + 2 * kTile * kRS + (tid >> 7) * 8;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double2 av = *reinterpret_cast<const double2*>(rowp + 2 * e);
          const double2 yv = *reinterpret_cast<const double2*>(yp + 2 * e);
          yacc = fma(av.x, yv.x, yacc);
          yacc = fma(av.y, yv.y, yacc);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;\n");

    // ---- piece -> workspace (row-major 128 x 128, then the 128 partial A y values) ------------------
    double* Pc = prm.pieces + ((int64_t)t * prm.slices + slice) * kPieceDoubles;
    if (!diag) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = wm * 32 + i * 8 + g, cc = wn * 64 + j * 8 + 2 * q;
          *reinterpret_cast<double2*>(Pc + r * kTile + cc) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    } else if (warp < 2) {
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            const int r = (2 * warp + s) * 32 + i * 8 + g, cc = (2 * warp + s) * 32 + j * 8 + 2 * q;
            *reinterpret_cast<double2*>(Pc + r * kTile + cc) = make_double2(acc[i][4 * s + j][0], acc[i][4 * s + j][1]);
          }
    } else {
      const int rb = (warp == 2) ? 1 : (warp <= 4 ? 2 : 3);
      const int cb = (warp == 2 || warp == 3 || warp == 5) ? 0 : (warp == 7 ? 2 : 1);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = rb * 32 + i * 8 + g, cc = cb * 32 + j * 8 + 2 * q;
          *reinterpret_cast<double2*>(Pc + r * kTile + cc) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
    if (ydot) {
      ysum[(tid >> 7) * kTile + (tid & (kTile - 1))] = yacc;
      __syncthreads();
      if (tid < kTile) Pc[kTile * kTile + tid] = ysum[tid] + ysum[kTile + tid];
    }
  }
}

// C (lower, column-major m x m) += sum of the S pieces of every tile, in slice order; kufy += the A y pieces of
// block column 0.  grid = (tiles, 4): block (t, s) owns rows [32 s, 32 s + 32) of tile t.
__global__ void __launch_bounds__(256) syrk2_reduce_kernel(const Syrk2Params prm, double* __restrict__ C,
                                                           double* __restrict__ kufy) {
  using namespace syrk2;
  __shared__ double sh[32][kTile + 1];
  const int t = blockIdx.x, strip = blockIdx.y;
  int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (bi * (bi + 1) / 2 > t) --bi;
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const int bj = t - bi * (bi + 1) / 2;
  const bool diag = (bi == bj);
  double v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = 0.0;
  double yv = 0.0;
  for (int sl = 0; sl < prm.slices; ++sl) {
    const double* Pc = prm.pieces + ((int64_t)t * prm.slices + sl) * kPieceDoubles;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int idx = threadIdx.x + 256 * e;  // 32 rows x 128 columns, column fastest
      const int r = 32 * strip + (idx >> 7), cc = idx & (kTile - 1);
      // diagonal tiles: only the 8 x 8 blocks on / below the diagonal were written
      if (!diag || cc <= r) v[e] += Pc[r * kTile + cc];
    }
    if (kufy != nullptr && bj == 0 && strip == 0 && threadIdx.x < kTile) yv += Pc[kTile * kTile + threadIdx.x];
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int idx = threadIdx.x + 256 * e;
    sh[idx >> 7][idx & (kTile - 1)] = v[e];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * kTile; idx += 256) {
    const int r = idx & 31, cc = idx >> 5;  // row fastest: contiguous in column-major C
    const int i = bi * kTile + 32 * strip + r, j = bj * kTile + cc;
    if (i < prm.m && j < prm.m && i >= j) C[(int64_t)j * prm.m + i] += sh[r][cc];
  }
  if (kufy != nullptr && bj == 0 && strip == 0 && threadIdx.x < kTile) {
    const int i = bi * kTile + threadIdx.x;
    if (i < prm.m) kufy[i] += yv;
  }
}

// ---- host side ----------------------------------------------------------------------------------------
static int syrk2_sms(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int sms = 0, per_sm = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  if (cudaFuncSetAttribute(syrk2_lower_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)syrk2::kSmemBytes) != cudaSuccess)
    return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, syrk2_lower_dmma_kernel, syrk2::kThreads,
                                                    syrk2::kSmemBytes) != cudaSuccess || per_sm < 1)
    return 0;
  if (device >= 0 && device < 64) cached[device] = sms;
  return sms;
}

constexpr int kSyrk2MaxSlices = 96;

// slices: ~17 units per CTA (tail of a few per cent), at least 32 stages per unit
static int syrk2_pick_slices(int tiles, int k_steps, int G) {
  int s = (17 * G + tiles - 1) / tiles;
  if (s > kSyrk2MaxSlices) s = kSyrk2MaxSlices;
  if (s > k_steps / 32) s = k_steps / 32;
  return s < 1 ? 1 : s;
}

size_t syrk2_work_bytes(int m, int device) {
  const int G = syrk2_sms(device);
  if (G <= 0 || m <= 0) return 0;
  const int nb = (m + syrk2::kTile - 1) / syrk2::kTile;
  const int tiles = nb * (nb + 1) / 2;
  int s = (17 * G + tiles - 1) / tiles;
  if (s > kSyrk2MaxSlices) s = kSyrk2MaxSlices;
  return (size_t)tiles * s * syrk2::kPieceDoubles * sizeof(double) + 64;
}

// C (lower, column-major m x m) += A A^T and, with y, kufy += A y; A row-major [m][lda] with k valid columns
// (columns [k, roundup16(k)) are zeroed here).  `work` must hold syrk2_work_bytes(m, device).
int syrk2_lower_dmma(int m, int64_t k, double* A, int64_t lda, double* C, const double* y, double* kufy, double* work,
                     size_t work_bytes, int device, cudaStream_t stream, double* A_alt, const int* d_route) {
  using namespace syrk2;
  if (m <= 0 || k <= 0) return 0;
  const int G = syrk2_sms(device);
  OAK_REQUIRE(G > 0, "syrk2_lower_dmma: kernel does not fit on this device");
  OAK_REQUIRE(syrk2_work_bytes(m, device) <= work_bytes, "syrk2_lower_dmma: workspace too small");
  const int64_t k_pad = (k + kKT - 1) / kKT * kKT;
  OAK_REQUIRE(k_pad <= lda, "syrk2_lower_dmma: the padded k range exceeds the leading dimension");
  OAK_REQUIRE(lda % 2 == 0 && (reinterpret_cast<uintptr_t>(A) % 16 == 0), "syrk2_lower_dmma: unaligned operand");
  if (k_pad > k) {
    OAK_CUDA(cudaMemset2DAsync(A + k, (size_t)lda * sizeof(double), 0, (size_t)(k_pad - k) * sizeof(double), (size_t)m,
                               stream));
    if (A_alt)
      OAK_CUDA(cudaMemset2DAsync(A_alt + k, (size_t)lda * sizeof(double), 0, (size_t)(k_pad - k) * sizeof(double),
                                 (size_t)m, stream));
  }
  Syrk2Params prm;
  prm.A = A;
  prm.A_alt = A_alt;
  prm.route = A_alt ? d_route : nullptr;
  prm.y = y;
  prm.pieces = work;
  prm.lda = lda;
  prm.k = k;
  prm.m = m;
  prm.nb = (m + kTile - 1) / kTile;
  prm.tiles = prm.nb * (prm.nb + 1) / 2;
  prm.k_steps = (int)(k_pad / kKT);
  prm.slices = syrk2_pick_slices(prm.tiles, prm.k_steps, G);
  // the unit counter lives behind the pieces
  prm.counter = reinterpret_cast<int*>(reinterpret_cast<char*>(work) +
                                       (size_t)prm.tiles * prm.slices * kPieceDoubles * sizeof(double));
  OAK_CUDA(cudaMemsetAsync(prm.counter, 0, sizeof(int), stream));
  const int units = prm.tiles * prm.slices;
  const int grid = units < G ? units : G;
  syrk2_lower_dmma_kernel<<<grid, kThreads, kSmemBytes, stream>>>(prm);
  OAK_LAUNCHED();
  syrk2_reduce_kernel<<<dim3(prm.tiles, 4), 256, 0, stream>>>(prm, C, (y != nullptr) ? kufy : nullptr);
  OAK_LAUNCHED();
  return 0;
}

}  // namespace oak
