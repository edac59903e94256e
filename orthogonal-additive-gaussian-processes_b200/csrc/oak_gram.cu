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
//     register micro-tile (rows ty*RM + r, cols tx + 16c) -> CTA tile (16 RM) x (16 RN); the tile
//     index is decoded once per CTA and advanced incrementally.
//   * the prepared row/col point tiles (double2 per point and dim) are staged in shared memory
//     by cp.async, double buffered over (tile, dim-chunk) stages: one __syncthreads per stage.
//   * FP64 exp is table driven (oak_common.cuh): the 2^(j/1024) table is replicated 8x in shared
//     memory; a clamp-free body runs when the min/max keys of the prepared coordinates prove the
//     distances bounded (and s^2 == 1), a general clamped body otherwise -- one uniform decision
//     per stage.
//   * epilogue: all Newton-Girard recurrences first, then direct register stores (interior tiles
//     without bounds checks).
//   * symmetric mode (X2 = X, full row range): only tiles that intersect the lower triangle are
//     evaluated; the transposed tile is staged in 128-byte-swizzled shared memory and leaves through
//     2-D TMA tensor stores issued after the next stage barrier (no epilogue barrier).
//   * gram_matvec_kernel: the same tiles contracted with a vector on the fly (prediction mean).
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)
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
  int tables_len;       // doubles in the discrete-table blob (staged in shared memory when <= kTableStage)
  const double* exptab;
  double* K;
  int64_t n_row_pad, n_col_pad, ldk;
  int64_t row_begin, row_end, col_begin, n2;  // n2 = number of output columns
  int64_t tiles_n, num_tiles;
  int D, Dc, device;
  int symmetric;        // 0 general, 1 lower-triangle tiles + mirrored stores, 2 lower trapezoid only,
                        // 3 lower trapezoid of a row strip + its mirror image into a second buffer (Kt)
  double* Kt;           // mode 3: (n2 x rows) block, Kt[j * ldkt + (i - row_begin)] = K(i, j) of the off-diagonal tiles
  int64_t ldkt;
  const double* ydot_y;  // YDOT kernels (SGPR Kuf tiles): y of the columns of this call, y[j - col_begin]
  double* ydot_part;     // [column tile][rows] partial sums  sum_{j in tile} K(i, j) y_j  (general mode only)
  int64_t tile_row0;    // global TN-row-block index of row_begin (modes 1, 2)
  const unsigned long long* mm_row;  // [2 D] min / max keys of the prepared coordinates, or null
  const unsigned long long* mm_col;
  int use_tma;  // symmetric mode: the transposed tile leaves through shared memory + 2-D TMA stores
                // (needs an even leading dimension and a 16-byte aligned K)
  alignas(64) CUtensorMap tm_mir;  // K as (cols, rows) f64 tensor, box 16 x TN, 128-byte swizzle
};

constexpr int kFastWords = 8;  // clamp-free exp flags for the first 256 continuous dims
constexpr int kTableStage = 2048;  // doubles of shared memory for the discrete B tables (16 KB)

template <int TXD, int TYD, int RM, int RN>
struct SmemLayout {
  static constexpr int TM = TYD * RM, TN = TXD * RN;
  static constexpr int kTabDoubles = kExpTab * kExpRepl;
  static constexpr int kStageDouble2 = kDimChunk * (TM + TN);
  // Staging of the transposed tile (symmetric mode) for the TMA stores: TM/16 sub-tiles of
  // [TN rows][16 doubles] in the 128-byte swizzle the tensor map declares; two buffers,
  // 1024-byte aligned.
  static constexpr int kOutBytes = TM * TN * (int)sizeof(double);
  static constexpr size_t base_bytes = sizeof(double) * kTabDoubles + 2 * sizeof(double2) * kStageDouble2 +
                                       2 * sizeof(double) * kDimChunk + sizeof(unsigned) * kFastWords +
                                       sizeof(double) * kTableStage;
  static constexpr size_t bytes(bool tma) { return base_bytes + (tma ? 1024 + 2 * (size_t)kOutBytes : 0); }
};

// ---- bulk (TMA) shared -> global copies ---------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, unsigned smem_addr) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tm), "r"(c0),
               "r"(c1), "r"(smem_addr)
               : "memory");
}
__device__ __forceinline__ void sts64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void sts128(unsigned addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

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

// DISC: the spec has discrete dimensions (table gathers).  Continuous-only specs run an instantiation without that
// code: the depth-4 tile's schedule (and 2 % of its time) turned out to depend on what else the function contains.
template <int P, int TXD, int TYD, int RM, int RN, int ALGO, int MINB, bool YDOT = false, bool DISC = true>
__global__ void __launch_bounds__(TXD * TYD, MINB) gram_kernel(const __grid_constant__ GramParams prm) {
  using L = SmemLayout<TXD, TYD, RM, RN>;
  constexpr int TM = L::TM, TN = L::TN;
  constexpr int RS = TN / TM;  // row sub-tiles per TN-sized triangle block (symmetric modes)
  constexpr int kThreads = TXD * TYD;
  static_assert(TXD % kExpRepl == 0, "the exp-table replicas are indexed by lane % kExpRepl");
  static_assert(TN % TM == 0, "triangle blocks are TN x TN");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sTab = reinterpret_cast<double*>(smem_raw);
  double2* sStage = reinterpret_cast<double2*>(sTab + L::kTabDoubles);
  double* sAux = reinterpret_cast<double*>(sStage + 2 * L::kStageDouble2);
  unsigned* sFast = reinterpret_cast<unsigned*>(sAux + 2 * kDimChunk);
  double* sTables = reinterpret_cast<double*>(sFast + kFastWords);
  // output staging (TMA path only), 1024-byte aligned shared-window addresses
  const unsigned out_mir = ((unsigned)__cvta_generic_to_shared(smem_raw + L::base_bytes) + 1023u) & ~1023u;
  int mirror_buf = 0;

  const int tid = threadIdx.x;
  const int tx = tid % TXD, ty = tid / TXD;
  const bool issuer = (tid & 31) == 0 && (tid >> 5) < TM / 16;  // lanes that issue the TMA stores
  int pend = 0, pend_row0 = 0, pend_col0 = 0;                  // mirrored tile staged, not yet issued
  unsigned pend_stage = 0;

  // replicate the exp table: entry j, replica r at sTab[j*kExpRepl + r]
  // (high words pre-compensated by -(j << 12), see exp_tail)
  for (int i = tid; i < L::kTabDoubles; i += kThreads) {
    const int j = i / kExpRepl;
    const double v = prm.exptab[j];
    sTab[i] = __hiloint2double(__double2hiint(v) - (j << (20 - kExpBits)), __double2loint(v));
  }
  const unsigned char* tab_bytes = smem_raw;
  const unsigned lane_bits = (unsigned)(tx & (kExpRepl - 1)) * 8u;
  // discrete B tables: a few dozen doubles gathered 16 times per thread, dimension and tile -- from shared memory
  // when the blob fits (global / L1 gathers held the mixed-input configuration at 0.76 of its roofline)
  const bool tables_in_smem = DISC && prm.tables_len > 0 && prm.tables_len <= kTableStage;
  if (tables_in_smem)
    for (int i = tid; i < prm.tables_len; i += kThreads) sTables[i] = prm.tables[i];

  const int D = prm.D, Dc = prm.Dc;
  const int num_chunks = (D + kDimChunk - 1) / kDimChunk;

  // Which continuous dims may take the clamp-free exp: s^2 == 1 and every |a_i - b_j| of the two
  // point sets inside kFastSpan (from the min/max keys written by the prepare kernel).
  for (int i = tid; i < kFastWords; i += kThreads) sFast[i] = 0u;
  __syncthreads();
  if (prm.mm_row != nullptr) {
    for (int d = tid; d < min(Dc, kFastWords * 32); d += kThreads) {
      const double rmin = order_key_decode(prm.mm_row[d]), rmax = order_key_decode(prm.mm_row[D + d]);
      const double cmin = order_key_decode(prm.mm_col[d]), cmax = order_key_decode(prm.mm_col[D + d]);
      const double span = fmax(rmax - cmin, cmax - rmin);
      if (prm.dim_aux[d] == 0.0 && span <= kFastSpan) atomicOr(&sFast[d >> 5], 1u << (d & 31));
    }
  }
  // (made visible by the first __syncthreads of the stage loop)

  // Tile enumeration.  Symmetric modes: TN x TN blocks of the lower triangle, row-block major
  // (t' = gb (gb+1)/2 + bj with the global row-block index gb; tile_row0 > 0 for a sharded lower
  // trapezoid), each block cut into RS row sub-tiles of TM rows.  General mode: row-major tiles.
  // The closed form (one FP64 sqrt / one 64-bit division) runs once per CTA; a CTA then advances
  // by gridDim.x tiles with a few integer operations.
  int64_t tb = 0, tj = 0;  // block row (global) / block column, or tile row / tile column
  int th = 0;              // row sub-tile inside a triangle block
  int64_t step_q = 0, step_r = 0;
  {
    const int64_t u0 = blockIdx.x, G = gridDim.x;
    if (prm.symmetric) {
      int64_t t = u0 / RS;
      th = (int)(u0 - t * RS);
      const int64_t g0 = prm.tile_row0;
      t += g0 * (g0 + 1) / 2;
      int64_t b = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while (b * (b + 1) / 2 > t) --b;
      while ((b + 1) * (b + 2) / 2 <= t) ++b;
      tb = b;
      tj = t - b * (b + 1) / 2;
      step_q = G / RS;
      step_r = G - step_q * RS;
    } else {
      const int64_t tn = prm.tiles_n > 0 ? prm.tiles_n : 1;
      tb = u0 / tn;
      tj = u0 - tb * tn;
      step_q = G / tn;
      step_r = G - step_q * tn;
    }
  }
  auto tile_origin = [&](int64_t& row0, int64_t& col0, bool& diag) {
    if (prm.symmetric) {
      row0 = (tb - prm.tile_row0) * TN + th * TM;  // relative to row_begin
      col0 = tj * TN;
      diag = (tb == tj);
    } else {
      row0 = tb * TM;
      col0 = tj * TN;
      diag = false;
    }
  };
  auto tile_advance = [&]() {
    if (prm.symmetric) {
      th += (int)step_r;
      tj += step_q;
      if (th >= RS) {
        th -= RS;
        ++tj;
      }
      while (tj > tb) {
        tj -= tb + 1;
        ++tb;
      }
    } else {
      tj += step_r;
      tb += step_q;
      if (tj >= prm.tiles_n) {
        tj -= prm.tiles_n;
        ++tb;
      }
    }
  };

  // (row0, col0) of a tile are relative to (row_begin, col_begin)
  auto issue_stage = [&](int64_t row0, int64_t col0, int ch, int buf) {
    const int d0 = ch * kDimChunk;
    const int nd = min(kDimChunk, D - d0);
    double2* dst = sStage + buf * L::kStageDouble2;
    row0 += prm.row_begin;  // global row index (tiles never leave the padded block because
    col0 += prm.col_begin;  // TM, TN divide kPointPad)
    for (int o = tid; o < TM + TN; o += kThreads) {
      const double2* src = (o < TM) ? prm.pts_row + (int64_t)d0 * prm.n_row_pad + row0 + o
                                    : prm.pts_col + (int64_t)d0 * prm.n_col_pad + col0 + (o - TM);
      const int64_t stride = (o < TM) ? prm.n_row_pad : prm.n_col_pad;
      for (int dl = 0; dl < nd; ++dl) cp_async16(dst + dl * (TM + TN) + o, src + dl * stride);
    }
    if (tid < nd) cp_async8(sAux + buf * kDimChunk + tid, prm.dim_aux + d0 + tid);
    cp_async_commit();
  };

  double acc[RM][RN][P];

  int64_t u = blockIdx.x;
  int buf = 0;
  int64_t nxt_row0 = 0, nxt_col0 = 0;
  bool nxt_diag = false;
  if (u < prm.num_tiles) {
    tile_origin(nxt_row0, nxt_col0, nxt_diag);
    issue_stage(nxt_row0, nxt_col0, 0, 0);
  }

  for (; u < prm.num_tiles; u += gridDim.x) {
    const int64_t row0 = nxt_row0, col0 = nxt_col0;
    const bool diag = nxt_diag;
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < RN; ++c)
#pragma unroll
        for (int p = 0; p < P; ++p) acc[r][c][p] = 0.0;

    for (int ch = 0; ch < num_chunks; ++ch) {
      cp_async_wait_all();
      __syncthreads();  // stage landed; everyone is done with the other buffer
      if (pend) {       // ... and the staged transpose of the previous tile is complete
        if (issuer) {
          tma_store_2d(&prm.tm_mir, pend_row0 + 16 * (tid >> 5), pend_col0, pend_stage + (tid >> 5) * (TN * 128));
          bulk_commit();
        }
        pend = 0;
      }
      // prefetch the next stage into the other buffer
      if (ch + 1 < num_chunks) {
        issue_stage(row0, col0, ch + 1, buf ^ 1);
      } else if (u + gridDim.x < prm.num_tiles) {
        tile_advance();
        tile_origin(nxt_row0, nxt_col0, nxt_diag);
        issue_stage(nxt_row0, nxt_col0, 0, buf ^ 1);
      }

      const double2* sRow = sStage + buf * L::kStageDouble2;
      const double* aux = sAux + buf * kDimChunk;
      const int d0 = ch * kDimChunk;
      const int nd = min(kDimChunk, D - d0);
      const int nc = max(0, min(nd, Dc - d0));  // continuous dims in this chunk
      const unsigned fast_bits =
          (d0 < kFastWords * 32) ? __funnelshift_r(sFast[d0 >> 5], (d0 >> 5) + 1 < kFastWords ? sFast[(d0 >> 5) + 1] : 0u, d0 & 31)
                                 : 0u;

      // One decision per stage (16 dims): the clamp-free body when every continuous dim of the
      // stage qualifies, else the general body (valid for all dims).  Keeping the branch outside
      // the dim loops leaves them straight-line (uniform loop counter, no reconvergence points).
      const unsigned cmask = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
      if (nc > 0 && (fast_bits & cmask) == cmask) {
#pragma unroll 1
        for (int dl = 0; dl < nc; ++dl) {
          const double2* rowp = sRow + dl * (TM + TN);
          const double2* colp = rowp + TM;
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
              const double e = exp_neg_sq_fast(d, tab_bytes, lane_bits);  // exp(-(x-y)^2/(2 l^2))
              const double k = fma(-rv[r].y, cv[c].y, e);  // - cov_X_s(x) cov_X_s(y) / var_s
              accumulate<P, ALGO>(acc[r][c], k);
            }
        }
      } else {
#pragma unroll 1
        for (int dl = 0; dl < nc; ++dl) {
          const double2* rowp = sRow + dl * (TM + TN);
          const double2* colp = rowp + TM;
          const double ax = aux[dl];  // -ln(s^2) 256/ln2
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
              const double zs = fma(d, d, ax);
              const double e = exp_neg_scaled(zs, tab_bytes, lane_bits);  // s^2 exp(-(x-y)^2/(2 l^2))
              const double k = fma(-rv[r].y, cv[c].y, e);
              accumulate<P, ALGO>(acc[r][c], k);
            }
        }
      }
      if constexpr (DISC) {
      if (tables_in_smem) {
        // discrete dims, tables staged in shared memory: 32-bit shared-window byte offsets (row part and column part
        // + table base formed once per dimension), one IADD + one LDS.64 per entry.  The generic-pointer form below
        // costs 64-bit address arithmetic and a generic LD per entry: 148 instead of 85 instructions per dimension
        // for 32 FP64 ones, which is most of what held the mixed-input configuration back.
        const unsigned tables_saddr = (unsigned)__cvta_generic_to_shared(sTables);
#pragma unroll 1
        for (int dl = nc; dl < nd; ++dl) {
          const double2* rowp = sRow + dl * (TM + TN);
          const double2* colp = rowp + TM;
          const unsigned tb = tables_saddr + 8u * (unsigned)(int)__double_as_longlong(aux[dl]);
          unsigned ro[RM], co[RN];
#pragma unroll
          for (int r = 0; r < RM; ++r) ro[r] = 8u * (unsigned)__double2hiint(rowp[ty * RM + r].x);
#pragma unroll
          for (int c = 0; c < RN; ++c) co[c] = 8u * (unsigned)__double2loint(colp[tx + TXD * c].x) + tb;
#pragma unroll
          for (int r = 0; r < RM; ++r)
#pragma unroll
            for (int c = 0; c < RN; ++c) {
              double k;
              asm("ld.shared.f64 %0, [%1];" : "=d"(k) : "r"(ro[r] + co[c]));
              accumulate<P, ALGO>(acc[r][c], k);
            }
        }
      } else {
#pragma unroll 1
        for (int dl = nc; dl < nd; ++dl) {  // discrete dims: table gather from global memory (large table blobs)
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
      }
      }  // DISC
      buf ^= 1;
    }

    // ---- epilogue: Newton-Girard + variance-weighted sum, stores -------------------------
    // All RM x RN results first (independent recurrences -> ILP), kept in acc[r][c][0].
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < RN; ++c) acc[r][c][0] = finish<P, ALGO>(acc[r][c], prm.sigma2);

    const bool mirror = (prm.symmetric == 1 || prm.symmetric == 3) && !diag;
    const int64_t nrows = prm.row_end - prm.row_begin;
    const int64_t ldk = prm.ldk;
    const int64_t trow = row0 + ty * RM;  // first row / first column of this thread
    const int64_t tcol = col0 + tx;
    if (mirror && prm.use_tma) {
      // The transposed tile would cost every warp 16 partial lines per store instruction.
      // Instead: registers -> swizzled shared-memory staging (double buffered) -> TM/16 2-D TMA
      // stores of [TN rows][16 doubles], one per lane 0 of the first TM/16 warps.  The stores are
      // issued after the NEXT stage barrier (which orders the staging writes), so the epilogue
      // needs no barrier of its own; the tensor map clips partial tiles.
      static_assert(RM % 2 == 0, "the transposed tile is staged as 16-byte pairs of consecutive rows");
      static_assert(kThreads / 32 >= TM / 16, "one issuing warp per 16-column box");
      const unsigned stage = out_mir + (unsigned)mirror_buf * L::kOutBytes;
#pragma unroll
      for (int c = 0; c < RN; ++c)
#pragma unroll
        for (int r = 0; r + 1 < RM; r += 2) {
          const int i = ty * RM + r, j = tx + TXD * c;
          const unsigned off = (unsigned)((i >> 4) * (TN * 128) + j * 128 + ((((i & 15) >> 1) ^ (j & 7)) << 4));
          sts128(stage + off, acc[r][c][0], acc[r + 1][c][0]);
        }
      fence_async_smem();
      // the group issued at the top of this tile read the OTHER buffer: it must be drained before
      // the next stage barrier lets anybody write that buffer again
      if (issuer) bulk_wait_read<0>();
      pend = 1;
      pend_row0 = (int)row0;
      pend_col0 = (int)col0;
      pend_stage = stage;
      mirror_buf ^= 1;
    }
    if constexpr (YDOT) {
      // Kuf y folded into the tile epilogue (north_star: warp-shuffle reduction): every thread dots its RM x RN
      // entries with the y values of its columns, the TXD threads of a row group are folded by shuffles, and
      // lane tx == 0 writes the tile's row sums -- one slot per (column tile, row), summed in a fixed order later.
      static_assert(TXD == 16, "row groups are half warps");
      double yv[RN];
#pragma unroll
      for (int c = 0; c < RN; ++c) yv[c] = (tcol + TXD * c < prm.n2) ? __ldg(prm.ydot_y + tcol + TXD * c) : 0.0;
      double* const part = prm.ydot_part + (col0 / TN) * nrows;
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        double v = acc[r][0][0] * yv[0];
#pragma unroll
        for (int c = 1; c < RN; ++c) v = fma(acc[r][c][0], yv[c], v);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (tx == 0 && trow + r < nrows) part[trow + r] = v;
      }
    }
    double* const kp = prm.K + trow * ldk + tcol;
    const bool interior = row0 + TM <= nrows && col0 + TN <= prm.n2;
    // the tile itself: straight from registers (a warp store covers two 128-byte row segments)
    if (interior) {
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < RN; ++c) __stcs(kp + r * ldk + TXD * c, acc[r][c][0]);
    } else {
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < RN; ++c)
          if (trow + r < nrows && tcol + TXD * c < prm.n2) __stcs(kp + r * ldk + TXD * c, acc[r][c][0]);
    }
    if (mirror && !prm.use_tma) {
      // direct mirrored stores (odd leading dimension / unaligned output)
      // row = column index (< n2), column = row index relative to the strip (< nrows)
      const int64_t ldm = prm.symmetric == 3 ? prm.ldkt : ldk;
      double* const mp = (prm.symmetric == 3 ? prm.Kt : prm.K) + tcol * ldm + trow;
#pragma unroll
      for (int c = 0; c < RN; ++c) {
        if (tcol + TXD * c < prm.n2) {
#pragma unroll
          for (int r = 0; r < RM; ++r)
            if (trow + r < nrows) __stcs(mp + (int64_t)(TXD * c) * ldm + r, acc[r][c][0]);
        }
      }
    }
  }
  if (pend) {  // the last mirrored tile of this CTA
    __syncthreads();
    if (issuer) {
      tma_store_2d(&prm.tm_mir, pend_row0 + 16 * (tid >> 5), pend_col0, pend_stage + (tid >> 5) * (TN * 128));
      bulk_commit();
    }
  }
  cp_async_wait_all();
  if (issuer) bulk_wait_all();
}

// ---- launch plumbing -------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// K viewed as a (cols, rows) FP64 tensor with row pitch ldk; box = 16 columns x box_rows rows,
// 128-byte swizzle (the staging layout of the kernel's epilogue).  Returns 0 on success.
static int encode_output_map(CUtensorMap* tm, double* K, int64_t cols, int64_t rows, int64_t ldk, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return 1;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ldk * sizeof(double)};
  const cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, K, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

template <int P, int TXD, int TYD, int RM, int RN, int ALGO, int MINB, bool YDOT, bool DISC>
static int launch_gram_impl(GramParams prm, int sms, int device, cudaStream_t stream) {
  using L = SmemLayout<TXD, TYD, RM, RN>;
  constexpr int kThreads = TXD * TYD;
  constexpr int TM = L::TM, TN = L::TN;
  const int64_t rows = prm.row_end - prm.row_begin;
  if (prm.symmetric) {
    const int64_t blocks_m = (rows + TN - 1) / TN;
    prm.tile_row0 = prm.row_begin / TN;
    prm.tiles_n = 0;
    prm.num_tiles = (blocks_m * (prm.tile_row0 + 1) + blocks_m * (blocks_m - 1) / 2) * (TN / TM);
  } else {
    prm.tile_row0 = 0;
    prm.tiles_n = (prm.n2 + TN - 1) / TN;
    prm.num_tiles = ((rows + TM - 1) / TM) * prm.tiles_n;
  }
  // 2-D TMA stores need a 16-byte aligned output with an even leading dimension
  static const int no_tma = env_int("OAK_GRAM_NOTMA", 0);
  prm.use_tma = 0;
  if ((prm.symmetric == 1 || prm.symmetric == 3) && !no_tma && (TM % 16 == 0) && kThreads / 32 >= TM / 16 &&
      prm.n2 < (1ll << 31) && rows < (1ll << 31)) {
    // the mirror image: memory rows = columns of K, memory columns = rows of the strip
    double* const mbase = prm.symmetric == 3 ? prm.Kt : prm.K;
    const int64_t mld = prm.symmetric == 3 ? prm.ldkt : prm.ldk;
    if (mld % 2 == 0 && reinterpret_cast<uintptr_t>(mbase) % 16 == 0 &&
        encode_output_map(&prm.tm_mir, mbase, rows, prm.n2, mld, TN) == 0)
      prm.use_tma = 1;
  }
  const size_t smem = L::bytes(prm.use_tma != 0);
  auto kern = gram_kernel<P, TXD, TYD, RM, RN, ALGO, MINB, YDOT, DISC>;
  static int cached_per_sm[64][2] = {{0}};  // per instantiation, device and shared-memory footprint
  const int fp = prm.use_tma ? 1 : 0;
  int per_sm = (device >= 0 && device < 64) ? cached_per_sm[device][fp] : 0;
  if (per_sm == 0) {
    OAK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes(true)));
    OAK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    if (device >= 0 && device < 64) cached_per_sm[device][fp] = per_sm;
  }
  const int64_t resident = (int64_t)sms * per_sm;
  const int64_t grid = prm.num_tiles < resident ? prm.num_tiles : resident;
  kern<<<(unsigned)grid, kThreads, smem, stream>>>(prm);
  OAK_LAUNCHED();
  return 0;
}

template <int P, int TXD, int TYD, int RM, int RN, int ALGO, int MINB, bool YDOT = false>
static int launch_gram(const GramParams& prm, int sms, int device, cudaStream_t stream) {
  if (prm.D > prm.Dc) return launch_gram_impl<P, TXD, TYD, RM, RN, ALGO, MINB, YDOT, true>(prm, sms, device, stream);
  return launch_gram_impl<P, TXD, TYD, RM, RN, ALGO, MINB, YDOT, false>(prm, sms, device, stream);
}

template <int P, int TXD, int TYD, int RM, int RN, int MINB>
static int launch_algo(const GramParams& prm, int algo, int sms, cudaStream_t stream) {
  const int device = prm.device;
  if (algo == OAK_ESP_DIRECT)
    return launch_gram<P, TXD, TYD, RM, RN, OAK_ESP_DIRECT, MINB>(prm, sms, device, stream);
  return launch_gram<P, TXD, TYD, RM, RN, OAK_ESP_NEWTON_GIRARD, MINB>(prm, sms, device, stream);
}

// depth <= 4, 4x4 register micro-tile.  Geometries (OAK_GRAM_VARIANT, for A/B experiments):
//   0: 256 threads (16x16), 64x64 tile, 1 CTA / SM
//   1: 512 threads (32x16), 4x2 micro-tile, 64x64 tile, <= 128 registers
//   2: 128 threads (16x8), 32x64 tile, 2 CTAs / SM: the two CTAs run out of phase, so the
//      latency-bound epilogue (Newton-Girard + stores) of one overlaps the FP64 loop of the other
//   3: 256 threads (16x16), 2x4 micro-tile, 32x64 tile, 2 CTAs / SM
//   4: 256 threads (32x8), 4x2 micro-tile, 32x64 tile, 2 CTAs / SM
template <int P>
static int launch_small_depth(const GramParams& prm, int algo, int sms, cudaStream_t stream) {
  if (prm.ydot_y != nullptr) {  // the SGPR Kuf tiles with the folded Kuf y (general mode)
    if (algo == OAK_ESP_DIRECT)
      return launch_gram<P, 16, 16, 4, 4, OAK_ESP_DIRECT, 1, true>(prm, sms, prm.device, stream);
    return launch_gram<P, 16, 16, 4, 4, OAK_ESP_NEWTON_GIRARD, 1, true>(prm, sms, prm.device, stream);
  }
#ifdef OAK_GRAM_EXPERIMENTS  // the losing geometries of profiles/r01_ab_gram_variants_*.txt; not built by default
  static const int variant = env_int("OAK_GRAM_VARIANT", 0);
  if (variant == 1) return launch_algo<P, 32, 16, 4, 2, 1>(prm, algo, sms, stream);
  if (variant == 2) return launch_algo<P, 16, 8, 4, 4, 2>(prm, algo, sms, stream);
#endif
  return launch_algo<P, 16, 16, 4, 4, 1>(prm, algo, sms, stream);
}

int tile_rows_for_depth(int depth) { return depth <= 8 ? 64 : 32; }

static int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}

// ---- fused K(X*, Z) alpha (prediction mean) ------------------------------------------------------
// out[i] = sum_j K(x_i, z_j) alpha_j without the N* x M matrix (gpflow predict_f mean = Kus^T alpha,
// oak/model_utils.py:429-443).  One CTA owns a block of TM rows and walks all column
// tiles; every thread keeps the running sums of its RM rows, the 16 threads of a row group are
// folded by shuffles at the end: deterministic, no atomics.
template <int P, int ALGO>
__global__ void __launch_bounds__(256, 1) gram_matvec_kernel(const GramParams prm, const double* __restrict__ alpha,
                                                             double* __restrict__ out) {
  constexpr int TXD = 16, TYD = 16, RM = 2, RN = 4;
  using L = SmemLayout<TXD, TYD, RM, RN>;
  constexpr int TM = L::TM, TN = L::TN;
  constexpr int kThreads = TXD * TYD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sTab = reinterpret_cast<double*>(smem_raw);
  double2* sStage = reinterpret_cast<double2*>(sTab + L::kTabDoubles);
  double* sAux = reinterpret_cast<double*>(sStage + 2 * L::kStageDouble2);
  const int tid = threadIdx.x;
  const int tx = tid % TXD, ty = tid / TXD;
  for (int i = tid; i < L::kTabDoubles; i += kThreads) {
    const int j = i / kExpRepl;
    const double v = prm.exptab[j];
    sTab[i] = __hiloint2double(__double2hiint(v) - (j << (20 - kExpBits)), __double2loint(v));
  }
  const unsigned char* tab_bytes = smem_raw;
  const unsigned lane_bits = (unsigned)(tx & (kExpRepl - 1)) * 8u;
  const int D = prm.D, Dc = prm.Dc;
  const int num_chunks = (D + kDimChunk - 1) / kDimChunk;
  const int64_t nrows = prm.row_end - prm.row_begin;
  __syncthreads();

  auto issue_stage = [&](int64_t row0, int64_t col0, int ch, int buf) {
    const int d0 = ch * kDimChunk;
    const int nd = min(kDimChunk, D - d0);
    double2* dst = sStage + buf * L::kStageDouble2;
    for (int o = tid; o < TM + TN; o += kThreads) {
      const double2* src = (o < TM) ? prm.pts_row + (int64_t)d0 * prm.n_row_pad + prm.row_begin + row0 + o
                                    : prm.pts_col + (int64_t)d0 * prm.n_col_pad + col0 + (o - TM);
      const int64_t stride = (o < TM) ? prm.n_row_pad : prm.n_col_pad;
      for (int dl = 0; dl < nd; ++dl) cp_async16(dst + dl * (TM + TN) + o, src + dl * stride);
    }
    if (tid < nd) cp_async8(sAux + buf * kDimChunk + tid, prm.dim_aux + d0 + tid);
    cp_async_commit();
  };

  const int64_t row_blocks = (nrows + TM - 1) / TM;
  for (int64_t rb = blockIdx.x; rb < row_blocks; rb += gridDim.x) {
    const int64_t row0 = rb * TM;
    double rs[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) rs[r] = 0.0;
    for (int64_t cb = 0; cb < prm.tiles_n; ++cb) {
      const int64_t col0 = cb * TN;
      double acc[RM][RN][P];
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < RN; ++c)
#pragma unroll
          for (int p = 0; p < P; ++p) acc[r][c][p] = 0.0;
      int buf = 0;
      __syncthreads();  // the previous tile's last buffer is free
      issue_stage(row0, col0, 0, 0);
      for (int ch = 0; ch < num_chunks; ++ch) {
        cp_async_wait_all();
        __syncthreads();
        if (ch + 1 < num_chunks) issue_stage(row0, col0, ch + 1, buf ^ 1);
        const double2* sRow = sStage + buf * L::kStageDouble2;
        const double* aux = sAux + buf * kDimChunk;
        const int d0 = ch * kDimChunk;
        const int nd = min(kDimChunk, D - d0);
#pragma unroll 1
        for (int dl = 0; dl < nd; ++dl) {
          const double2* rowp = sRow + dl * (TM + TN);
          const double2* colp = rowp + TM;
          double2 rv[RM], cv[RN];
#pragma unroll
          for (int r = 0; r < RM; ++r) rv[r] = rowp[ty * RM + r];
#pragma unroll
          for (int c = 0; c < RN; ++c) cv[c] = colp[tx + TXD * c];
          if (d0 + dl < Dc) {
            const double ax = aux[dl];
#pragma unroll
            for (int r = 0; r < RM; ++r)
#pragma unroll
              for (int c = 0; c < RN; ++c) {
                const double d = rv[r].x - cv[c].x;
                const double e = exp_neg_scaled(fma(d, d, ax), tab_bytes, lane_bits);
                accumulate<P, ALGO>(acc[r][c], fma(-rv[r].y, cv[c].y, e));
              }
          } else {
            const double* tbl = prm.tables + (int)__double_as_longlong(aux[dl]);
#pragma unroll
            for (int r = 0; r < RM; ++r)
#pragma unroll
              for (int c = 0; c < RN; ++c)
                accumulate<P, ALGO>(acc[r][c], __ldg(tbl + __double2hiint(rv[r].x) + __double2loint(cv[c].x)));
          }
        }
        buf ^= 1;
      }
#pragma unroll
      for (int c = 0; c < RN; ++c) {
        const int64_t col = col0 + tx + TXD * c;
        const double a = col < prm.n2 ? __ldg(alpha + col) : 0.0;
#pragma unroll
        for (int r = 0; r < RM; ++r) rs[r] = fma(finish<P, ALGO>(acc[r][c], prm.sigma2), a, rs[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      double v = rs[r];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int64_t row = row0 + ty * RM + r;
      if (tx == 0 && row < nrows) out[row] = v;
    }
  }
  cp_async_wait_all();
}

template <int P>
static int launch_matvec(GramParams prm, int algo, int sms, const double* alpha, double* out, cudaStream_t stream) {
  using L = SmemLayout<16, 16, 2, 4>;
  prm.tiles_n = (prm.n2 + L::TN - 1) / L::TN;
  const int64_t row_blocks = (prm.row_end - prm.row_begin + L::TM - 1) / L::TM;
  const size_t smem = L::bytes(false);
  const int grid = (int)(row_blocks < sms ? row_blocks : sms);
  if (algo == OAK_ESP_DIRECT) {
    auto kern = gram_matvec_kernel<P, OAK_ESP_DIRECT>;
    OAK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, stream>>>(prm, alpha, out);
  } else {
    auto kern = gram_matvec_kernel<P, OAK_ESP_NEWTON_GIRARD>;
    OAK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, stream>>>(prm, alpha, out);
  }
  OAK_LAUNCHED();
  return 0;
}

int gram_launch(const oak_spec* spec, const double2* prow, int64_t n_row_pad, int64_t row_begin,
                int64_t row_end, const double2* pcol, int64_t n_col_pad, int64_t col_begin,
                int64_t col_end, int mode, double* K, int64_t ldk, cudaStream_t stream, double* Kt, int64_t ldkt,
                const double* ydot_y, double* ydot_part, int sm_reserve) {
  GramParams prm;
  prm.Kt = Kt;
  prm.ldkt = ldkt;
  prm.ydot_y = nullptr;
  prm.ydot_part = nullptr;
  if (ydot_y != nullptr) {
    OAK_REQUIRE(mode == 0 && spec->depth <= 4 && ydot_part != nullptr,
                "gram: the folded K y needs the general mode and max_interaction_depth <= 4");
    prm.ydot_y = ydot_y;
    prm.ydot_part = ydot_part;
  }
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) prm.sigma2[p] = spec->sigma2[p];
  prm.pts_row = prow;
  prm.pts_col = pcol;
  prm.n_row_pad = n_row_pad;
  prm.n_col_pad = n_col_pad;
  prm.dim_aux = spec->d_gram_aux;
  prm.tables = spec->d_tables;
  prm.tables_len = spec->tables_len;
  prm.exptab = spec->d_exptab;
  prm.K = K;
  prm.ldk = ldk;
  prm.row_begin = row_begin;
  prm.row_end = row_end;
  prm.col_begin = col_begin;
  prm.n2 = col_end - col_begin;
  prm.D = spec->D;
  prm.Dc = spec->Dc;
  prm.device = spec->device;
  static const int no_fast = env_int("OAK_GRAM_NOFAST", 0);
  prm.mm_row = no_fast ? nullptr : points_minmax(spec, prow, n_row_pad);
  prm.mm_col = no_fast ? nullptr : points_minmax(spec, pcol, n_col_pad);
  const int depth = spec->depth;
  const int T = tile_rows_for_depth(depth);
  // tile origins must sit on tile boundaries so that tile loads stay inside the padded block
  OAK_REQUIRE(row_begin % T == 0 && col_begin % T == 0,
              "gram: row/column range must start on a multiple of the tile size (64)");
  prm.symmetric = mode;
  prm.tile_row0 = 0;
  prm.tiles_n = 0;
  prm.num_tiles = 0;
  int sms = sm_count(spec->device);
  if (sm_reserve > 0) sms = sms - sm_reserve > 1 ? sms - sm_reserve : 1;
  if (sm_reserve < 0 && -sm_reserve < sms) sms = -sm_reserve;  // cap: the kernel itself is the small one
  const int algo = spec->algo;
  switch (depth) {
    case 0:
    case 1: return launch_small_depth<1>(prm, algo, sms, stream);
    case 2: return launch_small_depth<2>(prm, algo, sms, stream);
    case 3: return launch_small_depth<3>(prm, algo, sms, stream);
    case 4: return launch_small_depth<4>(prm, algo, sms, stream);
    // depth 5..8: 2 x 4 micro-tile (32 x 64 tiles): 8 entries x P accumulators, the same register
    // budget as 4 x 4 at depth 4, twice the work per staged operand of the former 2 x 2 geometry
    case 5: return launch_algo<5, 16, 16, 2, 4, 1>(prm, algo, sms, stream);
    case 6: return launch_algo<6, 16, 16, 2, 4, 1>(prm, algo, sms, stream);
    case 7: return launch_algo<7, 16, 16, 2, 4, 1>(prm, algo, sms, stream);
    case 8: return launch_algo<8, 16, 16, 2, 4, 1>(prm, algo, sms, stream);
    default:
      if (depth <= 12) return launch_algo<12, 16, 16, 2, 2, 1>(prm, algo, sms, stream);
      return launch_algo<16, 16, 16, 2, 2, 1>(prm, algo, sms, stream);
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

// out[i - row_begin] = sum_j K(x_i, x2_j) alpha_j for rows [row_begin, row_end): the mean of
// predict_f (Kus^T alpha) and any other Gram-matrix/vector product, without forming the matrix.
extern "C" int oak_gram_matvec_f64(const oak_spec* spec, const void* d_points, int64_t n, int64_t row_begin,
                                   int64_t row_end, const void* d_points2, int64_t n2, const double* d_alpha,
                                   double* d_out, void* stream_) {
  OAK_REQUIRE(spec && d_points && d_points2 && d_alpha && d_out, "oak_gram_matvec_f64: null argument");
  OAK_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n, "oak_gram_matvec_f64: bad row range");
  OAK_REQUIRE(row_begin % 64 == 0, "oak_gram_matvec_f64: row_begin must be a multiple of 64");
  if (row_end == row_begin) return 0;
  OAK_REQUIRE(n2 >= 1, "oak_gram_matvec_f64: empty conditioning set");
  const int depth = spec->depth < 1 ? 1 : spec->depth;
  OAK_REQUIRE(depth <= 8, "oak_gram_matvec_f64: max_interaction_depth > 8 is not supported");
  GramParams prm;
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) prm.sigma2[p] = spec->sigma2[p];
  prm.pts_row = (const double2*)d_points;
  prm.pts_col = (const double2*)d_points2;
  prm.n_row_pad = padded(n);
  prm.n_col_pad = padded(n2);
  prm.dim_aux = spec->d_gram_aux;
  prm.tables = spec->d_tables;
  prm.tables_len = spec->tables_len;
  prm.exptab = spec->d_exptab;
  prm.K = nullptr;
  prm.ldk = 0;
  prm.row_begin = row_begin;
  prm.row_end = row_end;
  prm.col_begin = 0;
  prm.n2 = n2;
  prm.D = spec->D;
  prm.Dc = spec->Dc;
  prm.device = spec->device;
  prm.symmetric = 0;
  prm.Kt = nullptr;
  prm.ldkt = 0;
  prm.ydot_y = nullptr;
  prm.ydot_part = nullptr;
  prm.tile_row0 = 0;
  prm.num_tiles = 0;
  prm.mm_row = prm.mm_col = nullptr;
  prm.use_tma = 0;
  const int sms = sm_count(spec->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  switch (depth) {
    case 1: return launch_matvec<1>(prm, spec->algo, sms, d_alpha, d_out, stream);
    case 2: return launch_matvec<2>(prm, spec->algo, sms, d_alpha, d_out, stream);
    case 3: return launch_matvec<3>(prm, spec->algo, sms, d_alpha, d_out, stream);
    case 4: return launch_matvec<4>(prm, spec->algo, sms, d_alpha, d_out, stream);
    case 5: return launch_matvec<5>(prm, spec->algo, sms, d_alpha, d_out, stream);
    case 6: return launch_matvec<6>(prm, spec->algo, sms, d_alpha, d_out, stream);
    case 7: return launch_matvec<7>(prm, spec->algo, sms, d_alpha, d_out, stream);
    default: return launch_matvec<8>(prm, spec->algo, sms, d_alpha, d_out, stream);
  }
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

// The same lower trapezoid plus its mirror image: d_Kt is an (row_end x (row_end - row_begin)) block with
// d_Kt[j * ldkt + (i - row_begin)] = K(i, j) for every entry of the off-diagonal tiles (the diagonal tiles are
// written in full to d_K), so that the strips of all ranks together hold the whole symmetric matrix -- the
// product of the single-GPU call -- without a collective.
extern "C" int oak_gram_lower_mirror_f64(const oak_spec* spec, const void* d_points, int64_t n, int64_t row_begin,
                                         int64_t row_end, double* d_K, int64_t ldk, double* d_Kt, int64_t ldkt,
                                         void* stream_) {
  OAK_REQUIRE(spec && d_points, "oak_gram_lower_mirror_f64: null argument");
  OAK_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n,
              "oak_gram_lower_mirror_f64: row range outside [0, n]");
  if (row_end == row_begin) return 0;
  OAK_REQUIRE(d_K && d_Kt, "oak_gram_lower_mirror_f64: null output");
  OAK_REQUIRE(ldk >= row_end && ldkt >= row_end - row_begin, "oak_gram_lower_mirror_f64: leading dimension too small");
  return gram_launch(spec, (const double2*)d_points, padded(n), row_begin, row_end, (const double2*)d_points, padded(n),
                     0, row_end, 3, d_K, ldk, (cudaStream_t)stream_, d_Kt, ldkt);
}
