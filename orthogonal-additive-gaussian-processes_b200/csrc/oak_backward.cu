// Backward pass of the fused OAK tiles: the reference trains by
// TensorFlow autodiff through OAKKernel.K (oak/oak_kernel.py:223-265) inside gpflow's objectives
// (oak/model_utils.py:168-175).  Here the same contraction is done tile by tile:
//
//   grad[l_d]       += sum_ij W_ij * dK_ij/dk~_d * dk~_d/dl_d        (RBF dims, Gaussian measure or none)
//   grad[sigma2_n]  += sum_ij W_ij * e_n(k~_1..k~_D)_ij              (n = 0..P)
//
// for a caller-supplied cotangent W (d objective / d K).  Per entry, sweep 1 over the dimensions
// builds e_1..e_P with the direct recurrence (registers), sweep 2 recomputes each k~_d and forms
//   dK/dk~_d = sum_n sigma2_n e_{n-1}^{(-d)},  e_m^{(-d)} = e_m - k~_d e_{m-1}^{(-d)}   (removal recurrence)
//   dk~/dl   = ex * z * 2/l - c^i c^j (kappa + u_i + u_j)
// with ex = s^2 exp(-z), z = (x-y)^2/(2 l^2), c^ = cov_X_s / sqrt(var_s), and for the Gaussian measure
// N(mu, delta^2): kappa = 1/l + l/(l^2+2 delta^2) - 2l/(l^2+delta^2), u(x) = (x-mu)^2 l/(l^2+delta^2)^2
// (ortho_rbf_kernel.py:82-97 differentiated by hand; checked against torch autograd in tests/).
// No per-dimension N x N2 derivative matrix is ever formed; W is read once per entry.
// Reductions are deterministic: warp shuffles -> per-warp shared slots -> per-CTA partials ->
// one fixed-order pass over the CTAs.
//
// Row-point gradients (the inducing points Z of SGPR when zfixed=False, model_utils.py:98-101):
//   d k~_d(z, x)/dz = -ex (z - x)/l^2 - c^'(z) c^(x)
// so per (row, dim) the tiles accumulate A = sum_j W dK/dk~ ex d' and B = sum_j W dK/dk~ c^(x_j)
// (d' the prepared coordinate difference) over a contiguous segment of column tiles in shared
// memory, write one partial per (segment, row, dim) and a finishing kernel forms
//   dZ[m, d] = -A / (xscale l^2) - c^'(z_m) B      (c^' per measure, ortho_rbf_kernel.py:49-136).
#include <cmath>

#include "oak_common.cuh"

namespace oak {

namespace bw {
#ifndef OAK_BW_DIMCHUNK
#define OAK_BW_DIMCHUNK 16  // dims per pipeline stage (A/B aid: 8 halves the staging buffers for two CTAs per SM)
#endif
constexpr int kDimChunk = OAK_BW_DIMCHUNK;
constexpr int kThreads = 256;
constexpr int kTXD = 16, kTYD = 16;
constexpr int kMaxDims = 512;  // per-warp shared slots for the lengthscale partials
}  // namespace bw

struct BwDim {      // per sub-kernel (kernel order)
  double half_kappa;  // Gaussian measure: d c^/dl = c^ (kappa/2 + u(x))
  double uc;        // l / (l^2 + delta^2)^2
  double mu;
  double inv_xscale;  // prepared coordinate -> x
  double c2;        // (2 / l) * ln2 / kExpTab : ex * d'^2 * c2 = ex * z * 2 / l
  double inv_s2;    // 1 / s^2 (base variance)
  double zc;        // 1 / (xscale l^2): prepared-coordinate difference -> (z - x)/l^2
  double kind;      // 0: no lengthscale gradient (discrete dim / uniform / MOG measure)
                    // 1: closed form above (Gaussian measure, or no measure: c^ = 0)
                    // 2: d c^/dl read from the per-point array written by oak_prepare_backward_f64
                    //    (empirical measure)
};

struct BwParams {
  double sigma2[OAK_MAX_DEPTH + 1];
  const double2* pts_row;
  const double2* pts_col;
  const double* dim_aux;
  const double* tables;
  const double* exptab;
  const BwDim* bwdims;
  const double* dch_row;  // [D][n_row_pad] d c^/dl per point (kind 2 dims), may be null
  const double* dch_col;
  const unsigned long long* mm_row;  // [2 D] min / max keys of the prepared coordinates (see oak_prepare.cu)
  const unsigned long long* mm_col;
  const double* W;
  double* partial;  // [grid][D + P + 1 + tables_len]
  double* zpartial;  // row-gradient variant: [segs][zrows][D][2] partial (A, B) sums
  int64_t zrows;     // row blocks * tile rows
  int segs;          // column segments per row block
  int64_t n_row_pad, n_col_pad, ldw;
  int64_t row_begin, row_end, n2;
  int64_t tiles_n, num_tiles;
  int D, Dc, tables_len;  // gradient layout: [D lengthscales | P + 1 order variances | tables_len table
                          // entries | D base variances s^2 of the RBF sub-kernels]
};

// s^2 exp(-z) of one entry: FAST = clamp-free body (s^2 == 1 and bounded distances, proven per launch
// from the min/max keys exactly as in the forward kernel), else the general clamped body
template <bool FAST>
__device__ __forceinline__ double entry_exp(double d, double ax, const unsigned char* tab_bytes, unsigned lane_bits) {
  if constexpr (FAST) return exp_neg_sq_fast(d, tab_bytes, lane_bits);
  return exp_neg_scaled(fma(d, d, ax), tab_bytes, lane_bits);
}

#ifndef OAK_BW_MINB
#define OAK_BW_MINB 1  // A/B aid: 2 = two CTAs per SM (needs the 2 x 4 micro-tile to stay within 128 registers)
#endif
#ifndef OAK_BW_RM3
#define OAK_BW_RM3 4   // rows of the micro-tile at depth <= 3
#endif
#ifndef OAK_BW_RN3
#define OAK_BW_RN3 4
#endif
template <int P, int RM, int RN, bool ZG>
__global__ void __launch_bounds__(bw::kThreads, OAK_BW_MINB) gram_backward_kernel(const BwParams prm) {
  using namespace bw;
  constexpr int TM = kTYD * RM, TN = kTXD * RN;
  constexpr int kTabDoubles = kExpTab * kExpRepl;
  constexpr int kStageDouble2 = kDimChunk * (TM + TN);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sTab = reinterpret_cast<double*>(smem_raw);
  double2* sStage = reinterpret_cast<double2*>(sTab + kTabDoubles);
  double* sAux = reinterpret_cast<double*>(sStage + 2 * kStageDouble2);
  double* sG = sAux + 2 * kDimChunk;  // [8 warps][D + P + 1 + tables_len]

  const int tid = threadIdx.x;
  const int tx = tid % kTXD, ty = tid / kTXD;
  const int lane = tid & 31, warp = tid >> 5;
  const int D = prm.D, Dc = prm.Dc;
  const int nout = 2 * D + P + 1 + prm.tables_len;
  const int vbase = D + P + 1 + prm.tables_len;  // first base-variance slot
  double* sZ = sG + 8 * nout;  // ZG: [TM][D][2] (A, B) sums of the current (row block, segment)
  for (int i = tid; i < kTabDoubles; i += kThreads) {
    const int j = i / kExpRepl;
    const double v = prm.exptab[j];
    sTab[i] = __hiloint2double(__double2hiint(v) - (j << (20 - kExpBits)), __double2loint(v));
  }
  for (int i = tid; i < 8 * nout; i += kThreads) sG[i] = 0.0;
  const unsigned char* tab_bytes = smem_raw;
  const unsigned lane_bits = (unsigned)(tx & (kExpRepl - 1)) * 8u;
  const int num_chunks = (D + kDimChunk - 1) / kDimChunk;
  double* myG = sG + warp * nout;
  __shared__ int sSlow;
  if (tid == 0) sSlow = 0;
  __syncthreads();
  for (int d = tid; d < Dc; d += kThreads) {
    const double rmin = order_key_decode(prm.mm_row[d]), rmax = order_key_decode(prm.mm_row[D + d]);
    const double cmin = order_key_decode(prm.mm_col[d]), cmax = order_key_decode(prm.mm_col[D + d]);
    const double span = fmax(rmax - cmin, cmax - rmin);
    if (!(prm.dim_aux[d] == 0.0 && span <= kFastSpan)) atomicOr(&sSlow, 1);
  }
  __syncthreads();
  const bool fast = sSlow == 0;

  auto issue_stage = [&](int64_t row0, int64_t col0, int ch, int buf) {
    const int d0 = ch * kDimChunk;
    const int nd = min(kDimChunk, D - d0);
    double2* dst = sStage + buf * kStageDouble2;
    for (int o = tid; o < TM + TN; o += kThreads) {
      const double2* src = (o < TM) ? prm.pts_row + (int64_t)d0 * prm.n_row_pad + prm.row_begin + row0 + o
                                    : prm.pts_col + (int64_t)d0 * prm.n_col_pad + col0 + (o - TM);
      const int64_t stride = (o < TM) ? prm.n_row_pad : prm.n_col_pad;
      for (int dl = 0; dl < nd; ++dl) cp_async16(dst + dl * (TM + TN) + o, src + dl * stride);
    }
    if (tid < nd) cp_async8(sAux + buf * kDimChunk + tid, prm.dim_aux + d0 + tid);
    cp_async_commit();
  };

  const int64_t nrows = prm.row_end - prm.row_begin;
  // work units: one tile, or (ZG) one row block x one contiguous segment of column tiles
  const int64_t units = ZG ? (prm.zrows / TM) * prm.segs : prm.num_tiles;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    int64_t bi, ct0, ct1;
    int seg = 0;
    if constexpr (ZG) {
      bi = u / prm.segs;
      seg = (int)(u - bi * prm.segs);
      ct0 = prm.tiles_n * seg / prm.segs;
      ct1 = prm.tiles_n * (seg + 1) / prm.segs;
      for (int i = tid; i < TM * D * 2; i += kThreads) sZ[i] = 0.0;
      __syncthreads();
    } else {
      bi = u / prm.tiles_n;
      ct0 = u - bi * prm.tiles_n;
      ct1 = ct0 + 1;
    }
    const int64_t row0 = bi * TM;
    for (int64_t ct = ct0; ct < ct1; ++ct) {
      const int64_t col0 = ct * TN;

      double E[RM][RN][P];  // e_1..e_P (direct recurrence)
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < RN; ++c)
#pragma unroll
          for (int p = 0; p < P; ++p) E[r][c][p] = 0.0;

      // ---- sweep 1: elementary symmetric polynomials ------------------------------------------
      int buf = 0;
      issue_stage(row0, col0, 0, 0);
      for (int ch = 0; ch < num_chunks; ++ch) {
        cp_async_wait_all();
        __syncthreads();
        if (ch + 1 < num_chunks) issue_stage(row0, col0, ch + 1, buf ^ 1);
        const double2* sRow = sStage + buf * kStageDouble2;
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
          for (int c = 0; c < RN; ++c) cv[c] = colp[tx + kTXD * c];
          const bool cont = d0 + dl < Dc;
          const double ax = aux[dl];
          auto fold = [&](int r, int c, double k) {
#pragma unroll
            for (int p = P - 1; p >= 1; --p) E[r][c][p] = fma(k, E[r][c][p - 1], E[r][c][p]);
            E[r][c][0] += k;
          };
          if (cont && fast) {
#pragma unroll
            for (int r = 0; r < RM; ++r)
#pragma unroll
              for (int c = 0; c < RN; ++c)
                fold(r, c, fma(-rv[r].y, cv[c].y, entry_exp<true>(rv[r].x - cv[c].x, ax, tab_bytes, lane_bits)));
          } else if (cont) {
#pragma unroll
            for (int r = 0; r < RM; ++r)
#pragma unroll
              for (int c = 0; c < RN; ++c)
                fold(r, c, fma(-rv[r].y, cv[c].y, entry_exp<false>(rv[r].x - cv[c].x, ax, tab_bytes, lane_bits)));
          } else {
            const double* tbl = prm.tables + (int)__double_as_longlong(ax);
#pragma unroll
            for (int r = 0; r < RM; ++r)
#pragma unroll
              for (int c = 0; c < RN; ++c)
                fold(r, c, __ldg(tbl + __double2hiint(rv[r].x) + __double2loint(cv[c].x)));
          }
        }
        buf ^= 1;
      }

      // ---- cotangent tile + order-variance gradients --------------------------------------------
      double wv[RM][RN];
      double gs[P + 1];
#pragma unroll
      for (int p = 0; p <= P; ++p) gs[p] = 0.0;
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < RN; ++c) {
          const int64_t row = row0 + ty * RM + r, col = col0 + tx + kTXD * c;
          wv[r][c] = (row < nrows && col < prm.n2) ? __ldcs(prm.W + row * prm.ldw + col) : 0.0;
          gs[0] += wv[r][c];
#pragma unroll
          for (int p = 1; p <= P; ++p) gs[p] = fma(wv[r][c], E[r][c][p - 1], gs[p]);
        }
#pragma unroll
      for (int p = 0; p <= P; ++p) {
        double v = gs[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) myG[D + p] += v;
      }

      // ---- sweep 2: lengthscale gradients ---------------------------------------------------------
      __syncthreads();  // everybody has left sweep 1's last buffer before it is overwritten
      buf = 0;
      issue_stage(row0, col0, 0, 0);
      // The warp sums of a dimension's two partials (d/dl, d/ds^2) are folded one butterfly round per micro-tile row
      // of the NEXT dimension: five dependent shuffle + add rounds per dimension would otherwise sit between two
      // dimension bodies with nothing to cover them but the scheduler's second warp, which is in step.  Same
      // rounds, same order: the sums are bit-identical to the immediate reduction.  (Measured: -2.5 % on the plain
      // tiles; the row-gradient variant, at 255 registers with its own half-warp sums, loses 2 % and keeps the
      // immediate reduction: profiles/r02bd_ab_backward_deferred_butterfly.txt.)
      constexpr bool kDefer = !ZG;
      double pend_a = 0.0, pend_v = 0.0, pend_inv_s2 = 0.0;
      int pend_slot = -1;      // dimension whose partials are in flight
      bool pend_ls = false;    // ... and whether it has a lengthscale gradient
      auto pend_round = [&](int o) {
        pend_a += __shfl_xor_sync(0xffffffffu, pend_a, o);
        pend_v += __shfl_xor_sync(0xffffffffu, pend_v, o);
      };
      auto pend_commit = [&]() {  // after the last round
        if (lane == 0 && pend_slot >= 0) {
          if (pend_ls) myG[pend_slot] += pend_a;
          myG[vbase + pend_slot] += pend_v * pend_inv_s2;
        }
      };
      for (int ch = 0; ch < num_chunks; ++ch) {
        cp_async_wait_all();
        __syncthreads();
        if (ch + 1 < num_chunks) issue_stage(row0, col0, ch + 1, buf ^ 1);
        const double2* sRow = sStage + buf * kStageDouble2;
        const double* aux = sAux + buf * kDimChunk;
        const int d0 = ch * kDimChunk;
        const int nc = max(0, min(min(kDimChunk, D - d0), Dc - d0));  // continuous dims of this chunk
#pragma unroll 1
        for (int dl = 0; dl < nc; ++dl) {
          const BwDim bd = prm.bwdims[d0 + dl];
          const double2* rowp = sRow + dl * (TM + TN);
          const double2* colp = rowp + TM;
          double2 rv[RM], cv[RN];
          double dr[RM], dc[RN];  // d c^/dl of the row / column points
#pragma unroll
          for (int r = 0; r < RM; ++r) rv[r] = rowp[ty * RM + r];
#pragma unroll
          for (int c = 0; c < RN; ++c) cv[c] = colp[tx + kTXD * c];
          if (bd.kind == 0.0) {  // no lengthscale gradient (uniform / MOG measure): base variance only
#pragma unroll
            for (int r = 0; r < RM; ++r) dr[r] = 0.0;
#pragma unroll
            for (int c = 0; c < RN; ++c) dc[c] = 0.0;
          } else if (bd.kind == 1.0) {
#pragma unroll
            for (int r = 0; r < RM; ++r) {
              const double xm = fma(rv[r].x, bd.inv_xscale, -bd.mu);
              dr[r] = rv[r].y * fma(xm * xm, bd.uc, bd.half_kappa);
            }
#pragma unroll
            for (int c = 0; c < RN; ++c) {
              const double xm = fma(cv[c].x, bd.inv_xscale, -bd.mu);
              dc[c] = cv[c].y * fma(xm * xm, bd.uc, bd.half_kappa);
            }
          } else {
            const double* gr = prm.dch_row + (int64_t)(d0 + dl) * prm.n_row_pad + prm.row_begin + row0 + ty * RM;
            const double* gc = prm.dch_col + (int64_t)(d0 + dl) * prm.n_col_pad + col0 + tx;
#pragma unroll
            for (int r = 0; r < RM; ++r) dr[r] = __ldg(gr + r);
#pragma unroll
            for (int c = 0; c < RN; ++c) dc[c] = __ldg(gc + kTXD * c);
          }
          const double ax = aux[dl];
          double acc = 0.0, accv = 0.0;
          double za[RM], zb[RM];
#pragma unroll
          for (int r = 0; r < RM; ++r) za[r] = zb[r] = 0.0;
          auto entry = [&](int r, int c, double d, double ex) {
            const double cc = rv[r].y * cv[c].y;
            const double k = ex - cc;
            // removal recurrence: g_m = e_m - k g_{m-1}; dK/dk = sum_n sigma2_n g_{n-1}
            double g = 1.0;
            double dKdk = prm.sigma2[1];
#pragma unroll
            for (int m = 1; m < P; ++m) {
              g = fma(-k, g, E[r][c][m - 1]);
              dKdk = fma(prm.sigma2[m + 1], g, dKdk);
            }
            const double dkdl = fma(ex * (d * d), bd.c2, -fma(dr[r], cv[c].y, rv[r].y * dc[c]));
            const double wk = wv[r][c] * dKdk;
            acc = fma(wk, dkdl, acc);
            accv = fma(wk, k, accv);  // k~ is homogeneous of degree one in s^2: d k~ / d s^2 = k~ / s^2
            if constexpr (ZG) {
              za[r] = fma(wk * ex, d, za[r]);
              zb[r] = fma(wk, cv[c].y, zb[r]);
            }
          };
          // rounds 16, 8, 4, 2 of the PREVIOUS dimension's butterfly after the rows of this one (RM = 4: one each,
          // RM = 2: two each), round 1 and the commit behind the last row
          constexpr int kRoundsPerRow = 4 / RM;
          static_assert(RM == 2 || RM == 4, "the deferred butterfly deals four rounds over the micro-tile rows");
          if (fast) {
#pragma unroll
            for (int r = 0; r < RM; ++r) {
#pragma unroll
              for (int c = 0; c < RN; ++c) {
                const double d = rv[r].x - cv[c].x;
                entry(r, c, d, entry_exp<true>(d, ax, tab_bytes, lane_bits));
              }
              if constexpr (kDefer) {
#pragma unroll
                for (int q = 0; q < kRoundsPerRow; ++q) pend_round(16 >> (r * kRoundsPerRow + q));
              }
            }
          } else {
#pragma unroll
            for (int r = 0; r < RM; ++r) {
#pragma unroll
              for (int c = 0; c < RN; ++c) {
                const double d = rv[r].x - cv[c].x;
                entry(r, c, d, entry_exp<false>(d, ax, tab_bytes, lane_bits));
              }
              if constexpr (kDefer) {
#pragma unroll
                for (int q = 0; q < kRoundsPerRow; ++q) pend_round(16 >> (r * kRoundsPerRow + q));
              }
            }
          }
          if constexpr (kDefer) {
            pend_round(1);
            pend_commit();
          }
          pend_a = acc;
          pend_v = accv;
          pend_inv_s2 = bd.inv_s2;
          pend_slot = d0 + dl;
          pend_ls = bd.kind != 0.0;
          if constexpr (!kDefer) {  // immediate reduction
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pend_round(o);
            pend_commit();
            pend_slot = -1;
          }
          if constexpr (ZG) {  // the 16 threads of a half warp share their rows
#pragma unroll
            for (int r = 0; r < RM; ++r) {
              double a = za[r], b = zb[r];
#pragma unroll
              for (int o = 8; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
              }
              if (tx == 0) {
                double* z = sZ + ((ty * RM + r) * D + d0 + dl) * 2;
                z[0] += a;
                z[1] += b;
              }
            }
          }
        }
        // discrete dims: cotangent of the table blob (the host chains it to W / kappa / variance)
        const int nd2 = min(kDimChunk, D - d0);
#pragma unroll 1
        for (int dl = nc; dl < nd2; ++dl) {
          const double2* rowp = sRow + dl * (TM + TN);
          const double2* colp = rowp + TM;
          const int toff = (int)__double_as_longlong(aux[dl]);
          const double* tbl = prm.tables + toff;
          double* gt = myG + D + P + 1 + toff;
          int ro[RM], co[RN];
#pragma unroll
          for (int r = 0; r < RM; ++r) ro[r] = __double2hiint(rowp[ty * RM + r].x);
#pragma unroll
          for (int c = 0; c < RN; ++c) co[c] = __double2loint(colp[tx + kTXD * c].x);
#pragma unroll
          for (int r = 0; r < RM; ++r)
#pragma unroll
            for (int c = 0; c < RN; ++c) {
              if (wv[r][c] == 0.0) continue;  // out-of-range entries carry zero cotangent
              const double k = __ldg(tbl + ro[r] + co[c]);
              double g = 1.0;
              double dKdk = prm.sigma2[1];
#pragma unroll
              for (int m = 1; m < P; ++m) {
                g = fma(-k, g, E[r][c][m - 1]);
                dKdk = fma(prm.sigma2[m + 1], g, dKdk);
              }
              atomicAdd(gt + ro[r] + co[c], wv[r][c] * dKdk);  // per-warp slots: intra-warp order only
            }
        }
        buf ^= 1;
      }
      // the last continuous dimension's partials
      if constexpr (kDefer) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pend_round(o);
        pend_commit();
      }
      __syncthreads();  // stage buffers are re-issued by the next tile
    }
    if constexpr (ZG) {
      double* dst = prm.zpartial + ((int64_t)seg * prm.zrows + row0) * (2 * D);
      for (int i = tid; i < TM * D * 2; i += kThreads) dst[i] = sZ[i];
      __syncthreads();
    }
  }

  // ---- per-CTA partials (fixed order over the warps) ---------------------------------------------
  __syncthreads();
  for (int i = tid; i < nout; i += kThreads) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += sG[w * nout + i];
    prm.partial[(int64_t)blockIdx.x * nout + i] = v;
  }
}

// grad[i] += sum over CTAs (fixed order); the lengthscale part is scattered to the caller's
// dimension order
__global__ void backward_reduce_kernel(const double* __restrict__ partial, int grid, int nout, int D,
                                       const int* __restrict__ orig_of_pos, double* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout) return;
  double v = 0.0;
  for (int b = 0; b < grid; ++b) v += partial[(int64_t)b * nout + i];
  int dst = i;
  if (i < D) dst = orig_of_pos[i];                                   // lengthscales
  else if (i >= nout - D) dst = nout - D + orig_of_pos[i - (nout - D)];  // base variances
  grad[dst] += v;
}

// K_diag backward: one thread per point; grad += sum_i w_i dK_diag(x_i)/d theta
struct DiagBwParams {
  double sigma2[OAK_MAX_DEPTH + 1];
  int D, Dc, depth, tables_len;
};

__global__ void __launch_bounds__(256) diag_backward_kernel(DiagBwParams prm, const DimDev* __restrict__ dims,
                                                            const BwDim* __restrict__ bwdims,
                                                            const double2* __restrict__ pts,
                                                            const double* __restrict__ dch, int64_t n,
                                                            int64_t n_pad, const double* __restrict__ w,
                                                            double wscale, double* __restrict__ partial) {
  extern __shared__ double sh[];  // [8 warps][nout]
  const int P = prm.depth;
  const int nout = 2 * prm.D + P + 1 + prm.tables_len;
  const int vbase = prm.D + P + 1 + prm.tables_len;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 8 * nout; i += 256) sh[i] = 0.0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 256 + tid;
  const bool live = i < n;
  double e[OAK_MAX_DEPTH + 1];
#pragma unroll
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) e[p] = 0.0;
  e[0] = 1.0;
  const double wi = live ? (w ? w[i] : 1.0) * wscale : 0.0;
  for (int k = 0; k < prm.D; ++k) {
    const double2 v = pts[(int64_t)k * n_pad + (live ? i : 0)];
    const double kd = (k < prm.Dc) ? dims[k].s2 - v.y * v.y : v.y;
#pragma unroll
    for (int p = OAK_MAX_DEPTH; p >= 1; --p)
      if (p <= P) e[p] = fma(kd, e[p - 1], e[p]);
  }
  for (int p = 0; p <= P; ++p) {
    double v = wi * e[p];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp * nout + prm.D + p] += v;
  }
  for (int k = 0; k < prm.Dc; ++k) {
    const BwDim bd = bwdims[k];
    const double2 v = pts[(int64_t)k * n_pad + (live ? i : 0)];
    const double cc = v.y * v.y;
    const double kd = dims[k].s2 - cc;
    double g = 1.0, dKdk = prm.sigma2[1];
    for (int m = 1; m < P; ++m) {
      g = fma(-kd, g, e[m]);
      dKdk = fma(prm.sigma2[m + 1], g, dKdk);
    }
    double dchat = 0.0;
    if (bd.kind == 1.0) {
      const double xm = fma(v.x, bd.inv_xscale, -bd.mu);
      dchat = v.y * fma(xm * xm, bd.uc, bd.half_kappa);
    } else if (bd.kind == 2.0) {
      dchat = dch[(int64_t)k * n_pad + (live ? i : 0)];
    }
    double c = wi * dKdk * (-2.0 * v.y * dchat);
    double cv = wi * dKdk * kd * bd.inv_s2;
    for (int o = 16; o > 0; o >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, o);
      cv += __shfl_xor_sync(0xffffffffu, cv, o);
    }
    if (lane == 0) {
      if (bd.kind != 0.0) sh[warp * nout + k] += c;
      sh[warp * nout + vbase + k] += cv;
    }
  }
  // discrete dims: K_diag reads tables[table_off + C*C + idx]
  for (int k = prm.Dc; k < prm.D; ++k) {
    if (!live || wi == 0.0) continue;
    const DimDev dd = dims[k];
    const double2 v = pts[(int64_t)k * n_pad + i];
    double g = 1.0, dKdk = prm.sigma2[1];
    for (int m = 1; m < P; ++m) {
      g = fma(-v.y, g, e[m]);
      dKdk = fma(prm.sigma2[m + 1], g, dKdk);
    }
    atomicAdd(sh + warp * nout + prm.D + P + 1 + dd.table_off + dd.count * dd.count + __double2loint(v.x),
              wi * dKdk);
  }
  __syncthreads();
  for (int j = tid; j < nout; j += 256) {
    double v = 0.0;
    for (int wq = 0; wq < 8; ++wq) v += sh[wq * nout + j];
    partial[(int64_t)blockIdx.x * nout + j] = v;
  }
}

// ---- d c^/dl for empirical-measure dims (ortho_rbf_kernel.py:101-120 differentiated) -------------
//   c(x) = s^2 sum_q w_q exp(-t_q^2),  t_q = (x - s_q)/(sqrt(2) l):   c'(x) = s^2 sum_q w_q exp(-t_q^2) 2 t_q^2 / l
//   v    = s^2 sum_pq w_p w_q exp(-t_pq^2):                            v'    = s^2 sum_pq w_p w_q exp(-t_pq^2) 2 t_pq^2 / l
//   c^ = c / sqrt(v):   d c^/dl = c'/sqrt(v) - c^ v'/(2 v)
// v' uses the same deterministic two-pass reduction as var_s in oak_spec.cu.
__global__ void empirical_dvar_partial(const double* __restrict__ loc, const double* __restrict__ w, int m,
                                       double inv_sqrt2_l, double scale, double* __restrict__ partial) {
  extern __shared__ double sh[];
  const int tile = blockDim.x;
  double* sl = sh;
  double* sw = sh + tile;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double zi = (i < m) ? loc[i] * inv_sqrt2_l : 0.0;
  const double wi = (i < m) ? w[i] : 0.0;
  double acc = 0.0;
  for (int base = 0; base < m; base += tile) {
    const int j = base + threadIdx.x;
    sl[threadIdx.x] = (j < m) ? loc[j] * inv_sqrt2_l : 0.0;
    sw[threadIdx.x] = (j < m) ? w[j] : 0.0;
    __syncthreads();
    const int lim = min(tile, m - base);
    for (int k = 0; k < lim; ++k) {
      const double t = zi - sl[k];
      acc = fma(sw[k] * (t * t), exp(-t * t), acc);
    }
    __syncthreads();
  }
  acc *= wi * scale;  // scale = s^2 * 2 / l
  sl[threadIdx.x] = acc;
  __syncthreads();
  for (int s = tile / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sl[threadIdx.x] += sl[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sl[0];
}

__global__ void __launch_bounds__(256) empirical_dch_kernel(const double* __restrict__ loc, const double* __restrict__ w,
                                                            int m, double inv_sqrt2_l, double s2, double two_over_l,
                                                            const double* __restrict__ inv_sqrt_v_slot,
                                                            const double* __restrict__ dvar_partial, int nblocks,
                                                            const double2* __restrict__ pts_k, int64_t n,
                                                            double* __restrict__ out_k) {
  __shared__ double sl[256], sw[256];
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const double isv = *inv_sqrt_v_slot;
  double dv = 0.0;
  for (int b = 0; b < nblocks; ++b) dv += dvar_partial[b];
  const double half_dlogv = 0.5 * dv * isv * isv;
  const double2 p = (i < n) ? pts_k[i] : make_double2(0.0, 0.0);
  const double a = p.x * (1.0 / kXScale);  // x / (sqrt(2) l)
  double acc = 0.0;
  for (int base = 0; base < m; base += 256) {
    const int j = base + threadIdx.x;
    sl[threadIdx.x] = (j < m) ? loc[j] * inv_sqrt2_l : 0.0;
    sw[threadIdx.x] = (j < m) ? w[j] : 0.0;
    __syncthreads();
    const int lim = min(256, m - base);
    for (int q = 0; q < lim; ++q) {
      const double t = a - sl[q];
      acc = fma(sw[q] * (t * t), exp(-t * t), acc);
    }
    __syncthreads();
  }
  if (i < n) out_k[i] = s2 * two_over_l * acc * isv - p.y * half_dlogv;
}

// d c^/dl per point for the uniform and mixture-of-Gaussians measures (ortho_rbf_kernel.py:49-63,
// 124-136 differentiated); half_dlogv = v'/(2 v) comes from the host (closed forms, :65-78, :138-152).
//   uniform [a, b]:  c = s^2 l sqrt(pi/2)/(b-a) [erf(tb) - erf(ta)],  t. = (. - x)/(sqrt(2) l)
//                    c' = c/l - s^2 sqrt(2)/(b-a) [tb exp(-tb^2) - ta exp(-ta^2)]
//   MOG:             c = s^2 l sum_k w_k g_k,  g_k = exp(-(x-m_k)^2/(2 S_k))/sqrt(S_k),  S_k = l^2 + var_k
//                    c' = c/l + s^2 l sum_k w_k g_k [(x-m_k)^2 l/S_k^2 - l/S_k]
__global__ void __launch_bounds__(256) measure_dch_kernel(DimDev d, const double* __restrict__ inv_sqrt_v_slot,
                                                          double half_dlogv, const double2* __restrict__ pts_k,
                                                          int64_t n, double* __restrict__ out_k) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double2 p = pts_k[i];
  const double x = p.x / d.xscale;
  const double l = d.lengthscale;
  const double isv = *inv_sqrt_v_slot;
  double dc = 0.0;
  if (d.measure == OAK_MEASURE_UNIFORM) {
    const double a = d.c1, b = d.c2;
    const double ta = (a - x) * d.inv_sqrt2_l, tb = (b - x) * d.inv_sqrt2_l;
    const double c = d.c0 * (erf(tb) - erf(ta));
    dc = c / l - d.s2 * sqrt(2.0) / (b - a) * (tb * exp(-tb * tb) - ta * exp(-ta * ta));
  } else {  // MOG
    double c = 0.0, extra = 0.0;
    for (int q = 0; q < d.count; ++q) {
      const double S = l * l + d.v1[q];
      const double t = x - d.v0[q];
      const double g = exp(-0.5 * t * t / S) / sqrt(S) * d.v2[q];
      c += g;
      extra = fma(g, t * t * l / (S * S) - l / S, extra);
    }
    dc = d.c0 * c / l + d.c0 * extra;  // c0 = s^2 l
  }
  out_k[i] = dc * isv - p.y * half_dlogv;
}

// ---- row-point gradients: finishing pass ---------------------------------------------------------
// One warp per (row m, continuous dim k): fixed-order sum of the segment partials, c^'(z_m) of the
// dim's measure (lanes stride the locations / components), dZ[m, orig] += -A zc - c^' B.
//   Gaussian:  c = c0 exp(-(x-mu)^2 c2)                      c' = -2 (x-mu) c2 c
//   uniform:   c = c0 [erf(tb) - erf(ta)]                    c' = -c0 (2/sqrt(pi)) [exp(-tb^2) - exp(-ta^2)] / (sqrt(2) l)
//   MOG:       c = c0 sum_q w_q g_q                          c' = -c0 sum_q w_q g_q (x-m_q)/S_q
//   empirical: c = c0 sum_q w_q exp(-t_q^2)                  c' = -c0 sum_q w_q exp(-t_q^2) 2 t_q / (sqrt(2) l)
// with c^ = c / sqrt(var_s) (oak_prepare.cu).
__global__ void __launch_bounds__(256) rows_finish_kernel(const double* __restrict__ zpartial, int segs, int64_t zrows,
                                                          int64_t n, int D, int Dc, const DimDev* __restrict__ dims,
                                                          const BwDim* __restrict__ bwdims,
                                                          const double2* __restrict__ pts, int64_t n_pad,
                                                          const double* __restrict__ inv_sqrt_v,
                                                          double* __restrict__ grad_rows, int64_t ldg) {
  const int64_t gw = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= n * Dc) return;  // whole warps leave together
  const int64_t m = gw / Dc;
  const int k = (int)(gw - m * Dc);
  double A = 0.0, B = 0.0;
  for (int s = lane; s < segs; s += 32) {
    const double* z = zpartial + (((int64_t)s * zrows + m) * D + k) * 2;
    A += z[0];
    B += z[1];
  }
  const DimDev d = dims[k];
  const double2 p = pts[(int64_t)k * n_pad + m];
  const double x = p.x / d.xscale;
  double dc = 0.0;  // per-lane share of c'(x)
  switch (d.measure) {
    case OAK_MEASURE_GAUSSIAN:
      if (lane == 0) dc = -2.0 * (x - d.c1) * d.c2 * (p.y / inv_sqrt_v[k]);
      break;
    case OAK_MEASURE_UNIFORM:
      if (lane == 0) {
        const double ta = (d.c1 - x) * d.inv_sqrt2_l, tb = (d.c2 - x) * d.inv_sqrt2_l;
        dc = -d.c0 * 1.1283791670955126 * (exp(-tb * tb) - exp(-ta * ta)) * d.inv_sqrt2_l;
      }
      break;
    case OAK_MEASURE_MOG: {
      const double l2 = d.lengthscale * d.lengthscale;
      for (int q = lane; q < d.count; q += 32) {
        const double sc = l2 + d.v1[q];
        const double t = x - d.v0[q];
        dc -= exp(-0.5 * (t * t) / sc) / sqrt(sc) * d.v2[q] * (t / sc);
      }
      dc *= d.c0;
      break;
    }
    case OAK_MEASURE_EMPIRICAL:
      for (int q = lane; q < d.count; q += 32) {
        const double t = (x - d.v0[q]) * d.inv_sqrt2_l;
        dc = fma(-2.0 * t * d.v1[q], exp(-t * t), dc);
      }
      dc *= d.c0 * d.inv_sqrt2_l;
      break;
    default:
      break;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    A += __shfl_xor_sync(0xffffffffu, A, o);
    B += __shfl_xor_sync(0xffffffffu, B, o);
    dc += __shfl_xor_sync(0xffffffffu, dc, o);
  }
  if (lane == 0) grad_rows[m * ldg + d.orig] += -A * bwdims[k].zc - dc * inv_sqrt_v[k] * B;
}

// ---- host side ------------------------------------------------------------------------------
static int build_bwdims(const oak_spec* spec, std::vector<BwDim>& out, std::vector<int>& orig_of_pos) {
  out.assign(spec->D, BwDim{0, 0, 0, 0, 0, 0, 0, 0});
  orig_of_pos.assign(spec->D, 0);
  for (int k = 0; k < spec->D; ++k) {
    const DimDev& dd = spec->h_dims[k];
    orig_of_pos[k] = dd.orig;
    if (dd.type != OAK_DIM_RBF) continue;
    const double l = dd.lengthscale;
    BwDim b;
    b.inv_xscale = 1.0 / dd.xscale;
    b.c2 = (2.0 / l) * kInvXScale2;
    b.inv_s2 = 1.0 / dd.s2;
    b.zc = 1.0 / (dd.xscale * l * l);
    b.mu = 0.0;
    b.half_kappa = 0.0;
    b.uc = 0.0;
    b.kind = 0.0;
    if (dd.measure == OAK_MEASURE_NONE) {
      b.kind = 1.0;  // c^ = 0: only the ex z 2/l term survives
    } else if (dd.measure == OAK_MEASURE_GAUSSIAN) {
      const double l2d = 0.5 / dd.c2;      // l^2 + delta^2
      const double delta2 = l2d - l * l;
      b.mu = dd.c1;
      b.half_kappa = 0.5 * (1.0 / l + l / (l * l + 2.0 * delta2) - 2.0 * l / l2d);
      b.uc = l / (l2d * l2d);
      b.kind = 1.0;
    } else {
      b.kind = 2.0;  // empirical / uniform / MOG: per-point block from oak_prepare_backward_f64
    }
    out[k] = b;
  }
  return 0;
}

// tile shape of the backward kernel for a depth (the launch table of oak_gram_backward_f64)
static void bw_tile_shape(int depth, int* tm, int* tn) {
  *tm = depth <= 3 ? 64 : 32;
  *tn = depth <= 4 ? 64 : 32;
}

// column segments per row block of the row-gradient variant: about four work units per SM
static int bw_row_segments(int64_t rows, int64_t n2, int depth, int sms) {
  int tm, tn;
  bw_tile_shape(depth, &tm, &tn);
  const int64_t row_blocks = (rows + tm - 1) / tm, tiles_n = (n2 + tn - 1) / tn;
  int64_t segs = (4 * (int64_t)sms + row_blocks - 1) / row_blocks;
  if (segs > tiles_n) segs = tiles_n;
  return (int)(segs < 1 ? 1 : segs);
}

template <int P, int RM, int RN, bool ZG>
static int launch_backward_as(BwParams prm, int grid_max, cudaStream_t stream, int* grid_out) {
  using namespace bw;
  constexpr int TM = kTYD * RM, TN = kTXD * RN;
  const int64_t rows = prm.row_end - prm.row_begin;
  prm.tiles_n = (prm.n2 + TN - 1) / TN;
  const int64_t row_blocks = (rows + TM - 1) / TM;
  prm.num_tiles = row_blocks * prm.tiles_n;
  size_t smem = sizeof(double) * kExpTab * kExpRepl + 2 * sizeof(double2) * kDimChunk * (TM + TN) +
                2 * sizeof(double) * kDimChunk + 8 * sizeof(double) * (2 * prm.D + P + 1 + prm.tables_len);
  int64_t units = prm.num_tiles;
  if (ZG) {
    smem += sizeof(double) * TM * prm.D * 2;
    prm.zrows = row_blocks * TM;
    units = row_blocks * prm.segs;
  }
  OAK_REQUIRE(smem <= 227 * 1024, "backward tiles: too many dimensions / table entries for one CTA's shared memory");
  auto kern = gram_backward_kernel<P, RM, RN, ZG>;
  OAK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(units < grid_max ? units : grid_max);
  *grid_out = grid;
  kern<<<grid, kThreads, smem, stream>>>(prm);
  OAK_LAUNCHED();
  return 0;
}

template <int P, int RM, int RN>
static int launch_backward(const BwParams& prm, int grid_max, cudaStream_t stream, int* grid_out) {
  return prm.zpartial ? launch_backward_as<P, RM, RN, true>(prm, grid_max, stream, grid_out)
                      : launch_backward_as<P, RM, RN, false>(prm, grid_max, stream, grid_out);
}

}  // namespace oak

using namespace oak;

// workspace: per-dimension constants + dimension map + per-CTA partials
extern "C" size_t oak_gram_backward_work_bytes(const oak_spec* spec, int64_t n) {
  if (!spec || n < 0) return 0;
  const size_t nout = 2 * (size_t)spec->D + spec->depth + 2 + spec->tables_len;
  const size_t blocks = (size_t)((n + 255) / 256);
  const size_t ctas = blocks > 1024 ? blocks : 1024;
  return 256 + (size_t)spec->D * (sizeof(BwDim) + sizeof(int)) + 64 + ctas * nout * sizeof(double);
}

static int stage_bwdims(const oak_spec* spec, void* d_work, cudaStream_t stream, const BwDim** d_bw,
                        const int** d_map, double** d_partial) {
  std::vector<BwDim> h;
  std::vector<int> map;
  build_bwdims(spec, h, map);
  char* w = (char*)d_work;
  BwDim* dbw = (BwDim*)w;
  w += (size_t)spec->D * sizeof(BwDim);
  int* dmap = (int*)w;
  w += ((size_t)spec->D * sizeof(int) + 63) / 64 * 64;
  // pageable sources: the copies are staged before the calls return
  OAK_CUDA(cudaMemcpyAsync(dbw, h.data(), h.size() * sizeof(BwDim), cudaMemcpyHostToDevice, stream));
  OAK_CUDA(cudaMemcpyAsync(dmap, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  *d_bw = dbw;
  *d_map = dmap;
  *d_partial = (double*)w;
  return 0;
}

// Layout of the gradient vector: [num_dims lengthscales (caller's order) | depth + 1 order
// variances | cotangent of the discrete-kernel table blob | num_dims base variances s^2 of the RBF
// sub-kernels (caller's order; zero for discrete sub-kernels)].
extern "C" size_t oak_backward_grad_count(const oak_spec* spec) {
  if (!spec) return 0;
  return 2 * (size_t)spec->D + (size_t)(spec->depth < 1 ? 1 : spec->depth) + 1 + (size_t)spec->tables_len;
}

// Where sub-kernel `dim` (caller's order) keeps its table inside that blob: C x C entries B[a, b]
// (row-major) followed by the C diagonal entries read by K_diag.  count = 0 for RBF sub-kernels.
extern "C" int oak_spec_table_layout(const oak_spec* spec, int32_t dim, int32_t* offset, int32_t* count) {
  OAK_REQUIRE(spec && offset && count, "oak_spec_table_layout: null argument");
  OAK_REQUIRE(dim >= 0 && dim < spec->D, "oak_spec_table_layout: dim out of range");
  const DimDev& dd = spec->h_dims[spec->pos_of_orig[dim]];
  if (dd.type == OAK_DIM_RBF) {
    *offset = 0;
    *count = 0;
  } else {
    *offset = dd.table_off;
    *count = dd.count;
  }
  return 0;
}

// per-point derivative block: double dch[D][n_pad] (d c^/dl; written for empirical-measure dims only)
extern "C" size_t oak_backward_points_bytes(const oak_spec* spec, int64_t n) {
  if (!spec || n < 0) return 0;
  return (size_t)spec->D * (size_t)padded(n) * sizeof(double);
}

extern "C" int oak_prepare_backward_f64(const oak_spec* spec, const void* d_points, int64_t n, void* d_dpoints,
                                        void* stream_) {
  OAK_REQUIRE(spec && d_points && d_dpoints, "oak_prepare_backward_f64: null argument");
  if (n <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t n_pad = padded(n);
  for (int k = 0; k < spec->Dc; ++k) {
    const DimDev& dd = spec->h_dims[k];
    if (dd.measure == OAK_MEASURE_UNIFORM || dd.measure == OAK_MEASURE_MOG) {
      // v'/(2v) on the host: closed forms of var_s differentiated
      const double l = dd.lengthscale;
      double v = 0.0, dv = 0.0;
      if (dd.measure == OAK_MEASURE_UNIFORM) {
        const double delta = dd.c2 - dd.c1, y = delta / std::sqrt(2.0) / l;
        const double sp = std::sqrt(M_PI);
        v = 2.0 / (delta * delta) * dd.s2 * l * l * (sp * y * std::erf(y) + std::exp(-y * y) - 1.0);
        dv = 2.0 / (delta * delta) * dd.s2 * l * (sp * y * std::erf(y) + 2.0 * std::exp(-y * y) - 2.0);
      } else {
        const double* m = spec->h_blob.data() + spec->blob_off0[k];
        const double* var = spec->h_blob.data() + spec->blob_off1[k];
        const double* w = spec->h_blob.data() + spec->blob_off2[k];
        for (int i = 0; i < dd.count; ++i)
          for (int j = 0; j < dd.count; ++j) {
            const double T = l * l + var[i] + var[j];
            const double dist = (m[i] - m[j]) * (m[i] - m[j]);
            const double h = w[i] * w[j] * std::exp(-0.5 * dist / T) / std::sqrt(T);
            v += dd.s2 * l * h;
            dv += dd.s2 * h + dd.s2 * l * h * (dist * l / (T * T) - l / T);
          }
      }
      measure_dch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
          dd, spec->d_inv_sqrt_v + k, 0.5 * dv / v, (const double2*)d_points + (int64_t)k * n_pad, n,
          (double*)d_dpoints + (int64_t)k * n_pad);
      OAK_LAUNCHED();
      continue;
    }
    if (dd.measure != OAK_MEASURE_EMPIRICAL) continue;
    const int threads = 256;
    const int blocks = (dd.count + threads - 1) / threads;
    double* partial = nullptr;
    OAK_CUDA(cudaMallocAsync(&partial, blocks * sizeof(double), stream));
    const double two_over_l = 2.0 / dd.lengthscale;
    empirical_dvar_partial<<<blocks, threads, 2 * threads * sizeof(double), stream>>>(
        dd.v0, dd.v1, dd.count, dd.inv_sqrt2_l, dd.s2 * two_over_l, partial);
    OAK_LAUNCHED();
    empirical_dch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        dd.v0, dd.v1, dd.count, dd.inv_sqrt2_l, dd.s2, two_over_l, spec->d_inv_sqrt_v + k, partial, blocks,
        (const double2*)d_points + (int64_t)k * n_pad, n, (double*)d_dpoints + (int64_t)k * n_pad);
    OAK_LAUNCHED();
    OAK_CUDA(cudaFreeAsync(partial, stream));
  }
  return 0;
}

// grad[0..D) += d/d lengthscale of sub-kernel i (caller's order; unsupported ones untouched),
// grad[D..D+depth] += d/d sigma2_n, for rows [row_begin,row_end) of points x all points2
// (points2 == NULL: the same point set), cotangent W (rows x n2, pitch ldw).  d_dpoints /
// d_dpoints2: the blocks written by oak_prepare_backward_f64 (needed when the kernel has
// empirical-measure dims; may be NULL otherwise).
static int gram_backward_impl(const oak_spec* spec, const void* d_points, const void* d_dpoints, int64_t n,
                              int64_t row_begin, int64_t row_end, const void* d_points2, const void* d_dpoints2,
                              int64_t n2, const double* d_W, int64_t ldw, double* d_grad, double* d_grad_rows,
                              int64_t ldg, void* d_work, void* stream_) {
  OAK_REQUIRE(spec && d_points && d_grad && d_work, "oak_gram_backward_f64: null argument");
  const bool same = d_points2 == nullptr;
  if (same) n2 = n;
  OAK_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n, "oak_gram_backward_f64: bad row range");
  if (row_end == row_begin || n2 <= 0) return 0;
  OAK_REQUIRE(d_W && ldw >= n2, "oak_gram_backward_f64: bad cotangent");
  OAK_REQUIRE(row_begin % 64 == 0, "oak_gram_backward_f64: row_begin must be a multiple of 64");
  OAK_REQUIRE(spec->D <= bw::kMaxDims, "oak_gram_backward_f64: too many dimensions");
  const int depth = spec->depth < 1 ? 1 : spec->depth;
  OAK_REQUIRE(depth <= 8, "oak_gram_backward_f64: max_interaction_depth > 8 is not supported by the backward tiles");
  cudaStream_t stream = (cudaStream_t)stream_;
  const BwDim* d_bw;
  const int* d_map;
  double* d_partial;
  if (int rc = stage_bwdims(spec, d_work, stream, &d_bw, &d_map, &d_partial)) return rc;
  BwParams prm;
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) prm.sigma2[p] = spec->sigma2[p];
  prm.pts_row = (const double2*)d_points;
  prm.pts_col = same ? (const double2*)d_points : (const double2*)d_points2;
  prm.n_row_pad = padded(n);
  prm.n_col_pad = padded(n2);
  prm.dim_aux = spec->d_gram_aux;
  prm.tables = spec->d_tables;
  prm.exptab = spec->d_exptab;
  prm.bwdims = d_bw;
  prm.dch_row = (const double*)d_dpoints;
  prm.dch_col = same ? (const double*)d_dpoints : (const double*)d_dpoints2;
  bool need_dch = false;
  for (int k = 0; k < spec->Dc; ++k) {
    const int ms = spec->h_dims[k].measure;
    need_dch |= ms == OAK_MEASURE_EMPIRICAL || ms == OAK_MEASURE_UNIFORM || ms == OAK_MEASURE_MOG;
  }
  OAK_REQUIRE(!need_dch || (prm.dch_row && prm.dch_col),
              "oak_gram_backward_f64: empirical / uniform / MOG dims need the oak_prepare_backward_f64 blocks");
  prm.mm_row = points_minmax(spec, prm.pts_row, prm.n_row_pad);
  prm.mm_col = points_minmax(spec, prm.pts_col, prm.n_col_pad);
  prm.W = d_W;
  prm.ldw = ldw;
  prm.partial = d_partial;
  prm.row_begin = row_begin;
  prm.row_end = row_end;
  prm.n2 = n2;
  prm.D = spec->D;
  prm.Dc = spec->Dc;
  prm.tables_len = spec->tables_len;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, spec->device);
  prm.zpartial = nullptr;
  prm.zrows = 0;
  prm.segs = 0;
  if (d_grad_rows) {  // (A, B) partials behind the per-CTA parameter partials
    const size_t nout_ = 2 * (size_t)spec->D + depth + 1 + spec->tables_len;
    prm.zpartial = d_partial + ((size_t)sms * nout_ + 7) / 8 * 8;
    prm.segs = bw_row_segments(row_end - row_begin, n2, depth, sms);
  }
  int grid = 0, rc = 0;
  switch (depth) {
    case 1: rc = launch_backward<1, OAK_BW_RM3, OAK_BW_RN3>(prm, sms * OAK_BW_MINB, stream, &grid); break;
    case 2: rc = launch_backward<2, OAK_BW_RM3, OAK_BW_RN3>(prm, sms * OAK_BW_MINB, stream, &grid); break;
    case 3: rc = launch_backward<3, OAK_BW_RM3, OAK_BW_RN3>(prm, sms * OAK_BW_MINB, stream, &grid); break;
    case 4: rc = launch_backward<4, 2, 4>(prm, sms, stream, &grid); break;
    case 5: rc = launch_backward<5, 2, 2>(prm, sms, stream, &grid); break;
    case 6: rc = launch_backward<6, 2, 2>(prm, sms, stream, &grid); break;
    case 7: rc = launch_backward<7, 2, 2>(prm, sms, stream, &grid); break;
    default: rc = launch_backward<8, 2, 2>(prm, sms, stream, &grid); break;
  }
  if (rc) return rc;
  const int nout = 2 * spec->D + depth + 1 + spec->tables_len;
  backward_reduce_kernel<<<(nout + 127) / 128, 128, 0, stream>>>(d_partial, grid, nout, spec->D, d_map, d_grad);
  OAK_LAUNCHED();
  if (d_grad_rows && spec->Dc > 0) {
    int tm, tn;
    bw_tile_shape(depth, &tm, &tn);
    const int64_t rows = row_end - row_begin;
    const int64_t zrows = (rows + tm - 1) / tm * tm;
    const int64_t warps = rows * spec->Dc;
    rows_finish_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(
        prm.zpartial, prm.segs, zrows, rows, spec->D, spec->Dc, spec->d_dims, d_bw, prm.pts_row, prm.n_row_pad,
        spec->d_inv_sqrt_v, d_grad_rows, ldg);
    OAK_LAUNCHED();
  }
  return 0;
}

extern "C" int oak_gram_backward_f64(const oak_spec* spec, const void* d_points, const void* d_dpoints, int64_t n,
                                     int64_t row_begin, int64_t row_end, const void* d_points2,
                                     const void* d_dpoints2, int64_t n2, const double* d_W, int64_t ldw,
                                     double* d_grad, void* d_work, void* stream_) {
  return gram_backward_impl(spec, d_points, d_dpoints, n, row_begin, row_end, d_points2, d_dpoints2, n2, d_W, ldw,
                            d_grad, nullptr, 0, d_work, stream_);
}

// Same contraction over all rows of `points`, and additionally the gradient with respect to the row
// points themselves (the inducing points Z of SGPR with zfixed=False, model_utils.py:98-101):
//   grad_rows[i * ldg + k] += sum_j W_ij dK(x_i, y_j)/d x_{i,k}     (k = sub-kernel index, caller's order;
// discrete sub-kernels receive nothing: tf.cast / tf.gather carry no gradient).  Only the FIRST argument
// of K is differentiated: for a symmetric objective over K(Z, Z) pass W + W^T.
extern "C" size_t oak_gram_backward_rows_work_bytes(const oak_spec* spec, int64_t n, int64_t n2) {
  if (!spec || n < 0 || n2 < 0) return 0;
  const int depth = spec->depth < 1 ? 1 : spec->depth;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, spec->device);
  int tm, tn;
  bw_tile_shape(depth, &tm, &tn);
  const size_t zrows = (size_t)((n + tm - 1) / tm * tm);
  const size_t segs = (size_t)bw_row_segments(n, n2 > 0 ? n2 : n, depth, sms);
  return oak_gram_backward_work_bytes(spec, n > n2 ? n : n2) + 64 + segs * zrows * 2 * (size_t)spec->D * sizeof(double);
}

extern "C" int oak_gram_backward_rows_f64(const oak_spec* spec, const void* d_points, const void* d_dpoints,
                                          int64_t n, const void* d_points2, const void* d_dpoints2, int64_t n2,
                                          const double* d_W, int64_t ldw, double* d_grad, double* d_grad_rows,
                                          int64_t ldg, void* d_work, void* stream_) {
  OAK_REQUIRE(d_grad_rows && spec && ldg >= spec->D, "oak_gram_backward_rows_f64: bad row-gradient buffer");
  return gram_backward_impl(spec, d_points, d_dpoints, n, 0, n, d_points2, d_dpoints2, n2, d_W, ldw, d_grad,
                            d_grad_rows, ldg, d_work, stream_);
}

// grad += d/d theta of  wscale * sum_i w_i K_diag(x_i)   (w == NULL: all ones)
extern "C" int oak_gram_diag_backward_f64(const oak_spec* spec, const void* d_points, const void* d_dpoints,
                                          int64_t n, const double* d_w, double wscale, double* d_grad,
                                          void* d_work, void* stream_) {
  OAK_REQUIRE(spec && d_points && d_grad && d_work, "oak_gram_diag_backward_f64: null argument");
  if (n <= 0) return 0;
  for (int k = 0; k < spec->Dc; ++k) {
    const int ms = spec->h_dims[k].measure;
    OAK_REQUIRE(!(ms == OAK_MEASURE_EMPIRICAL || ms == OAK_MEASURE_UNIFORM || ms == OAK_MEASURE_MOG) || d_dpoints,
                "oak_gram_diag_backward_f64: empirical / uniform / MOG dims need the oak_prepare_backward_f64 block");
  }
  const int depth = spec->depth < 1 ? 1 : spec->depth;
  cudaStream_t stream = (cudaStream_t)stream_;
  const BwDim* d_bw;
  const int* d_map;
  double* d_partial;
  if (int rc = stage_bwdims(spec, d_work, stream, &d_bw, &d_map, &d_partial)) return rc;
  DiagBwParams prm;
  for (int p = 0; p <= OAK_MAX_DEPTH; ++p) prm.sigma2[p] = spec->sigma2[p];
  prm.D = spec->D;
  prm.Dc = spec->Dc;
  prm.depth = depth;
  prm.tables_len = spec->tables_len;
  const int nout = 2 * spec->D + depth + 1 + spec->tables_len;
  const int blocks = (int)((n + 255) / 256);
  diag_backward_kernel<<<blocks, 256, 8 * nout * sizeof(double), stream>>>(
      prm, spec->d_dims, d_bw, (const double2*)d_points, (const double*)d_dpoints, n, padded(n), d_w, wscale,
      d_partial);
  OAK_LAUNCHED();
  backward_reduce_kernel<<<(nout + 127) / 128, 128, 0, stream>>>(d_partial, blocks, nout, spec->D, d_map, d_grad);
  OAK_LAUNCHED();
  return 0;
}
