// Blocked Cholesky factorisation with border rows in ONE cooperative launch (sm_100a).
//
// The M x M factorisations of the SGPR bound (L = chol(Kuu + jitter I), LB = chol(I + A A^T),
// oak/utils.py:188-193) sit on the critical path of every ELBO evaluation and are replicated on every
// rank; cuSOLVER's potrf needs ~1 ms at M = 1024 (a dozen launches with two ~450 us panel kernels).
// Here the whole factorisation is a single persistent kernel:
//
//   for each 64-column panel
//     phase 1  every CTA that owns rows of the panel factors the 64 x 64 diagonal block redundantly in
//              registers (right-looking on the bordered block [S ; I]: one barrier per pivot, L11 and
//              inv(L11) come out together) and solves its 16-row tiles of the panel as a small matrix
//              product X = A21 inv(L11)^T; CTA 0 writes L11 back
//     grid barrier
//     phase 2  trailing update C -= X X^T on the 64 x 64 tiles of the lower triangle (4 x 4 register
//              micro-tiles, FP64 FMA), one tile per CTA at M = 1024
//     grid barrier
//
// Border rows: rows [n, rows) below the symmetric block take part in the panel solves and in the
// trailing updates but are never factored, so on exit they hold  Border * L^-T.  Two uses:
//   * border = identity  ->  L^-T, i.e. the explicit inverse factor (whitened SGPR statistics, the
//     condition estimate, the un-whitened tail); rows whose unit entry lies right of the panel are
//     still zero there and are skipped, which halves the work;
//   * border = one row v^T  ->  (L^-1 v)^T, the forward substitution of the bound's `c` for free.
// The sum of log diag(L) is accumulated by CTA 0 in pivot order (deterministic).
// Column-major storage (leading dimension ld); only the lower triangle of the symmetric block is read
// and written.  `gap` unused rows may separate the symmetric block from the border (alignment).
#include <cooperative_groups.h>

#include <type_traits>

#include "oak_common.cuh"

namespace cg = cooperative_groups;

namespace oak {

namespace chol {
constexpr int NB = 64;              // panel width
constexpr int LD = NB + 1;          // shared-memory leading dimension (conflict-free rows and columns)
constexpr int kThreads = 256;
constexpr int TR = 16;              // rows per panel-solve tile
constexpr int kBlk = NB * LD;       // doubles per staged 64 x 64 block
constexpr size_t kSmemBytes = (3 * (size_t)kBlk + 2 * NB) * sizeof(double);  // S | Lc | Inv | rd | lg
}  // namespace chol

struct CholParams {
  double* A;
  int64_t ld;
  int n, rows, gap;        // rows = n + number of border rows; border row r >= n lives at row r + gap
  int border_identity;     // border row n + i holds e_i on entry
  int* info;               // 0, or k > 0: the leading minor of order k is not positive definite
  double* logdet;          // optional: sum_i log L_ii
};

#ifdef OAK_CHOL_TIMING  // development aid: per-panel phase boundaries (cycles) of CTA 0
__device__ long long g_chol_t[64 * 8];
__device__ long long g_chol_t2[256 * 4];  // per CTA, panel 1: phase-1 cycles, phase-2 cycles, tiles
#define OAK_CHOL_T(slot)                                                                      \
  do {                                                                                        \
    if (blockIdx.x == 0 && tid == 0 && j0 / NB < 64) g_chol_t[(j0 / NB) * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define OAK_CHOL_T(slot) \
  do {                   \
  } while (0)
#endif

__global__ void __launch_bounds__(chol::kThreads, 1) chol_bordered_kernel(const CholParams prm) {
  using namespace chol;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double sm[];
  double* S = sm;
  double* Lc = sm + kBlk;
  double* Inv = sm + 2 * kBlk;
  double* rd = sm + 3 * kBlk;
  double* lg = rd + NB;
  const int tid = threadIdx.x;
  const int n = prm.n;
  const int64_t ld = prm.ld;
  double* const A = prm.A;
  auto grow = [&](int r) -> int64_t { return r < n ? r : (int64_t)r + prm.gap; };
  double logsum = 0.0;
  if (blockIdx.x == 0 && tid == 0) *prm.info = 0;

  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = min(NB, n - j0);
    const int j1 = j0 + nb;
    const int r_end = prm.border_identity ? min(prm.rows, n + j1) : prm.rows;
    const int trsm_tiles = (r_end - j1 + TR - 1) / TR;
    OAK_CHOL_T(0);
    if (blockIdx.x == 0 || (int)blockIdx.x < trsm_tiles) {
      // ---- diagonal block: L11 and inv(L11) together, trailing matrix in registers -----------------
      // Thread (i, kq) = (tid & 63, tid >> 6) owns row i, columns k = kq + 4u (u = 0..15) of the UNIFIED
      // 64 x 64 matrix W: W(i, k) = S(i, k) for k <= i (the symmetric block being factored) and the border
      // E(i, k) for k > i, E = I on entry (its unit diagonal is implicit).  Factoring [S ; E] right-looking
      // leaves L11 below the diagonal and E L11^-T = inv(L11)^T above it, so the triangular inverse costs no
      // separate pass.  Per pivot j: the owners of column j publish it through a double-buffered shared
      // vector, ONE barrier, every thread forms its multiplier m_i and updates its registers:
      //   W(i, k) -= m_i l_kj   for k > j and (k <= i  [S part, i > j]  or  i <= j  [E part]).
      const int i = tid & (NB - 1), kq = tid >> 6;
      double w[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int k = kq + 4 * u;
        double v = 0.0;
        if (k <= i) {
          v = (i == k) ? 1.0 : 0.0;  // padding: identity
          if (i < nb) v = A[(int64_t)(j0 + k) * ld + j0 + i];
        }
        w[u] = v;
      }
      OAK_CHOL_T(1);
      int failj = -1;
      double* const Cs = rd;  // 2 x NB doubles (rd | lg): the published column, double buffered
      // Groups of four pivots; the register slots ROTATE so that the group's own columns always sit in slot 0
      // and the loop body is the same code for every group (a fully unrolled 64-pivot body is ~100 KB of
      // straight-line code executed once per panel: instruction-fetch bound, measured 22 us against 6).
      // Two copies of the body: 16 live slots for the first eight groups, 8 for the rest.
      auto pivot_group = [&](const int g, auto slots_tag) -> bool {
        constexpr int kSlots = decltype(slots_tag)::value;
#pragma unroll
        for (int jq = 0; jq < 4; ++jq) {
          const int j = 4 * g + jq;
          double* const C = Cs + (jq & 1) * NB;
          if (kq == jq) C[i] = w[0];
          __syncthreads();
          const double ajj = C[j];
          if (!(ajj > 0.0) || !(ajj < 1.0e300)) {  // uniform: every thread reads the same value
            failj = j;
            return false;
          }
          const double rs = rsqrt(ajj);
          const double mi = (i == j) ? rs : C[i] * rs;
          if (kq == jq) {
            if (i > j) {
              Lc[j * LD + i] = mi;            // L(i, j)
            } else if (i == j) {
              Lc[j * LD + j] = ajj * rs;      // sqrt(a_jj) without a second long-latency chain in this warp
              Inv[j * LD + j] = rs;
            } else {
              Inv[i * LD + j] = mi;           // inv(L11)(j, i) = (L11^-T)(i, j)
            }
          }
          const double mrs = -mi * rs;
          const int klim = (i <= j) ? NB - 1 : i;
#pragma unroll
          for (int u = 0; u < kSlots; ++u) {
            const int k = kq + 4 * (g + u);
            if (k > j && k <= klim) w[u] = fma(mrs, C[k], w[u]);
          }
        }
#pragma unroll
        for (int u = 0; u + 1 < 16; ++u) w[u] = w[u + 1];
        w[15] = 0.0;
        return true;
      };
      {
        bool ok = true;
#pragma unroll 1
        for (int g = 0; g < 8 && ok; ++g) ok = pivot_group(g, std::integral_constant<int, 16>{});
#pragma unroll 1
        for (int g = 8; g < 16 && ok; ++g) ok = pivot_group(g, std::integral_constant<int, 8>{});
      }
      __syncthreads();
      OAK_CHOL_T(2);
      if (failj >= 0) {
        if (blockIdx.x == 0 && tid == 0) *prm.info = j0 + failj + 1;
      } else {
        if (blockIdx.x == 0) {
          for (int e = tid; e < NB * NB; e += kThreads) {
            const int r = e & (NB - 1), k = e >> 6;
            if (r < nb && k < nb && r >= k) A[(int64_t)(j0 + k) * ld + j0 + r] = Lc[k * LD + r];
          }
          if (tid < NB) S[tid] = log(Lc[tid * LD + tid]);  // padded pivots are 1 -> 0
          __syncthreads();
          if (tid == 0) {
            for (int r = 0; r < nb; ++r) logsum += S[r];
          }
          __syncthreads();
        }
        OAK_CHOL_T(3);
        // ---- panel solve: X = A21 inv(L11)^T on this CTA's 16-row tiles ----------------------
        for (int t = blockIdx.x; t < trsm_tiles; t += gridDim.x) {
          const int r0 = j1 + t * TR;
          for (int e = tid; e < NB * TR; e += kThreads) {
            const int r = e & (TR - 1), k = e >> 4;
            const int row = r0 + r;
            S[k * LD + r] = (row < r_end && k < nb) ? A[(int64_t)(j0 + k) * ld + grow(row)] : 0.0;
          }
          __syncthreads();
          const int c = tid & (NB - 1), rq = tid >> 6;
          double x[4] = {0.0, 0.0, 0.0, 0.0};
          for (int k = 0; k <= c; ++k) {
            const double iv = Inv[k * LD + c];  // inv(L11)(c, k)
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = fma(S[k * LD + rq + 4 * u], iv, x[u]);
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < 4; ++u) S[c * LD + rq + 4 * u] = x[u];
          __syncthreads();
          for (int e = tid; e < NB * TR; e += kThreads) {
            const int r = e & (TR - 1), k = e >> 4;
            const int row = r0 + r;
            if (row < r_end && k < nb) A[(int64_t)(j0 + k) * ld + grow(row)] = S[k * LD + r];
          }
          __syncthreads();
        }
      }
    }
    OAK_CHOL_T(4);
#ifdef OAK_CHOL_TIMING
    long long tb_ = clock64();
    if (j0 == NB && tid == 0 && blockIdx.x < 256) g_chol_t2[blockIdx.x * 4 + 0] = tb_ - g_chol_t2[blockIdx.x * 4 + 3];
#endif
    __threadfence();
    grid.sync();
    OAK_CHOL_T(5);
#ifdef OAK_CHOL_TIMING
    tb_ = clock64();
#endif
    if (*(volatile int*)prm.info != 0) break;  // uniform: written before the barrier
    if (j1 >= n) break;
    // ---- trailing update: C -= X X^T on the tiles that touch the lower triangle / the border ---
    // FP64 tensor cores (DMMA m8n8k4): 8 warps as 4 x 2, warp tile 16 x 32; both operands are panel rows staged
    // k-major with stride 68 (conflict-free fragment loads: lane (g, q) reads [4 ks + q][.. + g]).
    {
      constexpr int LX = NB + 4;
      double* const Xa = sm;
      double* const Xb = sm + NB * LX;
      const int rt_count = (r_end - j1 + NB - 1) / NB, ct_count = (n - j1 + NB - 1) / NB;
      const int total = ct_count * rt_count - ct_count * (ct_count - 1) / 2;
      const int lane = tid & 31, warp = tid >> 5;
      const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
      for (int u = blockIdx.x; u < total; u += gridDim.x) {
        int ct = 0, rem = u;
        while (rem >= rt_count - ct) {
          rem -= rt_count - ct;
          ++ct;
        }
        const int rbase = j1 + NB * (ct + rem), cbase = j1 + NB * ct;
        __syncthreads();
        for (int e = tid; e < NB * NB; e += kThreads) {
          const int r = e & (NB - 1), k = e >> 6;
          const bool kk = k < nb;
          Xa[k * LX + r] = (kk && rbase + r < r_end) ? A[(int64_t)(j0 + k) * ld + grow(rbase + r)] : 0.0;
          Xb[k * LX + r] = (kk && cbase + r < n) ? A[(int64_t)(j0 + k) * ld + cbase + r] : 0.0;
        }
        __syncthreads();
        double acc[2][4][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        const double* pa = Xa + q * LX + wm * 16 + g;
        const double* pb = Xb + q * LX + wn * 32 + g;
#pragma unroll 4
        for (int ks = 0; ks < NB / 4; ++ks) {
          double av[2], bv[4];
#pragma unroll
          for (int a = 0; a < 2; ++a) av[a] = pa[ks * 4 * LX + a * 8];
#pragma unroll
          for (int b = 0; b < 4; ++b) bv[b] = pb[ks * 4 * LX + b * 8];
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                           : "+d"(acc[a][b][0]), "+d"(acc[a][b][1])
                           : "d"(av[a]), "d"(bv[b]));
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int r = rbase + wm * 16 + a * 8 + g;
          if (r >= r_end) continue;
          const int64_t gr = grow(r);
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = cbase + wn * 32 + b * 8 + 2 * q + e;
              if (c < n && r >= c) A[(int64_t)c * ld + gr] -= acc[a][b][e];
            }
        }
      }
    }
    OAK_CHOL_T(6);
#ifdef OAK_CHOL_TIMING
    if (j0 == NB && tid == 0 && blockIdx.x < 256) g_chol_t2[blockIdx.x * 4 + 1] = clock64() - tb_;
#endif
    __threadfence();
    grid.sync();
    OAK_CHOL_T(7);
#ifdef OAK_CHOL_TIMING
    if (j0 == 0 && tid == 0 && blockIdx.x < 256) g_chol_t2[blockIdx.x * 4 + 3] = clock64();  // start of panel 1
#endif
  }
  if (blockIdx.x == 0 && tid == 0 && prm.logdet) *prm.logdet = logsum;
}

// Launcher shared with oak_sgpr.cu.  Column-major A (ld), symmetric block n x n (lower triangle), border
// rows [n, rows) stored `gap` rows further down.
int chol_bordered(double* A, int64_t ld, int n, int rows, int gap, int border_identity, int* d_info,
                  double* d_logdet, int device, cudaStream_t stream) {
  using namespace chol;
  if (n <= 0) return 0;
  OAK_REQUIRE(rows >= n && gap >= 0 && ld >= (int64_t)rows + gap, "chol_bordered: bad shape");
  static int sms_cached[64] = {0};
  int sms = (device >= 0 && device < 64) ? sms_cached[device] : 0;
  if (sms == 0) {
    int coop = 0;
    OAK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    OAK_REQUIRE(coop, "chol_bordered: the device does not support cooperative launches");
    OAK_CUDA(cudaFuncSetAttribute(chol_bordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSmemBytes));
    int per_sm = 0;
    OAK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_bordered_kernel, kThreads, kSmemBytes));
    OAK_REQUIRE(per_sm >= 1, "chol_bordered: kernel does not fit on an SM");
    OAK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (device >= 0 && device < 64) sms_cached[device] = sms;
  }
  // widest phase over the panels -> number of CTAs (never more than one per SM: all must be co-resident)
  int want = 1;
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int j1 = j0 + (n - j0 < NB ? n - j0 : NB);
    const int r_end = border_identity ? (rows < n + j1 ? rows : n + j1) : rows;
    const int t1 = (r_end - j1 + TR - 1) / TR;
    const int rt = (r_end - j1 + NB - 1) / NB, ct = (n - j1 + NB - 1) / NB;
    const int t2 = ct * rt - ct * (ct - 1) / 2;
    if (t1 > want) want = t1;
    if (t2 > want) want = t2;
  }
  const int grid = want < sms ? want : sms;
  CholParams prm{A, ld, n, rows, gap, border_identity, d_info, d_logdet};
  void* args[] = {&prm};
  OAK_CUDA(cudaLaunchCooperativeKernel((const void*)chol_bordered_kernel, dim3(grid), dim3(kThreads), args,
                                       kSmemBytes, stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace oak

using namespace oak;

#ifdef OAK_CHOL_TIMING
extern "C" int oak_debug_chol_timing(long long* h_out) {
  OAK_CUDA(cudaDeviceSynchronize());
  OAK_CUDA(cudaMemcpyFromSymbol(h_out, g_chol_t, sizeof(long long) * 64 * 8));
  OAK_CUDA(cudaMemcpyFromSymbol(h_out + 64 * 8, g_chol_t2, sizeof(long long) * 256 * 4));
  return 0;
}
#endif

// Cholesky factorisation with border rows of a column-major matrix (equivalently: the UPPER triangle of
// the row-major view).  Replaces the tf.linalg.cholesky + tf.linalg.triangular_solve pairs of
// oak/utils.py:188-195 (and of gpflow's SGPR.elbo / GPR.log_marginal_likelihood).
extern "C" int oak_chol_f64(double* d_A, int64_t n, int64_t rows, int64_t gap, int64_t ld, int border_identity,
                            int32_t* d_info, double* d_logdet, void* stream) {
  OAK_REQUIRE(d_A && d_info, "oak_chol_f64: null argument");
  OAK_REQUIRE(n >= 0 && n <= INT32_MAX / 4 && rows >= n && rows <= INT32_MAX / 4 && gap >= 0 && gap <= INT32_MAX / 4,
              "oak_chol_f64: bad size");
  int dev = 0;
  OAK_CUDA(cudaGetDevice(&dev));
  return chol_bordered(d_A, ld, (int)n, (int)rows, (int)gap, border_identity, (int*)d_info, d_logdet, dev,
                       (cudaStream_t)stream);
}
