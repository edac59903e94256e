// Blocked Cholesky factorisation with border rows in ONE cooperative launch (sm_100a).
//
// The M x M factorisations of the SGPR bound (L = chol(Kuu + jitter I), LB = chol(I + A A^T),
// oak/utils.py:188-193) sit on the critical path of every ELBO evaluation and are replicated on every
// rank; cuSOLVER's potrf needs ~1 ms at M = 1024 (a dozen launches with two ~450 us panel kernels).
// Here the whole factorisation is a single persistent kernel:
//
//   for each 64-column panel
//     phase 1  every CTA that owns rows of the panel factors the 64 x 64 diagonal block redundantly in
//              shared memory (right-looking, one barrier per pivot), inverts it (16 -> 32 -> 64 blocked
//              triangular inverse) and solves its 16-row tiles of the panel as a small matrix product
//              X = A21 inv(L11)^T; CTA 0 writes L11 back
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
    if (blockIdx.x == 0 || (int)blockIdx.x < trsm_tiles) {
      // ---- diagonal block -> shared memory (padded with the identity) ----------------------
      for (int e = tid; e < NB * NB; e += kThreads) {
        const int i = e & (NB - 1), k = e >> 6;
        double v = (i == k) ? 1.0 : 0.0;
        if (i < nb && k < nb && i >= k) v = A[(int64_t)(j0 + k) * ld + j0 + i];
        S[k * LD + i] = v;
      }
      __syncthreads();
      int failj = -1;
      {
        const int i = tid & (NB - 1);
        for (int j = 0; j < NB; ++j) {
          const double ajj = S[j * LD + j];
          if (!(ajj > 0.0) || !(ajj < 1.0e300)) {  // uniform: every thread reads the same value
            failj = j;
            break;
          }
          const double ljj = sqrt(ajj);
          const double rs = rsqrt(ajj);
          const double sij = S[j * LD + i];
          if (tid < NB) Lc[j * LD + i] = (i > j) ? sij * rs : (i == j ? ljj : 0.0);
          if (i > j) {
            const double lij = sij * rs;
            for (int k = j + 1 + (tid >> 6); k <= i; k += 4) {
              const double lkj = S[j * LD + k] * rs;
              S[k * LD + i] = fma(-lij, lkj, S[k * LD + i]);
            }
          }
          __syncthreads();
        }
      }
      if (failj >= 0) {
        if (blockIdx.x == 0 && tid == 0) *prm.info = j0 + failj + 1;
      } else {
        if (blockIdx.x == 0) {
          for (int e = tid; e < NB * NB; e += kThreads) {
            const int i = e & (NB - 1), k = e >> 6;
            if (i < nb && k < nb && i >= k) A[(int64_t)(j0 + k) * ld + j0 + i] = Lc[k * LD + i];
          }
          if (tid < NB) lg[tid] = log(Lc[tid * LD + tid]);  // padded pivots are 1 -> 0
        }
        // ---- inv(L11): 16 x 16 diagonal blocks by substitution, then 32, then 64 by products -
        if (tid < NB) rd[tid] = 1.0 / Lc[tid * LD + tid];
        for (int e = tid; e < kBlk; e += kThreads) Inv[e] = 0.0;
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) {
          for (int i = 0; i < nb; ++i) logsum += lg[i];
        }
        if (tid < NB) {
          const int c = tid, b0 = c & ~15, cl = c & 15;
          double x[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            double s = (i == cl) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) s = fma(-Lc[(b0 + k) * LD + b0 + i], x[k], s);
            x[i] = (i >= cl) ? s * rd[b0 + i] : 0.0;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) Inv[c * LD + b0 + i] = x[i];
        }
        __syncthreads();
        // 32-level: Inv21 = -Inv22 (L21 Inv11) for both 32-blocks; T1 staged in S
        {
          const int hb = tid >> 7, o = tid & 127;
          const int R0 = 32 * hb + 16, C0 = 32 * hb;
          const int r = o & 15;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = (o >> 4) + 8 * u;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k) s = fma(Lc[(C0 + k) * LD + R0 + r], Inv[(C0 + c) * LD + C0 + k], s);
            S[(C0 + c) * LD + R0 + r] = s;
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = (o >> 4) + 8 * u;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k) s = fma(Inv[(R0 + k) * LD + R0 + r], S[(C0 + c) * LD + R0 + k], s);
            Inv[(C0 + c) * LD + R0 + r] = -s;
          }
          __syncthreads();
        }
        // 64-level
        {
          const int r = tid & 31;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = (tid >> 5) + 8 * u;
            double s = 0.0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) s = fma(Lc[k * LD + 32 + r], Inv[c * LD + k], s);
            S[c * LD + 32 + r] = s;
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = (tid >> 5) + 8 * u;
            double s = 0.0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) s = fma(Inv[(32 + k) * LD + 32 + r], S[c * LD + 32 + k], s);
            Inv[c * LD + 32 + r] = -s;
          }
          __syncthreads();
        }
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
    __threadfence();
    grid.sync();
    if (*(volatile int*)prm.info != 0) break;  // uniform: written before the barrier
    if (j1 >= n) break;
    // ---- trailing update: C -= X X^T on the tiles that touch the lower triangle / the border ---
    {
      const int rt_count = (r_end - j1 + NB - 1) / NB, ct_count = (n - j1 + NB - 1) / NB;
      const int total = ct_count * rt_count - ct_count * (ct_count - 1) / 2;
      const int tx = tid & 15, ty = tid >> 4;
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
          S[k * LD + r] = (kk && rbase + r < r_end) ? A[(int64_t)(j0 + k) * ld + grow(rbase + r)] : 0.0;
          Lc[k * LD + r] = (kk && cbase + r < n) ? A[(int64_t)(j0 + k) * ld + cbase + r] : 0.0;
        }
        __syncthreads();
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
        for (int k = 0; k < NB; ++k) {
          double av[4], bv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) av[a] = S[k * LD + tx + 16 * a];
#pragma unroll
          for (int b = 0; b < 4; ++b) bv[b] = Lc[k * LD + ty + 16 * b];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int c = cbase + ty + 16 * b;
          if (c >= n) continue;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int r = rbase + tx + 16 * a;
            if (r < r_end && r >= c) A[(int64_t)c * ld + grow(r)] -= acc[a][b];
          }
        }
      }
    }
    __threadfence();
    grid.sync();
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
