// Blocked Cholesky factorisation with border rows in ONE cooperative launch (sm_100a).
//
// The M x M factorisations of the SGPR bound (L = chol(Kuu + jitter I), LB = chol(I + A A^T),
// oak/utils.py:188-193) sit on the critical path of every ELBO evaluation and are replicated on every
// rank; cuSOLVER's potrf needs ~1 ms at M = 1024 (a dozen launches with two ~450 us panel kernels).
// Here the whole factorisation is a single persistent kernel:
//
//   for each 64-column panel
//     phase 1  every CTA that owns rows of the panel factors the 64 x 64 diagonal block redundantly in
//              registers (right-looking on the bordered block [S ; I], FOUR pivots per pair of barriers: the
//              4 x 4 diagonal block is factored redundantly by every thread, the rank-4 update follows; L11 and
//              inv(L11) come out together) and solves its 16-row tiles of the panel as a small matrix
//              product X = A21 inv(L11)^T; CTA 0 writes L11 back
//     grid barrier
//     phase 2  trailing update C -= X X^T on the 64 x 64 tiles of the lower triangle (DMMA m8n8k4), one tile
//              per CTA at M = 1024; operand loads and the read-modify-write of C are batched (all loads of a
//              thread in flight together: element-wise loops were chains of 16 dependent L2 round trips)
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
constexpr size_t kSmemBytes = (3 * (size_t)kBlk + 8 * NB) * sizeof(double);  // S | Lc | Inv | group columns | multipliers
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
      double* const Cs = rd;           // [4][NB]: the four columns of the group as published
      double* const Ms = rd + 4 * NB;  // [NB][4]: the four multipliers of every row
      // Groups of FOUR pivots per pair of barriers (one barrier per pivot before: the 64 dependent
      // publish -> barrier -> rsqrt -> update rounds were 47 % of the factorisation).  The owners publish the
      // group's four columns; every thread factors the 4 x 4 diagonal block redundantly (four dependent rsqrt
      // chains -- the irreducible part) and forms the four multipliers of its own row,
      //   m_q = (W(i, j+q) - sum_{q' < q} m_q' l_{q q'}) / l_qq,
      // with the rule of the unified matrix deciding which terms exist: pivot column q' reaches column k of row i
      // iff (i > j+q' and k <= i) [S part] or i <= j+q' [E part].  The multipliers are published, second barrier,
      // and every thread applies the rank-4 update to its registers.  The register slots ROTATE so that the
      // group's own columns always sit in slot 0 and the loop body is the same code for every group (a fully
      // unrolled body is ~100 KB of straight-line code executed once per panel: instruction-fetch bound).
      // Two copies of the body: 16 live slots for the first eight groups, 8 for the rest.
      auto pivot_group = [&](const int g, auto slots_tag) -> bool {
        constexpr int kSlots = decltype(slots_tag)::value;
        const int j = 4 * g;
        Cs[kq * NB + i] = w[0];
        __syncthreads();
        // diagonal block, lower part: a_pq = W(j+p, j+q)
        const double a00 = Cs[j], a10 = Cs[j + 1], a20 = Cs[j + 2], a30 = Cs[j + 3];
        double a11 = Cs[NB + j + 1];
        const double a21 = Cs[NB + j + 2], a31 = Cs[NB + j + 3];
        double a22 = Cs[2 * NB + j + 2];
        const double a32 = Cs[2 * NB + j + 3];
        double a33 = Cs[3 * NB + j + 3];
        const double c0 = Cs[i], c1 = Cs[NB + i], c2 = Cs[2 * NB + i], c3 = Cs[3 * NB + i];
        auto bad = [](double v) { return !(v > 0.0) || !(v < 1.0e300); };  // uniform: same value in every thread
        if (bad(a00)) { failj = j; return false; }
        const double r0 = rsqrt(a00);
        const double l10 = a10 * r0, l20 = a20 * r0, l30 = a30 * r0;
        a11 = fma(-l10, l10, a11);
        if (bad(a11)) { failj = j + 1; return false; }
        const double r1 = rsqrt(a11);
        const double l21 = fma(-l20, l10, a21) * r1, l31 = fma(-l30, l10, a31) * r1;
        a22 = fma(-l21, l21, fma(-l20, l20, a22));
        if (bad(a22)) { failj = j + 2; return false; }
        const double r2 = rsqrt(a22);
        const double l32 = fma(-l31, l21, fma(-l30, l20, a32)) * r2;
        a33 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, a33)));
        if (bad(a33)) { failj = j + 3; return false; }
        const double r3 = rsqrt(a33);
        // this row's multipliers; rule(q', q): does pivot q' reach column j+q of row i
        const int p = i - j;  // < 0: E row above the block, 0..3: inside it, > 3: S row below it
        auto reach = [&](int qp, int q) { return (p > qp && q <= p) || p <= qp; };
        const double m0 = (p == 0) ? r0 : c0 * r0;
        double t1 = c1;
        if (reach(0, 1)) t1 = fma(-m0, l10, t1);
        const double m1 = (p == 1) ? r1 : t1 * r1;
        double t2 = c2;
        if (reach(0, 2)) t2 = fma(-m0, l20, t2);
        if (reach(1, 2)) t2 = fma(-m1, l21, t2);
        const double m2 = (p == 2) ? r2 : t2 * r2;
        double t3 = c3;
        if (reach(0, 3)) t3 = fma(-m0, l30, t3);
        if (reach(1, 3)) t3 = fma(-m1, l31, t3);
        if (reach(2, 3)) t3 = fma(-m2, l32, t3);
        const double m3 = (p == 3) ? r3 : t3 * r3;
        {
          // thread (i, kq) files pivot j + kq: L(i, j+kq) below the diagonal, inv(L11)(j+kq, i) above it
          const double mq = kq == 0 ? m0 : kq == 1 ? m1 : kq == 2 ? m2 : m3;
          const int jj = j + kq;
          if (i > jj) {
            Lc[jj * LD + i] = mq;
          } else if (i == jj) {
            const double ad = kq == 0 ? a00 : kq == 1 ? a11 : kq == 2 ? a22 : a33;
            Lc[jj * LD + jj] = ad * mq;  // sqrt(a_jj) = a_jj * rsqrt(a_jj)
            Inv[jj * LD + jj] = mq;
          } else {
            Inv[i * LD + jj] = mq;       // inv(L11)(jj, i) = (L11^-T)(i, jj)
          }
          Ms[i * 4 + kq] = mq;
        }
        __syncthreads();
        // rank-4 update of the columns right of the block (k >= j + 4): pivot q acts on this row unless the row
        // lies inside the block below pivot q's own row (its S part ends at column i < k)
        const double u0 = (p >= 1 && p <= 3) ? 0.0 : -m0;
        const double u1 = (p >= 2 && p <= 3) ? 0.0 : -m1;
        const double u2 = (p == 3) ? 0.0 : -m2;
        const double u3 = -m3;
        const int klim = (p > 3) ? i : NB - 1;
#pragma unroll
        for (int u = 1; u < kSlots; ++u) {
          const int k = kq + 4 * (g + u);
          if (k <= klim) {
            const double2 ma = *reinterpret_cast<const double2*>(Ms + 4 * k);
            const double2 mb = *reinterpret_cast<const double2*>(Ms + 4 * k + 2);
            w[u] = fma(u3, mb.y, fma(u2, mb.x, fma(u1, ma.y, fma(u0, ma.x, w[u]))));
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
          // (all global loads of a tile are issued before the first use: one L2 round trip, not one per element)
          constexpr int kPer = NB * TR / kThreads;
          const int lr = tid & (TR - 1), lk = tid >> 4;  // element e = tid + it * kThreads: row lr, column lk + 16 it
          const int row = r0 + lr;
          double* const gp = A + (int64_t)(j0 + lk) * ld + grow(row);
          double vin[kPer];
#pragma unroll
          for (int it = 0; it < kPer; ++it)
            vin[it] = (row < r_end && lk + 16 * it < nb) ? gp[(int64_t)(16 * it) * ld] : 0.0;
#pragma unroll
          for (int it = 0; it < kPer; ++it) S[(lk + 16 * it) * LD + lr] = vin[it];
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
#pragma unroll
          for (int it = 0; it < kPer; ++it)
            if (row < r_end && lk + 16 * it < nb) gp[(int64_t)(16 * it) * ld] = S[(lk + 16 * it) * LD + lr];
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
        {
          // both operand blocks: 2 x 16 independent loads per thread in flight, then the shared-memory stores
          // (element-by-element the loop was a chain of 16 L2 round trips: 4-6 us of the 6-11 us per tile)
          const int lr = tid & (NB - 1), lk = tid >> 6;  // element (row lr, column lk + 4 it)
          const bool ra = rbase + lr < r_end, rb = cbase + lr < n;
          const double* const pa_g = A + (int64_t)(j0 + lk) * ld + grow(rbase + lr);
          const double* const pb_g = A + (int64_t)(j0 + lk) * ld + cbase + lr;
          double va[NB / 4], vb[NB / 4];
#pragma unroll
          for (int it = 0; it < NB / 4; ++it) {
            const bool kk = lk + 4 * it < nb;
            va[it] = (kk && ra) ? pa_g[(int64_t)(4 * it) * ld] : 0.0;
            vb[it] = (kk && rb) ? pb_g[(int64_t)(4 * it) * ld] : 0.0;
          }
#pragma unroll
          for (int it = 0; it < NB / 4; ++it) {
            Xa[(lk + 4 * it) * LX + lr] = va[it];
            Xb[(lk + 4 * it) * LX + lr] = vb[it];
          }
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
        {
          // C -= acc: the 16 loads first, then the 16 stores (a load -> subtract -> store chain per element is
          // sixteen dependent L2 round trips)
          double cv[2][4][2];
          double* cp[2][4];
          bool ok[2][4][2];
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int r = rbase + wm * 16 + a * 8 + g;
            const int64_t gr = grow(r);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int c = cbase + wn * 32 + b * 8 + 2 * q;
              cp[a][b] = A + (int64_t)c * ld + gr;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                ok[a][b][e] = r < r_end && c + e < n && r >= c + e;
                cv[a][b][e] = ok[a][b][e] ? cp[a][b][(int64_t)e * ld] : 0.0;
              }
            }
          }
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                if (ok[a][b][e]) cp[a][b][(int64_t)e * ld] = cv[a][b][e] - acc[a][b][e];
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
                  double* d_logdet, int device, cudaStream_t stream, int max_ctas) {
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
  int grid = want < sms ? want : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
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
