// C[M x n] = T[M x Kd] B[Kd x n] (+ u v^T) on the FP64 tensor cores (DMMA m8n8k4), all row-major.
//
// Two callers, one shape -- a small square left factor against a wide M x n panel of Kuf:
//   * whitened SGPR statistics: A_r = L^-1 Kuf_r with the replicated explicit inverse factor
//     (T lower triangular: k blocks right of the diagonal are skipped), gpflow's operation order
//     A = L^-1 Kuf / sigma before A A^T (oak/utils.py:186-190);
//   * the cotangent of the training step: W = 2 G_Phi Kuf + g_b y^T (training.py), formerly a cuBLAS
//     DGEMM (cutlass_80 d884gemm) plus a rank-1 update.
// CTA tile 64 x 128 (4 warps as 2 x 2, warp tile 32 x 64: 64 accumulators per lane; two CTAs per SM so that one
// CTA's stage barrier is covered by the other -- the 128 x 128 / one-CTA-per-SM build, OAK_PGEMM_BM=128, idles the
// FP64 tensor pipe at every barrier), 16-wide k stages in a 3-stage cp.async pipeline.  The A fragment ("8 rows x 4 k") is read as 16-byte pairs of consecutive k, the
// B fragment ("4 k x 8 columns") as 8-byte words from the k-major staged panel; the k values of a stage are
// dealt to the four k-lanes as kappa(q, s) = 8 (s >> 1) + 2 q + (s & 1) for both operands, which makes both
// access patterns bank-conflict free (row strides 24 and 130 doubles).  Tiles are handed out by an atomic
// counter (row blocks of a triangular T differ 8x in cost), heaviest row block first inside each column
// tile so that the 16 row blocks of one panel column run close together and share it through L2.
// A device-side gate lets the launch be a no-op without a host decision (the SGPR route flag).
#include "oak_common.cuh"

namespace oak {

namespace pgemm {
#ifndef OAK_PGEMM_STAGES
#define OAK_PGEMM_STAGES 3
#endif
#ifndef OAK_PGEMM_BM
#define OAK_PGEMM_BM 64  // 64: 4 warps (2 x 2), two CTAs per SM; 128: 8 warps (4 x 2), one CTA per SM
#endif
constexpr int BM = OAK_PGEMM_BM, BN = 128, KT = 16, kStages = OAK_PGEMM_STAGES, kThreads = 2 * BM;
constexpr int kMinBlocks = BM == 64 ? 2 : 1;
static_assert(BM == 64 || BM == 128, "OAK_PGEMM_BM must be 64 or 128");
constexpr int SA = 24;            // doubles per staged T row (16 + 8: conflict-free LDS.128)
constexpr int SB = BN + 2;        // doubles per staged B row (conflict-free LDS.64)
constexpr int kADoubles = BM * SA, kBDoubles = KT * SB;
constexpr int kStageDoubles = kADoubles + kBDoubles;
constexpr size_t kSmemBytes = (size_t)kStages * kStageDoubles * sizeof(double) + 16;
}  // namespace pgemm

struct PGemmParams {
  const double* T;
  const double* B;
  double* C;
  const double* u;   // optional rank-1 update C += u v^T
  const double* v;
  int64_t ldt, ldb, ldc;
  int64_t n, n_read;  // output columns; readable columns of B (even, zero-filled beyond)
  int M, Kd, lower;
  int row_blocks;
  int64_t units;
  const int* gate;    // optional: the kernel returns at once when *gate == 0
  int* counter;       // tile scheduler, zeroed by the launcher
};

__device__ __forceinline__ void pg_cp16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void pg_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(pgemm::kThreads, pgemm::kMinBlocks) panel_gemm_dmma_kernel(const PGemmParams prm) {
  using namespace pgemm;
  if (prm.gate != nullptr && *prm.gate == 0) return;
  extern __shared__ __align__(16) double smem[];
  int* s_unit = reinterpret_cast<int*>(smem + kStages * kStageDoubles);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, q = lane & 3;

  for (;;) {
    __syncthreads();  // everybody is done with the previous unit's buffers and s_unit
    if (tid == 0) *s_unit = atomicAdd(prm.counter, 1);
    __syncthreads();
    const int64_t unit = *s_unit;
    if (unit >= prm.units) break;
    const int64_t ct = unit / prm.row_blocks;
    const int rb = prm.row_blocks - 1 - (int)(unit - ct * prm.row_blocks);
    const int64_t n0 = ct * BN;
    const int m0 = rb * BM;
    int kend = prm.Kd;
    if (prm.lower && m0 + BM < kend) kend = m0 + BM;
    const int steps = (kend + KT - 1) / KT;

    auto load_stage = [&](int step, int buf) {
      double* dA = smem + buf * kStageDoubles;
      double* dB = dA + kADoubles;
      const int kbase = step * KT;
#pragma unroll
      for (int i = 0; i < (BM * (KT / 2)) / kThreads; ++i) {
        const int idx = tid + i * kThreads;
        const int row = idx >> 3, ch = idx & 7;
        const bool ok = (m0 + row < prm.M) && (kbase + 2 * ch < prm.Kd);
        const double* src = prm.T + (int64_t)(ok ? m0 + row : 0) * prm.ldt + (ok ? kbase + 2 * ch : 0);
        pg_cp16(dA + row * SA + 2 * ch, src, ok);
      }
#pragma unroll
      for (int i = 0; i < (KT * (BN / 2)) / kThreads; ++i) {
        const int idx = tid + i * kThreads;
        const int kr = idx >> 6, ch = idx & 63;
        const bool ok = (kbase + kr < prm.Kd) && (n0 + 2 * ch < prm.n_read);
        const double* src = prm.B + (int64_t)(ok ? kbase + kr : 0) * prm.ldb + (ok ? n0 + 2 * ch : 0);
        pg_cp16(dB + kr * SB + 2 * ch, src, ok);
      }
      asm volatile("cp.async.commit_group;\n");
    };

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

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
      const double* As = smem + (step % kStages) * kStageDoubles + (wm * 32 + g) * SA + 2 * q;
      const double* Bs = smem + (step % kStages) * kStageDoubles + kADoubles + (2 * q) * SB + wn * 64 + g;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        double a[4][2], b[8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 x = *reinterpret_cast<const double2*>(As + i * 8 * SA + 8 * t);
          a[i][0] = x.x;
          a[i][1] = x.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          b[j][0] = Bs[(8 * t) * SB + j * 8];
          b[j][1] = Bs[(8 * t + 1) * SB + j * 8];
        }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) pg_dmma(acc[i][j][0], acc[i][j][1], a[i][s], b[j][s]);
      }
    }
    asm volatile("cp.async.wait_group 0;\n");

    // epilogue: C fragment = row g, columns 2q, 2q + 1 of each 8 x 8 block
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + wm * 32 + i * 8 + g;
      if (row >= prm.M) continue;
      const double uu = prm.u ? prm.u[row] : 0.0;
      double* crow = prm.C + (int64_t)row * prm.ldc;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t col = n0 + wn * 64 + j * 8 + 2 * q;
        double c0 = acc[i][j][0], c1 = acc[i][j][1];
        if (col + 1 < prm.n) {
          if (prm.u) {
            const double2 vv = *reinterpret_cast<const double2*>(prm.v + col);
            c0 = fma(uu, vv.x, c0);
            c1 = fma(uu, vv.y, c1);
          }
          *reinterpret_cast<double2*>(crow + col) = make_double2(c0, c1);
        } else if (col < prm.n) {
          if (prm.u) c0 = fma(uu, prm.v[col], c0);
          crow[col] = c0;
        }
      }
    }
  }
}

// C (M x n, ldc) = T (M x Kd, ldt) B (Kd x n, ldb) [+ u v^T]; `lower`: T is lower triangular (its upper part
// must hold zeros).  All leading dimensions even, all bases 16-byte aligned; columns [n, n_read) of B only
// have to be readable.  d_counter: one int of scratch.
int panel_gemm_dmma(const double* T, int64_t ldt, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int Kd,
                    int64_t n, int lower, const double* u, const double* v, const int* d_gate, int* d_counter,
                    int device, cudaStream_t stream) {
  using namespace pgemm;
  if (M <= 0 || n <= 0 || Kd <= 0) return 0;
  OAK_REQUIRE(ldt % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0, "panel_gemm_dmma: leading dimensions must be even");
  OAK_REQUIRE((reinterpret_cast<uintptr_t>(T) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) % 16 == 0,
              "panel_gemm_dmma: operands must be 16-byte aligned");
  OAK_REQUIRE(u == nullptr || (v != nullptr && reinterpret_cast<uintptr_t>(v) % 16 == 0),
              "panel_gemm_dmma: rank-1 factors missing or unaligned");
  static int sms_cached[64] = {0};
  int sms = (device >= 0 && device < 64) ? sms_cached[device] : 0;
  if (sms == 0) {
    OAK_CUDA(cudaFuncSetAttribute(panel_gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSmemBytes));
    OAK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (device >= 0 && device < 64) sms_cached[device] = sms;
  }
  PGemmParams prm;
  prm.T = T; prm.B = B; prm.C = C; prm.u = u; prm.v = v;
  prm.ldt = ldt; prm.ldb = ldb; prm.ldc = ldc;
  prm.n = n;
  prm.n_read = (n + 1) / 2 * 2;
  OAK_REQUIRE(prm.n_read <= ldb, "panel_gemm_dmma: ldb must cover the column count rounded up to even");
  prm.M = M; prm.Kd = Kd; prm.lower = lower;
  prm.row_blocks = (M + BM - 1) / BM;
  prm.units = (int64_t)prm.row_blocks * ((n + BN - 1) / BN);
  prm.gate = d_gate;
  prm.counter = d_counter;
  OAK_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), stream));
  const int64_t resident = (int64_t)sms * kMinBlocks;
  const int64_t grid = prm.units < resident ? prm.units : resident;
  panel_gemm_dmma_kernel<<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(prm);
  OAK_LAUNCHED();
  return 0;
}

}  // namespace oak

using namespace oak;

extern "C" size_t oak_panel_gemm_work_bytes(void) { return 64; }

extern "C" int oak_panel_gemm_f64(const double* d_T, int64_t ldt, const double* d_B, int64_t ldb, double* d_C,
                                  int64_t ldc, int64_t m, int64_t kd, int64_t n, int lower, const double* d_u,
                                  const double* d_v, void* d_work, void* stream) {
  OAK_REQUIRE(d_T && d_B && d_C && d_work, "oak_panel_gemm_f64: null argument");
  OAK_REQUIRE(m >= 0 && kd >= 0 && n >= 0 && m <= INT32_MAX / 2 && kd <= INT32_MAX / 2, "oak_panel_gemm_f64: bad size");
  OAK_REQUIRE(ldt >= kd && ldb >= n && ldc >= n, "oak_panel_gemm_f64: leading dimension too small");
  int dev = 0;
  OAK_CUDA(cudaGetDevice(&dev));
  return panel_gemm_dmma(d_T, ldt, d_B, ldb, d_C, ldc, (int)m, (int)kd, n, lower, d_u, d_v, nullptr, (int*)d_work, dev,
                         (cudaStream_t)stream);
}
