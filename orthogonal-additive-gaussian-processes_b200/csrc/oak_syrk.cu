// Phi += Kuf Kuf^T on the FP64 tensor cores: the "genuinely dense" contraction of the SGPR statistics
// (gpflow SGPR.elbo forms A A^T with A = L^-1 Kuf / sigma, oak/utils.py:186-190; here the
// un-whitened Phi = Kuf Kuf^T is accumulated and whitened once in the M^3 tail).
//
// Why a kernel of our own: cuBLAS DSYRK runs a 32x32-tile sm_80 kernel at full-GEMM cost on this
// shape, and a batched DGEMM over the lower-triangle blocks still computes the diagonal blocks in
// full and leaves 148 - 136 SMs idle at M = 1024.  This kernel
//   * enumerates only the 64 x 64 tiles of the lower triangle (136 at M = 1024),
//   * cuts the k axis (the N points of the chunk) into S slices and deals the (slice, tile) units
//     round-robin to persistent CTAs (stream-K): S is chosen so that the units fill the resident
//     CTAs almost exactly (e.g. S = 13 -> 1768 units on 296 resident CTAs, 99.5 %),
//   * keeps each unit's 64 x 64 partial sum in registers (DMMA m8n8k4, 4 warps x 32 x 32) and
//     writes it once; a second kernel adds the S partials of every tile into Phi in a fixed order,
//     so the result is deterministic (no floating-point atomics).
// Operands are row-major Kuf rows (contiguous along k): both the A and the B fragment of
// mma.m8n8k4 read "8 rows x 4 consecutive k", so one shared-memory layout serves both; the k values
// of a 16-wide stage are dealt to the four k-lanes as contiguous groups of four (any permutation
// of k applied to both operands leaves the product unchanged), which turns the fragment loads
// into conflict-free 16-byte loads.
#include "oak_common.cuh"

namespace oak {

namespace syrk {
constexpr int kTile = 64;         // CTA tile (rows and columns of Phi)
#ifndef OAK_SYRK_KT
#define OAK_SYRK_KT 16
#endif
#ifndef OAK_SYRK_STAGES
#define OAK_SYRK_STAGES 3  // measured (65536-point chunks): 3 stages 56.7 ms per 10^6 points, 4: 57.9, 6: 56.8; KT=32: 58.0
#endif
constexpr int kKT = OAK_SYRK_KT;  // k values per pipeline stage (16 or 32)
constexpr int kStages = OAK_SYRK_STAGES;
constexpr int kRowStride = kKT + 2;  // doubles per staged row (+2: conflict-free LDS.128)
constexpr int kChunks = kKT / 2;     // 16-byte chunks per staged row
constexpr int kThreads = 128;
constexpr int kStageDoubles = 2 * kTile * kRowStride;  // A rows then B rows
constexpr size_t kSmemBytes = (size_t)kStages * kStageDoubles * sizeof(double);
}  // namespace syrk

struct SyrkParams {
  const double* A;   // row-major [m][lda], k contiguous
  const double* A_alt;  // operand used instead of A when *route != 0 (the whitened block L^-1 Kuf), or null
  const int* route;     // device-side route flag of the SGPR statistics, or null
  double* partial;   // [units][64][64]
  int64_t lda;
  int m, k_steps;    // k_steps = number of 16-wide k stages (the caller zero-pads the last one)
  int tiles, nb, slices, units;
};

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

#ifndef OAK_SYRK_MINB
#define OAK_SYRK_MINB 2  // 182 registers, 2 CTAs / SM: 54.3 ms per 10^6 points; 3 (153 registers): 57.2; 4: 58.0
#endif
__global__ void __launch_bounds__(syrk::kThreads, OAK_SYRK_MINB) syrk_lower_dmma_kernel(const SyrkParams prm) {
  using namespace syrk;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;  // 2 x 2 warps, 32 x 32 each
  const int g = lane >> 2, q = lane & 3;    // fragment row / k-lane
  const double* const Aop = (prm.route != nullptr && *prm.route != 0) ? prm.A_alt : prm.A;

  for (int u = blockIdx.x; u < prm.units; u += gridDim.x) {
    const int s = u / prm.tiles, t = u - s * prm.tiles;
    // tile (bi >= bj) of the lower triangle, row-block major
    int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (bi * (bi + 1) / 2 > t) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    const int bj = t - bi * (bi + 1) / 2;
    const int k0 = (int)((int64_t)s * prm.k_steps / prm.slices);
    const int k1 = (int)((int64_t)(s + 1) * prm.k_steps / prm.slices);
    const int steps = k1 - k0;
    const bool skip_mma = (bi == bj) && wm == 0 && wn == 1;

    // stage loader: 128 rows x kChunks chunks of 16 bytes; rows >= m are zero-filled
    auto load_stage = [&](int step, int buf) {
      double* dst = smem + buf * kStageDoubles;
      const int64_t kbase = (int64_t)(k0 + step) * kKT;
#pragma unroll
      for (int i = 0; i < kChunks; ++i) {
        const int idx = tid + i * kThreads;
        const int row = idx / kChunks, chunk = idx % kChunks;  // row 0..127 (A rows 0..63, B rows 64..127)
        const int grow = (row < kTile) ? bi * kTile + row : bj * kTile + (row - kTile);
        const bool ok = grow < prm.m;
        const double* src = Aop + (int64_t)(ok ? grow : 0) * prm.lda + kbase + chunk * 2;
        cp_async16_zfill(dst + row * kRowStride + chunk * 2, src, ok);
      }
      asm volatile("cp.async.commit_group;\n");
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // pipeline prologue
#pragma unroll
    for (int p = 0; p < kStages - 1; ++p) {
      if (p < steps) load_stage(p, p);
      else asm volatile("cp.async.commit_group;\n");
    }
    for (int step = 0; step < steps; ++step) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(kStages - 2));
      __syncthreads();  // stage `step` has landed; everybody is done with stage step-1's buffer
      {
        const int nxt = step + kStages - 1;
        if (nxt < steps) load_stage(nxt, nxt % kStages);
        else asm volatile("cp.async.commit_group;\n");
      }
      const double* As = smem + (step % kStages) * kStageDoubles + (wm * 32 + g) * kRowStride + q * 4;
      const double* Bs = smem + (step % kStages) * kStageDoubles + (kTile + wn * 32 + g) * kRowStride + q * 4;
      // the upper-right 32 x 32 quadrant of a diagonal tile lies strictly above the diagonal: its
      // warp only takes part in the staging, and leaves its DMMA slots to the co-resident CTAs
      if (skip_mma) continue;
      // this lane's k values: groups of four consecutive ones, 16 apart (kk -> (kk/4)*16 + 4q + kk%4)
#pragma unroll
      for (int h = 0; h < kKT / 16; ++h) {
#ifndef OAK_SYRK_HALF_FRAGS
#define OAK_SYRK_HALF_FRAGS 0  // 1: hold two k values per fragment load; with OAK_SYRK_MINB=4 (128 registers,
                               // 4 CTAs / SM) measured SLOWER: 58.0-58.6 vs 55.7 ms per 10^6 points (profiles/)
#endif
#if OAK_SYRK_HALF_FRAGS
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          double a[4][2], b[4][2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double2 x = *reinterpret_cast<const double2*>(As + i * 8 * kRowStride + h * 16 + 2 * h2);
            a[i][0] = x.x; a[i][1] = x.y;
            const double2 y = *reinterpret_cast<const double2*>(Bs + i * 8 * kRowStride + h * 16 + 2 * h2);
            b[i][0] = y.x; b[i][1] = y.y;
          }
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i][kk], b[j][kk]);
        }
#else
        double a[4][4], b[4][4];  // [row block][k step]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 x0 = *reinterpret_cast<const double2*>(As + i * 8 * kRowStride + h * 16);
          const double2 x1 = *reinterpret_cast<const double2*>(As + i * 8 * kRowStride + h * 16 + 2);
          a[i][0] = x0.x; a[i][1] = x0.y; a[i][2] = x1.x; a[i][3] = x1.y;
          const double2 y0 = *reinterpret_cast<const double2*>(Bs + i * 8 * kRowStride + h * 16);
          const double2 y1 = *reinterpret_cast<const double2*>(Bs + i * 8 * kRowStride + h * 16 + 2);
          b[i][0] = y0.x; b[i][1] = y0.y; b[i][2] = y1.x; b[i][3] = y1.y;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i][kk], b[j][kk]);
#endif
      }
    }
    asm volatile("cp.async.wait_group 0;\n");
    __syncthreads();  // the buffers are reused by the next unit's prologue

    // partial tile -> workspace (row-major 64 x 64); C fragment: row g, columns 2q, 2q+1
    double* P = prm.partial + (int64_t)u * (kTile * kTile);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = wm * 32 + i * 8 + g, c = wn * 32 + j * 8 + 2 * q;
        *reinterpret_cast<double2*>(P + r * kTile + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
  }
}

// Phi(lower, column-major m x m) += sum over slices of the tile partials, fixed order.
// grid = tiles, 256 threads; thread -> (r = tid % 64 fastest: contiguous in column-major Phi)
__global__ void __launch_bounds__(256) syrk_reduce_kernel(const double* __restrict__ partial, int tiles, int slices,
                                                          int m, double* __restrict__ C) {
  using namespace syrk;
  __shared__ double sh[kTile][kTile + 1];
  const int t = blockIdx.x;
  int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (bi * (bi + 1) / 2 > t) --bi;
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const int bj = t - bi * (bi + 1) / 2;
  // coalesced reads of the row-major partials, transposed through shared memory
  for (int e = threadIdx.x; e < kTile * kTile; e += 256) {
    double v = 0.0;
    for (int s = 0; s < slices; ++s) v += partial[((int64_t)s * tiles + t) * (kTile * kTile) + e];
    sh[e >> 6][e & 63] = v;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kTile * kTile; e += 256) {
    const int c = e >> 6, r = e & 63;  // r fastest
    const int i = bi * kTile + r, j = bj * kTile + c;
    if (i < m && j < m && i >= j) C[(int64_t)j * m + i] += sh[r][c];
  }
}

// ---- host side ----------------------------------------------------------------------------
static int syrk_resident_ctas(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int sms = 148, per_sm = 2;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaFuncSetAttribute(syrk_lower_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syrk::kSmemBytes);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, syrk_lower_dmma_kernel, syrk::kThreads,
                                                    syrk::kSmemBytes) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int r = sms * per_sm;
  if (device >= 0 && device < 64) cached[device] = r;
  return r;
}

// number of k slices: the units (tiles x slices) should fill the resident CTAs in whole rounds;
// the smallest slice count within 1 % of the best fill is taken (fewer partials to write and add)
static int syrk_pick_slices(int tiles, int k_steps, int resident, size_t work_bytes) {
  const size_t unit_bytes = (size_t)syrk::kTile * syrk::kTile * sizeof(double);
  int max_s = k_steps / 32;  // at least 32 pipeline stages per unit
  if (max_s > 48) max_s = 48;
  if (max_s < 1) max_s = 1;
  double best_eff = 0.0;
  double eff[49] = {0.0};
  for (int s = 1; s <= max_s; ++s) {
    const int64_t units = (int64_t)tiles * s;
    eff[s] = 0.0;
    if ((size_t)units * unit_bytes > work_bytes) break;
    const int64_t rounds = (units + resident - 1) / resident;
    eff[s] = (double)units / (double)(rounds * resident);
    if (eff[s] > best_eff) best_eff = eff[s];
  }
  for (int s = 1; s <= max_s; ++s)
    if (eff[s] >= best_eff - 0.01 && eff[s] > 0.0) return s;
  return 1;
}

// workspace for the partial tiles: enough for any slice count the picker may choose (<= 48), capped
size_t syrk_dmma_work_bytes(int m) {
  const int nb = (m + syrk::kTile - 1) / syrk::kTile;
  const size_t tiles = (size_t)nb * (nb + 1) / 2;
  const size_t unit_bytes = (size_t)syrk::kTile * syrk::kTile * sizeof(double);
  size_t b = tiles * 48 * unit_bytes;
  const size_t cap = 256ull << 20, floor_b = tiles * unit_bytes;
  if (b > cap) b = cap;
  if (b < floor_b) b = floor_b;  // one slice always fits
  return b;
}

// C(lower, column-major m x m) += A A^T, A row-major [m][lda] with k valid columns; columns
// [k, roundup16(k)) of A must be readable (they are zeroed here).  `work` holds the partials.
int syrk_lower_dmma(int m, int64_t k, double* A, int64_t lda, double* C, double* work, size_t work_bytes,
                    int device, cudaStream_t stream, double* A_alt, const int* d_route) {
  using namespace syrk;
  if (m <= 0 || k <= 0) return 0;
  const int64_t k_pad = (k + kKT - 1) / kKT * kKT;
  OAK_REQUIRE(k_pad <= lda, "syrk_lower_dmma: the padded k range exceeds the leading dimension");
  OAK_REQUIRE(lda % 2 == 0 && (reinterpret_cast<uintptr_t>(A) % 16 == 0), "syrk_lower_dmma: unaligned operand");
  if (k_pad > k)
    OAK_CUDA(cudaMemset2DAsync(A + k, (size_t)lda * sizeof(double), 0, (size_t)(k_pad - k) * sizeof(double), (size_t)m,
                               stream));
  if (k_pad > k && A_alt)
    OAK_CUDA(cudaMemset2DAsync(A_alt + k, (size_t)lda * sizeof(double), 0, (size_t)(k_pad - k) * sizeof(double),
                               (size_t)m, stream));
  SyrkParams prm;
  prm.A = A;
  prm.A_alt = A_alt;
  prm.route = A_alt ? d_route : nullptr;
  prm.partial = work;
  prm.lda = lda;
  prm.m = m;
  prm.k_steps = (int)(k_pad / kKT);
  prm.nb = (m + kTile - 1) / kTile;
  prm.tiles = prm.nb * (prm.nb + 1) / 2;
  const int resident = syrk_resident_ctas(device);
  prm.slices = syrk_pick_slices(prm.tiles, prm.k_steps, resident, work_bytes);
  prm.units = prm.tiles * prm.slices;
  OAK_REQUIRE((size_t)prm.units * kTile * kTile * sizeof(double) <= work_bytes,
              "syrk_lower_dmma: workspace too small");
  const int grid = prm.units < resident ? prm.units : resident;
  syrk_lower_dmma_kernel<<<grid, kThreads, kSmemBytes, stream>>>(prm);
  OAK_LAUNCHED();
  syrk_reduce_kernel<<<prm.tiles, 256, 0, stream>>>(work, prm.tiles, prm.slices, m, C);
  OAK_LAUNCHED();
  return 0;
}

}  // namespace oak
