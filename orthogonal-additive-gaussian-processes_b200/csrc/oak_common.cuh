// Shared declarations for the OAK B200 hot-path library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "oak_b200.h"

namespace oak {

// ---- error plumbing ------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;

#define OAK_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::oak::set_error(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" + \
                       __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define OAK_REQUIRE(cond, msg)                   \
  do {                                           \
    if (!(cond)) {                               \
      ::oak::set_error(std::string(msg));        \
      return 2;                                  \
    }                                            \
  } while (0)

#define OAK_LAUNCHED()                                                                      \
  do {                                                                                      \
    ::oak::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ::oak::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " (" + \
                       __FILE__ + ":" + std::to_string(__LINE__) + ")");                    \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

// ---- device-side parameter records ---------------------------------------------------
// One record per sub-kernel, in KERNEL ORDER: all RBF dims first, then the discrete dims.
struct DimDev {
  int32_t type;       // OAK_DIM_*
  int32_t column;     // column of X
  int32_t measure;    // OAK_MEASURE_*
  int32_t count;      // locations / components / categories
  int32_t table_off;  // discrete: offset (doubles) of the CxC table in the table blob; the
                      // C-entry diagonal follows it
  int32_t orig;       // index of this sub-kernel in the caller's ordering
  double inv_sqrt2_l; // 1 / (sqrt(2) l)
  double neg_log_s2;  // -ln(s^2)
  double s2;          // base variance
  double lengthscale;
  double c0, c1, c2;  // measure constants (see oak_prepare.cu)
  const double* v0;   // device copies of the per-measure arrays
  const double* v1;
  const double* v2;
};

constexpr int kPointPad = 128;  // prepared-point rows are padded to a multiple of this
constexpr int kExpTab = 256;    // 2^(j/256) table entries

inline int64_t padded(int64_t n) { return (n + kPointPad - 1) / kPointPad * kPointPad; }

}  // namespace oak

// Host-visible spec object behind the opaque handle.
struct oak_spec {
  int D = 0, Dc = 0, Dd = 0;  // total / continuous / discrete sub-kernels
  int depth = 0;              // max_interaction_depth as given
  int share_var = 1;
  int algo = 0;
  int device = 0;
  double sigma2[OAK_MAX_DEPTH + 1];    // coefficient of e_n for n = 0..max(depth,1)
  std::vector<oak::DimDev> h_dims;     // kernel order
  std::vector<int> pos_of_orig;        // original index -> kernel-order position
  std::vector<double> h_tables;        // discrete tables (host copy)
  std::vector<double> h_sobolG;        // categorical dims: G = B diag(p) B^T (unit variance)
  std::vector<int> sobol_off;          // [D] offset of G per kernel-order position
  double* d_sobolG = nullptr;
  oak::DimDev* d_dims = nullptr;
  double* d_inv_sqrt_v = nullptr;      // [D] 1/sqrt(var_s) per RBF dim (0 if unconstrained)
  double* d_neg_log_s2 = nullptr;      // [D]
  double* d_tables = nullptr;          // discrete tables blob
  int tables_len = 0;
  double* d_blob = nullptr;            // per-measure arrays
  const double* d_exptab = nullptr;    // 2^(j/256)
};

namespace oak {
const double* exp_table_device();  // lazily uploaded per device; nullptr on failure

// Internal launchers shared between translation units (device pointers; see oak_gram.cu).
int tile_rows_for_depth(int depth);
int gram_launch(const oak_spec* spec, const double2* prow, int64_t n_row_pad, int64_t row_begin,
                int64_t row_end, const double2* pcol, int64_t n_col_pad, int64_t col_begin,
                int64_t col_end, int mode, double* K, int64_t ldk, cudaStream_t stream);
int gram_diag_launch(const oak_spec* spec, const double2* pts, int64_t n, int64_t n_pad,
                     double* out, cudaStream_t stream);

// ---- fast FP64 exp(-z) ---------------------------------------------------------------
// Table-driven: -z = n ln2/256 + r, |r| <= ln2/512, exp(-z) = 2^(n>>8) * T[n&255] * e^r with a
// degree-4 Taylor polynomial (truncation 3.8e-17 relative).  8 FP64-pipe instructions; the
// index arithmetic, the clamp and the exponent insertion run on the integer pipe.  The
// single-step reduction leaves |z| * 1.1e-16 relative error (<= 8e-14 at the clamp), far
// inside the 1e-9 parity budget.  `tab` points at this lane's replica of the table in
// shared memory: entry j lives at tab[j * 16] so that a half-warp never conflicts.
#ifdef __CUDACC__
// Variant used by the Gram tile.  Measured on B200 (scripts/ubench): a DFMA blocks the issue
// port for two cycles and integer-ALU instructions (LOP3/SHF/LEA/IADD3) compete with it, while
// IMAD (FMA pipe) co-issues for free.  The index math is therefore phrased as
//   off = ((n << 7) & 0x7F80) | lane_bits          IMAD.SHL + one LOP3
//   hi  = n * 4096 + T'hi[j]                        one IMAD
// where the table's high words are stored pre-compensated, T'hi[j] = hi(2^(j/256)) - (j << 12),
// so that adding n << 12 = (k << 20) + (j << 12) inserts the binary exponent k without masking.
// `tab_bytes` is the table base in shared memory, `lane_bits` = (lane % 16) * 8.
__device__ __forceinline__ double exp_neg_tile(double z, const unsigned char* __restrict__ tab_bytes,
                                               unsigned lane_bits) {
  constexpr double kMagic = 6755399441055744.0;     // 1.5 * 2^52
  constexpr double kScale = -369.3299304675746271;  // -256 / ln 2
  constexpr double kStep = 0.0027076061740622863;   // ln 2 / 256
  constexpr int kHiClamp = 0x40862000;              // hi word of 708.0
  int hi = __double2hiint(z);
  hi = min(hi, kHiClamp);  // z <= 708 (sign bit set => negative int => untouched)
  z = __hiloint2double(hi, __double2loint(z));
  const double nd = fma(z, kScale, kMagic);
  const int ni = __double2loint(nd);
  const double n = nd - kMagic;
  const double rp = fma(n, kStep, z);  // = -r
  double p = fma(rp, 1.0 / 24.0, -1.0 / 6.0);
  p = fma(p, rp, 0.5);
  p = fma(p, rp, -1.0);
  const double q = p * rp;  // e^r - 1
  const unsigned off = (((unsigned)ni * 128u) & 0x7F80u) | lane_bits;
  const uint2 tv = *reinterpret_cast<const uint2*>(tab_bytes + off);
  int thi;
  asm("mad.lo.s32 %0, %1, 4096, %2;" : "=r"(thi) : "r"(ni), "r"((int)tv.y));
  const double t = __hiloint2double(thi, (int)tv.x);
  return fma(t, q, t);
}

__device__ __forceinline__ double exp_neg(double z, const double* __restrict__ tab) {
  constexpr double kMagic = 6755399441055744.0;            // 1.5 * 2^52
  constexpr double kScale = -369.3299304675746271;         // -256 / ln 2
  constexpr double kStep = 0.0027076061740622863;          // ln 2 / 256
  constexpr int kHiClamp = 0x40862000;                     // hi word of 708.0
#ifndef OAK_ABLATE
#define OAK_ABLATE 0  // development only: 1 no clamp, 2 no table load, 3 no table + no exponent
#endif
#if OAK_ABLATE != 1
  int hi = __double2hiint(z);
  hi = min(hi, kHiClamp);  // z <= 708 (sign bit set => negative int => untouched)
  z = __hiloint2double(hi, __double2loint(z));
#endif
  double nd = fma(z, kScale, kMagic);
  int ni = __double2loint(nd);
  double n = nd - kMagic;
  double rp = fma(n, kStep, z);  // = -r
  double p = fma(rp, 1.0 / 24.0, -1.0 / 6.0);
  p = fma(p, rp, 0.5);
  p = fma(p, rp, -1.0);
  double q = p * rp;             // e^r - 1
#if OAK_ABLATE == 2 || OAK_ABLATE == 3
  double t = tab[0];
#else
  double t = tab[(ni & (kExpTab - 1)) * 16];
#endif
#if OAK_ABLATE != 3
  int thi = __double2hiint(t) + ((ni >> 8) << 20);
  t = __hiloint2double(thi, __double2loint(t));
#endif
  return fma(t, q, t);
}
#endif

}  // namespace oak
