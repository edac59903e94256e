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
  double xscale;      // sqrt(256 / ln 2) / (sqrt(2) l): factor of the prepared coordinate
  double neg_log_s2;  // -ln(s^2)
  double s2;          // base variance
  double lengthscale;
  double c0, c1, c2;  // measure constants (see oak_prepare.cu)
  const double* v0;   // device copies of the per-measure arrays
  const double* v1;
  const double* v2;
};

constexpr int kPointPad = 128;  // prepared-point rows are padded to a multiple of this
#ifndef OAK_EXP_BITS
#define OAK_EXP_BITS 10  // 8: 256-entry table + degree-4 Taylor; 9 / 10: 512 / 1024 entries + degree-3 near-minimax
#endif
constexpr int kExpBits = OAK_EXP_BITS;
constexpr int kExpTab = 1 << kExpBits;  // 2^(j/kExpTab) table entries

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
  double* d_neg_log_s2 = nullptr;      // [D] RBF: -ln s^2 ; discrete: bits(table offset)
  double* d_gram_aux = nullptr;        // [D] RBF: -ln(s^2) 256/ln2 ; discrete: bits(table offset)
  double* d_tables = nullptr;          // discrete tables blob
  int tables_len = 0;
  double* d_blob = nullptr;            // per-measure arrays
  std::vector<double> h_blob;          // host copy (backward pass: closed-form d var_s / dl of MOG measures)
  std::vector<size_t> blob_off0, blob_off1, blob_off2;  // [D] offsets of v0 / v1 / v2 per kernel-order dim
  const double* d_exptab = nullptr;    // 2^(j/256)
  void* arena = nullptr;               // pooled device + pinned staging block holding every d_* array above
  void* arena_stream = nullptr;        // stream the spec was created on (orders the reuse of the arena)
};

namespace oak {
const double* exp_table_device();  // lazily uploaded per device; nullptr on failure

// Internal launchers shared between translation units (device pointers; see oak_gram.cu).
int tile_rows_for_depth(int depth);
// `prow` / `pcol` are the BASES of two prepared-point blocks (their min/max keys follow the
// coordinates, see points_minmax); row/column ranges select the part to evaluate.
int gram_launch(const oak_spec* spec, const double2* prow, int64_t n_row_pad, int64_t row_begin,
                int64_t row_end, const double2* pcol, int64_t n_col_pad, int64_t col_begin,
                int64_t col_end, int mode, double* K, int64_t ldk, cudaStream_t stream, double* Kt = nullptr,
                int64_t ldkt = 0, const double* ydot_y = nullptr, double* ydot_part = nullptr, int sm_reserve = 0);
// sm_reserve > 0: the persistent grid leaves that many SMs free (for a small kernel running next to the tiles);
// sm_reserve < 0: the grid is capped at -sm_reserve CTAs (this launch IS the small kernel)
// the folded K y of the general-mode tiles (depth <= 4): ydot_part holds ceil(cols / 64) x rows partial sums
inline bool gram_can_fold_ky(const oak_spec* spec) { return spec->depth <= 4; }
// per-dimension min / max keys of the prepared coordinate: [D] min keys then [D] max keys
inline const unsigned long long* points_minmax(const oak_spec* spec, const double2* pts, int64_t n_pad) {
  return reinterpret_cast<const unsigned long long*>(pts + (int64_t)spec->D * n_pad);
}
// Phi(lower, column-major) += A A^T on the FP64 tensor cores (oak_syrk.cu)
size_t syrk_dmma_work_bytes(int m);
int syrk_lower_dmma(int m, int64_t k, double* A, int64_t lda, double* C, double* work, size_t work_bytes,
                    int device, cudaStream_t stream, double* A_alt = nullptr, const int* d_route = nullptr);
// second-generation contraction (oak_syrk2.cu): 128 x 128 tiles, weighted stream-K, A y folded in
size_t syrk2_work_bytes(int m, int device);
int syrk2_lower_dmma(int m, int64_t k, double* A, int64_t lda, double* C, const double* y, double* kufy, double* work,
                     size_t work_bytes, int device, cudaStream_t stream, double* A_alt = nullptr,
                     const int* d_route = nullptr);
// C = T B (+ u v^T) on the FP64 tensor cores (oak_pgemm.cu); d_gate: optional device flag, 0 = no-op
int panel_gemm_dmma(const double* T, int64_t ldt, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int Kd,
                    int64_t n, int lower, const double* u, const double* v, const int* d_gate, int* d_counter,
                    int device, cudaStream_t stream);
// blocked Cholesky with border rows, one cooperative launch (oak_chol.cu)
// max_ctas > 0 caps the grid (the factorisation then shares the GPU with another kernel)
int chol_bordered(double* A, int64_t ld, int n, int rows, int gap, int border_identity, int* d_info,
                  double* d_logdet, int device, cudaStream_t stream, int max_ctas = 0);
int gram_diag_launch(const oak_spec* spec, const double2* pts, int64_t n, int64_t n_pad,
                     double* out, cudaStream_t stream);

// ---- fast FP64 exp(-z) on pre-scaled distances ---------------------------------------
// T = kExpTab table entries of 2^(j/T).  Prepared RBF coordinates carry the factor sqrt(T / ln 2):
// with d = a_i - b_j,
//   d^2 = z * T / ln 2,  z = (x - y)^2 / (2 l^2)
// so that n = -round(d^2) is the table/exponent index and w = d^2 + n (|w| <= 1/2, EXACT: both
// products are fused) gives the reduced argument r = -w ln2/T:
//   exp(-z) = 2^(n >> bits) * T[n & (T-1)] * e^r,   e^r - 1 = w * poly(w)
// Default: T = 1024 with a degree-3 near-minimax polynomial (max error 9.5e-17 relative, measured
// max Gram error vs the oracle 6e-16): 8 FP64-pipe instructions from d.  The table is replicated
// 8x in shared memory (64 KB).  Alternatives kept for A/B runs (profiles/): T = 512, 16 replicas
// (1.5e-15, 0.4 % faster); T = 256 with the degree-4 Taylor polynomial (3.8e-17, one FP64
// instruction more: 0.933 vs 0.99 of the roofline on config B).
// Fast form: no clamp; valid for |d| <= kFastSpan (z <= 704), which the launcher proves per
// dimension from the min/max of the prepared coordinates.
// General form (s^2 != 1 or unbounded distance): zs = d^2 - ln(s^2) T/ln2, clamped at z ~ 707.
//
// Measured on B200 (scripts/ubench, profiles/): an FP64 instruction holds the issue port for two
// cycles and every other instruction for one, without overlap; integer-ALU forms (LOP3/SHF/VIMNMX)
// are the expensive ones, IMAD (FMA pipe) the cheap one.  The index math is therefore
//   off = mulhi(n << (32-bits), 8R 2^bits) + lane_bits = (n & (T-1)) * 8R + lane_bits   IMAD.SHL + LEA.HI
//   hi  = n * 2^(20-bits) + T'hi[j]                                                      IMAD
// with the table's high words stored pre-compensated, T'hi[j] = hi(2^(j/T)) - (j << (20-bits)), so
// that adding n << (20-bits) = (k << 20) + (j << (20-bits)) inserts the binary exponent k without
// masking.  `tab_bytes` is the table base in shared memory (entry j, replica lane % R at
// j*8R + (lane%R)*8; R = 16: a half-warp never bank-conflicts, R = 8: two-way at worst),
// `lane_bits` = (lane % R) * 8.
#if OAK_EXP_BITS == 8
constexpr double kXScale = 19.217958540583197;      // sqrt(256 / ln 2)
constexpr double kXScale2 = 369.3299304675746271;   // 256 / ln 2
constexpr double kInvXScale2 = 0.0027076061740622863;  // ln 2 / 256
constexpr double kFastSpan = 510.0;                 // |d| bound of the clamp-free form (z <= 704)
constexpr int kHiClampScaled = 0x410fe000;          // hi word of 261120.0 (z ~ 707)
#elif OAK_EXP_BITS == 9
constexpr double kXScale = 27.17829760921661;       // sqrt(512 / ln 2)
constexpr double kXScale2 = 738.6598609351493;      // 512 / ln 2
constexpr double kInvXScale2 = 0.0013538030870311431;  // ln 2 / 512
constexpr double kFastSpan = 721.0;                 // d^2 <= 519841 (z <= 703.8)
constexpr int kHiClampScaled = 0x411fe000;          // hi word of 522240.0 (z ~ 707)
#elif OAK_EXP_BITS == 10
constexpr double kXScale = 38.435917081166394;      // sqrt(1024 / ln 2)
constexpr double kXScale2 = 1477.3197218702985;     // 1024 / ln 2
constexpr double kInvXScale2 = 0.0006769015435155716;  // ln 2 / 1024
constexpr double kFastSpan = 1019.0;                // d^2 <= 1038361 (z <= 702.9)
constexpr int kHiClampScaled = 0x412fe000;          // hi word of 1044480.0 (z ~ 707)
#else
#error "OAK_EXP_BITS must be 8, 9 or 10"
#endif
#ifndef OAK_EXP_REPL
#define OAK_EXP_REPL 8  // shared-memory replicas of the table (entry stride = replicas * 8 bytes)
#endif
constexpr int kExpRepl = OAK_EXP_REPL;
constexpr int kExpReplLog2 = kExpRepl == 16 ? 4 : (kExpRepl == 8 ? 3 : -1);
static_assert(kExpReplLog2 > 0, "OAK_EXP_REPL must be 8 or 16");
#ifdef __CUDACC__
__device__ __forceinline__ double exp_tail(double w, int ni, const unsigned char* __restrict__ tab_bytes,
                                           unsigned lane_bits) {
#if OAK_EXP_BITS == 8
  // degree-4 Taylor in r = -w ln2/256: truncation 3.8e-17
  constexpr double C1 = -0.0027076061740622863;
  constexpr double C2 = 3.6655655969101062e-06;
  constexpr double C3 = -3.3083026805413713e-09;
  constexpr double C4 = 2.239395190875157e-12;
  double p = fma(w, C4, C3);
  p = fma(p, w, C2);
  p = fma(p, w, C1);
#elif OAK_EXP_BITS == 9
  // degree-3 near-minimax (Chebyshev least squares on |w| <= 1/2) in r = -w ln2/512: max error
  // 1.5e-15 relative (Taylor would give 8.8e-15); one FP64 instruction fewer per entry
  constexpr double C1 = -0.0013538030870311425;
  constexpr double C2 = 9.163914283863184e-07;
  constexpr double C3 = -4.1353784691025013e-10;
  double p = fma(w, C3, C2);
  p = fma(p, w, C1);
#else
  // degree-3 near-minimax in r = -w ln2/1024: max error 9.5e-17 relative
  constexpr double C1 = -0.0006769015435155716;
  constexpr double C2 = 2.290978516293061e-07;
  constexpr double C3 = -5.1692229753539507e-11;
  double p = fma(w, C3, C2);
  p = fma(p, w, C1);
#endif
  const double q = p * w;  // e^r - 1
#ifndef OAK_IDX_MODE
#define OAK_IDX_MODE 0  // development switch, see scripts/ubench/fp64_mix3.cu
#endif
  constexpr unsigned kShl = 1u << (32 - kExpBits);  // n << (32 - bits): the table index in the top bits
  constexpr unsigned kMulHi = 1u << (kExpBits + 3 + kExpReplLog2);  // ... byte offset j * (replicas * 8)
  constexpr int kExpIns = 1 << (20 - kExpBits);     // n * 2^(20 - bits) = (k << 20) + (j << (20 - bits))
  unsigned off;
#if OAK_IDX_MODE == 0    // IMAD.SHL + LEA.HI
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(off) : "r"((unsigned)ni * kShl), "n"(kMulHi), "r"(lane_bits));
#elif OAK_IDX_MODE == 1  // IMAD.SHL + IMAD.HI.U32 (multiplier kept opaque in a register)
  unsigned m15;
  asm volatile("mov.u32 %0, %1;" : "=r"(m15) : "n"(kMulHi));
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(off) : "r"((unsigned)ni * kShl), "r"(m15), "r"(lane_bits));
#elif OAK_IDX_MODE == 2  // IMAD.SHL + IMAD.WIDE.U32, high word
  unsigned m15;
  asm volatile("mov.u32 %0, %1;" : "=r"(m15) : "n"(kMulHi));
  unsigned long long wide;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(wide) : "r"((unsigned)ni * kShl), "r"(m15),
      "l"((unsigned long long)lane_bits << 32));
  off = (unsigned)(wide >> 32);
#else                    // IMAD.SHL + LOP3 (round-1 form)
  off = (((unsigned)ni * (8u * kExpRepl)) & (unsigned)((kExpTab - 1) << (3 + kExpReplLog2))) | lane_bits;
#endif
  const uint2 tv = *reinterpret_cast<const uint2*>(tab_bytes + off);
  int thi;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(thi) : "r"(ni), "n"(kExpIns), "r"((int)tv.y));
  const double t = __hiloint2double(thi, (int)tv.x);
  return fma(t, q, t);
}

// exp(-d^2 ln2/256) for |d| <= kFastSpan
__device__ __forceinline__ double exp_neg_sq_fast(double d, const unsigned char* __restrict__ tab_bytes,
                                                  unsigned lane_bits) {
  constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double nd = fma(-d, d, kMagic);
  const double n = nd - kMagic;   // -round(d^2)
  const double w = fma(d, d, n);  // exact
  return exp_tail(w, __double2loint(nd), tab_bytes, lane_bits);
}

// exp(-zs ln2/256), zs = d^2 + aux, clamped at zs <= ~261121 (z ~ 707)
__device__ __forceinline__ double exp_neg_scaled(double zs, const unsigned char* __restrict__ tab_bytes,
                                                 unsigned lane_bits) {
  constexpr double kMagic = 6755399441055744.0;
  constexpr int kHiClamp = kHiClampScaled;
  int hi = __double2hiint(zs);
  hi = min(hi, kHiClamp);  // (sign bit set => negative int => untouched)
  zs = __hiloint2double(hi, __double2loint(zs));
  const double nd = kMagic - zs;
  const double n = nd - kMagic;
  const double w = zs + n;
  return exp_tail(w, __double2loint(nd), tab_bytes, lane_bits);
}

// ---- cp.async (LDGSTS) staging helpers ------------------------------------------------------
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

// monotone uint64 keys for atomicMin / atomicMax on doubles
__device__ __forceinline__ unsigned long long order_key(double x) {
  const long long b = __double_as_longlong(x);
  return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ULL));
}
__device__ __forceinline__ double order_key_decode(unsigned long long k) {
  const long long b = (k & 0x8000000000000000ULL) ? (long long)(k ^ 0x8000000000000000ULL) : (long long)~k;
  return __longlong_as_double(b);
}
#endif

}  // namespace oak
