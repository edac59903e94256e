// oak_spec_create / destroy: packs the hyper-parameters of an OAKKernel into a device-resident
// block.  Replaces the state read at call time by the reference
// (oak/oak_kernel.py:59-221, 251-265).
#include <cmath>
#include <cstring>
#include <mutex>

#include "oak_common.cuh"

namespace oak {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }

// 2^(j/kExpTab) table, one copy per device.
static std::mutex g_tab_mu;
static double* g_exptab[64] = {nullptr};

const double* exp_table_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_tab_mu);
  if (g_exptab[dev]) return g_exptab[dev];
  double h[kExpTab];
  for (int j = 0; j < kExpTab; ++j) h[j] = std::exp2((double)j / (double)kExpTab);
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
  g_exptab[dev] = d;
  return d;
}

// ---- var_s for the empirical measure: w^T K(z,z) w  (ortho_rbf_kernel.py:109-120) -------
// Deterministic two-pass reduction: one partial per block, summed in index order.
__global__ void empirical_var_partial(const double* __restrict__ loc, const double* __restrict__ w,
                                      int m, double inv_sqrt2_l, double s2,
                                      double* __restrict__ partial) {
  extern __shared__ double sh[];  // [2*tile] locations then weights, then reduction scratch
  const int tile = blockDim.x;
  double* sl = sh;
  double* sw = sh + tile;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double zi = (i < m) ? loc[i] * inv_sqrt2_l : 0.0;
  double wi = (i < m) ? w[i] : 0.0;
  double acc = 0.0;
  for (int base = 0; base < m; base += tile) {
    int j = base + threadIdx.x;
    sl[threadIdx.x] = (j < m) ? loc[j] * inv_sqrt2_l : 0.0;
    sw[threadIdx.x] = (j < m) ? w[j] : 0.0;
    __syncthreads();
    int lim = min(tile, m - base);
    for (int k = 0; k < lim; ++k) {
      double t = zi - sl[k];
      acc = fma(sw[k], exp(-t * t), acc);
    }
    __syncthreads();
  }
  acc *= wi * s2;
  // block tree reduction in a fixed order
  sl[threadIdx.x] = acc;
  __syncthreads();
  for (int s = tile / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sl[threadIdx.x] += sl[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sl[0];
}

__global__ void empirical_var_final(const double* __restrict__ partial, int nblocks,
                                    double* __restrict__ inv_sqrt_v_slot) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += partial[b];
    *inv_sqrt_v_slot = 1.0 / sqrt(v);
  }
}

static double var_s_closed_form(const oak_dim_desc& d) {
  const double l = d.lengthscale, s2 = d.variance;
  switch (d.measure) {
    case OAK_MEASURE_GAUSSIAN:  // ortho_rbf_kernel.py:94-97
      return s2 * l / std::sqrt(l * l + 2.0 * d.m1);
    case OAK_MEASURE_UNIFORM: {  // ortho_rbf_kernel.py:65-78
      const double a = d.m0, b = d.m1;
      const double y = (b - a) / std::sqrt(2.0) / l;
      return 2.0 / ((b - a) * (b - a)) * s2 * l * l *
             (std::sqrt(M_PI) * y * std::erf(y) + std::exp(-y * y) - 1.0);
    }
    case OAK_MEASURE_MOG: {  // ortho_rbf_kernel.py:138-152
      double acc = 0.0;
      for (int i = 0; i < d.count; ++i) {
        double row = 0.0;
        for (int j = 0; j < d.count; ++j) {
          const double dist = (d.v0[i] - d.v0[j]) * (d.v0[i] - d.v0[j]);
          const double scale = l * l + d.v1[i] + d.v1[j];
          row += s2 * l / std::sqrt(scale) * std::exp(-0.5 * dist / scale) * d.v2[j];
        }
        acc += d.v2[i] * row;
      }
      return acc;
    }
    default:
      return 0.0;
  }
}

// B table and its diagonal for the discrete kernels.
static void discrete_table(const oak_dim_desc& d, std::vector<double>& out) {
  if (d.type == OAK_DIM_BINARY) {  // ortho_binary_kernel.py:29-38
    const double p0 = d.m0, p1 = 1.0 - d.m0, v = d.variance;
    const double t[4] = {p1 * p1 * v, -p0 * p1 * v, -p0 * p1 * v, p0 * p0 * v};
    out.insert(out.end(), t, t + 4);
    out.push_back(p1 * p1 * v);
    out.push_back(p0 * p0 * v);
    return;
  }
  // ortho_categorical_kernel.py:34-53
  const int C = d.count, R = d.rank;
  std::vector<double> A((size_t)C * C), Ap(C, 0.0);
  for (int i = 0; i < C; ++i)
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
      for (int r = 0; r < R; ++r) s += d.v0[i * R + r] * d.v0[j * R + r];
      A[(size_t)i * C + j] = s + (i == j ? d.v1[i] : 0.0);
    }
  double pAp = 0.0;
  for (int i = 0; i < C; ++i) {
    for (int j = 0; j < C; ++j) Ap[i] += A[(size_t)i * C + j] * d.v2[j];
  }
  for (int i = 0; i < C; ++i) pAp += d.v2[i] * Ap[i];
  for (int i = 0; i < C; ++i)
    for (int j = 0; j < C; ++j)
      out.push_back((A[(size_t)i * C + j] - Ap[i] * Ap[j] / pAp) * d.variance);
  for (int i = 0; i < C; ++i) {
    double adiag = d.v1[i];
    for (int r = 0; r < R; ++r) adiag += d.v0[i * R + r] * d.v0[i * R + r];
    out.push_back((adiag - Ap[i] * Ap[i] / pAp) * d.variance);
  }
}

}  // namespace oak

using namespace oak;

extern "C" const char* oak_last_error(void) { return t_error.c_str(); }
extern "C" int oak_version(void) { return 100; }
extern "C" int64_t oak_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int oak_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ---- parameter arena pool -----------------------------------------------------------------------------
// A spec is rebuilt for every objective evaluation (the hyper-parameters are gpflow Parameters read at call
// time).  Seven cudaMalloc + pageable copies + seven cudaFree per evaluation cost ~0.17 ms, and every cudaFree
// synchronises the device in the middle of the evaluation.  Instead: one device block + one pinned staging
// block per spec, recycled through a per-device free list; one asynchronous copy; nothing is freed.
// Reuse is ordered by events: the copy must have consumed the staging block, and the kernels of the previous
// owner (on the stream it was created on) must have finished before the new copy lands.
namespace oak {
struct SpecArena {
  char* dev = nullptr;
  char* host = nullptr;
  size_t cap = 0;
  int device = 0;
  cudaEvent_t copied = nullptr;    // the H2D copy of the last owner has completed
  cudaEvent_t released = nullptr;  // recorded on the owner's stream when the spec was destroyed
  bool has_released = false;
};
static std::mutex g_arena_mu;
static std::vector<SpecArena*> g_arena_free;

static SpecArena* arena_acquire(size_t need, int device, cudaStream_t stream) {
  SpecArena* a = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_arena_mu);
    for (size_t i = 0; i < g_arena_free.size(); ++i)
      if (g_arena_free[i]->device == device && g_arena_free[i]->cap >= need) {
        a = g_arena_free[i];
        g_arena_free.erase(g_arena_free.begin() + i);
        break;
      }
  }
  if (!a) {
    a = new SpecArena();
    a->device = device;
    a->cap = (need + 65535) / 65536 * 65536;
    if (cudaMalloc(&a->dev, a->cap) != cudaSuccess || cudaMallocHost(&a->host, a->cap) != cudaSuccess ||
        cudaEventCreateWithFlags(&a->copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&a->released, cudaEventDisableTiming) != cudaSuccess) {
      if (a->dev) cudaFree(a->dev);
      if (a->host) cudaFreeHost(a->host);
      delete a;
      return nullptr;
    }
    return a;
  }
  cudaEventSynchronize(a->copied);  // staging block free again (normally long done)
  if (a->has_released) cudaStreamWaitEvent(stream, a->released, 0);
  return a;
}

static void arena_release(SpecArena* a, cudaStream_t stream) {
  if (!a) return;
  a->has_released = cudaEventRecord(a->released, stream) == cudaSuccess;
  std::lock_guard<std::mutex> lock(g_arena_mu);
  g_arena_free.push_back(a);
}
}  // namespace oak

extern "C" int oak_spec_destroy(oak_spec* spec) {
  if (!spec) return 0;
  arena_release((SpecArena*)spec->arena, (cudaStream_t)spec->arena_stream);
  delete spec;
  return 0;
}

extern "C" int oak_spec_create(const oak_kernel_desc* desc, void* stream_, oak_spec** out) {
  OAK_REQUIRE(desc && out, "oak_spec_create: null argument");
  OAK_REQUIRE(desc->num_dims >= 1, "oak_spec_create: num_dims must be >= 1");
  OAK_REQUIRE(desc->depth >= 0 && desc->depth <= OAK_MAX_DEPTH,
              "oak_spec_create: max_interaction_depth outside [0, OAK_MAX_DEPTH]");
  OAK_REQUIRE(desc->variances && desc->dims, "oak_spec_create: null parameter arrays");
  OAK_REQUIRE(oak_device_count() > 0,
              "oak_spec_create: no CUDA device visible (this library has no CPU fallback)");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int D = desc->num_dims;

  oak_spec* s = new oak_spec();
  s->D = D;
  s->depth = desc->depth;
  s->share_var = desc->share_var_across_orders;
  s->algo = desc->esp_algorithm;
  OAK_CUDA(cudaGetDevice(&s->device));
  // oak_kernel.py:255-265: sum_n sigma2_n e_n, or sigma2_0 e_0 + sum_{n>=1} e_n
  for (int n = 0; n <= OAK_MAX_DEPTH; ++n) s->sigma2[n] = 0.0;
  s->sigma2[0] = desc->variances[0];
  for (int n = 1; n <= desc->depth; ++n)
    s->sigma2[n] = desc->share_var_across_orders ? desc->variances[n] : 1.0;

  // kernel order: RBF dims first, discrete after
  std::vector<int> order;
  for (int d = 0; d < D; ++d)
    if (desc->dims[d].type == OAK_DIM_RBF) order.push_back(d);
  s->Dc = (int)order.size();
  for (int d = 0; d < D; ++d)
    if (desc->dims[d].type != OAK_DIM_RBF) order.push_back(d);
  s->Dd = D - s->Dc;
  s->pos_of_orig.assign(D, 0);
  s->sobol_off.assign(D, 0);

  // blob of per-measure arrays
  std::vector<double> blob;
  std::vector<size_t> off0(D, 0), off1(D, 0), off2(D, 0);
  std::vector<double> isv(D, 0.0), nls(D, 0.0);
  s->h_dims.resize(D);
  for (int k = 0; k < D; ++k) {
    const oak_dim_desc& d = desc->dims[order[k]];
    DimDev& dd = s->h_dims[k];
    std::memset(&dd, 0, sizeof(dd));
    s->pos_of_orig[order[k]] = k;
    dd.type = d.type;
    dd.column = d.column;
    dd.measure = d.measure;
    dd.count = d.count;
    dd.orig = order[k];
    dd.s2 = d.variance;
    dd.lengthscale = d.lengthscale;
    if (d.type == OAK_DIM_RBF) {
      if (!(d.lengthscale > 0.0) || !(d.variance > 0.0)) {
        delete s;
        OAK_REQUIRE(false, "oak_spec_create: RBF lengthscale and variance must be positive");
      }
      const double l = d.lengthscale, s2 = d.variance;
      dd.inv_sqrt2_l = 1.0 / (std::sqrt(2.0) * l);
      dd.xscale = kXScale / (std::sqrt(2.0) * l);
      dd.neg_log_s2 = -std::log(s2);
      nls[k] = dd.neg_log_s2;
      switch (d.measure) {
        case OAK_MEASURE_NONE:
          break;
        case OAK_MEASURE_GAUSSIAN:  // ortho_rbf_kernel.py:82-92
          dd.c0 = s2 * l / std::sqrt(l * l + d.m1);
          dd.c1 = d.m0;
          dd.c2 = 0.5 / (l * l + d.m1);
          isv[k] = 1.0 / std::sqrt(var_s_closed_form(d));
          break;
        case OAK_MEASURE_UNIFORM:  // ortho_rbf_kernel.py:49-63
          dd.c0 = s2 * l / (d.m1 - d.m0) * std::sqrt(M_PI / 2.0);
          dd.c1 = d.m0;
          dd.c2 = d.m1;
          isv[k] = 1.0 / std::sqrt(var_s_closed_form(d));
          break;
        case OAK_MEASURE_MOG:  // ortho_rbf_kernel.py:124-136
          if (d.count < 1 || !d.v0 || !d.v1 || !d.v2) {
            delete s;
            OAK_REQUIRE(false, "oak_spec_create: MOG measure needs means/variances/weights");
          }
          dd.c0 = s2 * l;
          isv[k] = 1.0 / std::sqrt(var_s_closed_form(d));
          off0[k] = blob.size();
          blob.insert(blob.end(), d.v0, d.v0 + d.count);
          off1[k] = blob.size();
          blob.insert(blob.end(), d.v1, d.v1 + d.count);
          off2[k] = blob.size();
          blob.insert(blob.end(), d.v2, d.v2 + d.count);
          break;
        case OAK_MEASURE_EMPIRICAL:  // ortho_rbf_kernel.py:101-120
          if (d.count < 1 || !d.v0 || !d.v1) {
            delete s;
            OAK_REQUIRE(false, "oak_spec_create: empirical measure needs locations/weights");
          }
          dd.c0 = s2;
          off0[k] = blob.size();
          blob.insert(blob.end(), d.v0, d.v0 + d.count);
          off1[k] = blob.size();
          blob.insert(blob.end(), d.v1, d.v1 + d.count);
          break;
        default:
          delete s;
          OAK_REQUIRE(false, "oak_spec_create: unknown measure kind");  // ortho_rbf_kernel.py:36-45
      }
    } else if (d.type == OAK_DIM_BINARY || d.type == OAK_DIM_CATEGORICAL) {
      if (d.type == OAK_DIM_BINARY) dd.count = 2;
      if (d.type == OAK_DIM_CATEGORICAL &&
          (d.count < 1 || d.rank < 1 || !d.v0 || !d.v1 || !d.v2)) {
        delete s;
        OAK_REQUIRE(false, "oak_spec_create: categorical kernel needs W, kappa and p");
      }
      dd.table_off = (int)s->h_tables.size();
      discrete_table(d, s->h_tables);
      dd.c0 = d.m0;  // binary: p0 (used by the Sobol L formula, oak/utils.py:266-269)
      s->sobol_off[k] = (int)s->h_sobolG.size();
      if (d.type == OAK_DIM_CATEGORICAL) {
        // G = B1 diag(p) B1^T with the unit-variance table B1: L[i,j] = G[x_i, x_j]
        // (compute_L_categorical_kernel, oak/utils.py:292-307)
        const int Cn = d.count;
        const double* B = s->h_tables.data() + dd.table_off;
        const double inv_var = 1.0 / d.variance;
        for (int a = 0; a < Cn; ++a)
          for (int b = 0; b < Cn; ++b) {
            double g = 0.0;
            for (int c = 0; c < Cn; ++c)
              g += (B[a * Cn + c] * inv_var) * d.v2[c] * (B[b * Cn + c] * inv_var);
            s->h_sobolG.push_back(g);
          }
      }
      // the per-dim auxiliary double carries the table offset (bit pattern) for discrete dims
      const int64_t off_bits = dd.table_off;
      std::memcpy(&nls[k], &off_bits, sizeof(double));
    } else {
      delete s;
      OAK_REQUIRE(false, "oak_spec_create: unknown sub-kernel type");
    }
  }
  s->tables_len = (int)s->h_tables.size();

  // upload
  auto fail = [&](const char* what) {
    set_error(std::string("oak_spec_create: ") + what + ": " +
              cudaGetErrorString(cudaGetLastError()));
    oak_spec_destroy(s);
    return 1;
  };
  s->h_blob = blob;
  s->blob_off0 = off0;
  s->blob_off1 = off1;
  s->blob_off2 = off2;
  // Gram-tile flavour: the exponent of the RBF dims in units of ln2/256 (0 when s^2 == 1)
  std::vector<double> gaux(nls);
  for (int k = 0; k < s->Dc; ++k) gaux[k] = (s->h_dims[k].s2 == 1.0) ? 0.0 : nls[k] * kXScale2;
  // one arena: blob | dims | inv_sqrt_v | neg_log_s2 | gram_aux | tables | sobolG, each 256-byte aligned
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_blob = 0;
  const size_t o_dims = o_blob + up(blob.size() * sizeof(double));
  const size_t o_isv = o_dims + up(D * sizeof(DimDev));
  const size_t o_nls = o_isv + up(D * sizeof(double));
  const size_t o_gaux = o_nls + up(D * sizeof(double));
  const size_t o_tab = o_gaux + up(D * sizeof(double));
  const size_t o_sob = o_tab + up((size_t)s->tables_len * sizeof(double));
  const size_t total = o_sob + up(s->h_sobolG.size() * sizeof(double)) + 256;
  SpecArena* arena = arena_acquire(total, s->device, stream);
  if (!arena) return fail("arena allocation");
  s->arena = arena;
  s->arena_stream = stream;
  s->d_blob = blob.empty() ? nullptr : (double*)(arena->dev + o_blob);
  s->d_dims = (DimDev*)(arena->dev + o_dims);
  s->d_inv_sqrt_v = (double*)(arena->dev + o_isv);
  s->d_neg_log_s2 = (double*)(arena->dev + o_nls);
  s->d_gram_aux = (double*)(arena->dev + o_gaux);
  s->d_tables = s->tables_len > 0 ? (double*)(arena->dev + o_tab) : nullptr;
  s->d_sobolG = s->h_sobolG.empty() ? nullptr : (double*)(arena->dev + o_sob);
  for (int k = 0; k < D; ++k) {
    DimDev& dd = s->h_dims[k];
    if (dd.type == OAK_DIM_RBF &&
        (dd.measure == OAK_MEASURE_MOG || dd.measure == OAK_MEASURE_EMPIRICAL)) {
      dd.v0 = s->d_blob + off0[k];
      dd.v1 = s->d_blob + off1[k];
      dd.v2 = dd.measure == OAK_MEASURE_MOG ? s->d_blob + off2[k] : nullptr;
    }
  }
  if (!blob.empty()) std::memcpy(arena->host + o_blob, blob.data(), blob.size() * sizeof(double));
  std::memcpy(arena->host + o_dims, s->h_dims.data(), D * sizeof(DimDev));
  std::memcpy(arena->host + o_isv, isv.data(), D * sizeof(double));
  std::memcpy(arena->host + o_nls, nls.data(), D * sizeof(double));
  std::memcpy(arena->host + o_gaux, gaux.data(), D * sizeof(double));
  if (s->tables_len > 0) std::memcpy(arena->host + o_tab, s->h_tables.data(), (size_t)s->tables_len * sizeof(double));
  if (!s->h_sobolG.empty())
    std::memcpy(arena->host + o_sob, s->h_sobolG.data(), s->h_sobolG.size() * sizeof(double));
  if (cudaMemcpyAsync(arena->dev, arena->host, total - 256, cudaMemcpyHostToDevice, stream) != cudaSuccess)
    return fail("cudaMemcpyAsync");
  if (cudaEventRecord(arena->copied, stream) != cudaSuccess) return fail("cudaEventRecord");
  s->d_exptab = exp_table_device();
  if (!s->d_exptab) return fail("exp table upload");

  // var_s of the empirical dims on the device
  for (int k = 0; k < s->Dc; ++k) {
    const DimDev& dd = s->h_dims[k];
    if (dd.measure != OAK_MEASURE_EMPIRICAL) continue;
    const int threads = 256;
    const int blocks = (dd.count + threads - 1) / threads;
    double* partial = nullptr;
    if (cudaMallocAsync(&partial, blocks * sizeof(double), stream) != cudaSuccess)
      return fail("cudaMallocAsync");
    empirical_var_partial<<<blocks, threads, 2 * threads * sizeof(double), stream>>>(
        dd.v0, dd.v1, dd.count, dd.inv_sqrt2_l, dd.s2, partial);
    g_launches.fetch_add(1);
    empirical_var_final<<<1, 32, 0, stream>>>(partial, blocks, s->d_inv_sqrt_v + k);
    g_launches.fetch_add(1);
    cudaFreeAsync(partial, stream);
    if (cudaGetLastError() != cudaSuccess) return fail("empirical var_s kernels");
  }
  *out = s;
  return 0;
}

extern "C" int oak_spec_var_s_f64(const oak_spec* spec, int32_t dim, double* h_var_s, void* stream_) {
  OAK_REQUIRE(spec && h_var_s, "oak_spec_var_s_f64: null argument");
  OAK_REQUIRE(dim >= 0 && dim < spec->D, "oak_spec_var_s_f64: dim out of range");
  const int pos = spec->pos_of_orig[dim];
  OAK_REQUIRE(pos < spec->Dc && spec->h_dims[pos].measure != OAK_MEASURE_NONE,
              "oak_spec_var_s_f64: not a constrained RBF dimension");
  cudaStream_t stream = (cudaStream_t)stream_;
  double isv = 0.0;
  OAK_CUDA(cudaMemcpyAsync(&isv, spec->d_inv_sqrt_v + pos, sizeof(double), cudaMemcpyDeviceToHost,
                           stream));
  OAK_CUDA(cudaStreamSynchronize(stream));
  *h_var_s = 1.0 / (isv * isv);
  return 0;
}
