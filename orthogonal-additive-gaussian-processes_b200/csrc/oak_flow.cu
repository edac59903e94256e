// Input pipeline on the device: the per-column normalising flow of
// oak/normalising_flow.py.  The reference chains tfb.Shift(-offset) -> Log -> Shift -> Scale -> SinhArcsinh
// (:46-56; without the first two when log=False), SinhArcsinh as in its TFP 0.11 pin:
//     y = sinh((asinh(z) + skewness) * tailweight),   z = (u + shift) * scale,   u = log(x - offset) or x
// and fits (log scale, shift, skewness, log tailweight) by minimising KL_objective (:76-81)
//     J = mean(y^2 / 2) - mean(log |dy/dx|)
// with scipy L-BFGS-B (model_utils.py:313-317).  One evaluation of J and of its four derivatives is one
// pass over the column: a grid-stride kernel with a deterministic two-stage reduction.  The forward
// transform of a column is one elementwise kernel.
#include <cmath>

#include "oak_common.cuh"

namespace oak {

struct FlowParams {
  double offset, shift, scale, skewness, tailweight;
  int use_log;
};

__global__ void __launch_bounds__(256) flow_forward_kernel(const double* __restrict__ x, int64_t n, int64_t sx,
                                                           FlowParams p, double* __restrict__ y, int64_t sy) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double u = p.use_log ? log(x[i * sx] - p.offset) : x[i * sx];
  const double z = (u + p.shift) * p.scale;
  y[i * sy] = sinh((asinh(z) + p.skewness) * p.tailweight);
}

constexpr int kFlowSums = 5;

// partial[b][0..4] = sum over the block's points of
//   J_i = y^2/2 - ldj,  dJ/dz z,  dJ/dz,  dJ/dw,  dJ/dw w
// with w = (asinh z + skewness) tailweight, y = sinh w,
//   ldj  = log cosh w + log tailweight - log1p(z^2)/2 + log scale - (u when use_log)
//   dJ/dw = y cosh w - tanh w,   dJ/dz = dJ/dw tailweight / sqrt(1 + z^2) + z / (1 + z^2)
__global__ void __launch_bounds__(256) flow_objective_partial_kernel(const double* __restrict__ x, int64_t n, int64_t sx,
                                                                     FlowParams p, double log_scale, double log_tail,
                                                                     double* __restrict__ partial) {
  __shared__ double red[8][kFlowSums];
  double s[kFlowSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double u = p.use_log ? log(x[i * sx] - p.offset) : x[i * sx];
    const double z = (u + p.shift) * p.scale;
    const double w = (asinh(z) + p.skewness) * p.tailweight;
    const double y = sinh(w), ch = cosh(w);
    const double q = 1.0 + z * z;
    const double ldj = log(ch) + log_tail - 0.5 * log1p(z * z) + log_scale - (p.use_log ? u : 0.0);
    const double dw = y * ch - tanh(w);
    const double dz = dw * p.tailweight / sqrt(q) + z / q;
    s[0] += 0.5 * y * y - ldj;
    s[1] = fma(dz, z, s[1]);
    s[2] += dz;
    s[3] += dw;
    s[4] = fma(dw, w, s[4]);
  }
#pragma unroll
  for (int k = 0; k < kFlowSums; ++k) {
    double v = s[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kFlowSums) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partial[(int64_t)blockIdx.x * kFlowSums + threadIdx.x] = v;
  }
}

// out = [J, dJ/d log scale, dJ/d shift, dJ/d skewness, dJ/d log tailweight] (fixed-order sum over the blocks)
__global__ void flow_objective_finish_kernel(const double* __restrict__ partial, int blocks, double inv_n, double scale,
                                             double tailweight, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= kFlowSums) return;
  double v = 0.0;
  for (int b = 0; b < blocks; ++b) v += partial[(int64_t)b * kFlowSums + k];
  v *= inv_n;
  if (k == 1) v -= 1.0;             // d/d log scale: mean(dJ/dz z) - 1
  if (k == 2) v *= scale;           // d/d shift
  if (k == 3) v *= tailweight;      // d/d skewness
  if (k == 4) v -= 1.0;             // d/d log tailweight: mean(dJ/dw w) - 1
  out[k] = v;
}

static int flow_blocks(int64_t n) {
  const int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));  // 8 CTAs x 148 SMs
}

}  // namespace oak

using namespace oak;

extern "C" int oak_flow_forward_f64(const double* d_x, int64_t n, int64_t stride_in, double offset, int32_t use_log,
                                    double shift, double scale, double skewness, double tailweight, double* d_y,
                                    int64_t stride_out, void* stream_) {
  OAK_REQUIRE(d_x && d_y && stride_in >= 1 && stride_out >= 1, "oak_flow_forward_f64: bad argument");
  if (n <= 0) return 0;
  const FlowParams p{offset, shift, scale, skewness, tailweight, use_log};
  flow_forward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(d_x, n, stride_in, p, d_y,
                                                                                     stride_out);
  OAK_LAUNCHED();
  return 0;
}

extern "C" size_t oak_flow_objective_work_bytes(int64_t n) {
  return n < 0 ? 0 : (size_t)flow_blocks(n) * kFlowSums * sizeof(double);
}

extern "C" int oak_flow_objective_f64(const double* d_x, int64_t n, int64_t stride, double offset, int32_t use_log,
                                      double log_scale, double shift, double skewness, double log_tailweight,
                                      double* d_out5, void* d_work, void* stream_) {
  OAK_REQUIRE(d_x && d_out5 && d_work && stride >= 1 && n >= 1, "oak_flow_objective_f64: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const FlowParams p{offset, shift, std::exp(log_scale), skewness, std::exp(log_tailweight), use_log};
  const int blocks = flow_blocks(n);
  flow_objective_partial_kernel<<<blocks, 256, 0, stream>>>(d_x, n, stride, p, log_scale, log_tailweight,
                                                            (double*)d_work);
  OAK_LAUNCHED();
  flow_objective_finish_kernel<<<1, 32, 0, stream>>>((const double*)d_work, blocks, 1.0 / (double)n, p.scale,
                                                     p.tailweight, d_out5);
  OAK_LAUNCHED();
  return 0;
}
