// Development microbenchmark: DMMA (mma.sync m8n8k4 f64) issue behaviour with FEW warps per SM sub-partition,
// as in the SGPR contraction kernel (2 CTAs x 4 warps per SM = 2 warps per scheduler).
//   DIST = number of independent accumulators a warp cycles through (dependent DMMAs are DIST issues apart)
//   warps per scheduler = 1 or 2 (blocks per SM), 128 threads per block
#include <cstdio>
#include <cuda_runtime.h>
constexpr int INNER = 8192;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int DIST>
__global__ void __launch_bounds__(128) k(double* out, const double* in, long long* cyc) {
  double c[2 * DIST];
  for (int i = 0; i < 2 * DIST; ++i) c[i] = in[i & 31];
  const double a = in[40] + (threadIdx.x & 3) * 1e-9, b = in[41];
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < INNER; ++it) {
#pragma unroll
    for (int r = 0; r < 32 / DIST; ++r)
#pragma unroll
      for (int i = 0; i < DIST; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 2 * DIST; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int DIST>
void run(double* out, double* in, long long* cyc, int sms, int per_sm) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * per_sm;
  k<DIST><<<blocks, 128>>>(out, in, cyc); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<DIST><<<blocks, 128>>>(out, in, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)blocks * 4 * 5 * INNER * 32;
  printf("dist %2d, %d warp(s)/scheduler: %.3f ms  %.2f TFLOP/s  | one warp: %.1f cycles per DMMA\n", DIST, per_sm, ms,
         n * 512 / (ms * 1e-3) / 1e12, (double)h / (INNER * 32.0));
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&in, 64 * 8); cudaMalloc(&cyc, 8);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5 + i * 1e-9;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int per_sm = 1; per_sm <= 4; per_sm *= 2) {
    run<1>(out, in, cyc, sms, per_sm);
    run<2>(out, in, cyc, sms, per_sm);
    run<4>(out, in, cyc, sms, per_sm);
    run<8>(out, in, cyc, sms, per_sm);
    run<16>(out, in, cyc, sms, per_sm);
  }
  return 0;
}
