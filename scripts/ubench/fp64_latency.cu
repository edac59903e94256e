// Dependent-issue latencies that bound the in-register Cholesky panel (development aid):
// DFMA chain, rsqrt(double) chain, MUFU.RSQ64H alone, __syncthreads with 256 threads, shared-memory round trip.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a0, double b0) {
  const int N = 1024;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  long long t0, t1;
  __shared__ double sh[512];
  // 1. DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = fma(a, b, b);
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  // 2. rsqrt chain
  double r = fabs(a) + 1.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) r = rsqrt(r) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[1] = t1 - t0;
  // 3. MUFU.RSQ64H chain (approximation only)
  double q = r;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    double y;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    q = y + 1.5;
  }
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[2] = t1 - t0;
  // 4. barrier
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[3] = t1 - t0;
  // 5. STS -> barrier -> LDS (neighbour) round trip
  double v = q;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    sh[threadIdx.x] = v;
    __syncthreads();
    v = sh[(threadIdx.x + 33) & 255] + 1.0;
    __syncthreads();
  }
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[4] = t1 - t0;
  // 6. DADD chain, DMUL chain
  double s = v;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) s = s + b;
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[5] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) s = s * b;
  t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[6] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + r + q + v + s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 8); cudaMallocManaged(&cyc, 64);
  for (int threads : {32, 256}) {
    for (int rep = 0; rep < 2; ++rep) { lat<<<1, threads>>>(out, cyc, 0.3, 0.999); cudaDeviceSynchronize(); }
    printf("%3d threads/CTA: DFMA %.1f | rsqrt(double)+add %.1f | MUFU.RSQ64H+add %.1f | __syncthreads %.1f | STS-BAR-LDS-BAR %.1f | DADD %.1f | DMUL %.1f cycles per dependent step\n",
           threads, cyc[0] / 1024.0, cyc[1] / 1024.0, cyc[2] / 1024.0, cyc[3] / 1024.0, cyc[4] / 1024.0, cyc[5] / 1024.0, cyc[6] / 1024.0);
  }
  return 0;
}
