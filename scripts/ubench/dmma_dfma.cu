// Development microbenchmark: is DMMA (mma.sync m8n8k4 f64) a pipe of its own on sm_100a, i.e. can
// it run concurrently with DFMA?  Kernels: DFMA only, DMMA only, both from the same warp, and the
// two kinds on different warps of the same SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int INNER = 4096;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: 16 DFMA / iter, 1: 8 DMMA / iter, 2: both in every warp, 3: even warps DFMA, odd warps DMMA
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, const double* in) {
  double f[16], c[16];
  for (int i = 0; i < 16; ++i) { f[i] = in[i] + threadIdx.x * 1e-9; c[i] = in[16 + i]; }
  const double a = in[40] + (threadIdx.x & 3) * 1e-9, b = in[41], bs = in[42];
  const int w = threadIdx.x >> 5;
  const bool do_f = MODE == 0 || MODE == 2 || (MODE == 3 && (w & 1) == 0);
  const bool do_m = MODE == 1 || MODE == 2 || (MODE == 3 && (w & 1) == 1);
#pragma unroll 1
  for (int it = 0; it < INNER; ++it) {
    if (do_f) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fma(f[i], f[i], bs);
    }
    if (do_m) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += f[i] + c[i];
  if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, double* out, double* in, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * 4;
  k<MODE><<<blocks, 256>>>(out, in); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 10; ++r) k<MODE><<<blocks, 256>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)blocks * 8 * 10 * INNER;
  double fw = (MODE == 0 || MODE == 2) ? warps : (MODE == 3 ? warps / 2 : 0);
  double mw = (MODE == 1 || MODE == 2) ? warps : (MODE == 3 ? warps / 2 : 0);
  const double dfma_flop = fw * 16 * 32 * 2, dmma_flop = mw * 8 * 256 * 2;
  printf("%-44s %.3f ms  DFMA %.2f TF + DMMA %.2f TF = %.2f TFLOP/s\n", name, ms / 10,
         dfma_flop / (ms * 1e-3) / 1e12, dmma_flop / (ms * 1e-3) / 1e12, (dfma_flop + dmma_flop) / (ms * 1e-3) / 1e12);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in; cudaMalloc(&out, 8); cudaMalloc(&in, 64 * 8);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5 + i * 1e-9; h[42] = 0.1; h[40] = 1e-3; h[41] = 1e-3;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("DFMA only (16 chains / warp)", out, in, sms);
  run<1>("DMMA only (8 accumulators / warp)", out, in, sms);
  run<2>("DFMA + DMMA in every warp", out, in, sms);
  run<3>("even warps DFMA, odd warps DMMA", out, in, sms);
  return 0;
}
