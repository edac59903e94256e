// Development microbenchmark: FP64 issue rate vs operand pattern on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 16, INNER = 1024;

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, const double* in) {
  double a[CH], b[CH], c[CH];
  for (int i = 0; i < CH; ++i) { a[i] = in[i] + threadIdx.x * 1e-9; b[i] = in[16 + i]; c[i] = in[32 + i]; }
  const double bs = in[50], cs = in[51];
#pragma unroll 1
  for (int it = 0; it < INNER; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (MODE == 0) a[i] = fma(a[i], 0.999999999999, 1e-13);        // reg, const, const
        if (MODE == 1) a[i] = fma(a[i], bs, cs);                       // 3 regs, b/c shared by all chains
        if (MODE == 2) a[i] = fma(a[i], b[i], c[i]);                   // 3 regs, all distinct per chain
        if (MODE == 3) a[i] = a[i] + bs;                               // DADD 2 regs
        if (MODE == 4) a[i] = a[i] * bs;                               // DMUL 2 regs
        if (MODE == 5) a[i] = fma(a[i], a[i], cs);                     // 2 distinct
        if (MODE == 6) a[i] = fma(a[i], b[i], a[i]);                   // 2 distinct
        if (MODE == 7) a[i] = fma(b[i], c[i], a[i]);                   // accumulate form, 3 distinct
        if (MODE == 8) a[i] = fma(b[i], b[(i + 1) % CH], a[i]);        // accumulate, neighbours
      }
  }
  double s = 0;
  for (int i = 0; i < CH; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, double* out, double* in, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = sms * 8;
  k<MODE><<<blocks, 256>>>(out, in); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 10; ++r) k<MODE><<<blocks, 256>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double slots = (double)blocks * 256 * CH * INNER * 4 * 10;
  double rate = slots / (ms * 1e-3);
  printf("%-44s %.3e slots/s  = %.3f inst/cycle/SMSP @1.965GHz\n", name, rate, rate / 32 / (sms * 4) / 1.965e9);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in; cudaMalloc(&out, 8); cudaMalloc(&in, 64 * 8);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.9999999 + i * 1e-9; h[51] = 1e-13;
  for (int i = 32; i < 48; ++i) h[i] = 1e-13;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("fma(a, const, const)", out, in, sms);
  run<1>("fma(a, bs, cs) shared regs", out, in, sms);
  run<2>("fma(a[i], b[i], c[i]) distinct regs", out, in, sms);
  run<3>("a + bs (DADD)", out, in, sms);
  run<4>("a * bs (DMUL)", out, in, sms);
  run<5>("fma(a, a, cs)", out, in, sms);
  run<6>("fma(a, b[i], a)", out, in, sms);
  run<7>("fma(b[i], c[i], a[i]) accumulate", out, in, sms);
  run<8>("fma(b[i], b[i+1], a[i]) accumulate", out, in, sms);
  return 0;
}
