// Development microbenchmark 2: which integer forms co-issue with DFMA on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 16, INNER = 1024;

template <int KIND>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int* iout) {
  __shared__ double sh[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sh[i] = in[i & 63];
  __syncthreads();
  double a[CH];
  unsigned x[CH];
  for (int i = 0; i < CH; ++i) { a[i] = in[i] + threadIdx.x * 1e-9; x[i] = threadIdx.x * 77 + i; }
  const double bs = in[50];
  const unsigned lane16 = (threadIdx.x & 15) * 8;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sh);
#pragma unroll 1
  for (int it = 0; it < INNER; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        a[i] = fma(a[i], a[i], bs);
        if (KIND == 1) x[i] = __umulhi(x[i] << 24, 0x8000u) + x[i];            // IMAD.SHL + IMAD.HI
        if (KIND == 2) x[i] = (x[i] & 255u) + (unsigned)it;                   // LOP3 + IADD
        if (KIND == 3) x[i] = min(x[i] + 1u, 0x40862000u);                    // IADD + VIMNMX
        if (KIND == 4) x[i] = x[i] * 4096u + (unsigned)it;                    // IMAD
        if (KIND == 5) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sbase + ((x[i] & 0x7F80u)) + lane16));
                         x[i] = (unsigned)__double2loint(v) * 3u + x[i]; }    // LOP3 + LDS + IMAD
        if (KIND == 6) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sbase + lane16 + i * 128));
                         x[i] = (unsigned)__double2loint(v) * 3u + x[i]; }    // LDS + IMAD (fixed address)
        if (KIND == 7) x[i] = __byte_perm(x[i], 0, 0x4440) + (unsigned)it;    // PRMT + IADD
      }
  }
  double s = 0; unsigned t = 0;
  for (int i = 0; i < CH; ++i) { s += a[i]; t += x[i]; }
  if (s == 123.456) out[0] = s;
  if (t == 123456789u) iout[0] = t;
}

template <int KIND>
void run(const char* name, double* out, double* in, int* iout, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = sms * 8;
  k<KIND><<<blocks, 256>>>(out, in, iout); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 10; ++r) k<KIND><<<blocks, 256>>>(out, in, iout);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double slots = (double)blocks * 256 * CH * INNER * 4 * 10;
  double rate = slots / (ms * 1e-3);
  printf("%-50s DFMA %.3f inst/cycle/SMSP  (%.1f cycles per 16 DFMA)\n", name, rate / 32 / (sms * 4) / 1.965e9, 16.0 / (rate / 32 / (sms * 4) / 1.965e9));
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in; int* iout; cudaMalloc(&out, 8); cudaMalloc(&in, 64 * 8); cudaMalloc(&iout, 4);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5 + i * 1e-9; h[50] = 0.1;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("16 DFMA only", out, in, iout, sms);
  run<1>("+16x(IMAD.SHL, IMAD.HI, IADD)", out, in, iout, sms);
  run<2>("+16x(LOP3, IADD)", out, in, iout, sms);
  run<3>("+16x(IADD, VIMNMX)", out, in, iout, sms);
  run<4>("+16x IMAD", out, in, iout, sms);
  run<5>("+16x(LOP3, LDS.64 var addr, IMAD)", out, in, iout, sms);
  run<6>("+16x(LDS.64 fixed addr, IMAD)", out, in, iout, sms);
  run<7>("+16x(PRMT, IADD)", out, in, iout, sms);
  return 0;
}
