// Development microbenchmark: does integer / LDS work steal issue cycles from the FP64 pipe?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 16, INNER = 1024;

template <int NINT, int NLDS, int KIND>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int* iout) {
  __shared__ double sh[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sh[i] = in[i & 63];
  __syncthreads();
  double a[CH];
  int x[CH];
  for (int i = 0; i < CH; ++i) { a[i] = in[i] + threadIdx.x * 1e-9; x[i] = threadIdx.x + i; }
  const double bs = in[50];
  const int lane16 = threadIdx.x & 15;
#pragma unroll 1
  for (int it = 0; it < INNER; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        a[i] = fma(a[i], a[i], bs);
        if (i < NINT) {
          if (KIND == 0) x[i] = (x[i] ^ (x[i] >> 3)) + it;          // LOP3/SHF/IADD (alu pipe)
          if (KIND == 1) x[i] = x[i] * 5 + it;                        // IMAD (fma pipe)
        }
        if (i < NLDS) a[i] += sh[((x[i] & 255) << 4) + lane16];
      }
  }
  double s = 0; int t = 0;
  for (int i = 0; i < CH; ++i) { s += a[i]; t += x[i]; }
  if (s == 123.456) out[0] = s;
  if (t == 123456789) iout[0] = t;
}

template <int NINT, int NLDS, int KIND>
void run(const char* name, double* out, double* in, int* iout, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = sms * 8;
  k<NINT, NLDS, KIND><<<blocks, 256>>>(out, in, iout); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 10; ++r) k<NINT, NLDS, KIND><<<blocks, 256>>>(out, in, iout);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double slots = (double)blocks * 256 * CH * INNER * 4 * 10;
  double rate = slots / (ms * 1e-3);
  printf("%-50s DFMA %.3f inst/cycle/SMSP\n", name, rate / 32 / (sms * 4) / 1.965e9);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in; int* iout; cudaMalloc(&out, 8); cudaMalloc(&in, 64 * 8); cudaMalloc(&iout, 4);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5 + i * 1e-9; h[50] = 0.1;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0, 0, 0>("16 DFMA only", out, in, iout, sms);
  run<4, 0, 0>("16 DFMA + 4x(xor,shift,add)", out, in, iout, sms);
  run<8, 0, 0>("16 DFMA + 8x(xor,shift,add)", out, in, iout, sms);
  run<16, 0, 0>("16 DFMA + 16x(xor,shift,add)", out, in, iout, sms);
  run<8, 0, 1>("16 DFMA + 8 IMAD", out, in, iout, sms);
  run<16, 0, 1>("16 DFMA + 16 IMAD", out, in, iout, sms);
  run<16, 4, 0>("16 DFMA + 16x(alu) + 4 LDS.64(+DADD)", out, in, iout, sms);
  run<16, 16, 0>("16 DFMA + 16x(alu) + 16 LDS.64(+DADD)", out, in, iout, sms);
  return 0;
}
