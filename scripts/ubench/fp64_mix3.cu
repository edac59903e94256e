// Development microbenchmark 3: candidates for the exp-table index math next to a saturated FP64
// pipe (sm_100a).  Multipliers come from memory so that ptxas cannot strength-reduce them to
// LEA / SHF.  Each loop body: 16 independent DFMA chains + 16x the candidate integer sequence.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 16, INNER = 1024;

template <int KIND>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int* iout, const unsigned* cst) {
  double a[CH];
  unsigned x[CH];
  for (int i = 0; i < CH; ++i) { a[i] = in[i] + threadIdx.x * 1e-9; x[i] = threadIdx.x * 77 + i; }
  const double bs = in[50];
  const unsigned m15 = cst[0];           // 32768, opaque
  const unsigned lane = (threadIdx.x & 15) * 8;
#pragma unroll 1
  for (int it = 0; it < INNER; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        a[i] = fma(a[i], a[i], bs);
        unsigned r = 0;
        if (KIND == 1) {  // IMAD.SHL + IMAD.HI.U32 (register multiplier)
          asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x[i] * 16777216u), "r"(m15), "r"(lane));
        }
        if (KIND == 2) {  // IMAD.SHL + IMAD.WIDE.U32, take the high word
          unsigned long long w;
          asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(x[i] * 16777216u), "r"(m15), "l"((unsigned long long)lane << 32));
          r = (unsigned)(w >> 32);
        }
        if (KIND == 3) {  // IMAD.SHL + LOP3 (round-1 kernel)
          r = ((x[i] * 128u) & 0x7F80u) | lane;
        }
        if (KIND == 4) {  // IMAD.SHL + LEA.HI (what ptxas makes of mad.hi with an immediate 2^15)
          asm volatile("mad.hi.u32 %0, %1, 32768, %2;" : "=r"(r) : "r"(x[i] * 16777216u), "r"(lane));
        }
        if (KIND == 5) {  // + VIMNMX (the clamp)
          r = min(x[i], 0x40862000u);
        }
        if (KIND == 6) {  // two IMADs (cost of pure FMA-pipe integer work)
          r = x[i] * m15 + lane; r = r * m15 + lane;
        }
        if (KIND != 0) x[i] = x[i] + r;
      }
  }
  double s = 0; unsigned t = 0;
  for (int i = 0; i < CH; ++i) { s += a[i]; t += x[i]; }
  if (s == 123.456) out[0] = s;
  if (t == 123456789u) iout[0] = t;
}

// DMNMX / 3-operand forms on the FP64 pipe
template <int KIND>
__global__ void __launch_bounds__(256) k2(double* out, const double* in) {
  double a[CH], b[CH], c[CH];
  for (int i = 0; i < CH; ++i) { a[i] = in[i] + threadIdx.x * 1e-9; b[i] = in[16 + i]; c[i] = in[32 + i]; }
  const double bs = in[50];
#pragma unroll 1
  for (int it = 0; it < INNER; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (KIND == 0) a[i] = fma(a[i], a[i], bs);
        if (KIND == 1) { a[i] = fma(a[i], a[i], bs); a[i] = fmin(a[i], 708.0); }     // + DMNMX?
        if (KIND == 2) a[i] = fma(b[i], c[i], a[i]);                                   // 3 distinct
        if (KIND == 3) a[i] = fma(b[i], c[(i + 1) % CH], a[i]);                        // 3 distinct, b reused by neighbour?
        if (KIND == 4) a[i] = fma(b[i / 4], c[i % 4], a[i]);                           // outer-product pattern
      }
  }
  double s = 0;
  for (int i = 0; i < CH; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
void timeit(const char* name, F launch, int sms, double per_launch_dfma) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 10; ++r) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double rate = per_launch_dfma * 10 / (ms * 1e-3) / 32 / (sms * 4) / 1.965e9;
  printf("%-58s DFMA %.3f inst/cycle/SMSP  (%.1f cycles per 16 DFMA)\n", name, rate, 16.0 / rate);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in; int* iout; unsigned* cst;
  cudaMalloc(&out, 8); cudaMalloc(&in, 64 * 8); cudaMalloc(&iout, 4); cudaMalloc(&cst, 16);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5 + i * 1e-9; h[50] = 0.1;
  for (int i = 16; i < 48; ++i) h[i] = 1e-9;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  unsigned hc[4] = {32768u, 0, 0, 0}; cudaMemcpy(cst, hc, sizeof(hc), cudaMemcpyHostToDevice);
  const int blocks = sms * 8;
  const double per = (double)blocks * 256 * CH * INNER * 4;
#define RUN(K, NAME) timeit(NAME, [&] { k<K><<<blocks, 256>>>(out, in, iout, cst); }, sms, per)
  RUN(0, "16 DFMA only");
  RUN(1, "+16x(IMAD.SHL, IMAD.HI.U32 reg multiplier, IADD)");
  RUN(2, "+16x(IMAD.SHL, IMAD.WIDE.U32, IADD)");
  RUN(3, "+16x(IMAD.SHL, LOP3, IADD)   [round-1 kernel]");
  RUN(4, "+16x(IMAD.SHL, LEA.HI, IADD)");
  RUN(5, "+16x(VIMNMX, IADD)");
  RUN(6, "+16x(IMAD, IMAD, IADD)");
#define RUN2(K, NAME) timeit(NAME, [&] { k2<K><<<blocks, 256>>>(out, in); }, sms, per)
  RUN2(0, "fma(a,a,bs)");
  RUN2(1, "fma(a,a,bs); fmin(a,708)  (per DFMA)");
  RUN2(2, "fma(b[i],c[i],a[i]) 3 distinct");
  RUN2(3, "fma(b[i],c[i+1],a[i]) 3 distinct");
  RUN2(4, "fma(b[i/4],c[i%4],a[i]) outer product");
  return 0;
}
