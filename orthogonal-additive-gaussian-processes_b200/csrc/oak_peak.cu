// Measured FP64 roofline denominator: a register-resident DFMA throughput microbenchmark.
// MEASURED_PEAKS.json carries no FP64 figure, so bench.py calls this on the same box, in the same
// run, and quotes every FP64 fraction "of measured".  Several shapes (independent chains per
// thread, threads per block) are tried and the best sustained rate is reported.
#include "oak_common.cuh"

namespace oak {

constexpr int kInner = 1024;

template <int CHAINS>
__global__ void dfma_peak_kernel(double* out, double seed) {
  double a[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) a[i] = seed + (double)(threadIdx.x + i) * 1e-9;
  const double m = 1.0 - 1e-12, c = 1e-13;
#pragma unroll 1
  for (int it = 0; it < kInner; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

template <int CHAINS>
static int time_shape(int blocks, int threads, double seconds, double* d_out, cudaEvent_t e0,
                      cudaEvent_t e1, cudaStream_t stream, double* best) {
  const double slots_per_launch = (double)blocks * threads * CHAINS * kInner * 4.0;
  for (int i = 0; i < 2; ++i) dfma_peak_kernel<CHAINS><<<blocks, threads, 0, stream>>>(d_out, 1.0);
  OAK_CUDA(cudaStreamSynchronize(stream));
  double elapsed = 0.0;
  while (elapsed < seconds) {
    const int reps = 10;
    OAK_CUDA(cudaEventRecord(e0, stream));
    for (int i = 0; i < reps; ++i) dfma_peak_kernel<CHAINS><<<blocks, threads, 0, stream>>>(d_out, 1.0);
    OAK_CUDA(cudaEventRecord(e1, stream));
    OAK_CUDA(cudaEventSynchronize(e1));
    g_launches.fetch_add(reps);
    float ms = 0.f;
    OAK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double rate = slots_per_launch * reps / (ms * 1e-3);
    if (rate > *best) *best = rate;
    elapsed += ms * 1e-3;
  }
  return 0;
}

}  // namespace oak

using namespace oak;

extern "C" int oak_measure_fp64_peak(double seconds, double* h_slots_per_s, void* stream_) {
  OAK_REQUIRE(h_slots_per_s, "oak_measure_fp64_peak: null output");
  OAK_REQUIRE(oak_device_count() > 0, "oak_measure_fp64_peak: no CUDA device visible");
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  OAK_CUDA(cudaGetDevice(&dev));
  OAK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d_out = nullptr;
  OAK_CUDA(cudaMalloc(&d_out, sizeof(double)));
  cudaEvent_t e0, e1;
  OAK_CUDA(cudaEventCreate(&e0));
  OAK_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  const double each = seconds / 6.0;
  int rc = 0;
  if (!rc) rc = time_shape<8>(sms * 8, 256, each, d_out, e0, e1, stream, &best);
  if (!rc) rc = time_shape<8>(sms * 4, 512, each, d_out, e0, e1, stream, &best);
  if (!rc) rc = time_shape<16>(sms * 8, 256, each, d_out, e0, e1, stream, &best);
  if (!rc) rc = time_shape<16>(sms * 4, 512, each, d_out, e0, e1, stream, &best);
  if (!rc) rc = time_shape<32>(sms * 4, 256, each, d_out, e0, e1, stream, &best);
  if (!rc) rc = time_shape<32>(sms * 8, 128, each, d_out, e0, e1, stream, &best);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  if (rc) return rc;
  *h_slots_per_s = best;
  return 0;
}
