// Measured FP64 roofline denominator: a register-resident DFMA throughput microbenchmark.
// MEASURED_PEAKS.json carries no FP64 figure, so bench.py calls this on the same box, in the same
// run, and quotes every FP64 fraction "of measured".
#include "oak_common.cuh"

namespace oak {

constexpr int kChains = 16;
constexpr int kInner = 2048;

__global__ void __launch_bounds__(512, 1) dfma_peak_kernel(double* out, double seed) {
  double a[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) a[i] = seed + (double)(threadIdx.x + i) * 1e-9;
  const double m = 1.0 - 1e-12, c = 1e-13;
#pragma unroll 1
  for (int it = 0; it < kInner; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;  // never true; keeps the chain alive
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
  const int blocks = sms * 8, threads = 512;
  const double slots_per_launch = (double)blocks * threads * kChains * kInner;
  // warm-up
  for (int i = 0; i < 3; ++i) dfma_peak_kernel<<<blocks, threads, 0, stream>>>(d_out, 1.0);
  OAK_CUDA(cudaStreamSynchronize(stream));
  double best = 0.0, elapsed = 0.0;
  while (elapsed < seconds) {
    const int reps = 20;
    OAK_CUDA(cudaEventRecord(e0, stream));
    for (int i = 0; i < reps; ++i) dfma_peak_kernel<<<blocks, threads, 0, stream>>>(d_out, 1.0);
    OAK_CUDA(cudaEventRecord(e1, stream));
    OAK_CUDA(cudaEventSynchronize(e1));
    g_launches.fetch_add(reps);
    float ms = 0.f;
    OAK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double rate = slots_per_launch * reps / (ms * 1e-3);
    if (rate > best) best = rate;
    elapsed += ms * 1e-3;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *h_slots_per_s = best;
  return 0;
}
