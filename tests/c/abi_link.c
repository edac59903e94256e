/* A plain C99 program that LINKS liboak_b200.so through include/oak_b200.h -- what a maintainer's cgo / JNI /
 * ctypes-free binding does ("no C program links the ABI").
 *   abi_link --no-gpu : housekeeping calls only (CPU test suite)
 *   abi_link          : K(X, X) and K_diag(X) of a two-dimensional constrained OAK kernel (Gaussian measure, depth 2)
 *                       through oak_spec_create / oak_prepare_points_f64 / oak_gram_f64 / oak_gram_diag_f64, compared
 *                       with the closed forms of oak/ortho_rbf_kernel.py:82-97, 157-177 and oak/oak_kernel.py:251-278
 *                       evaluated here in C. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oak_b200.h"

#define N 37
#define D 2

static double ktilde(double x, double y, double l, double mu, double var) {
  const double base = exp(-0.5 * (x - y) * (x - y) / (l * l));
  const double cx = l / sqrt(l * l + var) * exp(-0.5 * (x - mu) * (x - mu) / (l * l + var));
  const double cy = l / sqrt(l * l + var) * exp(-0.5 * (y - mu) * (y - mu) / (l * l + var));
  return base - cx * cy / (l / sqrt(l * l + 2.0 * var));
}

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      fprintf(stderr, "%s failed: %s\n", #call, oak_last_error());         \
      return 2;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  printf("oak_version %d, devices %d, stats_count(4) %zu\n", oak_version(), oak_device_count(), oak_sgpr_stats_count(4));
  if (oak_version() < 100 || oak_sgpr_stats_count(4) != 4 * 4 + 4 + 2) return 1;
  if (argc > 1 && strcmp(argv[1], "--no-gpu") == 0) return 0;

  const double ls[D] = {0.8, 1.7}, sig[3] = {0.4, 1.1, 0.6}, mu = 0.2, var = 1.5;
  oak_dim_desc dims[D];
  memset(dims, 0, sizeof(dims));
  for (int d = 0; d < D; ++d) {
    dims[d].type = OAK_DIM_RBF;
    dims[d].column = d;
    dims[d].measure = OAK_MEASURE_GAUSSIAN;
    dims[d].lengthscale = ls[d];
    dims[d].variance = 1.0;
    dims[d].m0 = mu;
    dims[d].m1 = var;
  }
  oak_kernel_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.num_dims = D;
  desc.depth = 2;
  desc.share_var_across_orders = 1;
  desc.esp_algorithm = OAK_ESP_NEWTON_GIRARD;
  desc.variances = sig;
  desc.dims = dims;

  double X[N * D];
  unsigned s = 12345u;
  for (int i = 0; i < N * D; ++i) {
    s = s * 1664525u + 1013904223u;
    X[i] = ((double)(s >> 8) / 16777216.0 - 0.5) * 5.0;
  }
  oak_spec* spec = NULL;
  CHECK(oak_spec_create(&desc, NULL, &spec));
  double *dX = NULL, *dK = NULL, *dDiag = NULL;
  void* dP = NULL;
  if (cudaMalloc((void**)&dX, sizeof(X)) || cudaMalloc(&dP, oak_points_bytes(spec, N)) ||
      cudaMalloc((void**)&dK, sizeof(double) * N * N) || cudaMalloc((void**)&dDiag, sizeof(double) * N)) {
    fprintf(stderr, "cudaMalloc failed\n");
    return 3;
  }
  cudaMemcpy(dX, X, sizeof(X), cudaMemcpyHostToDevice);
  CHECK(oak_prepare_points_f64(spec, dX, N, D, dP, NULL));
  CHECK(oak_gram_f64(spec, dP, N, NULL, 0, 0, N, dK, N, NULL));
  CHECK(oak_gram_diag_f64(spec, dP, N, dDiag, NULL));
  static double K[N * N], diag[N];
  if (cudaMemcpy(K, dK, sizeof(K), cudaMemcpyDeviceToHost) || cudaMemcpy(diag, dDiag, sizeof(diag), cudaMemcpyDeviceToHost)) {
    fprintf(stderr, "copy back failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 3;
  }
  double worst = 0.0, scale = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      const double k1 = ktilde(X[i * D], X[j * D], ls[0], mu, var), k2 = ktilde(X[i * D + 1], X[j * D + 1], ls[1], mu, var);
      const double ref = sig[0] + sig[1] * (k1 + k2) + sig[2] * k1 * k2; /* e_0, e_1, e_2 weighted */
      const double err = fabs(K[i * N + j] - ref);
      if (err > worst) worst = err;
      if (fabs(ref) > scale) scale = fabs(ref);
      if (i == j && fabs(diag[i] - ref) > 1e-12 * (1.0 + fabs(ref))) {
        fprintf(stderr, "K_diag[%d] = %.17g, expected %.17g\n", i, diag[i], ref);
        return 4;
      }
    }
  printf("max |K - closed form| / max |K| = %.3e\n", worst / scale);
  CHECK(oak_spec_destroy(spec));
  cudaFree(dX); cudaFree(dP); cudaFree(dK); cudaFree(dDiag);
  return worst / scale < 1e-12 ? 0 : 5;
}
