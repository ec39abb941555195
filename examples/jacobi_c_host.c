/*
 * A host that is NOT Python: plain C against include/glab.h + the CUDA runtime.
 * Builds the 5-point (negative) Laplacian of an N x N grid as int64 COO on the host (the layout
 * of UtilsGNN.py:53-67), uploads it, creates a plan, runs `sweeps` fused weighted-Jacobi sweeps
 * (JacobiGNN.py:119) on the GPU and compares with the same sweeps done by a scalar CPU loop in
 * the reference's operation order.  Prints the max abs difference (expected: 0, bit-exact).
 *
 *   nvcc -O2 -o jacobi_c_host examples/jacobi_c_host.c -Iinclude \
 *        -Lgnn-applied-linear-algebra_b200 -lglab_b200 -Xlinker -rpath=$PWD/gnn-applied-linear-algebra_b200
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "glab.h"

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s failed: %s\n", #x, glab_error_string(rc_)); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 300, sweeps = argc > 2 ? atoi(argv[2]) : 7;
  const int64_t n = (int64_t)N * N;
  int64_t *row = malloc(sizeof(int64_t) * 5 * n), *col = malloc(sizeof(int64_t) * 5 * n);
  float* val = malloc(sizeof(float) * 5 * n);
  int64_t z = 0;
  for (int64_t i = 0; i < n; ++i) {           /* row-major sorted, columns ascending */
    const int y = (int)(i / N), x = (int)(i % N);
    if (y > 0)     { row[z] = i; col[z] = i - N; val[z++] = 1.f; }
    if (x > 0)     { row[z] = i; col[z] = i - 1; val[z++] = 1.f; }
                   { row[z] = i; col[z] = i;     val[z++] = -4.f; }
    if (x < N - 1) { row[z] = i; col[z] = i + 1; val[z++] = 1.f; }
    if (y < N - 1) { row[z] = i; col[z] = i + N; val[z++] = 1.f; }
  }
  float *b = malloc(4 * n), *x0 = malloc(4 * n), *diag = malloc(4 * n), *xa = malloc(4 * n), *xb = malloc(4 * n);
  srand(24601);
  for (int64_t i = 0; i < n; ++i) { b[i] = rand() / (float)RAND_MAX; x0[i] = rand() / (float)RAND_MAX; diag[i] = -4.f; xa[i] = x0[i]; }
  const float w = 0.7f;
  /* CPU: x_i + (w * (b_i - sum_j A_ij x_j)) / A_ii, sum in edge order, no FMA contraction */
  for (int s = 0; s < sweeps; ++s) {
    int64_t e = 0;
    for (int64_t i = 0; i < n; ++i) {
      volatile float acc = 0.f;
      while (e < z && row[e] == i) { volatile float p = val[e] * xa[col[e]]; acc = acc + p; ++e; }
      volatile float t = b[i] - acc; t = w * t; t = t / diag[i];
      xb[i] = xa[i] + t;
    }
    float* tmp = xa; xa = xb; xb = tmp;
  }
  /* GPU through the C ABI */
  int64_t *d_row, *d_col; float *d_val, *d_b, *d_diag, *d_x, *d_y, *d_w;
  CU(cudaMalloc((void**)&d_row, 8 * z)); CU(cudaMalloc((void**)&d_col, 8 * z)); CU(cudaMalloc((void**)&d_val, 4 * z + 64));
  CU(cudaMalloc((void**)&d_b, 4 * n + 64)); CU(cudaMalloc((void**)&d_diag, 4 * n + 64));
  CU(cudaMalloc((void**)&d_x, 4 * n + 64)); CU(cudaMalloc((void**)&d_y, 4 * n + 64)); CU(cudaMalloc((void**)&d_w, 64));
  CU(cudaMemcpy(d_row, row, 8 * z, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_col, col, 8 * z, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_val, val, 4 * z, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_b, b, 4 * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_diag, diag, 4 * n, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_x, x0, 4 * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_w, &w, 4, cudaMemcpyHostToDevice));
  glab_plan* plan = NULL;
  CK(glab_plan_create(n, n, z, d_row, d_col, NULL, &plan));
  int32_t maxrow = 0, ident = 0;
  CK(glab_plan_info(plan, NULL, NULL, NULL, &maxrow, &ident));
  for (int s = 0; s < sweeps; ++s) {
    CK(glab_jacobi_f32(plan, d_val, d_diag, d_b, d_x, d_y, d_w, 1, 0, n, NULL));
    float* t = d_x; d_x = d_y; d_y = t;
  }
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(xb, d_x, 4 * n, cudaMemcpyDeviceToHost));
  double maxdiff = 0;
  for (int64_t i = 0; i < n; ++i) { double d = fabs((double)xb[i] - (double)xa[i]); if (d > maxdiff) maxdiff = d; }
  printf("glab %d: N=%d rows=%lld nnz=%lld max_row_nnz=%d identity_perm=%d sweeps=%d max|gpu-cpu|=%g\n", glab_version(), N,
         (long long)n, (long long)z, maxrow, ident, sweeps, maxdiff);
  glab_plan_destroy(plan);
  return maxdiff == 0.0 ? 0 : 1;
}
