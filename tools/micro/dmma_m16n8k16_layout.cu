#include <cstdio>
#include <cstdlib>
#include <cmath>
// probe: mma.sync m16n8k16 f64 fragment layout on sm_100a. D(16x8) = A(16x16, row) * B(16x8, col)
__global__ void k(const double* A, const double* B, double* D, int variant) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[8], b[4], c[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) {
    int row, col;
    if (variant == 0) { row = g + 8 * (i & 1); col = t + 4 * (i >> 1); }
    else { row = g + 8 * ((i >> 1) & 1); col = t + 4 * (i & 1) + 8 * (i >> 2); }
    a[i] = A[row * 16 + col];
  }
  for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];   // B[k][n], k = t + 4i, n = g
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  D[g * 8 + 2 * t] = c[0]; D[g * 8 + 2 * t + 1] = c[1]; D[(g + 8) * 8 + 2 * t] = c[2]; D[(g + 8) * 8 + 2 * t + 1] = c[3];
}
int main() {
  double hA[256], hB[128], hD[128], ref[128];
  for (int i = 0; i < 256; ++i) hA[i] = (double)((i * 37) % 101) / 7.0;
  for (int i = 0; i < 128; ++i) hB[i] = (double)((i * 53) % 89) / 3.0;
  for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int kk = 0; kk < 16; ++kk) s += hA[m * 16 + kk] * hB[kk * 8 + n]; ref[m * 8 + n] = s; }
  double *dA, *dB, *dD; cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  for (int v = 0; v < 2; ++v) {
    k<<<1, 32>>>(dA, dB, dD, v); cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    double e = 0; for (int i = 0; i < 128; ++i) e = fmax(e, fabs(hD[i] - ref[i]));
    printf("variant %d max err %g (%s)\n", v, e, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
