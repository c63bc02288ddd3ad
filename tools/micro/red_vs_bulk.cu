// Microbenchmark: the camera-side scatter of the Schur operator as (A) red.global.add.f64 per (segment, dof) - what the chunk
// kernel's phase 4 does - against (B) one cp.reduce.async.bulk (TMA, UBLKRED.ADD.F64) of a padded 80-byte row per segment.
// 158 segments per "chunk", destinations spread pseudo-randomly over ncam rows, 3 CTAs per SM, many chunks per CTA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_vs_bulk red_vs_bulk.cu ; run: ./red_vs_bulk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NSEG = 158, DC = 9, YS = 10, NCAM = 1778;
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void __launch_bounds__(256, 3) k_red(double* y, int chunks) {
  const int tid = threadIdx.x;
  for (int c = 0; c < chunks; ++c) {
    const uint32_t base = hash(blockIdx.x * 7919u + c);
    for (int idx = tid; idx < NSEG * 3; idx += 256) {
      const int sg = idx / 3, kg = (idx % 3) * 3;
      const uint32_t cam = (base + hash(sg + c * 131u) % 320u) % NCAM;
      double* yr = y + (size_t)cam * DC + kg;
      const double v = 1e-9 * (double)(idx + 1);
      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(yr), "d"(v) : "memory");
      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(yr + 1), "d"(v) : "memory");
      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(yr + 2), "d"(v) : "memory");
    }
  }
}
__global__ void __launch_bounds__(256, 3) k_bulk(double* ypad, int chunks) {
  __shared__ __align__(16) double rows[NSEG * YS];
  const int tid = threadIdx.x;
  for (int c = 0; c < chunks; ++c) {
    const uint32_t base = hash(blockIdx.x * 7919u + c);
    for (int i = tid; i < NSEG * YS; i += 256) rows[i] = (i % YS) < DC ? 1e-9 * (double)(i + 1) : 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int sg = tid; sg < NSEG; sg += 256) {
      const uint32_t cam = (base + hash(sg + c * 131u) % 320u) % NCAM;
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(rows + sg * YS);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 80;" ::"l"(ypad + (size_t)cam * YS), "r"(src) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  double *y, *ypad;
  cudaMalloc(&y, NCAM * DC * 8); cudaMalloc(&ypad, NCAM * YS * 8);
  cudaMemset(y, 0, NCAM * DC * 8); cudaMemset(ypad, 0, NCAM * YS * 8);
  const int grid = 148 * 3, chunks = 48;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_red<<<grid, 256>>>(y, chunks); else k_bulk<<<grid, 256>>>(ypad, chunks);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double per_chunk_sm_us = ms * 1e3 / (chunks * 3.0);   // 3 CTAs per SM share the SM
      printf("%s rep %d: %.3f ms, %.3f us per chunk per SM (%.0f cycles at 1.9 GHz), err %s\n", which ? "bulk(TMA)" : "red", rep, ms, per_chunk_sm_us,
             per_chunk_sm_us * 1900.0, cudaGetErrorString(cudaGetLastError()));
    }
  double h[2] = {0, 0};
  cudaMemcpy(&h[0], y, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&h[1], ypad, 8, cudaMemcpyDeviceToHost);
  printf("y[0] %.6e ypad[0] %.6e\n", h[0], h[1]);
  return 0;
}
