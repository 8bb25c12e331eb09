// latency microbenchmarks for the numbers that bound the block-SVD kernel (one warp, dependent chains)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, double seed) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = seed + i;
  __syncthreads();
  double x = seed + threadIdx.x, y = 1.0000001;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) x = fma(x, y, 1e-9);
  long long t1 = clock64();
  double z = x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) z += __shfl_xor_sync(0xffffffffu, z, 8);
  long long t2 = clock64();
  double r = fabs(z) + 1.0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) r = rsqrt(r) + 1.5;
  long long t3 = clock64();
  double q = r;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) q = sqrt(q) + 2.0;
  long long t4 = clock64();
  double d = q;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) d = 3.0 / d + 1.0;
  long long t5 = clock64();
  int idx = threadIdx.x;
  double acc = 0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { double v = sm[idx]; idx = ((int)v + i) & 1023; acc += v; }
  long long t6 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { __syncthreads(); }
  long long t7 = clock64();
  if (threadIdx.x == 0) {
    clk[0] = (t1 - t0) / 1024; clk[1] = (t2 - t1) / 256; clk[2] = (t3 - t2) / 256; clk[3] = (t4 - t3) / 256; clk[4] = (t5 - t4) / 256;
    clk[5] = (t6 - t5) / 256; clk[6] = (t7 - t6) / 256;
  }
  out[threadIdx.x] = x + z + r + q + d + acc;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 8 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, clk, 1.5);
  long long h[8];
  cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
  printf("1 warp:  DFMA dep %lld clk | shfl64+DADD dep %lld | rsqrt+DADD dep %lld | sqrt+DADD %lld | div+DADD %lld | LDS dep (+cvt) %lld | bar.sync(1 warp) %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  k<<<1, 512>>>(out, clk, 1.5);
  cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
  printf("16 warps: DFMA dep %lld clk | shfl64+DADD dep %lld | rsqrt+DADD dep %lld | sqrt+DADD %lld | div+DADD %lld | LDS dep (+cvt) %lld | bar.sync(16 warps) %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  return 0;
}
