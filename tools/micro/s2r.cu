// Cost of reading SR_CgaCtaId (what the compiler emits, S2R / S2UR, whenever it re-materialises the address of a __shared__
// variable on sm_90+ instead of keeping it in a register): clocks per read, dependent and independent, for 1..16 warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o s2r s2r.cu && ./s2r
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dep_kernel(int iters, unsigned* out, long long* t) {
  unsigned acc = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    acc = acc * 3u + r;                       // the read is consumed at once
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) t[blockIdx.x] = t1 - t0;
}

__global__ void indep_kernel(int iters, unsigned* out, long long* t) {
  unsigned a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    unsigned r0, r1, r2, r3;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r0));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r2));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r3));
    a0 += r0; a1 += r1; a2 += r2; a3 += r3;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) t[blockIdx.x] = t1 - t0;
}

// the same loop without the special-register read (loop skeleton)
__global__ void base_kernel(int iters, unsigned* out, long long* t) {
  unsigned acc = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    unsigned r;
    asm volatile("mov.u32 %0, 7;" : "=r"(r));
    acc = acc * 3u + r;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) t[blockIdx.x] = t1 - t0;
}

int main() {
  unsigned* out; long long* t;
  cudaMalloc(&out, 1024 * 4 * 4); cudaMalloc(&t, 64);
  const int iters = 4096;
  for (int threads : {32, 128, 256, 512, 1024}) {
    long long h[3];
    base_kernel<<<1, threads>>>(iters, out, t);  cudaMemcpy(&h[0], t, 8, cudaMemcpyDeviceToHost);
    dep_kernel<<<1, threads>>>(iters, out, t);   cudaMemcpy(&h[1], t, 8, cudaMemcpyDeviceToHost);
    indep_kernel<<<1, threads>>>(iters, out, t); cudaMemcpy(&h[2], t, 8, cudaMemcpyDeviceToHost);
    printf("threads %4d: skeleton %.1f clk/iter | dependent read %.1f clk/read | 4 independent reads %.1f clk/read (per warp)\n", threads,
           (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters / 4);
  }
  return 0;
}
