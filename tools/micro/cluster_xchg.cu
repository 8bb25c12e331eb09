// Microbenchmark behind the cluster-split Jacobi: latency of one "all-CTAs exchange partial sums" round inside a thread-block
// cluster on B200.  Every CTA of a C-CTA cluster sends NP 16-byte values to every peer with st.async (data + mbarrier
// complete_tx in one message), waits on its own mbarrier, sums the C partials in rank order, and feeds the sum into the next
// round (so rounds are dependent, like Jacobi rounds).  Prints clocks per round and checks the sums.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/cluster_xchg tools/micro/cluster_xchg.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v2(unsigned raddr, double a, double b, unsigned rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "d"(a), "d"(b), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int MAXC = 16;
constexpr int NP = 32;      // values exchanged per round and CTA

__global__ void xchg_kernel(int C, int rounds, int mode, double* out, long long* clk) {
  __shared__ __align__(16) double2 buf[2][MAXC][NP];
  __shared__ __align__(8) unsigned long long mbar[2];
  const unsigned rank = cluster_rank();
  const int tid = threadIdx.x;
  const unsigned mb0 = smem_u32(&mbar[0]), mb1 = smem_u32(&mbar[1]);
  if (tid == 0) {
    mbar_init(mb0, 1); mbar_init(mb1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect(mb0, (C - 1) * NP * 16);
    mbar_expect(mb1, (C - 1) * NP * 16);
  }
  __syncthreads();
  cluster_sync();
  double vx = 1.0 + tid, vy = 0.5 * rank;
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const int b = r & 1;
    const unsigned par = (r >> 1) & 1;
    if (tid < NP) {
      if (mode == 0) {
        buf[b][rank][tid] = make_double2(vx, vy);
        const unsigned la = smem_u32(&buf[b][rank][tid]);
        for (int p = 1; p < C; ++p) {
          const unsigned peer = (rank + p) % C;
          st_async_v2(mapa(la, peer), vx, vy, mapa(b ? mb1 : mb0, peer));
        }
      }
    }
    if (mode == 0) {
      mbar_wait(b ? mb1 : mb0, par);
    } else {            // hardware cluster barrier + plain remote stores
      if (tid < NP) {
        for (int p = 0; p < C; ++p) {
          const unsigned ra = mapa(smem_u32(&buf[b][rank][tid]), p);
          asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(ra), "d"(vx), "d"(vy) : "memory");
        }
      }
      cluster_sync();
    }
    double sx = 0.0, sy = 0.0;
    if (tid < NP) {
      for (int p = 0; p < C; ++p) { const double2 v = buf[b][p][tid]; sx += v.x; sy += v.y; }
    }
    if (mode == 0) {
      __syncthreads();                              // everyone has read buffer b: re-arm its barrier for round r+2
      if (tid == 0) mbar_expect(b ? mb1 : mb0, (C - 1) * NP * 16);
    }
    vx = sx * (1.0 / C);                            // dependent on the exchange
    vy = sy * (1.0 / C) + 1.0;
  }
  const long long t1 = clock64();
  cluster_sync();
  if (tid < NP) { out[(blockIdx.x * NP + tid) * 2] = vx; out[(blockIdx.x * NP + tid) * 2 + 1] = vy; }
  if (tid == 0) clk[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
  double* d_out; long long* d_clk;
  cudaMalloc(&d_out, sizeof(double) * 2 * NP * 64);
  cudaMalloc(&d_clk, sizeof(long long) * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int C : {1, 2, 4, 8, 16}) {
      for (int threads : {32, 128}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C, 1, 1); cfg.blockDim = dim3(threads, 1, 1);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (C > 8) cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaError_t e = cudaLaunchKernelEx(&cfg, xchg_kernel, C, rounds, mode, d_out, d_clk);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("mode %d C %d threads %d: %s / %s\n", mode, C, threads, cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); continue; }
        long long clk[64]; double out[2 * NP * 16];
        cudaMemcpy(clk, d_clk, sizeof(long long) * C, cudaMemcpyDeviceToHost);
        cudaMemcpy(out, d_out, sizeof(double) * 2 * NP * C, cudaMemcpyDeviceToHost);
        // expected: all CTAs hold the same vx (mean over ranks of the same value = itself) ; vy converges to mean + r
        bool same = true;
        for (int c = 1; c < C; ++c) for (int t = 0; t < NP; ++t) same = same && out[(c * NP + t) * 2] == out[t * 2] && out[(c * NP + t) * 2 + 1] == out[t * 2 + 1];
        printf("mode %s C %2d threads %3d: %7.1f clk/round  (vx[3] %.3f vy[3] %.3f, identical across CTAs: %s)\n", mode == 0 ? "st.async+mbarrier" : "st.cluster+barrier.cluster",
               C, threads, (double)clk[0] / rounds, out[6], out[7], same ? "yes" : "NO");
      }
    }
  return 0;
}
