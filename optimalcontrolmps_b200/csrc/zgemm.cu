// Complex-FP64 grouped GEMM on the FP64 tensor cores (DMMA, mma.sync m8n8k4.f64).
//
// Every contraction of the engine (two-site merge, projections U^H.theta, gauge pushes,
// transfer matrices) goes through this kernel.  Problem sizes are read from device-resident
// descriptors because bond dimensions are decided on the GPU; the grid is sized for the
// capacity and surplus CTAs exit immediately.
//
// Complex product on real tensor cores: Cr += Ar.Br - Ai.Bi ; Ci += Ar.Bi + Ai.Br  (4 DMMA per
// k4-step and 8x8 tile).  Operands are split into re/im planes while being staged in shared
// memory; HBM keeps interleaved complex128 so that slices are plain numpy-compatible arrays.
#include "ocmps_internal.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int APAD = 1, BPAD = 8;   // shared-memory paddings (doubles)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) zgemm_kernel(const GemmDesc* __restrict__ descs) {
  const GemmDesc d = descs[blockIdx.z];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= d.M || n0 >= d.N) return;

  __shared__ double As_re[BM][BK + APAD], As_im[BM][BK + APAD];
  __shared__ double Bs_re[BK][BN + BPAD], Bs_im[BK][BN + BPAD];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;   // groupID, thread-in-group (mma fragment layout)

  double cr[8][2], ci[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { cr[i][0] = cr[i][1] = ci[i][0] = ci[i][1] = 0.0; }

  // global -> registers -> shared staging: the loads of tile k+1 are in flight while tile k is multiplied
  constexpr int PER = BM * BK / 256;      // elements of each operand per thread and tile (= 4)
  cplx ra[PER], rb[PER];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = tid + u * 256;
      {
        int i, kk;
        if (d.opA == 0) { i = idx / BK; kk = idx % BK; } else { i = idx % BM; kk = idx / BM; }
        const int gi = m0 + i, gk = k0 + kk;
        cplx v = make_double2(0.0, 0.0);
        if (gi < d.M && gk < d.K) v = d.opA == 0 ? d.A[(size_t)gi * d.lda + gk] : d.A[(size_t)gk * d.lda + gi];
        if (d.opA) v.y = -v.y;
        ra[u] = v;
      }
      {
        int kk, j;
        if (d.opB == 0) { kk = idx / BN; j = idx % BN; } else { kk = idx % BK; j = idx / BK; }
        const int gk = k0 + kk, gj = n0 + j;
        cplx v = make_double2(0.0, 0.0);
        if (gk < d.K && gj < d.N) v = d.opB == 0 ? d.B[(size_t)gk * d.ldb + gj] : d.B[(size_t)gj * d.ldb + gk];
        if (d.opB) v.y = -v.y;
        rb[u] = v;
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = tid + u * 256;
      int i, kk;
      if (d.opA == 0) { i = idx / BK; kk = idx % BK; } else { i = idx % BM; kk = idx / BM; }
      As_re[i][kk] = ra[u].x; As_im[i][kk] = ra[u].y;
      int kb, j;
      if (d.opB == 0) { kb = idx / BN; j = idx % BN; } else { kb = idx % BK; j = idx / BK; }
      Bs_re[kb][j] = rb[u].x; Bs_im[kb][j] = rb[u].y;
    }
  };
  load_tiles(0);
  for (int k0 = 0; k0 < d.K; k0 += BK) {
    // MPS tensors are block sparse in the boson number and stored dense with exact zeros: a staged tile of A or of B
    // that is entirely zero contributes nothing, and most (tile, k-chunk) combinations are of that kind
    bool nzA = false, nzB = false;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      nzA |= (ra[u].x != 0.0) | (ra[u].y != 0.0);
      nzB |= (rb[u].x != 0.0) | (rb[u].y != 0.0);
    }
    store_tiles();
    const int anyA = __syncthreads_or(nzA ? 1 : 0);      // (the barrier also publishes the staged tiles)
    const int anyB = __syncthreads_or(nzB ? 1 : 0);
    if (k0 + BK < d.K) load_tiles(k0 + BK);
    if (anyA && anyB) {
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      const double ar = As_re[warp * 8 + g][k4 + t];
      const double ai = As_im[warp * 8 + g][k4 + t];
      const double nai = -ai;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const double br = Bs_re[k4 + t][nt * 8 + g];
        const double bi = Bs_im[k4 + t][nt * 8 + g];
        dmma(cr[nt][0], cr[nt][1], ar, br);
        dmma(cr[nt][0], cr[nt][1], nai, bi);
        dmma(ci[nt][0], ci[nt][1], ar, bi);
        dmma(ci[nt][0], ci[nt][1], ai, br);
      }
    }
    }
    __syncthreads();
  }
  // ---- epilogue: C fragment element (row g, cols 2t, 2t+1) of each 8x8 tile ----
  const int gi = m0 + warp * 8 + g;
  if (gi < d.M) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      int gj = n0 + nt * 8 + 2 * t;
      if (gj < d.N) d.C[(size_t)gi * d.ldc + gj] = make_double2(cr[nt][0], ci[nt][0]);
      if (gj + 1 < d.N) d.C[(size_t)gi * d.ldc + gj + 1] = make_double2(cr[nt][1], ci[nt][1]);
    }
  }
}

}  // namespace

void launch_zgemm(const GemmDesc* d_descs, int batch, int maxM, int maxN, cudaStream_t s) {
  if (batch <= 0 || maxM <= 0 || maxN <= 0) return;
  dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, batch);
  zgemm_kernel<<<grid, 256, 0, s>>>(d_descs);
}
