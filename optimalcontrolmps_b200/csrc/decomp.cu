// Truncated decompositions of the engine: charge-blocked one-sided Jacobi SVD in shared memory.
//
// The reference truncates with ITensor's denmatDecomp (eigen-decomposition of the two-site reduced
// density matrix, i.e. the Gram matrix theta.theta^H) per particle-number block, followed by a
// global sort of all eigenvalues and the Cutoff/Maxm rule, and moves the
// orthogonality centre with block SVDs (A.4).  The eigenvalues of the Gram matrix are the squared
// singular values of theta, and its eigenvectors are the singular vectors, so here each charge block
// X_q is orthogonalised directly with Hestenes' one-sided Jacobi: pairs of vectors are rotated until
// all Gram entries <y_p|y_q> vanish; the Gram diagonal (squared norms) is the spectrum the truncation
// rule sees.  One CTA per block, the block lives in shared memory, one warp per vector pair, dot
// products by warp shuffles.
#include "ocmps_internal.h"

namespace {

constexpr int JAC_THREADS = 512;
constexpr int JAC_MAX_SWEEPS = 60;
constexpr double JAC_TOL2 = 1e-28;        // rotate while |<p|q>|^2 > tol^2 <p|p><q|q>, tol = 1e-14
constexpr int NV_MAX = 2048;              // max number of vectors in one decomposition

// ------------------------------------------------------------------------------------------------
// setup: charges of rows / columns, block table, sorted index lists
// ------------------------------------------------------------------------------------------------
struct Geometry { int n, m, mode; };

__device__ __forceinline__ Geometry geometry(const DecompArgs& a, int chiL, int chiR) {
  Geometry g;
  switch (a.kind) {
    case DK_GATE_LEFT:  g.n = chiL * a.D; g.m = a.D * chiR; g.mode = 0; break;
    case DK_GATE_RIGHT: g.n = chiL * a.D; g.m = a.D * chiR; g.mode = 1; break;
    case DK_ORTH_LEFT:  g.n = chiL * a.D; g.m = chiR;       g.mode = 0; break;
    default:            g.n = chiL;       g.m = a.D * chiR; g.mode = 1; break;
  }
  return g;
}
__device__ __forceinline__ int row_charge(const DecompArgs& a, int i) {
  return (a.kind == DK_ORTH_RIGHT) ? a.qL[i] : a.qL[i / a.D] + i % a.D;
}
__device__ __forceinline__ int col_charge(const DecompArgs& a, int j, int chiR) {
  return (a.kind == DK_ORTH_LEFT) ? a.qR[j] : a.qR[j % chiR] - j / chiR;
}

// comp_blk / comp_rank live behind comp_idx in the same allocation (see engine: 3 * NV_MAX ints)
__global__ void __launch_bounds__(OCMPS_MAX_Q) decomp_setup_kernel(DecompArgs a, DecompBuffers b) {
  __shared__ int cnt_v[OCMPS_MAX_Q], cnt_c[OCMPS_MAX_Q], off_v[OCMPS_MAX_Q], off_c[OCMPS_MAX_Q], blk_of_q[OCMPS_MAX_Q];
  const int chiL = *a.dimL, chiR = *a.dimR;
  const Geometry g = geometry(a, chiL, chiR);
  const int q = threadIdx.x;
  const int nvec = g.mode == 0 ? g.m : g.n;
  const int ncomp = g.mode == 0 ? g.n : g.m;
  int* comp_blk = b.comp_idx + NV_MAX;
  int* comp_rank = b.comp_idx + 2 * NV_MAX;

  int cv = 0, cc = 0, bad = 0;
  for (int i = 0; i < g.n; ++i) {
    int c = row_charge(a, i);
    if (c >= OCMPS_MAX_Q) bad = 1;
    if (c == q) { if (g.mode == 0) ++cc; else ++cv; }
  }
  for (int j = 0; j < g.m; ++j) {
    int c = col_charge(a, j, chiR);
    if (c >= OCMPS_MAX_Q) bad = 1;
    if (c == q) { if (g.mode == 0) ++cv; else ++cc; }
  }
  cnt_v[q] = cv; cnt_c[q] = cc;
  if (bad && q == 0) atomicOr(b.status, OCMPS_ST_CHARGE);
  __syncthreads();
  if (q == 0) {
    DecompWork* w = b.dw;
    int nb = 0, ov = 0, oc = 0, ows = 0;
    for (int c = 0; c < OCMPS_MAX_Q; ++c) {
      blk_of_q[c] = -1;
      if (cnt_v[c] > 0 && cnt_c[c] > 0) {
        if (nb < OCMPS_MAX_BLK) {
          DecompBlock& B = w->blk[nb];
          B.q = c; B.nv = cnt_v[c]; B.len = cnt_c[c];
          B.vec_off = ov; B.comp_off = oc; B.ws_off = ows; B.p_off = ov;
          blk_of_q[c] = nb;
          off_v[c] = ov; off_c[c] = oc;
          ov += cnt_v[c]; oc += cnt_c[c]; ows += cnt_v[c] * cnt_c[c];
          ++nb;
        } else {
          atomicOr(b.status, OCMPS_ST_TOOMANYBLK);
        }
      }
    }
    w->n = g.n; w->m = g.m; w->ld = g.m; w->mode = g.mode;
    w->nblocks = nb; w->nvtot = ov; w->newdim = 0; w->norm_count = 0;
  }
  __syncthreads();
  // comps / vectors that belong to no block
  for (int i = q; i < ncomp; i += blockDim.x) { comp_blk[i] = -1; comp_rank[i] = 0; }
  __syncthreads();
  const int mb = blk_of_q[q];
  if (mb >= 0) {
    int ov = off_v[q], oc = off_c[q], kv = 0, kc = 0;
    for (int i = 0; i < g.n; ++i) {
      if (row_charge(a, i) == q) {
        if (g.mode == 0) { b.comp_idx[oc + kc] = i; comp_blk[i] = mb; comp_rank[i] = kc; ++kc; }
        else { b.vec_idx[ov + kv] = i; b.vecq[ov + kv] = q; ++kv; }
      }
    }
    for (int j = 0; j < g.m; ++j) {
      if (col_charge(a, j, chiR) == q) {
        if (g.mode == 0) { b.vec_idx[ov + kv] = j; b.vecq[ov + kv] = q; ++kv; }
        else { b.comp_idx[oc + kc] = j; comp_blk[j] = mb; comp_rank[j] = kc; ++kc; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// one-sided Jacobi on one charge block per CTA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(JAC_THREADS) jacobi_blocks_kernel(DecompArgs a, DecompBuffers b, int smem_elems) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_rot;
  __shared__ double s_part[JAC_THREADS / 32];
  __shared__ double s_thr;
  const DecompWork* w = b.dw;
  if ((int)blockIdx.x >= w->nblocks) return;
  const DecompBlock B = w->blk[blockIdx.x];
  const int nv = B.nv, len = B.len, ld = w->ld, mode = w->mode;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = JAC_THREADS / 32;
  cplx* Yg = b.ywork + B.ws_off;
  const bool in_smem = nv * len <= smem_elems;
  cplx* Y = in_smem ? reinterpret_cast<cplx*>(smem_raw) : Yg;
  const int* vidx = b.vec_idx + B.vec_off;
  const int* cidx = b.comp_idx + B.comp_off;

  // gather the block: Y[v][c]
  if (mode == 0) {   // vectors are columns: consecutive threads take consecutive vectors (coalesced over columns)
    for (int e = tid; e < nv * len; e += JAC_THREADS) {
      int c = e / nv, v = e % nv;
      Y[v * len + c] = a.X[(size_t)cidx[c] * ld + vidx[v]];
    }
  } else {
    for (int e = tid; e < nv * len; e += JAC_THREADS) {
      int v = e / len, c = e % len;
      Y[v * len + c] = a.X[(size_t)vidx[v] * ld + cidx[c]];
    }
  }
  __syncthreads();

  // deflation threshold: vectors whose squared norm falls below 1e-30 of the block's Frobenius norm are
  // numerically zero (a set of nv > rank vectors can only become mutually orthogonal if the surplus is 0)
  {
    double f = 0.0;
    for (int e = tid; e < nv * len; e += JAC_THREADS) { cplx u = Y[e]; f += u.x * u.x + u.y * u.y; }
    f = warp_sum(f);
    if (lane == 0) s_part[warp] = f;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < nwarps; ++i) t += s_part[i];
      s_thr = t * 1e-30;
    }
    __syncthreads();
  }
  const double thr = s_thr;

  const int npad = (nv + 1) & ~1;
  bool converged = (nv < 2);
  for (int sweep = 0; sweep < JAC_MAX_SWEEPS && !converged; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int r = 0; r < npad - 1; ++r) {
      for (int k = warp; k < npad / 2; k += nwarps) {
        int p = (r + k) % (npad - 1);
        int q = (k == 0) ? (npad - 1) : (r + npad - 1 - k) % (npad - 1);
        if (p >= nv || q >= nv) continue;
        if (p > q) { int tmp = p; p = q; q = tmp; }
        cplx* yp = Y + p * len;
        cplx* yq = Y + q * len;
        double aa = 0.0, bb = 0.0, cre = 0.0, cim = 0.0;
        for (int c = lane; c < len; c += 32) {
          cplx u = yp[c], v = yq[c];
          aa += u.x * u.x + u.y * u.y;
          bb += v.x * v.x + v.y * v.y;
          cre += u.x * v.x + u.y * v.y;      // conj(u) * v
          cim += u.x * v.y - u.y * v.x;
        }
        aa = warp_sum(aa); bb = warp_sum(bb); cre = warp_sum(cre); cim = warp_sum(cim);
        const double c2 = cre * cre + cim * cim;
        if (aa <= thr || bb <= thr) {
          if (aa <= thr && aa > 0.0) for (int c = lane; c < len; c += 32) yp[c] = make_double2(0.0, 0.0);
          if (bb <= thr && bb > 0.0) for (int c = lane; c < len; c += 32) yq[c] = make_double2(0.0, 0.0);
        } else if (c2 > JAC_TOL2 * aa * bb) {
          const double cabs = sqrt(c2);
          const double zeta = (bb - aa) / (2.0 * cabs);
          const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + tt * tt);
          const double sn = cs * tt;
          const double phr = cre / cabs, phi = cim / cabs;       // e^{i phi}
          // yp' = cs yp - sn conj(ph) yq ;  yq' = sn ph yp + cs yq
          const double s1r = sn * phr, s1i = sn * phi;
          for (int c = lane; c < len; c += 32) {
            cplx u = yp[c], v = yq[c];
            cplx nu, nvv;
            nu.x = cs * u.x - (s1r * v.x + s1i * v.y);
            nu.y = cs * u.y - (s1r * v.y - s1i * v.x);
            nvv.x = (s1r * u.x - s1i * u.y) + cs * v.x;
            nvv.y = (s1r * u.y + s1i * u.x) + cs * v.y;
            yp[c] = nu; yq[c] = nvv;
          }
          if (lane == 0) s_rot = 1;
        }
      }
      __syncthreads();
    }
    converged = (s_rot == 0);
    __syncthreads();
    if (!converged && sweep == JAC_MAX_SWEEPS - 1 && tid == 0) atomicOr(b.status, OCMPS_ST_NOCONV);
  }

  // spectrum + normalised vectors
  for (int v = warp; v < nv; v += nwarps) {
    cplx* y = Y + v * len;
    double s = 0.0;
    for (int c = lane; c < len; c += 32) { cplx u = y[c]; s += u.x * u.x + u.y * u.y; }
    s = warp_sum(s);
    const double inv = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
    for (int c = lane; c < len; c += 32) {
      cplx u = y[c];
      Yg[v * len + c] = make_double2(u.x * inv, u.y * inv);
    }
    if (lane == 0) b.P[B.p_off + v] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// global truncation (ITensor truncate()) + new bond bookkeeping + follow-up descriptors
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) truncate_kernel(DecompArgs a, DecompBuffers b, TruncParams tp) {
  __shared__ double sP[NV_MAX];
  __shared__ double sSorted[NV_MAX];
  __shared__ int sQ[NV_MAX];
  __shared__ unsigned char sKeep[NV_MAX];
  __shared__ double s_docut;
  __shared__ int s_total;
  DecompWork* w = b.dw;
  const int nv = w->nvtot;
  const int tid = threadIdx.x;
  for (int i = tid; i < nv; i += blockDim.x) { sP[i] = b.P[i]; sQ[i] = b.vecq[i]; }
  if (tid == 0) s_total = 0;
  __syncthreads();
  for (int i = tid; i < nv; i += blockDim.x) {
    const double pi = sP[i];
    int rank = 0;
    for (int j = 0; j < nv; ++j) {
      const double pj = sP[j];
      rank += (pj > pi) || (pj == pi && j < i);
    }
    sSorted[rank] = pi;
  }
  __syncthreads();
  if (tid == 0) {
    double docut = 0.0;
    const int origm = nv;
    if (origm == 1) {
      docut = sSorted[0] / 2.0;
    } else if (origm > 1) {
      int n = origm - 1;
      double truncerr = 0.0;
      while (n >= tp.maxm) { truncerr += sSorted[n]; --n; }
      double scale = 1.0;
      if (tp.rel_cutoff) {
        double sum = 0.0;
        for (int i = 0; i < origm; ++i) sum += sSorted[i];
        scale = (sum == 0.0) ? 1.0 : sum;
      }
      while (n >= tp.minm && truncerr + sSorted[n] < tp.cutoff * scale) { truncerr += sSorted[n]; --n; }
      if (n < 0) n = 0;
      const int m = n + 1;
      if (m < origm) {
        docut = (sSorted[m] + sSorted[m - 1]) / 2.0;
        if (fabs(sSorted[m] - sSorted[m - 1]) < 1e-3 * sSorted[m - 1]) docut += 1e-3 * sSorted[m - 1];
      }
    }
    s_docut = docut;
  }
  __syncthreads();
  const double docut = s_docut;
  for (int i = tid; i < nv; i += blockDim.x) {
    const unsigned char k = sP[i] > docut;
    sKeep[i] = k;
    if (k) atomicAdd(&s_total, 1);
  }
  __syncthreads();
  if (s_total == 0 && nv > 0) {        // zero tensor: keep one arbitrary state
    if (tid == 0) { sKeep[0] = 1; s_total = 1; }
    __syncthreads();
  }
  const int total = s_total;
  int* inv_blk = b.pos + NV_MAX;       // per new index: block id, vector-in-block
  int* inv_v = b.pos + 2 * NV_MAX;
  for (int i = tid; i < nv; i += blockDim.x) {
    int pos = -1;
    if (sKeep[i]) {
      pos = 0;
      const int qi = sQ[i];
      const double pi = sP[i];
      for (int j = 0; j < nv; ++j) {
        if (!sKeep[j]) continue;
        const int qj = sQ[j];
        const double pj = sP[j];
        pos += (qj < qi) || (qj == qi && (pj > pi || (pj == pi && j < i)));
      }
      if (pos < tp.cap) {
        a.qNew[pos] = qi;
        // block of vector i: scan the (short) block table
        int bi = 0;
        while (bi + 1 < w->nblocks && w->blk[bi + 1].p_off <= i) ++bi;
        inv_blk[pos] = bi;
        inv_v[pos] = i - w->blk[bi].p_off;
      }
    }
    b.pos[i] = pos;
  }
  if (tid == 0) {
    int k = total;
    if (k > tp.cap) { atomicOr(b.status, OCMPS_ST_CAPACITY); k = tp.cap; }
    w->newdim = k;
    *a.dimNew = k;
    const int n = w->n, m = w->m;
    GemmDesc g0, g1;
    g0.pad = g1.pad = 0;
    g1.A = g1.B = nullptr; g1.C = nullptr; g1.M = g1.N = g1.K = 0; g1.lda = g1.ldb = g1.ldc = 1; g1.opA = g1.opB = 0;
    switch (a.kind) {
      case DK_GATE_LEFT:   // partner (k x m) = iso^H (n x k) . X (n x m)
      case DK_ORTH_LEFT:
        g0.A = a.iso; g0.opA = 1; g0.lda = k;
        g0.B = a.X; g0.opB = 0; g0.ldb = m;
        g0.C = a.partner; g0.ldc = m; g0.M = k; g0.N = m; g0.K = n;
        w->norm_count = k * m;
        if (a.kind == DK_ORTH_LEFT) {     // neighbour (k x D*chiFar) = C (k x m) . nb_in (m x D*chiFar)
          const int far = a.D * (*a.dimNb);
          g1.A = a.partner; g1.opA = 0; g1.lda = m;
          g1.B = a.nb_in; g1.opB = 0; g1.ldb = far;
          g1.C = a.nb_out; g1.ldc = far; g1.M = k; g1.N = far; g1.K = m;
        }
        break;
      default:             // partner (n x k) = X (n x m) . iso^H (k x m)^H
        g0.A = a.X; g0.opA = 0; g0.lda = m;
        g0.B = a.iso; g0.opB = 1; g0.ldb = m;
        g0.C = a.partner; g0.ldc = k; g0.M = n; g0.N = k; g0.K = m;
        w->norm_count = n * k;
        if (a.kind == DK_ORTH_RIGHT) {    // neighbour (chiFar*D x k) = nb_in (chiFar*D x n) . C (n x k)
          const int far = a.D * (*a.dimNb);
          g1.A = a.nb_in; g1.opA = 0; g1.lda = n;
          g1.B = a.partner; g1.opB = 0; g1.ldb = k;
          g1.C = a.nb_out; g1.ldc = k; g1.M = far; g1.N = k; g1.K = n;
        }
        break;
    }
    b.descs[0] = g0;
    b.descs[1] = g1;
  }
}

// isometry assembly: every output element looks up its source (no zero-fill pass, coalesced writes)
__global__ void scatter_iso_kernel(DecompArgs a, DecompBuffers b) {
  const DecompWork* w = b.dw;
  const int k = w->newdim, n = w->n, m = w->m, mode = w->mode;
  const int* comp_blk = b.comp_idx + NV_MAX;
  const int* comp_rank = b.comp_idx + 2 * NV_MAX;
  const int* inv_blk = b.pos + NV_MAX;
  const int* inv_v = b.pos + 2 * NV_MAX;
  const long long total = (long long)(mode == 0 ? n : m) * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int comp, kk;
    if (mode == 0) { comp = (int)(e / k); kk = (int)(e % k); }      // iso[n][k]
    else { kk = (int)(e / m); comp = (int)(e % m); }                // iso[k][m]
    cplx v = make_double2(0.0, 0.0);
    const int bi = inv_blk[kk];
    if (comp_blk[comp] == bi) {
      const DecompBlock& B = w->blk[bi];
      v = b.ywork[B.ws_off + inv_v[kk] * B.len + comp_rank[comp]];
    }
    a.iso[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Frobenius normalisation, deterministic two-stage reduction
// ------------------------------------------------------------------------------------------------
constexpr int NORM_CTAS = 32;
constexpr int NORM_THREADS = 256;

__device__ __forceinline__ void norm_partial_body(const cplx* x, long long count, double* partial) {
  __shared__ double red[NORM_THREADS];
  double s = 0.0;
  for (long long e = blockIdx.x * (long long)NORM_THREADS + threadIdx.x; e < count; e += (long long)NORM_CTAS * NORM_THREADS) {
    cplx u = x[e];
    s += u.x * u.x + u.y * u.y;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = NORM_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__device__ __forceinline__ double norm_total(const double* partial) {
  double s = 0.0;
  for (int i = 0; i < NORM_CTAS; ++i) s += partial[i];
  return s;
}

__global__ void __launch_bounds__(NORM_THREADS) norm_partial_dw_kernel(const cplx* x, const DecompWork* w, double* partial) {
  norm_partial_body(x, w->norm_count, partial);
}
__global__ void scale_dw_kernel(cplx* x, const DecompWork* w, const double* partial) {
  const double nrm = sqrt(norm_total(partial));
  if (!(nrm > 1e-16)) return;          // src/BH_tDMRG.cpp:184
  const double inv = 1.0 / nrm;
  const long long count = w->norm_count;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
    cplx u = x[e];
    x[e] = make_double2(u.x * inv, u.y * inv);
  }
}
__global__ void __launch_bounds__(NORM_THREADS) norm_partial_site_kernel(const cplx* x, const int* dimL, const int* dimR, int D,
                                                                        double* partial) {
  norm_partial_body(x, (long long)(*dimL) * D * (*dimR), partial);
}
__global__ void scale_site_kernel(cplx* x, const int* dimL, const int* dimR, int D, const double* partial) {
  const double nrm = sqrt(norm_total(partial));
  if (!(nrm > 0.0)) return;
  const double inv = 1.0 / nrm;
  const long long count = (long long)(*dimL) * D * (*dimR);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
    cplx u = x[e];
    x[e] = make_double2(u.x * inv, u.y * inv);
  }
}
__global__ void norm_finish_kernel(const double* partial, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = sqrt(norm_total(partial));
}

int grid_for(long long elems, int threads, int cap = 1184) {
  long long g = (elems + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

static bool g_jac_attr_set[64] = {false};

void launch_decomp_setup(const DecompArgs& a, const DecompBuffers& b, cudaStream_t s) {
  decomp_setup_kernel<<<1, OCMPS_MAX_Q, 0, s>>>(a, b);
}

void launch_jacobi_blocks(const DecompArgs& a, const DecompBuffers& b, int nblk_launch, size_t smem_limit, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !g_jac_attr_set[dev]) {
    cudaFuncSetAttribute(jacobi_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    g_jac_attr_set[dev] = true;
  }
  jacobi_blocks_kernel<<<nblk_launch, JAC_THREADS, smem_limit, s>>>(a, b, (int)(smem_limit / sizeof(cplx)));
}

void launch_truncate(const DecompArgs& a, const DecompBuffers& b, const TruncParams& tp, cudaStream_t s) {
  truncate_kernel<<<1, 1024, 0, s>>>(a, b, tp);
}

void launch_scatter_iso(const DecompArgs& a, const DecompBuffers& b, int max_elems, cudaStream_t s) {
  scatter_iso_kernel<<<grid_for(max_elems, 256), 256, 0, s>>>(a, b);
}

void launch_normalize(cplx* x, const DecompBuffers& b, int max_elems, cudaStream_t s) {
  norm_partial_dw_kernel<<<NORM_CTAS, NORM_THREADS, 0, s>>>(x, b.dw, b.partial);
  scale_dw_kernel<<<grid_for(max_elems, 256), 256, 0, s>>>(x, b.dw, b.partial);
}

void launch_norm_only(const cplx* x, const int* dimL, const int* dimR, int D, double* partial, double* out, int max_elems,
                      cudaStream_t s) {
  (void)max_elems;
  norm_partial_site_kernel<<<NORM_CTAS, NORM_THREADS, 0, s>>>(x, dimL, dimR, D, partial);
  norm_finish_kernel<<<1, 32, 0, s>>>(partial, out);
}

void launch_normalize_site(cplx* x, const int* dimL, const int* dimR, int D, double* partial, int max_elems, cudaStream_t s) {
  norm_partial_site_kernel<<<NORM_CTAS, NORM_THREADS, 0, s>>>(x, dimL, dimR, D, partial);
  scale_site_kernel<<<grid_for(max_elems, 256), 256, 0, s>>>(x, dimL, dimR, D, partial);
}
