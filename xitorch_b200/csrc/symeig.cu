// Block Rayleigh-Ritz eigensolver on a growing Krylov subspace for a dense Hermitian operator --
// B200-native restatement of davidson (xitorch/_impls/linalg/symeig.py:100-227) and its orthogonaliser
// tallqr (xitorch/_utils/tensor.py:8-19), plus the block-Lanczos variant ("lanczos", BASELINE.json C5).
//
// One iteration = one subspace expansion:
//   1. W = A Q_j                      one pass over A (matvec.cu)                      [symeig.py:221]
//   2. C = V^T W                      new block column of T = V^T A V (fp64 accumulate) [symeig.py:170, incremental]
//   3. eigh(T), keep k extreme pairs  one-CTA cyclic Jacobi in fp64                      [symeig.py:174-175]
//   4. X = V S, R = AV S - X Lambda,  max|R| -> stop test / best-pair bookkeeping        [symeig.py:178-199]
//   5. expansion block Z: R (expansion=0, the reference's Davidson step, symeig.py:207) or W (expansion=1,
//      block Lanczos: same Krylov space), orthogonalised against V by block classical Gram-Schmidt
//      (twice for W) and orthonormalised by Cholesky-QR with an fp64 Gram matrix     [symeig.py:210-220, tensor.py:8-19]
// Differences from the reference, all result-preserving: T and the basis are updated incrementally (old
// basis vectors are not re-orthonormalised every iteration), Gram/projection matrices are accumulated in
// fp64 (the reference's fp32 tallqr breaks down), the subspace is thick-restarted when it
// reaches `max_basis`, and convergence is tested on the device (no per-iteration host sync).
//
// Layout in HBM: basis V and AV as blocks [block][n][k] (each block is an (n,k) row-major array, so
// block j is directly the X / Y operand of the matvec kernel); T, S row-major fp64.
#include "matvec.cuh"

#include <cstring>
#include <cmath>

namespace xt {

constexpr int SE_MAXK = 16;
constexpr int SE_THREADS = 256;
constexpr int SE_ROWS = 64;         // rows per CTA chunk in the tall-skinny kernels
constexpr int EIG_THREADS = 1024;

struct EigCtl {
  int done;
  int converged;
  int breakdown;
  int niter;
  int best_slot;        // which X / evals slot holds the best pair so far
  unsigned int counter;
  unsigned int resmax_bits;   // max |R| of the current iteration (float bits, atomicMax)
  float best_resid;
};

// ---------------------------------------------------------------------------- tall-skinny kernels
// Zp = Z - V Cin  (Cin may be null);  Cout += V^T Zp (m x k);  G += Zp^T Zp (k x k).  Zp is written
// to Zout when Zout != nullptr.  V: m basis vectors in blocks of k.  fp64 accumulation, fp64 atomics.
template <typename TV>
__global__ void __launch_bounds__(SE_THREADS)
subproj_kernel(const TV* __restrict__ V, int n, int k, int m, const TV* __restrict__ Z, const double* __restrict__ Cin,
               TV* __restrict__ Zout, double* __restrict__ Cout, double* __restrict__ G, const EigCtl* ctl) {
  if (ctl->done) return;
  __shared__ double Zs[SE_ROWS * SE_MAXK];
  __shared__ double Vs[SE_ROWS * SE_MAXK];
  __shared__ double Cs[SE_MAXK * SE_MAXK];
  const int row0 = blockIdx.x * SE_ROWS;
  const int rows = min(SE_ROWS, n - row0);
  const int tid = threadIdx.x;
  const int nblk = m / k;
  // load the Z chunk
  for (int i = tid; i < rows * k; i += SE_THREADS) Zs[i] = (double)Z[(int64_t)row0 * k + i];
  __syncthreads();
  // subtract V Cin
  if (Cin != nullptr) {
    for (int blk = 0; blk < nblk; ++blk) {
      const TV* Vb = V + ((int64_t)blk * n + row0) * k;
      for (int i = tid; i < rows * k; i += SE_THREADS) Vs[i] = (double)Vb[i];
      for (int i = tid; i < k * k; i += SE_THREADS) Cs[i] = Cin[(int64_t)blk * k * k + i];   // rows blk*k.. of Cin (m x k)
      __syncthreads();
      for (int e = tid; e < rows * k; e += SE_THREADS) {
        const int r = e / k, j = e - r * k;
        double acc = 0.0;
        for (int i = 0; i < k; ++i) acc += Vs[r * k + i] * Cs[i * k + j];
        Zs[e] -= acc;
      }
      __syncthreads();
    }
  }
  if (Zout != nullptr)
    for (int i = tid; i < rows * k; i += SE_THREADS) Zout[(int64_t)row0 * k + i] = (TV)Zs[i];
  // the orthonormalisation works with the ROUNDED block (what is stored), so re-read the rounded values
  if (Zout != nullptr && sizeof(TV) < sizeof(double)) {
    __syncthreads();
    for (int i = tid; i < rows * k; i += SE_THREADS) Zs[i] = (double)(TV)Zs[i];
  }
  __syncthreads();
  // projections onto every basis block and the Gram matrix: thread <-> (i, j) pairs x row slices
  const int npair = k * k;
  const int nslice = SE_THREADS / npair > 0 ? SE_THREADS / npair : 1;
  const int pr = tid % npair, sl = tid / npair;
  const int pi = pr / k, pj = pr - pi * k;
  const bool worker = (npair <= SE_THREADS) ? (sl < nslice) : true;
  for (int blk = 0; blk <= nblk; ++blk) {
    const bool gram = (blk == nblk);
    if (!gram) {
      const TV* Vb = V + ((int64_t)blk * n + row0) * k;
      for (int i = tid; i < rows * k; i += SE_THREADS) Vs[i] = (double)Vb[i];
      __syncthreads();
    }
    const double* L = gram ? Zs : Vs;
    if (npair <= SE_THREADS) {
      if (worker) {
        double acc = 0.0;
        for (int r = sl; r < rows; r += nslice) acc += L[r * k + pi] * Zs[r * k + pj];
        if (gram) {
          if (G != nullptr) atomicAdd(&G[pi * k + pj], acc);
        } else if (Cout != nullptr) {
          atomicAdd(&Cout[((int64_t)blk * k + pi) * k + pj], acc);
        }
      }
    } else {   // k*k > threads cannot happen for k <= 16 with 256 threads; kept for safety
      for (int e = tid; e < npair; e += SE_THREADS) {
        const int i = e / k, j = e - i * k;
        double acc = 0.0;
        for (int r = 0; r < rows; ++r) acc += L[r * k + i] * Zs[r * k + j];
        if (gram) {
          if (G != nullptr) atomicAdd(&G[i * k + j], acc);
        } else if (Cout != nullptr) {
          atomicAdd(&Cout[((int64_t)blk * k + i) * k + j], acc);
        }
      }
    }
    __syncthreads();
  }
}

// Q = (Zp - V C2) Rinv with R = chol(G - C2^T C2)^T (upper), every CTA recomputes the tiny k x k factor.
// Writes Q to Qout (an (n,k) block).  Breakdown (non-positive pivot) -> ctl->breakdown = done = 1.
template <typename TV>
__global__ void __launch_bounds__(SE_THREADS)
orth_finish_kernel(const TV* __restrict__ V, int n, int k, int m, const TV* __restrict__ Zp,
                   const double* __restrict__ C2, const double* __restrict__ G, TV* __restrict__ Qout, EigCtl* ctl) {
  if (ctl->done) return;
  __shared__ double Zs[SE_ROWS * SE_MAXK];
  __shared__ double Vs[SE_ROWS * SE_MAXK];
  __shared__ double Cs[SE_MAXK * SE_MAXK];
  __shared__ double Gs[SE_MAXK * SE_MAXK];     // Gram -> L (lower Cholesky) -> Rinv
  __shared__ double Ri[SE_MAXK * SE_MAXK];
  __shared__ int bad;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * SE_ROWS;
  const int rows = min(SE_ROWS, n - row0);
  const int nblk = m / k;
  // Gs = G - C2^T C2
  for (int e = tid; e < k * k; e += SE_THREADS) {
    const int i = e / k, j = e - i * k;
    double acc = 0.5 * (G[i * k + j] + G[j * k + i]);
    for (int t = 0; t < m; ++t) acc -= C2[(int64_t)t * k + i] * C2[(int64_t)t * k + j];
    Gs[e] = acc;
  }
  if (tid == 0) bad = 0;
  __syncthreads();
  if (tid == 0) {
    // Cholesky G = L L^T (lower, in place), then Rinv = (L^T)^-1 = (L^-1)^T
    double scale = 0.0;
    for (int i = 0; i < k; ++i) scale = fmax(scale, Gs[i * k + i]);
    for (int j = 0; j < k && !bad; ++j) {
      double d = Gs[j * k + j];
      for (int t = 0; t < j; ++t) d -= Gs[j * k + t] * Gs[j * k + t];
      if (!(d > 1e-24 * scale) || !(d == d)) { bad = 1; break; }
      const double ljj = sqrt(d);
      Gs[j * k + j] = ljj;
      for (int i = j + 1; i < k; ++i) {
        double s = Gs[i * k + j];
        for (int t = 0; t < j; ++t) s -= Gs[i * k + t] * Gs[j * k + t];
        Gs[i * k + j] = s / ljj;
      }
    }
    if (!bad) {
      // Linv (lower) by forward substitution, stored transposed: Ri[c][r'] ... we need Rinv = Linv^T, i.e.
      // Q[:, j] = sum_i Z[:, i] * Rinv[i][j],  Rinv[i][j] = Linv[j][i]
      for (int c = 0; c < k; ++c) {          // column c of Linv
        for (int r = 0; r < k; ++r) {
          if (r < c) { Ri[c * k + r] = 0.0; continue; }   // Linv[r][c] = 0 for r < c ; store Rinv[c][r] = Linv[r][c]
          double s = (r == c) ? 1.0 : 0.0;
          for (int t = c; t < r; ++t) s -= Gs[r * k + t] * Ri[c * k + t];
          Ri[c * k + r] = s / Gs[r * k + r];
        }
      }
    }
  }
  __syncthreads();
  if (bad) {
    if (blockIdx.x == 0 && tid == 0) { ctl->breakdown = 1; ctl->done = 1; }
    return;
  }
  for (int i = tid; i < rows * k; i += SE_THREADS) Zs[i] = (double)Zp[(int64_t)row0 * k + i];
  __syncthreads();
  for (int blk = 0; blk < nblk; ++blk) {
    const TV* Vb = V + ((int64_t)blk * n + row0) * k;
    for (int i = tid; i < rows * k; i += SE_THREADS) Vs[i] = (double)Vb[i];
    for (int i = tid; i < k * k; i += SE_THREADS) Cs[i] = C2[(int64_t)blk * k * k + i];
    __syncthreads();
    for (int e = tid; e < rows * k; e += SE_THREADS) {
      const int r = e / k, j = e - r * k;
      double acc = 0.0;
      for (int i = 0; i < k; ++i) acc += Vs[r * k + i] * Cs[i * k + j];
      Zs[e] -= acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < rows * k; e += SE_THREADS) {
    const int r = e / k, j = e - r * k;
    double acc = 0.0;
    for (int i = 0; i <= j; ++i) acc += Zs[r * k + i] * Ri[i * k + j];
    Qout[(int64_t)row0 * k + e] = (TV)acc;
  }
}

// X = V S_k, R = AV S_k - X theta, max|R| -> ctl->resmax_bits.  The last CTA to finish does the
// reference's bookkeeping (symeig.py:196-201): best pair (slot flip) and the stop test.
template <typename TV>
__global__ void __launch_bounds__(SE_THREADS)
ritz_kernel(const TV* __restrict__ V, const TV* __restrict__ AV, int n, int k, int m, const double* __restrict__ Sk,
            const double* __restrict__ theta, TV* __restrict__ Xslots, double* __restrict__ evals_slots,
            TV* __restrict__ Rout, EigCtl* ctl, int iter, float min_eps) {
  if (ctl->done) return;
  extern __shared__ double sh[];
  double* Ss = sh;                 // [m][k]
  __shared__ float red[32];
  const int tid = threadIdx.x;
  for (int i = tid; i < m * k; i += SE_THREADS) Ss[i] = Sk[i];
  __syncthreads();
  const int slot = 1 - ctl->best_slot;
  TV* X = Xslots + (int64_t)slot * n * k;
  const int row0 = blockIdx.x * SE_ROWS;
  const int rows = min(SE_ROWS, n - row0);
  const int nblk = m / k;
  float lmax = 0.f;
  for (int e = tid; e < rows * k; e += SE_THREADS) {
    const int r = e / k, j = e - r * k;
    double x = 0.0, ax = 0.0;
    for (int blk = 0; blk < nblk; ++blk) {
      const TV* vr = V + ((int64_t)blk * n + row0 + r) * k;
      const TV* ar = AV + ((int64_t)blk * n + row0 + r) * k;
      for (int i = 0; i < k; ++i) {
        const double s = Ss[(blk * k + i) * k + j];
        x += (double)vr[i] * s;
        ax += (double)ar[i] * s;
      }
    }
    const double res = ax - x * theta[j];
    X[(int64_t)(row0 + r) * k + j] = (TV)x;
    Rout[(int64_t)(row0 + r) * k + j] = (TV)res;
    lmax = fmaxf(lmax, fabsf((float)res));
    if (!(res == res)) lmax = INFINITY;
  }
  lmax = block_max(lmax, red);
  if (tid == 0) {
    atomicMax(&ctl->resmax_bits, __float_as_uint(lmax));
    __threadfence();
    const unsigned int ticket = atomicAdd(&ctl->counter, 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      const float rmax = __uint_as_float(atomicAdd(&ctl->resmax_bits, 0u));
      ctl->niter = iter;
      for (int j = 0; j < k; ++j) evals_slots[slot * SE_MAXK + j] = theta[j];
      if (rmax < ctl->best_resid) {
        ctl->best_resid = rmax;
        ctl->best_slot = slot;
      }
      if (rmax < min_eps) {
        ctl->converged = 1;
        ctl->done = 1;
      }
      ctl->counter = 0;
      ctl->resmax_bits = 0;
      __threadfence();
    }
  }
}

// Out[:, 0..p) = In(:, 0..m) * S[:, order[first + c]]   (thick restart: rotate the basis onto kept Ritz vectors)
template <typename TV>
__global__ void __launch_bounds__(SE_THREADS)
rotate_kernel(const TV* __restrict__ In, int n, int k, int m, const double* __restrict__ S, int lds,
              const int* __restrict__ order, int first, int p, TV* __restrict__ Out, const EigCtl* ctl) {
  if (ctl->done) return;
  // one thread per (row, output column); S accessed through L1/L2 (m*p doubles, shared by all threads)
  const int64_t e = (int64_t)blockIdx.x * SE_THREADS + threadIdx.x;
  if (e >= (int64_t)n * p) return;
  const int64_t row = e / p;
  const int c = (int)(e - row * p);
  const int col = order[first + c];
  double acc = 0.0;
  const int nblk = m / k;
  for (int blk = 0; blk < nblk; ++blk) {
    const TV* vr = In + ((int64_t)blk * n + row) * k;
    for (int i = 0; i < k; ++i) acc += (double)vr[i] * S[(int64_t)(blk * k + i) * lds + col];
  }
  // output in block layout [c / k][row][c % k]
  Out[((int64_t)(c / k) * n + row) * k + (c % k)] = (TV)acc;
}

// T <- diag(w[order[first + i]]) after a restart
__global__ void restart_T_kernel(double* T, int ldt, const double* w, const int* order, int first, int p,
                                 const EigCtl* ctl) {
  if (ctl->done) return;
  for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
    const int i = e / p, j = e - i * p;
    T[(int64_t)i * ldt + j] = (i == j) ? w[order[first + i]] : 0.0;
  }
}

// ---------------------------------------------------------------------------- small dense eigh (one CTA)
// Cyclic parallel-ordered Jacobi in fp64 on Tw (m x m, row-major, leading dimension ld), eigenvectors
// accumulated in S (m x m, ld).  On exit w[i] = Tw[i][i] and order[] sorts them ascending.
__device__ void jacobi_eigh_device(double* Tw, double* S, int m, int ld, double* w, int* order, double* sc,
                                   float2* cs /* smem [m/2+1] */, int* prs /* smem [m+2] */) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e - i * m;
    S[(int64_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int M = (m + 1) & ~1;        // even number of players (index m is a dummy when m is odd)
  const int npairs = M / 2;
  double* cd = reinterpret_cast<double*>(cs);   // [npairs][2] doubles (c, s)
  for (int sweep = 0; sweep < 40; ++sweep) {
    // convergence: off-diagonal Frobenius norm relative to the diagonal
    double off = 0.0, dia = 0.0;
    for (int e = tid; e < m * m; e += nt) {
      const int i = e / m, j = e - i * m;
      const double v = Tw[(int64_t)i * ld + j];
      if (i == j) dia += v * v; else off += v * v;
    }
    off = block_sum(off, sc);
    dia = block_sum(dia, sc);
    if (off <= 1e-30 * dia || off == 0.0) break;
    for (int rd = 0; rd < M - 1; ++rd) {
      // pairing (round robin): player M-1 is fixed, the others rotate
      if (tid < npairs) {
        int p, q;
        if (tid == 0) { p = M - 1; q = rd; }
        else { p = (rd + tid) % (M - 1); q = (rd - tid + (M - 1)) % (M - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < m) {
          const double apq = Tw[(int64_t)p * ld + q];
          const double app = Tw[(int64_t)p * ld + p], aqq = Tw[(int64_t)q * ld + q];
          if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * (fabs(app) + fabs(aqq))) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        } else {
          q = -1;   // dummy pair
        }
        prs[2 * tid] = p;
        prs[2 * tid + 1] = q;
        cd[2 * tid] = c;
        cd[2 * tid + 1] = s;
      }
      __syncthreads();
      // rows p, q of Tw
      for (int e = tid; e < npairs * m; e += nt) {
        const int pr = e / m, j = e - pr * m;
        const int p = prs[2 * pr], q = prs[2 * pr + 1];
        const double s = cd[2 * pr + 1];
        if (q >= 0 && s != 0.0) {
          const double c = cd[2 * pr];
          const double tp = Tw[(int64_t)p * ld + j], tq = Tw[(int64_t)q * ld + j];
          Tw[(int64_t)p * ld + j] = c * tp - s * tq;
          Tw[(int64_t)q * ld + j] = s * tp + c * tq;
        }
      }
      __syncthreads();
      // columns p, q of Tw and S
      for (int e = tid; e < npairs * m; e += nt) {
        const int pr = e % npairs, i = e / npairs;
        const int p = prs[2 * pr], q = prs[2 * pr + 1];
        const double s = cd[2 * pr + 1];
        if (q >= 0 && s != 0.0) {
          const double c = cd[2 * pr];
          double tp = Tw[(int64_t)i * ld + p], tq = Tw[(int64_t)i * ld + q];
          Tw[(int64_t)i * ld + p] = c * tp - s * tq;
          Tw[(int64_t)i * ld + q] = s * tp + c * tq;
          tp = S[(int64_t)i * ld + p]; tq = S[(int64_t)i * ld + q];
          S[(int64_t)i * ld + p] = c * tp - s * tq;
          S[(int64_t)i * ld + q] = s * tp + c * tq;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < m; i += nt) w[i] = Tw[(int64_t)i * ld + i];
  __syncthreads();
  for (int i = tid; i < m; i += nt) {
    const double wi = w[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const double wj = w[j];
      rank += (wj < wi || (wj == wi && j < i)) ? 1 : 0;
    }
    order[rank] = i;
  }
  __syncthreads();
}

// T[:, new block] = C (and its transpose), Tw = T, eigh, select k extreme pairs -> theta (k), Sk (m x k)
__global__ void __launch_bounds__(EIG_THREADS)
rr_kernel(double* T, int ldt, const double* C, int m, int k, double* Tw, double* S, double* w, int* order,
          double* Sk, double* theta, int mode, int have_new_block, const EigCtl* ctl) {
  if (ctl->done) return;
  __shared__ double sc[64];
  extern __shared__ double dyn[];
  float2* cs = reinterpret_cast<float2*>(dyn);                       // (m/2+1) * 2 doubles
  int* prs = reinterpret_cast<int*>(dyn + 2 * (m / 2 + 2));          // m + 2 ints
  const int tid = threadIdx.x, nt = blockDim.x;
  if (have_new_block) {
    const int c0 = m - k;
    for (int e = tid; e < m * k; e += nt) {
      const int i = e / k, j = e - i * k;
      double v = C[e];
      if (i >= c0) v = 0.5 * (C[e] + C[(int64_t)(c0 + j) * k + (i - c0)]);   // symmetrise the diagonal block
      T[(int64_t)i * ldt + c0 + j] = v;
      T[(int64_t)(c0 + j) * ldt + i] = v;
    }
    __syncthreads();
  }
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e - i * m;
    Tw[(int64_t)i * ldt + j] = T[(int64_t)i * ldt + j];
  }
  __syncthreads();
  jacobi_eigh_device(Tw, S, m, ldt, w, order, sc, cs, prs);
  const int first = (mode == 0) ? 0 : (m - k);
  for (int e = tid; e < m * k; e += nt) {
    const int i = e / k, j = e - i * k;
    Sk[e] = S[(int64_t)i * ldt + order[first + j]];
  }
  for (int j = tid; j < k; j += nt) theta[j] = w[order[first + j]];
}

__global__ void __launch_bounds__(EIG_THREADS)
small_eigh_kernel(double* T, int m, double* w_sorted, double* S_sorted, double* S, double* w, int* order) {
  __shared__ double sc[64];
  extern __shared__ double dyn[];
  float2* cs = reinterpret_cast<float2*>(dyn);
  int* prs = reinterpret_cast<int*>(dyn + 2 * (m / 2 + 2));
  jacobi_eigh_device(T, S, m, m, w, order, sc, cs, prs);
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e - i * m;
    S_sorted[e] = S[(int64_t)i * m + order[j]];
  }
  for (int j = threadIdx.x; j < m; j += blockDim.x) w_sorted[j] = w[order[j]];
}

// final copy of the best pair into the caller's tensors
template <typename TV>
__global__ void output_kernel(const TV* Xslots, const double* evals_slots, int n, int k, TV* evecs, int64_t ldv,
                              TV* evals, const EigCtl* ctl) {
  const int slot = ctl->best_slot;
  const TV* X = Xslots + (int64_t)slot * n * k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)n * k;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / k;
    const int j = (int)(e - row * k);
    evecs[row * ldv + j] = X[e];
  }
  if (blockIdx.x == 0 && threadIdx.x < k) evals[threadIdx.x] = (TV)evals_slots[slot * SE_MAXK + threadIdx.x];
}

template <typename TV>
__global__ void gather_block_kernel(const TV* src, int64_t ld, int n, int k, TV* dst) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)n * k;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / k;
    dst[e] = src[row * ld + (e - row * k)];
  }
}

__global__ void init_ctl_kernel(EigCtl* ctl) {
  ctl->done = 0; ctl->converged = 0; ctl->breakdown = 0; ctl->niter = 0; ctl->best_slot = 0;
  ctl->counter = 0; ctl->resmax_bits = 0; ctl->best_resid = INFINITY;
}

// ============================================================================ host driver
struct EigWs {
  void *V, *AV, *Zbuf, *Rblk, *Xslots, *Vtmp;
  double *T, *Tw, *S, *w, *Sk, *theta, *C, *C2, *G, *evals_slots;
  int* order;
  EigCtl* ctl;
};

static bool carve(Arena& ar, EigWs& W, size_t vs, int n, int k, int mb) {
  const size_t blk = (size_t)n * k * vs;
  W.V = ar.take<char>((size_t)(mb / k) * blk);
  W.AV = ar.take<char>((size_t)(mb / k) * blk);
  W.Zbuf = ar.take<char>(blk);
  W.Rblk = ar.take<char>(blk);
  W.Xslots = ar.take<char>(2 * blk);
  W.Vtmp = ar.take<char>((size_t)(mb / k) * blk);
  W.T = ar.take<double>((size_t)mb * mb);
  W.Tw = ar.take<double>((size_t)mb * mb);
  W.S = ar.take<double>((size_t)mb * mb);
  W.w = ar.take<double>(mb);
  W.Sk = ar.take<double>((size_t)mb * k);
  W.theta = ar.take<double>(SE_MAXK);
  W.C = ar.take<double>((size_t)mb * k);
  W.C2 = ar.take<double>((size_t)mb * k);
  W.G = ar.take<double>(2 * SE_MAXK * SE_MAXK);
  W.evals_slots = ar.take<double>(2 * SE_MAXK);
  W.order = ar.take<int>(mb);
  W.ctl = ar.take<EigCtl>(1);
  return ar.ok();
}

template <typename TV> static int run_symeig(const xt_symeig_args* g) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(g->stream);
  const int n = g->n, k = g->neig;
  int mb = g->max_basis;
  if (mb > n) mb = (n / k) * k;
  mb = (mb / k) * k;
  XT_REQUIRE(mb >= 2 * k, "symeig: max_basis=%d too small for neig=%d (n=%d)", mb, k, n);
  Arena ar(g->workspace, g->workspace_bytes);
  EigWs W;
  if (!carve(ar, W, sizeof(TV), n, k, mb)) {
    set_last_error("symeig: workspace too small (%zu needed, %zu given)", ar.off, ar.cap);
    return XT_ERR_WORKSPACE;
  }
  TV* V = static_cast<TV*>(W.V);
  TV* AV = static_cast<TV*>(W.AV);
  TV* Zbuf = static_cast<TV*>(W.Zbuf);
  TV* Rblk = static_cast<TV*>(W.Rblk);
  TV* Xslots = static_cast<TV*>(W.Xslots);
  TV* Vtmp = static_cast<TV*>(W.Vtmp);
  const int64_t blk = (int64_t)n * k;
  const int grid_rows = (n + SE_ROWS - 1) / SE_ROWS;
  const int ce = g->check_every > 0 ? g->check_every : 1;
  const int keep = ((mb / 2) / k) * k >= k ? ((mb / 2) / k) * k : k;   // Ritz vectors kept at a restart
  const size_t eig_smem = (size_t)(2 * (mb / 2 + 2)) * sizeof(double) + (size_t)(mb + 2) * sizeof(int) + 64;

  int64_t napply = 0;
  int all_conv = 1;
  double worst_resid = 0.0;
  int last_niter = 0;

  for (int b = 0; b < g->nbatch; ++b) {
    const void* Ab = static_cast<const char*>(g->A) +
                     (size_t)b * g->a_bstride * (g->dtype == XT_F32 ? 4 : (g->dtype == XT_BF16 ? 2 : 8));
    init_ctl_kernel<<<1, 1, 0, st>>>(W.ctl); XT_LAUNCHED();
    // ---- orthonormalise the start block (Cholesky-QR twice; tensor.py:8-19 / symeig.py:249-252)
    gather_block_kernel<TV><<<grid_rows, 256, 0, st>>>(static_cast<const TV*>(g->V0) + (int64_t)b * g->v0_bstride,
                                                       g->ldv0, n, k, Rblk); XT_LAUNCHED();
    for (int pass = 0; pass < 2; ++pass) {
      XT_CUDA_OK(cudaMemsetAsync(W.G, 0, sizeof(double) * SE_MAXK * SE_MAXK, st));
      subproj_kernel<TV><<<grid_rows, SE_THREADS, 0, st>>>(V, n, k, 0, pass == 0 ? Rblk : V, nullptr, Zbuf, nullptr,
                                                            W.G, W.ctl); XT_LAUNCHED();
      orth_finish_kernel<TV><<<grid_rows, SE_THREADS, 0, st>>>(V, n, k, 0, Zbuf, W.C2, W.G, V, W.ctl); XT_LAUNCHED();
    }
    XT_CUDA_OK(cudaGetLastError());

    int m = k;          // current basis size
    int iter = 0;
    bool stop = false;
    while (!stop) {
      ++iter;
      const int j = m / k - 1;     // newest block
      // 1. W = A Q_j
      MvArgs a;
      memset(&a, 0, sizeof(a));
      a.dtype = g->dtype;
      a.nbatch = 1; a.nrows = n; a.ncolsA = n; a.k = k;
      a.A = Ab; a.lda = g->lda; a.a_bstride = 0;
      a.X = V + j * blk; a.ldx = k; a.x_bstride = 0;
      a.Y = AV + j * blk; a.ldy = k; a.y_bstride = 0;
      a.done_flag = &W.ctl->done;
      int rc = mv_launch(a, st);
      if (rc != XT_OK) return rc;
      ++napply;
      // 2. C = V^T W  (new block column of T)
      XT_CUDA_OK(cudaMemsetAsync(W.C, 0, sizeof(double) * (size_t)m * k, st));
      subproj_kernel<TV><<<grid_rows, SE_THREADS, 0, st>>>(V, n, k, m, AV + j * blk, nullptr, nullptr, W.C, nullptr,
                                                            W.ctl); XT_LAUNCHED();
      // 3. Rayleigh-Ritz on T
      rr_kernel<<<1, EIG_THREADS, eig_smem, st>>>(W.T, mb, W.C, m, k, W.Tw, W.S, W.w, W.order, W.Sk, W.theta, g->mode,
                                                  1, W.ctl); XT_LAUNCHED();
      // 4. Ritz vectors, residual, bookkeeping
      ritz_kernel<TV><<<grid_rows, SE_THREADS, (size_t)m * k * sizeof(double), st>>>(
          V, AV, n, k, m, W.Sk, W.theta, Xslots, W.evals_slots, Rblk, W.ctl, iter, (float)g->min_eps); XT_LAUNCHED();
      XT_CUDA_OK(cudaGetLastError());
      if (iter >= g->max_niter) break;
      if (m + k > n) break;                      // the basis cannot grow any further (symeig.py:202-203)
      if (iter % ce == 0) {
        int done = 0;
        XT_CUDA_OK(cudaMemcpyAsync(&done, &W.ctl->done, sizeof(int), cudaMemcpyDeviceToHost, st));
        XT_CUDA_OK(cudaStreamSynchronize(st));
        if (done) break;
      }
      // 5. expansion block, orthogonalised against V
      XT_CUDA_OK(cudaMemsetAsync(W.C2, 0, sizeof(double) * (size_t)m * k, st));
      XT_CUDA_OK(cudaMemsetAsync(W.G, 0, sizeof(double) * SE_MAXK * SE_MAXK, st));
      if (g->expansion == 1) {
        subproj_kernel<TV><<<grid_rows, SE_THREADS, 0, st>>>(V, n, k, m, AV + j * blk, W.C, Zbuf, W.C2, W.G, W.ctl);
      } else {
        subproj_kernel<TV><<<grid_rows, SE_THREADS, 0, st>>>(V, n, k, m, Rblk, nullptr, Zbuf, W.C2, W.G, W.ctl);
      }
      XT_LAUNCHED();
      if (m + k > mb) {
        // thick restart: finish the new block against the OLD basis first, then compress V / AV / T
        orth_finish_kernel<TV><<<grid_rows, SE_THREADS, 0, st>>>(V, n, k, m, Zbuf, W.C2, W.G, Rblk, W.ctl); XT_LAUNCHED();
        const int first = (g->mode == 0) ? 0 : (m - keep);
        const int64_t tot = (int64_t)n * keep;
        const int rg = (int)((tot + SE_THREADS - 1) / SE_THREADS);
        rotate_kernel<TV><<<rg, SE_THREADS, 0, st>>>(V, n, k, m, W.S, mb, W.order, first, keep, Vtmp, W.ctl); XT_LAUNCHED();
        XT_CUDA_OK(cudaMemcpyAsync(V, Vtmp, (size_t)tot * sizeof(TV), cudaMemcpyDeviceToDevice, st));
        rotate_kernel<TV><<<rg, SE_THREADS, 0, st>>>(AV, n, k, m, W.S, mb, W.order, first, keep, Vtmp, W.ctl); XT_LAUNCHED();
        XT_CUDA_OK(cudaMemcpyAsync(AV, Vtmp, (size_t)tot * sizeof(TV), cudaMemcpyDeviceToDevice, st));
        restart_T_kernel<<<1, 256, 0, st>>>(W.T, mb, W.w, W.order, first, keep, W.ctl); XT_LAUNCHED();
        XT_CUDA_OK(cudaMemcpyAsync(V + (int64_t)(keep / k) * blk, Rblk, (size_t)blk * sizeof(TV),
                                   cudaMemcpyDeviceToDevice, st));
        m = keep + k;
      } else {
        orth_finish_kernel<TV><<<grid_rows, SE_THREADS, 0, st>>>(V, n, k, m, Zbuf, W.C2, W.G, V + (int64_t)(m / k) * blk,
                                                                  W.ctl); XT_LAUNCHED();
        m += k;
      }
      XT_CUDA_OK(cudaGetLastError());
    }
    // ---- output
    output_kernel<TV><<<grid_rows, 256, 0, st>>>(Xslots, W.evals_slots, n, k,
                                                 static_cast<TV*>(g->evecs) + (int64_t)b * g->evecs_bstride, g->ldv,
                                                 static_cast<TV*>(g->evals) + (int64_t)b * g->evals_bstride, W.ctl); XT_LAUNCHED();
    EigCtl h;
    XT_CUDA_OK(cudaMemcpyAsync(&h, W.ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    XT_CUDA_OK(cudaStreamSynchronize(st));
    if (!h.converged) all_conv = 0;
    if (h.best_resid > worst_resid || !(h.best_resid == h.best_resid)) worst_resid = h.best_resid;
    last_niter = h.niter;
  }
  if (g->niter_out) *g->niter_out = last_niter;
  if (g->converged_out) *g->converged_out = all_conv;
  if (g->best_resid_out) *g->best_resid_out = worst_resid;
  if (g->napply_out) *g->napply_out = napply;
  return XT_OK;
}

}  // namespace xt

extern "C" {

size_t xt_symeig_workspace_bytes(int32_t dtype, int32_t n, int32_t neig, int32_t max_basis) {
  if (neig < 1 || n < 1) return 0;
  int mb = max_basis;
  if (mb > n) mb = (n / neig) * neig;
  mb = (mb / neig) * neig;
  if (mb < 2 * neig) return 0;
  xt::Arena ar(nullptr, 0);
  xt::EigWs W;
  xt::carve(ar, W, dtype == XT_F64 ? 8 : 4, n, neig, mb);
  return ar.off + 1024;
}

int xt_symeig_krylov(const xt_symeig_args* g) {
  XT_REQUIRE(g != nullptr, "symeig: null args");
  XT_REQUIRE(g->neig >= 1 && g->neig <= xt::SE_MAXK, "symeig: neig=%d outside 1..%d", g->neig, xt::SE_MAXK);
  XT_REQUIRE(g->n >= 2 * g->neig, "symeig: n=%d too small for neig=%d", g->n, g->neig);
  XT_REQUIRE(g->dtype == XT_F32 || g->dtype == XT_F64, "symeig: only fp32 / fp64 operators are supported");
  XT_REQUIRE(g->A && g->V0 && g->evals && g->evecs && g->workspace, "symeig: null pointer");
  XT_REQUIRE(g->mode == 0 || g->mode == 1, "symeig: mode must be 0 (lowest) or 1 (uppest)");
  XT_REQUIRE(g->max_basis <= 1024, "symeig: max_basis=%d exceeds 1024", g->max_basis);
  return g->dtype == XT_F64 ? xt::run_symeig<double>(g) : xt::run_symeig<float>(g);
}

int xt_small_eigh(double* T, int32_t m, double* w, double* S, void* stream) {
  // workspace-free test hook: w/S double as outputs; scratch is taken from the tail of S's caller buffer
  // layout expected from the caller: S has room for 2*m*m doubles, w for 2*m doubles + m ints (as doubles)
  XT_REQUIRE(T && w && S && m >= 1 && m <= 1024, "small_eigh: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* Sraw = S + (size_t)m * m;
  double* wraw = w + m;
  int* order = reinterpret_cast<int*>(w + 2 * m);
  const size_t smem = (size_t)(2 * (m / 2 + 2)) * sizeof(double) + (size_t)(m + 2) * sizeof(int) + 64;
  xt::small_eigh_kernel<<<1, xt::EIG_THREADS, smem, st>>>(T, m, w, S, Sraw, wraw, order); XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

}  // extern "C"
