// Conjugate gradient for problems that FIT ON CHIP -- the whole solve in ONE launch of one thread-block cluster
// (BASELINE configs[0]: cg on a 256 x 256 SPD fp64 operator; reference xitorch/_impls/linalg/solve.py:69-190).
// The general path is two launches per iteration (block matvec + vector step, ~15 us per iteration at this size, all of
// it launch latency: 512 KiB of A fits the shared memory of four SMs).  Here 8 CTAs of a cluster each keep n/8 rows of A
// in shared memory for the whole solve; per iteration a CTA computes its rows of q = A p from a replicated p, the dot
// products go through distributed shared memory (every CTA stores its partial into every CTA's slot, one cluster barrier,
// summed in rank order: the same bits everywhere), the new p is all-gathered the same way.  Three cluster barriers per
// iteration, no global memory traffic, no launch.  Same recurrences, stop test (all columns, ||r|| < max(rtol ||b||,
// atol)), true-residual refresh every `resid_calc_every` iterations and best-iterate bookkeeping as the general path.
#include "common.cuh"

#include <cooperative_groups.h>
#include <cstring>

namespace cg = cooperative_groups;

namespace xt {

constexpr int SC_CLUSTER = 8;
constexpr int SC_THREADS = 256;
constexpr int SC_MAXCOLS = 4;

struct SmallCgArgs {
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* X; int64_t ldx;
  int n, ncols, max_niter, resid_every;
  double rtol, atol, eps;
  int* out;              // device: [0] niter, [1] converged; out + 2 (as double*, 8-byte aligned): best residual
};

template <typename T> __device__ __forceinline__ T block_sum_small(T v, T* scratch) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T s = T(0);
  for (int w = 0; w < SC_THREADS / 32; ++w) s += scratch[w];
  return s;
}

template <typename TV>
__global__ void __cluster_dims__(SC_CLUSTER, 1, 1) __launch_bounds__(SC_THREADS, 1)
cg_cluster_kernel(const SmallCgArgs g) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = g.n, nc = g.ncols;
  const int RL = (n + SC_CLUSTER - 1) / SC_CLUSTER;          // rows per CTA
  const int row0 = rank * RL;
  const int rows = max(0, min(RL, n - row0));
  extern __shared__ __align__(16) unsigned char sc_raw[];
  TV* As = reinterpret_cast<TV*>(sc_raw);                      // [RL][n]
  TV* pf = As + (size_t)RL * n;                                // [nc][n]   replicated search directions
  TV* xf = pf + (size_t)SC_MAXCOLS * n;                        // [nc][n]   replicated iterate (true-residual refresh only)
  TV* xl = xf + (size_t)SC_MAXCOLS * n;                        // [nc][RL]  local slices
  TV* rl = xl + SC_MAXCOLS * RL;
  TV* pl = rl + SC_MAXCOLS * RL;
  TV* ql = pl + SC_MAXCOLS * RL;
  TV* bl = ql + SC_MAXCOLS * RL;
  TV* xb = bl + SC_MAXCOLS * RL;                               // best iterate so far
  double* red = reinterpret_cast<double*>(xb + SC_MAXCOLS * RL);   // [2 slots][SC_CLUSTER][SC_MAXCOLS]
  double* scratch = red + 2 * SC_CLUSTER * SC_MAXCOLS;         // [8]
  int slot = 0;

  // all-reduce of nc per-column partial sums over the cluster (deterministic: summed in rank order on every CTA)
  auto allreduce = [&](const double (&part)[SC_MAXCOLS], double (&tot)[SC_MAXCOLS]) {
    double mine[SC_MAXCOLS];
#pragma unroll
    for (int c = 0; c < SC_MAXCOLS; ++c) mine[c] = block_sum_small<double>(c < nc ? part[c] : 0.0, scratch);
    if (tid < SC_CLUSTER) {
      double* dst = cluster.map_shared_rank(red, tid) + ((size_t)slot * SC_CLUSTER + rank) * SC_MAXCOLS;
#pragma unroll
      for (int c = 0; c < SC_MAXCOLS; ++c) dst[c] = mine[c];
    }
    cluster.sync();
#pragma unroll
    for (int c = 0; c < SC_MAXCOLS; ++c) {
      double s = 0.0;
      for (int q = 0; q < SC_CLUSTER; ++q) s += red[((size_t)slot * SC_CLUSTER + q) * SC_MAXCOLS + c];
      tot[c] = s;
    }
    slot ^= 1;
  };
  // this CTA's rows of `src` (local slices [nc][RL]) into every CTA's replicated copy `dstfull` ([nc][n]); the caller
  // follows with a cluster barrier
  auto allgather = [&](const TV* src, TV* dstfull) {
    for (int q = 0; q < SC_CLUSTER; ++q) {
      TV* dst = cluster.map_shared_rank(dstfull, q);
      for (int e = tid; e < nc * rows; e += SC_THREADS) {
        const int c = e / rows, r = e - c * rows;
        dst[(size_t)c * n + row0 + r] = src[c * RL + r];
      }
    }
  };
  // out_local[c][r] = sum_j As[r][j] * vfull[c][j]   (warp per row, lanes over the columns of A)
  auto matvec = [&](const TV* vfull, TV* out_local) {
    for (int r = warp; r < rows; r += SC_THREADS / 32) {
      double acc[SC_MAXCOLS];
#pragma unroll
      for (int c = 0; c < SC_MAXCOLS; ++c) acc[c] = 0.0;
      const TV* ar = As + (size_t)r * n;
      for (int j = lane; j < n; j += 32) {
        const double a = (double)ar[j];
#pragma unroll
        for (int c = 0; c < SC_MAXCOLS; ++c)
          if (c < nc) acc[c] = fma(a, (double)vfull[(size_t)c * n + j], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < SC_MAXCOLS; ++c) {
        if (c < nc) {
          const double s = warp_sum(acc[c]);
          if (lane == 0) out_local[c * RL + r] = (TV)s;
        }
      }
    }
  };

  // ---- load: A rows, b, x = 0, r = p = b
  const TV* A = static_cast<const TV*>(g.A);
  const TV* B = static_cast<const TV*>(g.B);
  for (int e = tid; e < rows * n; e += SC_THREADS) {
    const int r = e / n, j = e - r * n;
    As[e] = A[(int64_t)(row0 + r) * g.lda + j];
  }
  for (int e = tid; e < nc * RL; e += SC_THREADS) {
    const int c = e / RL, r = e - c * RL;
    const TV b = (r < rows) ? B[(int64_t)(row0 + r) * g.ldb + c] : TV(0);
    bl[e] = b; rl[e] = b; pl[e] = b; xl[e] = TV(0); xb[e] = TV(0); ql[e] = TV(0);
  }
  __syncthreads();
  cluster.sync();                                             // every CTA's shared memory is set up before remote stores
  allgather(pl, pf);
  double part[SC_MAXCOLS], rr[SC_MAXCOLS], tot[SC_MAXCOLS], stop[SC_MAXCOLS];
#pragma unroll
  for (int c = 0; c < SC_MAXCOLS; ++c) {
    double s = 0.0;
    if (c < nc)
      for (int r = tid; r < rows; r += SC_THREADS) s = fma((double)rl[c * RL + r], (double)rl[c * RL + r], s);
    part[c] = s;
  }
  allreduce(part, rr);                                        // (its barrier also closes the all-gather of p)
  double best = 0.0;
#pragma unroll
  for (int c = 0; c < SC_MAXCOLS; ++c) {
    stop[c] = c < nc ? fmax(g.rtol * sqrt(rr[c]), g.atol) : 0.0;
    if (c < nc) best = fmax(best, sqrt(rr[c]));
  }
  int niter = 0, converged = 0;
  for (int k = 1; k <= g.max_niter; ++k) {
    niter = k;
    matvec(pf, ql);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < SC_MAXCOLS; ++c) {
      double s = 0.0;
      if (c < nc)
        for (int r = tid; r < rows; r += SC_THREADS) s = fma((double)pl[c * RL + r], (double)ql[c * RL + r], s);
      part[c] = s;
    }
    allreduce(part, tot);
    double alpha[SC_MAXCOLS];
#pragma unroll
    for (int c = 0; c < SC_MAXCOLS; ++c) alpha[c] = c < nc ? rr[c] / (tot[c] == 0.0 ? g.eps : tot[c]) : 0.0;
    const bool refresh = g.resid_every != 0 && (k % g.resid_every == 0);
    for (int e = tid; e < nc * RL; e += SC_THREADS) {
      const int c = e / RL;
      xl[e] = (TV)((double)xl[e] + alpha[c] * (double)pl[e]);
      if (!refresh) rl[e] = (TV)((double)rl[e] - alpha[c] * (double)ql[e]);
    }
    __syncthreads();
    if (refresh) {                                            // r = b - A x from the iterate itself (solve.py:148-149)
      allgather(xl, xf);
      cluster.sync();
      matvec(xf, ql);
      __syncthreads();
      for (int e = tid; e < nc * RL; e += SC_THREADS) rl[e] = (TV)((double)bl[e] - (double)ql[e]);
      __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < SC_MAXCOLS; ++c) {
      double s = 0.0;
      if (c < nc)
        for (int r = tid; r < rows; r += SC_THREADS) s = fma((double)rl[c * RL + r], (double)rl[c * RL + r], s);
      part[c] = s;
    }
    allreduce(part, tot);
    double mx = 0.0;
    bool all_below = true;
#pragma unroll
    for (int c = 0; c < SC_MAXCOLS; ++c) {
      if (c < nc) {
        const double rn = sqrt(tot[c]);
        mx = fmax(mx, rn);
        all_below = all_below && (rn < stop[c]);
      }
    }
    if (mx < best) {                                          // best iterate by the largest column residual (solve.py:157-160)
      best = mx;
      for (int e = tid; e < nc * RL; e += SC_THREADS) xb[e] = xl[e];
    }
    if (all_below) { converged = 1; break; }
    for (int e = tid; e < nc * RL; e += SC_THREADS) {
      const int c = e / RL;
      const double beta = tot[c] / (rr[c] == 0.0 ? g.eps : rr[c]);
      pl[e] = (TV)((double)rl[e] + beta * (double)pl[e]);
    }
#pragma unroll
    for (int c = 0; c < SC_MAXCOLS; ++c) rr[c] = tot[c];
    __syncthreads();
    allgather(pl, pf);
    cluster.sync();
  }
  __syncthreads();
  TV* X = static_cast<TV*>(g.X);
  for (int e = tid; e < nc * rows; e += SC_THREADS) {
    const int c = e / rows, r = e - c * rows;
    X[(int64_t)(row0 + r) * g.ldx + c] = xb[c * RL + r];
  }
  if (rank == 0 && tid == 0) {
    g.out[0] = niter;
    g.out[1] = converged;
    *reinterpret_cast<double*>(g.out + 2) = best;
  }
  cluster.sync();                                             // nobody leaves while a peer may still store into its memory
}

template <typename TV> static size_t small_cg_smem(int n) {
  const int RL = (n + SC_CLUSTER - 1) / SC_CLUSTER;
  return ((size_t)RL * n + (size_t)2 * SC_MAXCOLS * n + (size_t)6 * SC_MAXCOLS * RL) * sizeof(TV) +
         (size_t)(2 * SC_CLUSTER * SC_MAXCOLS + 8) * sizeof(double) + 64;
}

// true when the whole problem fits the shared memory of one cluster and nothing but the plain recurrences is asked for
bool small_cg_applies(const xt_solve_args* g) {
  if (g->nbatch != 1 || g->ncols < 1 || g->ncols > SC_MAXCOLS || g->n < SC_CLUSTER) return false;
  if (g->E != nullptr || g->M != nullptr || g->apply != nullptr || g->precond_l != nullptr || g->precond_r != nullptr)
    return false;
  if (g->dtype != XT_F32 && g->dtype != XT_F64) return false;
  if (getenv("XT_NO_SMALL_CG") != nullptr) return false;
  const size_t smem = g->dtype == XT_F64 ? small_cg_smem<double>(g->n) : small_cg_smem<float>(g->n);
  return smem <= (size_t)200 * 1024;
}

template <typename TV> static int run_small_cg_t(const xt_solve_args* g) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(g->stream);
  XT_REQUIRE(g->workspace != nullptr && g->workspace_bytes >= 64, "cg (on-chip): workspace too small");
  int* out = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(g->workspace) + 15) & ~uintptr_t(15));
  SmallCgArgs a;
  memset(&a, 0, sizeof(a));
  a.A = g->A; a.lda = g->lda; a.B = g->B; a.ldb = g->ldb; a.X = g->X; a.ldx = g->ldx;
  a.n = g->n; a.ncols = g->ncols; a.max_niter = g->max_niter; a.resid_every = g->resid_calc_every;
  a.rtol = g->rtol; a.atol = g->atol; a.eps = g->eps;
  a.out = out;
  const size_t smem = small_cg_smem<TV>(g->n);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    XT_CUDA_OK(cudaFuncSetAttribute(cg_cluster_kernel<TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.mark();
  }
  cg_cluster_kernel<TV><<<SC_CLUSTER, SC_THREADS, smem, st>>>(a);
  XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  struct { int niter, converged; double best; } h;
  XT_CUDA_OK(cudaMemcpyAsync(&h, out, sizeof(h), cudaMemcpyDeviceToHost, st));
  XT_CUDA_OK(cudaStreamSynchronize(st));
  if (g->niter_out) *g->niter_out = h.niter;
  if (g->converged_out) *g->converged_out = h.converged;
  if (g->best_resid_out) *g->best_resid_out = h.best;
  if (g->napply_out) *g->napply_out = h.niter + (g->resid_calc_every ? h.niter / g->resid_calc_every : 0);
  return XT_OK;
}

int run_small_cg(const xt_solve_args* g) {
  return g->dtype == XT_F64 ? run_small_cg_t<double>(g) : run_small_cg_t<float>(g);
}

}  // namespace xt
