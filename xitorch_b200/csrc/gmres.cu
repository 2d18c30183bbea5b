// Unrestarted GMRES for dense operators -- B200-native restatement of
//   gmres  xitorch/_impls/linalg/solve.py:326-433
// Per iteration: one fused block-matvec pass over A (matvec.cu) and ONE vector kernel that does the
// Arnoldi orthogonalisation (classical Gram-Schmidt applied twice instead of the reference's modified
// Gram-Schmidt loop of k dependent dot/axpy launches, solve.py:391-393), the Givens update of the
// Hessenberg least-squares problem (instead of a dense lstsq per iteration, solve.py:403) and the
// implicit residual norm |g_{k+1}| (instead of an extra explicit-residual matvec per iteration, :414).
// Each (batch item, column) owns its Krylov space, as in the reference (q[k] has shape (*batch, n, ncols)).
// GMRES residual norms are non-increasing, so the last iterate is the reference's "best" iterate.
#include "solve_common.cuh"

namespace xt {

constexpr int GM_JCHUNK = 4;

template <typename TV> struct GmState {
  int n, nbatch, ncols, maxk;
  TV* Q;            // [maxk+1][nbatch][n][ncols]
  TV* w;            // [nbatch][n][ncols]
  int64_t qstride;  // nbatch*n*ncols
  const TV* B; int64_t ldb, b_bstride;
  double* R;        // [nbatch*ncols][maxk][maxk+1]  column j of the triangular factor at R[bc][j][0..j]
  double* cs;       // [nbatch*ncols][maxk][2]
  double* gvec;     // [nbatch*ncols][maxk+1]
  double* hcol;     // [nbatch*ncols][maxk+2]   scratch Hessenberg column
  double* stop;     // [nbatch*ncols]
  double* cta_max; int* cta_bad;
  SolveCtl* ctl;
  double eps;
  // sliced Arnoldi step (see gm_step_sliced_kernel)
  int nslices;
  double* slice_part;        // [3][nbatch * nslices][(maxk + 1) * ncols]
  unsigned int* slice_bar;   // [nbatch]
};

template <typename TV>
__device__ __forceinline__ void gm_bookkeeping(const GmState<TV>& S, int b, const double* resn, int iter) {
  if (threadIdx.x == 0) {
    double mx = 0.0;
    int bad = 0;
    for (int c = 0; c < S.ncols; ++c) {
      const double nr = resn[c];
      mx = nr > mx ? nr : mx;
      if (!(nr < S.stop[b * S.ncols + c])) ++bad;
      if (!(nr == nr)) mx = INFINITY;
    }
    S.cta_max[b] = mx;
    S.cta_bad[b] = bad;
    __threadfence();
    const unsigned int ticket = atomicAdd(&S.ctl->counter, 1u);
    if (ticket == (unsigned int)(S.nbatch - 1)) {
      __threadfence();
      double gmax = 0.0;
      int gbad = 0;
      for (int i = 0; i < S.nbatch; ++i) {
        const double m = __ldcg(&S.cta_max[i]);
        gmax = m > gmax ? m : gmax;
        gbad += __ldcg(&S.cta_bad[i]);
      }
      SolveCtl* ctl = S.ctl;
      ctl->last_iter = iter;
      ctl->niter = iter;
      if (gmax < ctl->best_resid) ctl->best_resid = gmax;
      if (gbad == 0) { ctl->converged = 1; ctl->done = 1; }
      ctl->counter = 0;
      __threadfence();
    }
  }
}

// Q[0] = B / ||B||, g = ||B|| e1, stop = max(rtol ||b||, atol)
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) gm_init_kernel(GmState<TV> S, double rtol, double atol) {
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  const Geo g = geo(S.ncols);
  double* res = sm;
  double* scr = sm + S.ncols;
  const int64_t len = (int64_t)S.n * S.ncols;
  const TV* Bb = S.B + (int64_t)b * S.b_bstride;
  double part[1][SV_MAXCS];
#pragma unroll
  for (int i = 0; i < SV_MAXCS; ++i) part[0][i] = 0.0;
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols)
      for (int row = g.ty; row < S.n; row += g.TY) {
        const double v = (double)Bb[(int64_t)row * S.ldb + c];
        part[0][cs] += v * v;
      }
  }
  col_reduce<1>(part, S.ncols, g.tx, g.ty, g.TX, g.TY, scr, res);
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols) {
      const double bn = sqrt(res[c]);
      const TV inv = (TV)(1.0 / safedenom(bn, S.eps));
      for (int row = g.ty; row < S.n; row += g.TY)
        S.Q[(int64_t)b * len + (int64_t)row * S.ncols + c] = Bb[(int64_t)row * S.ldb + c] * inv;
    }
  }
  if (threadIdx.x == 0) {
    double mx = 0.0;
    for (int c = 0; c < S.ncols; ++c) {
      const double bn = sqrt(res[c]);
      S.stop[b * S.ncols + c] = rtol * bn > atol ? rtol * bn : atol;
      S.gvec[(int64_t)(b * S.ncols + c) * (S.maxk + 1)] = bn;
      mx = bn > mx ? bn : mx;
    }
    atomicMax(reinterpret_cast<unsigned long long*>(&S.ctl->best_resid), (unsigned long long)__double_as_longlong(mx));
  }
}

// Arnoldi step k: w = A q_k has been computed by the driver.  Orthogonalise w against q_0..q_k (CGS2),
// normalise into q_{k+1}, update the Givens QR of the Hessenberg matrix and the residual estimate.
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) gm_step_kernel(GmState<TV> S, int k) {
  extern __shared__ double sm[];
  if (S.ctl->done) return;
  const int b = blockIdx.x;
  const Geo g = geo(S.ncols);
  const int nc = S.ncols;
  double* res = sm;                              // [GM_JCHUNK][nc]
  double* hs = sm + GM_JCHUNK * nc;              // [(k+2)][nc]  accumulated h
  double* resn = hs + (size_t)(S.maxk + 2) * nc; // [nc]
  double* scr = resn + nc;                       // [TY][GM_JCHUNK][TX]
  const int64_t len = (int64_t)S.n * nc;
  TV* wb = S.w + (int64_t)b * len;

  for (int i = threadIdx.x; i < (k + 2) * nc; i += blockDim.x) hs[i] = 0.0;
  __syncthreads();

  for (int pass = 0; pass < 2; ++pass) {
    // h_j = q_j . w   for j = 0..k  (chunks of GM_JCHUNK basis vectors)
    for (int j0 = 0; j0 <= k; j0 += GM_JCHUNK) {
      double part[GM_JCHUNK][SV_MAXCS];
#pragma unroll
      for (int jj = 0; jj < GM_JCHUNK; ++jj)
#pragma unroll
        for (int i = 0; i < SV_MAXCS; ++i) part[jj][i] = 0.0;
      for (int cs = 0; cs * g.TX < nc; ++cs) {
        const int c = cs * g.TX + g.tx;
        if (c < nc)
          for (int row = g.ty; row < S.n; row += g.TY) {
            const int64_t o = (int64_t)row * nc + c;
            const double wv = (double)wb[o];
#pragma unroll
            for (int jj = 0; jj < GM_JCHUNK; ++jj)
              if (j0 + jj <= k) part[jj][cs] += (double)S.Q[(int64_t)(j0 + jj) * S.qstride + (int64_t)b * len + o] * wv;
          }
      }
      col_reduce<GM_JCHUNK>(part, nc, g.tx, g.ty, g.TX, g.TY, scr, res);
      // w -= sum_j h_j q_j for this chunk;  accumulate h
      for (int cs = 0; cs * g.TX < nc; ++cs) {
        const int c = cs * g.TX + g.tx;
        if (c < nc)
          for (int row = g.ty; row < S.n; row += g.TY) {
            const int64_t o = (int64_t)row * nc + c;
            double acc = 0.0;
#pragma unroll
            for (int jj = 0; jj < GM_JCHUNK; ++jj)
              if (j0 + jj <= k)
                acc += res[jj * nc + c] * (double)S.Q[(int64_t)(j0 + jj) * S.qstride + (int64_t)b * len + o];
            wb[o] = (TV)((double)wb[o] - acc);
          }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < GM_JCHUNK * nc; i += blockDim.x) {
        const int jj = i / nc, c = i - jj * nc;
        if (j0 + jj <= k) hs[(j0 + jj) * nc + c] += res[i];
      }
      __syncthreads();
    }
  }
  // note: subtracting chunk by chunk within a pass is block-modified Gram-Schmidt across chunks and classical
  // within a chunk; two passes give orthogonality to rounding either way.
  {
    double part[1][SV_MAXCS];
#pragma unroll
    for (int i = 0; i < SV_MAXCS; ++i) part[0][i] = 0.0;
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc)
        for (int row = g.ty; row < S.n; row += g.TY) {
          const double wv = (double)wb[(int64_t)row * nc + c];
          part[0][cs] += wv * wv;
        }
    }
    col_reduce<1>(part, nc, g.tx, g.ty, g.TX, g.TY, scr, res);
  }
  // q_{k+1} = w / ||w||
  if (k + 1 <= S.maxk) {
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc) {
        const double hn = sqrt(res[c]);
        const TV inv = (TV)(hn > 0.0 ? 1.0 / hn : 0.0);
        for (int row = g.ty; row < S.n; row += g.TY) {
          const int64_t o = (int64_t)row * nc + c;
          S.Q[(int64_t)(k + 1) * S.qstride + (int64_t)b * len + o] = wb[o] * inv;
        }
      }
    }
  }
  // Givens update per column (tiny, serial in k)
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    const int64_t bc = (int64_t)b * nc + c;
    double* Rcol = S.R + (bc * S.maxk + k) * (S.maxk + 1);
    double* csv = S.cs + bc * S.maxk * 2;
    double* gv = S.gvec + bc * (S.maxk + 1);
    double hk1 = sqrt(res[c]);
    // apply the previous rotations to the new column
    double prev = hs[0 * nc + c];
    for (int j = 0; j < k; ++j) {
      const double cj = csv[2 * j], sj = csv[2 * j + 1];
      const double nxt = hs[(j + 1) * nc + c];
      Rcol[j] = cj * prev + sj * nxt;
      prev = -sj * prev + cj * nxt;
    }
    const double denom = sqrt(prev * prev + hk1 * hk1);
    double ck = 1.0, sk = 0.0;
    if (denom > 0.0) { ck = prev / denom; sk = hk1 / denom; }
    csv[2 * k] = ck;
    csv[2 * k + 1] = sk;
    Rcol[k] = ck * prev + sk * hk1;
    const double gk = gv[k];
    gv[k] = ck * gk;
    gv[k + 1] = -sk * gk;
    resn[c] = fabs(gv[k + 1]);
  }
  __syncthreads();
  gm_bookkeeping(S, b, resn, k + 1);
}

// ---------------------------------------------------------------------------- sliced Arnoldi step
// One CTA per system pushes 2 x 2 x (k + 1) basis vectors through one SM per step (330 us at n = 16384, k = 128: twice
// the matvec).  Here the rows of a system are split over `nslices` co-resident CTAs (cooperative launch).  Each keeps its
// rows of w in shared memory (fp64) and does classical Gram-Schmidt twice over ALL basis vectors at once, so a step has
// three cross-slice reductions -- the (k + 1) x ncols projections of each pass and the norm -- instead of one per chunk
// of basis vectors.  Reductions go through per-slice tables (one per reduction of the step) summed in slice order by
// one warp per entry: the same bits in every slice.  Slice 0 does the Givens update and the bookkeeping.
template <typename TV>
__device__ __forceinline__ void gm_slice_allreduce(const GmState<TV>& S, int b, int sl, double* vals, int count, int table,
                                                   unsigned int arrival) {
  const size_t ps = (size_t)(S.maxk + 1) * S.ncols;
  double* tab = S.slice_part + (size_t)table * S.nbatch * S.nslices * ps;
  double* mine = tab + ((size_t)b * S.nslices + sl) * ps;
  for (int i = threadIdx.x; i < count; i += blockDim.x) mine[i] = vals[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&S.slice_bar[b], 1u);
    const unsigned int target = (arrival + 1u) * (unsigned int)S.nslices;
    while (*reinterpret_cast<volatile unsigned int*>(&S.slice_bar[b]) < target) {
      XT_SPIN_PAUSE();
    }
    __threadfence();
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < count; i += nw) {
    double sum = 0.0;
    for (int j = lane; j < S.nslices; j += 32) sum += __ldcg(&tab[((size_t)b * S.nslices + j) * ps + i]);
    sum = warp_sum(sum);
    if (lane == 0) vals[i] = sum;
  }
  __syncthreads();
}

template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) gm_step_sliced_kernel(GmState<TV> S, int k) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double sm[];
  if (S.ctl->done) return;
  const int b = blockIdx.x / S.nslices, sl = blockIdx.x - b * S.nslices;
  const int per = (S.n + S.nslices - 1) / S.nslices;
  const int lo = sl * per < S.n ? sl * per : S.n;
  const int hi = lo + per < S.n ? lo + per : S.n;
  const int R = hi - lo, nc = S.ncols, np = (k + 1) * nc, ne = R * nc;
  const size_t ps = (size_t)(S.maxk + 1) * nc;
  double* hs = sm;                 // [(maxk+1) nc]  h accumulated over the two passes
  double* hp = hs + ps;            // [(maxk+1) nc]  h of this pass
  double* nrm = hp + ps;           // [nc]
  double* resn = nrm + nc;         // [nc]
  double* grp = resn + nc;         // [SV_THREADS]   partial sums of the subtraction, one row per group of basis vectors
  double* wsm = grp + SV_THREADS;  // [per nc]       this slice of w
  const int64_t len = (int64_t)S.n * nc;
  const int64_t off = (int64_t)b * len + (int64_t)lo * nc;        // this slice inside a (nbatch, n, ncols) block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;

  for (int e = threadIdx.x; e < ne; e += blockDim.x) wsm[e] = (double)S.w[off + e];
  for (int i = threadIdx.x; i < np; i += blockDim.x) hs[i] = 0.0;
  __syncthreads();

  // groups of basis vectors for the subtraction: EP threads cover the elements of the slice, G groups split j
  int EP = 32;
  while (EP < ne && EP < SV_THREADS) EP <<= 1;
  const int G = SV_THREADS / EP;
  const int eg = threadIdx.x % EP, jg = threadIdx.x / EP;

  for (int pass = 0; pass < 2; ++pass) {
    // h_j = q_j . w over this slice: one warp per (j, column), lanes over the rows
    for (int p = warp; p < np; p += nw) {
      const int j = p / nc, c = p - j * nc;
      const TV* q = S.Q + (int64_t)j * S.qstride + off;
      double s = 0.0;
      for (int row = lane; row < R; row += 32) s += (double)q[row * nc + c] * wsm[row * nc + c];
      s = warp_sum(s);
      if (lane == 0) hp[p] = s;
    }
    __syncthreads();
    gm_slice_allreduce(S, b, sl, hp, np, pass, 3u * (unsigned int)k + (unsigned int)pass);
    for (int i = threadIdx.x; i < np; i += blockDim.x) hs[i] += hp[i];
    // w -= sum_j h_j q_j
    for (int e0 = 0; e0 < ne; e0 += EP) {
      const int e = e0 + eg;
      double acc = 0.0;
      if (e < ne) {
        const int c = e % nc;
        for (int j = jg; j <= k; j += G) acc += hp[j * nc + c] * (double)S.Q[(int64_t)j * S.qstride + off + e];
      }
      grp[jg * EP + eg] = acc;
      __syncthreads();
      if (jg == 0 && e < ne) {
        double tot = 0.0;
        for (int gi = 0; gi < G; ++gi) tot += grp[gi * EP + eg];
        wsm[e] -= tot;
      }
      __syncthreads();
    }
  }
  // ||w||^2 per column
  for (int c = warp; c < nc; c += nw) {
    double s = 0.0;
    for (int row = lane; row < R; row += 32) s += wsm[row * nc + c] * wsm[row * nc + c];
    s = warp_sum(s);
    if (lane == 0) nrm[c] = s;
  }
  __syncthreads();
  gm_slice_allreduce(S, b, sl, nrm, nc, 2, 3u * (unsigned int)k + 2u);
  // q_{k+1} = w / ||w||
  if (k + 1 <= S.maxk) {
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
      const double hn = sqrt(nrm[e % nc]);
      S.Q[(int64_t)(k + 1) * S.qstride + off + e] = (TV)(hn > 0.0 ? wsm[e] / hn : 0.0);
    }
  }
  if (sl != 0) return;
  // Givens update per column (tiny, serial in k)
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    const int64_t bc = (int64_t)b * nc + c;
    double* Rcol = S.R + (bc * S.maxk + k) * (S.maxk + 1);
    double* csv = S.cs + bc * S.maxk * 2;
    double* gv = S.gvec + bc * (S.maxk + 1);
    const double hk1 = sqrt(nrm[c]);
    double prev = hs[0 * nc + c];
    for (int j = 0; j < k; ++j) {
      const double cj = csv[2 * j], sj = csv[2 * j + 1];
      const double nxt = hs[(j + 1) * nc + c];
      Rcol[j] = cj * prev + sj * nxt;
      prev = -sj * prev + cj * nxt;
    }
    const double denom = sqrt(prev * prev + hk1 * hk1);
    double ck = 1.0, sk = 0.0;
    if (denom > 0.0) { ck = prev / denom; sk = hk1 / denom; }
    csv[2 * k] = ck;
    csv[2 * k + 1] = sk;
    Rcol[k] = ck * prev + sk * hk1;
    const double gk = gv[k];
    gv[k] = ck * gk;
    gv[k + 1] = -sk * gk;
    resn[c] = fabs(gv[k + 1]);
  }
  __syncthreads();
  gm_bookkeeping(S, b, resn, k + 1);
}

template <typename TV> static size_t gm_sliced_smem(const GmState<TV>& S) {
  const int per = (S.n + S.nslices - 1) / S.nslices;
  return ((size_t)2 * (S.maxk + 1) * S.ncols + 2 * S.ncols + SV_THREADS + (size_t)per * S.ncols + 64) * sizeof(double);
}

// x = sum_j y_j q_j with R y = g (back substitution over the first `kk` Arnoldi vectors, kk = ctl->niter)
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) gm_final_kernel(GmState<TV> S, TV* X, int64_t ldx, int64_t x_bstride) {
  extern __shared__ double sm[];
  const int b = blockIdx.x / S.nslices, sl = blockIdx.x - b * S.nslices;   // every slice solves the small triangular system
  const int nc = S.ncols;
  const int kk = S.ctl->niter;
  double* y = sm;   // [kk][nc]
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    const int64_t bc = (int64_t)b * nc + c;
    const double* gv = S.gvec + bc * (S.maxk + 1);
    for (int i = kk - 1; i >= 0; --i) {
      double s = gv[i];
      for (int j = i + 1; j < kk; ++j) s -= S.R[(bc * S.maxk + j) * (S.maxk + 1) + i] * y[j * nc + c];
      const double d = S.R[(bc * S.maxk + i) * (S.maxk + 1) + i];
      y[i * nc + c] = s / (d == 0.0 ? S.eps : d);
    }
  }
  __syncthreads();
  const int64_t len = (int64_t)S.n * nc;
  TV* Xb = X + (int64_t)b * x_bstride;
  const int per = (S.n + S.nslices - 1) / S.nslices;
  const int64_t e_lo = (int64_t)(sl * per < S.n ? sl * per : S.n) * nc;
  const int64_t e_hi = (int64_t)((sl + 1) * per < S.n ? (sl + 1) * per : S.n) * nc;
  for (int64_t e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
    const int64_t row = e / nc;
    const int c = (int)(e - row * nc);
    double acc = 0.0;
    for (int j = 0; j < kk; ++j) acc += y[j * nc + c] * (double)S.Q[(int64_t)j * S.qstride + (int64_t)b * len + e];
    Xb[row * ldx + c] = (TV)acc;
  }
}

template <typename TV> static int run_gmres(const xt_solve_args* g) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(g->stream);
  Arena ar(g->workspace, g->workspace_bytes);
  const int64_t len = (int64_t)g->nbatch * g->n * g->ncols;
  int maxk = g->max_niter;
  if (maxk > g->n) maxk = g->n;
  XT_REQUIRE(maxk >= 1, "gmres: max_niter must be >= 1");
  GmState<TV> S;
  memset(&S, 0, sizeof(S));
  S.n = g->n; S.nbatch = g->nbatch; S.ncols = g->ncols; S.maxk = maxk;
  S.qstride = len;
  S.Q = ar.take<TV>((size_t)(maxk + 1) * len);
  S.w = ar.take<TV>(len);
  TV* mx = ar.take<TV>(len);
  const int64_t nbc = (int64_t)g->nbatch * g->ncols;
  S.R = ar.take<double>((size_t)nbc * maxk * (maxk + 1));
  S.cs = ar.take<double>((size_t)nbc * maxk * 2);
  S.gvec = ar.take<double>((size_t)nbc * (maxk + 1));
  S.hcol = nullptr;
  S.stop = ar.take<double>(nbc);
  S.cta_max = ar.take<double>(g->nbatch);
  S.cta_bad = ar.take<int>(g->nbatch);
  S.ctl = ar.take<SolveCtl>(1);
  S.nslices = step_slices(g->n, g->nbatch);
  if (S.nslices > 1 && gm_sliced_smem(S) > 200 * 1024) S.nslices = 1;
  if (S.nslices > 1) {
    S.slice_part = ar.take<double>((size_t)3 * g->nbatch * S.nslices * (maxk + 1) * g->ncols);
    S.slice_bar = ar.take<unsigned int>(g->nbatch);
  }
  S.B = static_cast<const TV*>(g->B); S.ldb = g->ldb; S.b_bstride = g->b_bstride;
  S.eps = g->eps;
  if (!ar.ok()) {
    set_last_error("gmres: workspace too small (%zu needed, %zu given)", ar.off, ar.cap);
    return XT_ERR_WORKSPACE;
  }
  OpDesc op{g->dtype, g->n, g->nbatch, g->ncols, g->A, g->lda, g->a_bstride, nullptr, 0, 0, nullptr, 0};
  op.apply = g->apply; op.apply_user = g->apply_user; op.abort = g->abort;
  op.pdl = solve_pdl_enabled() ? 1 : 0;
  XT_CUDA_OK(cudaMemsetAsync(S.ctl, 0, sizeof(SolveCtl), st));
  if (S.nslices > 1) XT_CUDA_OK(cudaMemsetAsync(S.slice_bar, 0, (size_t)g->nbatch * sizeof(unsigned int), st));
  const size_t smem_init = (size_t)(2 * g->ncols + 2 * SV_THREADS + 64) * sizeof(double);
  const size_t smem_step =
      (size_t)((GM_JCHUNK + maxk + 3) * g->ncols + GM_JCHUNK * SV_THREADS + 64) * sizeof(double);
  const size_t smem_fin = (size_t)(maxk + 1) * g->ncols * sizeof(double) + 64;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    XT_CUDA_OK(cudaFuncSetAttribute(gm_step_kernel<TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    XT_CUDA_OK(cudaFuncSetAttribute(gm_final_kernel<TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    XT_CUDA_OK(cudaFuncSetAttribute(gm_step_sliced_kernel<TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.mark();
  }
  XT_REQUIRE(smem_step <= 200 * 1024 && smem_fin <= 200 * 1024,
             "gmres: max_niter*ncols = %d*%d too large for the on-chip Hessenberg state", maxk, g->ncols);
  gm_init_kernel<TV><<<g->nbatch, SV_THREADS, smem_init, st>>>(S, g->rtol, g->atol); XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  int64_t napply = 0;
  const int ce = g->check_every > 0 ? g->check_every : 1;
  int next_check = ce < 4 ? ce : 4;      // poll the device flag at 4, 8, 16, ... iterations, then every `ce`
  const int* done_flag = &S.ctl->done;
  for (int k = 0; k < maxk; ++k) {
    int rc = apply_op<TV>(op, S.Q + (int64_t)k * len, S.w, mx, nullptr, nullptr, 0, done_flag, st, &napply);
    if (rc != XT_OK) return rc;
    bool stepped = false;
    if (S.nslices > 1) {
      stepped = coop_launch(gm_step_sliced_kernel<TV>, g->nbatch * S.nslices, SV_THREADS, gm_sliced_smem(S), st, S, k);
      // a refused cooperative launch: one CTA per system from here on.  The arrivals counted so far stay consistent
      // because the sliced kernel is never launched again in this solve.
      if (!stepped) S.nslices = 1;
    }
    if (!stepped) gm_step_kernel<TV><<<g->nbatch, SV_THREADS, smem_step, st>>>(S, k);
    XT_LAUNCHED();
    XT_CUDA_OK(cudaGetLastError());
    if (k + 1 == next_check || k + 1 == maxk) {
      next_check += (next_check < ce) ? next_check : ce;
      int done = 0;
      rc = poll_done(S.ctl, st, &done);
      if (rc != XT_OK) return rc;
      if (done) break;
    }
  }
  gm_final_kernel<TV><<<g->nbatch * S.nslices, SV_THREADS, smem_fin, st>>>(S, static_cast<TV*>(g->X), g->ldx, g->x_bstride); XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  SolveCtl h;
  XT_CUDA_OK(cudaMemcpyAsync(&h, S.ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
  XT_CUDA_OK(cudaStreamSynchronize(st));
  if (g->niter_out) *g->niter_out = h.niter;
  if (g->converged_out) *g->converged_out = h.converged;
  if (g->best_resid_out) *g->best_resid_out = h.best_resid;
  if (g->napply_out) *g->napply_out = napply;
  return XT_OK;
}

size_t gmres_ws_bytes(size_t vs, int n, int nbatch, int ncols, int max_niter) {
  int maxk = max_niter > n ? n : max_niter;
  if (maxk < 1) maxk = 1;
  const size_t len = (size_t)nbatch * n * ncols;
  const size_t nbc = (size_t)nbatch * ncols;
  size_t bytes = (size_t)(maxk + 3) * (len * vs + 256);
  bytes += nbc * maxk * (maxk + 1) * 8 + nbc * maxk * 16 + nbc * (maxk + 1) * 8 + nbc * 8 + (size_t)nbatch * 16;
  bytes += (size_t)3 * (nbatch + SV_MAX_SLICED_CTAS) * (maxk + 1) * ncols * 8 + (size_t)nbatch * 4 + 2 * 256;   // slice tables
  return bytes + 16 * 256 + 1024;
}

}  // namespace xt

extern "C" int xt_gmres(const xt_solve_args* g) {
  XT_REQUIRE(g != nullptr, "gmres: null args");
  XT_REQUIRE(g->n >= 1 && g->nbatch >= 1 && g->ncols >= 1, "gmres: empty problem");
  XT_REQUIRE(g->ncols <= 32 * xt::SV_MAXCS, "gmres: ncols=%d exceeds %d", g->ncols, 32 * xt::SV_MAXCS);
  XT_REQUIRE((g->A || g->apply) && g->B && g->X && g->workspace, "gmres: null pointer");
  XT_REQUIRE(g->E == nullptr && g->M == nullptr, "gmres: E / M are not supported (as in the reference method)");
  return g->dtype == XT_F64 ? xt::run_gmres<double>(g) : xt::run_gmres<float>(g);
}
