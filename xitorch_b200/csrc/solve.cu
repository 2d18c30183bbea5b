// CG / BiCGSTAB for dense operators  A X - M X diag(E) = B  -- B200-native restatement of
//   cg        xitorch/_impls/linalg/solve.py:69-190
//   bicgstab  xitorch/_impls/linalg/solve.py:192-324
// Every operator application is ONE pass over A (matvec.cu) with the following dot products fused
// into its epilogue (p.Ap | r0hat.v | t.s, t.t); the remaining O(n) vector recurrences of an
// iteration run in one small kernel per matvec.  All convergence control lives on the device:
// the reference's two host syncs per iteration (solve.py:157,166 / :300,310) become a device flag
// polled every `check_every` iterations; kernels launched after convergence exit immediately.
//
// Reference semantics kept: x0 = 0 (so r0 = B), all columns and batch items iterate in lock-step,
// stop when EVERY column has ||r|| < max(rtol ||b||, atol), the iterate with the smallest max-norm
// residual is returned (best_xk), `_safedenom` replaces exact-zero denominators by eps, the true
// residual B - A x is recomputed every `resid_calc_every` iterations.
#include "solve_common.cuh"

namespace xt {

// deferred copy of the best iterate (decided by the last CTA of the previous norm evaluation)
template <typename TV>
__device__ __forceinline__ void deferred_best_copy(const SolveState<TV>& S, int b, int prev_iter, RowRange rr) {
  if (S.ctl->improved_iter == prev_iter) {
    const int64_t len = (int64_t)S.n * S.ncols;
    const TV* src = S.x + (int64_t)b * len;
    TV* dst = S.bestx + (int64_t)b * len;
    for (int64_t i = (int64_t)rr.lo * S.ncols + threadIdx.x; i < (int64_t)rr.hi * S.ncols; i += blockDim.x) dst[i] = src[i];
  }
}

// residual-norm bookkeeping shared by cg and bicgstab: rn2[c] = ||r_c||^2 of this batch item
template <typename TV>
__device__ __forceinline__ void norm_bookkeeping(const SolveState<TV>& S, int b, const double* rn2, int iter,
                                                 double* scr) {
  // per-CTA max / count of unconverged columns
  if (threadIdx.x == 0) {
    double mx = 0.0;
    int bad = 0;
    for (int c = 0; c < S.ncols; ++c) {
      const double nr = sqrt(rn2[c]);
      mx = nr > mx ? nr : mx;
      if (!(nr < S.stop[b * S.ncols + c])) ++bad;
      if (!(nr == nr)) mx = INFINITY;   // NaN never improves / never converges
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
      if (gmax < ctl->best_resid) {
        ctl->best_resid = gmax;
        ctl->improved_iter = iter;
      }
      if (gbad == 0) {
        ctl->converged = 1;
        ctl->niter = iter;
        ctl->done = 1;
      } else {
        ctl->niter = iter;
      }
      ctl->counter = 0;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------- init
// r = p = rhat = B (x0 = 0), x = bestx = 0, rz = r.r, stop = max(rtol ||b||, atol), best_resid = max ||r||
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) solve_init_kernel(SolveState<TV> S, double rtol, double atol, int bicg) {
  extern __shared__ double sm[];
  const int b = blockIdx.x / S.nslices, sl = blockIdx.x - b * S.nslices;
  const RowRange rr = slice_rows(S, sl);
  const Geo g = geo(S.ncols);
  double* res = sm;                       // [1][ncols]
  double* scr = sm + S.ncols;             // [TY][1][TX]
  const int64_t len = (int64_t)S.n * S.ncols;
  const TV* Bb = S.B + (int64_t)b * S.b_bstride;
  double part[1][SV_MAXCS];
#pragma unroll
  for (int i = 0; i < SV_MAXCS; ++i) part[0][i] = 0.0;
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols) {
      for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
        const TV v = Bb[(int64_t)row * S.ldb + c];
        const int64_t o = (int64_t)b * len + (int64_t)row * S.ncols + c;
        S.r[o] = v;
        S.x[o] = TV(0);
        S.bestx[o] = TV(0);
        if (bicg) {
          S.rhat[o] = v;
          S.p[o] = TV(0);
          S.q[o] = TV(0);
        } else {
          S.p[o] = v;
        }
        part[0][cs] += (double)v * (double)v;
      }
    }
  }
  col_reduce<1>(part, S.ncols, g.tx, g.ty, g.TX, g.TY, scr, res);
  slice_allreduce(S, b, sl, res, S.ncols, 0u);            // the solve's reduction number 0
  if (sl == 0 && threadIdx.x == 0) {
    double mx = 0.0;
    for (int c = 0; c < S.ncols; ++c) {
      const double bn = sqrt(res[c]);
      const double st = rtol * bn > atol ? rtol * bn : atol;
      S.stop[b * S.ncols + c] = st;
      S.rz[b * S.ncols + c] = res[c];
      if (bicg) {
        S.alpha[b * S.ncols + c] = 1.0;
        S.omega[b * S.ncols + c] = 1.0;
      }
      mx = bn > mx ? bn : mx;
    }
    // best_resid = max over batch of the initial residual norm: atomicMax on the bit pattern (non-negative doubles)
    atomicMax(reinterpret_cast<unsigned long long*>(&S.ctl->best_resid), (unsigned long long)__double_as_longlong(mx));
  }
}

// ---------------------------------------------------------------------------- CG
// phase 0: alpha; x += alpha p; r -= alpha Ap; norms; beta; p = r + beta p
// phase 1: alpha; x += alpha p                                   (then the driver computes q = A x)
// phase 2: r = B - q; norms; beta; p = r + beta p
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) cg_step_kernel(SolveState<TV> S, int iter, int phase, unsigned int epoch, int rel) {
  pdl_wait();        // may have been scheduled while the matvec before it drains
  pdl_trigger();     // the matvec after it may start streaming A now (it waits for this grid before reading vectors)
  if (rel) {         // replayed from a graph: numbers relative to the period's base
    iter += S.ctl->graph_base;
    epoch += S.ctl->epoch_base;
  }
  extern __shared__ double sm[];
  const int b = blockIdx.x / S.nslices, sl = blockIdx.x - b * S.nslices;
  const RowRange rr = slice_rows(S, sl);
  const Geo g = geo(S.ncols);
  double* res = sm;                       // [ncols]   (r.r)
  double* alp = sm + S.ncols;             // [ncols]
  double* rzo = sm + 2 * S.ncols;         // [ncols]   r.z of the previous iteration (slice 0 overwrites it at the end)
  double* scr = sm + 3 * S.ncols;
  const int64_t len = (int64_t)S.n * S.ncols;
  const int64_t base = (int64_t)b * len;

  if (phase != 2) deferred_best_copy(S, b, iter - 1, rr);
  if (S.ctl->done) return;

  if (phase != 2) tile_dots_all(S, b, false, alp);        // p.Ap per column
  for (int c = threadIdx.x; c < S.ncols; c += blockDim.x) {
    rzo[c] = S.rz[b * S.ncols + c];
    if (phase != 2) alp[c] = rzo[c] / safedenom(alp[c], S.eps);
  }
  __syncthreads();
  double part[1][SV_MAXCS];
#pragma unroll
  for (int i = 0; i < SV_MAXCS; ++i) part[0][i] = 0.0;
  const TV* Bb = S.B + (int64_t)b * S.b_bstride;
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols) {
      const TV a = (phase != 2) ? (TV)alp[c] : TV(0);
      for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
        const int64_t o = base + (int64_t)row * S.ncols + c;
        if (phase != 2) S.x[o] = S.x[o] + a * S.p[o];
        if (phase == 0) {
          const TV rn = S.r[o] - a * S.q[o];
          S.r[o] = rn;
          part[0][cs] += (double)rn * (double)rn;
        } else if (phase == 2) {
          const TV rn = Bb[(int64_t)row * S.ldb + c] - S.q[o];
          S.r[o] = rn;
          part[0][cs] += (double)rn * (double)rn;
        }
      }
    }
  }
  if (phase == 1) return;
  col_reduce<1>(part, S.ncols, g.tx, g.ty, g.TX, g.TY, scr, res);
  slice_allreduce(S, b, sl, res, S.ncols, epoch);
  if (sl == 0) norm_bookkeeping(S, b, res, iter, scr);
  if (S.precond) return;          // z = P r, beta and p follow in cg_precond_kernel
  // beta = rz_new / safedenom(rz);  p = r + beta p
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols) {
      const TV beta = (TV)(res[c] / safedenom(rzo[c], S.eps));
      for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
        const int64_t o = base + (int64_t)row * S.ncols + c;
        S.p[o] = S.r[o] + beta * S.p[o];
      }
    }
  }
  if (sl == 0)
    for (int c = threadIdx.x; c < S.ncols; c += blockDim.x) S.rz[b * S.ncols + c] = res[c];
}

// preconditioned CG tail (solve.py:170-180): rz_new = r.z (fused into the application of the preconditioner),
// beta = rz_new / safedenom(rz), p = z + beta p  (first call: p = z)
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) cg_precond_kernel(SolveState<TV> S, const TV* __restrict__ z, int first) {
  pdl_wait();        // may have been scheduled while the matvec before it drains
  pdl_trigger();     // the matvec after it may start streaming A now (it waits for this grid before reading vectors)
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  const Geo g = geo(S.ncols);
  double* bet = sm;                       // [ncols]
  if (S.ctl->done) return;
  const int64_t len = (int64_t)S.n * S.ncols;
  const int64_t base = (int64_t)b * len;
  tile_dots_all(S, b, false, bet);
  for (int c = threadIdx.x; c < S.ncols; c += blockDim.x) {
    const double rzn = bet[c];
    bet[c] = first ? 0.0 : rzn / safedenom(S.rz[b * S.ncols + c], S.eps);
    S.rz[b * S.ncols + c] = rzn;
  }
  __syncthreads();
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols) {
      const TV beta = (TV)bet[c];
      for (int row = g.ty; row < S.n; row += g.TY) {
        const int64_t o = base + (int64_t)row * S.ncols + c;
        S.p[o] = first ? z[o] : z[o] + beta * S.p[o];
      }
    }
  }
}

// ---------------------------------------------------------------------------- BiCGSTAB
// stage 1 (before v = A p):   [tail of the previous iteration when iter > 1 -- see stage 3]
//                             rho_new = rhat.r; beta = rho_new/sd(rho) * alpha/sd(omega); p = r + beta (p - omega v)
// stage 2 (before t = A s):   alpha = rho_new / sd(rhat.v); x += alpha p; s = r - alpha v
// stage 3 (after t = A s):    omega = t.s / sd(t.t); x += omega s; r = s - omega t; norms; rho = rho_new
// stage 4 (true residual, before q=A x): omega...; x += omega s        (then the driver computes t = A x)
// stage 5 (true residual, after):        r = B - t; norms; rho = rho_new
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) bicg_step_kernel(SolveState<TV> S, int iter, int stage, unsigned int epoch, int rel) {
  pdl_wait();        // may have been scheduled while the matvec before it drains
  pdl_trigger();     // the matvec after it may start streaming A now (it waits for this grid before reading vectors)
  if (rel) {         // replayed from a graph: numbers relative to the period's base
    iter += S.ctl->graph_base;
    epoch += S.ctl->epoch_base;
  }
  extern __shared__ double sm[];
  const int b = blockIdx.x / S.nslices, sl = blockIdx.x - b * S.nslices;
  const RowRange rr = slice_rows(S, sl);
  const Geo g = geo(S.ncols);
  double* res = sm;                       // [2][ncols]
  double* sc = sm + 2 * S.ncols;          // [2][ncols] scalar scratch
  double* scr = sm + 4 * S.ncols;
  const int64_t len = (int64_t)S.n * S.ncols;
  const int64_t base = (int64_t)b * len;
  const int nc = S.ncols;

  if (stage == 1) deferred_best_copy(S, b, iter - 1, rr);
  if (S.ctl->done) return;

  double part[1][SV_MAXCS];
#pragma unroll
  for (int i = 0; i < SV_MAXCS; ++i) part[0][i] = 0.0;

  if (stage == 1) {
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc)
        for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          part[0][cs] += (double)S.rhat[o] * (double)S.r[o];
        }
    }
    col_reduce<1>(part, nc, g.tx, g.ty, g.TX, g.TY, scr, res);
    slice_allreduce(S, b, sl, res, nc, epoch);
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
      const int i = b * nc + c;
      // reference order (solve.py:273-275): omega and rho are "safed" in place before use (every slice computes the
      // same values; an exact zero is replaced by eps whichever slice's store lands first)
      double om = S.omega[i];
      if (om == 0.0) { om = S.eps; S.omega[i] = om; }
      double rho = S.rz[i];
      if (rho == 0.0) { rho = S.eps; S.rz[i] = rho; }
      if (sl == 0) S.rhonew[i] = res[c];
      sc[c] = res[c] / rho * (S.alpha[i] / om);
      sc[nc + c] = om;
    }
    __syncthreads();
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc) {
        const TV beta = (TV)sc[c];
        const TV om = (TV)sc[nc + c];
        for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          S.p[o] = S.r[o] + beta * (S.p[o] - om * S.q[o]);
        }
      }
    }
    return;
  }
  if (stage == 2) {
    tile_dots_all(S, b, false, sc);                     // rhat.v per column
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
      const int i = b * nc + c;
      const double a = S.rhonew[i] / safedenom(sc[c], S.eps);
      if (sl == 0) S.alpha[i] = a;
      sc[c] = a;
    }
    __syncthreads();
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc) {
        const TV a = (TV)sc[c];
        for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          S.x[o] = S.x[o] + a * S.p[o];
          S.s[o] = S.r[o] - a * S.q[o];
        }
      }
    }
    return;
  }
  if (stage == 3 || stage == 4) {
    tile_dots_all(S, b, true, sc);                      // t.s and t.t per column
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
      const int i = b * nc + c;
      const double om = sc[c] / safedenom(sc[nc + c], S.eps);
      if (sl == 0) S.omega[i] = om;
      sc[c] = om;
    }
    __syncthreads();
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc) {
        const TV om = (TV)sc[c];
        for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          S.x[o] = S.x[o] + om * S.s[o];
          if (stage == 3) {
            const TV rn = S.s[o] - om * S.t[o];
            S.r[o] = rn;
            part[0][cs] += (double)rn * (double)rn;
          }
        }
      }
    }
    if (stage == 4) return;
  } else {   // stage 5
    const TV* Bb = S.B + (int64_t)b * S.b_bstride;
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc)
        for (int row = rr.lo + g.ty; row < rr.hi; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          const TV rn = Bb[(int64_t)row * S.ldb + c] - S.t[o];
          S.r[o] = rn;
          part[0][cs] += (double)rn * (double)rn;
        }
    }
  }
  col_reduce<1>(part, nc, g.tx, g.ty, g.TX, g.TY, scr, res);
  slice_allreduce(S, b, sl, res, nc, epoch);
  if (sl != 0) return;
  norm_bookkeeping(S, b, res, iter, scr);
  __syncthreads();
  for (int c = threadIdx.x; c < nc; c += blockDim.x) S.rz[b * nc + c] = S.rhonew[b * nc + c];
}

// ---------------------------------------------------------------------------- finalize: X_out = best iterate
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) solve_final_kernel(SolveState<TV> S, TV* X, int64_t ldx,
                                                                int64_t x_bstride) {
  const int b = blockIdx.x / S.nslices, sl = blockIdx.x - b * S.nslices;
  const RowRange rr = slice_rows(S, sl);
  const int64_t len = (int64_t)S.n * S.ncols;
  const bool pending = (S.ctl->improved_iter == S.ctl->last_iter) && (S.ctl->last_iter > 0);
  const TV* src = (pending ? S.x : S.bestx) + (int64_t)b * len;
  TV* Xb = X + (int64_t)b * x_bstride;
  for (int64_t i = (int64_t)rr.lo * S.ncols + threadIdx.x; i < (int64_t)rr.hi * S.ncols; i += blockDim.x) {
    const int64_t row = i / S.ncols, c = i - row * S.ncols;
    Xb[row * ldx + c] = src[i];
  }
}

// ============================================================================ host drivers
static size_t solve_ws_bytes(int nvecs, size_t vs, int n, int nbatch, int ncols) {
  const MvTiling til = mv_tiling(nbatch, n);
  const int ngroups = (ncols + MV_MAXK - 1) / MV_MAXK;
  size_t bytes = 0;
  bytes += (size_t)nvecs * (align_up((size_t)nbatch * n * ncols * vs, 256) + 256);
  bytes += align_up((size_t)ngroups * til.ntiles * 2 * MV_MAXK * sizeof(double), 256) + 256;
  bytes += 6 * (align_up((size_t)nbatch * ncols * sizeof(double), 256) + 256);
  bytes += 2 * (align_up((size_t)nbatch * sizeof(double), 256) + 256);
  bytes += align_up((size_t)(nbatch + SV_MAX_SLICED_CTAS) * ncols * sizeof(double), 256) + 256;   // slice_part
  bytes += align_up((size_t)nbatch * sizeof(unsigned int), 256) + 256;                           // slice_bar
  bytes += 1024;
  return bytes;
}

template <typename TV>
static int setup_state(const xt_solve_args* g, int nvecs, Arena& ar, SolveState<TV>& S, TV** extra) {
  const int64_t len = (int64_t)g->nbatch * g->n * g->ncols;
  const MvTiling til = mv_tiling(g->nbatch, g->n);
  const int ngroups = (g->ncols + MV_MAXK - 1) / MV_MAXK;
  memset(&S, 0, sizeof(S));
  S.n = g->n; S.nbatch = g->nbatch; S.ncols = g->ncols;
  S.tiles_per_batch = til.tiles_per_batch;
  TV* vecs[16];
  for (int i = 0; i < nvecs; ++i) vecs[i] = ar.take<TV>(len);
  S.x = vecs[0]; S.r = vecs[1]; S.p = vecs[2]; S.q = vecs[3]; S.bestx = vecs[4];
  *extra = vecs[5];                       // M x scratch (or unused)
  if (nvecs > 6) S.s = vecs[6];
  if (nvecs > 8) { S.t = vecs[7]; S.rhat = vecs[8]; }
  if (nvecs > 11) { S.ex[0] = vecs[9]; S.ex[1] = vecs[10]; S.ex[2] = vecs[11]; }
  S.dots_gstride = (int64_t)til.ntiles * 2 * MV_MAXK;
  S.dots = ar.take<double>((size_t)ngroups * S.dots_gstride);
  S.rz = ar.take<double>((size_t)g->nbatch * g->ncols);
  S.alpha = ar.take<double>((size_t)g->nbatch * g->ncols);
  S.omega = ar.take<double>((size_t)g->nbatch * g->ncols);
  S.rhonew = ar.take<double>((size_t)g->nbatch * g->ncols);
  S.stop = ar.take<double>((size_t)g->nbatch * g->ncols);
  S.cta_max = ar.take<double>(g->nbatch);
  S.cta_bad = ar.take<int>(g->nbatch);
  S.ctl = ar.take<SolveCtl>(1);
  S.nslices = step_slices(g->n, g->nbatch);
  S.slice_part = ar.take<double>((size_t)g->nbatch * S.nslices * g->ncols);
  S.slice_bar = ar.take<unsigned int>(g->nbatch);
  S.B = static_cast<const TV*>(g->B); S.ldb = g->ldb; S.b_bstride = g->b_bstride;
  S.eps = g->eps;
  if (!ar.ok()) {
    set_last_error("solve: workspace too small (%zu needed, %zu given)", ar.off, ar.cap);
    return XT_ERR_WORKSPACE;
  }
  return XT_OK;
}

static int check_solve_args(const xt_solve_args* g) {
  XT_REQUIRE(g != nullptr, "solve: null args");
  XT_REQUIRE(g->n >= 1 && g->nbatch >= 1 && g->ncols >= 1, "solve: empty problem");
  XT_REQUIRE(g->ncols <= 32 * SV_MAXCS, "solve: ncols=%d exceeds %d", g->ncols, 32 * SV_MAXCS);
  XT_REQUIRE((g->A || g->apply) && g->B && g->X && g->workspace, "solve: null pointer");
  XT_REQUIRE(g->max_niter >= 0, "solve: negative max_niter");
  XT_REQUIRE(g->M == nullptr || g->E != nullptr, "solve: M without E");
  return XT_OK;
}

template <typename TV> static size_t step_smem(int ncols) {
  int TX = 1;
  while (TX < ncols && TX < 32) TX <<= 1;
  return (size_t)(4 * ncols + 2 * SV_THREADS + 64) * sizeof(double);
}

template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) copy_out_kernel(const TV* __restrict__ src, int n, int ncols, TV* X,
                                                             int64_t ldx, int64_t x_bstride) {
  const int64_t len = (int64_t)n * ncols;
  const TV* s = src + (int64_t)blockIdx.x * len;
  TV* Xb = X + (int64_t)blockIdx.x * x_bstride;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
    const int64_t row = i / ncols, c = i - row * ncols;
    Xb[row * ldx + c] = s[i];
  }
}

template <typename TV> static int finish(const xt_solve_args* g, SolveState<TV>& S, int64_t napply, cudaStream_t st,
                                         void* pre = nullptr) {
  if (pre != nullptr) {
    // x = P_r xt for the best iterate
    solve_final_kernel<TV><<<g->nbatch * S.nslices, SV_THREADS, 0, st>>>(S, S.ex[1], g->ncols, (int64_t)g->n * g->ncols); XT_LAUNCHED();
    reinterpret_cast<xt_apply_fn>(pre)(g->precond_user, S.ex[1], S.ex[2], st);
    XT_CHECK_ABORT(g->abort);
    copy_out_kernel<TV><<<g->nbatch, SV_THREADS, 0, st>>>(S.ex[2], g->n, g->ncols, static_cast<TV*>(g->X), g->ldx,
                                                          g->x_bstride); XT_LAUNCHED();
  } else {
    solve_final_kernel<TV><<<g->nbatch * S.nslices, SV_THREADS, 0, st>>>(S, static_cast<TV*>(g->X), g->ldx, g->x_bstride); XT_LAUNCHED();
  }
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

// step kernels go out as programmatic dependents of the matvec before them (and the matvec after them as theirs):
// launch latency and CTA scheduling of every kernel of an iteration overlap the tail of the one before
template <typename TV>
static inline void launch_init(SolveState<TV>& S, size_t smem, cudaStream_t st, double rtol, double atol, int bicg) {
  if (S.nslices > 1) {
    if (coop_launch(solve_init_kernel<TV>, S.nbatch * S.nslices, SV_THREADS, smem, st, S, rtol, atol, bicg)) return;
    S.nslices = 1;
  }
  solve_init_kernel<TV><<<S.nbatch, SV_THREADS, smem, st>>>(S, rtol, atol, bicg);
}
template <typename TV>
static inline void launch_cg_step(SolveState<TV>& S, size_t smem, cudaStream_t st, int k, int phase, unsigned int& epoch,
                                  int rel = 0) {
  const unsigned int e = epoch;     // cross-slice reductions of the solve so far (every reducing launch is one)
  if (phase != 1) ++epoch;
  if (S.nslices > 1) {
    if (coop_launch(cg_step_kernel<TV>, S.nbatch * S.nslices, SV_THREADS, smem, st, S, k, phase, e, rel)) return;
    S.nslices = 1;          // slices only partition the rows: from here on one CTA per batch item, no spinning
  }
#ifdef __CUDACC__
  if (dep_launch(cg_step_kernel<TV>, S.nbatch, SV_THREADS, smem, st, S, k, phase, e, rel)) return;
#endif
  cg_step_kernel<TV><<<S.nbatch, SV_THREADS, smem, st>>>(S, k, phase, e, rel);
}
template <typename TV>
static inline void launch_bicg_step(SolveState<TV>& S, size_t smem, cudaStream_t st, int k, int stage, unsigned int& epoch,
                                    int rel = 0) {
  const unsigned int e = epoch;
  if (stage == 1 || stage == 3 || stage == 5) ++epoch;
  if (S.nslices > 1) {
    if (coop_launch(bicg_step_kernel<TV>, S.nbatch * S.nslices, SV_THREADS, smem, st, S, k, stage, e, rel)) return;
    S.nslices = 1;
  }
#ifdef __CUDACC__
  if (dep_launch(bicg_step_kernel<TV>, S.nbatch, SV_THREADS, smem, st, S, k, stage, e, rel)) return;
#endif
  bicg_step_kernel<TV><<<S.nbatch, SV_THREADS, smem, st>>>(S, k, stage, e, rel);
}

// ---------------------------------------------------------------------------- launch-bound solves: CUDA graphs
// Below a few thousand rows one iteration is a handful of microseconds of GPU work behind 2-5 launches that cost the
// host more than that.  After the first `P` iterations (plain launches: short solves never pay for a graph) the next
// iterations are replayed from a graph of one PERIOD -- P iterations, the true-residual iteration (solve.py:160) last --
// whose step kernels number themselves relative to `SolveCtl::graph_base / epoch_base`, advanced by the last node.
// The device `done` flag of period g is read back while period g + 1 runs.  Instantiated graphs are kept per thread,
// keyed by everything the captured launches contain (state pointers = workspace, operator, sizes).
__global__ void graph_base_kernel(SolveCtl* ctl, int set, int diter, unsigned int depoch) {
  if (set) {
    ctl->graph_base = diter;
    ctl->epoch_base = depoch;
  } else {
    ctl->graph_base += diter;
    ctl->epoch_base += depoch;
  }
}

constexpr int XT_NO_GRAPH = 1;

// iterations per graph period for this solve's `resid_calc_every` (0: none): the true-residual iteration closes a period
static int graph_period_len(const xt_solve_args* g) {
  if (g->apply != nullptr || g->precond_l != nullptr || g->precond_r != nullptr) return 0;   // host code in the loop
  const int rce = g->resid_calc_every;
  const int P = rce <= 0 ? 8 : (rce >= 8 ? rce : rce * ((8 + rce - 1) / rce));
  if (P > 32 || g->max_niter < 3 * P) return 0;
  return P;
}

#ifdef __CUDACC__
struct GraphEntry {
  std::string key;
  cudaGraphExec_t exec;
  uint64_t stamp;
  int64_t napply;      // operator applications in one period
};
struct GraphTools {          // per thread and device
  int dev = -1;
  cudaStream_t capture = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int* pinned = nullptr;     // [2]
  std::vector<GraphEntry> cache;
  uint64_t clock = 0;
};
static GraphTools* graph_tools() {
  static thread_local std::vector<GraphTools*> all;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (GraphTools* t : all)
    if (t->dev == dev) return t;
  GraphTools* t = new GraphTools();
  t->dev = dev;
  if (cudaStreamCreateWithFlags(&t->capture, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&t->ev[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&t->ev[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void**>(&t->pinned), 2 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
    (void)cudaGetLastError();
    delete t;
    return nullptr;
  }
  all.push_back(t);
  return t;
}

// iterations per graph period (0: no graph for this solve)
static int graph_period(const xt_solve_args* g) {
  const char* e = getenv("XT_NO_SOLVE_GRAPH");         // read per solve: the tests switch it
  if ((e && e[0] == '1') || prof_on()) return 0;
  const double esize = g->dtype == XT_F64 ? 8.0 : (g->dtype == XT_BF16 ? 2.0 : 4.0);
  if ((double)g->nbatch * g->n * g->n * esize / 6.0e12 > 60e-6) return 0;     // not launch-bound
  return graph_period_len(g);
}

template <typename T> static void key_add(std::string& k, const T& v) {
  k.append(reinterpret_cast<const char*>(&v), sizeof(T));
}
template <typename TV>
static std::string graph_key(const char* method, const xt_solve_args* g, const SolveState<TV>& S, const OpDesc& op, int P) {
  std::string k(method);
  key_add(k, S);                 // zero-filled before it was set up: no stray padding bytes
  key_add(k, op.A); key_add(k, op.lda); key_add(k, op.a_bstride);
  key_add(k, op.M); key_add(k, op.ldm); key_add(k, op.m_bstride);
  key_add(k, op.E); key_add(k, op.e_bstride); key_add(k, op.pdl); key_add(k, op.dtype);
  key_add(k, g->resid_calc_every); key_add(k, P);
  return k;
}

// replays up to `nper` periods on `st`; enq(stream) enqueues one period (relative numbering) and returns its status.
// *launched = periods put on the stream, *done = 1 when the stop flag was seen.  XT_NO_GRAPH means that nothing was
// launched and the caller carries on with plain launches; negative values are errors.
template <typename EnqFn>
static int graph_phase(const std::string& key, EnqFn&& enq, SolveCtl* ctl, int P, unsigned int E, unsigned int epoch_now,
                       int nper, cudaStream_t st, int* launched, int* done, int64_t* napply) {
  *launched = 0;
  GraphTools* t = graph_tools();
  if (t == nullptr) return XT_NO_GRAPH;
  GraphEntry* ge = nullptr;
  for (GraphEntry& e : t->cache)
    if (e.key == key) ge = &e;
  if (ge == nullptr) {
    if (cudaStreamBeginCapture(t->capture, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      (void)cudaGetLastError();
      return XT_NO_GRAPH;
    }
    const int64_t before = *napply;
    const int rc = enq(t->capture);
    graph_base_kernel<<<1, 1, 0, t->capture>>>(ctl, 0, P, E);
    cudaGraph_t graph = nullptr;
    const cudaError_t err = cudaStreamEndCapture(t->capture, &graph);
    const int64_t per = *napply - before;
    *napply = before;
    cudaGraphExec_t exec = nullptr;
    if (rc != XT_OK || err != cudaSuccess || graph == nullptr ||
        cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      (void)cudaGetLastError();
      if (graph != nullptr) cudaGraphDestroy(graph);
      return XT_NO_GRAPH;
    }
    cudaGraphDestroy(graph);
    if (t->cache.size() >= 8) {          // evict the least recently used
      size_t lru = 0;
      for (size_t i = 1; i < t->cache.size(); ++i)
        if (t->cache[i].stamp < t->cache[lru].stamp) lru = i;
      cudaGraphExecDestroy(t->cache[lru].exec);
      t->cache.erase(t->cache.begin() + lru);
    }
    t->cache.push_back(GraphEntry{key, exec, 0, per});
    ge = &t->cache.back();
  }
  ge->stamp = ++t->clock;
  graph_base_kernel<<<1, 1, 0, st>>>(ctl, 1, P, epoch_now);      // the first P iterations went out as plain launches
  XT_CUDA_OK(cudaGetLastError());
  for (int gi = 0; gi < nper; ++gi) {
    XT_CUDA_OK(cudaGraphLaunch(ge->exec, st));
    XT_CUDA_OK(cudaMemcpyAsync(&t->pinned[gi & 1], &ctl->done, sizeof(int), cudaMemcpyDeviceToHost, st));
    XT_CUDA_OK(cudaEventRecord(t->ev[gi & 1], st));
    ++*launched;
    *napply += ge->napply;
    if (gi >= 1) {                                       // the flag of the period before, while this one runs
      XT_CUDA_OK(cudaEventSynchronize(t->ev[(gi - 1) & 1]));
      if (t->pinned[(gi - 1) & 1] != 0) {
        *done = 1;
        break;
      }
    }
  }
  return XT_OK;
}
#else
// host build (tools/emu_engine): with XT_EMU_GRAPH=1 a "replay" enqueues the period again with relative numbering, so the
// numbering, the base kernel, the lagged stop flag and the bookkeeping around the replay run on the CPU; capture and
// instantiation are the part only the hardware tests cover
static int graph_period(const xt_solve_args* g) {
  const char* e = getenv("XT_EMU_GRAPH");
  return (e && e[0] == '1') ? graph_period_len(g) : 0;
}
template <typename TV>
static std::string graph_key(const char*, const xt_solve_args*, const SolveState<TV>&, const OpDesc&, int) {
  return std::string();
}
template <typename EnqFn>
static int graph_phase(const std::string&, EnqFn&& enq, SolveCtl* ctl, int P, unsigned int E, unsigned int epoch_now,
                       int nper, cudaStream_t st, int* launched, int* done, int64_t*) {
  *launched = 0;
  int flag[2] = {0, 0};
  graph_base_kernel<<<1, 1, 0, st>>>(ctl, 1, P, epoch_now);
  for (int gi = 0; gi < nper; ++gi) {
    const int rc = enq(st);
    if (rc != XT_OK) return rc;
    graph_base_kernel<<<1, 1, 0, st>>>(ctl, 0, P, E);
    flag[gi & 1] = ctl->done;
    ++*launched;
    if (gi >= 1 && flag[(gi - 1) & 1] != 0) {
      *done = 1;
      break;
    }
  }
  return XT_OK;
}
#endif

template <typename TV> static int run_cg(const xt_solve_args* g) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(g->stream);
  Arena ar(g->workspace, g->workspace_bytes);
  SolveState<TV> S;
  TV* mx;
  int rc = setup_state<TV>(g, 7, ar, S, &mx);
  if (rc != XT_OK) return rc;
  OpDesc op{g->dtype, g->n, g->nbatch, g->ncols, g->A, g->lda, g->a_bstride, g->M, g->ldm, g->m_bstride,
            g->E, g->e_bstride};
  op.apply = g->apply; op.apply_user = g->apply_user; op.abort = g->abort;
  op.pdl = solve_pdl_enabled() ? 1 : 0;
  XT_CUDA_OK(cudaMemsetAsync(S.ctl, 0, sizeof(SolveCtl), st));
  XT_CUDA_OK(cudaMemsetAsync(S.slice_bar, 0, (size_t)g->nbatch * sizeof(unsigned int), st));
  const size_t smem = step_smem<TV>(g->ncols);
  launch_init<TV>(S, smem, st, g->rtol, g->atol, 0); XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  int64_t napply = 0;
  unsigned int epoch = 1;     // cross-slice reductions launched so far (number 0 was the init kernel's)
  const int ce = g->check_every > 0 ? g->check_every : 1;
  int next_check = ce < 4 ? ce : 4;      // poll the device flag at 4, 8, 16, ... iterations, then every `ce`
  const int* done_flag = &S.ctl->done;
  // preconditioner z = P r (solve.py:136,171): one more operator application per iteration, r.z fused into it
  OpDesc pc{g->dtype, g->n, g->nbatch, g->ncols, nullptr, 0, 0, nullptr, 0, 0, nullptr, 0};
  pc.apply = g->precond_l; pc.apply_user = g->precond_user; pc.abort = g->abort;
  S.precond = g->precond_l != nullptr ? 1 : 0;
  TV* zbuf = S.s;                         // unused by plain cg
  auto precond_step = [&](int first) -> int {
    int rc_ = apply_op<TV>(pc, S.r, zbuf, mx, S.r, S.dots, S.dots_gstride, done_flag, st, nullptr);
    if (rc_ != XT_OK) return rc_;
    cg_precond_kernel<TV><<<g->nbatch, SV_THREADS, smem, st>>>(S, zbuf, first); XT_LAUNCHED();
    return XT_OK;
  };
  if (S.precond) {
    rc = precond_step(1);
    if (rc != XT_OK) return rc;
  }
  // one iteration on stream `s`; rel = 1: k and the epoch are relative to the graph period (see graph_phase)
  auto iteration = [&](int k, int rel, cudaStream_t s, unsigned int& ep) -> int {
    int rc_ = apply_op<TV>(op, S.p, S.q, mx, S.p, S.dots, S.dots_gstride, done_flag, s, &napply);
    if (rc_ != XT_OK) return rc_;
    const bool true_resid = g->resid_calc_every != 0 && (k % g->resid_calc_every == 0);
    if (!true_resid) {
      launch_cg_step<TV>(S, smem, s, k, 0, ep, rel); XT_LAUNCHED();
    } else {
      launch_cg_step<TV>(S, smem, s, k, 1, ep, rel); XT_LAUNCHED();
      rc_ = apply_op<TV>(op, S.x, S.q, mx, nullptr, nullptr, 0, done_flag, s, &napply);
      if (rc_ != XT_OK) return rc_;
      launch_cg_step<TV>(S, smem, s, k, 2, ep, rel); XT_LAUNCHED();
    }
    return XT_OK;
  };
  const int P = graph_period(g);
  for (int k = 1; k <= g->max_niter; ++k) {
    if (P > 0 && k == P + 1) {
      // launch-bound solve still running after its first period: the following whole periods come from a graph
      int done = 0;
      rc = poll_done(S.ctl, st, &done);
      if (rc != XT_OK) return rc;
      if (done) break;
      const unsigned int E = epoch - 1u;                  // reductions per period (the first one just ran)
      int launched = 0;
      rc = graph_phase(graph_key<TV>("cg", g, S, op, P), [&](cudaStream_t cs) -> int {
        unsigned int ep = 0;
        for (int j = 1; j <= P; ++j) {
          const int rc_ = iteration(j, 1, cs, ep);
          if (rc_ != XT_OK) return rc_;
        }
        return XT_OK;
      }, S.ctl, P, E, epoch, (g->max_niter - P) / P, st, &launched, &done, &napply);
      if (rc < 0) return rc;
      if (rc == XT_OK) {
        k += launched * P;
        epoch += (unsigned int)launched * E;
        if (done || k > g->max_niter) break;
        next_check = g->max_niter;                        // fewer than P iterations left: one poll at the end
      }
    }
    rc = iteration(k, 0, st, epoch);
    if (rc != XT_OK) return rc;
    if (S.precond) {
      rc = precond_step(0);
      if (rc != XT_OK) return rc;
    }
    XT_CUDA_OK(cudaGetLastError());
    if (k == next_check || k == g->max_niter) {
      next_check += (next_check < ce) ? next_check : ce;
      int done = 0;
      rc = poll_done(S.ctl, st, &done);
      if (rc != XT_OK) return rc;
      if (done) break;
    }
  }
  return finish<TV>(g, S, napply, st);
}

template <typename TV> static int run_bicgstab(const xt_solve_args* g) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(g->stream);
  Arena ar(g->workspace, g->workspace_bytes);
  SolveState<TV> S;
  TV* mx;
  int rc = setup_state<TV>(g, 12, ar, S, &mx);
  if (rc != XT_OK) return rc;
  OpDesc op{g->dtype, g->n, g->nbatch, g->ncols, g->A, g->lda, g->a_bstride, g->M, g->ldm, g->m_bstride,
            g->E, g->e_bstride};
  op.apply = g->apply; op.apply_user = g->apply_user; op.abort = g->abort;
  op.pdl = solve_pdl_enabled() ? 1 : 0;
  // right preconditioner (solve.py:276,282): y = P_r p, z = P_r s and x = h + omega z are the plain recurrences of the
  // composed operator A o P_r on the iterate xt with x = P_r xt (applied once at the end); the residual is untouched
  op.pre = g->precond_r; op.pre_user = g->precond_user; op.pre_tmp = S.ex[0];
  // left preconditioner (solve.py:285-286): only omega = <K t, K s> / <K t, K t>
  OpDesc pl{g->dtype, g->n, g->nbatch, g->ncols, nullptr, 0, 0, nullptr, 0, 0, nullptr, 0};
  pl.apply = g->precond_l; pl.apply_user = g->precond_user; pl.abort = g->abort;
  XT_CUDA_OK(cudaMemsetAsync(S.ctl, 0, sizeof(SolveCtl), st));
  XT_CUDA_OK(cudaMemsetAsync(S.slice_bar, 0, (size_t)g->nbatch * sizeof(unsigned int), st));
  const size_t smem = step_smem<TV>(g->ncols);
  launch_init<TV>(S, smem, st, g->rtol, g->atol, 1); XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  int64_t napply = 0;
  unsigned int epoch = 1;     // cross-slice reductions launched so far (number 0 was the init kernel's)
  const int ce = g->check_every > 0 ? g->check_every : 1;
  int next_check = ce < 4 ? ce : 4;      // poll the device flag at 4, 8, 16, ... iterations, then every `ce`
  const int* done_flag = &S.ctl->done;
  // one iteration on stream `s`; rel = 1: k and the epoch are relative to the graph period (see graph_phase)
  auto iteration = [&](int k, int rel, cudaStream_t s, unsigned int& ep) -> int {
    launch_bicg_step<TV>(S, smem, s, k, 1, ep, rel); XT_LAUNCHED();
    int rc_ = apply_op<TV>(op, S.p, S.q, mx, S.rhat, S.dots, S.dots_gstride, done_flag, s, &napply);   // v = A p, rhat.v
    if (rc_ != XT_OK) return rc_;
    launch_bicg_step<TV>(S, smem, s, k, 2, ep, rel); XT_LAUNCHED();
    rc_ = apply_op<TV>(op, S.s, S.t, mx, S.s, S.dots, S.dots_gstride, done_flag, s, &napply);      // t = A s, t.s, t.t
    if (rc_ != XT_OK) return rc_;
    if (pl.apply != nullptr) {            // K s, then K t with <K s, K t> and <K t, K t> in place of t.s and t.t
      rc_ = apply_op<TV>(pl, S.s, S.ex[1], mx, nullptr, nullptr, 0, done_flag, s, nullptr);
      if (rc_ != XT_OK) return rc_;
      rc_ = apply_op<TV>(pl, S.t, S.ex[2], mx, S.ex[1], S.dots, S.dots_gstride, done_flag, s, nullptr);
      if (rc_ != XT_OK) return rc_;
    }
    const bool true_resid = g->resid_calc_every != 0 && (k % g->resid_calc_every == 0);
    if (!true_resid) {
      launch_bicg_step<TV>(S, smem, s, k, 3, ep, rel); XT_LAUNCHED();
    } else {
      launch_bicg_step<TV>(S, smem, s, k, 4, ep, rel); XT_LAUNCHED();
      rc_ = apply_op<TV>(op, S.x, S.t, mx, nullptr, nullptr, 0, done_flag, s, &napply);
      if (rc_ != XT_OK) return rc_;
      launch_bicg_step<TV>(S, smem, s, k, 5, ep, rel); XT_LAUNCHED();
    }
    return XT_OK;
  };
  const int P = graph_period(g);
  for (int k = 1; k <= g->max_niter; ++k) {
    if (P > 0 && k == P + 1) {
      int done = 0;
      rc = poll_done(S.ctl, st, &done);
      if (rc != XT_OK) return rc;
      if (done) break;
      const unsigned int E = epoch - 1u;
      int launched = 0;
      rc = graph_phase(graph_key<TV>("bicgstab", g, S, op, P), [&](cudaStream_t cs) -> int {
        unsigned int ep = 0;
        for (int j = 1; j <= P; ++j) {
          const int rc_ = iteration(j, 1, cs, ep);
          if (rc_ != XT_OK) return rc_;
        }
        return XT_OK;
      }, S.ctl, P, E, epoch, (g->max_niter - P) / P, st, &launched, &done, &napply);
      if (rc < 0) return rc;
      if (rc == XT_OK) {
        k += launched * P;
        epoch += (unsigned int)launched * E;
        if (done || k > g->max_niter) break;
        next_check = g->max_niter;
      }
    }
    rc = iteration(k, 0, st, epoch);
    if (rc != XT_OK) return rc;
    XT_CUDA_OK(cudaGetLastError());
    if (k == next_check || k == g->max_niter) {
      next_check += (next_check < ce) ? next_check : ce;
      int done = 0;
      rc = poll_done(S.ctl, st, &done);
      if (rc != XT_OK) return rc;
      if (done) break;
    }
  }
  return finish<TV>(g, S, napply, st, g->precond_r);
}

size_t gmres_ws_bytes(size_t vs, int n, int nbatch, int ncols, int max_niter);   // gmres.cu

}  // namespace xt

#ifdef __CUDACC__
namespace xt {
bool small_cg_applies(const xt_solve_args* g);
int run_small_cg(const xt_solve_args* g);
}  // namespace xt
#endif

extern "C" {

size_t xt_solve_workspace_bytes(const char* method, int32_t dtype, int32_t n, int32_t nbatch, int32_t ncols,
                                int32_t max_niter, int32_t has_M) {
  (void)has_M;
  const size_t vs = dtype == XT_F64 ? 8 : 4;
  if (method == nullptr) return 0;
  if (strcmp(method, "cg") == 0) return xt::solve_ws_bytes(7, vs, n, nbatch, ncols);
  if (strcmp(method, "bicgstab") == 0) return xt::solve_ws_bytes(12, vs, n, nbatch, ncols);
  if (strcmp(method, "gmres") == 0) return xt::gmres_ws_bytes(vs, n, nbatch, ncols, max_niter);
  return 0;
}

int xt_cg(const xt_solve_args* g) {
  int rc = xt::check_solve_args(g);
  if (rc != XT_OK) return rc;
#ifdef __CUDACC__
  // problems that fit the shared memory of one thread-block cluster: the whole solve in one launch (small_solve.cu)
  if (xt::small_cg_applies(g)) return xt::run_small_cg(g);
#endif
  return g->dtype == XT_F64 ? xt::run_cg<double>(g) : xt::run_cg<float>(g);
}

int xt_bicgstab(const xt_solve_args* g) {
  int rc = xt::check_solve_args(g);
  if (rc != XT_OK) return rc;
  return g->dtype == XT_F64 ? xt::run_bicgstab<double>(g) : xt::run_bicgstab<float>(g);
}

}  // extern "C"
