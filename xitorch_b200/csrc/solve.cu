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
__device__ __forceinline__ void deferred_best_copy(const SolveState<TV>& S, int b, int prev_iter) {
  if (S.ctl->improved_iter == prev_iter) {
    const int64_t len = (int64_t)S.n * S.ncols;
    const TV* src = S.x + (int64_t)b * len;
    TV* dst = S.bestx + (int64_t)b * len;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
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
  const int b = blockIdx.x;
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
      for (int row = g.ty; row < S.n; row += g.TY) {
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
  if (threadIdx.x == 0) {
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
__global__ void __launch_bounds__(SV_THREADS) cg_step_kernel(SolveState<TV> S, int iter, int phase) {
  pdl_wait();        // may have been scheduled while the matvec before it drains
  pdl_trigger();     // the matvec after it may start streaming A now (it waits for this grid before reading vectors)
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  const Geo g = geo(S.ncols);
  double* res = sm;                       // [ncols]   (r.r)
  double* alp = sm + S.ncols;             // [ncols]
  double* scr = sm + 2 * S.ncols;
  const int64_t len = (int64_t)S.n * S.ncols;
  const int64_t base = (int64_t)b * len;

  if (phase != 2) deferred_best_copy(S, b, iter - 1);
  if (S.ctl->done) return;

  if (phase != 2) {
    for (int c = threadIdx.x; c < S.ncols; c += blockDim.x)
      alp[c] = S.rz[b * S.ncols + c] / safedenom(tile_dot(S, b, c, 0), S.eps);
    __syncthreads();
  }
  double part[1][SV_MAXCS];
#pragma unroll
  for (int i = 0; i < SV_MAXCS; ++i) part[0][i] = 0.0;
  const TV* Bb = S.B + (int64_t)b * S.b_bstride;
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols) {
      const TV a = (phase != 2) ? (TV)alp[c] : TV(0);
      for (int row = g.ty; row < S.n; row += g.TY) {
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
  norm_bookkeeping(S, b, res, iter, scr);
  if (S.precond) return;          // z = P r, beta and p follow in cg_precond_kernel
  // beta = rz_new / safedenom(rz);  p = r + beta p
  for (int cs = 0; cs * g.TX < S.ncols; ++cs) {
    const int c = cs * g.TX + g.tx;
    if (c < S.ncols) {
      const TV beta = (TV)(res[c] / safedenom(S.rz[b * S.ncols + c], S.eps));
      for (int row = g.ty; row < S.n; row += g.TY) {
        const int64_t o = base + (int64_t)row * S.ncols + c;
        S.p[o] = S.r[o] + beta * S.p[o];
      }
    }
  }
  __syncthreads();
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
  for (int c = threadIdx.x; c < S.ncols; c += blockDim.x) {
    const double rzn = tile_dot(S, b, c, 0);
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
__global__ void __launch_bounds__(SV_THREADS) bicg_step_kernel(SolveState<TV> S, int iter, int stage) {
  pdl_wait();        // may have been scheduled while the matvec before it drains
  pdl_trigger();     // the matvec after it may start streaming A now (it waits for this grid before reading vectors)
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  const Geo g = geo(S.ncols);
  double* res = sm;                       // [2][ncols]
  double* sc = sm + 2 * S.ncols;          // [ncols] scalar scratch
  double* scr = sm + 3 * S.ncols;
  const int64_t len = (int64_t)S.n * S.ncols;
  const int64_t base = (int64_t)b * len;
  const int nc = S.ncols;

  if (stage == 1) deferred_best_copy(S, b, iter - 1);
  if (S.ctl->done) return;

  double part[1][SV_MAXCS];
#pragma unroll
  for (int i = 0; i < SV_MAXCS; ++i) part[0][i] = 0.0;

  if (stage == 1) {
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc)
        for (int row = g.ty; row < S.n; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          part[0][cs] += (double)S.rhat[o] * (double)S.r[o];
        }
    }
    col_reduce<1>(part, nc, g.tx, g.ty, g.TX, g.TY, scr, res);
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
      const int i = b * nc + c;
      // reference order (solve.py:273-275): omega and rho are "safed" in place before use
      double om = S.omega[i];
      if (om == 0.0) { om = S.eps; S.omega[i] = om; }
      double rho = S.rz[i];
      if (rho == 0.0) { rho = S.eps; S.rz[i] = rho; }
      S.rhonew[i] = res[c];
      sc[c] = res[c] / rho * (S.alpha[i] / om);
    }
    __syncthreads();
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc) {
        const TV beta = (TV)sc[c];
        const TV om = (TV)S.omega[b * nc + c];
        for (int row = g.ty; row < S.n; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          S.p[o] = S.r[o] + beta * (S.p[o] - om * S.q[o]);
        }
      }
    }
    return;
  }
  if (stage == 2) {
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
      const int i = b * nc + c;
      const double a = S.rhonew[i] / safedenom(tile_dot(S, b, c, 0), S.eps);
      S.alpha[i] = a;
      sc[c] = a;
    }
    __syncthreads();
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc) {
        const TV a = (TV)sc[c];
        for (int row = g.ty; row < S.n; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          S.x[o] = S.x[o] + a * S.p[o];
          S.s[o] = S.r[o] - a * S.q[o];
        }
      }
    }
    return;
  }
  if (stage == 3 || stage == 4) {
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
      const int i = b * nc + c;
      const double om = tile_dot(S, b, c, 0) / safedenom(tile_dot(S, b, c, 1), S.eps);
      S.omega[i] = om;
      sc[c] = om;
    }
    __syncthreads();
    for (int cs = 0; cs * g.TX < nc; ++cs) {
      const int c = cs * g.TX + g.tx;
      if (c < nc) {
        const TV om = (TV)sc[c];
        for (int row = g.ty; row < S.n; row += g.TY) {
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
        for (int row = g.ty; row < S.n; row += g.TY) {
          const int64_t o = base + (int64_t)row * nc + c;
          const TV rn = Bb[(int64_t)row * S.ldb + c] - S.t[o];
          S.r[o] = rn;
          part[0][cs] += (double)rn * (double)rn;
        }
    }
  }
  col_reduce<1>(part, nc, g.tx, g.ty, g.TX, g.TY, scr, res);
  norm_bookkeeping(S, b, res, iter, scr);
  __syncthreads();
  for (int c = threadIdx.x; c < nc; c += blockDim.x) S.rz[b * nc + c] = S.rhonew[b * nc + c];
}

// ---------------------------------------------------------------------------- finalize: X_out = best iterate
template <typename TV>
__global__ void __launch_bounds__(SV_THREADS) solve_final_kernel(SolveState<TV> S, TV* X, int64_t ldx,
                                                                int64_t x_bstride) {
  const int b = blockIdx.x;
  const int64_t len = (int64_t)S.n * S.ncols;
  const bool pending = (S.ctl->improved_iter == S.ctl->last_iter) && (S.ctl->last_iter > 0);
  const TV* src = (pending ? S.x : S.bestx) + (int64_t)b * len;
  TV* Xb = X + (int64_t)b * x_bstride;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
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
  return (size_t)(3 * ncols + 2 * SV_THREADS + 64) * sizeof(double);
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
    solve_final_kernel<TV><<<g->nbatch, SV_THREADS, 0, st>>>(S, S.ex[1], g->ncols, (int64_t)g->n * g->ncols); XT_LAUNCHED();
    reinterpret_cast<xt_apply_fn>(pre)(g->precond_user, S.ex[1], S.ex[2], st);
    XT_CHECK_ABORT(g->abort);
    copy_out_kernel<TV><<<g->nbatch, SV_THREADS, 0, st>>>(S.ex[2], g->n, g->ncols, static_cast<TV*>(g->X), g->ldx,
                                                          g->x_bstride); XT_LAUNCHED();
  } else {
    solve_final_kernel<TV><<<g->nbatch, SV_THREADS, 0, st>>>(S, static_cast<TV*>(g->X), g->ldx, g->x_bstride); XT_LAUNCHED();
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
static inline void launch_cg_step(const SolveState<TV>& S, int nbatch, size_t smem, cudaStream_t st, int k, int phase) {
#ifdef __CUDACC__
  if (dep_launch(cg_step_kernel<TV>, nbatch, SV_THREADS, smem, st, S, k, phase)) return;
#endif
  cg_step_kernel<TV><<<nbatch, SV_THREADS, smem, st>>>(S, k, phase);
}
template <typename TV>
static inline void launch_bicg_step(const SolveState<TV>& S, int nbatch, size_t smem, cudaStream_t st, int k, int stage) {
#ifdef __CUDACC__
  if (dep_launch(bicg_step_kernel<TV>, nbatch, SV_THREADS, smem, st, S, k, stage)) return;
#endif
  bicg_step_kernel<TV><<<nbatch, SV_THREADS, smem, st>>>(S, k, stage);
}

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
  const size_t smem = step_smem<TV>(g->ncols);
  solve_init_kernel<TV><<<g->nbatch, SV_THREADS, smem, st>>>(S, g->rtol, g->atol, 0); XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  int64_t napply = 0;
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
  for (int k = 1; k <= g->max_niter; ++k) {
    rc = apply_op<TV>(op, S.p, S.q, mx, S.p, S.dots, S.dots_gstride, done_flag, st, &napply);
    if (rc != XT_OK) return rc;
    const bool true_resid = g->resid_calc_every != 0 && (k % g->resid_calc_every == 0);
    if (!true_resid) {
      launch_cg_step<TV>(S, g->nbatch, smem, st, k, 0); XT_LAUNCHED();
    } else {
      launch_cg_step<TV>(S, g->nbatch, smem, st, k, 1); XT_LAUNCHED();
      rc = apply_op<TV>(op, S.x, S.q, mx, nullptr, nullptr, 0, done_flag, st, &napply);
      if (rc != XT_OK) return rc;
      launch_cg_step<TV>(S, g->nbatch, smem, st, k, 2); XT_LAUNCHED();
    }
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
  const size_t smem = step_smem<TV>(g->ncols);
  solve_init_kernel<TV><<<g->nbatch, SV_THREADS, smem, st>>>(S, g->rtol, g->atol, 1); XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  int64_t napply = 0;
  const int ce = g->check_every > 0 ? g->check_every : 1;
  int next_check = ce < 4 ? ce : 4;      // poll the device flag at 4, 8, 16, ... iterations, then every `ce`
  const int* done_flag = &S.ctl->done;
  for (int k = 1; k <= g->max_niter; ++k) {
    launch_bicg_step<TV>(S, g->nbatch, smem, st, k, 1); XT_LAUNCHED();
    rc = apply_op<TV>(op, S.p, S.q, mx, S.rhat, S.dots, S.dots_gstride, done_flag, st, &napply);   // v = A p, rhat.v
    if (rc != XT_OK) return rc;
    launch_bicg_step<TV>(S, g->nbatch, smem, st, k, 2); XT_LAUNCHED();
    rc = apply_op<TV>(op, S.s, S.t, mx, S.s, S.dots, S.dots_gstride, done_flag, st, &napply);      // t = A s, t.s, t.t
    if (rc != XT_OK) return rc;
    if (pl.apply != nullptr) {            // K s, then K t with <K s, K t> and <K t, K t> in place of t.s and t.t
      rc = apply_op<TV>(pl, S.s, S.ex[1], mx, nullptr, nullptr, 0, done_flag, st, nullptr);
      if (rc != XT_OK) return rc;
      rc = apply_op<TV>(pl, S.t, S.ex[2], mx, S.ex[1], S.dots, S.dots_gstride, done_flag, st, nullptr);
      if (rc != XT_OK) return rc;
    }
    const bool true_resid = g->resid_calc_every != 0 && (k % g->resid_calc_every == 0);
    if (!true_resid) {
      launch_bicg_step<TV>(S, g->nbatch, smem, st, k, 3); XT_LAUNCHED();
    } else {
      launch_bicg_step<TV>(S, g->nbatch, smem, st, k, 4); XT_LAUNCHED();
      rc = apply_op<TV>(op, S.x, S.t, mx, nullptr, nullptr, 0, done_flag, st, &napply);
      if (rc != XT_OK) return rc;
      launch_bicg_step<TV>(S, g->nbatch, smem, st, k, 5); XT_LAUNCHED();
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
