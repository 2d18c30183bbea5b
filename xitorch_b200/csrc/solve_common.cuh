// Shared pieces of the linear-solver drivers (solve.cu, gmres.cu).
#pragma once
#include "matvec.cuh"

#include <cstring>
#ifdef __CUDACC__
#define XT_SPIN_PAUSE() ((void)0)
#else
#include <thread>
#define XT_SPIN_PAUSE() std::this_thread::yield()      // host build: one host thread per CUDA thread
#endif
#include <vector>
#include <string>
#include <cmath>

namespace xt {

constexpr int SV_THREADS = 512;
constexpr int SV_MAX_SLICED_CTAS = 160;   // nbatch * nslices of a sliced step launch never exceeds this (one wave)
constexpr int SV_MAXCS = 8;     // column slots per thread  => ncols <= 32 * 8

// device-resident control block of one solve
struct SolveCtl {
  int done;            // 1 once the stop test passed
  int converged;
  int niter;           // iteration at which the stop test passed (or last executed)
  int improved_iter;   // iteration whose iterate is the best so far and still has to be copied to best_x
  int last_iter;       // last iteration whose residual norms were evaluated
  unsigned int counter;
  double best_resid;
  // launches replayed from a CUDA graph carry iteration numbers / reduction epochs relative to these (solve.cu)
  int graph_base;
  unsigned int epoch_base;
};

template <typename TV> struct SolveState {
  int n, nbatch, ncols;
  int tiles_per_batch;
  TV *x, *r, *p, *q, *s, *rhat, *bestx;   // contiguous (nbatch, n, ncols); q = A p (cg) / v (bicgstab); s,t bicgstab
  TV *t;
  const TV* B; int64_t ldb, b_bstride;
  double* dots;          // [ngroups][ntiles][2][16]
  int64_t dots_gstride;  // ntiles*2*16
  double *rz;            // cg: r.z ; bicgstab: rho            [nbatch*ncols]
  double *alpha, *omega; // bicgstab scalars                    [nbatch*ncols]
  double *rhonew;        //                                     [nbatch*ncols]
  double *stop;          // [nbatch*ncols]
  double *cta_max; int* cta_bad;   // [nbatch]
  SolveCtl* ctl;
  double eps;
  TV* ex[3];             // bicgstab scratch: pre(X) | K s | K t   (preconditioners)
  int precond;           // cg: z = P r comes from a separate operator application (the step kernel stops after the norms)
  // sliced step kernels: the rows of one batch item are split over `nslices` CTAs (grid = nbatch * nslices, all
  // co-resident: cooperative launch); column sums cross the slices through slice_part / slice_bar
  int nslices;
  double* slice_part;        // [nbatch * nslices][ncols]
  unsigned int* slice_bar;   // [nbatch] arrivals, monotone over the solve (zeroed with the control block)
};

struct RowRange { int lo, hi; };
template <typename TV> __device__ __forceinline__ RowRange slice_rows(const SolveState<TV>& S, int sl) {
  const int per = (S.n + S.nslices - 1) / S.nslices;
  RowRange r;
  r.lo = sl * per < S.n ? sl * per : S.n;
  r.hi = r.lo + per < S.n ? r.lo + per : S.n;
  return r;
}

// res[0..nc) holds this CTA's column sums over its rows; on return it holds the sums over all slices of batch item b,
// added in slice order (the same bits in every slice).  `epoch` = number of such reductions in earlier launches of
// this solve.  One reduction per launch: the partials of the previous one are no longer read when these are written.
template <typename TV>
__device__ __forceinline__ void slice_allreduce(const SolveState<TV>& S, int b, int sl, double* res, int nc,
                                                unsigned int epoch) {
  if (S.nslices == 1) return;
  double* mine = S.slice_part + ((size_t)b * S.nslices + sl) * nc;
  for (int c = threadIdx.x; c < nc; c += blockDim.x) mine[c] = res[c];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&S.slice_bar[b], 1u);
    const unsigned int target = (epoch + 1u) * (unsigned int)S.nslices;
    while (*reinterpret_cast<volatile unsigned int*>(&S.slice_bar[b]) < target) {
      XT_SPIN_PAUSE();
    }
    __threadfence();
  }
  __syncthreads();
  {   // one warp per column, lanes over the slices, butterfly: the same order (the same bits) in every slice
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = warp; c < nc; c += nw) {
      double sum = 0.0;
      for (int j = lane; j < S.nslices; j += 32) sum += __ldcg(&S.slice_part[((size_t)b * S.nslices + j) * nc + c]);
      sum = warp_sum(sum);
      if (lane == 0) res[c] = sum;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ double safedenom(double v, double eps) { return v == 0.0 ? eps : v; }

// per-column block reduction of up to NR quantities; result in smem res[NR][ncols] (double)
template <int NR>
__device__ __forceinline__ void col_reduce(const double (&part)[NR][SV_MAXCS], int ncols, int tx, int ty, int TX,
                                           int TY, double* scr /* [TY][NR][TX] */, double* res /* [NR][ncols] */) {
  for (int cs = 0; cs * TX < ncols; ++cs) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NR; ++k) scr[(ty * NR + k) * TX + tx] = part[k][cs];
    __syncthreads();
    const int c = cs * TX + tx;
    if (ty < NR && c < ncols) {
      double sum = 0.0;
      for (int y = 0; y < TY; ++y) sum += scr[(y * NR + ty) * TX + tx];
      res[ty * ncols + c] = sum;
    }
  }
  __syncthreads();
}

// sum of the fused matvec dot partials of batch item b, column c
template <typename TV>
__device__ __forceinline__ double tile_dot(const SolveState<TV>& S, int b, int c, int which) {
  const double* base = S.dots + (int64_t)(c / MV_MAXK) * S.dots_gstride;
  double sum = 0.0;
  for (int t = 0; t < S.tiles_per_batch; ++t)
    sum += base[((size_t)(b * S.tiles_per_batch + t) * 2 + which) * MV_MAXK + (c % MV_MAXK)];
  return sum;
}

// the same sums for every column at once, one warp per column (lanes stride over the tiles, butterfly at the end: a
// fixed order): out[c] for which = 0, out[ncols + c] for which = 1 when `both`.  All threads of the CTA call it; ends
// with a barrier.  The one-thread loop above is a chain of tiles_per_batch dependent-latency loads (~10 us at 148 tiles).
template <typename TV>
__device__ __forceinline__ void tile_dots_all(const SolveState<TV>& S, int b, bool both, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nq = both ? 2 * S.ncols : S.ncols;
  for (int q = warp; q < nq; q += nw) {
    const int which = q / S.ncols, c = q - which * S.ncols;
    const double* base = S.dots + (int64_t)(c / MV_MAXK) * S.dots_gstride;
    double sum = 0.0;
    for (int t = lane; t < S.tiles_per_batch; t += 32)
      sum += base[((size_t)(b * S.tiles_per_batch + t) * 2 + which) * MV_MAXK + (c % MV_MAXK)];
    sum = warp_sum(sum);
    if (lane == 0) out[q] = sum;
  }
  __syncthreads();
}

struct Geo {
  int tx, ty, TX, TY;
};
__device__ __forceinline__ Geo geo(int ncols) {
  Geo g;
  int TX = 1;
  while (TX < ncols && TX < 32) TX <<= 1;
  g.TX = TX;
  g.TY = SV_THREADS / TX;
  g.tx = threadIdx.x % TX;
  g.ty = threadIdx.x / TX;
  return g;
}

struct OpDesc {
  int dtype, n, nbatch, ncols;
  const void* A; int64_t lda, a_bstride;
  const void* M; int64_t ldm, m_bstride;
  const void* E; int64_t e_bstride;
  void* apply = nullptr;        // matrix-free operator callback (xt_solve_args.apply)
  void* apply_user = nullptr;
  void* pre = nullptr;          // right preconditioner callback: the operator applied is A o pre
  void* pre_user = nullptr;
  void* pre_tmp = nullptr;      // (nbatch, n, ncols) scratch for pre(X)
  const volatile int* abort = nullptr;   // host flag a callback sets to stop the solve (xt_solve_args.abort)
  int pdl = 0;                  // launch the dense matvecs as programmatic dependents of the step kernels (MvArgs.pdl)
};

// XT_NO_SOLVE_PDL=1: plain stream-ordered launches in cg / bicgstab (A/B switch)
static inline bool solve_pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("XT_NO_SOLVE_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

#ifdef __CUDACC__
// launch `kern` as a programmatic dependent of the kernel before it on `st`: its CTAs may be scheduled while the
// predecessor drains, and it must call pdl_wait() before touching anything the predecessor wrote.  false: not launched.
template <typename... KArgs, typename... Args>
static inline bool dep_launch(void (*kern)(KArgs...), int grid, int threads, size_t smem, cudaStream_t st, Args&&... args) {
  static std::atomic<int> ok{1};
  if (!ok.load(std::memory_order_relaxed) || !solve_pdl_enabled()) return false;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...) == cudaSuccess) return true;
  (void)cudaGetLastError();
  ok.store(0);
  return false;
}

// cooperative launch (all CTAs co-resident: the sliced step kernels spin on each other), as a programmatic dependent
// when that is enabled.  false: not launched (the caller falls back to one slice per batch item).
template <typename... KArgs, typename... Args>
static inline bool coop_launch(void (*kern)(KArgs...), int grid, int threads, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = solve_pdl_enabled() ? 2 : 1;
  if (cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...) == cudaSuccess) return true;
  (void)cudaGetLastError();
  return false;
}
#else
// host build (tools/emu_engine): all CTAs of the grid at once, one host thread per CUDA thread
template <typename F, typename... Args>
static inline bool coop_launch(F kern, int grid, int threads, size_t smem, cudaStream_t, Args... args) {
  emu_launch_coop(dim3(grid), dim3(threads), smem, [=]() { kern(args...); });
  return true;
}
#endif

#define XT_CHECK_ABORT(flag)                                                           \
  do {                                                                                 \
    if ((flag) != nullptr && *(flag) != 0) {                                           \
      xt::set_last_error("stopped by a callback (the operator's code failed)");         \
      return XT_ERR_ABORTED;                                                           \
    }                                                                                  \
  } while (0)

typedef void (*xt_apply_fn)(void* user, const void* X, void* Y, void* stream);

// per-tile partial dot products of a matrix-free operator application, in the layout the block matvec produces:
//   dots[group][tile][0][c] = sum_rows U Y,  dots[group][tile][1][c] = sum_rows Y^2      (one CTA per tile)
template <typename TV>
__global__ void __launch_bounds__(256)
tile_dots_kernel(const TV* __restrict__ U, const TV* __restrict__ Y, int n, int ncols, int tile_rows,
                 int tiles_per_batch, double* __restrict__ dots, int64_t dots_gstride, const int* done_flag) {
  if (done_flag != nullptr && *done_flag != 0) return;
  __shared__ double red[2][8];
  const int tile = blockIdx.x;
  const int b = tile / tiles_per_batch;
  const int row0 = (tile - b * tiles_per_batch) * tile_rows;
  const int rows = min(tile_rows, n - row0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = ((int64_t)b * n + row0) * ncols;
  for (int c = 0; c < ncols; ++c) {
    double d0 = 0.0, d1 = 0.0;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
      const double y = (double)Y[base + (int64_t)r * ncols + c];
      d1 += y * y;
      if (U != nullptr) d0 += (double)U[base + (int64_t)r * ncols + c] * y;
    }
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    __syncthreads();
    if (lane == 0) { red[0][warp] = d0; red[1][warp] = d1; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
      dots[(int64_t)(c / MV_MAXK) * dots_gstride + ((size_t)tile * 2 + threadIdx.x) * MV_MAXK + (c % MV_MAXK)] = s;
    }
  }
}

// Y = A X - (M X) E  (+ fused dots with U on the A pass); `mx` is scratch for M X
template <typename TV>
static inline int apply_op(const OpDesc& op, const TV* X, TV* Y, TV* mx, const TV* U, double* dots, int64_t dots_gstride,
                    const int* done_flag, cudaStream_t st, int64_t* napply) {
  const int64_t len = (int64_t)op.n * op.ncols;
  if (op.pre != nullptr) {
    // composed operator A o pre: the preconditioned block goes through scratch, then the operator proper
    reinterpret_cast<xt_apply_fn>(op.pre)(op.pre_user, X, op.pre_tmp, st);
    XT_CHECK_ABORT(op.abort);
    OpDesc inner = op;
    inner.pre = nullptr;
    return apply_op<TV>(inner, static_cast<const TV*>(op.pre_tmp), Y, mx, U, dots, dots_gstride, done_flag, st, napply);
  }
  if (op.apply != nullptr) {
    // matrix-free operator: the caller applies it, the dot products the dense path fuses into the matvec epilogue
    // come from one small kernel with the same per-tile layout
    reinterpret_cast<xt_apply_fn>(op.apply)(op.apply_user, X, Y, st);
    XT_CHECK_ABORT(op.abort);
    if (dots != nullptr) {
      const MvTiling til = mv_tiling(op.nbatch, op.n);
      tile_dots_kernel<TV><<<til.ntiles, 256, 0, st>>>(U, Y, op.n, op.ncols, til.tile_rows, til.tiles_per_batch, dots,
                                                       dots_gstride, done_flag); XT_LAUNCHED();
      XT_CUDA_OK(cudaGetLastError());
    }
    if (napply) ++(*napply);
    return XT_OK;
  }
  for (int c0 = 0, gi = 0; c0 < op.ncols; c0 += MV_MAXK, ++gi) {
    const int kg = (op.ncols - c0 < MV_MAXK) ? (op.ncols - c0) : MV_MAXK;
    MvArgs a;
    memset(&a, 0, sizeof(a));
    a.dtype = op.dtype;
    a.nbatch = op.nbatch; a.nrows = op.n; a.ncolsA = op.n; a.k = kg;
    a.X = X + c0; a.ldx = op.ncols; a.x_bstride = len;
    a.done_flag = done_flag;
    a.pdl = op.pdl;
    if (op.E != nullptr && op.M != nullptr) {
      a.A = op.M; a.lda = op.ldm; a.a_bstride = op.m_bstride;
      a.Y = mx + c0; a.ldy = op.ncols; a.y_bstride = len;
      int rc = mv_launch(a, st);
      if (rc != XT_OK) return rc;
    }
    a.A = op.A; a.lda = op.lda; a.a_bstride = op.a_bstride;
    a.Y = Y + c0; a.ldy = op.ncols; a.y_bstride = len;
    if (op.E != nullptr) {
      a.E = static_cast<const TV*>(op.E) + c0; a.e_bstride = op.e_bstride;
      if (op.M != nullptr) { a.Z = mx + c0; a.ldz = op.ncols; a.z_bstride = len; }
    }
    if (dots != nullptr) {
      a.U = U ? U + c0 : nullptr; a.ldu = op.ncols; a.u_bstride = len;
      a.dot_out = dots + gi * dots_gstride;
    }
    int rc = mv_launch(a, st);
    if (rc != XT_OK) return rc;
  }
  if (napply) ++(*napply);
  return XT_OK;
}

// CTAs per batch item of the step kernels: one wave of at most num_sms() co-resident CTAs, at least 128 rows each.
// XT_NO_SOLVE_SLICES=1: one slice.  Host build: one slice unless XT_EMU_SLICES asks for more (cooperative launches
// run all CTAs of the grid at once there, ordinary ones one after another).
static inline int step_slices(int n, int nbatch) {
#ifdef __CUDACC__
  const char* e = getenv("XT_NO_SOLVE_SLICES");        // read per solve: the tests switch it
  if (e && e[0] == '1') return 1;
  int cap = num_sms();
  if (cap > SV_MAX_SLICED_CTAS) cap = SV_MAX_SLICED_CTAS;
  int ns = cap / nbatch;
  if (ns > n / 128) ns = n / 128;
  return ns < 1 ? 1 : ns;
#else
  (void)nbatch;
  const char* e = getenv("XT_EMU_SLICES");             // the CPU tests run the sliced kernels with a few slices
  int ns = e ? atoi(e) : 1;
  if (ns > n / 8) ns = n / 8;
  return ns < 1 ? 1 : ns;
#endif
}

static inline int poll_done(SolveCtl* ctl, cudaStream_t st, int* done) {
  int h = 0;
  XT_CUDA_OK(cudaMemcpyAsync(&h, &ctl->done, sizeof(int), cudaMemcpyDeviceToHost, st));
  XT_CUDA_OK(cudaStreamSynchronize(st));
  *done = h;
  return XT_OK;
}


}  // namespace xt
