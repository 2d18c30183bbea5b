/*
 * xitorch_b200 -- C ABI of the B200-native (sm_100a) Krylov hot path of xitorch.
 *
 * The reference (xitorch @ ad9eec1) is pure Python and has no FFI; its plug-in
 * surface for this path is
 *   B1  the `method=` callables of linalg.symeig / linalg.solve
 *         xitorch/linalg/symeig.py:275-280   fcn(A, neig, mode, M, **opts) -> (evals, evecs)
 *         xitorch/linalg/solve.py:144-153    fcn(A, B, E, M, **opts)       -> X
 *   B2  the dense operator the methods call
 *         xitorch/_core/linop.py:692-702     MatrixLinearOperator._mv/_mm/_rmv/_rmm = torch.matmul
 * Every entry point below is what a binding for one of those would call; each cites the
 * reference function it replaces.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - all pointers except `*_out` / `xt_*_args` themselves are DEVICE pointers of torch CUDA tensors
 *   - the library never allocates, frees or retains device memory: the caller passes a workspace
 *     (size from xt_*_workspace_bytes) and owns every buffer
 *   - work is enqueued on `stream` (a cudaStream_t); solver entry points synchronise that stream
 *     only to read their convergence flag, matvec never synchronises
 *   - return value: XT_OK, or a negative xt_status; xt_last_error() gives the message
 *     (thread-local).  Non-convergence is NOT an error: converged_out = 0 and the best iterate
 *     is returned, the caller emits the ConvergenceWarning (reference convention,
 *     xitorch/_impls/linalg/solve.py:182-186)
 *   - dense matrices are row-major with a row stride `ld*` (elements) and a batch stride
 *     (elements, 0 = the same matrix for every batch item); vectors/blocks are (n, ncols)
 *     row-major -- torch's natural layout for a (*, n, ncols) tensor
 *   - dtype = storage type of A (and M); vectors, scalars and results are
 *     float for XT_F32 / XT_BF16 and double for XT_F64
 *   - re-entrant.  Process-wide state: the per-device SM-count cache, per-device "kernel attributes set" bits, a
 *     thread-local pool of side streams / events per device for the eigensolver, a thread-local cache of instantiated
 *     CUDA graphs (with its capture stream, two events and two pinned words) per device for launch-bound cg / bicgstab
 *     solves, and the xt_profile_* diagnostics (two atomic launch counters; the event list is mutex-protected and only
 *     filled while profiling is enabled)
 */
#ifndef XITORCH_B200_H
#define XITORCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { XT_F32 = 0, XT_BF16 = 1, XT_F64 = 2 } xt_dtype;

typedef enum {
  XT_OK = 0,
  XT_ERR_INVALID = -1,     /* bad argument / unsupported shape or alignment */
  XT_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed */
  XT_ERR_WORKSPACE = -3,   /* workspace too small */
  XT_ERR_BREAKDOWN = -4,   /* numerical breakdown that cannot be recovered (e.g. non-finite input) */
  XT_ERR_ABORTED = -5      /* a callback asked to stop through `abort` (e.g. user code raised): nothing else is called */
} xt_status;

int xt_version(void);
const char* xt_last_error(void);

/* Launch accounting and in-situ timing of the dominant kernel (used by bench.py for the roofline):
 * xt_profile_reset(enable) zeroes the counters; with enable != 0 every block-matvec launch is bracketed
 * by CUDA events on its stream.  xt_profile_read synchronises on the last event and returns the summed
 * matvec kernel time (ms), the number of matvec launches and the number of ALL kernel launches. */
void xt_profile_reset(int enable);
int xt_profile_read(double* matvec_ms, int64_t* matvec_launches, int64_t* total_launches);

/* ---------------------------------------------------------------------------------------------
 * Dense block matvec   Y_b = A_b X_b  [ - Z_b diag(E_b) ]        b = 0..nbatch-1
 * replaces MatrixLinearOperator._mm/_mv (xitorch/_core/linop.py:692-696) and the shifted operator
 * x -> A x - (M x) E of _setup_linear_problem (xitorch/_impls/linalg/solve.py:590-595; pass Z = M X,
 * or Z = NULL for Z = X).
 *   A: (nbatch, nrows, ncolsA) row-major, row stride lda;   X: (nbatch, ncolsA, k), row stride ldx
 *   Y: (nbatch, nrows, k), row stride ldy;                  E: (nbatch, k) or NULL
 *   k >= 1 (handled in column groups of <= 16).  `impl`: 0 = auto, 1 = force a TMA kernel (auto layout),
 *   2 = force the plain-load kernel, 3 = TMA row-slice layout, 4 = TMA column-slice layout (fp32, k > 4),
 *   5 = TMA row-slice layout with one row per thread (k = 8 defaults to two), 6 = tensor-core layout (fp32, k <= 16,
 *   error-compensated TF32 through mma.sync; exact to fp32 rounding but slower than the SIMT layouts on B200, see
 *   csrc/matvec.cu); 2-6 are used by the tests to cross-check the kernels.
 *   Bit 8 of `impl` (+256) makes the pass read A's columns last-to-first and bits 16..23 give the MB of the end of
 *   the pass to leave in L2 (evict-last): callers that apply the same A repeatedly alternate the direction so that
 *   each pass starts on the part of A the previous one left in the 126 MB L2.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t dtype;
  int32_t nbatch, nrows, ncolsA, k;
  const void* A; int64_t lda, a_bstride;
  const void* X; int64_t ldx, x_bstride;
  void* Y;       int64_t ldy, y_bstride;
  const void* E; int64_t e_bstride;
  const void* Z; int64_t ldz, z_bstride;
  int32_t impl;
  void* stream;
  /* trans != 0:  Y_b = A_b^T X_b  with X: (nbatch, nrows, k), Y: (nbatch, ncolsA, k) -- replaces the `mat^H @ x` of
   * MatrixLinearOperator._rmm/_rmv (xitorch/_core/linop.py:698-702) and the A^H (A x) of the normal equations
   * (xitorch/_impls/linalg/solve.py:637-643) without materialising the transpose: A is read once, by column strips.
   * fp32 / fp64, E must be NULL, A 16-byte aligned with a 16-byte multiple row stride. */
  int32_t trans;
} xt_matvec_args;

int xt_block_matvec(const xt_matvec_args* args);

/* ---------------------------------------------------------------------------------------------
 * Linear solvers  A X - M X diag(E) = B   (all columns and batch items in lock-step, global stop
 * test ||r||_2(col) < max(rtol ||b||, atol), best iterate returned) -- replace
 *   cg        xitorch/_impls/linalg/solve.py:69-190
 *   bicgstab  xitorch/_impls/linalg/solve.py:192-324
 *   gmres     xitorch/_impls/linalg/solve.py:326-433   (E/M unsupported, as in the reference)
 * The operator must already be the one to iterate on: the posdef probe / normal-equation
 * switch of _setup_linear_problem (solve.py:605-643) is host logic done by the caller.
 *   A, M: (nbatch, n, n);  B, X: (nbatch, n, ncols), row strides ldb/ldx;  E: (nbatch, ncols)|NULL
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t dtype;
  int32_t n, nbatch, ncols;
  const void* A; int64_t lda, a_bstride;
  const void* M; int64_t ldm, m_bstride;      /* NULL = identity */
  const void* E; int64_t e_bstride;           /* NULL = no shift */
  const void* B; int64_t ldb, b_bstride;
  void* X;       int64_t ldx, x_bstride;      /* out */
  double rtol, atol, eps;
  int32_t max_niter;
  int32_t resid_calc_every;                   /* 0 = never recompute the true residual */
  int32_t check_every;                        /* host polls the device convergence flag every this many iterations (>=1) */
  int32_t* niter_out;                         /* host, may be NULL */
  int32_t* converged_out;                     /* host, may be NULL */
  double* best_resid_out;                     /* host, may be NULL: max column residual norm of the returned iterate */
  int64_t* napply_out;                        /* host, may be NULL: number of block operator applications */
  void* workspace; size_t workspace_bytes;
  void* stream;
  /* matrix-free operator (reference: any LinearOperator whose _mv is user code, e.g. the autograd Jacobian of
   * xitorch/grad/jachess.py:98-208 that the rootfinder backward solves with, xitorch/optimize/rootfinder.py:346-348).
   * When `apply` is non-NULL, A / M / E are ignored and every operator application calls
   *     apply(apply_user, X, Y, stream):   Y = Op(X)   for X, Y contiguous (nbatch, n, ncols) device blocks inside
   * the workspace, ordered on `stream`.  The solver loop (all vector updates, dot products, the stop test) still
   * runs in the library's kernels; only the operator is the caller's. */
  void* apply;
  void* apply_user;
  /* preconditioners (reference: `precond` of cg, solve.py:122,136,171; `precond_l` / `precond_r` of bicgstab,
   * solve.py:247-248,276-287), applied through callbacks of the same signature as `apply` (a dense preconditioner's
   * callback is one xt_block_matvec).  cg: z = precond_l(r).  bicgstab: y = precond_r(p), z = precond_r(s) (carried
   * out as the composed operator A o precond_r on the un-preconditioned iterate), and
   * omega = <K t, K s> / <K t, K t> with K = precond_l.  NULL = identity.  Ignored by gmres. */
  void* precond_l;
  void* precond_r;
  void* precond_user;
  /* optional HOST flag a callback may set non-zero (user code failed): the library then returns XT_ERR_ABORTED right
   * after that callback instead of iterating on garbage until max_niter */
  const int32_t* abort;
} xt_solve_args;

size_t xt_solve_workspace_bytes(const char* method, int32_t dtype, int32_t n, int32_t nbatch,
                                int32_t ncols, int32_t max_niter, int32_t has_M);
int xt_cg(const xt_solve_args* args);
int xt_bicgstab(const xt_solve_args* args);
int xt_gmres(const xt_solve_args* args);

/* ---------------------------------------------------------------------------------------------
 * Extreme eigenpairs of a dense Hermitian operator by block Rayleigh-Ritz on a growing Krylov
 * subspace -- replaces davidson (xitorch/_impls/linalg/symeig.py:100-227) incl. tallqr
 * (xitorch/_utils/tensor.py:8-19); expansion = 0 appends the orthonormalised Ritz residuals
 * (the reference's Davidson step, symeig.py:207-220), expansion = 1 appends the orthonormalised
 * A*(last block) (block Lanczos with full reorthogonalisation: the same Krylov space, method
 * "lanczos" of BASELINE.json).  Stop: max |A X - X Lambda| < min_eps (symeig.py:188,200);
 * the best pair by that measure is returned (symeig.py:196-199).
 *   A: (nbatch, n, n) (batch items are solved one after another);
 *   V0: (nbatch, n, neig) start block, any full-rank block (it is orthonormalised here)
 *   evals: (nbatch, neig) ascending;  evecs: (nbatch, n, neig), row stride ldv
 *   max_basis: thick-restart cap on the subspace dimension (multiple of neig, >= 3*neig)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t dtype;
  int32_t n, nbatch, neig;
  int32_t mode;                               /* 0 = lowest, 1 = uppest */
  int32_t expansion;                          /* 0 = davidson (residuals), 1 = lanczos (A * last block) */
  const void* A; int64_t lda, a_bstride;
  const void* V0; int64_t ldv0, v0_bstride;
  void* evals;   int64_t evals_bstride;
  void* evecs;   int64_t ldv, evecs_bstride;
  int32_t max_niter, max_basis, check_every;  /* check_every: accepted and ignored -- the eigen-engine watches a host-mapped
                                                 copy of the stop flag through a 3-iteration run-ahead window */
  double min_eps;
  int32_t* niter_out;                         /* host, may be NULL: iterations of the last batch item */
  int32_t* converged_out;                     /* host, may be NULL: 1 iff every batch item met min_eps */
  double* best_resid_out;                     /* host, may be NULL: max over batch of the returned max|R| */
  int64_t* napply_out;                        /* host, may be NULL: block operator applications (all items) */
  void* workspace; size_t workspace_bytes;
  void* stream;
  /* row-partitioned operator over `world` GPUs (world <= 1: A is the whole matrix).  A then holds
   * this rank's rows [rank*n/world, (rank+1)*n/world) as an (n/world, n) row-major block, and `allgather` is
   * called once per operator application: void allgather(user, buf, count_per_rank, elem_size, stream) must
   * all-gather IN PLACE the `world` equal chunks of `count_per_rank` elements stored back to back in `buf`
   * (chunk r is this rank's contribution when r == rank), ordered on `stream`.  nbatch must be 1. */
  int32_t world, rank;
  void* allgather;
  void* allgather_user;
  /* matrix-free operator (any LinearOperator of the reference: user _mv, composite operators such as A^H A of svd,
   * autograd Jacobians / Hessians; xitorch/_impls/linalg/symeig.py:155,165 call A.mm on whatever operator they get).
   * When non-NULL, `A` is ignored and every block application is handed to the caller:
   *   void apply(void* user, const void* X, void* Y, void* stream)      Y = A X,
   * X and Y contiguous (n, neig) row-major blocks of the value type inside the workspace, ordered on `stream`.
   * Everything else of the iteration stays in the library's kernels.  The stop flag cannot gate user code: up to three
   * applications may run after convergence (their results are never read).  Needs nbatch = 1 and world <= 1. */
  void* apply;
  void* apply_user;
  /* row-SHARDED engine over peer memory (world >= 1, expansion = 1, nbatch = 1; takes precedence over `allgather`).
   * `peers[r]` is the address, in THIS process, of rank r's exchange region of xt_symeig_peer_bytes(...) bytes
   * (peers[rank] is this rank's own region; the others are CUDA-IPC mappings, see xt_peer_*).  Every rank keeps only
   * its n/world rows of the basis and of A*basis; per iteration the kernels exchange two small partial-sum blocks
   * (projection coefficients, Gram matrix) and all-gather the new n x neig basis block by direct stores into the
   * peers' regions -- no collective library call and no host round trip inside the iteration.  All ranks take the
   * same stop decision at the same iteration (bit-identical projected matrices).  `evecs` then receives this rank's
   * rows only: (n/world, neig).  `epoch` must be the same on all ranks and increase by one per call using the same
   * regions (the regions must be zero when first used).  restart_keep: Ritz vectors kept at a thick restart
   * (0 = 2*neig). */
  const void* const* peers;
  uint32_t epoch;
  int32_t restart_keep;
  const int32_t* abort;                       /* optional host flag set by the `apply` callback: see xt_solve_args.abort */
} xt_symeig_args;

size_t xt_symeig_workspace_bytes(int32_t dtype, int32_t n, int32_t neig, int32_t max_basis, int32_t world);
int xt_symeig_krylov(const xt_symeig_args* args);

/* workspace / exchange-region sizes of the row-sharded engine (0 = unsupported arguments) */
size_t xt_symeig_sharded_workspace_bytes(int32_t dtype, int32_t n, int32_t neig, int32_t max_basis, int32_t world);
size_t xt_symeig_peer_bytes(int32_t dtype, int32_t n, int32_t neig, int32_t max_basis, int32_t world);

/* Peer-visible device memory for the exchange regions (the one place where the library allocates: CUDA IPC needs
 * whole cudaMalloc allocations, which a caching allocator's sub-blocks are not).  xt_peer_alloc: cudaMalloc + zero,
 * `handle_out` receives the 64-byte cudaIpcMemHandle_t to be sent to the other ranks (any transport);
 * xt_peer_open maps another rank's handle into this process (device = current device, peer access enabled lazily);
 * xt_peer_close / xt_peer_free undo them. */
#define XT_PEER_HANDLE_BYTES 64
#define XT_MAX_WORLD 8
int xt_peer_alloc(size_t bytes, void** ptr_out, void* handle_out);
int xt_peer_open(const void* handle, void** ptr_out);
int xt_peer_close(void* ptr);
int xt_peer_free(void* ptr);

/* small dense symmetric eigensolver used for the projected problem (device, one CTA; replaces torch.linalg.eigh
 * at xitorch/_impls/linalg/symeig.py:174): the nev lowest (mode 0) or highest (mode 1) eigenpairs of the m x m
 * row-major fp64 matrix T -> w_out[nev] ascending, S_out (m x nev row-major, orthonormal columns).
 * scratch: >= m*(m|1) + 16 doubles (the last 16 receive phase clock stamps).  Exposed for the parity tests. */
int xt_small_eigh(const double* T, int32_t m, int32_t nev, int32_t mode, double* w_out, double* S_out,
                  double* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Hermiticity check of a dense real operator in one pass over the matrix -- replaces
 * torch.allclose(mat, mat.transpose(-2, -1).conj()) of LinearOperator.m (xitorch/_core/linop.py:96-103), with the same
 * rule: |A_ij - A_ji| <= atol + rtol |A_ji| for every ordered pair, NaN never passes.
 *   A: (nbatch, n, n) row-major, row stride lda, batch stride a_bstride (elements);
 *   mismatch: device int32, zeroed by the caller; set to 1 when a violating pair is found.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t dtype;                              /* XT_F32 or XT_F64 */
  int32_t n, nbatch;
  const void* A; int64_t lda, a_bstride;
  double rtol, atol;
  int32_t* mismatch;
  void* stream;
} xt_hermcheck_args;

int xt_hermitian_check(const xt_hermcheck_args* args);

#ifdef __cplusplus
}
#endif
#endif /* XITORCH_B200_H */
