// Internal interface of the dense block matvec (shared by the solver drivers).
#pragma once
#include "common.cuh"

namespace xt {

constexpr int MV_MAXK = 16;          // columns handled by one pass over A
constexpr int MV_TILE_ROWS = 128;    // max rows of a CTA tile
constexpr int MV_BOX_ROWS = 8;       // rows of one TMA box (= one 128B-swizzle atom: 8 x 128 B)
constexpr int MV_CONSUMERS = 256;    // consumer threads (8 warps)
constexpr int MV_THREADS = MV_CONSUMERS + 64;   // + TMA warp + X-staging warp
// L2 carry-over between alternating passes (tests/gpu_l2_sweep.py, N = 16384, k = 8): alternating the direction gives
// 167.9 us instead of 173.2 us per pass; the evict-last tail adds nothing measurable beyond that (167.0 us at 32 MB)
// because hits and misses share the same ~6.4 TB/s L2-slice throughput cap.
constexpr int MV_L2_KEEP_MB = 32;

// Work decomposition of one block matvec launch (identical for both kernels, so that the
// per-tile partial dot products have one layout).
struct MvTiling {
  int tile_rows;        // TH: 8 .. 128 rows (any integer)
  int tiles_per_batch;
  int ntiles;
  int grid;
};
MvTiling mv_tiling(int nbatch, int nrows, int reserve_sms = 0);

// Y_b = A_b X_b [- Z_b diag(E_b)], plus optional fused per-tile partial dot products
//   dot_out[tile][0][c] = sum_rows U[row][c] * Y[row][c]      (if U != nullptr)
//   dot_out[tile][1][c] = sum_rows Y[row][c]^2
// stored as double, layout [ntiles][2][MV_MAXK].
struct MvArgs {
  int dtype;                                   // xt_dtype of A
  int nbatch, nrows, ncolsA, k;                // k <= MV_MAXK
  const void* A; int64_t lda, a_bstride;
  const void* X; int64_t ldx, x_bstride;
  void* Y; int64_t ldy, y_bstride;
  const void* E; int64_t e_bstride;            // optional shift (nbatch, k)
  const void* Z; int64_t ldz, z_bstride;       // optional Z (defaults to X when E is given)
  const void* U; int64_t ldu, u_bstride;       // optional dot operand
  double* dot_out;                             // optional (needs U or self-dot), see above
  int impl;                                    // 0 auto, 1 TMA, 2 plain
  const int* done_flag;                        // optional device flag: kernel exits immediately when *done_flag != 0
                                               // (only ever written by earlier work of the SAME stream)
  const int* abort_flag;                       // optional device flag that may be raised ASYNCHRONOUSLY (another stream)
                                               // while the kernel runs: the TMA producer polls it every few chunks, stops
                                               // producing and the CTA drains and exits without storing its tile
  int* latch_out;                              // optional: set to 1 by a CTA that abandons its tile because of abort_flag
                                               // (a stream-ordered 'this pass is incomplete' mark for the kernels behind it)
  int reserve_sms;                             // leave this many SMs free (for kernels overlapped on another stream)
  int reverse;                                 // traverse A's column chunks last-to-first (alternate per call, see l2_keep_mb)
  int l2_keep_mb;                              // MB of the end of this pass to keep in L2 for the next, reversed pass
  int pdl;                                     // 1: launch as a programmatic dependent of the previous kernel in the stream
                                               // (row-slice TMA kernel with bulk X staging only; otherwise ignored): the
                                               // first A tiles are in flight before the predecessor has finished, X / flags
                                               // are only read after it has
};

// enqueue on `stream`; returns xt_status
int mv_launch(const MvArgs& a, cudaStream_t stream);

// Y_b = A_b^T X_b without a transposed copy (A: nrows x ncolsA, X: nrows x k, Y: ncolsA x k; fp32 / fp64, plain product)
int mv_launch_t(const MvArgs& a, cudaStream_t stream);

// true when the TMA kernel can take these arguments (alignment / stride rules)
bool mv_tma_ok(const MvArgs& a);

}  // namespace xt
