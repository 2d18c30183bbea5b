// Dense block matvec  Y = A X [- Z diag(E)]  for row-major A (fp32 / bf16 / fp64), k <= 16 columns,
// with optional fused per-tile dot products -- the kernel behind MatrixLinearOperator.mm
// (reference: xitorch/_core/linop.py:692-696 = torch.matmul) and behind every Krylov iteration.
//
// HBM-bound: A is read exactly once.  Design (B200 / sm_100a):
//   * one persistent CTA per SM, each owning tiles of TH <= 128 consecutive rows chosen so that
//     #tiles ~= #SMs (or >> #SMs) -- see mv_tiling()
//   * warp 0 (one elected lane) streams A through TMA: 3-D tensor map (cols, rows, batch), boxes of
//     tile_rows x 128 B with SWIZZLE_128B (two per stage), L2 evict-first hint, multi-stage mbarrier ring
//   * warp 1 stages the matching chunk of X into the same stage (plain coalesced loads; arbitrary
//     strides, zero padding) -- this is the hook where solver prologues get fused
//   * warps 2..9 (256 threads) consume: thread <-> (row, k-slice); 16-byte conflict-free LDS of its
//     row (swizzle-aware), X chunk broadcast from shared memory, packed FFMA2 accumulation with
//     per-stage blocked summation; k-slices are reduced through shared memory at tile end
//   * epilogue per row: shift term, store Y, fused partial dots (warp-shuffle + smem reduction,
//     one deterministic partial per tile, fp64)
// A second, plain-load kernel (one warp per row) covers shapes the TMA path cannot take
// (row stride not a multiple of 16 B) and cross-checks the TMA kernel in the tests.
#include "matvec.cuh"

#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <vector>
#include <type_traits>

namespace xt {

// ============================================================================ host utilities
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---------------------------------------------------------------------------- launch accounting / profiling
// Process-wide diagnostic state (xt_profile_* of the C ABI): two atomic launch counters, always on, and -- only while
// bench.py has switched the in-situ timing on -- a mutex-protected list of CUDA event pairs around matvec launches.
struct ProfState {
  std::atomic<bool> on{false};
  std::mutex mu;                   // guards ev / used
  std::vector<cudaEvent_t> ev;     // pairs (begin, end); created once and reused across resets
  size_t used = 0;
  std::atomic<int64_t> launches{0}, mv_launches{0};
};
static ProfState g_prof;
void note_launch(int n) { g_prof.launches.fetch_add(n, std::memory_order_relaxed); }
bool prof_on() { return g_prof.on.load(std::memory_order_relaxed); }
static void prof_record(cudaStream_t st) {          // caller holds g_prof.mu
  if (g_prof.used == g_prof.ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof.ev.push_back(e);
  }
  cudaEventRecord(g_prof.ev[g_prof.used++], st);
}
void prof_mv_begin(cudaStream_t st) {
  g_prof.mv_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof.on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  prof_record(st);
}
void prof_mv_end(cudaStream_t st) {
  if (!g_prof.on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if ((g_prof.used & 1) == 0) return;
  prof_record(st);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

MvTiling mv_tiling(int nbatch, int nrows, int reserve_sms) {
  MvTiling t;
  const int64_t total = (int64_t)nbatch * nrows;
  int G = num_sms() - reserve_sms;
  if (G < 1) G = 1;
  // rows per tile so that one wave of G CTAs covers everything.  Any row count works: a stage holds two TMA boxes
  // of tile_rows x 128 B at fixed 16 KB offsets, so the swizzle atoms stay 1024-B aligned (N = 16384 on 148 SMs:
  // 111 rows -> 148 tiles; leaving 3 SMs free costs 113 rows -> 145 tiles, < 2 %).
  int64_t th = (total + G - 1) / G;
  if (th > MV_TILE_ROWS) th = MV_TILE_ROWS;
  if (th > nrows) th = nrows;
  if (th < MV_BOX_ROWS) th = MV_BOX_ROWS;
  // several batches: the per-batch round-up can push the tile count just past one wave
  while (th < MV_TILE_ROWS && nbatch > 1 && (int64_t)nbatch * ((nrows + th - 1) / th) > G &&
         (int64_t)nbatch * ((nrows + th - 1) / th) < 2 * (int64_t)G) ++th;
  if (const char* ov = getenv("XT_MV_TILE_ROWS")) {      // tuning knob for experiments (tests/gpu_tile_sweep.py)
    const int v = atoi(ov);
    if (v >= MV_BOX_ROWS && v <= MV_TILE_ROWS) th = v;
  }
  t.tile_rows = (int)th;
  t.tiles_per_batch = (nrows + t.tile_rows - 1) / t.tile_rows;
  t.ntiles = t.tiles_per_batch * nbatch;
  t.grid = t.ntiles < G ? t.ntiles : G;
  return t;
}

// ---------------------------------------------------------------------------- tensor map creation
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static size_t dtype_size(int dt) { return dt == XT_F32 ? 4 : (dt == XT_BF16 ? 2 : 8); }

bool mv_tma_ok(const MvArgs& a) {
  const size_t es = dtype_size(a.dtype);
  if (reinterpret_cast<uintptr_t>(a.A) % 16) return false;
  if ((a.lda * es) % 16) return false;
  if (a.nbatch > 1 && a.a_bstride != 0 && (a.a_bstride * es) % 16) return false;
  if (a.lda < a.ncolsA) return false;
  if (get_encode_fn() == nullptr) return false;
  return true;
}

static int make_tmap(const MvArgs& a, int box_rows, CUtensorMap* tm, bool* batched) {
  const size_t es = dtype_size(a.dtype);
  const bool b3 = (a.nbatch > 1 && a.a_bstride != 0);
  *batched = b3;
  CUtensorMapDataType dt = a.dtype == XT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : (a.dtype == XT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
  cuuint64_t dims[3] = {(cuuint64_t)a.ncolsA, (cuuint64_t)a.nrows, (cuuint64_t)(b3 ? a.nbatch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)(a.lda * es),
                           (cuuint64_t)((b3 ? a.a_bstride : (int64_t)a.nrows * a.lda) * es)};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(tm, dt, 3, const_cast<void*>(a.A), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (ncols=%d nrows=%d lda=%lld)", (int)r, a.ncolsA,
                   a.nrows, (long long)a.lda);
    return XT_ERR_CUDA;
  }
  return XT_OK;
}

// ============================================================================ device code
struct MvDev {
  int nbatch, nrows, ncolsA, kvalid;
  int tile_rows, tiles_per_batch, ntiles;
  int rows_pad;     // tile_rows rounded up to a multiple of 16
  int nstages;
  int a_batched;
  int x_bulk;       // X chunks are contiguous and aligned: the TMA lane stages them with cp.async.bulk (no X warp)
  const void* X; int64_t ldx, x_bstride;
  void* Y; int64_t ldy, y_bstride;
  const void* E; int64_t e_bstride;
  const void* Z; int64_t ldz, z_bstride;
  const void* U; int64_t ldu, u_bstride;
  double* dot_out;
  const int* done_flag;
  const int* abort_flag;   // may be raised by another stream while the kernel runs (see MvArgs)
  int* latch_out;          // set by a CTA that abandons its tile (see MvArgs)
  int reverse;      // traverse the column chunks of A from the last to the first
  int keep_from;    // chunks (in traversal order) >= keep_from are loaded with an L2 evict-last hint: the next,
                    // oppositely ordered pass finds the tail of this one in L2
  int pdl;          // launched as a programmatic dependent: see MvArgs.pdl
  int y_atomic;           // Y += (atomicAdd) instead of Y =: the two column halves of a split pass (see mv_launch)
  uint32_t box_stride;    // bytes between the TMA boxes of a stage (shared-memory slot of one box)
  uint32_t stage_stride;  // bytes between stages; the X chunk of a stage sits at nbx * box_stride
  int nbx;                // TMA boxes (128 bytes of every row each) per stage: 2, or 4 in the wide-chunk row-slice variant
  int dbg;          // tcgen05 kernel: XT_TC5_DBG bit mask that switches single roles off (timing experiments only)
};

template <typename TA> struct ElemTraits;
template <> struct ElemTraits<float> {
  static constexpr int EPV = 4;  // elements per 16-byte vector
  __device__ static __forceinline__ void unpack(const float4& v, float (&a)[4]) {
    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
  }
};
template <> struct ElemTraits<__nv_bfloat16> {
  static constexpr int EPV = 8;
  __device__ static __forceinline__ void unpack(const float4& v, float (&a)[8]) {
    const uint32_t w[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[2 * i] = __uint_as_float(w[i] << 16);
      a[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct ElemTraits<double> {
  static constexpr int EPV = 2;
  __device__ static __forceinline__ void unpack(const float4& v, double (&a)[2]) {
    a[0] = __hiloint2double(__float_as_int(v.y), __float_as_int(v.x));
    a[1] = __hiloint2double(__float_as_int(v.w), __float_as_int(v.z));
  }
};

// K values of one X row from shared memory (explicit ld.shared, widest loads; no type punning through pointers
// so that the row stays in registers)
template <int K> __device__ __forceinline__ void load_xrow(uint32_t addr, float (&x)[K]) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int i = 0; i < K / 4; ++i) {
      const float4 v = lds128(addr + 16 * i);
      x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
  } else if constexpr (K == 2) {
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[0]), "=f"(x[1]) : "r"(addr));
  } else {
    static_assert(K == 1, "unexpected X row size");
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[0]) : "r"(addr));
  }
}
template <int K> __device__ __forceinline__ void load_xrow(uint32_t addr, double (&x)[K]) {
  if constexpr (K % 2 == 0) {
#pragma unroll
    for (int i = 0; i < K / 2; ++i)
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x[2 * i]), "=d"(x[2 * i + 1]) : "r"(addr + 16 * i));
  } else {
    static_assert(K == 1, "unexpected X row size");
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x[0]) : "r"(addr));
  }
}

template <int K> __device__ __forceinline__ void fma_row(float a, const float (&x)[K], float (&acc)[K]) {
  if constexpr (K % 2 == 0) {
    const float2 aa = make_float2(a, a);
#pragma unroll
    for (int i = 0; i < K / 2; ++i) {
      float2 r = __ffma2_rn(aa, make_float2(x[2 * i], x[2 * i + 1]), make_float2(acc[2 * i], acc[2 * i + 1]));
      acc[2 * i] = r.x;
      acc[2 * i + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = fmaf(a, x[i], acc[i]);
  }
}
template <int K> __device__ __forceinline__ void fma_row(double a, const double (&x)[K], double (&acc)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = fma(a, x[i], acc[i]);
}

// One consumer thread's share of one stage: NV 16-byte vectors of each of its RP rows (precomputed swizzled offsets
// aoff[], rows `hoff` bytes apart) against the matching X rows (xbase + v * EPV*K*sizeof(TV)).  With RP = 2 every X
// row fetched from shared memory feeds two rows of A, which halves the X share of the LDS traffic (for K = 8 the
// one-row form is bound by the shared-memory pipe: 0.85 wavefronts per clock in profiles/r1i_full_*).
// Fully unrolled: all shared-memory loads of a vector are independent of the previous vector's FMAs.
template <typename TA, typename TV, int K, int NV, int RP>
__device__ __forceinline__ void consume_stage(uint32_t a_s, uint32_t xbase, const uint32_t (&aoff)[8], uint32_t hoff,
                                              TV (&acc)[RP][K]) {
  using Tr = ElemTraits<TA>;
  constexpr int EPV = Tr::EPV;
  TV loc[RP][K];
#pragma unroll
  for (int h = 0; h < RP; ++h)
#pragma unroll
    for (int i = 0; i < K; ++i) loc[h][i] = TV(0);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    TV a[RP][EPV];
#pragma unroll
    for (int h = 0; h < RP; ++h) {
      const float4 raw = lds128(a_s + aoff[v] + (uint32_t)h * hoff);
      Tr::unpack(raw, a[h]);
    }
#pragma unroll
    for (int j = 0; j < EPV; ++j) {
      TV x[K];
      load_xrow<K>(xbase + (uint32_t)((v * EPV + j) * K * (int)sizeof(TV)), x);
#pragma unroll
      for (int h = 0; h < RP; ++h) fma_row<K>(a[h][j], x, loc[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < RP; ++h)
#pragma unroll
    for (int i = 0; i < K; ++i) acc[h][i] += loc[h][i];
}

constexpr int MV_STAGE_A_BYTES = MV_TILE_ROWS * 256;   // 128 rows x 2 boxes x 128 B

// ---- roles shared by both consumer layouts -------------------------------------------------------------------
// TMA producer (one elected lane): two boxes of tile_rows x 128 B per stage, L2 evict-first
// `abort_at` (shared memory, kernels that support MvArgs.abort_flag; else nullptr): sequence number of the first chunk
// that is NOT produced.  The producer is the only thread that looks at the global flag, so the decision is one per CTA:
// everything before `*abort_at` is produced and consumed as usual (no TMA transfer is ever left in flight), every
// waiter gives up at `*abort_at`.
template <typename TA, typename TV, int K, int STAGE_BYTES>
__device__ __forceinline__ void mv_producer(const CUtensorMap* tmA, const MvDev& p, uint8_t* stage_base, uint64_t* full,
                                            uint64_t* empty, int NS, int nchunks, int* abort_at = nullptr,
                                            int npre = 0) {
  // npre: the A boxes of chunks 0 .. npre-1 (stages 0 .. npre-1, first phase) are already in flight with their byte
  // counts registered (mv_preissue): only the arrival and the X chunk are still due for them
  constexpr int BOXC = 128 / (int)sizeof(TA);
  const int KC = p.nbx * BOXC;
  const char* Xg = reinterpret_cast<const char*>(p.X);
  const uint64_t pol_first = l2_policy_evict_first();
  const uint64_t pol_keep = l2_policy_evict_last();
  int s = 0;
  uint32_t ph = 0;
  int seq = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_batch;
    const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
    const int bA = p.a_batched ? b : 0;
    for (int ch = 0; ch < nchunks; ++ch, ++seq) {
      const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
      const int nb = min(p.nbx, (p.ncolsA - kc + BOXC - 1) / BOXC);
      if (abort_at != nullptr && (seq & 7) == 0 && *reinterpret_cast<const volatile int*>(p.abort_flag) != 0) {
        for (int q = seq; q < npre; ++q) {        // pre-issued A boxes must land before the CTA may leave
          mbar_arrive_expect_tx(&full[q], 0u);
          mbar_wait(&full[q], 0u);
        }
        *reinterpret_cast<volatile int*>(abort_at) = seq;
        if (p.latch_out != nullptr) *reinterpret_cast<volatile int*>(p.latch_out) = 1;
        return;
      }
      const bool pre = seq < npre;
      if (!pre) mbar_wait(&empty[s], ph ^ 1);
      uint8_t* dst = stage_base + (size_t)s * p.stage_stride;
      // one box = tile_rows x 128 B (rows past the end of the matrix are zero-filled by the TMA unit)
      uint32_t xbytes = 0;
      if (p.x_bulk) {
        const int cols = min(KC, p.ncolsA - kc);
        xbytes = (uint32_t)(cols * K * (int)sizeof(TV));
      }
      mbar_arrive_expect_tx(&full[s], (pre ? 0u : (uint32_t)(nb * p.tile_rows * 128)) + xbytes);
      if (!pre) {
        for (int bx = 0; bx < nb; ++bx)
          tma_load_3d(dst + (size_t)bx * p.box_stride, tmA, &full[s], kc + bx * BOXC, row0, bA,
                      ch >= p.keep_from ? pol_keep : pol_first);
      }
      if (p.x_bulk)
        bulk_load_1d(dst + (uint32_t)p.nbx * p.box_stride,
                     Xg + ((int64_t)b * p.x_bstride + (int64_t)kc * K) * (int64_t)sizeof(TV), xbytes, &full[s]);
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  }
}

// Programmatic dependent launch: the A boxes of the CTA's first chunks, issued before the predecessor kernel has
// finished (A does not depend on it).  Byte counts are registered without an arrival; the producer completes the
// stages (arrival + X chunk) after pdl_wait().  Returns the number of chunks issued.
template <typename TA, typename TV, int K, int STAGE_BYTES>
__device__ __forceinline__ int mv_preissue(const CUtensorMap* tmA, const MvDev& p, uint8_t* stage_base, uint64_t* full,
                                           int NS, int nchunks) {
  constexpr int BOXC = 128 / (int)sizeof(TA);
  const int KC = p.nbx * BOXC;
  const int tile = blockIdx.x;
  if (tile >= p.ntiles) return 0;
  const uint64_t pol_first = l2_policy_evict_first();
  const uint64_t pol_keep = l2_policy_evict_last();
  const int b = tile / p.tiles_per_batch;
  const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
  const int bA = p.a_batched ? b : 0;
  const int npre = min(NS, nchunks);
  for (int ch = 0; ch < npre; ++ch) {
    const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
    const int nb = min(p.nbx, (p.ncolsA - kc + BOXC - 1) / BOXC);
    uint8_t* dst = stage_base + (size_t)ch * p.stage_stride;
    mbar_expect_tx(&full[ch], (uint32_t)(nb * p.tile_rows * 128));
    for (int bx = 0; bx < nb; ++bx)
      tma_load_3d(dst + (size_t)bx * p.box_stride, tmA, &full[ch], kc + bx * BOXC, row0, bA,
                  ch >= p.keep_from ? pol_keep : pol_first);
  }
  return npre;
}

// X staging warp: all loads of a chunk are issued back to back (KC*K/32 independent loads per lane) and one
// chunk ahead of the shared-memory slot becoming free, so their L2 latency overlaps the wait.
template <typename TA, typename TV, int K, int STAGE_BYTES, int NBX = 2>
__device__ __forceinline__ void mv_xstager(const MvDev& p, uint8_t* stage_base, uint64_t* full, uint64_t* empty, int NS,
                                           int nchunks, int lane, const int* abort_at = nullptr) {
  constexpr int BOXC = 128 / (int)sizeof(TA);
  constexpr int KC = NBX * BOXC;
  const TV* __restrict__ Xg = reinterpret_cast<const TV*>(p.X);
  constexpr int NPL = KC * K / 32;
  TV vals[NPL];
  auto load_chunk = [&](int tile, int ch) {
    const int b = tile / p.tiles_per_batch;
    const TV* Xb = Xg + (int64_t)b * p.x_bstride;
    const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int idx = lane + 32 * i;
      const int c = idx / K, v = idx - c * K;
      vals[i] = (kc + c < p.ncolsA && v < p.kvalid) ? Xb[(int64_t)(kc + c) * p.ldx + v] : TV(0);
    }
  };
  int s = 0;
  uint32_t ph = 0;
  int tile = blockIdx.x, ch = 0, seq = 0;
  if (tile < p.ntiles) load_chunk(tile, 0);
  while (tile < p.ntiles) {
    if (abort_at != nullptr) {
      if (!mbar_wait_abortable(&empty[s], ph ^ 1, abort_at, seq)) return;
    } else {
      mbar_wait(&empty[s], ph ^ 1);
    }
    TV* xs = reinterpret_cast<TV*>(stage_base + (size_t)s * p.stage_stride + (uint32_t)p.nbx * p.box_stride);
#pragma unroll
    for (int i = 0; i < NPL; ++i) xs[lane + 32 * i] = vals[i];
    __syncwarp();
    if (lane == 0) mbar_arrive(&full[s]);
    if (++s == NS) { s = 0; ph ^= 1; }
    ++seq;
    if (++ch == nchunks) { ch = 0; tile += gridDim.x; }
    if (tile < p.ntiles) load_chunk(tile, ch);
  }
}

// per-row epilogue shared by both consumer layouts: shift term, store Y, partial dot products
template <typename TV, int K>
__device__ __forceinline__ void row_epilogue(const MvDev& p, int b, int64_t row, TV (&y)[K], double (&d0)[K],
                                             double (&d1)[K]) {
  if (p.E != nullptr) {
    const TV* Eb = reinterpret_cast<const TV*>(p.E) + (int64_t)b * p.e_bstride;
    const TV* Zr = (p.Z != nullptr) ? reinterpret_cast<const TV*>(p.Z) + (int64_t)b * p.z_bstride + row * p.ldz
                                    : reinterpret_cast<const TV*>(p.X) + (int64_t)b * p.x_bstride + row * p.ldx;
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i < p.kvalid) y[i] -= Eb[i] * Zr[i];
  }
  TV* Yr = reinterpret_cast<TV*>(p.Y) + (int64_t)b * p.y_bstride + row * p.ldy;
  if (p.y_atomic) {
    // column-split pass: exactly two partial sums land on a zeroed Y, fl(a + b) either way round (deterministic)
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i < p.kvalid) atomicAdd(&Yr[i], y[i]);
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i < p.kvalid) Yr[i] = y[i];
  }
  if (p.dot_out != nullptr) {
    const TV* Ur = (p.U != nullptr) ? reinterpret_cast<const TV*>(p.U) + (int64_t)b * p.u_bstride + row * p.ldu
                                    : nullptr;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (i < p.kvalid) {
        d1[i] = (double)y[i] * (double)y[i];
        if (Ur != nullptr) d0[i] = (double)Ur[i] * (double)y[i];
      }
    }
  }
}

template <typename TA, typename TV, int K, int NC, int RP, int NBX = 2>
__global__ void __launch_bounds__(NC + 64, 1)
mv_tma_kernel(const __grid_constant__ CUtensorMap tmA, const MvDev p) {
  using Tr = ElemTraits<TA>;
  constexpr int EPV = Tr::EPV;
  constexpr int BOXC = 128 / (int)sizeof(TA);   // columns per box
  constexpr int KC = NBX * BOXC;                // columns per stage (NBX boxes of 128 bytes per row)
  constexpr int VPR = NBX * 8;                  // 16-byte vectors of a row per stage
  constexpr int XBYTES = KC * K * (int)sizeof(TV);
  constexpr int STAGE_BYTES = (MV_STAGE_A_BYTES + XBYTES + 1023) / 1024 * 1024;   // swizzle atoms need 1024-B aligned stages

  const bool pdl = p.pdl != 0;      // flags written by the predecessor may only be read after pdl_wait()
  if (!pdl && p.done_flag != nullptr && *p.done_flag != 0) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.nstages;
  uint8_t* stage_base = smem;
  TV* red = reinterpret_cast<TV*>(smem + (size_t)NS * p.stage_stride);              // [NC][K]
  double* dscr = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(red) + NC * K * sizeof(TV));  // [NC/32][2][K]
  uint64_t* full = reinterpret_cast<uint64_t*>(dscr + (NC / 32) * 2 * K);
  uint64_t* empty = full + 16;
  // first chunk (sequence number) that is not produced; INT_MAX: none (see mv_producer).  Lives in the dynamic region
  // (32 mbarrier words are reserved: full[0..15], empty[16..29] for NS <= 14, these two at 30 / 31) so that the kernel
  // has NO static shared memory and the full 227 KB can be requested as dynamic.
  int& abort_at_s = *reinterpret_cast<int*>(full + 30);
  int& skip_s = *reinterpret_cast<int*>(full + 31);
  int* abort_at = p.abort_flag != nullptr ? &abort_at_s : nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (p.ncolsA + KC - 1) / KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], p.x_bulk ? 1 : 2);     // TMA lane (expect_tx) [+ X-staging warp]
      mbar_init(&empty[s], NC / 32);   // one arrival per consumer warp
    }
    fence_mbar_init();
    prefetch_tmap(&tmA);
    // one thread decides for the whole CTA whether a flag that was already up at launch skips the pass
    if (!pdl && abort_at != nullptr)
      abort_at_s = (*reinterpret_cast<const volatile int*>(p.abort_flag) != 0) ? 0 : 0x7fffffff;
  }
  if (p.x_bulk) {   // X slots start as zeros: a ragged last chunk copies fewer bytes and must never expose NaN garbage
    for (int s = 0; s < NS; ++s) {
      uint32_t* xz = reinterpret_cast<uint32_t*>(stage_base + (size_t)s * p.stage_stride + (uint32_t)NBX * p.box_stride);
      for (int i = threadIdx.x; i < XBYTES / 4; i += blockDim.x) xz[i] = 0u;
    }
    fence_proxy_async();
  }
  __syncthreads();

  int npre = 0;
  if (pdl) {
    if (warp == 0 && lane == 0) npre = mv_preissue<TA, TV, K, STAGE_BYTES>(&tmA, p, stage_base, full, NS, nchunks);
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0) {
      skip_s = (p.done_flag != nullptr && *p.done_flag != 0) ? 1 : 0;
      if (abort_at != nullptr)
        abort_at_s = (*reinterpret_cast<const volatile int*>(p.abort_flag) != 0) ? 0 : 0x7fffffff;
    }
    __syncthreads();
    const bool skip = skip_s != 0 || (abort_at != nullptr && abort_at_s == 0);
    if (skip) {
      if (threadIdx.x == 0 && p.latch_out != nullptr && skip_s == 0) *reinterpret_cast<volatile int*>(p.latch_out) = 1;
      if (warp == 0 && lane == 0) {
        for (int q = 0; q < npre; ++q) {          // the pre-issued A boxes must land before the CTA may leave
          mbar_arrive_expect_tx(&full[q], 0u);
          mbar_wait(&full[q], 0u);
        }
      }
      return;
    }
  }

  if (warp == 0) {
    if (lane == 0 && (abort_at == nullptr || abort_at_s != 0))
      mv_producer<TA, TV, K, STAGE_BYTES>(&tmA, p, stage_base, full, empty, NS, nchunks, abort_at, npre);
    else if (lane == 0 && p.latch_out != nullptr)
      *reinterpret_cast<volatile int*>(p.latch_out) = 1;            // the flag was up at launch: nothing is produced
  } else if (warp == 1) {
    if (!p.x_bulk) mv_xstager<TA, TV, K, STAGE_BYTES, NBX>(p, stage_base, full, empty, NS, nchunks, lane, abort_at);
  } else {
    // ------------------------------------------------------------------ consumers
    // thread <-> (row group r, k-slice q).  A thread owns RP rows: r, r + rows_pad, ... (rows_pad = ceil(tile_rows /
    // RP) rounded up to a multiple of 16 for RP = 1, of 8 for RP = 2, so that the rows of a thread share one
    // swizzle phase and quarter-warps never straddle a k-slice); ksplit = largest power of two with
    // ksplit * rows_pad <= NC; threads with q >= ksplit idle.
    const int tc = threadIdx.x - 64;
    const int r = tc % p.rows_pad;
    const int q = tc / p.rows_pad;
    int ksplit = 16;
    while (ksplit * p.rows_pad > NC) ksplit >>= 1;
    const int nvec = VPR / ksplit;        // <= 8 (the launcher picks NBX accordingly)
    const int cw = warp - 2;
    const uint32_t a_row_off = (uint32_t)(r * 128);
    const uint32_t hoff = (uint32_t)(p.rows_pad * 128);
    const uint32_t sw = (uint32_t)(r & 7);
    // per-thread constants: swizzled shared-memory offsets of its nvec vectors inside a stage, X offset
    uint32_t aoff[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int gv = q * nvec + (v < nvec ? v : 0);
      aoff[v] = (uint32_t)(gv >> 3) * p.box_stride + a_row_off + ((((uint32_t)gv & 7u) ^ sw) << 4);
    }
    const uint32_t x_q_off = (uint32_t)(q * nvec * EPV * K * (int)sizeof(TV));
    int s = 0;
    uint32_t ph = 0;
    int seq = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_batch;
      const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
      const int rows = min(p.tile_rows, p.nrows - row0);
      const bool active = (r < rows) && (q < ksplit);      // row r of a thread is its lowest: none active if r is not
      TV acc[RP][K];
#pragma unroll
      for (int h = 0; h < RP; ++h)
#pragma unroll
        for (int i = 0; i < K; ++i) acc[h][i] = TV(0);

      for (int ch = 0; ch < nchunks; ++ch, ++seq) {
        const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
        if (abort_at != nullptr) {
          // every consumer warp gives up at the same chunk (the first one the producer did not issue): the tile is
          // dropped, no barrier below is entered by anyone
          if (!mbar_wait_abortable(&full[s], ph, abort_at, seq)) return;
        } else {
          mbar_wait(&full[s], ph);
        }
        if (active) {
          const uint32_t a_s = smem_u32(stage_base + (size_t)s * p.stage_stride);
          const uint32_t xs = a_s + (uint32_t)NBX * p.box_stride + x_q_off;
          if (kc + KC <= p.ncolsA) {
            // full chunk: every vector of this thread is in bounds (rows past the tile read stale, finite-or-not
            // shared memory into accumulators that are never stored)
            if (nvec == 4) consume_stage<TA, TV, K, 4, RP>(a_s, xs, aoff, hoff, acc);
            else if (nvec == 8) consume_stage<TA, TV, K, 8, RP>(a_s, xs, aoff, hoff, acc);
            else if (nvec == 2) consume_stage<TA, TV, K, 2, RP>(a_s, xs, aoff, hoff, acc);
            else consume_stage<TA, TV, K, 1, RP>(a_s, xs, aoff, hoff, acc);
          } else {
            // ragged last chunk: skip vectors whose box was not loaded (columns past the end are zero-filled)
            for (int v = 0; v < nvec; ++v) {
              const int gv = q * nvec + v;
              if (kc + (gv >> 3) * BOXC < p.ncolsA) {
                const uint32_t off = (uint32_t)(gv >> 3) * p.box_stride + a_row_off +
                                     ((((uint32_t)gv & 7u) ^ sw) << 4);
                TV a[RP][EPV];
#pragma unroll
                for (int h = 0; h < RP; ++h) {
                  const float4 raw = lds128(a_s + off + (uint32_t)h * hoff);
                  Tr::unpack(raw, a[h]);
                }
#pragma unroll
                for (int j = 0; j < EPV; ++j) {
                  TV x[K];
                  load_xrow<K>(xs + (uint32_t)((v * EPV + j) * K * (int)sizeof(TV)), x);
#pragma unroll
                  for (int h = 0; h < RP; ++h) fma_row<K>(a[h][j], x, acc[h]);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NS) { s = 0; ph ^= 1; }
      }

      // ---- reduce the k-slices (one round per owned row through red[NC][K]), then the per-row epilogue (q == 0)
      double d0[K], d1[K];
#pragma unroll
      for (int i = 0; i < K; ++i) { d0[i] = 0.0; d1[i] = 0.0; }
#pragma unroll
      for (int h = 0; h < RP; ++h) {
        if (h > 0) named_bar_sync(1, NC);        // red[] is reused
        if (q > 0 && q < ksplit) {
#pragma unroll
          for (int i = 0; i < K; ++i) red[(size_t)tc * K + i] = acc[h][i];
        }
        named_bar_sync(1, NC);
        const int rr = r + h * p.rows_pad;
        if (q == 0 && rr < rows) {
          for (int qq = 1; qq < ksplit; ++qq) {
#pragma unroll
            for (int i = 0; i < K; ++i) acc[h][i] += red[(size_t)(qq * p.rows_pad + r) * K + i];
          }
          double e0[K], e1[K];
#pragma unroll
          for (int i = 0; i < K; ++i) { e0[i] = 0.0; e1[i] = 0.0; }
          row_epilogue<TV, K>(p, b, (int64_t)row0 + rr, acc[h], e0, e1);
#pragma unroll
          for (int i = 0; i < K; ++i) { d0[i] += e0[i]; d1[i] += e1[i]; }
        }
      }
      if (p.dot_out != nullptr) {
        // rows of this tile live in the q == 0 threads: tc < rows_pad  <=> the first consumer warps (a warp shared
        // with q = 1 threads only adds their zeros)
#pragma unroll
        for (int i = 0; i < K; ++i) {
          d0[i] = warp_sum(d0[i]);
          d1[i] = warp_sum(d1[i]);
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < K; ++i) {
            dscr[(cw * 2 + 0) * K + i] = d0[i];
            dscr[(cw * 2 + 1) * K + i] = d1[i];
          }
        }
      }
      named_bar_sync(1, NC);   // red[] / dscr[] hand-over
      if (p.dot_out != nullptr && tc < 2 * K) {
        const int which = tc / K, i = tc - which * K;
        double sum = 0.0;
        for (int w = 0; w < NC / 32; ++w) sum += dscr[(w * 2 + which) * K + i];
        p.dot_out[((size_t)tile * 2 + which) * MV_MAXK + i] = sum;
      }
      // the next tile's first red[]/dscr[] writes happen after its whole K sweep and a barrier: no hazard with
      // the reads above except dscr (written after the next first barrier) -> safe as well.
    }
  }
}

// ---------------------------------------------------------------------------- column-slice layout (wide blocks)
// For K >= 8 the row-slice layout above is bound by shared-memory traffic: every 16-byte vector of A needs
// K more LDS.128 for its X rows.  Here a consumer WARP owns a column slice (2 vectors = 8 columns of every
// stage) for ALL rows of the tile and each LANE a row (up to 4 row passes): the slice's X values live in
// registers for the whole stage, A is read with one conflict-free LDS.128 per (row, vector), and the 8 slices
// are summed through shared memory at tile end.  LDS per FFMA2 drops from 9:16 to 3:16.
template <int K>
__global__ void __launch_bounds__(256 + 64, 1)
mv_tma_colslice_kernel(const __grid_constant__ CUtensorMap tmA, const MvDev p) {
  using TA = float;
  using TV = float;
  constexpr int NC = 256, NW = NC / 32, PASSES = 4;
  constexpr int BOXC = 32, KC = 64, EPV = 4;
  constexpr int XBYTES = KC * K * (int)sizeof(TV);
  constexpr int STAGE_BYTES = (MV_STAGE_A_BYTES + XBYTES + 1023) / 1024 * 1024;
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.nstages;
  uint8_t* stage_base = smem;
  TV* red = reinterpret_cast<TV*>(smem + (size_t)NS * STAGE_BYTES);                       // [NW][MV_TILE_ROWS][K]
  double* dscr = reinterpret_cast<double*>(red + (size_t)NW * MV_TILE_ROWS * K);          // [NW][2][K]
  uint64_t* full = reinterpret_cast<uint64_t*>(dscr + NW * 2 * K);
  uint64_t* empty = full + NS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (p.ncolsA + KC - 1) / KC;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], p.x_bulk ? 1 : 2);
      mbar_init(&empty[s], NW);
    }
    fence_mbar_init();
    prefetch_tmap(&tmA);
  }
  if (p.x_bulk) {
    for (int s = 0; s < NS; ++s) {
      uint32_t* xz = reinterpret_cast<uint32_t*>(stage_base + (size_t)s * STAGE_BYTES + MV_STAGE_A_BYTES);
      for (int i = threadIdx.x; i < XBYTES / 4; i += blockDim.x) xz[i] = 0u;
    }
    fence_proxy_async();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) mv_producer<TA, TV, K, STAGE_BYTES>(&tmA, p, stage_base, full, empty, NS, nchunks);
  } else if (warp == 1) {
    if (!p.x_bulk) mv_xstager<TA, TV, K, STAGE_BYTES>(p, stage_base, full, empty, NS, nchunks, lane);
  } else {
    const int tc = threadIdx.x - 64;
    const int cw = warp - 2;
    const int npass = (p.tile_rows + 31) / 32;
    // vectors 2cw, 2cw+1 of the 16 per stage row: box, swizzled 16-byte slot for this lane's rows (r & 7 == lane & 7)
    uint32_t abase[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gv = 2 * cw + h;
      abase[h] = (uint32_t)((gv >> 3) * (MV_TILE_ROWS * 128)) + (uint32_t)(lane * 128) +
                 ((((uint32_t)gv & 7u) ^ ((uint32_t)lane & 7u)) << 4);
    }
    const uint32_t xoff = (uint32_t)(2 * cw * EPV * K * (int)sizeof(TV));
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_batch;
      const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
      const int rows = min(p.tile_rows, p.nrows - row0);
      // two-level (blocked) summation where the register budget allows it (K = 8); single level for K = 16
      constexpr bool TWO = (K <= 8);
      constexpr int LP = TWO ? PASSES : 1;
      TV acc[PASSES][K], loc2[LP][K];
#pragma unroll
      for (int q = 0; q < PASSES; ++q)
#pragma unroll
        for (int i = 0; i < K; ++i) acc[q][i] = 0.f;
#pragma unroll
      for (int q = 0; q < LP; ++q)
#pragma unroll
        for (int i = 0; i < K; ++i) loc2[q][i] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
        mbar_wait(&full[s], ph);
        const uint32_t a_s = smem_u32(stage_base + (size_t)s * STAGE_BYTES);
        const uint32_t xs = a_s + MV_STAGE_A_BYTES + xoff;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int gv = 2 * cw + h;
          if (kc + (gv >> 3) * BOXC < p.ncolsA) {          // warp-uniform: this slice's box was loaded
            TV x[EPV][K];
#pragma unroll
            for (int j = 0; j < EPV; ++j) load_xrow<K>(xs + (uint32_t)((h * EPV + j) * K * (int)sizeof(TV)), x[j]);
#pragma unroll
            for (int q = 0; q < PASSES; ++q) {
              if (q < npass && q * 32 + lane < rows) {
                const float4 raw = lds128(a_s + abase[h] + (uint32_t)(q * 32 * 128));
                TV(&dst)[K] = TWO ? loc2[TWO ? q : 0] : acc[q];
                fma_row<K>(raw.x, x[0], dst);
                fma_row<K>(raw.y, x[1], dst);
                fma_row<K>(raw.z, x[2], dst);
                fma_row<K>(raw.w, x[3], dst);
              }
            }
          }
        }
        if (TWO && (ch & 15) == 15) {                       // blocked summation: fold every 16 stages
#pragma unroll
          for (int q = 0; q < LP; ++q)
#pragma unroll
            for (int i = 0; i < K; ++i) { acc[q][i] += loc2[q][i]; loc2[q][i] = 0.f; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NS) { s = 0; ph ^= 1; }
      }
      // ---- sum the NW column slices through shared memory, then the per-row epilogue
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        const int r = q * 32 + lane;
        if (q < npass && r < rows) {
#pragma unroll
          for (int i = 0; i < K; ++i)
            red[((size_t)cw * MV_TILE_ROWS + r) * K + i] = TWO ? acc[q][i] + loc2[TWO ? q : 0][i] : acc[q][i];
        }
      }
      named_bar_sync(1, NC);
      double d0[K], d1[K];
#pragma unroll
      for (int i = 0; i < K; ++i) { d0[i] = 0.0; d1[i] = 0.0; }
      if (tc < rows) {
        TV y[K];
#pragma unroll
        for (int i = 0; i < K; ++i) y[i] = 0.f;
        for (int w = 0; w < NW; ++w) {
#pragma unroll
          for (int i = 0; i < K; ++i) y[i] += red[((size_t)w * MV_TILE_ROWS + tc) * K + i];
        }
        row_epilogue<TV, K>(p, b, (int64_t)row0 + tc, y, d0, d1);
      }
      if (p.dot_out != nullptr) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          d0[i] = warp_sum(d0[i]);
          d1[i] = warp_sum(d1[i]);
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < K; ++i) {
            dscr[(cw * 2 + 0) * K + i] = d0[i];
            dscr[(cw * 2 + 1) * K + i] = d1[i];
          }
        }
      }
      named_bar_sync(1, NC);
      if (p.dot_out != nullptr && tc < 2 * K) {
        const int which = tc / K, i = tc - which * K;
        double sum = 0.0;
        for (int w = 0; w < NW; ++w) sum += dscr[(w * 2 + which) * K + i];
        p.dot_out[((size_t)tile * 2 + which) * MV_MAXK + i] = sum;
      }
    }
  }
}

// ---------------------------------------------------------------------------- tensor-core layout (fp32, 8 < k <= 16)
// For k = 16 the SIMT kernels above are bound by the FP32 pipe (16 FMAs per element of A: 4.3 TB/s measured).  Here the
// block product is what it genuinely is -- a tall-skinny GEMM -- and runs on the tensor cores (legacy mma.sync path;
// selectable with impl = 6, see launch_tma() for why it is not the default) with error-compensated
// TF32 (A = A_hi + A_lo, X = X_hi + X_lo, each part a TF32 number; A_hi X_hi in one accumulator, A_lo X_hi + A_hi X_lo
// in a second one; the dropped A_lo X_lo term is ~2^-22 relative).  Tensor-core accumulation rounds toward zero, so the
// accumulators are flushed into round-to-nearest fp32 sums once per stage (8 k-steps): the result has the accuracy of
// the blocked fp32 summation of the SIMT kernels.
//   * same TMA producer / stage ring as above (two SWIZZLE_128B boxes of tile_rows x 32 floats per stage)
//   * warp 1 stages X: it splits the 64 x 16 chunk into hi / lo parts and stores them in mma.m16n8k8 B-fragment order,
//     one conflict-free LDS.128 {hi0, hi1, lo0, lo1} per (k-step, column tile, lane) for the consumers
//   * 8 consumer warps, warp w owns rows 16w .. 16w+15 of the tile; per k-step 4 conflict-free LDS.32 of the swizzled A
//     tile (fragment rows g, g+8 x columns t, t+4), the hi/lo split in registers, 6 mma.sync.m16n8k8 (2 column tiles)
// Used for plain products (no shift, no fused dots).
constexpr int TC_XBYTES = 8 * 2 * 32 * 16;            // 8 KB of fragment-ordered X per stage
constexpr int TC_STAGE_BYTES = MV_STAGE_A_BYTES + TC_XBYTES;   // 40 KB, 1024-byte aligned

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256 + 64, 1)
mv_tma_tc_kernel(const __grid_constant__ CUtensorMap tmA, const MvDev p) {
  constexpr int BOXC = 32, KC = 64;
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.nstages;
  uint8_t* stage_base = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NS * TC_STAGE_BYTES);
  uint64_t* empty = full + NS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (p.ncolsA + KC - 1) / KC;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 2);          // TMA lane (expect_tx) + X-staging warp
      mbar_init(&empty[s], 8);         // one arrival per consumer warp
    }
    fence_mbar_init();
    prefetch_tmap(&tmA);
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      // TMA producer (A only; X goes through warp 1)
      const uint64_t pol_first = l2_policy_evict_first();
      const uint64_t pol_keep = l2_policy_evict_last();
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_batch;
        const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
        const int bA = p.a_batched ? b : 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
          const int nb = (kc + BOXC < p.ncolsA) ? 2 : 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* dst = stage_base + (size_t)s * TC_STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], (uint32_t)(nb * p.tile_rows * 128));
          for (int bx = 0; bx < nb; ++bx)
            tma_load_3d(dst + bx * (MV_TILE_ROWS * 128), &tmA, &full[s], kc + bx * BOXC, row0, bA,
                        ch >= p.keep_from ? pol_keep : pol_first);
          if (++s == NS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // X stager: 16 fragment slots per lane and stage: slot (ks, nt, lane) = {hi(b0), hi(b1), lo(b0), lo(b1)} with
    // b0 = X[kc + 8 ks + t][8 nt + g], b1 = X[kc + 8 ks + t + 4][8 nt + g], t = lane & 3, g = lane >> 2
    const float* __restrict__ Xg = reinterpret_cast<const float*>(p.X);
    const int t = lane & 3, g = lane >> 2;
    float v0[16], v1[16];
    auto load_chunk = [&](int tile, int ch) {
      const int b = tile / p.tiles_per_batch;
      const float* Xb = Xg + (int64_t)b * p.x_bstride;
      const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int ks = i >> 1, nt = i & 1;
        const int r0 = kc + 8 * ks + t, col = 8 * nt + g;
        v0[i] = (r0 < p.ncolsA && col < p.kvalid) ? Xb[(int64_t)r0 * p.ldx + col] : 0.f;
        v1[i] = (r0 + 4 < p.ncolsA && col < p.kvalid) ? Xb[(int64_t)(r0 + 4) * p.ldx + col] : 0.f;
      }
    };
    int s = 0;
    uint32_t ph = 0;
    int tile = blockIdx.x, ch = 0;
    if (tile < p.ntiles) load_chunk(tile, 0);
    while (tile < p.ntiles) {
      mbar_wait(&empty[s], ph ^ 1);
      uint4* xs = reinterpret_cast<uint4*>(stage_base + (size_t)s * TC_STAGE_BYTES + MV_STAGE_A_BYTES);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t h0 = f2tf32(v0[i]), h1 = f2tf32(v1[i]);
        const uint32_t l0 = f2tf32(v0[i] - __uint_as_float(h0)), l1 = f2tf32(v1[i] - __uint_as_float(h1));
        xs[i * 32 + lane] = make_uint4(h0, h1, l0, l1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      if (++s == NS) { s = 0; ph ^= 1; }
      if (++ch == nchunks) { ch = 0; tile += gridDim.x; }
      if (tile < p.ntiles) load_chunk(tile, ch);
    }
  } else {
    const int cw = warp - 2;                     // rows 16 cw .. 16 cw + 15 of the tile
    const int t = lane & 3, g = lane >> 2;
    const int ra = 16 * cw + g, rb = ra + 8;
    // byte offsets of the fragment rows inside a box; 16-byte chunk index is XORed with (row & 7) (SWIZZLE_128B)
    const uint32_t offa = (uint32_t)(ra * 128 + t * 4), offb = (uint32_t)(rb * 128 + t * 4);
    const uint32_t swa = (uint32_t)(ra & 7), swb = (uint32_t)(rb & 7);
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_batch;
      const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
      const int rows = min(p.tile_rows, p.nrows - row0);
      float y[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) y[n][i] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
        mbar_wait(&full[s], ph);
        const uint32_t a_s = smem_u32(stage_base + (size_t)s * TC_STAGE_BYTES);
        const uint32_t x_s = a_s + MV_STAGE_A_BYTES + (uint32_t)lane * 16;
        float mainacc[2][4], corr[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) { mainacc[n][i] = 0.f; corr[n][i] = 0.f; }
        const int nks = (kc + KC <= p.ncolsA) ? 8 : (p.ncolsA - kc + 7) / 8;    // ragged last chunk: loaded boxes only
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks < nks) {
            const uint32_t box = (uint32_t)(ks >> 2) * (MV_TILE_ROWS * 128);
            const uint32_t c0 = (uint32_t)(ks & 3) * 2;                          // 16-byte chunk of columns t, chunk + 1 of t + 4
            float af[4];
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(af[0]) : "r"(a_s + box + offa + (((c0) ^ swa) << 4)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(af[1]) : "r"(a_s + box + offb + (((c0) ^ swb) << 4)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(af[2]) : "r"(a_s + box + offa + (((c0 + 1) ^ swa) << 4)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(af[3]) : "r"(a_s + box + offb + (((c0 + 1) ^ swb) << 4)));
            uint32_t ahi[4], alo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ahi[i] = f2tf32(af[i]);
              alo[i] = f2tf32(af[i] - __uint_as_float(ahi[i]));
            }
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              uint32_t xh0, xh1, xl0, xl1;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(xh0), "=r"(xh1), "=r"(xl0), "=r"(xl1)
                           : "r"(x_s + (uint32_t)((ks * 2 + n) * 32 * 16)));
              mma_tf32(mainacc[n], ahi, xh0, xh1);
              mma_tf32(corr[n], alo, xh0, xh1);
              mma_tf32(corr[n], ahi, xl0, xl1);
            }
          }
        }
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) y[n][i] += mainacc[n][i] + corr[n][i];      // round-to-nearest, once per stage
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NS) { s = 0; ph ^= 1; }
      }
      // accumulator layout: y[n][0], y[n][1] -> row ra, columns 8n + 2t, +1;  y[n][2], y[n][3] -> row rb
      float* Yb = reinterpret_cast<float*>(p.Y) + (int64_t)b * p.y_bstride;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int col = 8 * n + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = h == 0 ? ra : rb;
          if (rr < rows) {
            float* yr = Yb + ((int64_t)row0 + rr) * p.ldy + col;
            if (col < p.kvalid) yr[0] = y[n][2 * h];
            if (col + 1 < p.kvalid) yr[1] = y[n][2 * h + 1];
          }
        }
      }
    }
  }
}

static int launch_tc(const MvArgs& a, const MvDev& dev0, const MvTiling& til, cudaStream_t st) {
  const size_t fixed = 2 * 8 * sizeof(uint64_t) + 1024 + 64;
  int ns = (int)((227 * 1024 - fixed) / TC_STAGE_BYTES);
  if (ns > 6) ns = 6;
  if (ns < 2) {
    set_last_error("matvec: not enough shared memory for 2 stages");
    return XT_ERR_INVALID;
  }
  const size_t smem = (size_t)ns * TC_STAGE_BYTES + fixed;
  MvDev dev = dev0;
  dev.nstages = ns;
  CUtensorMap tm;
  bool batched = false;
  int rc = make_tmap(a, til.tile_rows, &tm, &batched);
  if (rc != XT_OK) return rc;
  dev.a_batched = batched ? 1 : 0;
  dev.x_bulk = 0;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    XT_CUDA_OK(set_max_dyn_smem(mv_tma_tc_kernel));
    attr_once.mark();
  }
  prof_mv_begin(st);
  mv_tma_tc_kernel<<<til.grid, 256 + 64, smem, st>>>(tm, dev);
  prof_mv_end(st);
  XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

// ---------------------------------------------------------------------------- tcgen05 layout (fp32, k = 16)
// The genuine tall-skinny GEMM of the path (block Lanczos with neig = 16, BASELINE config 5) on the fifth-generation
// tensor cores.  The SIMT layouts are bound by FP32 issue at k = 16 (4.2 TB/s); here the products run as
// error-compensated TF32 (the 3xTF32 scheme) with operands AND accumulators in tensor memory:
//     A = A_hi + A_lo,  X = X_hi + X_lo   (hi = the 19 leading bits the tensor core reads, lo = the exact remainder)
//     Y ~= A_hi X_hi + (A_hi X_lo + A_lo X_hi)          (A_lo X_lo is below fp32 rounding)
//   * four conversion warps (thread = row) read the TMA tile (tile_rows x 32 floats per box, SWIZZLE_128B: 16 conflict-
//     free LDS.128 per row and stage) and write the RAW row (kind::tf32 ignores the 13 low mantissa bits: that is A_hi) and
//     A_lo = A - trunc(A) into TENSOR MEMORY with tcgen05.st -- the shared-memory stage is free again as soon as they
//     are through, the tensor core never reads A from shared memory;
//   * B = [X_hi | X_lo]^T (32 x 64 per chunk, K-major SW128) is written by one warp from the bulk-copied X chunk into a
//     small ring of its own;
//   * one thread issues, per 8 columns of A:  D1 (+)= A_hi [X_hi | X_lo]  (M128 N32 K8)  and  D2 (+)= A_lo X_hi  (M128
//     N16 K8), both with A from tensor memory;
//   * the tensor core accumulates in fp32 with truncation, so D1 / D2 only ever hold MV5_FLUSH chunks: four epilogue
//     warps (thread = row) drain them with tcgen05.ld into round-to-nearest fp32 running sums while the issuer fills the
//     other accumulator pair, and apply the common row epilogue (shift, store, partial dots) at tile end.
// Two rings: NS shared-memory stages (TMA -> conversion / B staging) and MV5_NT tensor-memory operand slots + B tiles
// (conversion / B staging -> MMA).  Roles (12 warps): 0 TMA producer, 1 TMEM allocation + MMA issue (one thread),
// 2 B staging, 4-7 conversion, 8-11 epilogue (warp % 4 selects the 32 TMEM lanes a warp may touch).
constexpr int MV5_XRAW = 64 * 16 * 4;                                  // raw X chunk: 64 k-rows x 16 columns fp32
constexpr int MV5_BTILE = 2 * 32 * 128;                                // [2 k-slabs][32 rows (hi | lo)][128 B]
constexpr int MV5_NT = 3;                                              // operand slots: 96 + 3 x 128 TMEM columns
constexpr int MV5_FLUSH = 2;                                           // chunks per accumulator window
constexpr int MV5_THREADS = 384;

__device__ __forceinline__ uint64_t mv5_smem_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B canonical layout: rows 128 B apart, 8-row groups 1024 B apart (SBO), LBO unused (= 1),
  // descriptor version 1 (Blackwell), layout type 2
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ constexpr uint32_t mv5_idesc(int n) {
  // D = F32, A = B = TF32, both K-major, M = 128
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mv5_mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc),
      "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mv5_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db),
      "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mv5_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mv5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mv5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define MV5_R16(a, o) "=r"(a[o + 0]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), \
                      "=r"(a[o + 6]), "=r"(a[o + 7]), "=r"(a[o + 8]), "=r"(a[o + 9]), "=r"(a[o + 10]), "=r"(a[o + 11]), \
                      "=r"(a[o + 12]), "=r"(a[o + 13]), "=r"(a[o + 14]), "=r"(a[o + 15])
#define MV5_W16(a, o) "r"(a[o + 0]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), \
                      "r"(a[o + 6]), "r"(a[o + 7]), "r"(a[o + 8]), "r"(a[o + 9]), "r"(a[o + 10]), "r"(a[o + 11]), \
                      "r"(a[o + 12]), "r"(a[o + 13]), "r"(a[o + 14]), "r"(a[o + 15])
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void mv5_ld16(uint32_t taddr, uint32_t (&v)[48], int o) {
  if (o == 0)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : MV5_R16(v, 0) : "r"(taddr) : "memory");
  else if (o == 16)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : MV5_R16(v, 16) : "r"(taddr) : "memory");
  else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : MV5_R16(v, 32) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mv5_st16(uint32_t taddr, const uint32_t (&v)[64], int o) {
#define MV5_ST(O)                                                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(taddr), MV5_W16(v, O) : "memory")
  if (o == 0) MV5_ST(0);
  else if (o == 16) MV5_ST(16);
  else if (o == 32) MV5_ST(32);
  else MV5_ST(48);
#undef MV5_ST
}

__global__ void __launch_bounds__(MV5_THREADS, 1)
mv_tma_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const MvDev p) {
  using TV = float;
  constexpr int K = 16, BOXC = 32, KC = 64, NT = MV5_NT;
#define MV5_WAIT(bar, par) mbar_wait(bar, par)
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.nstages;
  const uint32_t box_tx = (uint32_t)p.tile_rows * 128u;                 // bytes one TMA box delivers
  const uint32_t box_bytes = ((uint32_t)p.tile_rows + 7u) / 8u * 1024u; // its slot: whole 8-row swizzle atoms (1024-byte aligned)
  const uint32_t stage_bytes = 2u * box_bytes + MV5_XRAW;
  uint8_t* stage_base = smem;
  uint8_t* bring = smem + (size_t)NS * stage_bytes;                    // [NT][MV5_BTILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bring + (size_t)NT * MV5_BTILE);
  uint64_t* full = bars;                 // [NS]  TMA landed
  uint64_t* empty = full + 8;            // [NS]  stage read by the 4 conversion warps + the B warp
  uint64_t* tready = empty + 8;          // [NT]  operand slot written (4 conversion warps + B warp)
  uint64_t* tfree = tready + 4;          // [NT]  operand slot consumed (tcgen05.commit)
  uint64_t* accfull = tfree + 4;         // [2]
  uint64_t* accempty = accfull + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
  double* dscr = reinterpret_cast<double*>(tmem_slot + 4);      // [4 warps][2][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (p.ncolsA + KC - 1) / KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 5);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(&tready[t], 5);
      mbar_init(&tfree[t], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&accfull[b], 1);
      mbar_init(&accempty[b], 4);        // one arrival per epilogue warp
    }
    fence_mbar_init();
    prefetch_tmap(&tmA);
  }
  for (int s = 0; s < NS; ++s) {         // X slots start as zeros (a ragged last chunk copies fewer bytes)
    uint32_t* xz = reinterpret_cast<uint32_t*>(stage_base + (size_t)s * stage_bytes + 2u * box_bytes);
    for (int i = threadIdx.x; i < MV5_XRAW / 4; i += blockDim.x) xz[i] = 0u;
  }
  fence_proxy_async();
  if (warp == 1) {                       // 512 TMEM columns: 2 x (32 + 16) accumulators, then NT x (64 A_hi + 64 A_lo)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  mv5_fence_before();
  __syncthreads();
  mv5_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t tm_acc = tmem_base;                 // + buf * 48: D1 (32 columns), + 32: D2 (16 columns)
  const uint32_t tm_op = tmem_base + 96;             // + slot * 128: A_hi (64 columns), + 64: A_lo

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    if (lane == 0) {
      const char* Xg = reinterpret_cast<const char*>(p.X);
      const uint64_t pol_first = l2_policy_evict_first();
      const uint64_t pol_keep = l2_policy_evict_last();
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_batch;
        const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
        const int bA = p.a_batched ? b : 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
          const int nb = (kc + BOXC < p.ncolsA) ? 2 : 1;
          MV5_WAIT(&empty[s], ph ^ 1);
          uint8_t* dst = stage_base + (size_t)s * stage_bytes;
          const int cols = min(KC, p.ncolsA - kc);
          const uint32_t xbytes = (uint32_t)(cols * K * (int)sizeof(TV));
          mbar_arrive_expect_tx(&full[s], (uint32_t)nb * box_tx + xbytes);
          for (int bx = 0; bx < nb; ++bx)
            tma_load_3d(dst + (size_t)bx * box_bytes, &tmA, &full[s], kc + bx * BOXC, row0, bA,
                        ch >= p.keep_from ? pol_keep : pol_first);
          bulk_load_1d(dst + 2u * box_bytes, Xg + ((int64_t)b * p.x_bstride + (int64_t)kc * K) * (int64_t)sizeof(TV),
                       xbytes, &full[s]);
          if (++s == NS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t id32 = mv5_idesc(32), id16 = mv5_idesc(16);
      const int flush = ((p.dbg >> 8) & 0xff) ? ((p.dbg >> 8) & 0xff) : MV5_FLUSH;
      int t = 0;
      uint32_t pt = 0;
      uint32_t win = 0;                  // accumulator windows so far (pair = win & 1)
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int in_win = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          const uint32_t buf = win & 1u;
          if (in_win == 0) MV5_WAIT(&accempty[buf], ((win >> 1) & 1u) ^ 1u);      // the epilogue has drained this pair
          MV5_WAIT(&tready[t], pt);
          mv5_fence_after();
          const uint32_t b_s = smem_u32(bring + (size_t)t * MV5_BTILE);
          const uint32_t d1 = tm_acc + buf * 48u, d2 = d1 + 32u;
          const uint32_t ahi = tm_op + (uint32_t)t * 128u, alo = ahi + 64u;
          const int kc = (p.reverse ? nchunks - 1 - ch : ch) * KC;
          const int nks = min(8, (p.ncolsA - kc + 7) / 8);          // k-steps of this chunk (ragged last chunk)
#pragma unroll 1
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t boff = (uint32_t)(ks >> 2) * (32 * 128) + (uint32_t)(ks & 3) * 32u;
            const uint64_t db = mv5_smem_desc(b_s + boff);
            const uint32_t acc = (in_win > 0 || ks > 0) ? 1u : 0u;
            if (!(p.dbg & 2)) mv5_mma_ts(d1, ahi + (uint32_t)ks * 8u, db, id32, acc);      // A_hi [X_hi | X_lo]
            if (!(p.dbg & 1)) mv5_mma_ts(d2, alo + (uint32_t)ks * 8u, db, id16, acc);      // A_lo X_hi
          }
          mv5_commit(&tfree[t]);                                    // operand slot + B tile reusable once these MMAs are done
          ++in_win;
          if (in_win == flush || ch == nchunks - 1) {
            mv5_commit(&accfull[buf]);
            in_win = 0;
            ++win;
          }
          if (++t == NT) { t = 0; pt ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ B staging: X chunk -> [X_hi | X_lo]^T, K-major SW128
    int s = 0, t = 0;
    uint32_t ph = 0, pt = 0;
    const int n = lane & 15, kh = lane >> 4;               // column of X, parity of the 16-byte chunk
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int ch = 0; ch < nchunks; ++ch) {
        MV5_WAIT(&full[s], ph);
        MV5_WAIT(&tfree[t], pt ^ 1);
        const float* xr = reinterpret_cast<const float*>(stage_base + (size_t)s * stage_bytes + 2u * box_bytes);
        uint8_t* bt = bring + (size_t)t * MV5_BTILE;
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2) {
            const int c = 2 * c2 + kh;                     // chunk = 4 consecutive k-rows, 0..15 over the stage
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x = xr[(4 * c + q) * K + n];
              const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
              hi[q] = h;
              lo[q] = x - h;
            }
            const int slab = c >> 3, cc = c & 7;
            uint8_t* rowh = bt + slab * (32 * 128) + n * 128 + ((cc ^ (n & 7)) << 4);
            uint8_t* rowl = rowh + 16 * 128;               // rows 16..31 (same swizzle phase: (n + 16) & 7 == n & 7)
            *reinterpret_cast<float4*>(rowh) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(rowl) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tready[t]);
          mbar_arrive(&empty[s]);
        }
        if (++s == NS) { s = 0; ph ^= 1; }
        if (++t == NT) { t = 0; pt ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ conversion: thread = row, A_hi / A_lo -> TMEM
    const int r = (warp - 4) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp - 4) * 32) << 16;
    const uint32_t sw = (uint32_t)(r & 7);
    const bool live = r < p.tile_rows;                     // rows past the tile: their TMEM lanes may hold anything
    int s = 0, t = 0;
    uint32_t ph = 0, pt = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int ch = 0; ch < nchunks; ++ch) {
        MV5_WAIT(&full[s], ph);
        MV5_WAIT(&tfree[t], pt ^ 1);
        mv5_fence_after();
        const uint32_t a_s = smem_u32(stage_base + (size_t)s * stage_bytes) + (uint32_t)r * 128u;
        const uint32_t tdst = tm_op + (uint32_t)t * 128u + lane_addr;
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {           // one TMA box (32 columns of A) at a time
            uint32_t hi[64], lo[64];                       // only [0, 32) used per half (the helpers take 64-wide arrays)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (live) v = lds128(a_s + (uint32_t)half * box_bytes + ((((uint32_t)c) ^ sw) << 4));
              const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t raw = __float_as_uint(f[q]);
                hi[4 * c + q] = raw;                                                       // the tensor core truncates
                lo[4 * c + q] = __float_as_uint(f[q] - __uint_as_float(raw & 0xffffe000u));
              }
            }
            mv5_st16(tdst + (uint32_t)half * 32u, hi, 0);
            mv5_st16(tdst + (uint32_t)half * 32u + 16u, hi, 16);
            mv5_st16(tdst + 64u + (uint32_t)half * 32u, lo, 0);
            mv5_st16(tdst + 64u + (uint32_t)half * 32u + 16u, lo, 16);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        mv5_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tready[t]);
          mbar_arrive(&empty[s]);
        }
        if (++s == NS) { s = 0; ph ^= 1; }
        if (++t == NT) { t = 0; pt ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ epilogue: thread = row
    const int ew = warp - 8;
    const int r = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    uint32_t win = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_batch;
      const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
      const int rows = min(p.tile_rows, p.nrows - row0);
      float y[K];
#pragma unroll
      for (int i = 0; i < K; ++i) y[i] = 0.f;
      const int flush = ((p.dbg >> 8) & 0xff) ? ((p.dbg >> 8) & 0xff) : MV5_FLUSH;
      const int nwin = (nchunks + flush - 1) / flush;
      for (int w = 0; w < nwin; ++w, ++win) {
        const uint32_t buf = win & 1u;
        MV5_WAIT(&accfull[buf], (win >> 1) & 1u);
        mv5_fence_after();
        uint32_t v[48];
        const uint32_t t = tm_acc + buf * 48u + lane_addr;
        if (!(p.dbg & 16)) {
          mv5_ld16(t, v, 0);
          mv5_ld16(t + 16, v, 16);
          mv5_ld16(t + 32, v, 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int i = 0; i < 48; ++i) v[i] = 0u;
        }
        mv5_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&accempty[buf]);
#pragma unroll
        for (int i = 0; i < K; ++i)
          y[i] += __uint_as_float(v[i]) + (__uint_as_float(v[16 + i]) + __uint_as_float(v[32 + i]));   // round to nearest
      }
      double d0[K], d1[K];
#pragma unroll
      for (int i = 0; i < K; ++i) { d0[i] = 0.0; d1[i] = 0.0; }
      if (r < rows) row_epilogue<TV, K>(p, b, (int64_t)row0 + r, y, d0, d1);
      if (p.dot_out != nullptr) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          d0[i] = warp_sum(d0[i]);
          d1[i] = warp_sum(d1[i]);
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < K; ++i) {
            dscr[(ew * 2 + 0) * K + i] = d0[i];
            dscr[(ew * 2 + 1) * K + i] = d1[i];
          }
        }
        named_bar_sync(1, 128);
        const int tc = ew * 32 + lane;
        if (tc < 2 * K) {
          const int which = tc / K, i = tc - which * K;
          double sum = 0.0;
          for (int q = 0; q < 4; ++q) sum += dscr[(q * 2 + which) * K + i];
          p.dot_out[((size_t)tile * 2 + which) * MV_MAXK + i] = sum;
        }
        named_bar_sync(1, 128);
      }
    }
  }
  // ---- teardown: every role is done with tensor memory
  mv5_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    mv5_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

#undef MV5_WAIT

static int launch_tc5(const MvArgs& a, const MvDev& dev0, const MvTiling& til, cudaStream_t st) {
  const size_t fixed = (size_t)MV5_NT * MV5_BTILE + (size_t)(8 + 8 + 4 + 4 + 2 + 2) * sizeof(uint64_t) + 16 +
                       4 * 2 * 16 * sizeof(double) + 1024 + 64;
  const size_t stage_bytes = (size_t)((til.tile_rows + 7) / 8) * 2048 + MV5_XRAW;
  int ns = (int)((227 * 1024 - fixed) / stage_bytes);
  if (ns > 8) ns = 8;
  if (ns < 2) {
    set_last_error("matvec: tcgen05 layout: tile of %d rows does not fit", til.tile_rows);
    return XT_ERR_INVALID;
  }
  const size_t smem = (size_t)ns * stage_bytes + fixed;
  MvDev dev = dev0;
  dev.nstages = ns;
  CUtensorMap tm;
  bool batched = false;
  int rc = make_tmap(a, til.tile_rows, &tm, &batched);
  if (rc != XT_OK) return rc;
  dev.a_batched = batched ? 1 : 0;
  dev.x_bulk = 1;
  dev.dbg = getenv("XT_TC5_DBG") ? atoi(getenv("XT_TC5_DBG")) : 0;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    XT_CUDA_OK(set_max_dyn_smem(mv_tma_tc5_kernel));
    attr_once.mark();
  }
  prof_mv_begin(st);
  mv_tma_tc5_kernel<<<til.grid, MV5_THREADS, smem, st>>>(tm, dev);
  prof_mv_end(st);
  XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

// ---------------------------------------------------------------------------- transposed access  Y = A^T X
// rmm / rmv of a non-Hermitian dense operator and the A^H (A x) of the normal equations (reference linop.py:698-702,
// solve.py:637-643) WITHOUT materialising A^T: a CTA owns a strip of 2 TMA boxes = 64 (fp32) / 32 (fp64) columns of A and
// streams ALL rows through the same SWIZZLE_128B box ring as the forward kernels (box = 128 rows x 128 B), so A is
// read once, in full 128-byte row segments.  Consumers: a warp takes one box and every fourth row pair, a lane two
// adjacent columns (one LDS.64 of A per row -- a 128-byte row of the box is conflict-free under the swizzle -- plus
// K/4 broadcast LDS.128 of the X row) and keeps 2 x K accumulators; the row groups are summed through shared memory
// at the end of the strip.  X rows are staged by one warp with plain loads (any row stride, rows past the end as
// zeros; TMA zero-fills the rows of A past the end).
constexpr int MVT_ROWS = 128;                 // rows of A per stage
constexpr int MVT_NCW = 8;                    // consumer warps: 2 boxes x 4 row groups

template <typename TA, typename TV, int K>
__global__ void __launch_bounds__(MVT_NCW * 32 + 64, 1)
mv_tma_t_kernel(const __grid_constant__ CUtensorMap tmA, const MvDev p) {
  constexpr int BOXC = 128 / (int)sizeof(TA);           // columns per box
  constexpr int STRIP = 2 * BOXC;
  constexpr int CPL = BOXC / 16;                        // columns per lane: 2 (4-byte elements) or 1 (8-byte)
  constexpr int XBYTES = MVT_ROWS * K * (int)sizeof(TV);
  constexpr int STAGE_BYTES = (MV_STAGE_A_BYTES + XBYTES + 1023) / 1024 * 1024;
  static_assert(sizeof(TA) == sizeof(TV), "transposed kernel: A and the vectors share one element type");
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.nstages;
  uint8_t* stage_base = smem;
  TV* red = reinterpret_cast<TV*>(smem + (size_t)NS * STAGE_BYTES);          // [4 row groups][STRIP][K]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 4 * STRIP * K);
  uint64_t* empty = full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrchunks = (p.nrows + MVT_ROWS - 1) / MVT_ROWS;
  const int strips_per_batch = (p.ncolsA + STRIP - 1) / STRIP;
  const int nstrips = strips_per_batch * p.nbatch;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 2);                 // TMA lane (expect_tx) + X staging warp
      mbar_init(&empty[s], MVT_NCW);
    }
    fence_mbar_init();
    prefetch_tmap(&tmA);
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      int s = 0;
      uint32_t ph = 0;
      for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
        const int b = strip / strips_per_batch;
        const int c0 = (strip - b * strips_per_batch) * STRIP;
        const int bA = p.a_batched ? b : 0;
        const int nb = (c0 + BOXC < p.ncolsA) ? 2 : 1;
        for (int rc = 0; rc < nrchunks; ++rc) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* dst = stage_base + (size_t)s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], (uint32_t)(nb * MVT_ROWS * 128));
          for (int bx = 0; bx < nb; ++bx)
            tma_load_3d(dst + bx * (MV_TILE_ROWS * 128), &tmA, &full[s], c0 + bx * BOXC, rc * MVT_ROWS, bA, pol);
          if (++s == NS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // X staging: rows [rc * 128, +128) of X_b, K values each, zero rows past the end
    constexpr int NPL = MVT_ROWS * K / 32;
    int s = 0;
    uint32_t ph = 0;
    for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
      const int b = strip / strips_per_batch;
      const TV* Xb = reinterpret_cast<const TV*>(p.X) + (int64_t)b * p.x_bstride;
      for (int rc = 0; rc < nrchunks; ++rc) {
        TV vals[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          const int idx = lane + 32 * i;
          const int r = idx / K, v = idx - r * K;
          const int row = rc * MVT_ROWS + r;
          vals[i] = (row < p.nrows && v < p.kvalid) ? Xb[(int64_t)row * p.ldx + v] : TV(0);
        }
        mbar_wait(&empty[s], ph ^ 1);
        TV* xs = reinterpret_cast<TV*>(stage_base + (size_t)s * STAGE_BYTES + MV_STAGE_A_BYTES);
#pragma unroll
        for (int i = 0; i < NPL; ++i) xs[lane + 32 * i] = vals[i];
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int cw = warp - 2;                  // consumer warp 0..7
    const int box = cw & 1, rg = cw >> 1;     // its box of the stage, its row group
    const int half = lane >> 4, cl = lane & 15;
    // byte offset of this lane's columns inside a 128-byte box row: 16-byte chunk (cl * CPL * sizeof / 16), swizzled per row
    const uint32_t chunk = (uint32_t)(cl * CPL * (int)sizeof(TA)) >> 4;
    const uint32_t within = (uint32_t)(cl * CPL * (int)sizeof(TA)) & 15u;
    int s = 0;
    uint32_t ph = 0;
    for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
      const int b = strip / strips_per_batch;
      const int c0 = (strip - b * strips_per_batch) * STRIP;
      TV acc[CPL][K];
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int i = 0; i < K; ++i) acc[c][i] = TV(0);
      const bool box_live = c0 + box * BOXC < p.ncolsA;
      for (int rc = 0; rc < nrchunks; ++rc) {
        mbar_wait(&full[s], ph);
        if (box_live) {
          const uint32_t a_s = smem_u32(stage_base + (size_t)s * STAGE_BYTES) + (uint32_t)box * (MV_TILE_ROWS * 128);
          const uint32_t x_s = smem_u32(stage_base + (size_t)s * STAGE_BYTES) + MV_STAGE_A_BYTES;
#pragma unroll 4
          for (int j = 0; j < MVT_ROWS / 8; ++j) {
            const int r = 8 * j + 2 * rg + half;
            const uint32_t aaddr = a_s + (uint32_t)r * 128u + (((chunk ^ (uint32_t)(r & 7)) << 4) | within);
            TV av[CPL];
            if constexpr (sizeof(TA) == 4) {
              float2 t;
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"(aaddr));
              av[0] = t.x; av[1] = t.y;
            } else {
              asm volatile("ld.shared.f64 %0, [%1];" : "=d"(av[0]) : "r"(aaddr));
            }
            TV x[K];
            load_xrow<K>(x_s + (uint32_t)(r * K * (int)sizeof(TV)), x);
#pragma unroll
            for (int c = 0; c < CPL; ++c) fma_row<K>(av[c], x, acc[c]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NS) { s = 0; ph ^= 1; }
      }
      // the two row parities of a warp (lanes l and l + 16 hold the same columns), then the 4 row groups
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int i = 0; i < K; ++i) acc[c][i] += __shfl_xor_sync(0xffffffffu, acc[c][i], 16);
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < CPL; ++c)
#pragma unroll
          for (int i = 0; i < K; ++i) red[((size_t)rg * STRIP + box * BOXC + cl * CPL + c) * K + i] = acc[c][i];
      }
      named_bar_sync(1, MVT_NCW * 32);
      TV* Yb = reinterpret_cast<TV*>(p.Y) + (int64_t)b * p.y_bstride;
      for (int e = cw * 32 + lane; e < STRIP * K; e += MVT_NCW * 32) {
        const int c = e / K, i = e - c * K;
        if (c0 + c < p.ncolsA && i < p.kvalid) {
          const TV sum = (red[(size_t)(0 * STRIP + c) * K + i] + red[(size_t)(1 * STRIP + c) * K + i]) +
                         (red[(size_t)(2 * STRIP + c) * K + i] + red[(size_t)(3 * STRIP + c) * K + i]);
          Yb[(int64_t)(c0 + c) * p.ldy + i] = sum;
        }
      }
      named_bar_sync(1, MVT_NCW * 32);       // red[] is reused by the next strip
    }
  }
}

template <typename TA, typename TV, int K>
static int launch_tma_t_k(const MvArgs& a, const MvDev& dev0, cudaStream_t st) {
  constexpr int BOXC = 128 / (int)sizeof(TA);
  constexpr int STRIP = 2 * BOXC;
  constexpr int XBYTES = MVT_ROWS * K * (int)sizeof(TV);
  constexpr int STAGE_BYTES = (MV_STAGE_A_BYTES + XBYTES + 1023) / 1024 * 1024;
  const size_t fixed = (size_t)4 * STRIP * K * sizeof(TV) + 16 * sizeof(uint64_t) + 1024 + 64;
  int ns = (int)((227 * 1024 - fixed) / STAGE_BYTES);
  if (ns > 8) ns = 8;
  if (ns < 2) {
    set_last_error("matvec (transposed): not enough shared memory for 2 stages");
    return XT_ERR_INVALID;
  }
  const size_t smem = (size_t)ns * STAGE_BYTES + fixed;
  MvDev dev = dev0;
  dev.nstages = ns;
  CUtensorMap tm;
  bool batched = false;
  int rc = make_tmap(a, MVT_ROWS, &tm, &batched);
  if (rc != XT_OK) return rc;
  dev.a_batched = batched ? 1 : 0;
  auto kern = mv_tma_t_kernel<TA, TV, K>;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    XT_CUDA_OK(set_max_dyn_smem(kern));
    attr_once.mark();
  }
  const int nstrips = ((a.ncolsA + STRIP - 1) / STRIP) * a.nbatch;
  int G = num_sms() - a.reserve_sms;
  if (G < 1) G = 1;
  const int grid = nstrips < G ? nstrips : G;
  prof_mv_begin(st);
  kern<<<grid, MVT_NCW * 32 + 64, smem, st>>>(tm, dev);
  prof_mv_end(st);
  XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

// Y_b = A_b^T X_b  (A: nrows x ncolsA, X: nrows x k, Y: ncolsA x k); fp32 / fp64, TMA-able A only
int mv_launch_t(const MvArgs& a, cudaStream_t st) {
  XT_REQUIRE(a.k >= 1 && a.k <= MV_MAXK, "matvec^T: k=%d outside 1..%d", a.k, MV_MAXK);
  XT_REQUIRE(a.nbatch >= 1 && a.nrows >= 1 && a.ncolsA >= 1 && a.A && a.X && a.Y, "matvec^T: bad arguments");
  XT_REQUIRE(a.E == nullptr && a.dot_out == nullptr, "matvec^T: plain products only (no shift, no fused dots)");
  XT_REQUIRE(a.dtype == XT_F32 || a.dtype == XT_F64, "matvec^T: fp32 / fp64 operators only");
  XT_REQUIRE(mv_tma_ok(a), "matvec^T: A must be 16-byte aligned with a 16-byte multiple row stride");
  MvDev d;
  memset(&d, 0, sizeof(d));
  d.nbatch = a.nbatch; d.nrows = a.nrows; d.ncolsA = a.ncolsA; d.kvalid = a.k;
  d.X = a.X; d.ldx = a.ldx; d.x_bstride = a.x_bstride;
  d.Y = a.Y; d.ldy = a.ldy; d.y_bstride = a.y_bstride;
  d.done_flag = a.done_flag;
  if (a.dtype == XT_F32) {
    if (a.k <= 1) return launch_tma_t_k<float, float, 1>(a, d, st);
    if (a.k <= 2) return launch_tma_t_k<float, float, 2>(a, d, st);
    if (a.k <= 4) return launch_tma_t_k<float, float, 4>(a, d, st);
    if (a.k <= 8) return launch_tma_t_k<float, float, 8>(a, d, st);
    return launch_tma_t_k<float, float, 16>(a, d, st);
  }
  if (a.k <= 1) return launch_tma_t_k<double, double, 1>(a, d, st);
  if (a.k <= 2) return launch_tma_t_k<double, double, 2>(a, d, st);
  if (a.k <= 4) return launch_tma_t_k<double, double, 4>(a, d, st);
  if (a.k <= 8) return launch_tma_t_k<double, double, 8>(a, d, st);
  return launch_tma_t_k<double, double, 16>(a, d, st);
}

// ---------------------------------------------------------------------------- plain-load kernel
// one CTA per tile (same tiling => same dot layout), one warp per row, lanes stride the columns.
template <typename TA, typename TV>
__global__ void __launch_bounds__(256)
mv_plain_kernel(const TA* __restrict__ A, int64_t lda, int64_t a_bstride, const MvDev p) {
  if (p.done_flag != nullptr && *p.done_flag != 0) return;
  __shared__ double dscr[8][2][MV_MAXK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_batch;
    const int row0 = (tile - b * p.tiles_per_batch) * p.tile_rows;
    const int rows = min(p.tile_rows, p.nrows - row0);
    const TA* Ab = A + (int64_t)b * a_bstride;
    const TV* Xb = reinterpret_cast<const TV*>(p.X) + (int64_t)b * p.x_bstride;
    double d0[MV_MAXK], d1[MV_MAXK];
#pragma unroll
    for (int i = 0; i < MV_MAXK; ++i) { d0[i] = 0.0; d1[i] = 0.0; }
    for (int rr = warp; rr < rows; rr += 8) {
      const int64_t row = row0 + rr;
      const TA* Ar = Ab + row * lda;
      TV acc[MV_MAXK];
#pragma unroll
      for (int i = 0; i < MV_MAXK; ++i) acc[i] = TV(0);
      for (int c = lane; c < p.ncolsA; c += 32) {
        const TV a = (TV)Ar[c];
        const TV* xr = Xb + (int64_t)c * p.ldx;
#pragma unroll
        for (int i = 0; i < MV_MAXK; ++i)
          if (i < p.kvalid) acc[i] += a * xr[i];
      }
#pragma unroll
      for (int i = 0; i < MV_MAXK; ++i) acc[i] = warp_sum(acc[i]);
      if (lane == 0) {
        if (p.E != nullptr) {
          const TV* Eb = reinterpret_cast<const TV*>(p.E) + (int64_t)b * p.e_bstride;
          const TV* Zr = (p.Z != nullptr)
                             ? reinterpret_cast<const TV*>(p.Z) + (int64_t)b * p.z_bstride + row * p.ldz
                             : Xb + row * p.ldx;
          for (int i = 0; i < p.kvalid; ++i) acc[i] -= Eb[i] * Zr[i];
        }
        TV* Yr = reinterpret_cast<TV*>(p.Y) + (int64_t)b * p.y_bstride + row * p.ldy;
        const TV* Ur = (p.U != nullptr)
                           ? reinterpret_cast<const TV*>(p.U) + (int64_t)b * p.u_bstride + row * p.ldu
                           : nullptr;
#pragma unroll
        for (int i = 0; i < MV_MAXK; ++i) {
          if (i < p.kvalid) {
            Yr[i] = acc[i];
            d1[i] += (double)acc[i] * (double)acc[i];
            if (Ur != nullptr) d0[i] += (double)Ur[i] * (double)acc[i];
          }
        }
      }
    }
    if (p.dot_out != nullptr) {
      __syncthreads();
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < MV_MAXK; ++i) { dscr[warp][0][i] = d0[i]; dscr[warp][1][i] = d1[i]; }
      }
      __syncthreads();
      if (threadIdx.x < 2 * MV_MAXK) {
        const int which = threadIdx.x / MV_MAXK, i = threadIdx.x % MV_MAXK;
        double sum = 0.0;
        for (int w = 0; w < 8; ++w) sum += dscr[w][which][i];
        p.dot_out[((size_t)tile * 2 + which) * MV_MAXK + i] = sum;
      }
    }
  }
}

// ============================================================================ launch
// X can be staged by the TMA lane (cp.async.bulk) when every chunk is one contiguous, 16-byte aligned run
template <typename TV, int K> static bool x_bulk_ok(const MvArgs& a) {
  if (a.k != K || a.ldx != K) return false;
  if (reinterpret_cast<uintptr_t>(a.X) % 16) return false;
  if (((size_t)a.ncolsA * K * sizeof(TV)) % 16) return false;
  if (a.nbatch > 1 && ((size_t)a.x_bstride * sizeof(TV)) % 16) return false;
  return true;
}

template <typename TA, typename TV, int K, int NC, int RP, int NBX = 2>
static int launch_tma_k(const MvArgs& a, const MvDev& dev0, const MvTiling& til, cudaStream_t st) {
  constexpr int BOXC = 128 / (int)sizeof(TA);
  constexpr int KC = NBX * BOXC;
  constexpr int XBYTES = KC * K * (int)sizeof(TV);
  // compact stages: a TMA box of tile_rows rows occupies whole 8-row swizzle atoms only, so that short tiles (few rows
  // per SM: row-partitioned operators, small matrices) get MORE stages instead of half-empty ones -- the bytes in flight
  // per SM stay the same (8192 x 65536 fp32, 56-row tiles: 6 stages x 14 KB before)
  const uint32_t box_stride = (uint32_t)((til.tile_rows + 7) / 8) * 1024u;
  const uint32_t stage_stride = ((uint32_t)NBX * box_stride + (uint32_t)XBYTES + 1023u) / 1024u * 1024u;
  const size_t fixed = NC * K * sizeof(TV) + (NC / 32) * 2 * K * sizeof(double) + 32 * sizeof(uint64_t) + 1024 + 64;
  int ns = (int)((227 * 1024 - fixed) / stage_stride);
  const int ns_cap = getenv("XT_MV_MAXSTAGES") ? atoi(getenv("XT_MV_MAXSTAGES")) : 14;
  if (ns > ns_cap) ns = ns_cap;
  if (ns > 14) ns = 14;
  if (ns < 2) {
    set_last_error("matvec: not enough shared memory for 2 stages");
    return XT_ERR_INVALID;
  }
  const size_t smem = (size_t)ns * stage_stride + fixed;
  MvDev dev = dev0;
  dev.nstages = ns;
  dev.box_stride = box_stride;
  dev.stage_stride = stage_stride;
  dev.nbx = NBX;
  if (NBX != 2) dev.keep_from = dev0.keep_from * 2 / NBX;      // mv_launch counted chunks of 2 boxes
  CUtensorMap tm;
  bool batched = false;
  int rc = make_tmap(a, til.tile_rows, &tm, &batched);
  if (rc != XT_OK) return rc;
  dev.a_batched = batched ? 1 : 0;
  dev.x_bulk = x_bulk_ok<TV, K>(a) ? 1 : 0;
  dev.rows_pad = RP == 1 ? (til.tile_rows + 15) / 16 * 16 : ((til.tile_rows + RP - 1) / RP + 7) / 8 * 8;
  auto kern = mv_tma_kernel<TA, TV, K, NC, RP, NBX>;
  static DeviceOnce attr_once;   // per instantiation
  if (attr_once.pending()) {
    XT_CUDA_OK(set_max_dyn_smem(kern));
    attr_once.mark();
  }
  prof_mv_begin(st);
  static std::atomic<int> pdl_ok{1};
  bool launched = false;
  if (a.pdl && dev.x_bulk && pdl_ok.load(std::memory_order_relaxed)) {
    dev.pdl = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(til.grid); cfg.blockDim = dim3(NC + 64); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, tm, dev) == cudaSuccess) {
      launched = true;
    } else {
      (void)cudaGetLastError();
      pdl_ok.store(0);
      dev.pdl = 0;
    }
  }
  if (!launched) kern<<<til.grid, NC + 64, smem, st>>>(tm, dev);
  prof_mv_end(st);
  XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

template <int K>
static int launch_colslice(const MvArgs& a, const MvDev& dev0, const MvTiling& til, cudaStream_t st) {
  constexpr int XBYTES = 64 * K * 4;
  constexpr int STAGE_BYTES = (MV_STAGE_A_BYTES + XBYTES + 1023) / 1024 * 1024;
  const size_t fixed = (size_t)8 * MV_TILE_ROWS * K * 4 + 8 * 2 * K * sizeof(double) + 2 * 8 * sizeof(uint64_t) + 1024 + 64;
  int ns = (int)((227 * 1024 - fixed) / STAGE_BYTES);
  if (ns > 6) ns = 6;
  if (ns < 2) {
    set_last_error("matvec: not enough shared memory for 2 stages");
    return XT_ERR_INVALID;
  }
  const size_t smem = (size_t)ns * STAGE_BYTES + fixed;
  MvDev dev = dev0;
  dev.nstages = ns;
  dev.box_stride = MV_TILE_ROWS * 128;
  dev.stage_stride = STAGE_BYTES;
  CUtensorMap tm;
  bool batched = false;
  int rc = make_tmap(a, til.tile_rows, &tm, &batched);
  if (rc != XT_OK) return rc;
  dev.a_batched = batched ? 1 : 0;
  dev.x_bulk = x_bulk_ok<float, K>(a) ? 1 : 0;
  auto kern = mv_tma_colslice_kernel<K>;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    XT_CUDA_OK(set_max_dyn_smem(kern));
    attr_once.mark();
  }
  prof_mv_begin(st);
  kern<<<til.grid, 256 + 64, smem, st>>>(tm, dev);
  prof_mv_end(st);
  XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

template <typename TA, typename TV>
static int launch_tma(const MvArgs& a, const MvDev& dev, const MvTiling& til, cudaStream_t st) {
  // The column-slice layout (fp32, impl == 4) is kept as a cross-check only.  Measured on B200 (N = 16384):
  // k = 8: row-slice with two rows per thread 6260 GB/s, one row 6230 (5690 for tiles > 112 rows), column-slice 5850;
  // k = 16: row-slice with two rows per thread 4220 GB/s, column-slice 3890, one row 3680.
  if constexpr (std::is_same<TA, float>::value) {
    // impl == 6: tensor-core layout (3xTF32 through mma.sync).  Exact to fp32 rounding level but NOT the default:
    // measured on B200 (N = 16384, k = 16) it reaches 1.95 TB/s against 4.3 TB/s of the SIMT row-slice layout -- the
    // legacy mma.sync TF32 path sustains only ~43 clk per m16n8k8 per SM sub-partition (~53 TFLOP/s per GPU); the
    // tcgen05 / TMEM form of the same scheme is the round-2 item.
    const bool tc_ok = a.E == nullptr && a.dot_out == nullptr;
    if (a.impl == 6 && !tc_ok) {
      set_last_error("matvec: the tensor-core layout takes plain products only (no shift, no fused dots)");
      return XT_ERR_INVALID;
    }
    if (a.impl == 6) return launch_tc(a, dev, til, st);
    // impl == 7: tcgen05 / TMEM form of the same 3xTF32 scheme (k = 16, contiguous X); XT_MV_TC5=1 makes it the default
    // for such blocks
    {
      static const int tc5_default = (getenv("XT_MV_TC5") != nullptr && atoi(getenv("XT_MV_TC5")) != 0) ? 1 : 0;
      const bool tc5_ok = a.k == 16 && x_bulk_ok<float, 16>(a);
      if (a.impl == 7 && !tc5_ok) {
        set_last_error("matvec: the tcgen05 layout takes k = 16 with a contiguous, 16-byte aligned X");
        return XT_ERR_INVALID;
      }
      if (a.impl == 7 || (a.impl == 0 && tc5_default && tc5_ok)) return launch_tc5(a, dev, til, st);
    }
    if (a.impl == 4 && a.k > 4) {
      if (a.k <= 8) return launch_colslice<8>(a, dev, til, st);
      return launch_colslice<16>(a, dev, til, st);
    }
  }
  constexpr int NCW = std::is_same<TV, double>::value ? 256 : 512;   // fp64 accumulators need the larger register cap
  // 512 consumer threads (16 warps) where the register budget allows it, 256 for the widest blocks
  if (a.k <= 1) return launch_tma_k<TA, TV, 1, 512, 1>(a, dev, til, st);
  if (a.k <= 2) return launch_tma_k<TA, TV, 2, 512, 1>(a, dev, til, st);
  if (a.k <= 4) return launch_tma_k<TA, TV, 4, 512, 1>(a, dev, til, st);
  // k = 8: two rows per thread (impl == 5 keeps the one-row form for comparison)
  if (a.k <= 8) {
    if (a.impl == 5) return launch_tma_k<TA, TV, 8, NCW, 1>(a, dev, til, st);
    if constexpr (std::is_same<TA, float>::value) {
      // wide chunks (4 boxes = 512 bytes of every row per stage): half as many DRAM row activations per byte.  Needs
      // at most 8 vectors per thread and stage: ksplit >= 4, i.e. two-row groups of <= 128 consumer threads
      static const int nbx = getenv("XT_MV_NBX") ? atoi(getenv("XT_MV_NBX")) : 4;   // 4: measured +4 % (6.15 -> 6.39 TB/s in situ)
      const int rows_pad2 = ((til.tile_rows + 1) / 2 + 7) / 8 * 8;
      if (nbx == 4 && 4 * rows_pad2 <= NCW && a.ncolsA >= 1024) return launch_tma_k<TA, TV, 8, NCW, 2, 4>(a, dev, til, st);
    }
    return launch_tma_k<TA, TV, 8, NCW, 2>(a, dev, til, st);
  }
  if (a.impl == 5) return launch_tma_k<TA, TV, 16, 256, 1>(a, dev, til, st);
  return launch_tma_k<TA, TV, 16, 256, 2>(a, dev, til, st);
}

template <typename TA, typename TV>
static int launch_plain(const MvArgs& a, const MvDev& dev, const MvTiling& til, cudaStream_t st) {
  int grid = til.ntiles < 8 * num_sms() ? til.ntiles : 8 * num_sms();
  prof_mv_begin(st);
  mv_plain_kernel<TA, TV><<<grid, 256, 0, st>>>(reinterpret_cast<const TA*>(a.A), a.lda, a.a_bstride, dev);
  prof_mv_end(st);
  XT_LAUNCHED();
  XT_CUDA_OK(cudaGetLastError());
  return XT_OK;
}

int mv_launch(const MvArgs& a0, cudaStream_t st) {
  XT_REQUIRE(a0.k >= 1 && a0.k <= MV_MAXK, "matvec: k=%d outside 1..%d", a0.k, MV_MAXK);
  XT_REQUIRE(a0.nbatch >= 1 && a0.nrows >= 1 && a0.ncolsA >= 1, "matvec: empty problem (%d,%d,%d)", a0.nbatch, a0.nrows,
             a0.ncolsA);
  XT_REQUIRE(a0.A && a0.X && a0.Y, "matvec: null pointer");
  XT_REQUIRE(a0.E == nullptr || a0.Z != nullptr || a0.nrows == a0.ncolsA, "matvec: shift with Z = X needs a square A");
  MvArgs a = a0;
  int y_atomic = 0;
  {
    // Column split for wide, short operators (a row block of a row-partitioned matrix: 8192 x 65536 per GPU at 8 GPUs).
    // One wave of row tiles would be only ~56 rows high there, and the consumer layouts lose a quarter of their speed on
    // such flat tiles (3.2 instead of 4.1 TB/s at k = 16).  Instead the pass runs as TWO virtual batch items -- the left
    // and the right half of the columns (A: batch stride = half a row; X: its lower half) -- over tiles of twice the
    // height, both adding into a zeroed Y.
    const int64_t es = a.dtype == XT_F64 ? 8 : (a.dtype == XT_BF16 ? 2 : 4);
    const int64_t vs = a.dtype == XT_F64 ? 8 : 4;
    const int64_t kc_cols = 2 * (128 / es);
    const MvTiling t1 = mv_tiling(1, a.nrows, a.reserve_sms);
    const bool want = a.nbatch == 1 && a.impl == 0 && a.E == nullptr && a.dot_out == nullptr && a.abort_flag == nullptr &&
                      !a.pdl && t1.tile_rows <= 80 && a.nrows >= 8 * MV_BOX_ROWS && a.ncolsA >= 4096 &&
                      a.ncolsA % (2 * kc_cols) == 0 && mv_tma_ok(a) && getenv("XT_MV_NO_CSPLIT") == nullptr;
    if (want) {
      const int64_t half = a.ncolsA / 2;
      if (a.ldy == a.k) {
        XT_CUDA_OK(cudaMemsetAsync(a.Y, 0, (size_t)a.nrows * a.k * vs, st));
      } else {
        XT_CUDA_OK(cudaMemset2DAsync(a.Y, (size_t)a.ldy * vs, 0, (size_t)a.k * vs, (size_t)a.nrows, st));
      }
      a.nbatch = 2;
      a.ncolsA = (int)half;
      a.a_bstride = half;
      a.x_bstride = half * a.ldx;
      a.y_bstride = 0;
      y_atomic = 1;
    }
  }
  const MvTiling til = mv_tiling(a.nbatch, a.nrows, a.reserve_sms);
  MvDev d;
  d.nbatch = a.nbatch; d.nrows = a.nrows; d.ncolsA = a.ncolsA; d.kvalid = a.k;
  d.tile_rows = til.tile_rows; d.tiles_per_batch = til.tiles_per_batch; d.ntiles = til.ntiles;
  d.rows_pad = (til.tile_rows + 15) / 16 * 16;
  d.nstages = 0; d.a_batched = 0; d.x_bulk = 0;
  d.X = a.X; d.ldx = a.ldx; d.x_bstride = a.x_bstride;
  d.Y = a.Y; d.ldy = a.ldy; d.y_bstride = a.y_bstride;
  d.E = a.E; d.e_bstride = a.e_bstride;
  d.Z = a.Z; d.ldz = a.ldz; d.z_bstride = a.z_bstride;
  d.U = a.U; d.ldu = a.ldu; d.u_bstride = a.u_bstride;
  d.dot_out = a.dot_out;
  d.done_flag = a.done_flag;
  d.abort_flag = a.abort_flag;
  d.latch_out = a.latch_out;
  d.pdl = 0;
  d.dbg = 0;
  d.y_atomic = y_atomic;
  d.box_stride = MV_TILE_ROWS * 128;
  d.stage_stride = 0;          // set by the launcher
  d.nbx = 2;
  d.reverse = a.reverse ? 1 : 0;
  {
    // L2 carry-over between oppositely ordered passes (single-wave launches only): keep the last `keep` MB
    const int64_t es = a.dtype == XT_F64 ? 8 : (a.dtype == XT_BF16 ? 2 : 4);
    const int64_t kc_cols = 2 * (128 / es);
    const int nchunks = (int)((a.ncolsA + kc_cols - 1) / kc_cols);
    int64_t keep_mb = a.l2_keep_mb;
    if (const char* ov = getenv("XT_MV_L2_KEEP_MB")) keep_mb = atoi(ov);
    d.keep_from = nchunks;
    if (keep_mb > 0 && til.ntiles <= til.grid) {
      const int64_t chunk_bytes = (int64_t)a.nbatch * a.nrows * kc_cols * es;
      const int64_t kch = (keep_mb << 20) / (chunk_bytes > 0 ? chunk_bytes : 1);
      d.keep_from = kch >= nchunks ? 0 : nchunks - (int)kch;
    }
  }

  if (a.abort_flag != nullptr && (a.impl == 4 || a.impl == 6 || a.impl == 7 || a.impl == 2))
    d.abort_flag = nullptr;               // only the row-slice TMA kernel polls it; the others run to completion
  const bool forced_tma = (a.impl == 1 || a.impl == 3 || a.impl == 4 || a.impl == 5 || a.impl == 6 || a.impl == 7);
  bool use_tma = forced_tma || (a.impl == 0 && mv_tma_ok(a));
  if (forced_tma && !mv_tma_ok(a)) {
    set_last_error("matvec: TMA kernel forced but A is not 16-byte aligned / strided (lda=%lld)", (long long)a.lda);
    return XT_ERR_INVALID;
  }
  switch (a.dtype) {
    case XT_F32:
      return use_tma ? launch_tma<float, float>(a, d, til, st) : launch_plain<float, float>(a, d, til, st);
    case XT_BF16:
      return use_tma ? launch_tma<__nv_bfloat16, float>(a, d, til, st)
                     : launch_plain<__nv_bfloat16, float>(a, d, til, st);
    case XT_F64:
      return use_tma ? launch_tma<double, double>(a, d, til, st) : launch_plain<double, double>(a, d, til, st);
    default:
      set_last_error("matvec: unknown dtype %d", a.dtype);
      return XT_ERR_INVALID;
  }
}

}  // namespace xt

// ============================================================================ C ABI
extern "C" {

int xt_version(void) { return 100; }

void xt_profile_reset(int enable) {
  std::lock_guard<std::mutex> lk(xt::g_prof.mu);
  xt::g_prof.used = 0;
  xt::g_prof.on.store(enable != 0);
  xt::g_prof.launches.store(0);
  xt::g_prof.mv_launches.store(0);
}

int xt_profile_read(double* matvec_ms, int64_t* matvec_launches, int64_t* total_launches) {
  // launches made after a solver's device-side `done` flag was raised return immediately (or abort part-way): they
  // are not counted as matvecs (duration below 20 % of the longest one) but their time is reported in the total
  std::lock_guard<std::mutex> lk(xt::g_prof.mu);
  double ms = 0.0;
  int64_t eff = xt::g_prof.mv_launches.load();
  const size_t np = xt::g_prof.used / 2;
  if (np > 0) {
    XT_CUDA_OK(cudaEventSynchronize(xt::g_prof.ev[2 * np - 1]));
    std::vector<float> t(np, 0.f);
    float tmax = 0.f;
    for (size_t i = 0; i < np; ++i) {
      XT_CUDA_OK(cudaEventElapsedTime(&t[i], xt::g_prof.ev[2 * i], xt::g_prof.ev[2 * i + 1]));
      tmax = t[i] > tmax ? t[i] : tmax;
    }
    eff = 0;
    for (size_t i = 0; i < np; ++i)
      if (t[i] >= 0.2f * tmax) { ms += t[i]; ++eff; }
  }
  if (matvec_ms) *matvec_ms = ms;
  if (matvec_launches) *matvec_launches = eff;
  if (total_launches) *total_launches = xt::g_prof.launches.load();
  return XT_OK;
}
const char* xt_last_error(void) { return xt::last_error(); }

int xt_block_matvec(const xt_matvec_args* g) {
  if (g == nullptr) return XT_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(g->stream);
  const size_t vs = (g->dtype == XT_F64) ? 8 : 4;
  if (g->trans) {
    // Y = A^T X without materialising the transpose (rmm / rmv, normal equations)
    for (int c0 = 0; c0 < g->k; c0 += xt::MV_MAXK) {
      xt::MvArgs a;
      memset(&a, 0, sizeof(a));
      a.dtype = g->dtype;
      a.nbatch = g->nbatch; a.nrows = g->nrows; a.ncolsA = g->ncolsA;
      a.k = (g->k - c0 < xt::MV_MAXK) ? (g->k - c0) : xt::MV_MAXK;
      a.A = g->A; a.lda = g->lda; a.a_bstride = g->a_bstride;
      a.X = static_cast<const char*>(g->X) + c0 * vs; a.ldx = g->ldx; a.x_bstride = g->x_bstride;
      a.Y = static_cast<char*>(g->Y) + c0 * vs; a.ldy = g->ldy; a.y_bstride = g->y_bstride;
      if (g->E != nullptr) {
        xt::set_last_error("matvec: the transposed pass takes no shift");
        return XT_ERR_INVALID;
      }
      int rc = xt::mv_launch_t(a, st);
      if (rc != XT_OK) return rc;
    }
    return XT_OK;
  }
  // column groups of <= 16: one pass over A per group
  for (int c0 = 0; c0 < g->k; c0 += xt::MV_MAXK) {
    xt::MvArgs a;
    memset(&a, 0, sizeof(a));
    a.dtype = g->dtype;
    a.nbatch = g->nbatch; a.nrows = g->nrows; a.ncolsA = g->ncolsA;
    a.k = (g->k - c0 < xt::MV_MAXK) ? (g->k - c0) : xt::MV_MAXK;
    a.A = g->A; a.lda = g->lda; a.a_bstride = g->a_bstride;
    a.X = static_cast<const char*>(g->X) + c0 * vs; a.ldx = g->ldx; a.x_bstride = g->x_bstride;
    a.Y = static_cast<char*>(g->Y) + c0 * vs; a.ldy = g->ldy; a.y_bstride = g->y_bstride;
    a.E = g->E ? static_cast<const char*>(g->E) + c0 * vs : nullptr; a.e_bstride = g->e_bstride;
    a.Z = g->Z ? static_cast<const char*>(g->Z) + c0 * vs : nullptr; a.ldz = g->ldz; a.z_bstride = g->z_bstride;
    a.U = nullptr; a.ldu = 0; a.u_bstride = 0; a.dot_out = nullptr;
    a.impl = g->impl & 0xff;
    a.done_flag = nullptr;
    a.abort_flag = nullptr;
    a.reserve_sms = 0;
    a.reverse = (g->impl >> 8) & 1;          // bit 8: last-to-first column traversal
    a.l2_keep_mb = (g->impl >> 16) & 0xff;   // bits 16..23: MB of the pass's tail to keep in L2 (see matvec.cuh)
    int rc = xt::mv_launch(a, st);
    if (rc != XT_OK) return rc;
  }
  return XT_OK;
}

}  // extern "C"
