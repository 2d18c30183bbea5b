// (all state is `inline`: the three engine sources are separate translation units of ONE library and must share the
// emulated thread / CTA context)
// Host stand-ins for the CUDA constructs csrc/symeig.cu uses, so that the WHOLE eigensolver engine -- its host loop
// (run_symeig: kernel sequencing, lagged Ritz checks, thick restart, run-ahead window, operator callback) and every
// kernel -- compiles as plain C++ and runs on CPU memory (TEST INFRASTRUCTURE; built by tests/test_engine_emulation.py
// from the preprocessed .cu text, never shipped).  Execution model: a kernel launch runs synchronously, CTA after CTA,
// one host thread per CUDA thread (std::barrier = __syncthreads / __syncwarp); the one cooperative launch runs all
// its CTAs concurrently.  Streams and events are no-ops (everything has completed when a launch returns), "device"
// pointers are host pointers.  The block matvec (TMA kernel) is replaced by a plain loop: it is not what is tested here.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <semaphore>
#include <string>
#include <thread>
#include <vector>

#include "xitorch_b200.h"

#define __device__
#define __global__
#define __host__
#define __noinline__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x)

struct double2 { double x, y; };
struct float4 { float x, y, z, w; };
inline double2 make_double2(double x, double y) { return {x, y}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct EmuIdx { unsigned x, y, z; };
inline thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
constexpr int cudaSuccess = 0;
constexpr int cudaErrorNotReady = 600;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

struct EmuCta {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<std::barrier<>>> wbar;
  std::vector<unsigned char> stat, dyn;
  std::vector<double> wscr;        // warp exchange scratch [nwarps][32]
};
inline thread_local EmuCta* t_cta;

inline void __syncthreads() { t_cta->bar->arrive_and_wait(); }
inline void __syncwarp() { t_cta->wbar[threadIdx.x >> 5]->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void emu_fence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline long long clock64() { return 0; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
constexpr size_t EMU_SLOT = 16384;      // bytes reserved per `__shared__` declaration site
template <typename T> T* emu_shared(int id, size_t count) {
  if (count * sizeof(T) > EMU_SLOT || (size_t)id >= 64) { fprintf(stderr, "emu_shared: slot %d too small\n", id); abort(); }
  return reinterpret_cast<T*>(t_cta->stat.data() + EMU_SLOT * id);
}
template <typename T> T* emu_dyn_smem() {
  return reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(t_cta->dyn.data()) + 15) & ~uintptr_t(15));
}
template <typename T> T __shfl_xor_sync(unsigned, T v, int o) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* line = t_cta->wscr.data() + 32 * w;
  line[l] = (double)v;
  t_cta->wbar[w]->arrive_and_wait();
  const T r = (T)line[l ^ o];
  t_cta->wbar[w]->arrive_and_wait();
  return r;
}
inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return std::atomic_ref<unsigned int>(*p).fetch_add(v); }
inline unsigned int atomicInc(unsigned int* p, unsigned int lim) {
  std::atomic_ref<unsigned int> a(*p);
  unsigned int cur = a.load();
  while (!a.compare_exchange_weak(cur, cur >= lim ? 0u : cur + 1u)) {}
  return cur;
}
inline unsigned int atomicMax(unsigned int* p, unsigned int v) {
  std::atomic_ref<unsigned int> a(*p);
  unsigned int cur = a.load();
  while (cur < v && !a.compare_exchange_weak(cur, v)) {}
  return cur;
}
inline int atomicExch(int* p, int v) { return std::atomic_ref<int>(*p).exchange(v); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  std::atomic_ref<unsigned long long> a(*p);
  unsigned long long cur = a.load();
  while (cur < v && !a.compare_exchange_weak(cur, v)) {}
  return cur;
}
inline long long __double_as_longlong(double d) { long long l; std::memcpy(&l, &d, 8); return l; }
inline double __longlong_as_double(long long l) { double d; std::memcpy(&d, &l, 8); return d; }
template <typename T> T __ldcg(const T* p) { return std::atomic_ref<T>(*const_cast<T*>(p)).load(); }
inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
using std::fma; using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::min; using std::max;
using std::ceil; using std::log2;

// ------------------------------------------------------------------------------------------------ launches
inline int g_emu_sms = 10;
// persistent worker threads (parked on a semaphore between launches: creating a thread per CUDA thread and launch made
// a solve spend most of its time in clone/exit)
struct EmuWorker { std::binary_semaphore go{0}; std::thread th; };
inline std::vector<std::unique_ptr<EmuWorker>>* g_pool = nullptr;          // leaked on purpose: workers never exit
inline std::function<void(unsigned)> g_task;
inline std::atomic<unsigned> g_left{0};
inline void emu_parallel(unsigned n, std::function<void(unsigned)> f) {
  if (!g_pool) g_pool = new std::vector<std::unique_ptr<EmuWorker>>();
  while (g_pool->size() < n) {
    const unsigned id = (unsigned)g_pool->size();
    g_pool->push_back(std::make_unique<EmuWorker>());
    EmuWorker* w = g_pool->back().get();
    w->th = std::thread([w, id]() {
      for (;;) {
        w->go.acquire();
        g_task(id);
        if (g_left.fetch_sub(1) == 1) g_left.notify_one();
      }
    });
    w->th.detach();
  }
  g_task = std::move(f);
  g_left.store(n);
  for (unsigned i = 0; i < n; ++i) (*g_pool)[i]->go.release();
  for (unsigned v = g_left.load(); v != 0; v = g_left.load()) g_left.wait(v);
}
inline void emu_init_cta(EmuCta& cta, unsigned block, size_t smem) {
  cta.bar = std::make_unique<std::barrier<>>(block);
  cta.wbar.clear();
  for (unsigned w = 0; w < (block + 31) / 32; ++w)
    cta.wbar.push_back(std::make_unique<std::barrier<>>(std::min(32u, block - 32 * w)));
  cta.stat.assign(EMU_SLOT * 64, 0);
  cta.dyn.assign(smem + 64, 0);
  cta.wscr.assign(32 * ((block + 31) / 32), 0.0);
}
// kernel + its arguments, evaluated by the launching thread and copied (what a real launch does with its parameters)
template <typename F, typename... A> auto emu_bind(F f, A... a) { return [=]() { f(a...); }; }
// ordinary launch: the CTAs run one after the other on the same `block` workers (they never wait for each other); a
// separate barrier closes each CTA so that a thread that returned early cannot be mistaken for one at __syncthreads
template <typename F> void emu_launch(dim3 grid, dim3 block, size_t smem, F&& body) {
  EmuCta cta;
  emu_init_cta(cta, block.x, smem);
  std::barrier<> cta_end(block.x);
  const unsigned G = grid.x * grid.y * grid.z, B = block.x;
  emu_parallel(B, [&cta, &cta_end, &body, G, B, grid](unsigned t) {
    t_cta = &cta;
    threadIdx = {t, 0, 0}; blockDim = {B, 1, 1}; gridDim = {grid.x, grid.y, grid.z};
    for (unsigned c = 0; c < G; ++c) {
      blockIdx = {c % grid.x, (c / grid.x) % grid.y, c / (grid.x * grid.y)};
      body();
      cta_end.arrive_and_wait();
    }
  });
}
// cooperative launch: all CTAs at once (grid barrier inside the kernel)
template <typename F> void emu_launch_coop(dim3 grid, dim3 block, size_t smem, F&& body) {
  std::vector<EmuCta> ctas(grid.x);
  for (unsigned c = 0; c < grid.x; ++c) emu_init_cta(ctas[c], block.x, smem);
  const unsigned G = grid.x, B = block.x;
  emu_parallel(G * B, [&ctas, &body, G, B](unsigned id) {
    const unsigned c = id / B, t = id % B;
    t_cta = &ctas[c];
    threadIdx = {t, 0, 0}; blockIdx = {c, 0, 0}; blockDim = {B, 1, 1}; gridDim = {G, 1, 1};
    body();
  });
}

// ------------------------------------------------------------------------------------------------ runtime API
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1, cudaHostAllocMapped = 2, cudaMemcpyDeviceToDevice = 3,
       cudaMemcpyDeviceToHost = 2, cudaDevAttrCooperativeLaunch = 95, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaFuncAttributes { size_t sharedSizeBytes = 0; };
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return 0; }
template <typename F> cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
template <typename F> cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->sharedSizeBytes = 0; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = reinterpret_cast<void*>(1); return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = reinterpret_cast<void*>(1); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return 0; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return 0; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, int) { *p = std::calloc(1, n); return 0; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, int) { *d = h; return 0; }
template <typename T> cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { std::memcpy(&sym, src, n); return 0; }

namespace xt {
inline thread_local char g_err[1024];
inline void set_last_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
#define XT_CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return XT_ERR_CUDA; } while (0)
#define XT_REQUIRE(cond, ...) do { if (!(cond)) { xt::set_last_error(__VA_ARGS__); return XT_ERR_INVALID; } } while (0)
#define XT_LAUNCHED() ((void)0)
inline int num_sms() { return g_emu_sms; }
template <typename K> inline cudaError_t set_max_dyn_smem(K, int = 0) { return 0; }
struct DeviceOnce { bool done = false; bool pending() { return !done; } void mark() { done = true; } };
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
struct Arena {
  char* base; size_t cap; size_t off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <typename T> T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base ? base + off : nullptr);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};
inline unsigned long long gtimer() { return 0; }
inline void cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
inline void cp_async_wait_all() {}
inline void sys_store_release(unsigned long long* p, unsigned long long v) {
  std::atomic_ref<unsigned long long>(*p).store(v, std::memory_order_release);
}
// (the loads below sit in spin loops: yield, or a few thousand emulated threads starve the one being waited for)
inline unsigned long long sys_load_acquire(const unsigned long long* p) {
  std::this_thread::yield();
  return std::atomic_ref<unsigned long long>(*const_cast<unsigned long long*>(p)).load(std::memory_order_acquire);
}
inline void sys_red_add_release(unsigned int* p, unsigned int v) {
  std::atomic_ref<unsigned int>(*p).fetch_add(v, std::memory_order_release);
}
inline unsigned int sys_load_acquire_u32(const unsigned int* p) {
  std::this_thread::yield();
  return std::atomic_ref<unsigned int>(*const_cast<unsigned int*>(p)).load(std::memory_order_acquire);
}
inline void pdl_wait() {}
inline void pdl_trigger() {}
inline double fast_rcp(double x) { return 1.0 / x; }
template <typename T> T warp_sum(T v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T> T warp_max(T v) {
  for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}
template <typename T> T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : T(0);
  return warp_sum(r);
}
template <typename T> T block_max(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : scratch[0];
  return warp_max(r);
}

// the block matvec as plain loops, with the full contract of csrc/matvec.cuh: batch strides, the fused shift
// Y = A X - Z diag(E), and the per-tile partial dot products the solvers consume
constexpr int MV_MAXK = 16;
constexpr int MV_L2_KEEP_MB = 32;
struct MvTiling { int tile_rows, tiles_per_batch, ntiles, grid; };
inline MvTiling mv_tiling(int nbatch, int nrows, int reserve_sms = 0) {
  (void)reserve_sms;
  MvTiling t;
  t.tile_rows = nrows < 48 ? nrows : 48;               // several tiles per batch item already at test sizes
  t.tiles_per_batch = (nrows + t.tile_rows - 1) / t.tile_rows;
  t.ntiles = nbatch * t.tiles_per_batch;
  t.grid = t.ntiles;
  return t;
}
struct MvArgs {
  int dtype; int nbatch, nrows, ncolsA, k;
  const void* A; int64_t lda, a_bstride;
  const void* X; int64_t ldx, x_bstride;
  void* Y; int64_t ldy, y_bstride;
  const void* E; int64_t e_bstride;
  const void* Z; int64_t ldz, z_bstride;
  const void* U; int64_t ldu, u_bstride;
  double* dot_out; int impl; const int* done_flag; const int* abort_flag; int* latch_out; int reserve_sms; int reverse; int l2_keep_mb; int pdl;
};
struct emu_bf16 {
  uint16_t bits;
  explicit operator float() const { uint32_t u = (uint32_t)bits << 16; float f; std::memcpy(&f, &u, 4); return f; }
  explicit operator double() const { return (double)(float)*this; }
};
template <typename TA, typename TV> void emu_mv(const MvArgs& a) {
  const MvTiling til = mv_tiling(a.nbatch, a.nrows, a.reserve_sms);
  for (int b = 0; b < a.nbatch; ++b) {
    const TA* A = static_cast<const TA*>(a.A) + (int64_t)b * a.a_bstride;
    const TV* X = static_cast<const TV*>(a.X) + (int64_t)b * a.x_bstride;
    TV* Y = static_cast<TV*>(a.Y) + (int64_t)b * a.y_bstride;
    const TV* E = a.E ? static_cast<const TV*>(a.E) + (int64_t)b * a.e_bstride : nullptr;
    const TV* Z = a.Z ? static_cast<const TV*>(a.Z) + (int64_t)b * a.z_bstride : X;
    const int64_t ldz = a.Z ? a.ldz : a.ldx;
    const TV* U = a.U ? static_cast<const TV*>(a.U) + (int64_t)b * a.u_bstride : nullptr;
    for (int t = 0; t < til.tiles_per_batch; ++t) {
      double d0[MV_MAXK] = {0}, d1[MV_MAXK] = {0};
      const int r0 = t * til.tile_rows, r1 = std::min(a.nrows, r0 + til.tile_rows);
      for (int r = r0; r < r1; ++r)
        for (int c = 0; c < a.k; ++c) {
          double s = 0.0;
          for (int j = 0; j < a.ncolsA; ++j) s += (double)A[(int64_t)r * a.lda + j] * (double)X[(int64_t)j * a.ldx + c];
          if (E) s -= (double)Z[(int64_t)r * ldz + c] * (double)E[c];
          const TV y = (TV)s;
          Y[(int64_t)r * a.ldy + c] = y;
          d1[c] += (double)y * (double)y;
          if (U) d0[c] += (double)U[(int64_t)r * a.ldu + c] * (double)y;
        }
      if (a.dot_out) {
        double* o = a.dot_out + (size_t)(b * til.tiles_per_batch + t) * 2 * MV_MAXK;
        for (int c = 0; c < a.k; ++c) { o[c] = d0[c]; o[MV_MAXK + c] = d1[c]; }
      }
    }
  }
}
inline int mv_launch(const MvArgs& a, cudaStream_t) {
  if (a.done_flag && *a.done_flag) return XT_OK;
  if (a.abort_flag && *a.abort_flag) return XT_OK;
  if (a.dtype == XT_F32) emu_mv<float, float>(a);
  else if (a.dtype == XT_F64) emu_mv<double, double>(a);
  else if (a.dtype == XT_BF16) emu_mv<emu_bf16, float>(a);
  else return XT_ERR_INVALID;
  return XT_OK;
}
}  // namespace xt

// cooperative launch of a `void kernel(Args)` taken by address
template <typename Args> cudaError_t emu_coop(const void* fn, dim3 grid, dim3 block, void** kargs, size_t smem) {
  auto f = reinterpret_cast<void (*)(Args)>(const_cast<void*>(fn));
  Args a = *static_cast<Args*>(kargs[0]);
  emu_launch_coop(grid, block, smem, [&]() { f(a); });
  return 0;
}
