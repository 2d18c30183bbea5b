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
static inline double2 make_double2(double x, double y) { return {x, y}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct EmuIdx { unsigned x, y, z; };
static thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
constexpr int cudaSuccess = 0;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

struct EmuCta {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<std::barrier<>>> wbar;
  std::vector<unsigned char> stat, dyn;
  std::vector<double> wscr;        // warp exchange scratch [nwarps][32]
};
static thread_local EmuCta* t_cta;

static inline void __syncthreads() { t_cta->bar->arrive_and_wait(); }
static inline void __syncwarp() { t_cta->wbar[threadIdx.x >> 5]->arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void emu_fence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline long long clock64() { return 0; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
template <typename T> T* emu_shared(int id, size_t count) {
  (void)count;
  return reinterpret_cast<T*>(t_cta->stat.data() + 4096 * id);
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
static inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
static inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return std::atomic_ref<unsigned int>(*p).fetch_add(v); }
static inline unsigned int atomicInc(unsigned int* p, unsigned int lim) {
  std::atomic_ref<unsigned int> a(*p);
  unsigned int cur = a.load();
  while (!a.compare_exchange_weak(cur, cur >= lim ? 0u : cur + 1u)) {}
  return cur;
}
static inline unsigned int atomicMax(unsigned int* p, unsigned int v) {
  std::atomic_ref<unsigned int> a(*p);
  unsigned int cur = a.load();
  while (cur < v && !a.compare_exchange_weak(cur, v)) {}
  return cur;
}
static inline int atomicExch(int* p, int v) { return std::atomic_ref<int>(*p).exchange(v); }
template <typename T> T __ldcg(const T* p) { return std::atomic_ref<T>(*const_cast<T*>(p)).load(); }
static inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
using std::fma; using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::min; using std::max;
using std::ceil; using std::log2;

// ------------------------------------------------------------------------------------------------ launches
static int g_emu_sms = 10;
// persistent worker threads (parked on a semaphore between launches: creating a thread per CUDA thread and launch made
// a solve spend most of its time in clone/exit)
struct EmuWorker { std::binary_semaphore go{0}; std::thread th; };
static std::vector<std::unique_ptr<EmuWorker>>* g_pool = nullptr;          // leaked on purpose: workers never exit
static std::function<void(unsigned)> g_task;
static std::atomic<unsigned> g_left{0};
static void emu_parallel(unsigned n, std::function<void(unsigned)> f) {
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
static void emu_init_cta(EmuCta& cta, unsigned block, size_t smem) {
  cta.bar = std::make_unique<std::barrier<>>(block);
  cta.wbar.clear();
  for (unsigned w = 0; w < (block + 31) / 32; ++w)
    cta.wbar.push_back(std::make_unique<std::barrier<>>(std::min(32u, block - 32 * w)));
  cta.stat.assign(4096 * 64, 0);
  cta.dyn.assign(smem + 64, 0);
  cta.wscr.assign(32 * ((block + 31) / 32), 0.0);
}
// kernel + its arguments, evaluated by the launching thread and copied (what a real launch does with its parameters)
template <typename F, typename... A> auto emu_bind(F f, A... a) { return [=]() { f(a...); }; }
// ordinary launch: the CTAs run one after the other on the same `block` workers (they never wait for each other); a
// separate barrier closes each CTA so that a thread that returned early cannot be mistaken for one at __syncthreads
template <typename F> static void emu_launch(dim3 grid, dim3 block, size_t smem, F&& body) {
  EmuCta cta;
  emu_init_cta(cta, block.x, smem);
  std::barrier<> cta_end(block.x);
  const unsigned G = grid.x, B = block.x;
  emu_parallel(B, [&cta, &cta_end, &body, G, B](unsigned t) {
    t_cta = &cta;
    threadIdx = {t, 0, 0}; blockDim = {B, 1, 1}; gridDim = {G, 1, 1};
    for (unsigned c = 0; c < G; ++c) {
      blockIdx = {c, 0, 0};
      body();
      cta_end.arrive_and_wait();
    }
  });
}
// cooperative launch: all CTAs at once (grid barrier inside the kernel)
template <typename F> static void emu_launch_coop(dim3 grid, dim3 block, size_t smem, F&& body) {
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
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return 0; }
template <typename F> cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
template <typename F> cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->sharedSizeBytes = 0; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = reinterpret_cast<void*>(1); return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = reinterpret_cast<void*>(1); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, int) { *p = std::calloc(1, n); return 0; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, int) { *d = h; return 0; }
template <typename T> cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { std::memcpy(&sym, src, n); return 0; }

namespace xt {
static thread_local char g_err[1024];
static inline void set_last_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
#define XT_CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return XT_ERR_CUDA; } while (0)
#define XT_REQUIRE(cond, ...) do { if (!(cond)) { xt::set_last_error(__VA_ARGS__); return XT_ERR_INVALID; } } while (0)
#define XT_LAUNCHED() ((void)0)
static inline int num_sms() { return g_emu_sms; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
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
static inline unsigned long long gtimer() { return 0; }
static inline void cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
static inline void cp_async_wait_all() {}
static inline double fast_rcp(double x) { return 1.0 / x; }
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

// the block matvec as a plain loop (fp32 / fp64 operator, one batch item, no shift)
constexpr int MV_MAXK = 16;
constexpr int MV_L2_KEEP_MB = 32;
struct MvArgs {
  int dtype; int nbatch, nrows, ncolsA, k;
  const void* A; int64_t lda, a_bstride;
  const void* X; int64_t ldx, x_bstride;
  void* Y; int64_t ldy, y_bstride;
  const void* E; int64_t e_bstride;
  const void* Z; int64_t ldz, z_bstride;
  const void* U; int64_t ldu, u_bstride;
  double* dot_out; int impl; const int* done_flag; int reserve_sms; int reverse; int l2_keep_mb;
};
template <typename TA> static void emu_mv(const MvArgs& a) {
  const TA* A = static_cast<const TA*>(a.A);
  const TA* X = static_cast<const TA*>(a.X);
  TA* Y = static_cast<TA*>(a.Y);
  for (int r = 0; r < a.nrows; ++r)
    for (int c = 0; c < a.k; ++c) {
      double s = 0.0;
      for (int j = 0; j < a.ncolsA; ++j) s += (double)A[(int64_t)r * a.lda + j] * (double)X[(int64_t)j * a.ldx + c];
      Y[(int64_t)r * a.ldy + c] = (TA)s;
    }
}
static inline int mv_launch(const MvArgs& a, cudaStream_t) {
  if (a.done_flag && *a.done_flag) return XT_OK;
  if (a.dtype == XT_F32) emu_mv<float>(a);
  else if (a.dtype == XT_F64) emu_mv<double>(a);
  else return XT_ERR_INVALID;
  return XT_OK;
}
}  // namespace xt

// cooperative launch of a `void kernel(Args)` taken by address
template <typename Args> static cudaError_t emu_coop(const void* fn, dim3 grid, dim3 block, void** kargs, size_t smem) {
  auto f = reinterpret_cast<void (*)(Args)>(const_cast<void*>(fn));
  Args a = *static_cast<Args*>(kargs[0]);
  emu_launch_coop(grid, block, smem, [&]() { f(a); });
  return 0;
}
