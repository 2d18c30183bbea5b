// Common device/host helpers for the xitorch_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

#include "../../include/xitorch_b200.h"

namespace xt {

// ----------------------------------------------------------------------------- host error plumbing
void set_last_error(const char* fmt, ...);

#define XT_CUDA_OK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      xt::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return XT_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define XT_REQUIRE(cond, ...)                                                             \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      xt::set_last_error(__VA_ARGS__);                                                    \
      return XT_ERR_INVALID;                                                              \
    }                                                                                     \
  } while (0)

int num_sms();  // SM count of the current device (cached)

// One-time set-up that is PER DEVICE (cudaFuncSetAttribute applies to the current device only): a bit per device ordinal.
// Two threads may both find the bit clear and both run the set-up -- it is idempotent -- but a launch never precedes it.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  int dev = -1;
  bool pending() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    return (mask.load(std::memory_order_acquire) >> d & 1ull) == 0;
  }
  void mark() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return;
    mask.fetch_or(1ull << d, std::memory_order_release);
  }
};

// Opt a kernel in to the largest dynamic shared-memory size the device allows next to the kernel's own static shared
// memory (227 KB per CTA on sm_100 in total; asking for 227 KB of dynamic memory on a kernel with any static
// __shared__ variable is rejected with "invalid argument").
template <typename K> static inline cudaError_t set_max_dyn_smem(K kern, int total = 227 * 1024) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, total - (int)fa.sharedSizeBytes);
}

// launch accounting / in-situ kernel timing (xt_profile_* in the C ABI)
void note_launch(int n = 1);                        // every kernel launch site calls this
bool prof_on();                                     // event profiling of the matvec launches is enabled (xt_profile_reset)
void prof_mv_begin(cudaStream_t st);                // CUDA events around each block-matvec launch when enabled
void prof_mv_end(cudaStream_t st);
#define XT_LAUNCHED() xt::note_launch(1)

template <typename T> struct VecOf { using type = float; };
template <> struct VecOf<double> { using type = double; };

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller-provided workspace (the library never allocates)
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <typename T> T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base ? base + off : nullptr);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

// ----------------------------------------------------------------------------- device PTX wrappers
#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// system-scope release / acquire on peer-visible memory (flags and counters other GPUs write over NVLink)
__device__ __forceinline__ void sys_store_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long sys_load_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sys_red_add_release(unsigned int* p, unsigned int v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int sys_load_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// add to the pending transaction count of the current phase WITHOUT arriving (the arrival follows later)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Programmatic dependent launch (a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream is still running): pdl_wait() returns once the predecessor has completed and its
// writes are visible -- nothing the predecessor produces may be touched before it; pdl_trigger() lets the successor's
// CTAs be scheduled as soon as this kernel's CTAs leave their SMs.  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "XT_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra XT_DONE;\n"
      "bra XT_WAIT;\n"
      "XT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// one probe of the barrier phase (the hardware suspends the thread for a bounded time before it reports failure)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// mbar_wait that gives up when `*abort_at` (shared memory, written by the CTA's TMA producer) says that chunk `seq`
// will never be produced; returns false in that case
__device__ __forceinline__ bool mbar_wait_abortable(uint64_t* bar, uint32_t parity, const int* abort_at, int seq) {
  while (!mbar_try_wait(bar, parity)) {
    if (*reinterpret_cast<const volatile int*>(abort_at) <= seq) return false;
  }
  return true;
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// TMA: 3-D tiled tensor load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
// TMA: 1-D bulk copy global -> shared (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// block-wide sum (all threads get the result); `scratch` holds >= 32 T's
template <typename T> __device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : T(0);
  r = warp_sum(r);
  return r;
}
template <typename T> __device__ __forceinline__ T block_max(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : scratch[0];
  r = warp_max(r);
  return r;
}

#endif  // __CUDACC__

}  // namespace xt
