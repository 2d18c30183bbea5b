// CPU emulation harness of the fused expansion kernel (expand_fused_kernel, csrc/symeig.cu): G CTAs x 512 host threads.
// Per-CTA std::barrier = __syncthreads, the kernel's own sense-reversing grid barrier runs as written (atomics on host
// memory), fp64 atomics = std::atomic_ref.  The device code (EigCtl, Cholesky helpers, PostArgs .. expand_fused_kernel,
// po_smem_bytes) is cut out of the .cu file into fused_body.inc by tests/test_expand_fused_emulation.py, which also
// rewrites the few constructs a host compiler cannot take (PTX asm, __shared__ declarations).
//   emu <input file> <output file>        (binary, see the test for the layout)
#include <atomic>
#include <barrier>
#include <thread>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#define __device__
#define __global__
#define __noinline__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
#define __align__(x)
struct double2 { double x, y; };
struct float4 { float x, y, z, w; };
struct TI { int x; };
static thread_local TI threadIdx, blockIdx;
static TI gridDim{1};
static thread_local std::barrier<>* t_cta_bar;
static thread_local std::barrier<>* t_warp_bar;
static thread_local unsigned char* t_static_smem;      // per-CTA arena for `__shared__` declarations
static thread_local unsigned char* t_dyn_smem;         // per-CTA dynamic shared memory
static thread_local float* t_red;                      // per-CTA scratch of block_max
static void __syncthreads() { t_cta_bar->arrive_and_wait(); }
static void __syncwarp() { t_warp_bar->arrive_and_wait(); }
static void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static void emu_fence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static long long clock64() { return 0; }
static unsigned long long gtimer() { return 0; }
static double rsqrt(double x) { return 1.0 / std::sqrt(x); }
template <typename T> T* emu_shared(int id, int count) { (void)count; return reinterpret_cast<T*>(t_static_smem + 256 * id); }
static unsigned char* emu_dyn_smem() { return t_dyn_smem; }
static double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
static unsigned int atomicAdd(unsigned int* p, unsigned int v) { return std::atomic_ref<unsigned int>(*p).fetch_add(v); }
static int atomicExch(int* p, int v) { return std::atomic_ref<int>(*p).exchange(v); }
static unsigned int atomicMax(unsigned int* p, unsigned int v) {
  std::atomic_ref<unsigned int> a(*p);
  unsigned int cur = a.load();
  while (cur < v && !a.compare_exchange_weak(cur, v)) {}
  return cur;
}
template <typename T> T __ldcg(const T* p) { return std::atomic_ref<T>(*const_cast<T*>(p)).load(); }
static unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
static float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
static void cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
static void cp_async_wait_all() {}
static void pdl_wait() {}
static void pdl_trigger() {}
template <typename T> T block_max(T v, T*) {
  t_red[threadIdx.x] = (float)v;
  t_cta_bar->arrive_and_wait();
  float r = t_red[0];
  for (int i = 1; i < 512; ++i) r = t_red[i] > r ? t_red[i] : r;
  t_cta_bar->arrive_and_wait();
  return (T)r;
}
using std::fma; using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::min; using std::max;
constexpr int SE_MAXK = 16;
#include "fused_body.inc"

template <typename T> static std::vector<T> rd(FILE* f, size_t cnt) {
  std::vector<double> tmp(cnt);
  if (cnt && fread(tmp.data(), 8, cnt, f) != cnt) { fprintf(stderr, "short read\n"); exit(2); }
  return std::vector<T>(tmp.begin(), tmp.end());
}
template <typename T> static void wr(FILE* f, const T* p, size_t cnt) {
  std::vector<double> tmp(p, p + cnt);
  fwrite(tmp.data(), 8, cnt, f);
}

template <typename TV, int KP> static int run(FILE* fi, FILE* fo, const int* hd) {
  const int n = hd[0], k = hd[1], m = hd[2], G = hd[3], stage_v = hd[5], rz_m = hd[6], rz_ld = hd[7], rz_coff = hd[8],
            iter = hd[9], launches = hd[10];
  double min_eps;
  if (fread(&min_eps, 8, 1, fi) != 1) return 2;
  auto V = rd<TV>(fi, (size_t)(m + k) * n);                 // room for the new block behind the m basis vectors
  auto AV = rd<TV>(fi, (size_t)m * n);
  auto S = rd<double>(fi, (size_t)std::max(rz_m, 1) * std::max(rz_ld, 1));
  auto theta = rd<double>(fi, std::max(rz_ld, 1));
  const int acc_stride = (m + SE_MAXK) * k;
  std::vector<double> acc((size_t)2 * PO_NCOPY * acc_stride, 0.0), T((size_t)m * m, 0.0), evals_slots(2 * SE_MAXK, 0.0);
  std::vector<TV> X((size_t)2 * n * k, TV(0));
  EigCtl ctl;
  std::memset(&ctl, 0, sizeof(ctl));
  ctl.best_resid = INFINITY;
  const int R = (n + G - 1) / G;
  const size_t smem = po_smem_bytes(sizeof(TV), KP, R, k, m, rz_m, stage_v != 0);
  gridDim.x = G;
  PostArgs pa;
  std::memset(&pa, 0, sizeof(pa));
  pa.V = V.data(); pa.AV = AV.data(); pa.W = AV.data() + (size_t)(m - k) * n;      // W = A * (last block)
  pa.Qout = V.data() + (size_t)m * n;
  pa.n = n; pa.k = k; pa.m = m; pa.R = R;
  pa.acc = acc.data(); pa.acc_stride = acc_stride; pa.T = T.data(); pa.ldt = m;
  pa.ctl = &ctl; pa.iter = iter; pa.stage_v = stage_v;
  pa.Xslots = X.data(); pa.evals_slots = evals_slots.data(); pa.min_eps = (float)min_eps;
  if (rz_m > 0) { pa.rz_m = rz_m; pa.rz_iter = iter - 1; pa.rz_ld = rz_ld; pa.rz_coff = rz_coff; pa.rz_S = S.data(); pa.rz_theta = theta.data(); }
  for (int l = 0; l < launches; ++l) {
    if (l > 0) { ctl.done = 0; ctl.converged = 0; ctl.local_done = 0; }       // repeat the launch on the same inputs
    std::vector<std::vector<unsigned char>> dyn(G, std::vector<unsigned char>(smem + 64)), stat(G, std::vector<unsigned char>(256 * 64));
    std::vector<std::vector<float>> red(G, std::vector<float>(512));
    std::vector<std::barrier<>*> cbar, wbar;
    for (int c = 0; c < G; ++c) cbar.push_back(new std::barrier<>(PO_THREADS));
    for (int c = 0; c < G * (PO_THREADS / 32); ++c) wbar.push_back(new std::barrier<>(32));
    std::vector<std::thread> th;
    for (int c = 0; c < G; ++c)
      for (int t = 0; t < PO_THREADS; ++t)
        th.emplace_back([&, c, t]() {
          threadIdx.x = t; blockIdx.x = c;
          t_cta_bar = cbar[c]; t_warp_bar = wbar[c * (PO_THREADS / 32) + t / 32];
          t_static_smem = stat[c].data(); t_red = red[c].data();
          t_dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn[c].data()) + 15) & ~uintptr_t(15));
          expand_fused_kernel<TV, KP>(pa);
        });
    for (auto& t : th) t.join();
    for (auto b : cbar) delete b;
    for (auto b : wbar) delete b;
  }
  wr(fo, V.data() + (size_t)m * n, (size_t)n * k);          // Q
  wr(fo, T.data(), (size_t)m * m);
  wr(fo, acc.data(), acc.size());
  wr(fo, X.data(), X.size());
  wr(fo, evals_slots.data(), evals_slots.size());
  const double c[8] = {(double)ctl.done, (double)ctl.converged, (double)ctl.niter, (double)ctl.best_slot,
                       (double)ctl.best_resid, (double)ctl.breakdown, (double)ctl.bar_count, (double)ctl.resmax_bits};
  fwrite(c, 8, 8, fo);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* fi = fopen(argv[1], "rb");
  FILE* fo = fopen(argv[2], "wb");
  int hd[11];
  if (!fi || !fo || fread(hd, 4, 11, fi) != 11) return 2;
  const int k = hd[1], tv = hd[4];
  const int KP = k <= 4 ? 4 : (k <= 8 ? 8 : 16);
  int rc = 3;
  if (tv == 4) rc = KP == 4 ? run<float, 4>(fi, fo, hd) : KP == 8 ? run<float, 8>(fi, fo, hd) : run<float, 16>(fi, fo, hd);
  else rc = KP == 4 ? run<double, 4>(fi, fo, hd) : KP == 8 ? run<double, 8>(fi, fo, hd) : run<double, 16>(fi, fo, hd);
  fclose(fi); fclose(fo);
  return rc;
}
