// CPU emulation harness of the one-CTA eigensolver (eig_extreme_device, csrc/symeig.cu): 256 host threads stand in
// for the CUDA threads -- std::barrier = __syncthreads / __syncwarp, warp shuffles and sums go through a per-warp
// scratch line.  The device code itself (eig_plan, sturm_count, tridiag_regs, eig_extreme_device) is cut out of the
// .cu file into eig_body.inc by tests/test_small_eigh_emulation.py, so what is checked here is the shipped source.
//   emu <m> <nev> <mode 0|1> <kind 0 random | 1 clustered | 2 degenerate | 3 graded>
#include <barrier>
#include <thread>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#define __device__
#define __noinline__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return {x, y}; }
struct TI { int x; };
static thread_local TI threadIdx;
static TI blockDim{256};
constexpr int EIG_THREADS = 256;
static std::barrier<>* g_bar;
static std::vector<std::barrier<>*> g_wbar;
static double g_wscr[8][32];
static double g_bscr[256];
static void __syncthreads() { g_bar->arrive_and_wait(); }
static void __syncwarp() { g_wbar[threadIdx.x >> 5]->arrive_and_wait(); }
static long long clock64() { return 0; }
static double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static double fast_rcp(double x) { return 1.0 / x; }
template <typename T> T __shfl_xor_sync(unsigned, T v, int o) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_wscr[w][l] = v;
  g_wbar[w]->arrive_and_wait();
  const T r = g_wscr[w][l ^ o];
  g_wbar[w]->arrive_and_wait();
  return r;
}
template <typename T> T warp_sum(T v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T> T block_max(T v, T*) {
  g_bscr[threadIdx.x] = v;
  g_bar->arrive_and_wait();
  T r = g_bscr[0];
  for (int i = 1; i < 256; ++i) r = g_bscr[i] > r ? g_bscr[i] : r;
  g_bar->arrive_and_wait();
  return r;
}
using std::fma; using std::fabs; using std::fmin; using std::fmax; using std::sqrt; using std::ceil; using std::log2;
#include "eig_body.inc"

int main(int argc, char** argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 33, nev = argc > 2 ? atoi(argv[2]) : 4, mode = argc > 3 ? atoi(argv[3]) : 0;
  const int kind = argc > 4 ? atoi(argv[4]) : 0;
  if (argc > 5) g_eig_debug = atoi(argv[5]);
  const EigPlan pl = eig_plan(m, nev);
  if (pl.inv_slots < 1) { printf("noplan\n"); return 0; }
  const int lds = pl.lds;
  // symmetric test matrix A = Q diag(w) Q^T from a prescribed spectrum (Q: product of random reflectors)
  std::vector<double> w(m), A((size_t)m * m, 0.0);
  srand(1000 * m + 10 * nev + kind);
  auto rnd = []() { return rand() / (double)RAND_MAX - 0.5; };
  for (int i = 0; i < m; ++i) {
    if (kind == 0) w[i] = 4.0 * rnd();
    else if (kind == 1) w[i] = 1.0 + (i / 3) + 1e-9 * (i % 3);          // clusters of three, 1e-9 apart
    else if (kind == 2) w[i] = 1.0 + (i / 2);                            // exactly double eigenvalues
    else w[i] = std::pow(10.0, -8.0 * i / (m > 1 ? m - 1 : 1));        // graded over eight decades
  }
  for (int i = 0; i < m; ++i) A[(size_t)i * m + i] = w[i];
  for (int r = 0; r < 3 && m > 1; ++r) {
    std::vector<double> v(m), t(m);
    double nn = 0.0;
    for (int i = 0; i < m; ++i) { v[i] = rnd(); nn += v[i] * v[i]; }
    for (int i = 0; i < m; ++i) v[i] /= std::sqrt(nn);
    // A <- H A H with H = I - 2 v v^T
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += (pass == 0 ? A[(size_t)i * m + j] : A[(size_t)j * m + i]) * v[j];
        t[i] = s;
      }
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
          if (pass == 0) A[(size_t)i * m + j] -= 2.0 * t[i] * v[j];
          else A[(size_t)j * m + i] -= 2.0 * t[i] * v[j];
        }
    }
  }
  for (int i = 0; i < m; ++i) for (int j = 0; j < i; ++j) { const double s = 0.5 * (A[(size_t)i * m + j] + A[(size_t)j * m + i]); A[(size_t)i * m + j] = A[(size_t)j * m + i] = s; }
  std::vector<double> As((size_t)m * lds, 0.0), sh(pl.smem_bytes / 8 + 64, 0.0), lam(nev + 2, 0.0), Y((size_t)m * nev, 0.0);
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) As[(size_t)i * lds + j] = A[(size_t)i * m + j];
  std::barrier<> bar(256); g_bar = &bar;
  for (int q = 0; q < 8; ++q) g_wbar.push_back(new std::barrier<>(32));
  std::vector<std::thread> th;
  for (int t = 0; t < 256; ++t) th.emplace_back([&, t]() {
    threadIdx.x = t;
    eig_extreme_device(As.data(), lds, m, nev, mode, sh.data(), pl.inv_slots, lam.data(), Y.data());
  });
  for (auto& t : th) t.join();
  printf("%d %d %d %d\n", m, nev, pl.as_in_smem, pl.inv_slots);
  auto dump = [](const std::vector<double>& v, size_t cnt) { for (size_t i = 0; i < cnt; ++i) printf("%.17g ", v[i]); printf("\n"); };
  dump(A, (size_t)m * m); dump(lam, nev); dump(Y, (size_t)m * nev);
  return 0;
}
