// CPU emulation harness of tridiag_regs (csrc/symeig.cu): 256 host threads, std::barrier for __syncthreads.
// Built and checked by tests/test_tridiag_emulation.py (the function body is cut out of the .cu file into tridiag_body.inc).
#include <barrier>
#include <thread>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <functional>
#define __device__
#define __noinline__
#define __forceinline__ inline
#define __restrict__
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return {x, y}; }
struct TI { int x; };
static thread_local TI threadIdx;
constexpr int EIG_THREADS = 256;
static std::barrier<>* g_bar;
static std::vector<std::barrier<>*> g_wbar;
static double g_wscr[16][32];
static void __syncthreads() { g_bar->arrive_and_wait(); }
static double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static double fast_rcp(double x) { return 1.0 / x; }
template <typename T> T warp_sum(T v) {
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_wscr[w][l] = v;
  g_wbar[w]->arrive_and_wait();
  T s = 0; for (int i = 0; i < 32; ++i) s += g_wscr[w][i];
  g_wbar[w]->arrive_and_wait();
  return s;
}
using std::fma; using std::fabs;
#include "tridiag_body.inc"

int main(int argc, char** argv) {
  int m = argc > 1 ? atoi(argv[1]) : 33;
  int lds = m | 1;
  std::vector<double> A(m * m), As(m * lds), d(m), e(m), tau(m), xcol(128), xnext(128), scal(4), ppart(16 * 128), sg(96), pvp(32);
  std::vector<double2> vp(130);
  srand(m);
  for (int i = 0; i < m; ++i) for (int j = 0; j <= i; ++j) { double v = rand() / (double)RAND_MAX - 0.5; A[i * m + j] = A[j * m + i] = v; }
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) As[i * lds + j] = A[i * m + j];
  int abort_s = 0;
  std::barrier<> bar(256); g_bar = &bar;
  for (int w = 0; w < 8; ++w) g_wbar.push_back(new std::barrier<>(32));
  std::vector<std::thread> th;
  for (int t = 0; t < 256; ++t) th.emplace_back([&, t]() {
    threadIdx.x = t;
    std::vector<double>& xn = xnext; double* sc = scal.data();
    if (m <= 32) tridiag_regs<4, 1>(As.data(), lds, m, d.data(), e.data(), tau.data(), xcol.data(), xn.data(), vp.data(), ppart.data(), sg.data(), pvp.data(), sc, nullptr, &abort_s);
    else if (m <= 64) tridiag_regs<8, 2>(As.data(), lds, m, d.data(), e.data(), tau.data(), xcol.data(), xn.data(), vp.data(), ppart.data(), sg.data(), pvp.data(), sc, nullptr, &abort_s);
    else if (m <= 96) tridiag_regs<12, 3>(As.data(), lds, m, d.data(), e.data(), tau.data(), xcol.data(), xn.data(), vp.data(), ppart.data(), sg.data(), pvp.data(), sc, nullptr, &abort_s);
    else if (m <= 104) tridiag_regs<13, 4>(As.data(), lds, m, d.data(), e.data(), tau.data(), xcol.data(), xn.data(), vp.data(), ppart.data(), sg.data(), pvp.data(), sc, nullptr, &abort_s);
    else tridiag_regs<16, 4>(As.data(), lds, m, d.data(), e.data(), tau.data(), xcol.data(), xn.data(), vp.data(), ppart.data(), sg.data(), pvp.data(), sc, nullptr, &abort_s);
  });
  for (auto& t : th) t.join();
  // dump everything the checker needs: m, then A (m x m), d, e, tau, and the reflector storage (m x lds)
  printf("%d %d\n", m, lds);
  auto dump = [](const std::vector<double>& v, size_t cnt) { for (size_t i = 0; i < cnt; ++i) printf("%.17g ", v[i]); printf("\n"); };
  dump(A, (size_t)m * m); dump(d, m); dump(e, m); dump(tau, m); dump(As, (size_t)m * lds);
  return 0;
}
