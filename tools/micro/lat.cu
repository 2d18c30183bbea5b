// latency micro-benchmarks for the one-CTA eigensolver design (fp64 chains, shuffles, barriers) -- tuning aid
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* clk, int iters, double seed) {
  __shared__ double sm[1024];
  const int tid = threadIdx.x;
  sm[tid] = seed + tid;
  sm[tid + 512] = seed;
  __syncthreads();
  double x = seed, y = seed * 0.5, z = 1.0000001;
  long long t0, t1;
  // 0: dependent DFMA
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, z, y);
  t1 = clock64(); if (tid == 0) clk[0] = t1 - t0;
  // 1: dependent DADD
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = x + y;
  t1 = clock64(); if (tid == 0) clk[1] = t1 - t0;
  // 2: DFMA + DSETP/select on the chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, z, y); if (x == 0.0) x = 1e-300; }
  t1 = clock64(); if (tid == 0) clk[2] = t1 - t0;
  // 3: sqrt
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = sqrt(x + 2.0);
  t1 = clock64(); if (tid == 0) clk[3] = t1 - t0;
  // 4: division
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = 3.0 / (x + 1.5);
  t1 = clock64(); if (tid == 0) clk[4] = t1 - t0;
  // 5: rsqrt
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = rsqrt(x + 2.0);
  t1 = clock64(); if (tid == 0) clk[5] = t1 - t0;
  // 6: 64-bit shuffle + add
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64(); if (tid == 0) clk[6] = t1 - t0;
  // 7: __syncthreads
  t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64(); if (tid == 0) clk[7] = t1 - t0;
  // 8: dependent LDS (pointer chase through values)
  int idx = tid & 511;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { double v = sm[idx]; idx = ((int)v + i) & 511; }
  t1 = clock64(); if (tid == 0) clk[8] = t1 - t0;
  // 9: st.shared + barrier + ld.shared round trip
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { sm[tid] = x; __syncthreads(); x += sm[(tid + 32) & 511]; __syncthreads(); }
  t1 = clock64(); if (tid == 0) clk[9] = t1 - t0;
  // 10: float FMA chain
  float f = (float)seed, g = 1.0001f;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) f = fmaf(f, g, 0.5f);
  t1 = clock64(); if (tid == 0) clk[10] = t1 - t0;
  // 11: 8 independent DFMA chains (throughput per warp)
  double c0 = x, c1 = x + 1, c2 = x + 2, c3 = x + 3, c4 = x + 4, c5 = x + 5, c6 = x + 6, c7 = x + 7;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    c0 = fma(c0, z, y); c1 = fma(c1, z, y); c2 = fma(c2, z, y); c3 = fma(c3, z, y);
    c4 = fma(c4, z, y); c5 = fma(c5, z, y); c6 = fma(c6, z, y); c7 = fma(c7, z, y);
  }
  t1 = clock64(); if (tid == 0) clk[11] = t1 - t0;
  out[tid] = x + idx + f + c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&clk, 16 * 8);
  const char* names[12] = {"DFMA dep", "DADD dep", "DFMA+zero-fix", "sqrt", "div", "rsqrt", "shfl64+add", "syncthreads",
                           "LDS dep(+cvt)", "STS+bar+LDS+bar", "FFMA dep", "8xDFMA indep"};
  for (int threads : {32, 128, 512}) {
    const int iters = 256;
    k_lat<<<1, threads>>>(out, clk, iters, 1.25);
    cudaDeviceSynchronize();
    k_lat<<<1, threads>>>(out, clk, iters, 1.25);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    printf("threads=%d:", threads);
    for (int i = 0; i < 12; ++i) printf("  %s %.1f", names[i], (double)clk[i] / iters);
    printf("\n");
  }
  return 0;
}
