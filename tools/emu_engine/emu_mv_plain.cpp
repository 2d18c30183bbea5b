// CPU emulation harness of the plain-load block matvec (mv_plain_kernel, csrc/matvec.cu): the fallback of the TMA kernels
// for operands they cannot take (unaligned pointers / strides).  `struct MvDev` and the kernel are cut out of the .cu
// file into mv_plain_body.inc by tests/test_mv_plain_emulation.py.
//   emu <input file> <output file>
#include "emu_cuda.h"
namespace xt {
#include "mv_plain_body.inc"
}
#include <cstdio>

template <typename T> static std::vector<T> rd(FILE* f, size_t cnt) {
  std::vector<double> tmp(cnt);
  if (cnt && fread(tmp.data(), 8, cnt, f) != cnt) { fprintf(stderr, "short read\n"); exit(2); }
  return std::vector<T>(tmp.begin(), tmp.end());
}
static std::vector<xt::emu_bf16> rd_bf16(FILE* f, size_t cnt) {
  std::vector<double> tmp(cnt);
  if (cnt && fread(tmp.data(), 8, cnt, f) != cnt) exit(2);
  std::vector<xt::emu_bf16> out(cnt);
  for (size_t i = 0; i < cnt; ++i) { float v = (float)tmp[i]; uint32_t u; std::memcpy(&u, &v, 4); out[i].bits = (uint16_t)(u >> 16); }
  return out;
}

template <typename TA, typename TV> static int run(FILE* fi, FILE* fo, const int* hd, std::vector<TA> A) {
  const int nb = hd[1], nrows = hd[2], ncols = hd[3], k = hd[4], lda = hd[5], ldx = hd[6], ldy = hd[7], tile_rows = hd[8],
            has_e = hd[9], has_z = hd[10], has_u = hd[11], a_batched = hd[12], grid = hd[13];
  auto X = rd<TV>(fi, (size_t)nb * ncols * ldx);
  auto E = rd<TV>(fi, has_e ? (size_t)nb * k : 0);
  auto Z = rd<TV>(fi, has_z ? (size_t)nb * nrows * k : 0);
  auto U = rd<TV>(fi, has_u ? (size_t)nb * nrows * k : 0);
  std::vector<TV> Y((size_t)nb * nrows * ldy, TV(-7));
  xt::MvDev p;
  std::memset(&p, 0, sizeof(p));
  p.nbatch = nb; p.nrows = nrows; p.ncolsA = ncols; p.kvalid = k;
  p.tile_rows = tile_rows; p.tiles_per_batch = (nrows + tile_rows - 1) / tile_rows; p.ntiles = nb * p.tiles_per_batch;
  p.X = X.data(); p.ldx = ldx; p.x_bstride = (int64_t)ncols * ldx;
  p.Y = Y.data(); p.ldy = ldy; p.y_bstride = (int64_t)nrows * ldy;
  if (has_e) { p.E = E.data(); p.e_bstride = k; }
  if (has_z) { p.Z = Z.data(); p.ldz = k; p.z_bstride = (int64_t)nrows * k; }
  if (has_u) { p.U = U.data(); p.ldu = k; p.u_bstride = (int64_t)nrows * k; }
  std::vector<double> dots((size_t)p.ntiles * 2 * xt::MV_MAXK, -1.0);
  p.dot_out = dots.data();
  emu_launch(dim3(grid), dim3(256), 0, emu_bind(xt::mv_plain_kernel<TA, TV>, (const TA*)A.data(), (int64_t)lda,
                                                (int64_t)(a_batched ? (int64_t)nrows * lda : 0), p));
  std::vector<double> yo(Y.begin(), Y.end());
  fwrite(yo.data(), 8, yo.size(), fo);
  fwrite(dots.data(), 8, dots.size(), fo);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* fi = fopen(argv[1], "rb");
  FILE* fo = fopen(argv[2], "wb");
  int hd[14];
  if (!fi || !fo || fread(hd, 4, 14, fi) != 14) return 2;
  const size_t acount = (size_t)(hd[12] ? hd[1] : 1) * hd[2] * hd[5];
  int rc;
  if (hd[0] == 0) rc = run<float, float>(fi, fo, hd, rd<float>(fi, acount));
  else if (hd[0] == 2) rc = run<double, double>(fi, fo, hd, rd<double>(fi, acount));
  else rc = run<xt::emu_bf16, float>(fi, fo, hd, rd_bf16(fi, acount));
  fclose(fi); fclose(fo);
  return rc;
}
