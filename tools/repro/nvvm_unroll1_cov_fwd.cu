#define LQGK_BIG
#include <cstdio>
#include <vector>
#include <cstdlib>
// Reproducer: with `#pragma unroll 1` on the FP64 loop nests (LQGK_UNROLL64), nvcc 12.9 -O3 produced device results that differ
// from the host run of the same code; correct with -Xcicc -O1 or without the pragma.  Build: nvcc -std=c++17 -O3 -gencode
// arch=compute_100a,code=sm_100a -DLQGK_UNROLL64_FORCE_1 tools/repro/nvvm_unroll1_cov_fwd.cu ; run on a B200: ./a.out 1
#include "../../lqg_b200/csrc/lqgk_stages.h"
using namespace lqgk;
using DM = Dims<12, 12, 1, 3, 2>;
using C = CovC<DM>;
using SR = CovSeqRev<DM>;
constexpr int T = 4;
struct Sink {
  float* row; int rec;
  LQGK_HD void put(int idx, float v) { row[idx] = v; }
  LQGK_HD void commit(int) { row += rec; }
};
struct Bufs { double *cst, *lc, *L, *K, *Cw, *FU, *JS, *J0; float* rec; };
LQGK_HD void body(const Bufs& b, int save) {
  Sink sink{b.rec, DM::REC};
  cov_fwd_body<DM>(GCst{b.cst, 1, 0}, WView{b.lc, 1}, T, WView{b.L, 1}, WView{b.K, 1}, save != 0, WView{b.Cw, 1}, WView{b.FU, 1}, WView{b.JS, 1}, WView{b.J0, 1}, sink);
}
__global__ void k(Bufs b, int save) { body(b, save); }
int main(int argc, char** argv) {
  int save = argc > 1 ? atoi(argv[1]) : 1;
  constexpr CLayout cl = DM::CL;
  std::vector<double> cst(cl.total, 0.0), L(T * DM::EL), K(T * DM::EK);
  srand(2);
  auto rnd = [] { return 0.3 * (rand() / (double)RAND_MAX - 0.5); };
  for (auto& a : cst) a = rnd();
  for (int i = 0; i < DM::X; ++i) for (int j = 0; j <= i; ++j) cst[cl.N11 + i * (i + 1) / 2 + j] = i == j ? 0.5 : 0.0;
  for (int i = 0; i < DM::Y; ++i) for (int j = 0; j <= i; ++j) cst[cl.Om + i * (i + 1) / 2 + j] = i == j ? 0.7 : 0.0;
  for (auto& a : L) a = rnd();
  for (auto& a : K) a = rnd();
  size_t nCw = T * DM::EC, nFU = T * SR::NSF, nJS = T * SR::NJS, nJ0 = DM::R * DM::D, nrec = T * DM::REC;
  std::vector<double> lc(C::n), Cw(nCw), FU(nFU), JS(nJS), J0(nJ0);
  std::vector<float> rec(nrec, 0.f);
  Bufs h{cst.data(), lc.data(), L.data(), K.data(), Cw.data(), FU.data(), JS.data(), J0.data(), rec.data()};
  body(h, save);
  Bufs d;
  auto dup = [](const void* src, size_t bytes) { void* p; cudaMalloc(&p, bytes); if (src) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice); else cudaMemset(p, 0, bytes); return p; };
  d.cst = (double*)dup(cst.data(), 8 * cst.size()); d.lc = (double*)dup(nullptr, 8 * C::n); d.L = (double*)dup(L.data(), 8 * L.size());
  d.K = (double*)dup(K.data(), 8 * K.size()); d.Cw = (double*)dup(nullptr, 8 * nCw); d.FU = (double*)dup(nullptr, 8 * nFU);
  d.JS = (double*)dup(nullptr, 8 * nJS); d.J0 = (double*)dup(nullptr, 8 * nJ0); d.rec = (float*)dup(nullptr, 4 * nrec);
  k<<<1, 1>>>(d, save);
  std::vector<double> Cd(nCw), JSd(nJS), J0d(nJ0), FUd(nFU);
  std::vector<float> recd(nrec);
  cudaMemcpy(Cd.data(), d.Cw, 8 * nCw, cudaMemcpyDeviceToHost); cudaMemcpy(JSd.data(), d.JS, 8 * nJS, cudaMemcpyDeviceToHost);
  cudaMemcpy(J0d.data(), d.J0, 8 * nJ0, cudaMemcpyDeviceToHost); cudaMemcpy(FUd.data(), d.FU, 8 * nFU, cudaMemcpyDeviceToHost);
  cudaMemcpy(recd.data(), d.rec, 4 * nrec, cudaMemcpyDeviceToHost);
  printf("err=%s save=%d\n", cudaGetErrorString(cudaGetLastError()), save);
  auto cmp = [](const char* n, const double* a, const double* b, size_t m) { double e = 0; for (size_t i = 0; i < m; ++i) e = fmax(e, fabs(a[i] - b[i])); printf("%s err %g\n", n, e); };
  cmp("Cw", Cd.data(), Cw.data(), nCw); cmp("JS", JSd.data(), JS.data(), nJS); cmp("J0", J0d.data(), J0.data(), nJ0); cmp("FU", FUd.data(), FU.data(), nFU);
  for (int t = 0; t < T; ++t) {
    double ef = 0, ej = 0, el = 0;
    for (int i = 0; i < DM::REC_J; ++i) ef = fmax(ef, fabs(recd[t * DM::REC + i] - rec[t * DM::REC + i]));
    for (int i = DM::REC_J; i < DM::REC_LINV; ++i) ej = fmax(ej, fabs(recd[t * DM::REC + i] - rec[t * DM::REC + i]));
    for (int i = DM::REC_LINV; i <= DM::REC_LOGDET; ++i) el = fmax(el, fabs(recd[t * DM::REC + i] - rec[t * DM::REC + i]));
    printf("t=%d rec F err %g J err %g Linv/logdet err %g\n", t, ef, ej, el);
  }
}
