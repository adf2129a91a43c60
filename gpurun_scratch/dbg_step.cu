#define LQGK_BIG
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../lqg_b200/csrc/lqgk_core.h"
using namespace lqgk;
using DM = Dims<12, 12, 1, 3, 2>;
using C = CovC<DM>;
struct V1 { const double* p; LQGK_HD double operator()(int e) const { return p[e]; } };
__global__ void k(const double* lc, const double* L, const double* K, const double* C0, float* rec, double* Cout) {
  double Cm[DM::R * DM::R];
  for (int i = 0; i < DM::R * DM::R; ++i) Cm[i] = C0[i];
  V1 v{lc};
  for (int t = 0; t < 3; ++t)
    CovFwd<DM>::step(v, L, K, Cm, [&](int idx, float x) { rec[t * DM::REC + idx] = x; }, [&](int, int, double) {});
  for (int i = 0; i < DM::R * DM::R; ++i) Cout[i] = Cm[i];
}
int main() {
  constexpr int R = DM::R, REC = DM::REC;
  std::vector<double> lc(C::n, 0.0), L(DM::U * DM::B), K(DM::B * DM::Y), C0(R * R, 0.0), Ch(R * R);
  srand(2);
  auto rnd = [] { return 0.3 * (rand() / (double)RAND_MAX - 0.5); };
  for (auto& a : lc) a = rnd();
  for (int i = 0; i < DM::X; ++i) lc[C::N11 + i * (i + 1) / 2 + i] = 0.5;
  for (int i = 0; i < DM::Y; ++i) lc[C::Om + i * (i + 1) / 2 + i] = 0.7;
  for (int i = 0; i < DM::X; ++i) for (int j = 0; j < i; ++j) lc[C::N11 + i * (i + 1) / 2 + j] = 0.0;
  for (int i = 0; i < DM::Y; ++i) for (int j = 0; j < i; ++j) lc[C::Om + i * (i + 1) / 2 + j] = 0.0;
  for (auto& a : L) a = rnd();
  for (auto& a : K) a = rnd();
  for (int i = 0; i < R; ++i) C0[i * R + i] = 1.0;
  std::vector<float> rh(3 * REC, 0.f), rd(3 * REC, 0.f);
  Ch = C0;
  V1 v{lc.data()};
  for (int t = 0; t < 3; ++t)
    CovFwd<DM>::step(v, L.data(), K.data(), Ch.data(), [&](int idx, float x) { rh[t * REC + idx] = x; }, [&](int, int, double) {});
  double *dlc, *dL, *dK, *dC0, *dCo; float* drec;
  cudaMalloc(&dlc, 8 * C::n); cudaMalloc(&dL, 8 * L.size()); cudaMalloc(&dK, 8 * K.size()); cudaMalloc(&dC0, 8 * R * R); cudaMalloc(&dCo, 8 * R * R);
  cudaMalloc(&drec, 4 * 3 * REC);
  cudaMemcpy(dlc, lc.data(), 8 * C::n, cudaMemcpyHostToDevice); cudaMemcpy(dL, L.data(), 8 * L.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dK, K.data(), 8 * K.size(), cudaMemcpyHostToDevice); cudaMemcpy(dC0, C0.data(), 8 * R * R, cudaMemcpyHostToDevice);
  k<<<1, 1>>>(dlc, dL, dK, dC0, drec, dCo);
  std::vector<double> Cd(R * R);
  cudaMemcpy(rd.data(), drec, 4 * 3 * REC, cudaMemcpyDeviceToHost); cudaMemcpy(Cd.data(), dCo, 8 * R * R, cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  for (int t = 0; t < 3; ++t) {
    double ef = 0, ej = 0, el = 0;
    for (int i = 0; i < DM::REC_J; ++i) ef = fmax(ef, fabs(rd[t * REC + i] - rh[t * REC + i]));
    for (int i = DM::REC_J; i < DM::REC_LINV; ++i) ej = fmax(ej, fabs(rd[t * REC + i] - rh[t * REC + i]));
    for (int i = DM::REC_LINV; i <= DM::REC_LOGDET; ++i) el = fmax(el, fabs(rd[t * REC + i] - rh[t * REC + i]));
    printf("t=%d F err %g J err %g Linv/logdet err %g   J[0..3] dev %g %g %g %g host %g %g %g %g\n", t, ef, ej, el, rd[t*REC+576], rd[t*REC+577], rd[t*REC+578], rd[t*REC+579],
           rh[t*REC+576], rh[t*REC+577], rh[t*REC+578], rh[t*REC+579]);
  }
  double ec = 0; for (int i = 0; i < R * R; ++i) ec = fmax(ec, fabs(Cd[i] - Ch[i]));
  printf("C err %g\n", ec);
}
