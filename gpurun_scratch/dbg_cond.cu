#define LQGK_BIG
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../lqg_b200/csrc/lqgk_core.h"
using namespace lqgk;
using DM = Dims<12, 12, 1, 3, 2>;
__global__ void k(const double* Sig, double* J, double* C, double* Li, double* ld) {
  double Linv[4], Jl[DM::R * DM::D], Cn[DM::R * DM::R], l;
  double S[DM::N * DM::N];
  for (int i = 0; i < DM::N * DM::N; ++i) S[i] = Sig[i];
  CovFwd<DM>::condition(S, Linv, l, Jl, Cn);
  for (int i = 0; i < DM::R * DM::D; ++i) J[i] = Jl[i];
  for (int i = 0; i < DM::R * DM::R; ++i) C[i] = Cn[i];
  for (int i = 0; i < 4; ++i) Li[i] = Linv[i];
  *ld = l;
}
int main() {
  constexpr int N = DM::N, R = DM::R, D = DM::D;
  std::vector<double> A(N * N), Sig(N * N);
  srand(1);
  for (auto& a : A) a = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { double s = i == j ? 0.1 : 0; for (int k = 0; k < N; ++k) s += A[i * N + k] * A[j * N + k]; Sig[i * N + j] = s; }
  double Lh[4], Jh[R * D], Ch[R * R], ldh;
  CovFwd<DM>::condition(Sig.data(), Lh, ldh, Jh, Ch);
  double *dS, *dJ, *dC, *dL, *dl;
  cudaMalloc(&dS, 8 * N * N); cudaMalloc(&dJ, 8 * R * D); cudaMalloc(&dC, 8 * R * R); cudaMalloc(&dL, 32); cudaMalloc(&dl, 8);
  cudaMemcpy(dS, Sig.data(), 8 * N * N, cudaMemcpyHostToDevice);
  k<<<1, 1>>>(dS, dJ, dC, dL, dl);
  std::vector<double> Jd(R * D), Cd(R * R); double Ld[4], ldd;
  cudaMemcpy(Jd.data(), dJ, 8 * R * D, cudaMemcpyDeviceToHost); cudaMemcpy(Cd.data(), dC, 8 * R * R, cudaMemcpyDeviceToHost);
  cudaMemcpy(Ld, dL, 32, cudaMemcpyDeviceToHost); cudaMemcpy(&ldd, dl, 8, cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  double ej = 0, ec = 0;
  for (int i = 0; i < R * D; ++i) ej = fmax(ej, fabs(Jd[i] - Jh[i]));
  for (int i = 0; i < R * R; ++i) ec = fmax(ec, fabs(Cd[i] - Ch[i]));
  printf("J err %g  C err %g  ld %g %g  Linv %g %g %g %g | %g %g %g %g\n", ej, ec, ldd, ldh, Ld[0], Ld[1], Ld[2], Ld[3], Lh[0], Lh[1], Lh[2], Lh[3]);
  for (int i = 0; i < 6; ++i) printf("J[%d] dev %g host %g\n", i, Jd[i], Jh[i]);
}
