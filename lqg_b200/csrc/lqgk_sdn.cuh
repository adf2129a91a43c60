// lqgk_sdn.cuh -- Todorov (2005) alternating control / estimator gain iterations under signal-dependent noise, ONE kernel.
//
// EXTENSION: the reference has no signal-dependent noise (docs/README.md:60-62), so this has no reference counterpart; the
// north star asks for it ("lqg/control's backward Riccati recursion ..., including Todorov-style alternating
// control/estimator gain iterations under signal-dependent noise, becomes one kernel").  Mathematical spec and validation:
// oracle/sdn_np.py, tests/test_sdn.py (Monte Carlo, reduction to lqr.backward when C = D = 0).  Predictor-form convention
// of the paper: u_t = -L_t xhat_t, xhat_{t+1} = A xhat_t + B u_t + K_t (y_t - H xhat_t).
//
// Mapping: one thread per parameter sample (b <= 6: every matrix in registers), all sweeps inside the kernel.  The gain
// sequences are the only O(T) state: the backward pass reads K_t and writes L_t, the forward pass reads L_t and writes K_t,
// both directly in the caller's output arrays L[S][T][u][b], K[S][T][b][y] (FP64).
#pragma once
#include "lqgk_core.h"

namespace lqgk {

struct SdnArgs {
  const double *A, *B, *H, *C, *D, *Q, *R, *Qf, *Omxi, *Omom, *Sig1, *xh1;
  long long sA, sB, sH, sC, sD, sQ, sR, sQf, sOmxi, sOmom, sSig1, sxh1;   // elements between samples (0 = shared)
  int S, T, nc, nd, sweeps;
  double *L, *K, *cost;
};

template <class DM>
struct Sdn {
  static constexpr int B = DM::B, U = DM::U, Y = DM::Y;

  // backward pass (Todorov 2005, eq. 4.2): K -> L; returns the expected cost when `want_cost`
  LQGK_HD static double backward(const SdnArgs& a, int s, const double* A, const double* Bm, const double* H, const double* Q,
                                 const double* R, const double* Qf, const double* Omxi, const double* Omom, const double* Sig1,
                                 const double* xh1, bool k_is_zero) {
    double Sx[B * B], Se[B * B], sc = 0.0;
    LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) { Sx[i] = Qf[i]; Se[i] = 0.0; }
    for (int t = a.T - 1; t >= 0; --t) {
      double Kt[B * Y];
      const double* kp = a.K + ((size_t)s * a.T + t) * (B * Y);
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) Kt[i] = k_is_zero ? 0.0 : kp[i];
      double SB[B * U], SA[B * B], M[U * U], G[U * B];
      mm<B, B, U>(Sx, Bm, SB);
      mm<B, B, B>(Sx, A, SA);
      LQGK_UNROLL64 for (int i = 0; i < U * U; ++i) M[i] = R[i];
      mm_tn<U, B, U, true>(Bm, SB, M);                       // R + B' Sx B
      for (int c = 0; c < a.nc; ++c) {                       // + sum_i C_i' (Sx + Se) C_i
        const double* Ci = a.C + (size_t)s * a.sC + (size_t)c * (B * U);
        double SC[B * U], Cl[B * U];
        LQGK_UNROLL64 for (int i = 0; i < B * U; ++i) Cl[i] = Ci[i];
        LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < U; ++j) {
          double v = 0.0;
          LQGK_UNROLL64 for (int k = 0; k < B; ++k) v += (Sx[i * B + k] + Se[i * B + k]) * Cl[k * U + j];
          SC[i * U + j] = v;
        }
        mm_tn<U, B, U, true>(Cl, SC, M);
      }
      mm_tn<U, B, B>(Bm, SA, G);                             // B' Sx A
      double Mi[U * U], Li[U * U], Lt[U * B];
      symmetrize<U>(M);
      chol_and_inverse<U>(M, Li);
      mm_tn<U, U, U>(Li, Li, Mi);
      mm<U, U, B>(Mi, G, Lt);                                // L_t = M^-1 B' Sx A
      double* lp = a.L + ((size_t)s * a.T + t) * (U * B);
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) lp[i] = Lt[i];
      // s += tr(Sx Om_xi + Se (Om_xi + K Om_om K'))
      double KO[B * Y], KOK[B * B];
      mm<B, Y, Y>(Kt, Omom, KO);
      mm_nt<B, Y, B>(KO, Kt, KOK);
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j)
        sc += Sx[i * B + j] * Omxi[j * B + i] + Se[i * B + j] * (Omxi[j * B + i] + KOK[j * B + i]);
      // Sx' = Q + A' Sx (A - B L) + sum_i D_i' K' Se K D_i ;  Se' = A' Sx B L + (A - K H)' Se (A - K H)
      double ABL[B * B], AKH[B * B], Sxn[B * B], Sen[B * B], T1[B * B];
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) { ABL[i] = A[i]; AKH[i] = A[i]; }
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double v = 0.0, w = 0.0;
        LQGK_UNROLL64 for (int k = 0; k < U; ++k) v += Bm[i * U + k] * Lt[k * B + j];
        LQGK_UNROLL64 for (int k = 0; k < Y; ++k) w += Kt[i * Y + k] * H[k * B + j];
        ABL[i * B + j] -= v;
        AKH[i * B + j] -= w;
      }
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sxn[i] = Q[i];
      mm_tn<B, B, B, true>(SA, ABL, Sxn);                    // (Sx A)' (A - B L) = A' Sx (A - B L)   (Sx symmetric)
      for (int c = 0; c < a.nd; ++c) {
        const double* Di = a.D + (size_t)s * a.sD + (size_t)c * (Y * B);
        double KDm[B * B], Dl[Y * B], SKD[B * B];
        LQGK_UNROLL64 for (int i = 0; i < Y * B; ++i) Dl[i] = Di[i];
        mm<B, Y, B>(Kt, Dl, KDm);
        mm<B, B, B>(Se, KDm, SKD);
        mm_tn<B, B, B, true>(KDm, SKD, Sxn);
      }
      mm<B, B, B>(Se, AKH, T1);
      mm_tn<B, B, B>(AKH, T1, Sen);                          // (A - K H)' Se (A - K H)
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double v = 0.0;                                      // A' Sx B L = (Sx A)' (B L) ; B L = A - ABL
        LQGK_UNROLL64 for (int k = 0; k < B; ++k) v += SA[k * B + i] * (A[k * B + j] - ABL[k * B + j]);
        Sen[i * B + j] += v;
      }
      symmetrize<B>(Sxn);
      symmetrize<B>(Sen);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) { Sx[i] = Sxn[i]; Se[i] = Sen[i]; }
    }
    double cost = sc;
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j)
      cost += xh1[i] * Sx[i * B + j] * xh1[j] + (Sx[i * B + j] + Se[i * B + j]) * Sig1[j * B + i];
    return cost;
  }

  // forward pass (eq. 5.2, no internal estimator noise): L -> K
  LQGK_HD static void forward(const SdnArgs& a, int s, const double* A, const double* Bm, const double* H, const double* Omxi,
                              const double* Omom, const double* Sig1, const double* xh1) {
    double Se[B * B], Sx[B * B], Sxe[B * B];
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
      Se[i * B + j] = Sig1[i * B + j];
      Sx[i * B + j] = xh1[i] * xh1[j];
      Sxe[i * B + j] = 0.0;
    }
    for (int t = 0; t < a.T; ++t) {
      double Lt[U * B];
      const double* lp = a.L + ((size_t)s * a.T + t) * (U * B);
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) Lt[i] = lp[i];
      double HS[Y * B], G[Y * Y];
      mm<Y, B, B>(H, Se, HS);
      LQGK_UNROLL64 for (int i = 0; i < Y * Y; ++i) G[i] = Omom[i];
      mm_nt<Y, B, Y, true>(HS, H, G);                        // H Se H' + Om_om
      if (a.nd > 0) {
        double tot[B * B];
        LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j)
          tot[i * B + j] = Se[i * B + j] + Sx[i * B + j] + Sxe[i * B + j] + Sxe[j * B + i];
        for (int c = 0; c < a.nd; ++c) {
          const double* Di = a.D + (size_t)s * a.sD + (size_t)c * (Y * B);
          double Dl[Y * B], DT[Y * B];
          LQGK_UNROLL64 for (int i = 0; i < Y * B; ++i) Dl[i] = Di[i];
          mm<Y, B, B>(Dl, tot, DT);
          mm_nt<Y, B, Y, true>(DT, Dl, G);
        }
      }
      double Li[Y * Y], Gi[Y * Y], ASH[B * Y], Kt[B * Y];
      symmetrize<Y>(G);
      chol_and_inverse<Y>(G, Li);
      mm_tn<Y, Y, Y>(Li, Li, Gi);
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < Y; ++j) {
        double v = 0.0;                                      // A Se H' = A (H Se)'
        LQGK_UNROLL64 for (int k = 0; k < B; ++k) v += A[i * B + k] * HS[j * B + k];
        ASH[i * Y + j] = v;
      }
      mm<B, Y, Y>(ASH, Gi, Kt);
      double* kp = a.K + ((size_t)s * a.T + t) * (B * Y);
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) kp[i] = Kt[i];
      double ABL[B * B], AKH[B * B], KH[B * B];
      mm<B, Y, B>(Kt, H, KH);
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double v = 0.0;
        LQGK_UNROLL64 for (int k = 0; k < U; ++k) v += Bm[i * U + k] * Lt[k * B + j];
        ABL[i * B + j] = A[i * B + j] - v;
        AKH[i * B + j] = A[i * B + j] - KH[i * B + j];
      }
      double SeA[B * B], Sen[B * B], Sxn[B * B], Sxen[B * B], T1[B * B], T2[B * B];
      mm_nt<B, B, B>(Se, A, SeA);                            // Se A'
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sen[i] = Omxi[i];
      mm<B, B, B, true>(AKH, SeA, Sen);                      // + (A - K H) Se A'
      for (int c = 0; c < a.nc; ++c) {
        const double* Ci = a.C + (size_t)s * a.sC + (size_t)c * (B * U);
        double CL[B * B], Cl[B * U], CLS[B * B];
        LQGK_UNROLL64 for (int i = 0; i < B * U; ++i) Cl[i] = Ci[i];
        mm<B, U, B>(Cl, Lt, CL);
        mm<B, B, B>(CL, Sx, CLS);
        mm_nt<B, B, B, true>(CLS, CL, Sen);                  // + C_i L Sx L' C_i'
      }
      mm<B, B, B>(KH, SeA, Sxn);                             // K H Se A'
      mm<B, B, B>(ABL, Sx, T1);
      mm_nt<B, B, B, true>(T1, ABL, Sxn);                    // + (A - B L) Sx (A - B L)'
      mm<B, B, B>(ABL, Sxe, T2);                             // (A - B L) Sxe
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double v = 0.0;                                      // T2 (K H)' + (K H) T2'
        LQGK_UNROLL64 for (int k = 0; k < B; ++k) v += T2[i * B + k] * KH[j * B + k] + KH[i * B + k] * T2[j * B + k];
        Sxn[i * B + j] += v;
      }
      mm_nt<B, B, B>(T2, AKH, Sxen);                         // (A - B L) Sxe (A - K H)'
      symmetrize<B>(Sen);
      symmetrize<B>(Sxn);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) { Se[i] = Sen[i]; Sx[i] = Sxn[i]; Sxe[i] = Sxen[i]; }
    }
  }
};

#if defined(__CUDACC__)
template <class DM>
__global__ void __launch_bounds__(32) k_sdn_gains(SdnArgs a) {
  constexpr int B = DM::B, U = DM::U, Y = DM::Y;
  const int s = blockIdx.x * 32 + threadIdx.x;
  if (s >= a.S) return;
  double A[B * B], Bm[B * U], H[Y * B], Q[B * B], R[U * U], Qf[B * B], Omxi[B * B], Omom[Y * Y], Sig1[B * B], xh1[B];
  auto ld = [&](const double* p, long long stride, double* out, int n) { for (int i = 0; i < n; ++i) out[i] = p[(size_t)s * stride + i]; };
  ld(a.A, a.sA, A, B * B); ld(a.B, a.sB, Bm, B * U); ld(a.H, a.sH, H, Y * B); ld(a.Q, a.sQ, Q, B * B); ld(a.R, a.sR, R, U * U);
  ld(a.Qf, a.sQf, Qf, B * B); ld(a.Omxi, a.sOmxi, Omxi, B * B); ld(a.Omom, a.sOmom, Omom, Y * Y); ld(a.Sig1, a.sSig1, Sig1, B * B);
  ld(a.xh1, a.sxh1, xh1, B);
  for (int sw = 0; sw < a.sweeps; ++sw) {
    Sdn<DM>::backward(a, s, A, Bm, H, Q, R, Qf, Omxi, Omom, Sig1, xh1, sw == 0);
    Sdn<DM>::forward(a, s, A, Bm, H, Omxi, Omom, Sig1, xh1);
  }
  const double cost = Sdn<DM>::backward(a, s, A, Bm, H, Q, R, Qf, Omxi, Omom, Sig1, xh1, a.sweeps == 0);
  if (a.cost) a.cost[s] = cost;
}
#endif

}  // namespace lqgk
