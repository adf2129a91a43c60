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
#include "../../include/lqgk.h"
#include "lqgk_core.h"
#include "lqgk_pack.h"

namespace lqgk {

struct SdnArgs {
  const double *A, *B, *H, *C, *D, *Q, *R, *Qf, *Omxi, *Omom, *Sig1, *xh1;
  long long sA, sB, sH, sC, sD, sQ, sR, sQf, sOmxi, sOmom, sSig1, sxh1;   // elements between samples (0 = shared)
  int S, T, nc, nd, sweeps;
  double *L, *K, *cost;
  int filter_form;   // 0: Todorov's predictor form (u = -L xhat); 1: the reference's filter form (u = L xhat, lqr.py / kf.py conventions)
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
  // ------------------------------------------------------------------------------------------------ FILTER form
  // The same alternating iterations in the reference's own conventions (spec and derivation: oracle/sdn_np.py,
  // filter_backward_pass / filter_forward_pass):  u = L xhat;  xp = A xhat + B u;  xhat' = xp + K (y' - H xp) with y' observing
  // x_{t+1}.  Without multiplicative noise the backward pass IS lqr.backward (lqg/control/lqr.py:16-42, P = 0, no affine
  // terms) and the forward pass IS kf.forward (lqg/belief/kf.py:6-21).
  LQGK_HD static double backward_filter(const SdnArgs& a, int s, const double* A, const double* Bm, const double* H, const double* Q,
                                        const double* R, const double* Qf, const double* Omxi, const double* Omom, const double* Sig1,
                                        const double* xh1, bool k_is_zero) {
    double Sx[B * B], Se[B * B], sc = 0.0;
    LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) { Sx[i] = Qf[i]; Se[i] = 0.0; }
    for (int t = a.T - 1; t >= 0; --t) {
      double Kt[B * Y], IKF[B * B], St[B * B], Wm[B * B], T1[B * B];
      const double* kp = a.K + ((size_t)s * a.T + t) * (B * Y);
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) Kt[i] = k_is_zero ? 0.0 : kp[i];
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double w = i == j ? 1.0 : 0.0;
        LQGK_UNROLL64 for (int k = 0; k < Y; ++k) w -= Kt[i * Y + k] * H[k * B + j];
        IKF[i * B + j] = w;                                  // I - K H
      }
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) St[i] = Sx[i];
      for (int c = 0; c < a.nd; ++c) {                       // St = Sx + sum_j D_j' K' Se K D_j
        const double* Di = a.D + (size_t)s * a.sD + (size_t)c * (Y * B);
        double KDm[B * B], Dl[Y * B], SKD[B * B];
        LQGK_UNROLL64 for (int i = 0; i < Y * B; ++i) Dl[i] = Di[i];
        mm<B, Y, B>(Kt, Dl, KDm);
        mm<B, B, B>(Se, KDm, SKD);
        mm_tn<B, B, B, true>(KDm, SKD, St);
      }
      mm<B, B, B>(Se, IKF, T1);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Wm[i] = St[i];
      mm_tn<B, B, B, true>(IKF, T1, Wm);                     // Wm = St + (I-KH)' Se (I-KH)
      double SB[B * U], SA[B * B], M[U * U], G[U * B];
      mm<B, B, U>(St, Bm, SB);
      mm<B, B, B>(St, A, SA);
      LQGK_UNROLL64 for (int i = 0; i < U * U; ++i) M[i] = R[i];
      mm_tn<U, B, U, true>(Bm, SB, M);                       // R + B' St B
      for (int c = 0; c < a.nc; ++c) {                       // + sum_i C_i' Wm C_i
        const double* Ci = a.C + (size_t)s * a.sC + (size_t)c * (B * U);
        double SC[B * U], Cl[B * U];
        LQGK_UNROLL64 for (int i = 0; i < B * U; ++i) Cl[i] = Ci[i];
        mm<B, B, U>(Wm, Cl, SC);
        mm_tn<U, B, U, true>(Cl, SC, M);
      }
      mm_tn<U, B, B>(Bm, SA, G);                             // B' St A
      double Mi[U * U], Li[U * U], Lt[U * B];
      symmetrize<U>(M);
      chol_and_inverse<U>(M, Li);
      mm_tn<U, U, U>(Li, Li, Mi);
      mm<U, U, B>(Mi, G, Lt);
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) Lt[i] = -Lt[i];   // L_t = -H^-1 B' St A
      double* lp = a.L + ((size_t)s * a.T + t) * (U * B);
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) lp[i] = Lt[i];
      // s += tr(St Om_xi) + tr(Se ((I-KH) Om_xi (I-KH)' + K Om_om K'))
      double KO[B * Y], KOK[B * B], IO[B * B];
      mm<B, Y, Y>(Kt, Omom, KO);
      mm_nt<B, Y, B>(KO, Kt, KOK);
      mm<B, B, B>(IKF, Omxi, IO);
      mm_nt<B, B, B, true>(IO, IKF, KOK);
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j)
        sc += St[i * B + j] * Omxi[j * B + i] + Se[i * B + j] * KOK[j * B + i];
      // Sx' = Q + A' St (A + B L) ;  Se' = -A' St B L + ((I-KH)A)' Se (I-KH)A
      double ABL[B * B], Ab[B * B], Sxn[B * B], Sen[B * B];
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) ABL[i] = A[i];
      mm<B, U, B, true>(Bm, Lt, ABL);
      mm<B, B, B>(IKF, A, Ab);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sxn[i] = Q[i];
      mm_tn<B, B, B, true>(SA, ABL, Sxn);                    // (St A)' (A + B L)
      mm<B, B, B>(Se, Ab, T1);
      mm_tn<B, B, B>(Ab, T1, Sen);
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double v = 0.0;                                      // A' St B L = (St A)' (B L) ; B L = ABL - A
        LQGK_UNROLL64 for (int k = 0; k < B; ++k) v += SA[k * B + i] * (ABL[k * B + j] - A[k * B + j]);
        Sen[i * B + j] -= v;
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

  LQGK_HD static void forward_filter(const SdnArgs& a, int s, const double* A, const double* Bm, const double* H, const double* Omxi,
                                     const double* Omom, const double* Sig1, const double* xh1) {
    double Se[B * B], Sx[B * B], Sxe[B * B];
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
      Se[i * B + j] = Sig1[i * B + j];
      Sx[i * B + j] = xh1[i] * xh1[j];
      Sxe[i * B + j] = 0.0;
    }
    for (int t = 0; t < a.T; ++t) {
      double Lt[U * B], ABL[B * B];
      const double* lp = a.L + ((size_t)s * a.T + t) * (U * B);
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) Lt[i] = lp[i];
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) ABL[i] = A[i];
      mm<B, U, B, true>(Bm, Lt, ABL);                        // A + B L
      // P- = A Se A' + Om_xi + sum_i C_i L Sx L' C_i'
      double T1[B * B], Pm[B * B], ASA[B * B];
      mm<B, B, B>(A, Se, T1);
      mm_nt<B, B, B>(T1, A, ASA);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Pm[i] = ASA[i] + Omxi[i];
      if (a.nc > 0) {
        double LS[U * B], U2[U * U];
        mm<U, B, B>(Lt, Sx, LS);
        mm_nt<U, B, U>(LS, Lt, U2);                          // L Sx L'
        for (int c = 0; c < a.nc; ++c) {
          const double* Ci = a.C + (size_t)s * a.sC + (size_t)c * (B * U);
          double Cl[B * U], CU[B * U];
          LQGK_UNROLL64 for (int i = 0; i < B * U; ++i) Cl[i] = Ci[i];
          mm<B, U, U>(Cl, U2, CU);
          mm_nt<B, U, B, true>(CU, Cl, Pm);
        }
      }
      symmetrize<B>(Pm);
      // Reff = Om_om + sum_j D_j E[x'x''] D_j' ,  E[x'x''] = (A+BL) Sx (A+BL)' + (A+BL) Sxe A' + A Sxe' (A+BL)' + P-
      double Reff[Y * Y], xa[B * B];
      mm_nt<B, B, B>(Sxe, A, xa);                            // E[xhat a'] = Sxe A'
      LQGK_UNROLL64 for (int i = 0; i < Y * Y; ++i) Reff[i] = Omom[i];
      if (a.nd > 0) {
        double X2[B * B], T2[B * B];
        mm<B, B, B>(ABL, Sx, T2);
        LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) X2[i] = Pm[i];
        mm_nt<B, B, B, true>(T2, ABL, X2);
        mm<B, B, B>(ABL, xa, T2);                            // (A+BL) Sxe A'
        LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) X2[i * B + j] += T2[i * B + j] + T2[j * B + i];
        for (int c = 0; c < a.nd; ++c) {
          const double* Di = a.D + (size_t)s * a.sD + (size_t)c * (Y * B);
          double Dl[Y * B], DT[Y * B];
          LQGK_UNROLL64 for (int i = 0; i < Y * B; ++i) Dl[i] = Di[i];
          mm<Y, B, B>(Dl, X2, DT);
          mm_nt<Y, B, Y, true>(DT, Dl, Reff);
        }
      }
      // K = P- H' (H P- H' + Reff)^-1
      double HP[Y * B], G[Y * Y], Li[Y * Y], Gi[Y * Y], Kt[B * Y];
      mm<Y, B, B>(H, Pm, HP);
      LQGK_UNROLL64 for (int i = 0; i < Y * Y; ++i) G[i] = Reff[i];
      mm_nt<Y, B, Y, true>(HP, H, G);
      symmetrize<Y>(G);
      chol_and_inverse<Y>(G, Li);
      mm_tn<Y, Y, Y>(Li, Li, Gi);
      mm_tn<B, Y, Y>(HP, Gi, Kt);                            // (H P-)' G^-1
      double* kp = a.K + ((size_t)s * a.T + t) * (B * Y);
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) kp[i] = Kt[i];
      // moments of (xhat', e'):  IKF = I - K H
      double IKF[B * B], KH[B * B], KRK[B * B], KR[B * Y];
      mm<B, Y, B>(Kt, H, KH);
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) IKF[i * B + j] = (i == j ? 1.0 : 0.0) - KH[i * B + j];
      mm<B, Y, Y>(Kt, Reff, KR);
      mm_nt<B, Y, B>(KR, Kt, KRK);                           // K Reff K'
      double Sen[B * B], Sxn[B * B], Sxen[B * B], T3[B * B], T4[B * B];
      mm<B, B, B>(IKF, Pm, T3);                              // (I-KH) P-
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sen[i] = KRK[i];
      mm_nt<B, B, B, true>(T3, IKF, Sen);                    // Se' = (I-KH) P- (I-KH)' + K Reff K'
      mm<B, B, B>(ABL, Sx, T4);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sxn[i] = KRK[i];
      mm_nt<B, B, B, true>(T4, ABL, Sxn);                    // (A+BL) Sx (A+BL)' + K Reff K'
      double KHP[B * B], AX[B * B];
      mm<B, B, B>(KH, Pm, KHP);                              // K H P-
      mm_nt<B, B, B, true>(KHP, KH, Sxn);                    // + K H P- H' K'
      mm<B, B, B>(ABL, xa, AX);                              // (A+BL) E[xhat a']
      mm_nt<B, B, B>(AX, KH, T4);                            // (A+BL) xa H' K'
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) Sxn[i * B + j] += T4[i * B + j] + T4[j * B + i];
      mm_nt<B, B, B>(AX, IKF, Sxen);                         // Sxe' = (A+BL) xa (I-KH)' + K H P- (I-KH)' - K Reff K'
      mm_nt<B, B, B, true>(KHP, IKF, Sxen);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sxen[i] -= KRK[i];
      symmetrize<B>(Sen);
      symmetrize<B>(Sxn);
      LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) { Se[i] = Sen[i]; Sx[i] = Sxn[i]; Sxe[i] = Sxen[i]; }
    }
  }
  // all sweeps of one parameter sample (shared by the kernel and the CPU test harness); returns the expected cost
  LQGK_HD static double solve(const SdnArgs& a, int s) {
    double A[B * B], Bm[B * U], H[Y * B], Q[B * B], R[U * U], Qf[B * B], Omxi[B * B], Omom[Y * Y], Sig1[B * B], xh1[B];
    auto ld = [&](const double* p, long long stride, double* out, int n) { for (int i = 0; i < n; ++i) out[i] = p[(size_t)s * stride + i]; };
    ld(a.A, a.sA, A, B * B); ld(a.B, a.sB, Bm, B * U); ld(a.H, a.sH, H, Y * B); ld(a.Q, a.sQ, Q, B * B); ld(a.R, a.sR, R, U * U);
    ld(a.Qf, a.sQf, Qf, B * B); ld(a.Omxi, a.sOmxi, Omxi, B * B); ld(a.Omom, a.sOmom, Omom, Y * Y); ld(a.Sig1, a.sSig1, Sig1, B * B);
    ld(a.xh1, a.sxh1, xh1, B);
    for (int sw = 0; sw < a.sweeps; ++sw) {
      if (a.filter_form) {
        backward_filter(a, s, A, Bm, H, Q, R, Qf, Omxi, Omom, Sig1, xh1, sw == 0);
        forward_filter(a, s, A, Bm, H, Omxi, Omom, Sig1, xh1);
      } else {
        backward(a, s, A, Bm, H, Q, R, Qf, Omxi, Omom, Sig1, xh1, sw == 0);
        forward(a, s, A, Bm, H, Omxi, Omom, Sig1, xh1);
      }
    }
    return a.filter_form ? backward_filter(a, s, A, Bm, H, Q, R, Qf, Omxi, Omom, Sig1, xh1, a.sweeps == 0)
                         : backward(a, s, A, Bm, H, Q, R, Qf, Omxi, Omom, Sig1, xh1, a.sweeps == 0);
  }
};

// ================================================================================================
// Likelihood under signal-dependent noise (EXTENSION, parity unpinned by the reference; spec: oracle/sdn_np.py
// `sdn_conditional_moments` / `sdn_log_likelihood`).  The reference's experimenter-side filter (lqg/system.py:142-248) in
// the reference's own filter-form conventions plus the multiplicative terms
//     x_{t+1} = A x_t + B u_t + V eps + sum_i eps'_i C_i u_t ,   y_t = F x_{t+1} + W eta + sum_j eta'_j D_j x_{t+1} ,
// with the gains (L_t, K_t) given.  The predictive covariance now depends on the second moment of the conditioned state,
// i.e. on the running mean of the TRIAL: one dense system per (parameter sample x trial), carried in the reduced
// condition-then-predict form of the main path (c = E[unobserved | x_0..t], C = its covariance), all FP64:
//     Z2 = C + c c^T ;  U2 = L Z2[xhat,xhat] L^T ;  g_i = [C_i ; K F_d C_i] ;  h_j = K D_j
//     mu' = F [x_t ; c] ;  Sig' = F[:,u] C F[:,u]^T + N + sum_i g_i U2 g_i^T ;  Sig'[xhat,xhat] += sum_j h_j (Sig'[x,x] + mu'_x mu'_x^T) h_j^T
//     e = x_{t+1} - mu'[o] ;  ll += log N(e; 0, Sig'[o,o]) ;  J = Sig'[u,o] Sig'[o,o]^-1 ;  c = mu'[u] + J e ;  C = Sig'[u,u] - J Sig'[o,o] J^T
// With nc = nd = 0 this is the main path's recursion (same joint_F / joint_N / predict / condition code) with the per-trial
// part in FP64 instead of FP32.
constexpr int SDN_MAX_TERMS = 4;   // multiplicative-noise matrices per kind

struct SdnLikArgs {
  LqgkSpec act, dyn;          // time-invariant; A, B, F of the actor and A, B, F, V, W of the dynamics are read
  const void *L, *K;          // gains [S][T][u][b], [S][T][b][y] in the I/O type
  LqgkMat C, D;               // C[nc][x][u], D[nd][y][x] per sample (sample_stride 0 = shared; time_stride ignored)
  const float* x_tm;          // observations [T+1][N][d] (per sample when x_sample_stride != 0)
  long long x_sample_stride;
  void* ll_out;               // [S][N]
  int S, N, T, nc, nd;
};

template <class DM>
struct SdnLik {
  static constexpr int X = DM::X, B = DM::B, U = DM::U, Y = DM::Y, D = DM::D, N = DM::N, R = DM::R, XU = DM::XU;
  using CF = CovFwd<DM>;
  using C = CovC<DM>;
  // The constants the joint system needs, in the CovC layout the main path's k_pack produces (lqgk_pack.h), straight from
  // the base matrices of sample s; Fd (Y x X) separately (the multiplicative control noise reaches xhat through K F_d C_i).
  template <class T>
  LQGK_HD static void load_consts(const LqgkSpec& act, const LqgkSpec& dyn, int s, double* lc, double* Fd) {
    double Ad[X * X], Bd[X * U], Vd[X * X], Wd[Y * Y], Aa[B * B], Ba[B * U], Fa[Y * B];
    for (int i = 0; i < X * X; ++i) { Ad[i] = mat_at<T>(dyn.A, s, 0, i); Vd[i] = mat_at<T>(dyn.V, s, 0, i); }
    for (int i = 0; i < X * U; ++i) Bd[i] = mat_at<T>(dyn.B, s, 0, i);
    for (int i = 0; i < Y * X; ++i) Fd[i] = mat_at<T>(dyn.F, s, 0, i);
    for (int i = 0; i < Y * Y; ++i) Wd[i] = mat_at<T>(dyn.W, s, 0, i);
    for (int i = 0; i < B * B; ++i) Aa[i] = mat_at<T>(act.A, s, 0, i);
    for (int i = 0; i < B * U; ++i) Ba[i] = mat_at<T>(act.B, s, 0, i);
    for (int i = 0; i < Y * B; ++i) Fa[i] = mat_at<T>(act.F, s, 0, i);
    for (int i = 0; i < X * X; ++i) lc[C::Ad + i] = Ad[i];
    for (int i = 0; i < X * U; ++i) lc[C::Bd + i] = Bd[i];
    for (int i = 0; i < B * B; ++i) lc[C::Aa + i] = Aa[i];
    for (int i = 0; i < B * U; ++i) lc[C::Ba + i] = Ba[i];
    mm<Y, X, X>(Fd, Ad, lc + C::FAd);
    mm<Y, B, B>(Fa, Aa, lc + C::FAa);
    for (int i = 0; i < Y; ++i) for (int j = 0; j < U; ++j) {          // Dm = F_d B_d - F_a B_a
      double a = 0.0;
      for (int k = 0; k < X; ++k) a += Fd[i * X + k] * Bd[k * U + j];
      for (int k = 0; k < B; ++k) a -= Fa[i * B + k] * Ba[k * U + j];
      lc[C::Dm + i * U + j] = a;
    }
    double N11[X * X], FN[Y * X];
    mm_nt<X, X, X>(Vd, Vd, N11);
    for (int i = 0; i < X; ++i) for (int j = 0; j <= i; ++j) lc[C::N11 + i * (i + 1) / 2 + j] = N11[i * X + j];
    mm<Y, X, X>(Fd, N11, FN);
    for (int i = 0; i < Y * X; ++i) lc[C::FN + i] = FN[i];
    for (int i = 0; i < Y; ++i) for (int j = 0; j <= i; ++j) {          // Om = F_d N11 F_d^T + W_d W_d^T
      double a = 0.0;
      for (int k = 0; k < X; ++k) a += FN[i * X + k] * Fd[j * X + k];
      for (int k = 0; k < Y; ++k) a += Wd[i * Y + k] * Wd[j * Y + k];
      lc[C::Om + i * (i + 1) / 2 + j] = a;
    }
  }
  // c_0 = 0, C_0 = cov(unobserved | x_0) under Sig_0 = G_0 G_0^T                         system.py:211-212
  template <class V>
  LQGK_HD static void init(const V& lc, const double* K0, double* c, double* Cm) {
    double J0[R * D];
    CF::init(lc, K0, Cm, J0);
    for (int i = 0; i < R; ++i) c[i] = 0.0;
  }
  // one step; returns the log-density of x1 = x_{t+1}[o]
  template <class V>
  LQGK_HD static double step(const V& lc, const double* Fd, int nc, const double* Cs, int nd, const double* Ds, const double* L,
                             const double* K, const double* x0, const double* x1, double* c, double* Cm) {
    double Fj[N * N], Nj[N * N], Sig[N * N], mu[N];
    CF::joint_F(lc, L, K, Fj);
    CF::joint_N(lc, K, Nj);
    CF::predict(Fj, Cm, Nj, Sig);
    for (int i = 0; i < N; ++i) {
      double a = 0.0;
      for (int j = 0; j < D; ++j) a += Fj[i * N + j] * x0[j];
      for (int j = 0; j < R; ++j) a += Fj[i * N + D + j] * c[j];
      mu[i] = a;
    }
    if (nc > 0) {
      double LZ[U * B], U2[U * U], KF[B * X];
      for (int i = 0; i < U; ++i) for (int j = 0; j < B; ++j) {         // L Z2[xhat,xhat]
        double a = 0.0;
        for (int k = 0; k < B; ++k) a += L[i * B + k] * (Cm[(XU + k) * R + XU + j] + c[XU + k] * c[XU + j]);
        LZ[i * B + j] = a;
      }
      mm_nt<U, B, U>(LZ, L, U2);
      mm<B, Y, X>(K, Fd, KF);
      for (int q = 0; q < nc; ++q) {
        const double* Cq = Cs + q * (X * U);
        double g[N * U], gU[N * U];
        for (int i = 0; i < X * U; ++i) g[i] = Cq[i];
        mm<B, X, U>(KF, Cq, g + X * U);
        mm<N, U, U>(g, U2, gU);
        for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) {
          double a = 0.0;
          for (int k = 0; k < U; ++k) a += gU[i * U + k] * g[j * U + k];
          Sig[i * N + j] += a;
          if (j != i) Sig[j * N + i] += a;
        }
      }
    }
    if (nd > 0) {
      double X2[X * X];
      for (int i = 0; i < X; ++i) for (int j = 0; j < X; ++j) X2[i * X + j] = Sig[i * N + j] + mu[i] * mu[j];
      for (int q = 0; q < nd; ++q) {
        double h[B * X], hX[B * X];
        mm<B, Y, X>(K, Ds + q * (Y * X), h);
        mm<B, X, X>(h, X2, hX);
        for (int i = 0; i < B; ++i) for (int j = 0; j <= i; ++j) {
          double a = 0.0;
          for (int k = 0; k < X; ++k) a += hX[i * X + k] * h[j * X + k];
          Sig[(X + i) * N + X + j] += a;
          if (j != i) Sig[(X + j) * N + X + i] += a;
        }
      }
    }
    double Linv[D * D], ld, J[R * D], Cn[R * R], e[D], z[D];
    CF::condition(Sig, Linv, ld, [&](int i, double v) { J[i] = v; }, Cn);
    for (int i = 0; i < D; ++i) e[i] = x1[i] - mu[i];
    double qf = 0.0;
    for (int i = 0; i < D; ++i) {
      double a = 0.0;
      for (int j = 0; j <= i; ++j) a += Linv[i * D + j] * e[j];
      z[i] = a;
      qf += a * a;
    }
    for (int i = 0; i < R; ++i) {
      double a = mu[D + i];
      for (int j = 0; j < D; ++j) a += J[i * D + j] * e[j];
      c[i] = a;
    }
    for (int i = 0; i < R * R; ++i) Cm[i] = Cn[i];
    return -0.5 * qf - ld - 0.91893853320467274178 * D;
  }
  // the whole trial `i` of sample `s` (shared by the kernel and the CPU test harness)
  template <class T>
  LQGK_HD static double trial(const SdnLikArgs& a, int s, int i) {
    double lc[C::n], Fd[Y * X], Cs[SDN_MAX_TERMS * X * U], Ds[SDN_MAX_TERMS * Y * X];
    load_consts<T>(a.act, a.dyn, s, lc, Fd);
    for (int q = 0; q < a.nc * X * U; ++q) Cs[q] = mat_at<T>(a.C, s, 0, q);
    for (int q = 0; q < a.nd * Y * X; ++q) Ds[q] = mat_at<T>(a.D, s, 0, q);
    const T* Lp = (const T*)a.L + (size_t)s * a.T * (U * B);
    const T* Kp = (const T*)a.K + (size_t)s * a.T * (B * Y);
    const float* xp = a.x_tm + (size_t)s * a.x_sample_stride + (size_t)i * D;
    const size_t xrow = (size_t)a.N * D;
    double Lt[U * B], Kt[B * Y], c[R], Cm[R * R], x0[D], x1[D];
    for (int k = 0; k < B * Y; ++k) Kt[k] = (double)Kp[k];
    WView lv{lc, 1};
    init(lv, Kt, c, Cm);
    for (int k = 0; k < D; ++k) x0[k] = (double)xp[k];
    double ll = 0.0;
    for (int t = 0; t < a.T; ++t) {
      for (int k = 0; k < U * B; ++k) Lt[k] = (double)Lp[(size_t)t * (U * B) + k];
      for (int k = 0; k < B * Y; ++k) Kt[k] = (double)Kp[(size_t)t * (B * Y) + k];
      for (int k = 0; k < D; ++k) x1[k] = (double)xp[(size_t)(t + 1) * xrow + k];
      ll += step(lv, Fd, a.nc, Cs, a.nd, Ds, Lt, Kt, x0, x1, c, Cm);
      for (int k = 0; k < D; ++k) x0[k] = x1[k];
    }
    return ll;
  }
};

#if defined(__CUDACC__)
template <class DM>
__global__ void __launch_bounds__(32) k_sdn_gains(SdnArgs a) {
  const int s = blockIdx.x * 32 + threadIdx.x;
  if (s >= a.S) return;
  const double cost = Sdn<DM>::solve(a, s);
  if (a.cost) a.cost[s] = cost;
}
// one thread per (parameter sample, trial) system
template <class DM, class T>
__global__ void __launch_bounds__(64) k_sdn_loglik(SdnLikArgs a) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)a.S * a.N) return;
  const int s = (int)(idx / a.N), i = (int)(idx - (size_t)s * a.N);
  ((T*)a.ll_out)[idx] = (T)SdnLik<DM>::template trial<T>(a, s, i);
}
#endif


}  // namespace lqgk
