// lqgk_core.h -- dimension-templated step functions of the LQG likelihood path.
//
// Every function here is __host__ __device__: the CUDA kernels (lqgk_kernels.cuh) call them with one
// thread per parameter sample (FP64 per-sample recursions) or one lane per trial (FP32 per-trial
// recursions); tests/emul compiles the very same code with g++ to check the math on CPU against the
// oracle before any GPU time is spent.  No code here is copied from the reference; each block cites the
// reference lines whose result it reproduces (paths relative to the reference repo).
//
//   LqrFwd   lqg/control/lqr.py:16-42      backward Riccati sweep         (per sample, FP64)
//   KfFwd    lqg/belief/kf.py:6-21         forward Kalman-gain sweep      (per sample, FP64)
//   CovFwd   lqg/system.py:163-212,223-230 joint system + covariance scan (per sample, FP64)
//   TrialFwd lqg/system.py:219-221,244-248 mean scan + MVN log-density    (per trial,  FP32)
//   *Rev     reverse-mode adjoints of the above (JAX autodiff in the reference)
//
// Mathematical spec of all stages: oracle/adjoint_np.py (validated against autograd to 1e-12).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <type_traits>

#if defined(__CUDACC__)
#define LQGK_HD __host__ __device__ __forceinline__
#define LQGK_UNROLL _Pragma("unroll")
#else
#define LQGK_HD inline
#define LQGK_UNROLL
#endif

namespace lqgk {

LQGK_HD constexpr int tri(int n) { return n * (n + 1) / 2; }
LQGK_HD constexpr int round4(int n) { return (n + 3) & ~3; }
LQGK_HD constexpr int sidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

template <int I, int E, class F>
LQGK_HD void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

// ------------------------------------------------------------------------------------------------
// Layout of the per-sample block of derived constants (double, sample-minor in HBM: elem * S + s).
// Usable at run time (generic pack/unpack kernels) and at compile time (templated kernels).
struct CLayout {
  int x, b, u, y;
  int Aa, Ba, Fa, Q, R, Qf, VVa, WWa, Sig0, Ad, Bd, FAd, FAa, Dm, N11, FN, Om, q, r, P, qf, total;
  LQGK_HD constexpr CLayout(int x_, int b_, int u_, int y_)
      : x(x_), b(b_), u(u_), y(y_),
        Aa(0), Ba(Aa + b_ * b_), Fa(Ba + b_ * u_), Q(Fa + y_ * b_), R(Q + tri(b_)), Qf(R + tri(u_)),
        VVa(Qf + tri(b_)), WWa(VVa + tri(b_)), Sig0(WWa + tri(y_)), Ad(Sig0 + tri(b_)), Bd(Ad + x_ * x_),
        FAd(Bd + x_ * u_), FAa(FAd + y_ * x_), Dm(FAa + y_ * b_), N11(Dm + y_ * u_), FN(N11 + tri(x_)),
        Om(FN + y_ * x_), q(Om + tri(y_)), r(q + b_), P(r + u_), qf(P + u_ * b_), total(qf + b_) {}
};

template <int X_, int B_, int U_, int Y_, int D_>
struct Dims {
  static constexpr int X = X_, B = B_, U = U_, Y = Y_, D = D_;
  static constexpr int N = X + B;        // joint state (x, xhat)
  static constexpr int R = N - D;        // unobserved joint dims
  static constexpr int XU = X - D;       // unobserved dynamics-state dims
  static constexpr CLayout CL = CLayout(X, B, U, Y);
  // per-step record the trial kernels consume (float): F (N*N) | J (R*D) | Linv (tri D) | logdet | pad
  static constexpr int REC_F = 0, REC_J = N * N, REC_LINV = REC_J + R * D, REC_LOGDET = REC_LINV + tri(D);
  static constexpr int REC = round4(REC_LOGDET + 1);
  // per-step sums over trials the covariance adjoint consumes (float): Fb (N*N) | Jb (R*D) | Wv (tri D)
  static constexpr int SUM_F = 0, SUM_J = N * N, SUM_W = SUM_J + R * D, NSUM = SUM_W + tri(D);
  static constexpr int SUMP = round4(NSUM);
  // sample-minor FP64 workspace rows per step
  static constexpr int EL = U * B, EK = B * Y, ES = tri(B), EP = tri(B), EC = tri(R);
  static_assert(D >= 1 && D <= X, "observed dims are a prefix of the dynamics state");
};
template <int X_, int B_, int U_, int Y_, int D_>
constexpr CLayout Dims<X_, B_, U_, Y_, D_>::CL;

// View of one sample's column in a sample-minor array: element e lives at p[e * stride].
struct CView {
  const double* p;
  size_t stride;
  LQGK_HD double operator()(int e) const { return p[(size_t)e * stride]; }
};
struct WView {
  double* p;
  size_t stride;
  LQGK_HD double& operator()(size_t e) const { return p[e * stride]; }
};

// ------------------------------------------------------------------------------------------------
// Small dense helpers on row-major statically sized arrays.  Fully unrolled on the device.
template <int M, int K, int N, bool ACC = false, class TA, class TB, class TC>
LQGK_HD void mm(const TA* A, const TB* Bm, TC* C) {  // C[M,N] (+)= A[M,K] B[K,N]
  LQGK_UNROLL for (int i = 0; i < M; ++i) LQGK_UNROLL for (int j = 0; j < N; ++j) {
    TC acc = ACC ? C[i * N + j] : TC(0);
    LQGK_UNROLL for (int k = 0; k < K; ++k) acc += A[i * K + k] * Bm[k * N + j];
    C[i * N + j] = acc;
  }
}
template <int M, int K, int N, bool ACC = false, class TA, class TB, class TC>
LQGK_HD void mm_nt(const TA* A, const TB* Bm, TC* C) {  // C[M,N] (+)= A[M,K] B[N,K]^T
  LQGK_UNROLL for (int i = 0; i < M; ++i) LQGK_UNROLL for (int j = 0; j < N; ++j) {
    TC acc = ACC ? C[i * N + j] : TC(0);
    LQGK_UNROLL for (int k = 0; k < K; ++k) acc += A[i * K + k] * Bm[j * K + k];
    C[i * N + j] = acc;
  }
}
template <int M, int K, int N, bool ACC = false, class TA, class TB, class TC>
LQGK_HD void mm_tn(const TA* A, const TB* Bm, TC* C) {  // C[M,N] (+)= A[K,M]^T B[K,N]
  LQGK_UNROLL for (int i = 0; i < M; ++i) LQGK_UNROLL for (int j = 0; j < N; ++j) {
    TC acc = ACC ? C[i * N + j] : TC(0);
    LQGK_UNROLL for (int k = 0; k < K; ++k) acc += A[k * M + i] * Bm[k * N + j];
    C[i * N + j] = acc;
  }
}
// C[M,M] (+)= A[M,K] B[M,K]^T where the result is known symmetric: lower triangle computed, mirrored.
template <int M, int K, bool ACC = false>
LQGK_HD void mm_nt_sym(const double* A, const double* Bm, double* C) {
  LQGK_UNROLL for (int i = 0; i < M; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
    double acc = ACC ? C[i * M + j] : 0.0;
    LQGK_UNROLL for (int k = 0; k < K; ++k) acc += A[i * K + k] * Bm[j * K + k];
    C[i * M + j] = acc;
    C[j * M + i] = acc;
  }
}
// C[M,M] (+)= A[K,M]^T B[K,M], result known symmetric.
template <int M, int K, bool ACC = false>
LQGK_HD void mm_tn_sym(const double* A, const double* Bm, double* C) {
  LQGK_UNROLL for (int i = 0; i < M; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
    double acc = ACC ? C[i * M + j] : 0.0;
    LQGK_UNROLL for (int k = 0; k < K; ++k) acc += A[k * M + i] * Bm[k * M + j];
    C[i * M + j] = acc;
    C[j * M + i] = acc;
  }
}
template <int M>
LQGK_HD void symmetrize(double* C) {  // C <- (C + C^T)/2
  LQGK_UNROLL for (int i = 0; i < M; ++i) LQGK_UNROLL for (int j = 0; j < i; ++j) {
    double v = 0.5 * (C[i * M + j] + C[j * M + i]);
    C[i * M + j] = v;
    C[j * M + i] = v;
  }
}
template <int M, class V>
LQGK_HD void load_sym(const V& v, int off, double* C) {  // packed lower -> full
  LQGK_UNROLL for (int i = 0; i < M; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
    double a = v(off + i * (i + 1) / 2 + j);
    C[i * M + j] = a;
    C[j * M + i] = a;
  }
}
template <int M, int N, class V>
LQGK_HD void load_mat(const V& v, int off, double* C) {
  LQGK_UNROLL for (int i = 0; i < M * N; ++i) C[i] = v(off + i);
}
// In-place Cholesky of a symmetric positive definite M x M matrix (lower factor; upper part untouched).
template <int M>
LQGK_HD void chol(double* A) {
  LQGK_UNROLL for (int j = 0; j < M; ++j) {
    double djj = A[j * M + j];
    LQGK_UNROLL for (int k = 0; k < j; ++k) djj -= A[j * M + k] * A[j * M + k];
    djj = sqrt(djj);
    A[j * M + j] = djj;
    double inv = 1.0 / djj;
    LQGK_UNROLL for (int i = j + 1; i < M; ++i) {
      double v = A[i * M + j];
      LQGK_UNROLL for (int k = 0; k < j; ++k) v -= A[i * M + k] * A[j * M + k];
      A[i * M + j] = v * inv;
    }
  }
}
// Inverse of a lower-triangular matrix (lower part of Lc) into Li (lower part; upper part zeroed).
template <int M>
LQGK_HD void tri_inv(const double* Lc, double* Li) {
  LQGK_UNROLL for (int i = 0; i < M * M; ++i) Li[i] = 0.0;
  LQGK_UNROLL for (int j = 0; j < M; ++j) {
    Li[j * M + j] = 1.0 / Lc[j * M + j];
    LQGK_UNROLL for (int i = j + 1; i < M; ++i) {
      double v = 0.0;
      LQGK_UNROLL for (int k = j; k < i; ++k) v -= Lc[i * M + k] * Li[k * M + j];
      Li[i * M + j] = v / Lc[i * M + i];
    }
  }
}
// Smallest eigenvalue of a symmetric M x M matrix (closed form for M <= 2, cyclic Jacobi otherwise).
template <int M>
LQGK_HD double lambda_min(const double* H) {
  if constexpr (M == 1) return H[0];
  else if constexpr (M == 2) {
    double a = H[0], c = H[3], o = 0.5 * (H[1] + H[2]);
    double h = 0.5 * (a - c);
    return 0.5 * (a + c) - sqrt(h * h + o * o);
  } else {
  double A[M * M];
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) A[i * M + j] = 0.5 * (H[i * M + j] + H[j * M + i]);
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < M; ++p)
      for (int q = p + 1; q < M; ++q) off += A[p * M + q] * A[p * M + q];
    if (off < 1e-300) break;
    for (int p = 0; p < M; ++p)
      for (int q = p + 1; q < M; ++q) {
        double apq = A[p * M + q];
        if (fabs(apq) < 1e-300) continue;
        double th = (A[q * M + q] - A[p * M + p]) / (2.0 * apq);
        double tt = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
        for (int k = 0; k < M; ++k) {
          double akp = A[k * M + p], akq = A[k * M + q];
          A[k * M + p] = cs * akp - sn * akq;
          A[k * M + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < M; ++k) {
          double apk = A[p * M + k], aqk = A[q * M + k];
          A[p * M + k] = cs * apk - sn * aqk;
          A[q * M + k] = sn * apk + cs * aqk;
        }
      }
  }
  double m = A[0];
  for (int i = 1; i < M; ++i) m = A[i * M + i] < m ? A[i * M + i] : m;
  return m;
  }
}

// ================================================================================================
// LQR backward sweep (lqr.py:16-42).  Symmetric S.  Local constant layout: Aa | Ba | Q | R | Qf
// (+ q | r | P | qf when AFFINE).  One call of step() maps S_{t+1} -> S_t and emits L_t (and l_t, Ht).
template <class DM>
struct LqrC {
  static constexpr int B = DM::B, U = DM::U;
  static constexpr int Aa = 0, Ba = Aa + B * B, Q = Ba + B * U, R = Q + tri(B), Qf = R + tri(U), n = Qf + tri(B);
  static constexpr int q = n, r = q + B, P = r + U, qf = P + U * B, n_affine = qf + B;
  // (global offset, local offset, length) triplets
  static constexpr int NSEG = 5, NSEG_AFF = 9;
  LQGK_HD static void seg(int i, int& g, int& l, int& len) {
    constexpr CLayout c = DM::CL;
    const int G[9] = {c.Aa, c.Ba, c.Q, c.R, c.Qf, c.q, c.r, c.P, c.qf};
    const int Lo[9] = {Aa, Ba, Q, R, Qf, q, r, P, qf};
    const int Le[9] = {B * B, B * U, tri(B), tri(U), tri(B), B, U, U * B, B};
    g = G[i]; l = Lo[i]; len = Le[i];
  }
};

template <class DM, bool AFFINE>
struct LqrFwd {
  static constexpr int B = DM::B, U = DM::U;
  using C = LqrC<DM>;
  // S (full symmetric B x B) in/out; s (B) in/out when AFFINE; L (U x B), l (U), Ht (U x U) out.
  template <class V>
  LQGK_HD static void step(const V& c, double eps, double* S, double* s, double* L, double* l, double* Ht,
                           double& shift) {
    double A[B * B], Bm[B * U], SA[B * B], SB[B * U], H[U * U], G[U * B];
    load_mat<B, B>(c, C::Aa, A);
    load_mat<B, U>(c, C::Ba, Bm);
    mm<B, B, B>(S, A, SA);
    mm<B, B, U>(S, Bm, SB);
    load_sym<U>(c, C::R, H);
    mm_tn_sym<U, B, true>(Bm, SB, H);            // H = R + B^T S B            lqr.py:22
    mm_tn<U, B, B>(Bm, SA, G);                   // G = B^T S A (+ P)          lqr.py:23
    double g[U];
    if (AFFINE) {
      LQGK_UNROLL for (int i = 0; i < U * B; ++i) G[i] += c(C::P + i);
      LQGK_UNROLL for (int i = 0; i < U; ++i) {
        double a = c(C::r + i);
        LQGK_UNROLL for (int k = 0; k < B; ++k) a += Bm[k * U + i] * s[k];   // g = r + B^T s   lqr.py:24
        g[i] = a;
      }
    }
    shift = eps - lambda_min<U>(H);              // lqr.py:27-28
    shift = shift > 0.0 ? shift : 0.0;
    double Lc[U * U];
    LQGK_UNROLL for (int i = 0; i < U * U; ++i) { Lc[i] = H[i]; Ht[i] = H[i]; }
    LQGK_UNROLL for (int i = 0; i < U; ++i) { Lc[i * U + i] += shift; Ht[i * U + i] += shift; }
    chol<U>(Lc);                                  // Ht is SPD after the shift
    double Li[U * U], Hi[U * U];
    tri_inv<U>(Lc, Li);
    mm_tn<U, U, U>(Li, Li, Hi);                   // Ht^-1
    mm<U, U, B>(Hi, G, L);
    LQGK_UNROLL for (int i = 0; i < U * B; ++i) L[i] = -L[i];                 // L = -Ht^-1 G   lqr.py:30
    // S <- Q + A^T S A + L^T H L + L^T G + G^T L      (un-shifted H)       lqr.py:33
    double HL[U * B];
    mm<U, U, B>(H, L, HL);
    LQGK_UNROLL for (int i = 0; i < U * B; ++i) HL[i] += 2.0 * G[i];          // L^T(HL + 2G) sym part
    double Sn[B * B];
    load_sym<B>(c, C::Q, Sn);
    mm_tn_sym<B, B, true>(A, SA, Sn);
    // L^T H L + L^T G + G^T L = sym(L^T (H L + 2 G))
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < U; ++k) a += L[k * B + i] * HL[k * B + j] + L[k * B + j] * HL[k * B + i];
      Sn[i * B + j] += 0.5 * a;
      if (i != j) Sn[j * B + i] += 0.5 * a;
    }
    if (AFFINE) {
      mm<U, U, 1>(Hi, g, l);
      LQGK_UNROLL for (int i = 0; i < U; ++i) l[i] = -l[i];                   // l = -Ht^-1 g   lqr.py:31
      double Hl[U], sn[B];
      mm<U, U, 1>(H, l, Hl);
      LQGK_UNROLL for (int i = 0; i < B; ++i) {                               // lqr.py:34
        double a = c(C::q + i);
        LQGK_UNROLL for (int k = 0; k < B; ++k) a += A[k * B + i] * s[k];
        LQGK_UNROLL for (int k = 0; k < U; ++k) a += G[k * B + i] * l[k] + L[k * B + i] * (Hl[k] + g[k]);
        sn[i] = a;
      }
      LQGK_UNROLL for (int i = 0; i < B; ++i) s[i] = sn[i];
    }
    LQGK_UNROLL for (int i = 0; i < B * B; ++i) S[i] = Sn[i];
  }
};

// ================================================================================================
// Kalman-gain forward sweep (kf.py:6-21).  Local constants: Aa | Fa | VVa | WWa | Sig0.
template <class DM>
struct KfC {
  static constexpr int B = DM::B, Y = DM::Y;
  static constexpr int Aa = 0, Fa = Aa + B * B, VVa = Fa + Y * B, WWa = VVa + tri(B), Sig0 = WWa + tri(Y),
                       n = Sig0 + tri(B);
  static constexpr int NSEG = 5;
  LQGK_HD static void seg(int i, int& g, int& l, int& len) {
    constexpr CLayout c = DM::CL;
    const int G[5] = {c.Aa, c.Fa, c.VVa, c.WWa, c.Sig0};
    const int Lo[5] = {Aa, Fa, VVa, WWa, Sig0};
    const int Le[5] = {B * B, Y * B, tri(B), tri(Y), tri(B)};
    g = G[i]; l = Lo[i]; len = Le[i];
  }
};

template <class DM>
struct KfFwd {
  static constexpr int B = DM::B, Y = DM::Y;
  using C = KfC<DM>;
  // Shared by forward and adjoint: from P (full symmetric) compute Pp, M = F Pp, Gi = Gm^-1, K.
  template <class V>
  LQGK_HD static void gain(const V& c, const double* P, double* Pp, double* M, double* Gi, double* K) {
    double A[B * B], F[Y * B], AP[B * B];
    load_mat<B, B>(c, C::Aa, A);
    load_mat<Y, B>(c, C::Fa, F);
    mm<B, B, B>(A, P, AP);
    load_sym<B>(c, C::VVa, Pp);
    mm_nt_sym<B, B, true>(AP, A, Pp);                // Pp = A P A^T + V V^T        kf.py:10
    mm<Y, B, B>(F, Pp, M);
    double Gm[Y * Y];
    load_sym<Y>(c, C::WWa, Gm);
    mm_nt_sym<Y, B, true>(M, F, Gm);                 // Gm = F Pp F^T + W W^T       kf.py:11
    chol<Y>(Gm);
    double Li[Y * Y];
    tri_inv<Y>(Gm, Li);
    mm_tn_sym<Y, Y>(Li, Li, Gi);
    mm_tn<B, Y, Y>(M, Gi, K);                        // K = Pp F^T Gm^-1            kf.py:12
  }
  template <class V>
  LQGK_HD static void step(const V& c, double* P, double* K) {
    double Pp[B * B], M[Y * B], Gi[Y * Y];
    gain(c, P, Pp, M, Gi, K);
    // P <- (I - K F) Pp = Pp - K M   (symmetric)                               kf.py:14
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = Pp[i * B + j];
      LQGK_UNROLL for (int k = 0; k < Y; ++k) a -= 0.5 * (K[i * Y + k] * M[k * B + j] + K[j * Y + k] * M[k * B + i]);
      P[i * B + j] = a;
      P[j * B + i] = a;
    }
  }
};

// ================================================================================================
// Joint system + covariance pass (system.py:163-212, 223-230), reduced condition-then-predict form.
// Local constants: Ad | Bd | FAd | Aa | Ba | FAa | Dm | N11 | FN | Om.
template <class DM>
struct CovC {
  static constexpr int X = DM::X, B = DM::B, U = DM::U, Y = DM::Y;
  static constexpr int Ad = 0, Bd = Ad + X * X, FAd = Bd + X * U, Aa = FAd + Y * X, Ba = Aa + B * B,
                       FAa = Ba + B * U, Dm = FAa + Y * B, N11 = Dm + Y * U, FN = N11 + tri(X),
                       Om = FN + Y * X, n = Om + tri(Y);
  static constexpr int NSEG = 10;
  LQGK_HD static void seg(int i, int& g, int& l, int& len) {
    constexpr CLayout c = DM::CL;
    const int G[10] = {c.Ad, c.Bd, c.FAd, c.Aa, c.Ba, c.FAa, c.Dm, c.N11, c.FN, c.Om};
    const int Lo[10] = {Ad, Bd, FAd, Aa, Ba, FAa, Dm, N11, FN, Om};
    const int Le[10] = {X * X, X * U, Y * X, B * B, B * U, Y * B, Y * U, tri(X), Y * X, tri(Y)};
    g = G[i]; l = Lo[i]; len = Le[i];
  }
};

template <class DM>
struct CovFwd {
  static constexpr int X = DM::X, B = DM::B, U = DM::U, Y = DM::Y, D = DM::D, N = DM::N, R = DM::R;
  using C = CovC<DM>;

  // Joint transition F_t (N x N, row-major)                                system.py:167-187
  template <class V>
  LQGK_HD static void joint_F(const V& c, const double* L, const double* K, double* Fj) {
    double KD[B * U];
    {
      double Dm[Y * U];
      load_mat<Y, U>(c, C::Dm, Dm);
      mm<B, Y, U>(K, Dm, KD);
    }
    LQGK_UNROLL for (int i = 0; i < X; ++i) {
      LQGK_UNROLL for (int j = 0; j < X; ++j) Fj[i * N + j] = c(C::Ad + i * X + j);
      LQGK_UNROLL for (int j = 0; j < B; ++j) {
        double a = 0.0;
        LQGK_UNROLL for (int k = 0; k < U; ++k) a += c(C::Bd + i * U + k) * L[k * B + j];
        Fj[i * N + X + j] = a;                                        // Bd L
      }
    }
    LQGK_UNROLL for (int i = 0; i < B; ++i) {
      LQGK_UNROLL for (int j = 0; j < X; ++j) {
        double a = 0.0;
        LQGK_UNROLL for (int k = 0; k < Y; ++k) a += K[i * Y + k] * c(C::FAd + k * X + j);
        Fj[(X + i) * N + j] = a;                                      // K Fd Ad
      }
      LQGK_UNROLL for (int j = 0; j < B; ++j) {
        double a = c(C::Aa + i * B + j);
        LQGK_UNROLL for (int k = 0; k < U; ++k) a += (c(C::Ba + i * U + k) + KD[i * U + k]) * L[k * B + j];
        LQGK_UNROLL for (int k = 0; k < Y; ++k) a -= K[i * Y + k] * c(C::FAa + k * B + j);
        Fj[(X + i) * N + X + j] = a;                                  // Aa + Ba L - K Fa Aa + K D L
      }
    }
  }
  // Joint noise covariance N_t = G_t G_t^T (full symmetric N x N)          system.py:190-207
  template <class V>
  LQGK_HD static void joint_N(const V& c, const double* K, double* Nj) {
    double Om[Y * Y], KO[B * Y];
    load_sym<Y>(c, C::Om, Om);
    mm<B, Y, Y>(K, Om, KO);
    LQGK_UNROLL for (int i = 0; i < X; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = c(C::N11 + i * (i + 1) / 2 + j);
      Nj[i * N + j] = a;
      Nj[j * N + i] = a;
    }
    LQGK_UNROLL for (int i = 0; i < B; ++i) {
      LQGK_UNROLL for (int j = 0; j < X; ++j) {
        double a = 0.0;
        LQGK_UNROLL for (int k = 0; k < Y; ++k) a += K[i * Y + k] * c(C::FN + k * X + j);
        Nj[(X + i) * N + j] = a;
        Nj[j * N + X + i] = a;
      }
      LQGK_UNROLL for (int j = 0; j <= i; ++j) {
        double a = 0.0;
        LQGK_UNROLL for (int k = 0; k < Y; ++k) a += KO[i * Y + k] * K[j * Y + k];
        Nj[(X + i) * N + X + j] = a;
        Nj[(X + j) * N + X + i] = a;
      }
    }
  }
  // Condition Sig (full symmetric N x N) on its first D coordinates:
  //   Linv = chol(S)^-1 (lower), logdet = sum log diag chol(S), J = Sig[u,o] S^-1 (R x D),
  //   C = Sig[u,u] - J S J^T (R x R full symmetric).
  LQGK_HD static void condition(const double* Sig, double* Linv, double& logdet, double* J, double* Cn) {
    double Lc[D * D];
    LQGK_UNROLL for (int i = 0; i < D; ++i) LQGK_UNROLL for (int j = 0; j < D; ++j) Lc[i * D + j] = Sig[i * N + j];
    chol<D>(Lc);
    logdet = 0.0;
    LQGK_UNROLL for (int i = 0; i < D; ++i) logdet += log(Lc[i * D + i]);
    tri_inv<D>(Lc, Linv);
    double Z[R * D];                                                 // Z = Sig[u,o] Linv^T
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j < D; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k <= j; ++k) a += Sig[(D + i) * N + k] * Linv[j * D + k];
      Z[i * D + j] = a;
    }
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j < D; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = j; k < D; ++k) a += Z[i * D + k] * Linv[k * D + j];
      J[i * D + j] = a;                                              // J = Z Linv
    }
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = Sig[(D + i) * N + D + j];
      LQGK_UNROLL for (int k = 0; k < D; ++k) a -= Z[i * D + k] * Z[j * D + k];
      Cn[i * R + j] = a;
      Cn[j * R + i] = a;
    }
  }
  // Sig' = F[:,u] C F[:,u]^T + N   (full symmetric N x N)
  LQGK_HD static void predict(const double* Fj, const double* Cm, const double* Nj, double* Sig) {
    LQGK_UNROLL for (int i = 0; i < N; ++i) {
      double t1[R];
      LQGK_UNROLL for (int k = 0; k < R; ++k) {
        double a = 0.0;
        LQGK_UNROLL for (int m = 0; m < R; ++m) a += Fj[i * N + D + m] * Cm[m * R + k];
        t1[k] = a;
      }
      LQGK_UNROLL for (int j = 0; j <= i; ++j) {
        double a = Nj[i * N + j];
        LQGK_UNROLL for (int k = 0; k < R; ++k) a += t1[k] * Fj[j * N + D + k];
        Sig[i * N + j] = a;
        Sig[j * N + i] = a;
      }
    }
  }
  // Initial C_0 from Sig_0 = N_0                                          system.py:211-212
  template <class V>
  LQGK_HD static void init(const V& c, const double* K0, double* Cm) {
    double Nj[N * N], Linv[D * D], J[R * D], ld;
    joint_N(c, K0, Nj);
    condition(Nj, Linv, ld, J, Cm);
  }
  // One step: C_t -> C_{t+1}; emits the float record for the trial kernels through `put(idx, value)`.
  template <class V, class Put>
  LQGK_HD static void step(const V& c, const double* L, const double* K, double* Cm, Put&& put) {
    double Fj[N * N], Nj[N * N], Sig[N * N];
    joint_F(c, L, K, Fj);
    LQGK_UNROLL for (int i = 0; i < N * N; ++i) put(DM::REC_F + i, (float)Fj[i]);
    joint_N(c, K, Nj);
    predict(Fj, Cm, Nj, Sig);
    double Linv[D * D], J[R * D], ld;
    condition(Sig, Linv, ld, J, Cm);
    LQGK_UNROLL for (int i = 0; i < R * D; ++i) put(DM::REC_J + i, (float)J[i]);
    LQGK_UNROLL for (int i = 0; i < D; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j)
      put(DM::REC_LINV + i * (i + 1) / 2 + j, (float)Linv[i * D + j]);
    put(DM::REC_LOGDET, (float)ld);
  }
};

// ================================================================================================
// Per-trial mean / likelihood recursion (system.py:219-221 + MVN log_prob), FP32, one lane per trial.
//   e = x1 - F[o,o] x0 - F[o,u] c ;  c' = F[u,o] x0 + F[u,u] c + J e ;  ll -= 1/2 |Linv e|^2 + logdet + const
// `rec` is the float record of this step (any indexable: shared-memory pointer on the device).
template <class DM>
struct Trial {
  static constexpr int D = DM::D, N = DM::N, R = DM::R;
  static constexpr float HALF_LOG2PI_D = 0.91893853320467274178f * D;

  template <class Rec>
  LQGK_HD static void residual(const Rec& rec, const float* x0, const float* x1, const float* c, float* e) {
    LQGK_UNROLL for (int i = 0; i < D; ++i) {
      float a = x1[i];
      LQGK_UNROLL for (int j = 0; j < D; ++j) a -= rec[DM::REC_F + i * N + j] * x0[j];
      LQGK_UNROLL for (int j = 0; j < R; ++j) a -= rec[DM::REC_F + i * N + D + j] * c[j];
      e[i] = a;
    }
  }
  template <class Rec>
  LQGK_HD static void whiten(const Rec& rec, const float* e, float* z) {   // z = Linv e
    LQGK_UNROLL for (int i = 0; i < D; ++i) {
      float a = 0.f;
      LQGK_UNROLL for (int j = 0; j <= i; ++j) a += rec[DM::REC_LINV + i * (i + 1) / 2 + j] * e[j];
      z[i] = a;
    }
  }
  // forward step; returns this step's log-density term
  template <class Rec>
  LQGK_HD static float fwd(const Rec& rec, const float* x0, const float* x1, float* c) {
    float e[D], z[D], cn[R];
    residual(rec, x0, x1, c, e);
    whiten(rec, e, z);
    float qf = 0.f;
    LQGK_UNROLL for (int i = 0; i < D; ++i) qf += z[i] * z[i];
    LQGK_UNROLL for (int i = 0; i < R; ++i) {
      float a = 0.f;
      LQGK_UNROLL for (int j = 0; j < D; ++j) a += rec[DM::REC_F + (D + i) * N + j] * x0[j];
      LQGK_UNROLL for (int j = 0; j < R; ++j) a += rec[DM::REC_F + (D + i) * N + D + j] * c[j];
      LQGK_UNROLL for (int j = 0; j < D; ++j) a += rec[DM::REC_J + i * D + j] * e[j];
      cn[i] = a;
    }
    LQGK_UNROLL for (int i = 0; i < R; ++i) c[i] = cn[i];
    return -0.5f * qf - rec[DM::REC_LOGDET] - HALF_LOG2PI_D;
  }
  // reverse step for one trial: given c (= c_t), x0 = x_t, x1 = x_{t+1}, weight w and the incoming
  // cotangent cb (of c_{t+1}); produces e, v = S'^-1 e, eb = J^T cb - w v and the outgoing cotangent
  // cbn = F[u,u]^T cb - F[o,u]^T eb (of c_t).  Sums over trials are formed by the caller (sum_term).
  template <class Rec>
  LQGK_HD static void rev(const Rec& rec, const float* x0, const float* x1, const float* c, float w,
                          const float* cb, float* e, float* v, float* eb, float* cbn) {
    float z[D];
    residual(rec, x0, x1, c, e);
    whiten(rec, e, z);
    LQGK_UNROLL for (int i = 0; i < D; ++i) {                         // v = Linv^T z
      float a = 0.f;
      LQGK_UNROLL for (int k = i; k < D; ++k) a += rec[DM::REC_LINV + k * (k + 1) / 2 + i] * z[k];
      v[i] = a;
    }
    LQGK_UNROLL for (int j = 0; j < D; ++j) {                         // eb = J^T cb - w v
      float a = -w * v[j];
      LQGK_UNROLL for (int i = 0; i < R; ++i) a += rec[DM::REC_J + i * D + j] * cb[i];
      eb[j] = a;
    }
    LQGK_UNROLL for (int j = 0; j < R; ++j) {
      float a = 0.f;
      LQGK_UNROLL for (int i = 0; i < R; ++i) a += rec[DM::REC_F + (D + i) * N + D + j] * cb[i];
      LQGK_UNROLL for (int i = 0; i < D; ++i) a -= rec[DM::REC_F + i * N + D + j] * eb[i];
      cbn[j] = a;
    }
  }
  LQGK_HD static constexpr int tri_row(int k) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= k) ++i;
    return i;
  }
  // Term IDX (compile-time) of the per-step sums for one trial (layout DM::SUM_*):
  //   Fb[i][j] = a_i b_j with a = [-eb ; cb], b = [x0 ; c];  Jb[i][j] = cb_i e_j;  Wv[i>=j] = w v_i v_j.
  template <int IDX>
  LQGK_HD static float sum_term(const float* cb, const float* eb, const float* x0, const float* c,
                                const float* e, const float* v, float w) {
    if constexpr (IDX >= DM::NSUM) {
      return 0.f;
    } else if constexpr (IDX < DM::SUM_J) {
      constexpr int i = IDX / N, j = IDX % N;
      float a, b;
      if constexpr (i < D) a = -eb[i]; else a = cb[i - D];
      if constexpr (j < D) b = x0[j]; else b = c[j - D];
      return a * b;
    } else if constexpr (IDX < DM::SUM_W) {
      constexpr int k = IDX - DM::SUM_J;
      return cb[k / D] * e[k % D];
    } else {
      constexpr int k = IDX - DM::SUM_W;
      constexpr int i = tri_row(k);
      constexpr int j = k - i * (i + 1) / 2;
      return w * v[i] * v[j];
    }
  }
};

// ================================================================================================
// Covariance adjoint (one step, t descending).  Local constants as CovC; accumulators (cotangents of the
// derived constants) share CovC's local layout and are updated through `acc(e)` (read-modify-write).
template <class DM>
struct CovRev {
  static constexpr int X = DM::X, B = DM::B, U = DM::U, Y = DM::Y, D = DM::D, N = DM::N, R = DM::R;
  using C = CovC<DM>;
  using F = CovFwd<DM>;

  // Push joint-level cotangents Fb (N x N) and symmetric Nb (N x N) into accumulators and Lb (+=), Kb (+=).
  template <class V, class A>
  LQGK_HD static void joint_bar(const V& c, A&& acc, const double* L, const double* K, const double* Fb,
                                const double* Nb, double* Lb, double* Kb) {
    // --- F blocks: F11 = Fb[:X,:X], F12 = Fb[:X,X:], F21 = Fb[X:,:X], F22 = Fb[X:,X:]
    LQGK_UNROLL for (int i = 0; i < X; ++i) LQGK_UNROLL for (int j = 0; j < X; ++j) acc(C::Ad + i * X + j) += Fb[i * N + j];
    LQGK_UNROLL for (int i = 0; i < X; ++i) LQGK_UNROLL for (int k = 0; k < U; ++k) {   // Bd += F12 L^T
      double a = 0.0;
      LQGK_UNROLL for (int j = 0; j < B; ++j) a += Fb[i * N + X + j] * L[k * B + j];
      acc(C::Bd + i * U + k) += a;
    }
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int j = 0; j < X; ++j) {   // FAd += K^T F21
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += K[i * Y + k] * Fb[(X + i) * N + j];
      acc(C::FAd + k * X + j) += a;
    }
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j < B; ++j) acc(C::Aa + i * B + j) += Fb[(X + i) * N + X + j];
    double F22Lt[B * U];                                                                 // F22 L^T
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int k = 0; k < U; ++k) {
      double a = 0.0;
      LQGK_UNROLL for (int j = 0; j < B; ++j) a += Fb[(X + i) * N + X + j] * L[k * B + j];
      F22Lt[i * U + k] = a;
      acc(C::Ba + i * U + k) += a;
    }
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int j = 0; j < B; ++j) {   // FAa -= K^T F22
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += K[i * Y + k] * Fb[(X + i) * N + X + j];
      acc(C::FAa + k * B + j) -= a;
    }
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int m = 0; m < U; ++m) {   // D += K^T F22 L^T
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += K[i * Y + k] * F22Lt[i * U + m];
      acc(C::Dm + k * U + m) += a;
    }
    // Lb += Bd^T F12 + (Ba + K D)^T F22
    double BKD[B * U];
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int m = 0; m < U; ++m) {
      double a = c(C::Ba + i * U + m);
      LQGK_UNROLL for (int k = 0; k < Y; ++k) a += K[i * Y + k] * c(C::Dm + k * U + m);
      BKD[i * U + m] = a;
    }
    LQGK_UNROLL for (int m = 0; m < U; ++m) LQGK_UNROLL for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < X; ++i) a += c(C::Bd + i * U + m) * Fb[i * N + X + j];
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += BKD[i * U + m] * Fb[(X + i) * N + X + j];
      Lb[m * B + j] += a;
    }
    // Kb += F21 FAd^T - F22 FAa^T + F22 (D L)^T = F21 FAd^T - F22 FAa^T + (F22 L^T) D^T
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int k = 0; k < Y; ++k) {
      double a = 0.0;
      LQGK_UNROLL for (int j = 0; j < X; ++j) a += Fb[(X + i) * N + j] * c(C::FAd + k * X + j);
      LQGK_UNROLL for (int j = 0; j < B; ++j) a -= Fb[(X + i) * N + X + j] * c(C::FAa + k * B + j);
      LQGK_UNROLL for (int m = 0; m < U; ++m) a += F22Lt[i * U + m] * c(C::Dm + k * U + m);
      Kb[i * Y + k] += a;
    }
    // --- N blocks (Nb symmetric): N11 += Nxx ; FN += 2 K^T Nbx ; Om += K^T Nbb K ; Kb += 2 Nbx FN^T + 2 Nbb K Om
    LQGK_UNROLL for (int i = 0; i < X; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) acc(C::N11 + i * (i + 1) / 2 + j) += Nb[i * N + j];
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int j = 0; j < X; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += K[i * Y + k] * Nb[(X + i) * N + j];
      acc(C::FN + k * X + j) += 2.0 * a;
    }
    double NK[B * Y];                                                                    // Nbb K
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int k = 0; k < Y; ++k) {
      double a = 0.0;
      LQGK_UNROLL for (int j = 0; j < B; ++j) a += Nb[(X + i) * N + X + j] * K[j * Y + k];
      NK[i * Y + k] = a;
    }
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int m = 0; m <= k; ++m) {
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += K[i * Y + k] * NK[i * Y + m] + K[i * Y + m] * NK[i * Y + k];
      acc(C::Om + k * (k + 1) / 2 + m) += 0.5 * a;
    }
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int k = 0; k < Y; ++k) {
      double a = 0.0;
      LQGK_UNROLL for (int j = 0; j < X; ++j) a += Nb[(X + i) * N + j] * c(C::FN + k * X + j);
      LQGK_UNROLL for (int m = 0; m < Y; ++m) a += NK[i * Y + m] * c(C::Om + sidx(m, k));
      Kb[i * Y + k] += 2.0 * a;
    }
  }

  // One reverse step.  In: C_t (Cm), L_t, K_t, the float sums of this step via get(idx), sw = sum of trial
  // weights, Cb = cotangent of C_{t+1} (R x R sym).  Out: Cb <- cotangent of C_t; Lb, Kb (overwritten).
  template <class V, class A, class Get>
  LQGK_HD static void step(const V& c, A&& acc, const double* L, const double* K, const double* Cm, Get&& get,
                           double sw, double* Cb, double* Lb, double* Kb) {
    double Fj[N * N], Sgb[N * N];
    F::joint_F(c, L, K, Fj);
    double Linv[D * D], J[R * D];
    {
      double Nj[N * N], Sig[N * N], Cn[R * R], ld;
      F::joint_N(c, K, Nj);
      F::predict(Fj, Cm, Nj, Sig);
      F::condition(Sig, Linv, ld, J, Cn);
    }
    double Sinv[D * D];
    mm_tn_sym<D, D>(Linv, Linv, Sinv);
    // JbS = Jb S^-1 (R x D)
    double JbS[R * D];
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j < D; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < D; ++k) a += (double)get(DM::SUM_J + i * D + k) * Sinv[k * D + j];
      JbS[i * D + j] = a;
    }
    // CbJ = Cb J (R x D)
    double CbJ[R * D];
    mm<R, R, D>(Cb, J, CbJ);
    // Sb = 1/2 Wv - 1/2 sw S^-1 + J^T Cb J - J^T Jb S^-1 ; symmetrised -> Sgb[o,o]
    LQGK_UNROLL for (int i = 0; i < D; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = 0.5 * (double)get(DM::SUM_W + i * (i + 1) / 2 + j) - 0.5 * sw * Sinv[i * D + j];
      LQGK_UNROLL for (int k = 0; k < R; ++k)
        a += J[k * D + i] * CbJ[k * D + j] - 0.5 * (J[k * D + i] * JbS[k * D + j] + J[k * D + j] * JbS[k * D + i]);
      Sgb[i * N + j] = a;
      Sgb[j * N + i] = a;
    }
    // Bb/2 = -Cb J + 1/2 Jb S^-1 -> Sgb[u,o] and its transpose
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j < D; ++j) {
      double a = -CbJ[i * D + j] + 0.5 * JbS[i * D + j];
      Sgb[(D + i) * N + j] = a;
      Sgb[j * N + D + i] = a;
    }
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j < R; ++j) Sgb[(D + i) * N + D + j] = Cb[i * R + j];
    // SF = Sgb F[:,u]  (N x R) ; Fb = sums.Fb ; Fb[:,u] += 2 SF C ; Cb <- F[:,u]^T SF (sym)
    double SF[N * R];
    LQGK_UNROLL for (int i = 0; i < N; ++i) LQGK_UNROLL for (int k = 0; k < R; ++k) {
      double a = 0.0;
      LQGK_UNROLL for (int m = 0; m < N; ++m) a += Sgb[i * N + m] * Fj[m * N + D + k];
      SF[i * R + k] = a;
    }
    double Fb[N * N];
    LQGK_UNROLL for (int i = 0; i < N; ++i) {
      LQGK_UNROLL for (int j = 0; j < D; ++j) Fb[i * N + j] = (double)get(DM::SUM_F + i * N + j);
      LQGK_UNROLL for (int j = 0; j < R; ++j) {
        double a = 0.0;
        LQGK_UNROLL for (int k = 0; k < R; ++k) a += SF[i * R + k] * Cm[k * R + j];
        Fb[i * N + D + j] = (double)get(DM::SUM_F + i * N + D + j) + 2.0 * a;
      }
    }
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int m = 0; m < N; ++m) a += Fj[m * N + D + i] * SF[m * R + j] + Fj[m * N + D + j] * SF[m * R + i];
      Cb[i * R + j] = 0.5 * a;
      Cb[j * R + i] = 0.5 * a;
    }
    LQGK_UNROLL for (int i = 0; i < U * B; ++i) Lb[i] = 0.0;
    LQGK_UNROLL for (int i = 0; i < B * Y; ++i) Kb[i] = 0.0;
    joint_bar(c, acc, L, K, Fb, Sgb, Lb, Kb);
  }
  // After t = 0: cotangent of C_0 = cond(N_0) flows into N_0 (adds to Lb0 (unchanged), Kb0 and accumulators).
  template <class V, class A>
  LQGK_HD static void init_bar(const V& c, A&& acc, const double* L0, const double* K0, const double* Cb, double* Lb,
                               double* Kb) {
    double Nj[N * N], Linv[D * D], J0[R * D], Cn[R * R], ld;
    F::joint_N(c, K0, Nj);
    F::condition(Nj, Linv, ld, J0, Cn);
    double CbJ[R * D], Sgb[N * N], Fb[N * N];
    mm<R, R, D>(Cb, J0, CbJ);
    LQGK_UNROLL for (int i = 0; i < N * N; ++i) Fb[i] = 0.0;
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j < R; ++j) Sgb[(D + i) * N + D + j] = Cb[i * R + j];
    LQGK_UNROLL for (int i = 0; i < R; ++i) LQGK_UNROLL for (int j = 0; j < D; ++j) {
      Sgb[(D + i) * N + j] = -CbJ[i * D + j];
      Sgb[j * N + D + i] = -CbJ[i * D + j];
    }
    LQGK_UNROLL for (int i = 0; i < D; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < R; ++k) a += J0[k * D + i] * CbJ[k * D + j];
      Sgb[i * N + j] = a;
      Sgb[j * N + i] = a;
    }
    joint_bar(c, acc, L0, K0, Fb, Sgb, Lb, Kb);
  }
};

// ================================================================================================
// Kalman-gain adjoint (one step, t descending).  Accumulators share KfC's local layout.
template <class DM>
struct KfRev {
  static constexpr int B = DM::B, Y = DM::Y;
  using C = KfC<DM>;
  // In: P_t (full sym), Kb_t (B x Y), Pnb = cotangent of P_{t+1} (sym).  Out: Pnb <- cotangent of P_t.
  template <class V, class A>
  LQGK_HD static void step(const V& c, A&& acc, const double* P, const double* Kb, double* Pnb) {
    double Pp[B * B], M[Y * B], Gi[Y * Y], K[B * Y];
    KfFwd<DM>::gain(c, P, Pp, M, Gi, K);
    double Am[B * B], F[Y * B];
    load_mat<B, B>(c, C::Aa, Am);
    load_mat<Y, B>(c, C::Fa, F);
    // Ktot = Kb - Pnb M^T ; Y = Ktot Gi
    double Ktot[B * Y], Yv[B * Y];
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int k = 0; k < Y; ++k) {
      double a = Kb[i * Y + k];
      LQGK_UNROLL for (int j = 0; j < B; ++j) a -= Pnb[i * B + j] * M[k * B + j];
      Ktot[i * Y + k] = a;
    }
    mm<B, Y, Y>(Ktot, Gi, Yv);
    // Mb = -K^T Pnb + Y^T  (Y x B)
    double Mb[Y * B];
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int j = 0; j < B; ++j) {
      double a = Yv[j * Y + k];
      LQGK_UNROLL for (int i = 0; i < B; ++i) a -= K[i * Y + k] * Pnb[i * B + j];
      Mb[k * B + j] = a;
    }
    // Gmb = -sym(K^T Y)  (Y x Y)
    double Gmb[Y * Y];
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int m = 0; m <= k; ++m) {
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += K[i * Y + k] * Yv[i * Y + m] + K[i * Y + m] * Yv[i * Y + k];
      Gmb[k * Y + m] = -0.5 * a;
      Gmb[m * Y + k] = -0.5 * a;
      acc(C::WWa + k * (k + 1) / 2 + m) += -0.5 * a;
    }
    // Fa += Mb Pp + 2 Gmb M
    LQGK_UNROLL for (int k = 0; k < Y; ++k) LQGK_UNROLL for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int i = 0; i < B; ++i) a += Mb[k * B + i] * Pp[i * B + j];
      LQGK_UNROLL for (int m = 0; m < Y; ++m) a += 2.0 * Gmb[k * Y + m] * M[m * B + j];
      acc(C::Fa + k * B + j) += a;
    }
    // Ppb = Pnb + sym(F^T Mb) + F^T Gmb F
    double GF[Y * B], Ppb[B * B];
    mm<Y, Y, B>(Gmb, F, GF);
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = Pnb[i * B + j];
      LQGK_UNROLL for (int k = 0; k < Y; ++k)
        a += 0.5 * (F[k * B + i] * Mb[k * B + j] + F[k * B + j] * Mb[k * B + i]) + F[k * B + i] * GF[k * B + j];
      Ppb[i * B + j] = a;
      Ppb[j * B + i] = a;
      acc(C::VVa + i * (i + 1) / 2 + j) += a;
    }
    // Aa += 2 Ppb A P ; Pnb <- A^T Ppb A
    double PA[B * B], AP[B * B];
    mm<B, B, B>(Ppb, Am, PA);                       // Ppb A
    mm<B, B, B>(Am, P, AP);                         // A P
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < B; ++k) a += Ppb[i * B + k] * AP[k * B + j];
      acc(C::Aa + i * B + j) += 2.0 * a;
    }
    mm_tn_sym<B, B>(Am, PA, Pnb);
  }
  // After t = 0: P_0 = Sig0
  template <class A>
  LQGK_HD static void finish(A&& acc, const double* Pnb) {
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) acc(C::Sig0 + i * (i + 1) / 2 + j) += Pnb[i * B + j];
  }
};

// ================================================================================================
// Riccati adjoint (one step, t ascending; eigen-shift treated as constant).  Accumulators share LqrC's layout.
template <class DM>
struct LqrRev {
  static constexpr int B = DM::B, U = DM::U;
  using C = LqrC<DM>;
  // In: S_{t+1} (full sym), L_t, Lb_t (cotangent of L_t from the covariance adjoint), shift_t,
  //     Sn = cotangent of S_t (sym).  Out: Sn <- cotangent of S_{t+1}.
  template <class V, class A>
  LQGK_HD static void step(const V& c, A&& acc, const double* S, const double* L, const double* Lbar, double shift,
                           double* Sn) {
    double Am[B * B], Bm[B * U], SA[B * B], SB[B * U], H[U * U], G[U * B];
    load_mat<B, B>(c, C::Aa, Am);
    load_mat<B, U>(c, C::Ba, Bm);
    mm<B, B, B>(S, Am, SA);
    mm<B, B, U>(S, Bm, SB);
    load_sym<U>(c, C::R, H);
    mm_tn_sym<U, B, true>(Bm, SB, H);
    mm_tn<U, B, B>(Bm, SA, G);
    double Hi[U * U];
    {
      double Lc[U * U], Li[U * U];
      LQGK_UNROLL for (int i = 0; i < U * U; ++i) Lc[i] = H[i];
      LQGK_UNROLL for (int i = 0; i < U; ++i) Lc[i * U + i] += shift;
      chol<U>(Lc);
      tri_inv<U>(Lc, Li);
      mm_tn_sym<U, U>(Li, Li, Hi);
    }
    // Q += Sn ; Aa += 2 SA Sn
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) acc(C::Q + i * (i + 1) / 2 + j) += Sn[i * B + j];
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < B; ++k) a += SA[i * B + k] * Sn[k * B + j];
      acc(C::Aa + i * B + j) += 2.0 * a;
    }
    // LS = L Sn (U x B) ; Lb = Lbar + 2 (H L + G) Sn ; Hb = LS L^T ; Gb = 2 LS
    double LS[U * B], HLG[U * B], Lb[U * B], Hb[U * U], Gb[U * B];
    mm<U, B, B>(L, Sn, LS);
    mm<U, U, B>(H, L, HLG);
    LQGK_UNROLL for (int i = 0; i < U * B; ++i) HLG[i] += G[i];
    mm<U, B, B>(HLG, Sn, Lb);
    LQGK_UNROLL for (int i = 0; i < U * B; ++i) Lb[i] = Lbar[i] + 2.0 * Lb[i];
    mm_nt<U, B, U>(LS, L, Hb);
    // HiLb = Ht^-1 Lb ; Gb = 2 LS - HiLb ; Hb -= HiLb L^T ; Hb <- sym
    double HiLb[U * B];
    mm<U, U, B>(Hi, Lb, HiLb);
    LQGK_UNROLL for (int i = 0; i < U * B; ++i) Gb[i] = 2.0 * LS[i] - HiLb[i];
    LQGK_UNROLL for (int i = 0; i < U; ++i) LQGK_UNROLL for (int j = 0; j < U; ++j) {
      double a = Hb[i * U + j];
      LQGK_UNROLL for (int k = 0; k < B; ++k) a -= HiLb[i * B + k] * L[j * B + k];
      Hb[i * U + j] = a;
    }
    symmetrize<U>(Hb);
    LQGK_UNROLL for (int i = 0; i < U; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) acc(C::R + i * (i + 1) / 2 + j) += Hb[i * U + j];
    // Ba += 2 SB Hb + SA Gb^T ; Aa += SB Gb
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int m = 0; m < U; ++m) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < U; ++k) a += 2.0 * SB[i * U + k] * Hb[k * U + m];
      LQGK_UNROLL for (int k = 0; k < B; ++k) a += SA[i * B + k] * Gb[m * B + k];
      acc(C::Ba + i * U + m) += a;
    }
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < U; ++k) a += SB[i * U + k] * Gb[k * B + j];
      acc(C::Aa + i * B + j) += a;
    }
    // Sn <- A Sn A^T + B Hb B^T + sym(B Gb A^T)
    double AS[B * B], BH[B * U], BG[B * B], Sb[B * B];
    mm<B, B, B>(Am, Sn, AS);
    mm_nt_sym<B, B>(AS, Am, Sb);
    mm<B, U, U>(Bm, Hb, BH);
    mm_nt_sym<B, U, true>(BH, Bm, Sb);
    mm<B, U, B>(Bm, Gb, BG);                                       // B Gb (B x B), then (B Gb) A^T
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) {
      double a = 0.0;
      LQGK_UNROLL for (int k = 0; k < B; ++k) a += BG[i * B + k] * Am[j * B + k] + BG[j * B + k] * Am[i * B + k];
      Sb[i * B + j] += 0.5 * a;
      if (i != j) Sb[j * B + i] += 0.5 * a;
    }
    LQGK_UNROLL for (int i = 0; i < B * B; ++i) Sn[i] = Sb[i];
  }
  template <class A>
  LQGK_HD static void finish(A&& acc, const double* Sn) {   // S_T = Qf
    LQGK_UNROLL for (int i = 0; i < B; ++i) LQGK_UNROLL for (int j = 0; j <= i; ++j) acc(C::Qf + i * (i + 1) / 2 + j) += Sn[i * B + j];
  }
};

}  // namespace lqgk
