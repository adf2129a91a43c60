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
// Loops of the per-sample FP64 recursions: fully unrolled (matrices in registers) for the small systems; systems
// compiled with -DLQGK_BIG (joint dim > 12, e.g. the 24-dim delayed point-mass model of config c4) keep them rolled --
// their matrices live in local memory and a fully unrolled body would neither fit registers nor compile in minutes.
#if defined(__CUDACC__) && !defined(LQGK_BIG)
#define LQGK_UNROLL64 _Pragma("unroll")
#elif defined(__CUDACC__) && defined(LQGK_UNROLL64_FORCE_1)
#define LQGK_UNROLL64 _Pragma("unroll 1")
#elif defined(__CUDACC__)
// No pragma at all (the compiler's own heuristics): with `#pragma unroll 1` on these loop nests the NVVM optimiser of
// CUDA 12.9 (-O3; correct at -Xcicc -O1) miscompiled cov_fwd_body -- device results differed from the host run of the
// same code (tools/repro/nvvm_unroll1_cov_fwd.cu reproduces it with -DLQGK_UNROLL64_FORCE_1) -- while the un-annotated loops compile correctly.
#define LQGK_UNROLL64
#else
#define LQGK_UNROLL64
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
  // (SUM_J is 16-byte aligned so the [SUM_J, SUMP) tail of a row can be fetched with one bulk copy)
  static constexpr int SUM_F = 0, SUM_J = round4(N * N), SUM_W = SUM_J + R * D, NSUM = SUM_W + tri(D);
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
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j < N; ++j) {
    TC acc = ACC ? C[i * N + j] : TC(0);
    LQGK_UNROLL64 for (int k = 0; k < K; ++k) acc += A[i * K + k] * Bm[k * N + j];
    C[i * N + j] = acc;
  }
}
template <int M, int K, int N, bool ACC = false, class TA, class TB, class TC>
LQGK_HD void mm_nt(const TA* A, const TB* Bm, TC* C) {  // C[M,N] (+)= A[M,K] B[N,K]^T
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j < N; ++j) {
    TC acc = ACC ? C[i * N + j] : TC(0);
    LQGK_UNROLL64 for (int k = 0; k < K; ++k) acc += A[i * K + k] * Bm[j * K + k];
    C[i * N + j] = acc;
  }
}
template <int M, int K, int N, bool ACC = false, class TA, class TB, class TC>
LQGK_HD void mm_tn(const TA* A, const TB* Bm, TC* C) {  // C[M,N] (+)= A[K,M]^T B[K,N]
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j < N; ++j) {
    TC acc = ACC ? C[i * N + j] : TC(0);
    LQGK_UNROLL64 for (int k = 0; k < K; ++k) acc += A[k * M + i] * Bm[k * N + j];
    C[i * N + j] = acc;
  }
}
// C[M,M] (+)= A[M,K] B[M,K]^T where the result is known symmetric: lower triangle computed, mirrored.
template <int M, int K, bool ACC = false>
LQGK_HD void mm_nt_sym(const double* A, const double* Bm, double* C) {
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
    double acc = ACC ? C[i * M + j] : 0.0;
    LQGK_UNROLL64 for (int k = 0; k < K; ++k) acc += A[i * K + k] * Bm[j * K + k];
    C[i * M + j] = acc;
    C[j * M + i] = acc;
  }
}
// C[M,M] (+)= A[K,M]^T B[K,M], result known symmetric.
template <int M, int K, bool ACC = false>
LQGK_HD void mm_tn_sym(const double* A, const double* Bm, double* C) {
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
    double acc = ACC ? C[i * M + j] : 0.0;
    LQGK_UNROLL64 for (int k = 0; k < K; ++k) acc += A[k * M + i] * Bm[k * M + j];
    C[i * M + j] = acc;
    C[j * M + i] = acc;
  }
}
template <int M>
LQGK_HD void symmetrize(double* C) {  // C <- (C + C^T)/2
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j < i; ++j) {
    double v = 0.5 * (C[i * M + j] + C[j * M + i]);
    C[i * M + j] = v;
    C[j * M + i] = v;
  }
}
template <int M, class V>
LQGK_HD void load_sym(const V& v, int off, double* C) {  // packed lower -> full
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
    double a = v(off + i * (i + 1) / 2 + j);
    C[i * M + j] = a;
    C[j * M + i] = a;
  }
}
template <int M, int N, class V>
LQGK_HD void load_mat(const V& v, int off, double* C) {
  LQGK_UNROLL64 for (int i = 0; i < M * N; ++i) C[i] = v(off + i);
}
// In-place Cholesky of a symmetric positive definite M x M matrix (lower factor; upper part untouched).
template <int M>
LQGK_HD void chol(double* A) {
  LQGK_UNROLL64 for (int j = 0; j < M; ++j) {
    double djj = A[j * M + j];
    LQGK_UNROLL64 for (int k = 0; k < j; ++k) djj -= A[j * M + k] * A[j * M + k];
    djj = sqrt(djj);
    A[j * M + j] = djj;
    double inv = 1.0 / djj;
    LQGK_UNROLL64 for (int i = j + 1; i < M; ++i) {
      double v = A[i * M + j];
      LQGK_UNROLL64 for (int k = 0; k < j; ++k) v -= A[i * M + k] * A[j * M + k];
      A[i * M + j] = v * inv;
    }
  }
}
// Inverse of a lower-triangular matrix (lower part of Lc) into Li (lower part; upper part zeroed).
template <int M>
LQGK_HD void tri_inv(const double* Lc, double* Li) {
  LQGK_UNROLL64 for (int i = 0; i < M * M; ++i) Li[i] = 0.0;
  LQGK_UNROLL64 for (int j = 0; j < M; ++j) {
    Li[j * M + j] = 1.0 / Lc[j * M + j];
    LQGK_UNROLL64 for (int i = j + 1; i < M; ++i) {
      double v = 0.0;
      LQGK_UNROLL64 for (int k = j; k < i; ++k) v -= Lc[i * M + k] * Li[k * M + j];
      Li[i * M + j] = v / Lc[i * M + i];
    }
  }
}
// Cholesky factor (in place, lower) AND its inverse in one pass.  The diagonal uses one reciprocal square root per column
// (rsqrt on the device) and every later division becomes a multiplication by that reciprocal: the double-precision sqrt /
// divide / log sequences are the longest links of the per-step dependency chains that bound the sequential kernels.
// `half_logdet` (optional): sum_i log L_ii, from a single log of the product of the reciprocals (M <= 4: no range issue).
LQGK_HD double rsqrt_f64(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
template <int M>
LQGK_HD void chol_and_inverse(double* A, double* Li, double* half_logdet = nullptr) {
  double rinv[M];
  LQGK_UNROLL64 for (int j = 0; j < M; ++j) {
    double djj = A[j * M + j];
    LQGK_UNROLL64 for (int k = 0; k < j; ++k) djj -= A[j * M + k] * A[j * M + k];
    const double r = rsqrt_f64(djj);
    rinv[j] = r;
    A[j * M + j] = djj * r;
    LQGK_UNROLL64 for (int i = j + 1; i < M; ++i) {
      double v = A[i * M + j];
      LQGK_UNROLL64 for (int k = 0; k < j; ++k) v -= A[i * M + k] * A[j * M + k];
      A[i * M + j] = v * r;
    }
  }
  LQGK_UNROLL64 for (int i = 0; i < M * M; ++i) Li[i] = 0.0;
  LQGK_UNROLL64 for (int j = 0; j < M; ++j) {
    Li[j * M + j] = rinv[j];
    LQGK_UNROLL64 for (int i = j + 1; i < M; ++i) {
      double v = 0.0;
      LQGK_UNROLL64 for (int k = j; k < i; ++k) v -= A[i * M + k] * Li[k * M + j];
      Li[i * M + j] = v * rinv[i];
    }
  }
  if (half_logdet) {
    double prod = 1.0;
    LQGK_UNROLL64 for (int j = 0; j < M; ++j) prod *= rinv[j];
    *half_logdet = -log(prod);
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
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) G[i] += c(C::P + i);
      LQGK_UNROLL64 for (int i = 0; i < U; ++i) {
        double a = c(C::r + i);
        LQGK_UNROLL64 for (int k = 0; k < B; ++k) a += Bm[k * U + i] * s[k];   // g = r + B^T s   lqr.py:24
        g[i] = a;
      }
    }
    shift = eps - lambda_min<U>(H);              // lqr.py:27-28
    shift = shift > 0.0 ? shift : 0.0;
    double Lc[U * U];
    LQGK_UNROLL64 for (int i = 0; i < U * U; ++i) { Lc[i] = H[i]; Ht[i] = H[i]; }
    LQGK_UNROLL64 for (int i = 0; i < U; ++i) { Lc[i * U + i] += shift; Ht[i * U + i] += shift; }
    double Li[U * U], Hi[U * U];
    chol_and_inverse<U>(Lc, Li);                  // Ht is SPD after the shift
    mm_tn<U, U, U>(Li, Li, Hi);                   // Ht^-1
    mm<U, U, B>(Hi, G, L);
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) L[i] = -L[i];                 // L = -Ht^-1 G   lqr.py:30
    // S <- Q + A^T S A + L^T H L + L^T G + G^T L      (un-shifted H)       lqr.py:33
    double HL[U * B];
    mm<U, U, B>(H, L, HL);
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) HL[i] += 2.0 * G[i];          // L^T(HL + 2G) sym part
    double Sn[B * B];
    load_sym<B>(c, C::Q, Sn);
    mm_tn_sym<B, B, true>(A, SA, Sn);
    // L^T H L + L^T G + G^T L = sym(L^T (H L + 2 G))
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < U; ++k) a += L[k * B + i] * HL[k * B + j] + L[k * B + j] * HL[k * B + i];
      Sn[i * B + j] += 0.5 * a;
      if (i != j) Sn[j * B + i] += 0.5 * a;
    }
    if (AFFINE) {
      mm<U, U, 1>(Hi, g, l);
      LQGK_UNROLL64 for (int i = 0; i < U; ++i) l[i] = -l[i];                   // l = -Ht^-1 g   lqr.py:31
      double Hl[U], sn[B];
      mm<U, U, 1>(H, l, Hl);
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) {                               // lqr.py:34
        double a = c(C::q + i);
        LQGK_UNROLL64 for (int k = 0; k < B; ++k) a += A[k * B + i] * s[k];
        LQGK_UNROLL64 for (int k = 0; k < U; ++k) a += G[k * B + i] * l[k] + L[k * B + i] * (Hl[k] + g[k]);
        sn[i] = a;
      }
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) s[i] = sn[i];
    }
    LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) S[i] = Sn[i];
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
    double Li[Y * Y];
    chol_and_inverse<Y>(Gm, Li);
    mm_tn_sym<Y, Y>(Li, Li, Gi);
    mm_tn<B, Y, Y>(M, Gi, K);                        // K = Pp F^T Gm^-1            kf.py:12
  }
  template <class V>
  LQGK_HD static void step(const V& c, double* P, double* K) {
    double Pp[B * B], M[Y * B], Gi[Y * Y];
    gain(c, P, Pp, M, Gi, K);
    // P <- (I - K F) Pp = Pp - K M   (symmetric)                               kf.py:14
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
      double a = Pp[i * B + j];
      LQGK_UNROLL64 for (int k = 0; k < Y; ++k) a -= 0.5 * (K[i * Y + k] * M[k * B + j] + K[j * Y + k] * M[k * B + i]);
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
    LQGK_UNROLL64 for (int i = 0; i < X; ++i) {
      LQGK_UNROLL64 for (int j = 0; j < X; ++j) Fj[i * N + j] = c(C::Ad + i * X + j);
      LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double a = 0.0;
        LQGK_UNROLL64 for (int k = 0; k < U; ++k) a += c(C::Bd + i * U + k) * L[k * B + j];
        Fj[i * N + X + j] = a;                                        // Bd L
      }
    }
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) {
      LQGK_UNROLL64 for (int j = 0; j < X; ++j) {
        double a = 0.0;
        LQGK_UNROLL64 for (int k = 0; k < Y; ++k) a += K[i * Y + k] * c(C::FAd + k * X + j);
        Fj[(X + i) * N + j] = a;                                      // K Fd Ad
      }
      LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
        double a = c(C::Aa + i * B + j);
        LQGK_UNROLL64 for (int k = 0; k < U; ++k) a += (c(C::Ba + i * U + k) + KD[i * U + k]) * L[k * B + j];
        LQGK_UNROLL64 for (int k = 0; k < Y; ++k) a -= K[i * Y + k] * c(C::FAa + k * B + j);
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
    LQGK_UNROLL64 for (int i = 0; i < X; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
      double a = c(C::N11 + i * (i + 1) / 2 + j);
      Nj[i * N + j] = a;
      Nj[j * N + i] = a;
    }
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) {
      LQGK_UNROLL64 for (int j = 0; j < X; ++j) {
        double a = 0.0;
        LQGK_UNROLL64 for (int k = 0; k < Y; ++k) a += K[i * Y + k] * c(C::FN + k * X + j);
        Nj[(X + i) * N + j] = a;
        Nj[j * N + X + i] = a;
      }
      LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
        double a = 0.0;
        LQGK_UNROLL64 for (int k = 0; k < Y; ++k) a += KO[i * Y + k] * K[j * Y + k];
        Nj[(X + i) * N + X + j] = a;
        Nj[(X + j) * N + X + i] = a;
      }
    }
  }
  // Condition Sig (full symmetric N x N) on its first D coordinates:
  //   Linv = chol(S)^-1 (lower), logdet = sum log diag chol(S), J = Sig[u,o] S^-1 (R x D, handed element by element to
  //   emit_j(index, value)), C = Sig[u,u] - J S J^T (R x R full symmetric).
  // J is emitted rather than returned in an array: with the rolled loops of the large-system build nvcc 12.9 gave a
  // caller-side J array the same local-memory slot as Cn (the record then held rows of C instead of J).
  template <class EmitJ>
  LQGK_HD static void condition(const double* Sig, double* Linv, double& logdet, EmitJ&& emit_j, double* Cn) {
    double Lc[D * D];
    LQGK_UNROLL64 for (int i = 0; i < D; ++i) LQGK_UNROLL64 for (int j = 0; j < D; ++j) Lc[i * D + j] = Sig[i * N + j];
    chol_and_inverse<D>(Lc, Linv, &logdet);
    double Z[R * D];                                                 // Z = Sig[u,o] Linv^T
    LQGK_UNROLL64 for (int i = 0; i < R; ++i) LQGK_UNROLL64 for (int j = 0; j < D; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k <= j; ++k) a += Sig[(D + i) * N + k] * Linv[j * D + k];
      Z[i * D + j] = a;
    }
    LQGK_UNROLL64 for (int i = 0; i < R; ++i) LQGK_UNROLL64 for (int j = 0; j < D; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = j; k < D; ++k) a += Z[i * D + k] * Linv[k * D + j];
      emit_j(i * D + j, a);                                          // J = Z Linv
    }
    LQGK_UNROLL64 for (int i = 0; i < R; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
      double a = Sig[(D + i) * N + D + j];
      LQGK_UNROLL64 for (int k = 0; k < D; ++k) a -= Z[i * D + k] * Z[j * D + k];
      Cn[i * R + j] = a;
      Cn[j * R + i] = a;
    }
  }
  // Sig' = F[:,u] C F[:,u]^T + N   (full symmetric N x N)
  LQGK_HD static void predict(const double* Fj, const double* Cm, const double* Nj, double* Sig) {
    LQGK_UNROLL64 for (int i = 0; i < N; ++i) {
      double t1[R];
      LQGK_UNROLL64 for (int k = 0; k < R; ++k) {
        double a = 0.0;
        LQGK_UNROLL64 for (int m = 0; m < R; ++m) a += Fj[i * N + D + m] * Cm[m * R + k];
        t1[k] = a;
      }
      LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
        double a = Nj[i * N + j];
        LQGK_UNROLL64 for (int k = 0; k < R; ++k) a += t1[k] * Fj[j * N + D + k];
        Sig[i * N + j] = a;
        Sig[j * N + i] = a;
      }
    }
  }
  // Initial C_0 from Sig_0 = N_0                                          system.py:211-212
  template <class V>
  LQGK_HD static void init(const V& c, const double* K0, double* Cm, double* J0) {
    double Nj[N * N], Linv[D * D], ld;
    joint_N(c, K0, Nj);
    condition(Nj, Linv, ld, [&](int i, double v) { J0[i] = v; }, Cm);
  }
  // One step: C_t -> C_{t+1}; emits the float record for the trial kernels through `put(idx, value)`.
  // save(which, e, v): optional FP64 outputs for the adjoint: which 0 = Fu_t (N x R row-major), 1 = J_t then S'^-1_t.
  struct NoSig {
    LQGK_HD void operator()(int, double) const {}
  };
  // savesig(e, v) (optional): the full predictive joint covariance Sigma'[e] of this step (system.py:223-230), row-major N x N
  // -- what conditional_moments / belief_tracking_distribution return (moments entry point only).
  template <class V, class Put, class Save, class SaveSig = NoSig>
  LQGK_HD static void step(const V& c, const double* L, const double* K, double* Cm, Put&& put, Save&& save, SaveSig&& savesig = SaveSig{}) {
    double Fj[N * N], Nj[N * N], Sig[N * N];
    joint_F(c, L, K, Fj);
    // the observed rows are stored negated so the trial kernels form e = x1 + (-F[o,:]) [x0; c] with plain FMAs, starting
    // from x1 (x1 - x0 is exact when A_d = I, which keeps the residual accurate in FP32)
    LQGK_UNROLL64 for (int i = 0; i < N * N; ++i) put(DM::REC_F + i, (float)(i < D * N ? -Fj[i] : Fj[i]));
    LQGK_UNROLL64 for (int i = 0; i < N; ++i) LQGK_UNROLL64 for (int j = 0; j < R; ++j) save(0, i * R + j, Fj[i * N + D + j]);
    joint_N(c, K, Nj);
    predict(Fj, Cm, Nj, Sig);
    LQGK_UNROLL64 for (int i = 0; i < N * N; ++i) savesig(i, Sig[i]);
    double Linv[D * D], ld;
    condition(Sig, Linv, ld, [&](int i, double v) { put(DM::REC_J + i, (float)v); save(1, i, v); }, Cm);
    LQGK_UNROLL64 for (int i = 0; i < D; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
      put(DM::REC_LINV + i * (i + 1) / 2 + j, (float)Linv[i * D + j]);
      double sv = 0.0;
      LQGK_UNROLL64 for (int k = i; k < D; ++k) sv += Linv[k * D + i] * Linv[k * D + j];   // (Linv^T Linv)[i][j], i >= j
      save(1, R * D + i * (i + 1) / 2 + j, sv);
    }
    put(DM::REC_LOGDET, (float)ld);
  }
};

// ================================================================================================
// Scalar types of the per-trial recursions: `float` (one trial) or `f32x2` (two trials packed in one 64-bit
// register pair, executed by Blackwell's packed FP32 instructions fma/mul/add.rn.f32x2 -- one issue slot for two
// FMAs; device only).
struct f32x2 {
  float x, y;
};
template <class S>
struct Ops;
template <>
struct Ops<float> {
  LQGK_HD static float bc(float r) { return r; }
  LQGK_HD static float zero() { return 0.f; }
  LQGK_HD static float fma(float a, float b, float c) { return fmaf(a, b, c); }
  LQGK_HD static float mul(float a, float b) { return a * b; }
  LQGK_HD static float add(float a, float b) { return a + b; }
  LQGK_HD static float sub(float a, float b) { return a - b; }
  LQGK_HD static float neg(float a) { return -a; }
  LQGK_HD static float hsum(float a) { return a; }
};
#if defined(__CUDACC__)
template <>
struct Ops<f32x2> {
  using U = unsigned long long;
  __device__ __forceinline__ static U u(const f32x2& a) { return *reinterpret_cast<const U*>(&a); }
  __device__ __forceinline__ static f32x2 f(U v) { return *reinterpret_cast<f32x2*>(&v); }
  __device__ __forceinline__ static f32x2 bc(float r) { return f32x2{r, r}; }
  __device__ __forceinline__ static f32x2 zero() { return f32x2{0.f, 0.f}; }
  __device__ __forceinline__ static f32x2 fma(f32x2 a, f32x2 b, f32x2 c) {
    U d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(u(a)), "l"(u(b)), "l"(u(c)));
    return f(d);
  }
  __device__ __forceinline__ static f32x2 mul(f32x2 a, f32x2 b) {
    U d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(u(a)), "l"(u(b)));
    return f(d);
  }
  __device__ __forceinline__ static f32x2 add(f32x2 a, f32x2 b) {
    U d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(u(a)), "l"(u(b)));
    return f(d);
  }
  __device__ __forceinline__ static f32x2 sub(f32x2 a, f32x2 b) {
    U d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(u(a)), "l"(u(b)));
    return f(d);
  }
  __device__ __forceinline__ static f32x2 neg(f32x2 a) { return f32x2{-a.x, -a.y}; }
  __device__ __forceinline__ static float hsum(f32x2 a) { return a.x + a.y; }
};
#endif

// ================================================================================================
// Per-trial mean / likelihood recursion (system.py:219-221 + MVN log_prob), FP32, one lane per trial (S = float) or
// per pair of trials (S = f32x2).
//   e = x1 - F[o,o] x0 - F[o,u] c ;  c' = F[u,o] x0 + F[u,u] c + J e ;  ll -= 1/2 |Linv e|^2 + logdet + const
// `rec` is the float record of this step (any indexable: shared-memory pointer or register copy).
template <class DM>
struct Trial {
  static constexpr int D = DM::D, N = DM::N, R = DM::R;
  static constexpr float HALF_LOG2PI_D = 0.91893853320467274178f * D;

  template <class S, class Rec>
  LQGK_HD static void residual(const Rec& rec, const S* x0, const S* x1, const S* c, S* e) {
    using O = Ops<S>;
    LQGK_UNROLL for (int i = 0; i < D; ++i) {                         // record rows o hold -F[o,:]
      S a = x1[i];
      LQGK_UNROLL for (int j = 0; j < D; ++j) a = O::fma(O::bc(rec[DM::REC_F + i * N + j]), x0[j], a);
      LQGK_UNROLL for (int j = 0; j < R; ++j) a = O::fma(O::bc(rec[DM::REC_F + i * N + D + j]), c[j], a);
      e[i] = a;
    }
  }
  template <class S, class Rec>
  LQGK_HD static void whiten(const Rec& rec, const S* e, S* z) {   // z = Linv e
    using O = Ops<S>;
    LQGK_UNROLL for (int i = 0; i < D; ++i) {
      S a = O::zero();
      LQGK_UNROLL for (int j = 0; j <= i; ++j) a = O::fma(O::bc(rec[DM::REC_LINV + i * (i + 1) / 2 + j]), e[j], a);
      z[i] = a;
    }
  }
  // c <- F[u,o] x0 + F[u,u] c + J e   (the carried conditional mean of the unobserved block)
  template <class S, class Rec>
  LQGK_HD static void update(const Rec& rec, const S* x0, const S* e, S* c) {
    using O = Ops<S>;
    S cn[R];
    LQGK_UNROLL for (int i = 0; i < R; ++i) {
      S a = O::zero();
      LQGK_UNROLL for (int j = 0; j < D; ++j) a = O::fma(O::bc(rec[DM::REC_F + (D + i) * N + j]), x0[j], a);
      LQGK_UNROLL for (int j = 0; j < R; ++j) a = O::fma(O::bc(rec[DM::REC_F + (D + i) * N + D + j]), c[j], a);
      LQGK_UNROLL for (int j = 0; j < D; ++j) a = O::fma(O::bc(rec[DM::REC_J + i * D + j]), e[j], a);
      cn[i] = a;
    }
    LQGK_UNROLL for (int i = 0; i < R; ++i) c[i] = cn[i];
  }
  // forward step; returns this step's log-density term
  template <class S, class Rec>
  LQGK_HD static S fwd(const Rec& rec, const S* x0, const S* x1, S* c) {
    using O = Ops<S>;
    S e[D], z[D];
    residual<S>(rec, x0, x1, c, e);
    whiten<S>(rec, e, z);
    S qf = O::zero();
    LQGK_UNROLL for (int i = 0; i < D; ++i) qf = O::fma(z[i], z[i], qf);
    update<S>(rec, x0, e, c);
    return O::fma(O::bc(-0.5f), qf, O::bc(-rec[DM::REC_LOGDET] - HALF_LOG2PI_D));
  }
  // predictive mean of the joint state (system.py:219-221): mu'[o] = x1 - e, mu'[u] = F[u,o] x0 + F[u,u] c (before the
  // correction J e that conditions on x1); then advances c like fwd().  `mu`: N values.
  template <class S, class Rec>
  LQGK_HD static void moments(const Rec& rec, const S* x0, const S* x1, S* c, S* mu) {
    using O = Ops<S>;
    S e[D];
    residual<S>(rec, x0, x1, c, e);
    LQGK_UNROLL for (int i = 0; i < D; ++i) mu[i] = O::sub(x1[i], e[i]);
    LQGK_UNROLL for (int i = 0; i < R; ++i) {
      S a = O::zero();
      LQGK_UNROLL for (int j = 0; j < D; ++j) a = O::fma(O::bc(rec[DM::REC_F + (D + i) * N + j]), x0[j], a);
      LQGK_UNROLL for (int j = 0; j < R; ++j) a = O::fma(O::bc(rec[DM::REC_F + (D + i) * N + D + j]), c[j], a);
      mu[D + i] = a;
    }
    update<S>(rec, x0, e, c);
  }
  // state-only forward step (no log-density): what the adjoint kernel re-runs between two checkpoints of c.  Same
  // operation sequence as fwd(), so the recomputed states are bit-identical to the ones the forward pass carried.
  template <class S, class Rec>
  LQGK_HD static void advance(const Rec& rec, const S* x0, const S* x1, S* c) {
    S e[D];
    residual<S>(rec, x0, x1, c, e);
    update<S>(rec, x0, e, c);
  }
  // reverse step: given c (= c_t), x0 = x_t, x1 = x_{t+1}, weight w and the incoming cotangent cb (of c_{t+1});
  // produces e, v = S'^-1 e, wv = w v, eb = J^T cb - w v, neb = -eb and the outgoing cotangent
  // cbn = F[u,u]^T cb - F[o,u]^T eb (of c_t; the record holds -F[o,:]).  Sums over trials: sum_acc.
  template <class S, class Rec>
  LQGK_HD static void rev(const Rec& rec, const S* x0, const S* x1, const S* c, S w, const S* cb, S* e, S* v, S* wv, S* neb,
                          S* cbn) {
    using O = Ops<S>;
    S z[D];
    residual<S>(rec, x0, x1, c, e);
    whiten<S>(rec, e, z);
    LQGK_UNROLL for (int i = 0; i < D; ++i) {                         // v = Linv^T z ; wv = w v
      S a = O::zero();
      LQGK_UNROLL for (int k = i; k < D; ++k) a = O::fma(O::bc(rec[DM::REC_LINV + k * (k + 1) / 2 + i]), z[k], a);
      v[i] = a;
      wv[i] = O::mul(w, a);
    }
    S eb[D];
    LQGK_UNROLL for (int j = 0; j < D; ++j) {                         // eb = J^T cb - w v
      S a = O::zero();
      LQGK_UNROLL for (int i = 0; i < R; ++i) a = O::fma(O::bc(rec[DM::REC_J + i * D + j]), cb[i], a);
      eb[j] = O::sub(a, wv[j]);
      neb[j] = O::sub(wv[j], a);
    }
    LQGK_UNROLL for (int j = 0; j < R; ++j) {                         // cbn = F[u,u]^T cb + (-F[o,u])^T eb
      S a = O::zero();
      LQGK_UNROLL for (int i = 0; i < R; ++i) a = O::fma(O::bc(rec[DM::REC_F + (D + i) * N + D + j]), cb[i], a);
      LQGK_UNROLL for (int i = 0; i < D; ++i) a = O::fma(O::bc(rec[DM::REC_F + i * N + D + j]), eb[i], a);
      cbn[j] = a;
    }
  }
  LQGK_HD static constexpr int tri_row(int k) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= k) ++i;
    return i;
  }
  // acc += term IDX (compile-time) of the per-step sums (layout DM::SUM_*):
  //   Fb[i][j] = a_i b_j with a = [neb ; cb], b = [x0 ; c];  Jb[i][j] = cb_i e_j;  Wv[i>=j] = (w v_i) v_j.
  template <int IDX, class S>
  LQGK_HD static S sum_acc(S acc, const S* cb, const S* neb, const S* x0, const S* c, const S* e, const S* v, const S* wv) {
    using O = Ops<S>;
    if constexpr (IDX >= DM::NSUM || (IDX >= N * N && IDX < DM::SUM_J)) {
      return acc;
    } else if constexpr (IDX < N * N) {
      constexpr int i = IDX / N, j = IDX % N;
      S a, b;
      if constexpr (i < D) a = neb[i]; else a = cb[i - D];
      if constexpr (j < D) b = x0[j]; else b = c[j - D];
      return O::fma(a, b, acc);
    } else if constexpr (IDX < DM::SUM_W) {
      constexpr int k = IDX - DM::SUM_J;
      return O::fma(cb[k / D], e[k % D], acc);
    } else {
      constexpr int k = IDX - DM::SUM_W;
      constexpr int i = tri_row(k);
      constexpr int j = k - i * (i + 1) / 2;
      return O::fma(wv[i], v[j], acc);
    }
  }
};

// ================================================================================================
// Covariance adjoint, restructured for the GPU: the only sequential quantity is the
// cotangent Cb of the carried covariance block.  CovSeqRev advances Cb by one step using only stored forward
// quantities (Fu_t, J_t, S'^-1_t) and emits the symmetric joint cotangent Sgb_t and SF_t = Sgb_t Fu_t; the
// expensive, time-independent contraction of those with (L_t, K_t, constants) is done by CovContrib for all
// time steps in parallel, which writes per-step contributions that a generic reduction sums over time.
template <class DM>
struct CovSeqRev {
  static constexpr int D = DM::D, N = DM::N, R = DM::R;
  static constexpr int NSGB = tri(N), NSF = N * R, NJS = R * D + tri(D);
  static constexpr int SC_BH = 0, SC_SS = R * D, SC_N = R * D + tri(D);   // scratch (shared memory) layout

  // fu(e): Fu_t row-major N x R;  js(e): J_t (R x D) then S'^-1_t (packed lower);  get(idx): float sums of step t
  // sc(e): per-lane scratch;  Cb: R x R full symmetric (in: cotangent of C_{t+1}; out: of C_t)
  // esgb(e, v): emit Sgb_t packed lower over the joint index;  esf(e, v): emit SF_t row-major N x R
  template <class FU, class JS, class Get, class SC, class ESgb, class ESf>
  LQGK_HD static void step(FU&& fu, JS&& js, Get&& get, double sw, SC&& sc, double* Cb, ESgb&& esgb, ESf&& esf) {
    {
      double J[R * D], Sinv[D * D], Ss[D * D];
      LQGK_UNROLL64 for (int e = 0; e < R * D; ++e) J[e] = js(e);
      LQGK_UNROLL64 for (int i = 0; i < D; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
        double a = js(R * D + i * (i + 1) / 2 + j);
        Sinv[i * D + j] = a;
        Sinv[j * D + i] = a;
        Ss[i * D + j] = 0.5 * (double)get(DM::SUM_W + i * (i + 1) / 2 + j) - 0.5 * sw * a;
      }
      LQGK_UNROLL64 for (int i = 0; i < R; ++i) {
        double z[D];
        LQGK_UNROLL64 for (int j = 0; j < D; ++j) {
          double cbj = 0.0, jbs = 0.0;
          LQGK_UNROLL64 for (int k = 0; k < R; ++k) cbj += Cb[i * R + k] * J[k * D + j];
          LQGK_UNROLL64 for (int k = 0; k < D; ++k) jbs += (double)get(DM::SUM_J + i * D + k) * Sinv[k * D + j];
          z[j] = cbj - jbs;
          double bh = -cbj + 0.5 * jbs;
          sc(SC_BH + i * D + j) = bh;
          esgb(sidx(D + i, j), bh);
        }
        LQGK_UNROLL64 for (int a = 0; a < D; ++a) LQGK_UNROLL64 for (int b = 0; b <= a; ++b)
          Ss[a * D + b] += 0.5 * (J[i * D + a] * z[b] + J[i * D + b] * z[a]);
      }
      LQGK_UNROLL64 for (int a = 0; a < D; ++a) LQGK_UNROLL64 for (int b = 0; b <= a; ++b) {
        sc(SC_SS + a * (a + 1) / 2 + b) = Ss[a * D + b];
        esgb(sidx(a, b), Ss[a * D + b]);
      }
      LQGK_UNROLL64 for (int i = 0; i < R; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) esgb(sidx(D + i, D + j), Cb[i * R + j]);
    }
    double Fo[D * R], Fuu[R * R], Cn[R * R];
    LQGK_UNROLL64 for (int e = 0; e < D * R; ++e) Fo[e] = fu(e);
    LQGK_UNROLL64 for (int e = 0; e < R * R; ++e) Fuu[e] = fu(D * R + e);
    LQGK_UNROLL64 for (int e = 0; e < R * R; ++e) Cn[e] = 0.0;
    LQGK_UNROLL64 for (int i = 0; i < R; ++i) {            // SF_u row i = Bh[i,:] Fo + Cb[i,:] Fuu
      double row[R];
      LQGK_UNROLL64 for (int c = 0; c < R; ++c) row[c] = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < D; ++k) {
        double b = sc(SC_BH + i * D + k);
        LQGK_UNROLL64 for (int c = 0; c < R; ++c) row[c] += b * Fo[k * R + c];
      }
      LQGK_UNROLL64 for (int k = 0; k < R; ++k) LQGK_UNROLL64 for (int c = 0; c < R; ++c) row[c] += Cb[i * R + k] * Fuu[k * R + c];
      LQGK_UNROLL64 for (int c = 0; c < R; ++c) esf((D + i) * R + c, row[c]);
      LQGK_UNROLL64 for (int a = 0; a < R; ++a) LQGK_UNROLL64 for (int b = 0; b <= a; ++b) Cn[a * R + b] += Fuu[i * R + a] * row[b];
    }
    LQGK_UNROLL64 for (int k = 0; k < D; ++k) {            // SF_o row k = Ss[k,:] Fo + Bh[:,k]^T Fuu
      double row[R];
      LQGK_UNROLL64 for (int c = 0; c < R; ++c) row[c] = 0.0;
      LQGK_UNROLL64 for (int m = 0; m < D; ++m) {
        double sv = sc(SC_SS + sidx(k, m));
        LQGK_UNROLL64 for (int c = 0; c < R; ++c) row[c] += sv * Fo[m * R + c];
      }
      LQGK_UNROLL64 for (int i = 0; i < R; ++i) {
        double b = sc(SC_BH + i * D + k);
        LQGK_UNROLL64 for (int c = 0; c < R; ++c) row[c] += b * Fuu[i * R + c];
      }
      LQGK_UNROLL64 for (int c = 0; c < R; ++c) esf(k * R + c, row[c]);
      LQGK_UNROLL64 for (int a = 0; a < R; ++a) LQGK_UNROLL64 for (int b = 0; b <= a; ++b) Cn[a * R + b] += Fo[k * R + a] * row[b];
    }
    LQGK_UNROLL64 for (int a = 0; a < R; ++a) LQGK_UNROLL64 for (int b = 0; b <= a; ++b) {
      Cb[a * R + b] = Cn[a * R + b];
      Cb[b * R + a] = Cn[a * R + b];
    }
  }
  // After t = 0: cotangent of C_0 = cond(N_0) as a joint symmetric cotangent (Sgb of the initial condition).
  template <class J0, class ESgb>
  LQGK_HD static void init(J0&& j0, const double* Cb, ESgb&& esgb) {
    double J[R * D], CbJ[R * D];
    LQGK_UNROLL64 for (int e = 0; e < R * D; ++e) J[e] = j0(e);
    mm<R, R, D>(Cb, J, CbJ);
    LQGK_UNROLL64 for (int i = 0; i < R; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) esgb(sidx(D + i, D + j), Cb[i * R + j]);
    LQGK_UNROLL64 for (int i = 0; i < R; ++i) LQGK_UNROLL64 for (int j = 0; j < D; ++j) esgb(sidx(D + i, j), -CbJ[i * D + j]);
    LQGK_UNROLL64 for (int a = 0; a < D; ++a) LQGK_UNROLL64 for (int b = 0; b <= a; ++b) {
      double v = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < R; ++k) v += J[k * D + a] * CbJ[k * D + b];
      esgb(sidx(a, b), v);
    }
  }
};

// Time-parallel contraction of (Sgb_t, SF_t, trial sums) with (L_t, K_t, constants): per-step contributions to the
// derived-constant cotangents (CovC layout) and the gain cotangents Lb_t, Kb_t.  Two passes keep the register
// footprint below 255: PASS 0 = noise part (N11, FN, Om, Kb noise part) + Ad, Bd, FAd, Aa, Ba, Lb;
// PASS 1 = FAa, Dm and the transition part of Kb (added to what PASS 0 stored).
template <class DM>
struct CovContrib {
  static constexpr int X = DM::X, B = DM::B, U = DM::U, Y = DM::Y, D = DM::D, N = DM::N, R = DM::R;
  using C = CovC<DM>;

  // Row m of the joint transition cotangent: trial sums + [0 | 2 SF_m C].
  template <class SFr, class Get>
  LQGK_HD static void fb_row(int m, SFr&& sf, Get&& get, const double* Cm, bool use_trial, double* fb) {
    double sfr[R];
    LQGK_UNROLL64 for (int k = 0; k < R; ++k) sfr[k] = sf(m * R + k);
    LQGK_UNROLL64 for (int j = 0; j < N; ++j) fb[j] = use_trial ? (double)get(DM::SUM_F + m * N + j) : 0.0;
    LQGK_UNROLL64 for (int j = 0; j < R; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < R; ++k) a += sfr[k] * Cm[k * R + j];
      fb[D + j] += 2.0 * a;
    }
  }
  // Noise part for one symmetric joint cotangent Sgb (read through sgb(e), packed lower).  Accumulates (+=) into
  // n11[tri X], fn[Y*X], om[tri Y] and kb[B*Y].
  template <class V, class SG>
  LQGK_HD static void noise_part(const V& c, SG&& sgb, const double* K, double* n11, double* fn, double* om, double* kb) {
    LQGK_UNROLL64 for (int i = 0; i < X; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) n11[i * (i + 1) / 2 + j] += sgb(sidx(i, j));
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) {
      double nbx[X], nbb[B], nk[Y];
      LQGK_UNROLL64 for (int j = 0; j < X; ++j) nbx[j] = sgb(sidx(X + i, j));
      LQGK_UNROLL64 for (int j = 0; j < B; ++j) nbb[j] = sgb(sidx(X + i, X + j));
      LQGK_UNROLL64 for (int k = 0; k < Y; ++k) {
        double a = 0.0;
        LQGK_UNROLL64 for (int j = 0; j < B; ++j) a += nbb[j] * K[j * Y + k];
        nk[k] = a;                                                           // (Nbb K)[i, k]
        LQGK_UNROLL64 for (int j = 0; j < X; ++j) fn[k * X + j] += 2.0 * K[i * Y + k] * nbx[j];
      }
      LQGK_UNROLL64 for (int k = 0; k < Y; ++k) LQGK_UNROLL64 for (int m = 0; m <= k; ++m)
        om[k * (k + 1) / 2 + m] += 0.5 * (K[i * Y + k] * nk[m] + K[i * Y + m] * nk[k]);
      LQGK_UNROLL64 for (int k = 0; k < Y; ++k) {
        double a = 0.0;
        LQGK_UNROLL64 for (int j = 0; j < X; ++j) a += nbx[j] * c(C::FN + k * X + j);
        LQGK_UNROLL64 for (int m = 0; m < Y; ++m) a += nk[m] * c(C::Om + sidx(m, k));
        kb[i * Y + k] += 2.0 * a;
      }
    }
  }
  // PASS 0.  out(e, v): store contribution e (CovC layout) of this step;  Lb, Kb: overwritten.
  // `init_sgb` (may be null): second Sgb (initial condition) whose noise part is added at t == 0.
  template <class V, class SG, class SGI, class SFr, class Get, class Out>
  LQGK_HD static void pass0(const V& c, SG&& sgb, bool has_init, SGI&& sgb_init, SFr&& sf, Get&& get, const double* Cm,
                            const double* L, const double* K, Out&& out, double* Lb, double* Kb) {
    {
      double n11[tri(X)], fn[Y * X], om[tri(Y)];
      LQGK_UNROLL64 for (int e = 0; e < tri(X); ++e) n11[e] = 0.0;
      LQGK_UNROLL64 for (int e = 0; e < Y * X; ++e) fn[e] = 0.0;
      LQGK_UNROLL64 for (int e = 0; e < tri(Y); ++e) om[e] = 0.0;
      LQGK_UNROLL64 for (int e = 0; e < B * Y; ++e) Kb[e] = 0.0;
      noise_part(c, sgb, K, n11, fn, om, Kb);
      if (has_init) noise_part(c, sgb_init, K, n11, fn, om, Kb);
      LQGK_UNROLL64 for (int e = 0; e < tri(X); ++e) out(C::N11 + e, n11[e]);
      LQGK_UNROLL64 for (int e = 0; e < Y * X; ++e) out(C::FN + e, fn[e]);
      LQGK_UNROLL64 for (int e = 0; e < tri(Y); ++e) out(C::Om + e, om[e]);
    }
    double KD[B * U], fad[Y * X];
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int m = 0; m < U; ++m) {
      double a = c(C::Ba + i * U + m);
      LQGK_UNROLL64 for (int k = 0; k < Y; ++k) a += K[i * Y + k] * c(C::Dm + k * U + m);
      KD[i * U + m] = a;                                                     // Ba + K D
    }
    LQGK_UNROLL64 for (int e = 0; e < Y * X; ++e) fad[e] = 0.0;
    LQGK_UNROLL64 for (int e = 0; e < U * B; ++e) Lb[e] = 0.0;
    LQGK_UNROLL64 for (int m = 0; m < N; ++m) {
      double fb[N];
      fb_row(m, sf, get, Cm, true, fb);
      if (m < X) {
        LQGK_UNROLL64 for (int j = 0; j < X; ++j) out(C::Ad + m * X + j, fb[j]);
        LQGK_UNROLL64 for (int k = 0; k < U; ++k) {
          double a = 0.0;
          LQGK_UNROLL64 for (int j = 0; j < B; ++j) a += fb[X + j] * L[k * B + j];
          out(C::Bd + m * U + k, a);
          double bd = c(C::Bd + m * U + k);
          LQGK_UNROLL64 for (int j = 0; j < B; ++j) Lb[k * B + j] += bd * fb[X + j];
        }
      } else {
        const int i = m - X;
        LQGK_UNROLL64 for (int k = 0; k < Y; ++k) LQGK_UNROLL64 for (int j = 0; j < X; ++j) fad[k * X + j] += K[i * Y + k] * fb[j];
        LQGK_UNROLL64 for (int j = 0; j < B; ++j) out(C::Aa + i * B + j, fb[X + j]);
        LQGK_UNROLL64 for (int k = 0; k < U; ++k) {
          double a = 0.0;
          LQGK_UNROLL64 for (int j = 0; j < B; ++j) a += fb[X + j] * L[k * B + j];
          out(C::Ba + i * U + k, a);
          LQGK_UNROLL64 for (int j = 0; j < B; ++j) Lb[k * B + j] += KD[i * U + k] * fb[X + j];
        }
      }
    }
    LQGK_UNROLL64 for (int e = 0; e < Y * X; ++e) out(C::FAd + e, fad[e]);
  }
  // PASS 1.  Kb: in = what PASS 0 stored (noise part), out = total.
  template <class V, class SFr, class Get, class Out>
  LQGK_HD static void pass1(const V& c, SFr&& sf, Get&& get, const double* Cm, const double* L, const double* K, Out&& out,
                            double* Kb) {
    double faa[Y * B], dm[Y * U];
    LQGK_UNROLL64 for (int e = 0; e < Y * B; ++e) faa[e] = 0.0;
    LQGK_UNROLL64 for (int e = 0; e < Y * U; ++e) dm[e] = 0.0;
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) {
      double fb[N], t2[U];
      fb_row(X + i, sf, get, Cm, true, fb);
      LQGK_UNROLL64 for (int k = 0; k < U; ++k) {
        double a = 0.0;
        LQGK_UNROLL64 for (int j = 0; j < B; ++j) a += fb[X + j] * L[k * B + j];
        t2[k] = a;                                                           // (F22 L^T)[i, k]
      }
      LQGK_UNROLL64 for (int k = 0; k < Y; ++k) {
        LQGK_UNROLL64 for (int j = 0; j < B; ++j) faa[k * B + j] -= K[i * Y + k] * fb[X + j];
        LQGK_UNROLL64 for (int m = 0; m < U; ++m) dm[k * U + m] += K[i * Y + k] * t2[m];
        double a = 0.0;
        LQGK_UNROLL64 for (int j = 0; j < X; ++j) a += fb[j] * c(C::FAd + k * X + j);
        LQGK_UNROLL64 for (int j = 0; j < B; ++j) a -= fb[X + j] * c(C::FAa + k * B + j);
        LQGK_UNROLL64 for (int m = 0; m < U; ++m) a += t2[m] * c(C::Dm + k * U + m);
        Kb[i * Y + k] += a;
      }
    }
    LQGK_UNROLL64 for (int e = 0; e < Y * B; ++e) out(C::FAa + e, faa[e]);
    LQGK_UNROLL64 for (int e = 0; e < Y * U; ++e) out(C::Dm + e, dm[e]);
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
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int k = 0; k < Y; ++k) {
      double a = Kb[i * Y + k];
      LQGK_UNROLL64 for (int j = 0; j < B; ++j) a -= Pnb[i * B + j] * M[k * B + j];
      Ktot[i * Y + k] = a;
    }
    mm<B, Y, Y>(Ktot, Gi, Yv);
    // Mb = -K^T Pnb + Y^T  (Y x B)
    double Mb[Y * B];
    LQGK_UNROLL64 for (int k = 0; k < Y; ++k) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
      double a = Yv[j * Y + k];
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) a -= K[i * Y + k] * Pnb[i * B + j];
      Mb[k * B + j] = a;
    }
    // Gmb = -sym(K^T Y)  (Y x Y)
    double Gmb[Y * Y];
    LQGK_UNROLL64 for (int k = 0; k < Y; ++k) LQGK_UNROLL64 for (int m = 0; m <= k; ++m) {
      double a = 0.0;
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) a += K[i * Y + k] * Yv[i * Y + m] + K[i * Y + m] * Yv[i * Y + k];
      Gmb[k * Y + m] = -0.5 * a;
      Gmb[m * Y + k] = -0.5 * a;
      acc(C::WWa + k * (k + 1) / 2 + m) += -0.5 * a;
    }
    // Fa += Mb Pp + 2 Gmb M
    LQGK_UNROLL64 for (int k = 0; k < Y; ++k) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int i = 0; i < B; ++i) a += Mb[k * B + i] * Pp[i * B + j];
      LQGK_UNROLL64 for (int m = 0; m < Y; ++m) a += 2.0 * Gmb[k * Y + m] * M[m * B + j];
      acc(C::Fa + k * B + j) += a;
    }
    // Ppb = Pnb + sym(F^T Mb) + F^T Gmb F
    double GF[Y * B], Ppb[B * B];
    mm<Y, Y, B>(Gmb, F, GF);
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
      double a = Pnb[i * B + j];
      LQGK_UNROLL64 for (int k = 0; k < Y; ++k)
        a += 0.5 * (F[k * B + i] * Mb[k * B + j] + F[k * B + j] * Mb[k * B + i]) + F[k * B + i] * GF[k * B + j];
      Ppb[i * B + j] = a;
      Ppb[j * B + i] = a;
      acc(C::VVa + i * (i + 1) / 2 + j) += a;
    }
    // Aa += 2 Ppb A P ; Pnb <- A^T Ppb A
    double PA[B * B], AP[B * B];
    mm<B, B, B>(Ppb, Am, PA);                       // Ppb A
    mm<B, B, B>(Am, P, AP);                         // A P
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < B; ++k) a += Ppb[i * B + k] * AP[k * B + j];
      acc(C::Aa + i * B + j) += 2.0 * a;
    }
    mm_tn_sym<B, B>(Am, PA, Pnb);
  }
  // After t = 0: P_0 = Sig0
  template <class A>
  LQGK_HD static void finish(A&& acc, const double* Pnb) {
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) acc(C::Sig0 + i * (i + 1) / 2 + j) += Pnb[i * B + j];
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
      LQGK_UNROLL64 for (int i = 0; i < U * U; ++i) Lc[i] = H[i];
      LQGK_UNROLL64 for (int i = 0; i < U; ++i) Lc[i * U + i] += shift;
      chol_and_inverse<U>(Lc, Li);
      mm_tn_sym<U, U>(Li, Li, Hi);
    }
    // Q += Sn ; Aa += 2 SA Sn
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) acc(C::Q + i * (i + 1) / 2 + j) += Sn[i * B + j];
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < B; ++k) a += SA[i * B + k] * Sn[k * B + j];
      acc(C::Aa + i * B + j) += 2.0 * a;
    }
    // LS = L Sn (U x B) ; Lb = Lbar + 2 (H L + G) Sn ; Hb = LS L^T ; Gb = 2 LS
    double LS[U * B], HLG[U * B], Lb[U * B], Hb[U * U], Gb[U * B];
    mm<U, B, B>(L, Sn, LS);
    mm<U, U, B>(H, L, HLG);
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) HLG[i] += G[i];
    mm<U, B, B>(HLG, Sn, Lb);
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) Lb[i] = Lbar[i] + 2.0 * Lb[i];
    mm_nt<U, B, U>(LS, L, Hb);
    // HiLb = Ht^-1 Lb ; Gb = 2 LS - HiLb ; Hb -= HiLb L^T ; Hb <- sym
    double HiLb[U * B];
    mm<U, U, B>(Hi, Lb, HiLb);
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) Gb[i] = 2.0 * LS[i] - HiLb[i];
    LQGK_UNROLL64 for (int i = 0; i < U; ++i) LQGK_UNROLL64 for (int j = 0; j < U; ++j) {
      double a = Hb[i * U + j];
      LQGK_UNROLL64 for (int k = 0; k < B; ++k) a -= HiLb[i * B + k] * L[j * B + k];
      Hb[i * U + j] = a;
    }
    symmetrize<U>(Hb);
    LQGK_UNROLL64 for (int i = 0; i < U; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) acc(C::R + i * (i + 1) / 2 + j) += Hb[i * U + j];
    // Ba += 2 SB Hb + SA Gb^T ; Aa += SB Gb
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int m = 0; m < U; ++m) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < U; ++k) a += 2.0 * SB[i * U + k] * Hb[k * U + m];
      LQGK_UNROLL64 for (int k = 0; k < B; ++k) a += SA[i * B + k] * Gb[m * B + k];
      acc(C::Ba + i * U + m) += a;
    }
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j < B; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < U; ++k) a += SB[i * U + k] * Gb[k * B + j];
      acc(C::Aa + i * B + j) += a;
    }
    // Sn <- A Sn A^T + B Hb B^T + sym(B Gb A^T)
    double AS[B * B], BH[B * U], BG[B * B], Sb[B * B];
    mm<B, B, B>(Am, Sn, AS);
    mm_nt_sym<B, B>(AS, Am, Sb);
    mm<B, U, U>(Bm, Hb, BH);
    mm_nt_sym<B, U, true>(BH, Bm, Sb);
    mm<B, U, B>(Bm, Gb, BG);                                       // B Gb (B x B), then (B Gb) A^T
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
      double a = 0.0;
      LQGK_UNROLL64 for (int k = 0; k < B; ++k) a += BG[i * B + k] * Am[j * B + k] + BG[j * B + k] * Am[i * B + k];
      Sb[i * B + j] += 0.5 * a;
      if (i != j) Sb[j * B + i] += 0.5 * a;
    }
    LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sn[i] = Sb[i];
  }
  template <class A>
  LQGK_HD static void finish(A&& acc, const double* Sn) {   // S_T = Qf
    LQGK_UNROLL64 for (int i = 0; i < B; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) acc(C::Qf + i * (i + 1) / 2 + j) += Sn[i * B + j];
  }
};

}  // namespace lqgk
