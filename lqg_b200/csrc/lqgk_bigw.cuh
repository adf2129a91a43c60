// lqgk_bigw.cuh -- WARP-PER-SAMPLE covariance kernels of the large-system path (joint dim n > 12).
//
// The joint-system covariance scan and its adjoint carry the big matrices of a large system (n x n, n x r, r x r with
// n = 24, r = 22 at config c4): one warp owns one parameter sample, the matrices live in that warp's slice of shared memory
// (FP64, row-major), every matrix product is spread over the 32 lanes by output element, and dependent products are
// separated by __syncwarp().  Per-step inputs/outputs of these three kernels use SAMPLE-MAJOR workspace arrays
// ([sample][t][element]) so a warp's loads and stores are contiguous; the small per-step gain arrays (L, K, Lbar, Kbar)
// stay sample-minor because the thread-per-sample Riccati / Kalman kernels produce and consume them.
//
// Mathematics: identical to CovFwd / CovSeqRev / CovContrib in lqgk_core.h (which the host emulation and the small-system
// kernels execute) and to oracle/adjoint_np.py; reference lines: lqg/system.py:163-212, 223-230 and their reverse mode.
#pragma once
#include "lqgk_run.cuh"

namespace lqgk {

constexpr int BW_WARPS = 4;   // samples (warps) per CTA

// C[M,N] (op)= A B with A addressed as a(i,k), B as b(k,j); out(i, j, value) consumes each element once.
template <int M, int N, int K, class FA, class FB, class Out>
__device__ __forceinline__ void wmm(int lane, FA&& a, FB&& b, Out&& out) {
  for (int e = lane; e < M * N; e += 32) {
    const int i = e / N, j = e - i * N;
    double acc = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) acc += a(i, k) * b(k, j);
    out(i, j, acc);
  }
}

template <class DM>
struct BigW {
  static constexpr int X = DM::X, B = DM::B, U = DM::U, Y = DM::Y, D = DM::D, N = DM::N, R = DM::R;
  using C = CovC<DM>;
  using SR = CovSeqRev<DM>;
  static constexpr int NC = C::n;

  // ------------------------------------------------------------------------------------------ forward
  static constexpr int FWD_DOUBLES = NC + U * B + B * Y + B * U + B * Y + 3 * N * N + R * R + N * R + R * D + 3 * D * D + 8;
  static size_t smem_fwd() { return sizeof(double) * FWD_DOUBLES * BW_WARPS; }
  static constexpr int SEQ_DOUBLES = 2 * R * R + 2 * N * R + 5 * R * D + 2 * D * D + 8;
  static size_t smem_seq() { return sizeof(double) * SEQ_DOUBLES * BW_WARPS; }
  static constexpr int CON_DOUBLES = 2 * NC + U * B + B * Y + R * R + N * R + 2 * N * N + 2 * B * Y + 2 * B * U + U * B + 8;
  static size_t smem_con() { return sizeof(double) * CON_DOUBLES * BW_WARPS; }
};

// Joint noise covariance N(K) (full symmetric) into Nj; KO is B*Y scratch.  Caller syncs afterwards.
template <class DM>
__device__ __forceinline__ void bw_joint_N(int lane, const double* c, const double* K, double* KO, double* Nj) {
  using W = BigW<DM>;
  using C = typename W::C;
  constexpr int X = W::X, B = W::B, Y = W::Y, N = W::N;
  wmm<B, Y, Y>(lane, [&](int i, int k) { return K[i * Y + k]; }, [&](int k, int j) { return c[C::Om + sidx(k, j)]; },
               [&](int i, int j, double v) { KO[i * Y + j] = v; });
  __syncwarp();
  for (int e = lane; e < N * N; e += 32) {
    const int i = e / N, j = e - i * N;
    if (j > i) continue;
    double a = 0.0;
    if (i < X) a = c[C::N11 + sidx(i, j)];
    else if (j < X) { for (int k = 0; k < Y; ++k) a += K[(i - X) * Y + k] * c[C::FN + k * X + j]; }
    else { for (int k = 0; k < Y; ++k) a += KO[(i - X) * Y + k] * K[(j - X) * Y + k]; }
    Nj[i * N + j] = a;
    Nj[j * N + i] = a;
  }
}

// Condition Sig (full symmetric, shared memory) on its first D coordinates (CovFwd::condition).  Linv/Z scratch in shared
// memory; emit_j(idx, value) for J (R x D), C written full symmetric.  All lanes call; ends synchronised.
template <class DM, class EmitJ>
__device__ __forceinline__ void bw_condition(int lane, const double* Sig, double* Linv, double* Z, double* ld, EmitJ&& emit_j, double* Cn) {
  using W = BigW<DM>;
  constexpr int D = W::D, N = W::N, R = W::R;
  if (lane == 0) {
    double Lc[D * D], Li[D * D];
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) Lc[i * D + j] = Sig[i * N + j];
    chol<D>(Lc);
    double l = 0.0;
    for (int i = 0; i < D; ++i) l += log(Lc[i * D + i]);
    tri_inv<D>(Lc, Li);
    for (int i = 0; i < D * D; ++i) Linv[i] = Li[i];
    *ld = l;
  }
  __syncwarp();
  for (int e = lane; e < R * D; e += 32) {
    const int i = e / D, j = e - i * D;
    double a = 0.0;
    for (int k = 0; k <= j; ++k) a += Sig[(D + i) * N + k] * Linv[j * D + k];
    Z[e] = a;
  }
  __syncwarp();
  for (int e = lane; e < R * D; e += 32) {
    const int i = e / D, j = e - i * D;
    double a = 0.0;
    for (int k = j; k < D; ++k) a += Z[i * D + k] * Linv[k * D + j];
    emit_j(e, a);
  }
  for (int e = lane; e < R * R; e += 32) {
    const int i = e / R, j = e - i * R;
    if (j > i) continue;
    double a = Sig[(D + i) * N + D + j];
    for (int k = 0; k < D; ++k) a -= Z[i * D + k] * Z[j * D + k];
    Cn[i * R + j] = a;
    Cn[j * R + i] = a;
  }
  __syncwarp();
}

// Covariance pass (forward), warp per sample.  L, K: sample-minor [t][e][Sc];  Cs, FU, JS, J0: sample-major;  rec [s][t][REC].
template <class DM>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_cov_fwd(const double* lcc, size_t Sc, int npad, int Tn, const double* L, const double* K,
                                                            int save_adj, double* Cs, double* FU, double* JS, double* J0, float* rec) {
  using W = BigW<DM>;
  using C = typename W::C;
  using SR = typename W::SR;
  constexpr int X = W::X, B = W::B, U = W::U, Y = W::Y, D = W::D, N = W::N, R = W::R, NC = W::NC;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  double* c = smw + (size_t)warp * W::FWD_DOUBLES;
  double* Lt = c + NC;
  double* Kt = Lt + U * B;
  double* KD = Kt + B * Y;
  double* KO = KD + B * U;
  double* Fj = KO + B * Y;
  double* Nj = Fj + N * N;
  double* Sig = Nj + N * N;
  double* Cm = Sig + N * N;
  double* T1 = Cm + R * R;
  double* Z = T1 + N * R;
  double* Linv = Z + R * D;
  double* ld = Linv + 3 * D * D;
  for (int e = lane; e < NC; e += 32) c[e] = lcc[(size_t)e * Sc + s];
  for (int e = lane; e < B * Y; e += 32) Kt[e] = K[(size_t)e * Sc + s];
  __syncwarp();
  bw_joint_N<DM>(lane, c, Kt, KO, Nj);
  __syncwarp();
  bw_condition<DM>(lane, Nj, Linv, Z, ld, [&](int e, double v) { if (save_adj) J0[s * (R * D) + e] = v; }, Cm);
  float* recs = rec + s * (size_t)Tn * DM::REC;
  for (int t = 0; t < Tn; ++t) {
    for (int e = lane; e < U * B; e += 32) Lt[e] = L[((size_t)t * DM::EL + e) * Sc + s];
    for (int e = lane; e < B * Y; e += 32) Kt[e] = K[((size_t)t * DM::EK + e) * Sc + s];
    if (save_adj) {
      double* cs = Cs + (s * Tn + t) * (size_t)DM::EC;
      for (int e = lane; e < R * R; e += 32) {
        const int i = e / R, j = e - i * R;
        if (j <= i) cs[i * (i + 1) / 2 + j] = Cm[e];
      }
    }
    __syncwarp();
    wmm<B, U, Y>(lane, [&](int i, int k) { return Kt[i * Y + k]; }, [&](int k, int j) { return c[C::Dm + k * U + j]; },
                 [&](int i, int j, double v) { KD[i * U + j] = v; });
    __syncwarp();
    float* rt = recs + (size_t)t * DM::REC;
    double* fu = FU + (s * Tn + t) * (size_t)SR::NSF;
    for (int e = lane; e < N * N; e += 32) {                               // joint transition F_t  (system.py:167-187)
      const int i = e / N, j = e - i * N;
      double a;
      if (i < X) {
        if (j < X) a = c[C::Ad + i * X + j];
        else { a = 0.0; for (int k = 0; k < U; ++k) a += c[C::Bd + i * U + k] * Lt[k * B + (j - X)]; }
      } else {
        const int ib = i - X;
        if (j < X) { a = 0.0; for (int k = 0; k < Y; ++k) a += Kt[ib * Y + k] * c[C::FAd + k * X + j]; }
        else {
          const int jb = j - X;
          a = c[C::Aa + ib * B + jb];
          for (int k = 0; k < U; ++k) a += (c[C::Ba + ib * U + k] + KD[ib * U + k]) * Lt[k * B + jb];
          for (int k = 0; k < Y; ++k) a -= Kt[ib * Y + k] * c[C::FAa + k * B + jb];
        }
      }
      Fj[e] = a;
      rt[DM::REC_F + e] = (float)(e < D * N ? -a : a);
      if (save_adj && j >= D) fu[i * R + (j - D)] = a;
    }
    bw_joint_N<DM>(lane, c, Kt, KO, Nj);
    __syncwarp();
    wmm<N, R, R>(lane, [&](int i, int k) { return Fj[i * N + D + k]; }, [&](int k, int j) { return Cm[k * R + j]; },
                 [&](int i, int j, double v) { T1[i * R + j] = v; });
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) {                               // Sig' = Fu C Fu^T + N (lower, mirrored)
      const int i = e / N, j = e - i * N;
      if (j > i) continue;
      double a = Nj[e];
      for (int k = 0; k < R; ++k) a += T1[i * R + k] * Fj[j * N + D + k];
      Sig[i * N + j] = a;
      Sig[j * N + i] = a;
    }
    __syncwarp();
    double* js = JS + (s * Tn + t) * (size_t)SR::NJS;
    bw_condition<DM>(lane, Sig, Linv, Z, ld, [&](int e, double v) { rt[DM::REC_J + e] = (float)v; if (save_adj) js[e] = v; }, Cm);
    for (int e = lane; e < D * D; e += 32) {
      const int i = e / D, j = e - i * D;
      if (j > i) continue;
      rt[DM::REC_LINV + i * (i + 1) / 2 + j] = (float)Linv[i * D + j];
      if (save_adj) {
        double sv = 0.0;
        for (int k = i; k < D; ++k) sv += Linv[k * D + i] * Linv[k * D + j];
        js[R * D + i * (i + 1) / 2 + j] = sv;
      }
    }
    if (lane == 0) rt[DM::REC_LOGDET] = (float)*ld;
    __syncwarp();
  }
}

// Sequential covariance adjoint (CovSeqRev), warp per sample, t descending.
template <class DM>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_cov_seq_rev(int npad, int Tn, int Ntr, const float* w, const double* FU, const double* JS,
                                                                const double* J0, const float* sums, double* SGB, double* SGBI, double* SFW) {
  using W = BigW<DM>;
  using SR = typename W::SR;
  constexpr int D = W::D, N = W::N, R = W::R;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  double* Cb = smw + (size_t)warp * W::SEQ_DOUBLES;
  double* Cn = Cb + R * R;
  double* Fu = Cn + R * R;          // N x R: rows < D = Fo, rows >= D = Fuu
  double* SF = Fu + N * R;          // N x R: rows < D = SFo, rows >= D = SFu
  double* J = SF + N * R;
  double* Jb = J + R * D;
  double* Zm = Jb + R * D;
  double* Bh = Zm + R * D;
  double* CJ = Bh + R * D;
  double* Sinv = CJ + R * D;
  double* Ss = Sinv + D * D;
  double sw = 0.0;
  for (int i = lane; i < Ntr; i += 32) sw += (double)w[s * Ntr + i];
  for (int m = 16; m >= 1; m >>= 1) sw += __shfl_xor_sync(FULL, sw, m);
  for (int e = lane; e < R * R; e += 32) Cb[e] = 0.0;
  __syncwarp();
  for (int t = Tn - 1; t >= 0; --t) {
    const double* fu = FU + (s * Tn + t) * (size_t)SR::NSF;
    const double* js = JS + (s * Tn + t) * (size_t)SR::NJS;
    const float* sm = sums + (s * Tn + t) * (size_t)DM::SUMP;
    for (int e = lane; e < N * R; e += 32) Fu[e] = fu[e];
    for (int e = lane; e < R * D; e += 32) { J[e] = js[e]; Jb[e] = (double)sm[DM::SUM_J + e]; }
    for (int e = lane; e < D * D; e += 32) {
      const int i = e / D, j = e - i * D;
      Sinv[e] = js[R * D + sidx(i, j)];
    }
    __syncwarp();
    for (int e = lane; e < R * D; e += 32) {
      const int i = e / D, j = e - i * D;
      double cbj = 0.0, jbs = 0.0;
      for (int k = 0; k < R; ++k) cbj += Cb[i * R + k] * J[k * D + j];
      for (int k = 0; k < D; ++k) jbs += Jb[i * D + k] * Sinv[k * D + j];
      Zm[e] = cbj - jbs;
      Bh[e] = -cbj + 0.5 * jbs;
    }
    __syncwarp();
    for (int e = lane; e < D * D; e += 32) {
      const int a = e / D, b = e - a * D;
      if (b > a) continue;
      double v = 0.5 * (double)sm[DM::SUM_W + a * (a + 1) / 2 + b] - 0.5 * sw * Sinv[a * D + b];
      for (int i = 0; i < R; ++i) v += 0.5 * (J[i * D + a] * Zm[i * D + b] + J[i * D + b] * Zm[i * D + a]);
      Ss[a * D + b] = v;
      Ss[b * D + a] = v;
    }
    __syncwarp();
    double* sgb = SGB + (s * Tn + t) * (size_t)SR::NSGB;
    for (int e = lane; e < N * N; e += 32) {                               // Sgb_t, packed lower over the joint index
      const int i = e / N, j = e - i * N;
      if (j > i) continue;
      sgb[i * (i + 1) / 2 + j] = i < D ? Ss[i * D + j] : (j < D ? Bh[(i - D) * D + j] : Cb[(i - D) * R + (j - D)]);
    }
    double* sfw = SFW + (s * Tn + t) * (size_t)SR::NSF;
    for (int e = lane; e < N * R; e += 32) {                               // SF_t = Sgb_t Fu_t
      const int m = e / R, cc = e - m * R;
      double a = 0.0;
      if (m < D) {
        for (int k = 0; k < D; ++k) a += Ss[m * D + k] * Fu[k * R + cc];
        for (int i = 0; i < R; ++i) a += Bh[i * D + m] * Fu[(D + i) * R + cc];
      } else {
        const int i = m - D;
        for (int k = 0; k < D; ++k) a += Bh[i * D + k] * Fu[k * R + cc];
        for (int k = 0; k < R; ++k) a += Cb[i * R + k] * Fu[(D + k) * R + cc];
      }
      SF[e] = a;
      sfw[e] = a;
    }
    __syncwarp();
    for (int e = lane; e < R * R; e += 32) {                               // Cb <- Fu^T SF (lower, mirrored)
      const int a = e / R, b = e - a * R;
      if (b > a) continue;
      double v = 0.0;
      for (int m = 0; m < N; ++m) v += Fu[m * R + a] * SF[m * R + b];
      Cn[a * R + b] = v;
      Cn[b * R + a] = v;
    }
    __syncwarp();
    double* tmp = Cb; Cb = Cn; Cn = tmp;
  }
  // cotangent of C_0 = cond(N_0) as a joint symmetric cotangent (CovSeqRev::init)
  for (int e = lane; e < R * D; e += 32) J[e] = J0[s * (R * D) + e];
  __syncwarp();
  for (int e = lane; e < R * D; e += 32) {
    const int i = e / D, j = e - i * D;
    double a = 0.0;
    for (int k = 0; k < R; ++k) a += Cb[i * R + k] * J[k * D + j];
    CJ[e] = a;
  }
  __syncwarp();
  double* sgi = SGBI + s * (size_t)SR::NSGB;
  for (int e = lane; e < N * N; e += 32) {
    const int i = e / N, j = e - i * N;
    if (j > i) continue;
    double v;
    if (i >= D && j >= D) v = Cb[(i - D) * R + (j - D)];
    else if (i >= D) v = -CJ[(i - D) * D + j];
    else { v = 0.0; for (int k = 0; k < R; ++k) v += J[k * D + i] * CJ[k * D + j]; }
    sgi[i * (i + 1) / 2 + j] = v;
  }
}

// Time-parallel contraction (CovContrib pass 0 + pass 1), warp per (sample, time range).  la: [NC][Sc] accumulators
// (CovC layout, zeroed by the caller) updated with one FP64 atomicAdd per element and warp.
template <class DM>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_cov_contrib(const double* lcc, size_t Sc, int npad, int Tn, const double* L, const double* K,
                                                                const double* Cs, const double* SGB, const double* SGBI, const double* SFW,
                                                                const float* sums, double* la, double* Lbar, double* Kbar) {
  using W = BigW<DM>;
  using C = typename W::C;
  using SR = typename W::SR;
  constexpr int X = W::X, B = W::B, U = W::U, Y = W::Y, D = W::D, N = W::N, R = W::R, NC = W::NC;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  const int nq = gridDim.y, q = blockIdx.y;
  const int per = (Tn + nq - 1) / nq;
  const int t0 = min(Tn, q * per), t1 = min(Tn, t0 + per);
  if (t0 >= t1) return;
  double* c = smw + (size_t)warp * W::CON_DOUBLES;
  double* acc = c + NC;
  double* Lt = acc + NC;
  double* Kt = Lt + U * B;
  double* Cm = Kt + B * Y;
  double* SF = Cm + R * R;
  double* Nb = SF + N * R;          // full symmetric joint cotangent
  double* Fb = Nb + N * N;
  double* NK = Fb + N * N;
  double* Kb = NK + B * Y;
  double* KD = Kb + B * Y;
  double* T2 = KD + B * U;
  double* Lb = T2 + B * U;
  for (int e = lane; e < NC; e += 32) { c[e] = lcc[(size_t)e * Sc + s]; acc[e] = 0.0; }
  __syncwarp();
  // noise part of one symmetric joint cotangent Nb (CovContrib::noise_part); first = overwrite Kb, else accumulate
  auto noise = [&](bool first) {
    wmm<B, Y, B>(lane, [&](int i, int k) { return Nb[(X + i) * N + X + k]; }, [&](int k, int j) { return Kt[k * Y + j]; },
                 [&](int i, int j, double v) { NK[i * Y + j] = v; });
    for (int e = lane; e < X * X; e += 32) {
      const int i = e / X, j = e - i * X;
      if (j <= i) acc[C::N11 + i * (i + 1) / 2 + j] += Nb[i * N + j];
    }
    wmm<Y, X, B>(lane, [&](int k, int i) { return Kt[i * Y + k]; }, [&](int i, int j) { return Nb[(X + i) * N + j]; },
                 [&](int k, int j, double v) { acc[C::FN + k * X + j] += 2.0 * v; });
    __syncwarp();
    for (int e = lane; e < Y * Y; e += 32) {
      const int k = e / Y, m = e - k * Y;
      if (m > k) continue;
      double a = 0.0;
      for (int i = 0; i < B; ++i) a += Kt[i * Y + k] * NK[i * Y + m] + Kt[i * Y + m] * NK[i * Y + k];
      acc[C::Om + k * (k + 1) / 2 + m] += 0.5 * a;
    }
    for (int e = lane; e < B * Y; e += 32) {
      const int i = e / Y, k = e - i * Y;
      double a = 0.0;
      for (int j = 0; j < X; ++j) a += Nb[(X + i) * N + j] * c[C::FN + k * X + j];
      for (int m = 0; m < Y; ++m) a += NK[i * Y + m] * c[C::Om + sidx(m, k)];
      Kb[e] = (first ? 0.0 : Kb[e]) + 2.0 * a;
    }
    __syncwarp();
  };
  auto load_sym_packed = [&](const double* src) {
    for (int e = lane; e < N * N; e += 32) {
      const int i = e / N, j = e - i * N;
      Nb[e] = src[sidx(i, j)];
    }
  };
  for (int t = t0; t < t1; ++t) {
    const float* sm = sums + (s * Tn + t) * (size_t)DM::SUMP;
    const double* cs = Cs + (s * Tn + t) * (size_t)DM::EC;
    const double* sfw = SFW + (s * Tn + t) * (size_t)SR::NSF;
    for (int e = lane; e < U * B; e += 32) Lt[e] = L[((size_t)t * DM::EL + e) * Sc + s];
    for (int e = lane; e < B * Y; e += 32) Kt[e] = K[((size_t)t * DM::EK + e) * Sc + s];
    for (int e = lane; e < R * R; e += 32) {
      const int i = e / R, j = e - i * R;
      Cm[e] = cs[sidx(i, j)];
    }
    for (int e = lane; e < N * R; e += 32) SF[e] = sfw[e];
    load_sym_packed(SGB + (s * Tn + t) * (size_t)SR::NSGB);
    __syncwarp();
    noise(true);
    if (t == 0) {
      load_sym_packed(SGBI + s * (size_t)SR::NSGB);
      __syncwarp();
      noise(false);
    }
    // Fb = trial sums + [0 | 2 SF C]   (CovContrib::fb_row)
    for (int e = lane; e < N * N; e += 32) {
      const int m = e / N, j = e - m * N;
      double a = (double)sm[DM::SUM_F + e];
      if (j >= D) {
        double v = 0.0;
        for (int k = 0; k < R; ++k) v += SF[m * R + k] * Cm[k * R + (j - D)];
        a += 2.0 * v;
      }
      Fb[e] = a;
    }
    wmm<B, U, Y>(lane, [&](int i, int k) { return Kt[i * Y + k]; }, [&](int k, int j) { return c[C::Dm + k * U + j]; },
                 [&](int i, int j, double v) { KD[i * U + j] = c[C::Ba + i * U + j] + v; });
    __syncwarp();
    // pass 0
    for (int e = lane; e < X * X; e += 32) { const int m = e / X, j = e - m * X; acc[C::Ad + e] += Fb[m * N + j]; }
    for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; acc[C::Aa + e] += Fb[(X + i) * N + X + j]; }
    wmm<X, U, B>(lane, [&](int m, int j) { return Fb[m * N + X + j]; }, [&](int j, int k) { return Lt[k * B + j]; },
                 [&](int m, int k, double v) { acc[C::Bd + m * U + k] += v; });
    wmm<B, U, B>(lane, [&](int i, int j) { return Fb[(X + i) * N + X + j]; }, [&](int j, int k) { return Lt[k * B + j]; },
                 [&](int i, int k, double v) { T2[i * U + k] = v; acc[C::Ba + i * U + k] += v; });
    wmm<Y, X, B>(lane, [&](int k, int i) { return Kt[i * Y + k]; }, [&](int i, int j) { return Fb[(X + i) * N + j]; },
                 [&](int k, int j, double v) { acc[C::FAd + k * X + j] += v; });
    for (int e = lane; e < U * B; e += 32) {
      const int k = e / B, j = e - k * B;
      double a = 0.0;
      for (int m = 0; m < X; ++m) a += c[C::Bd + m * U + k] * Fb[m * N + X + j];
      for (int i = 0; i < B; ++i) a += KD[i * U + k] * Fb[(X + i) * N + X + j];
      Lb[e] = a;
    }
    // pass 1
    wmm<Y, B, B>(lane, [&](int k, int i) { return Kt[i * Y + k]; }, [&](int i, int j) { return Fb[(X + i) * N + X + j]; },
                 [&](int k, int j, double v) { acc[C::FAa + k * B + j] -= v; });
    __syncwarp();                                                          // T2 complete
    wmm<Y, U, B>(lane, [&](int k, int i) { return Kt[i * Y + k]; }, [&](int i, int m) { return T2[i * U + m]; },
                 [&](int k, int m, double v) { acc[C::Dm + k * U + m] += v; });
    for (int e = lane; e < B * Y; e += 32) {
      const int i = e / Y, k = e - i * Y;
      double a = 0.0;
      for (int j = 0; j < X; ++j) a += Fb[(X + i) * N + j] * c[C::FAd + k * X + j];
      for (int j = 0; j < B; ++j) a -= Fb[(X + i) * N + X + j] * c[C::FAa + k * B + j];
      for (int m = 0; m < U; ++m) a += T2[i * U + m] * c[C::Dm + k * U + m];
      Kbar[((size_t)t * DM::EK + e) * Sc + s] = Kb[e] + a;
    }
    for (int e = lane; e < U * B; e += 32) Lbar[((size_t)t * DM::EL + e) * Sc + s] = Lb[e];
    __syncwarp();
  }
  for (int e = lane; e < NC; e += 32) atomicAdd(la + (size_t)e * Sc + s, acc[e]);
}

}  // namespace lqgk
