// lqgk_bigw.cuh -- WARP-PER-SAMPLE covariance kernels of the large-system path (joint dim n > 12).
//
// The joint-system covariance scan and its adjoint carry the big matrices of a large system (n x n, n x r, r x r with
// n = 24, r = 22 at config c4): one warp owns one parameter sample, the matrices live in that warp's slice of shared memory
// (FP64, row-major), every matrix product is spread over the 32 lanes by output element, and dependent products are
// separated by __syncwarp().  Per-step inputs/outputs use SAMPLE-MAJOR workspace arrays ([sample][t][element]) so a
// warp's loads and stores are contiguous; the small per-step gain arrays (L, K, Lbar, Kbar)
// are sample-major as well (every kernel of the large-system path is warp-per-sample).
//
// Mathematics: identical to CovFwd / CovSeqRev / CovContrib in lqgk_core.h (which the host emulation and the small-system
// kernels execute) and to oracle/adjoint_np.py; reference lines: lqg/system.py:163-212, 223-230 and their reverse mode.
#pragma once
#include "lqgk_kernels.cuh"

namespace lqgk {

constexpr int BW_WARPS = 2;   // samples (warps) per CTA: small CTAs pack the 227 KB of shared memory better (18-23 KB per warp)

// FP64 tensor-core tile: D(8x8) = A(8x4) B(4x8) + C, one warp (SASS: DMMA).  Fragments (PTX m8n8k4, A row / B col): lane
// holds A[lane/4][lane%4], B[lane%4][lane/4] and C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Products with a large enough output (see wmm_use_dmma) go to the FP64 tensor cores: the warp walks the 8 x 8 output tiles, every lane
// feeds ONE element of A and of B per 8x8x4 block (2 shared-memory loads per DMMA = 256 FMAs, against 4 loads per 4 FMAs in
// the scalar scheme below, which is shared-memory-latency bound at the 8 warps per SM these kernels reach).  Rows / columns /
// k beyond the matrix are fed as zeros, results outside it are dropped, so any M, N, K works (n = 24, r = 22 at config c4).
#ifndef LQGK_BW_DMMA
#define LQGK_BW_DMMA 1
#endif
template <int M, int N, int K>
__host__ __device__ constexpr bool wmm_use_dmma() {
  return LQGK_BW_DMMA && ((M >= 6 && N >= 6) || ((M >= 6 || N >= 6) && M >= 2 && N >= 2 && K >= 6));
}
template <int M, int N, int K, class FA, class FB, class Out>
__device__ __forceinline__ void wmm_dmma(int lane, FA&& a, FB&& b, Out&& out) {
  constexpr int MT = (M + 7) / 8, NT = (N + 7) / 8, KT = (K + 3) / 4;
  const int g = lane >> 2, q = lane & 3;
  for (int it = 0; it < MT; ++it) {
    const int i = it * 8 + g;
    const int ic = i < M ? i : M - 1;
    for (int jt = 0; jt < NT; ++jt) {
      const int j = jt * 8 + g;                    // B fragment column (the C fragment's columns are jt*8 + 2q, +1)
      const int jc = j < N ? j : N - 1;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll 2
      for (int kt = 0; kt < KT; ++kt) {
        const int k = kt * 4 + q;
        const int kc = k < K ? k : K - 1;
        const double av = a(ic, kc), bv = b(kc, jc);
        dmma884(c0, c1, (i < M && k < K) ? av : 0.0, (j < N && k < K) ? bv : 0.0);
      }
      const int j0 = jt * 8 + 2 * q;
      if (i < M) {
        if (j0 < N) out(i, j0, c0);
        if (j0 + 1 < N) out(i, j0 + 1, c1);
      }
    }
  }
}

// C[M,N] (op)= A B with A addressed as a(i,k), B as b(k,j); out(i, j, value) consumes each element once.  Small outputs: every
// lane owns 2 x 2 tiles of the output: 4 shared-memory loads feed 4 FMAs (an element-per-lane product needs 2 loads per FMA).
template <int M, int N, int K, class FA, class FB, class Out>
__device__ __forceinline__ void wmm(int lane, FA&& a, FB&& b, Out&& out) {
  if constexpr (wmm_use_dmma<M, N, K>()) {
    wmm_dmma<M, N, K>(lane, a, b, out);
    return;
  }
  constexpr int MT = (M + 1) / 2, NT = (N + 1) / 2;
  for (int e = lane; e < MT * NT; e += 32) {
    const int it = e / NT;
    const int i0 = 2 * it, j0 = 2 * (e - it * NT);
    const bool hi = i0 + 1 < M, hj = j0 + 1 < N;
    const int i1 = hi ? i0 + 1 : i0, j1 = hj ? j0 + 1 : j0;          // clamped: the duplicate results are discarded
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const double a0 = a(i0, k), a1 = a(i1, k), b0 = b(k, j0), b1 = b(k, j1);
      c00 += a0 * b0; c01 += a0 * b1; c10 += a1 * b0; c11 += a1 * b1;
    }
    out(i0, j0, c00);
    if (hj) out(i0, j1, c01);
    if (hi) { out(i1, j0, c10); if (hj) out(i1, j1, c11); }
  }
}

// Derived constants of one sample: global block (sample-minor, stride Sc, CLayout order) -> this warp's shared copy in the
// kernel-local layout KC (LqrC / KfC / CovC), and the reverse for cotangent accumulators (FP64 atomics: several warps --
// time ranges, or the Riccati and Kalman adjoints -- add into the same per-sample block).
template <class KC>
__device__ __forceinline__ void bw_load_consts(int lane, const double* g, size_t Sc, double* c, int nseg) {
  for (int i = 0; i < nseg; ++i) {
    int go, lo, len;
    KC::seg(i, go, lo, len);
    for (int e = lane; e < len; e += 32) c[lo + e] = g[(size_t)(go + e) * Sc];
  }
}
template <class KC>
__device__ __forceinline__ void bw_flush_acc(int lane, double* g, size_t Sc, const double* a, int nseg) {
  for (int i = 0; i < nseg; ++i) {
    int go, lo, len;
    KC::seg(i, go, lo, len);
    for (int e = lane; e < len; e += 32) atomicAdd(g + (size_t)(go + e) * Sc, a[lo + e]);
  }
}
// C[M,M] = base(i,j) + sum_k a(i,k) b(k,j), result known symmetric: lower triangle computed (2 x 2 tiles), mirrored.
template <int M, int K, class Base, class FA, class FB>
__device__ __forceinline__ void wmm_sym(int lane, Base&& base, FA&& a, FB&& b, double* Cm) {
  if constexpr (wmm_use_dmma<M, M, K>()) {
    // tensor-core version: the 8 x 8 tiles on and below the diagonal, mirrored (diagonal tiles write both halves: same values
    // up to the rounding of two summation orders are avoided by writing only j <= i there)
    constexpr int MT8 = (M + 7) / 8, KT = (K + 3) / 4;
    const int g = lane >> 2, q = lane & 3;
    for (int it = 0; it < MT8; ++it) {
      const int i = it * 8 + g;
      const int ic = i < M ? i : M - 1;
      for (int jt = 0; jt <= it; ++jt) {
        const int j = jt * 8 + g;
        const int jc = j < M ? j : M - 1;
        const int j0 = jt * 8 + 2 * q;
        double c0 = (i < M && j0 < M) ? base(i, j0) : 0.0, c1 = (i < M && j0 + 1 < M) ? base(i, j0 + 1) : 0.0;
#pragma unroll 2
        for (int kt = 0; kt < KT; ++kt) {
          const int k = kt * 4 + q;
          const int kc = k < K ? k : K - 1;
          const double av = a(ic, kc), bv = b(kc, jc);
          dmma884(c0, c1, (i < M && k < K) ? av : 0.0, (j < M && k < K) ? bv : 0.0);
        }
        if (i < M) {
          if (j0 < M && j0 <= i) { Cm[i * M + j0] = c0; Cm[j0 * M + i] = c0; }
          if (j0 + 1 < M && j0 + 1 <= i) { Cm[i * M + j0 + 1] = c1; Cm[(j0 + 1) * M + i] = c1; }
        }
      }
    }
    return;
  }
  constexpr int MT = (M + 1) / 2;
  for (int e = lane; e < MT * MT; e += 32) {
    const int it = e / MT, jt = e - it * MT;
    if (jt > it) continue;
    const int i0 = 2 * it, j0 = 2 * jt;
    const bool hi = i0 + 1 < M, hj = j0 + 1 < M;
    const int i1 = hi ? i0 + 1 : i0, j1 = hj ? j0 + 1 : j0;
    double c00 = base(i0, j0), c01 = base(i0, j1), c10 = base(i1, j0), c11 = base(i1, j1);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const double a0 = a(i0, k), a1 = a(i1, k), b0 = b(k, j0), b1 = b(k, j1);
      c00 += a0 * b0; c01 += a0 * b1; c10 += a1 * b0; c11 += a1 * b1;
    }
    Cm[i0 * M + j0] = c00; Cm[j0 * M + i0] = c00;
    if (hj && j1 <= i0) { Cm[i0 * M + j1] = c01; Cm[j1 * M + i0] = c01; }     // off-diagonal tiles only (j1 > i0 on the diagonal)
    if (hi) {
      Cm[i1 * M + j0] = c10; Cm[j0 * M + i1] = c10;
      if (hj) { Cm[i1 * M + j1] = c11; Cm[j1 * M + i1] = c11; }
    }
  }
}

template <class DM>
struct BigW {
  static constexpr int X = DM::X, B = DM::B, U = DM::U, Y = DM::Y, D = DM::D, N = DM::N, R = DM::R;
  using C = CovC<DM>;
  using SR = CovSeqRev<DM>;
  static constexpr int NC = C::n;

  // ------------------------------------------------------------------------------------------ forward
  static constexpr int FWD_DOUBLES = NC + 2 * (U * B + B * Y) + B * U + B * Y + 2 * N * N + R * R + N * R + R * D + 3 * D * D + 8;
  static size_t smem_fwd() { return sizeof(double) * FWD_DOUBLES * BW_WARPS; }
  // + two staging buffers for the step inputs (Fu_t, (J_t, S'^-1_t), the float sums tail) filled by cp.async one step ahead
  static constexpr int SEQ_STAGE = (N * R + R * D + D * (D + 1) / 2 + (DM::SUMP - DM::SUM_J) + 1) / 2;   // floats, counted in doubles
  static constexpr int SEQ_DOUBLES = 2 * R * R + 2 * N * R + 5 * R * D + 2 * D * D + N * N + 8 + 2 * ((SEQ_STAGE + 1) & ~1);
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
    double l;
    chol_and_inverse<D>(Lc, Li, &l);
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

// Covariance pass (forward), warp per sample.  cst: derived constants [element][Sc];  L, K, Cs, FU, JS, J0: sample-major;
// rec [s][t][REC].
// GAINS_MINOR: L, K are sample-minor [t][e][Sc] (written by the thread-per-sample Riccati / Kalman kernels of the small systems).
template <class DM, bool GAINS_MINOR = false>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_cov_fwd(const double* cst, size_t Sc, int npad, int Tn, const double* L, const double* K,
                                                            int save_adj, double* Cs, lin_t* FU, lin_t* JS, double* J0, float* rec,
                                                            int t0 = 0, int t1 = -1) {
  // [t0, t1): time range of this launch (t1 < 0: up to Tn); t0 > 0 continues from the C_{t0} saved in Cs (save_adj on)
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
  double* LtN = Kt + B * Y;         // next step's gains, in flight (cp.async) while this step is computed
  double* KtN = LtN + U * B;
  double* KD = KtN + B * Y;
  double* KO = KD + B * U;
  double* Fj = KO + B * Y;
  double* Nj = Fj + N * N;
  double* Sig = Nj;                 // Sig' = N + Fu C Fu^T is accumulated in place (every element is read and written by one lane)
  double* Cm = Nj + N * N;
  double* T1 = Cm + R * R;
  double* Z = T1 + N * R;
  double* Linv = Z + R * D;
  double* ld = Linv + 3 * D * D;
  auto gi = [&](int t, int E, int e) -> size_t { return GAINS_MINOR ? ((size_t)t * E + e) * Sc + s : (s * Tn + t) * (size_t)E + e; };
  bw_load_consts<C>(lane, cst + s, Sc, c, C::NSEG);
  if (t1 < 0) t1 = Tn;
  if (t0 == 0) {
    for (int e = lane; e < B * Y; e += 32) Kt[e] = K[gi(0, DM::EK, e)];
    __syncwarp();
    bw_joint_N<DM>(lane, c, Kt, KO, Nj);
    __syncwarp();
    bw_condition<DM>(lane, Nj, Linv, Z, ld, [&](int e, double v) { if (save_adj) J0[s * (R * D) + e] = v; }, Cm);
  } else {
    const double* cs = Cs + (s * Tn + t0) * (size_t)DM::EC;
    for (int e = lane; e < R * R; e += 32) {
      const int i = e / R, j = e - i * R;
      Cm[e] = cs[sidx(i, j)];
    }
    __syncwarp();
  }
  float* recs = rec + s * (size_t)Tn * DM::REC;
  // gains of step t arrive by per-lane 8-byte cp.async one step ahead (an unprefetched global round trip per step was a third of
  // this latency-bound kernel's step time)
  auto fetch_gains = [&](int t, double* Ld, double* Kd) {
    for (int e = lane; e < U * B; e += 32) cp_async<8>(Ld + e, L + gi(t, DM::EL, e));
    for (int e = lane; e < B * Y; e += 32) cp_async<8>(Kd + e, K + gi(t, DM::EK, e));
    cp_async_commit();
  };
  __syncwarp();                                    // (Kt was read by the initialisation above)
  if (t0 < t1) fetch_gains(t0, Lt, Kt);
  for (int t = t0; t < t1; ++t) {
    cp_async_wait<0>();
    __syncwarp();                                  // gains of step t visible to all lanes; the other buffer is free
    if (t + 1 < t1) fetch_gains(t + 1, LtN, KtN);
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
    lin_t* fu = FU + (s * Tn + t) * (size_t)SR::NSF;
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
    wmm_sym<N, R>(lane, [&](int i, int j) { return Nj[i * N + j]; }, [&](int i, int k) { return T1[i * R + k]; },
                  [&](int k, int j) { return Fj[j * N + D + k]; }, Sig);   // Sig' = Fu C Fu^T + N (lower, mirrored)
    __syncwarp();
    lin_t* js = JS + (s * Tn + t) * (size_t)SR::NJS;
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
    { double* tp = Lt; Lt = LtN; LtN = tp; tp = Kt; Kt = KtN; KtN = tp; }
  }
  if (save_adj && t1 < Tn) {   // where the next time segment continues from
    double* cs = Cs + (s * Tn + t1) * (size_t)DM::EC;
    for (int e = lane; e < R * R; e += 32) {
      const int i = e / R, j = e - i * R;
      if (j <= i) cs[i * (i + 1) / 2 + j] = Cm[e];
    }
  }
}

// Sequential covariance adjoint (CovSeqRev), warp per sample, t descending.
template <class DM>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_cov_seq_rev(int npad, int Tn, int Ntr, const float* w, const lin_t* FU, const lin_t* JS,
                                                                const double* J0, const float* sums, lin_t* SGB, double* SGBI, lin_t* SFW,
                                                                int t0 = 0, int t1 = -1, double* carry = nullptr) {
  // [t0, t1) walked downwards (t1 < 0: from Tn); `carry` ([s][R*R]) hands Cb from one launch to the next (pipelined sequence)
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
  double* Sg = Ss + D * D;          // full symmetric joint cotangent Sgb_t (N x N)
  double* stg0 = Sg + N * N;        // two staging buffers: [Fu (N R) | js (NJS) | float sums tail]
  double* stg1 = stg0 + ((W::SEQ_STAGE + 1) & ~1);
  constexpr int NFT = DM::SUMP - DM::SUM_J;   // floats of the sums row this kernel reads (Jb, Wv)
  auto fetch_step = [&](int t, double* stg) {
    const lin_t* fu = FU + (s * Tn + t) * (size_t)SR::NSF;
    const lin_t* js = JS + (s * Tn + t) * (size_t)SR::NJS;
    const float* sm = sums + (s * Tn + t) * (size_t)DM::SUMP + DM::SUM_J;
    float* sf = reinterpret_cast<float*>(stg);              // staging layout (floats): [Fu (N R) | js (NJS) | sums tail (NFT)]
    for (int e = lane; e < N * R; e += 32) cp_async<4>(sf + e, fu + e);
    for (int e = lane; e < SR::NJS; e += 32) cp_async<4>(sf + N * R + e, js + e);
    for (int e = lane; e < NFT; e += 32) cp_async<4>(sf + N * R + SR::NJS + e, sm + e);
    cp_async_commit();
  };
  double sw = 0.0;
  for (int i = lane; i < Ntr; i += 32) sw += (double)w[s * Ntr + i];
  for (int m = 16; m >= 1; m >>= 1) sw += __shfl_xor_sync(FULL, sw, m);
  if (t1 < 0) t1 = Tn;
  for (int e = lane; e < R * R; e += 32) Cb[e] = t1 == Tn ? 0.0 : carry[s * (R * R) + e];
  __syncwarp();
  if (t1 > t0) fetch_step(t1 - 1, stg0);
  for (int t = t1 - 1; t >= t0; --t) {
    cp_async_wait<0>();
    __syncwarp();                                  // inputs of step t staged; the other staging buffer is free
    if (t > t0) fetch_step(t - 1, stg1);
    {
      const float* fu = reinterpret_cast<const float*>(stg0);
      const float* js = fu + N * R;
      const float* smt = js + SR::NJS;                                              // sums[SUM_J ..]
      for (int e = lane; e < N * R; e += 32) Fu[e] = (double)fu[e];
      for (int e = lane; e < R * D; e += 32) { J[e] = (double)js[e]; Jb[e] = (double)smt[e]; }
      for (int e = lane; e < D * D; e += 32) {
        const int i = e / D, j = e - i * D;
        Sinv[e] = (double)js[R * D + sidx(i, j)];
      }
    }
    const float* sm = reinterpret_cast<const float*>(stg0) + N * R + SR::NJS - DM::SUM_J;   // so that sm[SUM_W + ..] below reads the staged tail
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
    lin_t* sgb = SGB + (s * Tn + t) * (size_t)SR::NSGB;
    for (int e = lane; e < N * N; e += 32) {                               // Sgb_t: full symmetric in shared memory, packed lower out
      const int i = e / N, j = e - i * N;
      if (j > i) continue;
      const double v = i < D ? Ss[i * D + j] : (j < D ? Bh[(i - D) * D + j] : Cb[(i - D) * R + (j - D)]);
      sgb[i * (i + 1) / 2 + j] = (lin_t)v;
      Sg[i * N + j] = v;
      Sg[j * N + i] = v;
    }
    __syncwarp();
    lin_t* sfw = SFW + (s * Tn + t) * (size_t)SR::NSF;
    wmm<N, R, N>(lane, [&](int m, int k) { return Sg[m * N + k]; }, [&](int k, int cc) { return Fu[k * R + cc]; },
                 [&](int m, int cc, double v) { SF[m * R + cc] = v; sfw[m * R + cc] = (lin_t)v; });   // SF_t = Sgb_t Fu_t
    __syncwarp();
    wmm_sym<R, N>(lane, [&](int, int) { return 0.0; }, [&](int a, int m) { return Fu[m * R + a]; },
                  [&](int m, int b) { return SF[m * R + b]; }, Cn);        // Cb <- Fu^T SF (lower, mirrored)
    __syncwarp();
    double* tmp = Cb; Cb = Cn; Cn = tmp;
    tmp = stg0; stg0 = stg1; stg1 = tmp;
  }
  if (t0 > 0) {
    for (int e = lane; e < R * R; e += 32) carry[s * (R * R) + e] = Cb[e];
    return;
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

// Time-parallel contraction (CovContrib pass 0 + pass 1), warp per (sample, time range).  gacc: the per-sample cotangent
// accumulators [element][Sc] (CLayout order, zeroed by the caller), updated with one FP64 atomicAdd per element and warp.
// GAINS_MINOR: L, K, Lbar, Kbar sample-minor (small systems); KbarF (their second Kalman-cotangent part) is then zeroed.
template <class DM, bool GAINS_MINOR = false>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_cov_contrib(const double* cst, size_t Sc, int npad, int Tn, const double* L, const double* K,
                                                                const double* Cs, const lin_t* SGB, const double* SGBI, const lin_t* SFW,
                                                                const float* sums, double* gacc, double* Lbar, double* Kbar, double* KbarF,
                                                                int ta = 0, int tb = -1) {
  using W = BigW<DM>;
  using C = typename W::C;
  using SR = typename W::SR;
  constexpr int X = W::X, B = W::B, U = W::U, Y = W::Y, D = W::D, N = W::N, R = W::R, NC = W::NC;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  if (tb < 0) tb = Tn;
  const int nq = gridDim.y, q = blockIdx.y;       // the launch covers [ta, tb), cut into nq time ranges
  const int per = (tb - ta + nq - 1) / nq;
  const int t0 = min(tb, ta + q * per), t1 = min(tb, t0 + per);
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
  auto gi = [&](int t, int E, int e) -> size_t { return GAINS_MINOR ? ((size_t)t * E + e) * Sc + s : (s * Tn + t) * (size_t)E + e; };
  bw_load_consts<C>(lane, cst + s, Sc, c, C::NSEG);
  for (int e = lane; e < NC; e += 32) acc[e] = 0.0;
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
  auto load_sym_packed = [&](const auto* src) {
    for (int e = lane; e < N * N; e += 32) {
      const int i = e / N, j = e - i * N;
      Nb[e] = (double)src[sidx(i, j)];
    }
  };
  for (int t = t0; t < t1; ++t) {
    const float* sm = sums + (s * Tn + t) * (size_t)DM::SUMP;
    const double* cs = Cs + (s * Tn + t) * (size_t)DM::EC;
    const lin_t* sfw = SFW + (s * Tn + t) * (size_t)SR::NSF;
    for (int e = lane; e < U * B; e += 32) Lt[e] = L[gi(t, DM::EL, e)];
    for (int e = lane; e < B * Y; e += 32) Kt[e] = K[gi(t, DM::EK, e)];
    for (int e = lane; e < R * R; e += 32) {
      const int i = e / R, j = e - i * R;
      Cm[e] = cs[sidx(i, j)];
    }
    for (int e = lane; e < N * R; e += 32) SF[e] = (double)sfw[e];
    load_sym_packed(SGB + (s * Tn + t) * (size_t)SR::NSGB);
    __syncwarp();
    noise(true);
    if (t == 0) {
      load_sym_packed(SGBI + s * (size_t)SR::NSGB);
      __syncwarp();
      noise(false);
    }
    // Fb = trial sums + [0 | 2 SF C]   (CovContrib::fb_row)
    for (int e = lane; e < N * D; e += 32) { const int m = e / D, j = e - m * D; Fb[m * N + j] = (double)sm[DM::SUM_F + m * N + j]; }
    wmm<N, R, R>(lane, [&](int m, int k) { return SF[m * R + k]; }, [&](int k, int j) { return Cm[k * R + j]; },
                 [&](int m, int j, double v) { Fb[m * N + D + j] = (double)sm[DM::SUM_F + m * N + D + j] + 2.0 * v; });
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
      Kbar[gi(t, DM::EK, e)] = Kb[e] + a;
      if (GAINS_MINOR && KbarF) KbarF[gi(t, DM::EK, e)] = 0.0;
    }
    for (int e = lane; e < U * B; e += 32) Lbar[gi(t, DM::EL, e)] = Lb[e];
    __syncwarp();
  }
  bw_flush_acc<C>(lane, gacc + s, Sc, acc, C::NSEG);
}

// =========================================================================================== Riccati / Kalman sweeps
// Warp-per-sample versions of LqrFwd / KfFwd / KfRev / LqrRev (lqgk_core.h): b x b matrices in shared memory, the u x u and
// y x y factorisations (u <= 2, y <= 4) by lane 0 with the scalar templates.  L, K, l, H, Sric, Pkf, Lbar, Kbar: sample-major.
template <class DM>
struct BigG {
  static constexpr int B = DM::B, U = DM::U, Y = DM::Y;
  static constexpr int LQR_DOUBLES = LqrC<DM>::n_affine + LqrC<DM>::n + 8 * B * B + 10 * B * U + 6 * U * U + 4 * B + 4 * U + 8;
  static constexpr int KF_DOUBLES = 2 * KfC<DM>::n + 7 * B * B + 9 * B * Y + 4 * Y * Y + 8;
  static size_t smem_lqr() { return sizeof(double) * LQR_DOUBLES * BW_WARPS; }
  static size_t smem_kf() { return sizeof(double) * KF_DOUBLES * BW_WARPS; }
};

// H = R + B^T S B, G = B^T S A (+P), shift, Ht^-1 from S (full symmetric) -- shared by the forward sweep and its adjoint.
// SA, SB, H, G, Hi: shared-memory outputs; returns the eigen-shift (lqr.py:22-28).  Ends synchronised.
template <class DM, bool AFFINE>
__device__ __forceinline__ double bw_lqr_common(int lane, const double* c, double eps, const double* S, double* SA, double* SB, double* H,
                                                double* G, double* Hi, double* shift_sm) {
  using C = LqrC<DM>;
  constexpr int B = DM::B, U = DM::U;
  wmm<B, B, B>(lane, [&](int i, int k) { return S[i * B + k]; }, [&](int k, int j) { return c[C::Aa + k * B + j]; },
               [&](int i, int j, double v) { SA[i * B + j] = v; });
  wmm<B, U, B>(lane, [&](int i, int k) { return S[i * B + k]; }, [&](int k, int j) { return c[C::Ba + k * U + j]; },
               [&](int i, int j, double v) { SB[i * U + j] = v; });
  __syncwarp();
  wmm_sym<U, B>(lane, [&](int i, int j) { return c[C::R + sidx(i, j)]; }, [&](int i, int k) { return c[C::Ba + k * U + i]; },
                [&](int k, int j) { return SB[k * U + j]; }, H);
  wmm<U, B, B>(lane, [&](int i, int k) { return c[C::Ba + k * U + i]; }, [&](int k, int j) { return SA[k * B + j]; },
               [&](int i, int j, double v) { G[i * B + j] = v + (AFFINE ? c[C::P + i * B + j] : 0.0); });
  __syncwarp();
  if (lane == 0) {
    double Hl[U * U], Lc[U * U], Li[U * U], Hv[U * U];
    for (int i = 0; i < U * U; ++i) Hl[i] = H[i];
    double shift = eps - lambda_min<U>(Hl);
    shift = shift > 0.0 ? shift : 0.0;
    for (int i = 0; i < U * U; ++i) Lc[i] = Hl[i];
    for (int i = 0; i < U; ++i) Lc[i * U + i] += shift;
    chol_and_inverse<U>(Lc, Li);
    mm_tn<U, U, U>(Li, Li, Hv);
    for (int i = 0; i < U * U; ++i) Hi[i] = Hv[i];
    *shift_sm = shift;
  }
  __syncwarp();
  return *shift_sm;
}

template <class DM, bool AFFINE>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_lqr_fwd(const double* cst, size_t Sc, int npad, int Tn, double eps, double* L, int save_S,
                                                            double* Sric, double* l, double* H_out) {
  using C = LqrC<DM>;
  constexpr int B = DM::B, U = DM::U;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  double* c = smw + (size_t)warp * BigG<DM>::LQR_DOUBLES;
  double* S = c + C::n_affine + C::n;
  double* Sn = S + B * B;
  double* SA = Sn + B * B;
  double* SB = SA + 6 * B * B;            // (unused b x b slots belong to the adjoint kernel)
  double* G = SB + B * U;
  double* Lm = G + B * U;
  double* HL = Lm + B * U;
  double* H = HL + 7 * B * U;
  double* Hi = H + U * U;
  double* sv = Hi + 5 * U * U;            // s (B) | sn (B) | 2 B spare
  double* gv = sv + 4 * B;                // g (U) | l (U) | Hl (U) | spare
  double* shift_sm = gv + 4 * U;
  bw_load_consts<C>(lane, cst + s, Sc, c, AFFINE ? C::NSEG_AFF : C::NSEG);
  __syncwarp();
  for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; S[e] = c[C::Qf + sidx(i, j)]; }
  if (AFFINE) { for (int e = lane; e < B; e += 32) sv[e] = c[C::qf + e]; }
  __syncwarp();
  for (int t = Tn - 1; t >= 0; --t) {
    if (save_S) {
      double* so = Sric + (s * Tn + t) * (size_t)DM::ES;
      for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; if (j <= i) so[i * (i + 1) / 2 + j] = S[e]; }
    }
    const double shift = bw_lqr_common<DM, AFFINE>(lane, c, eps, S, SA, SB, H, G, Hi, shift_sm);
    if (AFFINE) {
      for (int e = lane; e < U; e += 32) {
        double a = c[C::r + e];
        for (int k = 0; k < B; ++k) a += c[C::Ba + k * U + e] * sv[k];
        gv[e] = a;
      }
    }
    wmm<U, B, U>(lane, [&](int i, int k) { return Hi[i * U + k]; }, [&](int k, int j) { return G[k * B + j]; },
                 [&](int i, int j, double v) { Lm[i * B + j] = -v; L[(s * Tn + t) * (size_t)DM::EL + i * B + j] = -v; });
    __syncwarp();
    wmm<U, B, U>(lane, [&](int i, int k) { return H[i * U + k]; }, [&](int k, int j) { return Lm[k * B + j]; },
                 [&](int i, int j, double v) { HL[i * B + j] = v + 2.0 * G[i * B + j]; });
    if (AFFINE) {
      for (int e = lane; e < U; e += 32) {
        double a = 0.0;
        for (int k = 0; k < U; ++k) a += Hi[e * U + k] * gv[k];
        gv[U + e] = -a;
        l[(s * Tn + t) * (size_t)U + e] = -a;
      }
      for (int e = lane; e < U * U; e += 32) {
        const int i = e / U, j = e - i * U;
        H_out[(s * Tn + t) * (size_t)(U * U) + e] = H[e] + (i == j ? shift : 0.0);
      }
    }
    __syncwarp();
    // S <- Q + A^T S A + sym(L^T (H L + 2 G))                                    lqr.py:33
    for (int e = lane; e < B * B; e += 32) {
      const int i = e / B, j = e - i * B;
      if (j > i) continue;
      double a = c[C::Q + sidx(i, j)];
      for (int k = 0; k < B; ++k) a += c[C::Aa + k * B + i] * SA[k * B + j];
      double b2 = 0.0;
      for (int k = 0; k < U; ++k) b2 += Lm[k * B + i] * HL[k * B + j] + Lm[k * B + j] * HL[k * B + i];
      a += 0.5 * b2;
      Sn[i * B + j] = a;
      Sn[j * B + i] = a;
    }
    if (AFFINE) {
      for (int e = lane; e < U; e += 32) {
        double a = 0.0;
        for (int k = 0; k < U; ++k) a += H[e * U + k] * gv[U + k];
        gv[2 * U + e] = a;                                                     // H l
      }
      __syncwarp();
      for (int e = lane; e < B; e += 32) {                                     // lqr.py:34
        double a = c[C::q + e];
        for (int k = 0; k < B; ++k) a += c[C::Aa + k * B + e] * sv[k];
        for (int k = 0; k < U; ++k) a += G[k * B + e] * gv[U + k] + Lm[k * B + e] * (gv[2 * U + k] + gv[k]);
        sv[B + e] = a;
      }
      __syncwarp();
      for (int e = lane; e < B; e += 32) sv[e] = sv[B + e];
    }
    __syncwarp();
    double* tmp = S; S = Sn; Sn = tmp;
  }
}

// Kalman gain from P (full symmetric): Pp, M = F Pp, Gi = Gm^-1, K (KfFwd::gain, kf.py:10-12).  AP: b x b scratch.
template <class DM>
__device__ __forceinline__ void bw_kf_gain(int lane, const double* c, const double* P, double* AP, double* Pp, double* M, double* Gm, double* Gi,
                                           double* K) {
  using C = KfC<DM>;
  constexpr int B = DM::B, Y = DM::Y;
  wmm<B, B, B>(lane, [&](int i, int k) { return c[C::Aa + i * B + k]; }, [&](int k, int j) { return P[k * B + j]; },
               [&](int i, int j, double v) { AP[i * B + j] = v; });
  __syncwarp();
  wmm_sym<B, B>(lane, [&](int i, int j) { return c[C::VVa + sidx(i, j)]; }, [&](int i, int k) { return AP[i * B + k]; },
                [&](int k, int j) { return c[C::Aa + j * B + k]; }, Pp);
  __syncwarp();
  wmm<Y, B, B>(lane, [&](int i, int k) { return c[C::Fa + i * B + k]; }, [&](int k, int j) { return Pp[k * B + j]; },
               [&](int i, int j, double v) { M[i * B + j] = v; });
  __syncwarp();
  wmm_sym<Y, B>(lane, [&](int i, int j) { return c[C::WWa + sidx(i, j)]; }, [&](int i, int k) { return M[i * B + k]; },
                [&](int k, int j) { return c[C::Fa + j * B + k]; }, Gm);
  __syncwarp();
  if (lane == 0) {
    double Gl[Y * Y], Li[Y * Y], Gv[Y * Y];
    for (int i = 0; i < Y * Y; ++i) Gl[i] = Gm[i];
    chol_and_inverse<Y>(Gl, Li);
    mm_tn_sym<Y, Y>(Li, Li, Gv);
    for (int i = 0; i < Y * Y; ++i) Gi[i] = Gv[i];
  }
  __syncwarp();
  wmm<B, Y, Y>(lane, [&](int i, int k) { return M[k * B + i]; }, [&](int k, int j) { return Gi[k * Y + j]; },
               [&](int i, int j, double v) { K[i * Y + j] = v; });
  __syncwarp();
}

template <class DM>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_kf_fwd(const double* cst, size_t Sc, int npad, int Tn, double* Kout, int save_P, double* Pkf) {
  using C = KfC<DM>;
  constexpr int B = DM::B, Y = DM::Y;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  double* c = smw + (size_t)warp * BigG<DM>::KF_DOUBLES;
  double* P = c + 2 * C::n;
  double* Pn = P + B * B;
  double* AP = Pn + B * B;
  double* Pp = AP + B * B;
  double* M = Pp + 4 * B * B;
  double* K = M + B * Y;
  double* Gm = K + 8 * B * Y;
  double* Gi = Gm + Y * Y;
  bw_load_consts<C>(lane, cst + s, Sc, c, C::NSEG);
  __syncwarp();
  for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; P[e] = c[C::Sig0 + sidx(i, j)]; }
  __syncwarp();
  for (int t = 0; t < Tn; ++t) {
    if (save_P) {
      double* po = Pkf + (s * Tn + t) * (size_t)DM::EP;
      for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; if (j <= i) po[i * (i + 1) / 2 + j] = P[e]; }
    }
    bw_kf_gain<DM>(lane, c, P, AP, Pp, M, Gm, Gi, K);
    for (int e = lane; e < B * Y; e += 32) Kout[(s * Tn + t) * (size_t)DM::EK + e] = K[e];
    for (int e = lane; e < B * B; e += 32) {                                   // P <- Pp - K M (symmetric)   kf.py:14
      const int i = e / B, j = e - i * B;
      if (j > i) continue;
      double a = Pp[i * B + j];
      for (int k = 0; k < Y; ++k) a -= 0.5 * (K[i * Y + k] * M[k * B + j] + K[j * Y + k] * M[k * B + i]);
      Pn[i * B + j] = a;
      Pn[j * B + i] = a;
    }
    __syncwarp();
    double* tmp = P; P = Pn; Pn = tmp;
  }
}

// Kalman-gain adjoint (KfRev), t descending.  gacc: per-sample cotangent accumulators [element][Sc] (atomicAdd at the end).
template <class DM>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_kf_rev(const double* cst, size_t Sc, int npad, int Tn, const double* Pkf, const double* Kbar,
                                                           double* gacc) {
  using C = KfC<DM>;
  constexpr int B = DM::B, Y = DM::Y;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  double* c = smw + (size_t)warp * BigG<DM>::KF_DOUBLES;
  double* acc = c + C::n;
  double* P = acc + C::n;
  double* Pnb = P + B * B;
  double* AP = Pnb + B * B;
  double* Pp = AP + B * B;
  double* Ppb = Pp + B * B;
  double* PA = Ppb + B * B;
  double* Pn2 = PA + B * B;
  double* M = Pn2 + B * B;
  double* K = M + B * Y;
  double* Kb = K + B * Y;
  double* Ktot = Kb + B * Y;
  double* Yv = Ktot + B * Y;
  double* Mb = Yv + B * Y;
  double* GF = Mb + B * Y;
  double* Gm = GF + 3 * B * Y;
  double* Gi = Gm + Y * Y;
  double* Gmb = Gi + Y * Y;
  bw_load_consts<C>(lane, cst + s, Sc, c, C::NSEG);
  for (int e = lane; e < C::n; e += 32) acc[e] = 0.0;
  for (int e = lane; e < B * B; e += 32) Pnb[e] = 0.0;
  __syncwarp();
  for (int t = Tn - 1; t >= 0; --t) {
    const double* po = Pkf + (s * Tn + t) * (size_t)DM::EP;
    for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; P[e] = po[sidx(i, j)]; }
    for (int e = lane; e < B * Y; e += 32) Kb[e] = Kbar[(s * Tn + t) * (size_t)DM::EK + e];
    __syncwarp();
    bw_kf_gain<DM>(lane, c, P, AP, Pp, M, Gm, Gi, K);
    // Ktot = Kb - Pnb M^T ; Y = Ktot Gi
    wmm<B, Y, B>(lane, [&](int i, int j) { return Pnb[i * B + j]; }, [&](int j, int k) { return M[k * B + j]; },
                 [&](int i, int k, double v) { Ktot[i * Y + k] = Kb[i * Y + k] - v; });
    __syncwarp();
    wmm<B, Y, Y>(lane, [&](int i, int k) { return Ktot[i * Y + k]; }, [&](int k, int j) { return Gi[k * Y + j]; },
                 [&](int i, int j, double v) { Yv[i * Y + j] = v; });
    __syncwarp();
    // Mb = Y^T - K^T Pnb (Y x B) ; Gmb = -sym(K^T Y) (Y x Y)
    wmm<Y, B, B>(lane, [&](int k, int i) { return K[i * Y + k]; }, [&](int i, int j) { return Pnb[i * B + j]; },
                 [&](int k, int j, double v) { Mb[k * B + j] = Yv[j * Y + k] - v; });
    for (int e = lane; e < Y * Y; e += 32) {
      const int k = e / Y, m = e - k * Y;
      if (m > k) continue;
      double a = 0.0;
      for (int i = 0; i < B; ++i) a += K[i * Y + k] * Yv[i * Y + m] + K[i * Y + m] * Yv[i * Y + k];
      Gmb[k * Y + m] = -0.5 * a;
      Gmb[m * Y + k] = -0.5 * a;
      acc[C::WWa + k * (k + 1) / 2 + m] += -0.5 * a;
    }
    __syncwarp();
    // Fa += Mb Pp + 2 Gmb M ; GF = Gmb F
    for (int e = lane; e < Y * B; e += 32) {
      const int k = e / B, j = e - k * B;
      double a = 0.0, g = 0.0;
      for (int i = 0; i < B; ++i) a += Mb[k * B + i] * Pp[i * B + j];
      for (int m = 0; m < Y; ++m) { a += 2.0 * Gmb[k * Y + m] * M[m * B + j]; g += Gmb[k * Y + m] * c[C::Fa + m * B + j]; }
      acc[C::Fa + e] += a;
      GF[e] = g;
    }
    __syncwarp();
    // Ppb = Pnb + sym(F^T Mb) + F^T Gmb F
    for (int e = lane; e < B * B; e += 32) {
      const int i = e / B, j = e - i * B;
      if (j > i) continue;
      double a = Pnb[i * B + j];
      for (int k = 0; k < Y; ++k)
        a += 0.5 * (c[C::Fa + k * B + i] * Mb[k * B + j] + c[C::Fa + k * B + j] * Mb[k * B + i]) + c[C::Fa + k * B + i] * GF[k * B + j];
      Ppb[i * B + j] = a;
      Ppb[j * B + i] = a;
      acc[C::VVa + i * (i + 1) / 2 + j] += a;
    }
    __syncwarp();
    // Aa += 2 Ppb (A P) ; Pnb <- A^T (Ppb A)        (AP = A P is still valid from bw_kf_gain)
    wmm<B, B, B>(lane, [&](int i, int k) { return Ppb[i * B + k]; }, [&](int k, int j) { return c[C::Aa + k * B + j]; },
                 [&](int i, int j, double v) { PA[i * B + j] = v; });
    wmm<B, B, B>(lane, [&](int i, int k) { return Ppb[i * B + k]; }, [&](int k, int j) { return AP[k * B + j]; },
                 [&](int i, int j, double v) { acc[C::Aa + i * B + j] += 2.0 * v; });
    __syncwarp();
    wmm_sym<B, B>(lane, [&](int, int) { return 0.0; }, [&](int i, int k) { return c[C::Aa + k * B + i]; },
                  [&](int k, int j) { return PA[k * B + j]; }, Pn2);
    __syncwarp();
    double* tmp = Pnb; Pnb = Pn2; Pn2 = tmp;
  }
  for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; if (j <= i) acc[C::Sig0 + i * (i + 1) / 2 + j] += Pnb[e]; }
  __syncwarp();
  bw_flush_acc<C>(lane, gacc + s, Sc, acc, C::NSEG);
}

// Riccati adjoint (LqrRev), t ascending; the eigen-shift is a constant of the adjoint (zero whenever R > 0).
template <class DM>
__global__ void __launch_bounds__(32 * BW_WARPS) kw_lqr_rev(const double* cst, size_t Sc, int npad, int Tn, double eps, const double* L,
                                                            const double* Sric, const double* Lbar, double* gacc) {
  using C = LqrC<DM>;
  constexpr int B = DM::B, U = DM::U;
  extern __shared__ __align__(16) double smw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * BW_WARPS + warp;
  if (s >= (size_t)npad) return;
  double* c = smw + (size_t)warp * BigG<DM>::LQR_DOUBLES;
  double* acc = c + C::n_affine;
  double* S = acc + C::n;
  double* Sn = S + B * B;
  double* SA = Sn + B * B;
  double* AS = SA + B * B;
  double* BG = AS + B * B;
  double* Sb = BG + B * B;
  double* spare = Sb + B * B;             // 2 b x b spare
  double* SB = spare + 2 * B * B;
  double* G = SB + B * U;
  double* Lm = G + B * U;
  double* Lbr = Lm + B * U;
  double* LS = Lbr + B * U;
  double* HLG = LS + B * U;
  double* Lb = HLG + B * U;
  double* HiLb = Lb + B * U;
  double* Gb = HiLb + B * U;
  double* BH = Gb + B * U;
  double* H = BH + B * U;
  double* Hi = H + U * U;
  double* Hb = Hi + U * U;
  double* shift_sm = Hb + 4 * U * U + 4 * B + 4 * U;
  bw_load_consts<C>(lane, cst + s, Sc, c, C::NSEG);
  for (int e = lane; e < C::n; e += 32) acc[e] = 0.0;
  for (int e = lane; e < B * B; e += 32) Sn[e] = 0.0;
  __syncwarp();
  for (int t = 0; t < Tn; ++t) {
    const double* so = Sric + (s * Tn + t) * (size_t)DM::ES;
    for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; S[e] = so[sidx(i, j)]; }
    for (int e = lane; e < U * B; e += 32) { Lm[e] = L[(s * Tn + t) * (size_t)DM::EL + e]; Lbr[e] = Lbar[(s * Tn + t) * (size_t)DM::EL + e]; }
    __syncwarp();
    bw_lqr_common<DM, false>(lane, c, eps, S, SA, SB, H, G, Hi, shift_sm);
    // Q += Sn ; Aa += 2 SA Sn ; LS = L Sn ; HLG = H L + G
    for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; if (j <= i) acc[C::Q + i * (i + 1) / 2 + j] += Sn[e]; }
    wmm<B, B, B>(lane, [&](int i, int k) { return SA[i * B + k]; }, [&](int k, int j) { return Sn[k * B + j]; },
                 [&](int i, int j, double v) { acc[C::Aa + i * B + j] += 2.0 * v; });
    wmm<U, B, B>(lane, [&](int i, int k) { return Lm[i * B + k]; }, [&](int k, int j) { return Sn[k * B + j]; },
                 [&](int i, int j, double v) { LS[i * B + j] = v; });
    wmm<U, B, U>(lane, [&](int i, int k) { return H[i * U + k]; }, [&](int k, int j) { return Lm[k * B + j]; },
                 [&](int i, int j, double v) { HLG[i * B + j] = v + G[i * B + j]; });
    __syncwarp();
    // Lb = Lbar + 2 HLG Sn ; Hb = LS L^T
    wmm<U, B, B>(lane, [&](int i, int k) { return HLG[i * B + k]; }, [&](int k, int j) { return Sn[k * B + j]; },
                 [&](int i, int j, double v) { Lb[i * B + j] = Lbr[i * B + j] + 2.0 * v; });
    __syncwarp();
    // HiLb = Ht^-1 Lb ; Gb = 2 LS - HiLb
    wmm<U, B, U>(lane, [&](int i, int k) { return Hi[i * U + k]; }, [&](int k, int j) { return Lb[k * B + j]; },
                 [&](int i, int j, double v) { HiLb[i * B + j] = v; Gb[i * B + j] = 2.0 * LS[i * B + j] - v; });
    __syncwarp();
    // Hb = sym(LS L^T - HiLb L^T) ; R += Hb
    for (int e = lane; e < U * U; e += 32) {
      const int i = e / U, j = e - i * U;
      if (j > i) continue;
      double a = 0.0, b2 = 0.0;
      for (int k = 0; k < B; ++k) {
        a += (LS[i * B + k] - HiLb[i * B + k]) * Lm[j * B + k];
        b2 += (LS[j * B + k] - HiLb[j * B + k]) * Lm[i * B + k];
      }
      const double v = 0.5 * (a + b2);
      Hb[i * U + j] = v;
      Hb[j * U + i] = v;
      acc[C::R + i * (i + 1) / 2 + j] += v;
    }
    __syncwarp();
    // Ba += 2 SB Hb + SA Gb^T ; Aa += SB Gb ; AS = A Sn ; BH = B Hb ; BG = B Gb
    for (int e = lane; e < B * U; e += 32) {
      const int i = e / U, m = e - i * U;
      double a = 0.0, h = 0.0;
      for (int k = 0; k < U; ++k) { a += 2.0 * SB[i * U + k] * Hb[k * U + m]; h += c[C::Ba + i * U + k] * Hb[k * U + m]; }
      for (int k = 0; k < B; ++k) a += SA[i * B + k] * Gb[m * B + k];
      acc[C::Ba + e] += a;
      BH[e] = h;
    }
    __syncwarp();                                                            // acc.Aa: the 2 SA Sn pass above is complete
    wmm<B, B, U>(lane, [&](int i, int k) { return SB[i * U + k]; }, [&](int k, int j) { return Gb[k * B + j]; },
                 [&](int i, int j, double v) { acc[C::Aa + i * B + j] += v; });
    wmm<B, B, B>(lane, [&](int i, int k) { return c[C::Aa + i * B + k]; }, [&](int k, int j) { return Sn[k * B + j]; },
                 [&](int i, int j, double v) { AS[i * B + j] = v; });
    wmm<B, B, U>(lane, [&](int i, int k) { return c[C::Ba + i * U + k]; }, [&](int k, int j) { return Gb[k * B + j]; },
                 [&](int i, int j, double v) { BG[i * B + j] = v; });
    __syncwarp();
    // Sn <- A Sn A^T + B Hb B^T + sym(B Gb A^T)
    for (int e = lane; e < B * B; e += 32) {
      const int i = e / B, j = e - i * B;
      if (j > i) continue;
      double a = 0.0, b2 = 0.0;
      for (int k = 0; k < B; ++k) {
        a += AS[i * B + k] * c[C::Aa + j * B + k];
        b2 += BG[i * B + k] * c[C::Aa + j * B + k] + BG[j * B + k] * c[C::Aa + i * B + k];
      }
      for (int k = 0; k < U; ++k) a += BH[i * U + k] * c[C::Ba + j * U + k];
      a += 0.5 * b2;
      Sb[i * B + j] = a;
      Sb[j * B + i] = a;
    }
    __syncwarp();
    double* tmp = Sn; Sn = Sb; Sb = tmp;
  }
  for (int e = lane; e < B * B; e += 32) { const int i = e / B, j = e - i * B; if (j <= i) acc[C::Qf + i * (i + 1) / 2 + j] += Sn[e]; }
  __syncwarp();
  bw_flush_acc<C>(lane, gacc + s, Sc, acc, C::NSEG);
}

}  // namespace lqgk
