// lqgk_stages.h -- whole-sweep bodies of the per-sample (FP64) stages, one call per parameter sample.
//
// On the GPU each body is executed by one thread (lane = sample, 32 samples per warp-CTA) with the derived
// constants and cotangent accumulators in shared memory (view stride 32); tests/emul runs the same bodies on
// the CPU with stride-1 local arrays.  Workspace arrays are sample-minor: row (t * E + e) of W lives at
// W.p[(t * E + e) * W.stride], so a warp's accesses are coalesced.
#pragma once
#include "lqgk_core.h"

// Largest constant block (doubles) that is copied into registers (RegView below); larger systems read shared memory.
#ifndef LQGK_REG_CONSTS_MAX
#define LQGK_REG_CONSTS_MAX 48
#endif

namespace lqgk {

template <class KC, class G, class Lc>
LQGK_HD void load_consts(const G& g, Lc&& l, int nseg) {
  for (int i = 0; i < nseg; ++i) {
    int go, lo, len;
    KC::seg(i, go, lo, len);
    for (int e = 0; e < len; ++e) l(lo + e) = g(go + e);
  }
}
template <class KC, class G, class La>
LQGK_HD void flush_acc(G&& g, const La& l, int nseg) {
  for (int i = 0; i < nseg; ++i) {
    int go, lo, len;
    KC::seg(i, go, lo, len);
    for (int e = 0; e < len; ++e) g(go + e) += l(lo + e);
  }
}
template <int M, class W>
LQGK_HD void store_sym(W&& w, size_t off, const double* C) {
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) w(off + i * (i + 1) / 2 + j) = C[i * M + j];
}
template <int M, class W>
LQGK_HD void load_sym_ws(const W& w, size_t off, double* C) {
  LQGK_UNROLL64 for (int i = 0; i < M; ++i) LQGK_UNROLL64 for (int j = 0; j <= i; ++j) {
    double a = w(off + i * (i + 1) / 2 + j);
    C[i * M + j] = a;
    C[j * M + i] = a;
  }
}

// Register copy of a kernel-local constant block: with the fully unrolled step functions every index is a compile-time
// constant, so the block lives in registers and the per-step shared-memory reads of the constants (and their latency, which
// bounds the sequential kernels) disappear.  Used for time-invariant specs; time-varying ones keep reading the reloaded block.
template <int NN>
struct RegView {
  double v[NN];
  LQGK_HD double operator()(int e) const { return v[e]; }
};

// gcst: this sample's global constant block (stride = chunk size); tstride: elements between consecutive
// time steps of that block (0 = time-invariant).  lc: local (shared-memory) copy used by the step functions.
struct GCst {
  const double* p;
  size_t stride;
  size_t tstride;
  LQGK_HD WView at(int t) const { return WView{const_cast<double*>(p) + (size_t)t * tstride, stride}; }
};

// ---------------------------------------------------------------------------------------------- LQR fwd
template <class DM, bool AFFINE>
LQGK_HD void lqr_fwd_body(const GCst& g, WView lc, int T, double eps, WView Lw, bool save_S, WView Sw, WView lw, WView Hw) {
  constexpr int B = DM::B, U = DM::U;
  using C = LqrC<DM>;
  const int nseg = AFFINE ? C::NSEG_AFF : C::NSEG;
  load_consts<C>(g.at(g.tstride ? T - 1 : 0), lc, nseg);
  double S[B * B], s[B];
  load_sym<B>(lc, C::Qf, S);
  if (AFFINE) { LQGK_UNROLL64 for (int i = 0; i < B; ++i) s[i] = lc(C::qf + i); }
  auto sweep = [&](const auto& cv, bool reload) {
    for (int t = T - 1; t >= 0; --t) {
      if (reload && t != T - 1) load_consts<C>(g.at(t), lc, nseg);
      if (save_S) store_sym<B>(Sw, (size_t)t * DM::ES, S);
      double L[U * B], l[U], Ht[U * U], shift;
      LqrFwd<DM, AFFINE>::step(cv, eps, S, s, L, l, Ht, shift);
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) Lw((size_t)t * DM::EL + i) = L[i];
      if (AFFINE) {
        LQGK_UNROLL64 for (int i = 0; i < U; ++i) lw((size_t)t * U + i) = l[i];
        LQGK_UNROLL64 for (int i = 0; i < U * U; ++i) Hw((size_t)t * U * U + i) = Ht[i];
      }
    }
  };
  constexpr int NCST = AFFINE ? C::n_affine : C::n;
  bool done = false;
  if constexpr (NCST <= LQGK_REG_CONSTS_MAX) {
    if (g.tstride == 0) {
      RegView<NCST> rc;
      LQGK_UNROLL64 for (int e = 0; e < NCST; ++e) rc.v[e] = lc(e);
      sweep(rc, false);
      done = true;
    }
  }
  if (!done) sweep(lc, g.tstride != 0);
}

// ---------------------------------------------------------------------------------------------- KF fwd
// [t0, t1): time range of this call (t1 < 0: up to T).  A sweep that starts at t0 > 0 continues from the P_{t0} an earlier
// call saved (save_P must have been on): that is how the pipelined launch sequence cuts the sweep into segments.
template <class DM>
LQGK_HD void kf_fwd_body(const GCst& g, WView lc, int T, WView Kw, bool save_P, WView Pw, int t0 = 0, int t1 = -1) {
  constexpr int B = DM::B, Y = DM::Y;
  using C = KfC<DM>;
  if (t1 < 0) t1 = T;
  load_consts<C>(g.at(g.tstride ? t0 : 0), lc, C::NSEG);
  double P[B * B];
  if (t0 == 0) load_sym<B>(lc, C::Sig0, P);
  else load_sym_ws<B>(Pw, (size_t)t0 * DM::EP, P);
  auto sweep = [&](const auto& cv, bool reload) {
    for (int t = t0; t < t1; ++t) {
      if (reload && t != t0) load_consts<C>(g.at(t), lc, C::NSEG);
      if (save_P) store_sym<B>(Pw, (size_t)t * DM::EP, P);
      double K[B * Y];
      KfFwd<DM>::step(cv, P, K);
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) Kw((size_t)t * DM::EK + i) = K[i];
    }
    if (save_P && t1 < T) store_sym<B>(Pw, (size_t)t1 * DM::EP, P);   // where the next time segment continues from
  };
  bool done = false;
  if constexpr (C::n <= LQGK_REG_CONSTS_MAX) {
    if (g.tstride == 0) {
      RegView<C::n> rc;
      LQGK_UNROLL64 for (int e = 0; e < C::n; ++e) rc.v[e] = lc(e);
      sweep(rc, false);
      done = true;
    }
  }
  if (!done) sweep(lc, g.tstride != 0);
}

// ---------------------------------------------------------------------------------------------- COV fwd
// Sink: put(idx, v) stores one float of the current step's record; commit(t) publishes the record.
// When save_adj: C_t -> Cw, Fu_t -> FUw, (J_t, S'^-1_t) -> JSw, J_0 (initial conditioning gain) -> J0w.
struct NoSigSink {
  LQGK_HD void operator()(int, int, double) const {}
};
// sig(t, e, v) (optional): full predictive joint covariance of step t (moments entry point).
template <class DM, class Sink, class SigSink = NoSigSink>
LQGK_HD void cov_fwd_body(const GCst& g, WView lc, int T, WView Lw, WView Kw, bool save_adj, WView Cw, WView FUw, WView JSw,
                          WView J0w, Sink&& sink, SigSink&& sig = SigSink{}) {
  constexpr int B = DM::B, U = DM::U, Y = DM::Y, R = DM::R, D = DM::D;
  using C = CovC<DM>;
  using SR = CovSeqRev<DM>;
  load_consts<C>(g.at(0), lc, C::NSEG);
  double Cm[R * R], L[U * B], K[B * Y];
  LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) K[i] = Kw(i);
  {
    double J0[R * D];
    CovFwd<DM>::init(lc, K, Cm, J0);
    if (save_adj) { LQGK_UNROLL64 for (int i = 0; i < R * D; ++i) J0w(i) = J0[i]; }
  }
  for (int t = 0; t < T; ++t) {
    if (g.tstride && t != 0) load_consts<C>(g.at(t), lc, C::NSEG);
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) L[i] = Lw((size_t)t * DM::EL + i);
    LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) K[i] = Kw((size_t)t * DM::EK + i);
    if (save_adj) store_sym<R>(Cw, (size_t)t * DM::EC, Cm);
    CovFwd<DM>::step(lc, L, K, Cm, [&](int idx, float v) { sink.put(idx, v); },
                     [&](int which, int e, double v) {
                       if (save_adj) {
                         if (which == 0) FUw((size_t)t * SR::NSF + e) = v;
                         else JSw((size_t)t * SR::NJS + e) = v;
                       }
                     }, [&](int e, double v) { sig(t, e, v); });
    sink.commit(t);
  }
}

// ---------------------------------------------------------------------------------------------- COV rev (sequential)
// Source: fetch(t) makes the float sums of step t available (only SUM_J.. and SUM_W.. are read); get(idx) reads one.
template <class DM, class Source>
LQGK_HD void cov_seq_rev_body(int T, double sw, WView FUw, WView JSw, WView J0w, Source&& src, WView sc, WView SGBw, WView SGBIw,
                              WView SFw) {
  constexpr int R = DM::R;
  using SR = CovSeqRev<DM>;
  double Cb[R * R];
  LQGK_UNROLL64 for (int i = 0; i < R * R; ++i) Cb[i] = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    src.fetch(t);
    SR::step([&](int e) { return FUw((size_t)t * SR::NSF + e); }, [&](int e) { return JSw((size_t)t * SR::NJS + e); },
             [&](int idx) { return src.get(idx); }, sw, sc, Cb,
             [&](int e, double v) { SGBw((size_t)t * SR::NSGB + e) = v; }, [&](int e, double v) { SFw((size_t)t * SR::NSF + e) = v; });
  }
  SR::init([&](int e) { return J0w(e); }, Cb, [&](int e, double v) { SGBIw(e) = v; });
}

// ---------------------------------------------------------------------------------------------- COV rev (time-parallel)
// lc: local copy of the CovC constants (already loaded).  Processes steps [t0, t1).
// ct(t, e, v): sink of contribution e (CovC layout) of step t.
template <class DM, int PASS, class Source, class CT>
LQGK_HD void cov_contrib_body(WView lc, int t0, int t1, WView Lw, WView Kw, WView Cw, WView SGBw, WView SGBIw, WView SFw,
                              Source&& src, CT&& ct, WView Lbw, WView Kbw) {
  constexpr int B = DM::B, U = DM::U, Y = DM::Y, R = DM::R;
  using SR = CovSeqRev<DM>;
  using CC = CovContrib<DM>;
  for (int t = t0; t < t1; ++t) {
    double Cm[R * R], L[U * B], K[B * Y];
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) L[i] = Lw((size_t)t * DM::EL + i);
    LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) K[i] = Kw((size_t)t * DM::EK + i);
    load_sym_ws<R>(Cw, (size_t)t * DM::EC, Cm);
    src.fetch(t);
    auto sf = [&](int e) { return SFw((size_t)t * SR::NSF + e); };
    auto get = [&](int idx) { return src.get(idx); };
    auto out = [&](int e, double v) { ct(t, e, v); };
    if (PASS == 0) {
      double Lb[U * B], Kb[B * Y];
      CC::pass0(lc, [&](int e) { return SGBw((size_t)t * SR::NSGB + e); }, t == 0, [&](int e) { return SGBIw(e); }, sf, get, Cm,
                L, K, out, Lb, Kb);
      LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) Lbw((size_t)t * DM::EL + i) = Lb[i];
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) Kbw((size_t)t * DM::EK + i) = Kb[i];
    } else {
      double Kb[B * Y];
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) Kb[i] = Kbw((size_t)t * DM::EK + i);
      CC::pass1(lc, sf, get, Cm, L, K, out, Kb);
      LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) Kbw((size_t)t * DM::EK + i) = Kb[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------- KF rev
template <class DM>
LQGK_HD void kf_rev_body(const GCst& g, WView lc, WView la, int T, WView Pw, WView Kbw, WView gacc) {
  constexpr int B = DM::B, Y = DM::Y;
  using C = KfC<DM>;
  load_consts<C>(g.at(0), lc, C::NSEG);
  for (int e = 0; e < C::n; ++e) la(e) = 0.0;
  auto acc = [&](int e) -> double& { return la(e); };
  double Pnb[B * B];
  LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Pnb[i] = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    double P[B * B], Kb[B * Y];
    load_sym_ws<B>(Pw, (size_t)t * DM::EP, P);
    LQGK_UNROLL64 for (int i = 0; i < B * Y; ++i) Kb[i] = Kbw((size_t)t * DM::EK + i);
    KfRev<DM>::step(lc, acc, P, Kb, Pnb);
  }
  KfRev<DM>::finish(acc, Pnb);
  flush_acc<C>(gacc, la, C::NSEG);
}

// ---------------------------------------------------------------------------------------------- LQR rev
template <class DM>
LQGK_HD void lqr_rev_body(const GCst& g, WView lc, WView la, int T, double eps, WView Lw, WView Sw, WView Lbw, WView gacc) {
  constexpr int B = DM::B, U = DM::U;
  using C = LqrC<DM>;
  load_consts<C>(g.at(0), lc, C::NSEG);
  for (int e = 0; e < C::n; ++e) la(e) = 0.0;
  auto acc = [&](int e) -> double& { return la(e); };
  double Sn[B * B];
  LQGK_UNROLL64 for (int i = 0; i < B * B; ++i) Sn[i] = 0.0;
  for (int t = 0; t < T; ++t) {
    double S[B * B], L[U * B], Lb[U * B];
    load_sym_ws<B>(Sw, (size_t)t * DM::ES, S);
    LQGK_UNROLL64 for (int i = 0; i < U * B; ++i) { L[i] = Lw((size_t)t * DM::EL + i); Lb[i] = Lbw((size_t)t * DM::EL + i); }
    // eigen-shift recomputed from S_{t+1} (lqr.py:27-28); zero in every well-posed model
    double shift;
    {
      double Bm[B * U], SB[B * U], H[U * U];
      load_mat<B, U>(lc, C::Ba, Bm);
      mm<B, B, U>(S, Bm, SB);
      load_sym<U>(lc, C::R, H);
      mm_tn_sym<U, B, true>(Bm, SB, H);
      shift = eps - lambda_min<U>(H);
      shift = shift > 0.0 ? shift : 0.0;
    }
    LqrRev<DM>::step(lc, acc, S, L, Lb, shift, Sn);
  }
  LqrRev<DM>::finish(acc, Sn);
  flush_acc<C>(gacc, la, C::NSEG);
}

}  // namespace lqgk
