// lqgk_big.cuh -- per-sample (FP64) kernels and launch sequence for LARGE systems (joint dim n = x + b > 12, e.g. the
// 24-dim delayed point-mass model of BASELINE config c4: TemporalDelayModel(PointMassBoundedActor, delay=2)).
//
// The small-system kernels keep every matrix of a sample in registers (fully unrolled templates) and its constants in
// shared memory; neither fits at n = 24 (576-entry joint matrices, ~900 derived constants per sample).  This path runs
// the SAME step functions (lqgk_core.h / lqgk_stages.h, compiled with -DLQGK_BIG so their loops stay rolled) with
//   * one thread per parameter sample, BIG_TPB samples per CTA (few lanes per warp: these systems come in thousands,
//     not hundreds of thousands, of samples, and the recursions are latency-bound -- more warps beat fuller warps),
//   * matrices in per-thread local memory (L1-resident, interleaved per lane by the hardware),
//   * derived constants / cotangent accumulators in sample-minor global scratch (coalesced across a warp's samples),
//   * the time-parallel contraction of the covariance adjoint split over blockIdx.y time ranges, accumulating into the
//     per-sample accumulator block with FP64 atomics.
// The per-trial FP32 kernels (k_trial_fwd / k_trial_rev: warp per sample, lane = trial, TMA-staged records) are shared
// with the small-system path; only their ring geometry adapts to the record size.
#pragma once
#include "lqgk_bigw.cuh"
#include "lqgk_run.cuh"

namespace lqgk {

constexpr int BIG_TPB = 8;   // samples (threads) per CTA

template <class DM, bool AFFINE>
__global__ void __launch_bounds__(BIG_TPB) kb_lqr_fwd(const double* cst, size_t Sc, size_t tstride, int npad, int Tn, double eps,
                                                      double* L, int save_S, double* Sric, double* l, double* H, double* scr) {
  const size_t s = (size_t)blockIdx.x * BIG_TPB + threadIdx.x;
  if (s >= (size_t)npad) return;
  lqr_fwd_body<DM, AFFINE>(GCst{cst + s, Sc, tstride}, WView{scr + s, Sc}, Tn, eps, WView{L + s, Sc}, save_S != 0, WView{Sric + s, Sc},
                           WView{l + s, Sc}, WView{H + s, Sc});
}

template <class DM>
__global__ void __launch_bounds__(BIG_TPB) kb_kf_fwd(const double* cst, size_t Sc, size_t tstride, int npad, int Tn, double* K,
                                                     int save_P, double* Pkf, double* scr) {
  const size_t s = (size_t)blockIdx.x * BIG_TPB + threadIdx.x;
  if (s >= (size_t)npad) return;
  kf_fwd_body<DM>(GCst{cst + s, Sc, tstride}, WView{scr + s, Sc}, Tn, WView{K + s, Sc}, save_P != 0, WView{Pkf + s, Sc});
}

// lcc: CovC-layout constants [C::n][Sc] (filled by kb_load_cov_consts); la: accumulators [C::n][Sc] (zeroed by the caller).
template <class DM>
__global__ void __launch_bounds__(BIG_TPB) kb_load_cov_consts(const double* cst, size_t Sc, int npad, double* lcc) {
  const size_t s = (size_t)blockIdx.x * BIG_TPB + threadIdx.x;
  if (s >= (size_t)npad) return;
  load_consts<CovC<DM>>(WView{const_cast<double*>(cst) + s, Sc}, WView{lcc + s, Sc}, CovC<DM>::NSEG);
}
template <class DM>
__global__ void __launch_bounds__(BIG_TPB) kb_flush_cov(size_t Sc, int npad, const double* la, double* acc) {
  const size_t s = (size_t)blockIdx.x * BIG_TPB + threadIdx.x;
  if (s >= (size_t)npad) return;
  flush_acc<CovC<DM>>(WView{acc + s, Sc}, WView{const_cast<double*>(la) + s, Sc}, CovC<DM>::NSEG);
}

template <class DM>
__global__ void __launch_bounds__(BIG_TPB) kb_kf_rev(const double* cst, size_t Sc, int npad, int Tn, const double* Pkf,
                                                     const double* Kbar, double* acc, double* lc, double* la) {
  const size_t s = (size_t)blockIdx.x * BIG_TPB + threadIdx.x;
  if (s >= (size_t)npad) return;
  kf_rev_body<DM>(GCst{cst + s, Sc, 0}, WView{lc + s, Sc}, WView{la + s, Sc}, Tn, WView{const_cast<double*>(Pkf) + s, Sc},
                  WView{const_cast<double*>(Kbar) + s, Sc}, WView{acc + s, Sc});
}
template <class DM>
__global__ void __launch_bounds__(BIG_TPB) kb_lqr_rev(const double* cst, size_t Sc, int npad, int Tn, double eps, const double* L,
                                                      const double* Sric, const double* Lbar, double* acc, double* lc, double* la) {
  const size_t s = (size_t)blockIdx.x * BIG_TPB + threadIdx.x;
  if (s >= (size_t)npad) return;
  lqr_rev_body<DM>(GCst{cst + s, Sc, 0}, WView{lc + s, Sc}, WView{la + s, Sc}, Tn, eps, WView{const_cast<double*>(L) + s, Sc},
                   WView{const_cast<double*>(Sric) + s, Sc}, WView{const_cast<double*>(Lbar) + s, Sc}, WView{acc + s, Sc});
}

// rows of the shared scratch block (two halves: constants copy | accumulators)
template <class DM>
constexpr int big_scratch_rows() {
  int m = CovC<DM>::n;
  m = LqrC<DM>::n_affine > m ? LqrC<DM>::n_affine : m;
  m = KfC<DM>::n > m ? KfC<DM>::n : m;
  m = CovSeqRev<DM>::SC_N > m ? CovSeqRev<DM>::SC_N : m;
  return m;
}

template <class DM, class T>
int run_big(const Call& c) {
  const LqgkDims& d = *c.dims;
  constexpr CLayout cl = DM::CL;
  const bool has_dyn = c.dyn != nullptr;
  const bool tv = spec_time_varying(*c.act, true) || (has_dyn && spec_time_varying(*c.dyn, false));
  if (tv && c.mode == LQGK_MODE_VJP) return LQGK_E_UNSUPPORTED;
  if (!c.ws || ((uintptr_t)c.ws % ALIGN) != 0) return LQGK_E_INVALID;
  const size_t Sc = choose_chunk<DM>(d, c.mode, tv, c.ws_bytes / ALIGN * ALIGN, 0);
  if (Sc == 0) return LQGK_E_WORKSPACE;
  const Plan p = make_plan<DM>(d, c.mode, tv, Sc);
  char* base = (char*)c.ws;
  auto D = [&](size_t off) { return (double*)(base + off); };
  auto F = [&](size_t off) { return (float*)(base + off); };
  cudaStream_t st = c.stream;
  const int Tn = d.T, N = d.N;
  const size_t tstride = tv ? (size_t)cl.total * Sc : 0;
  constexpr int SROWS = big_scratch_rows<DM>();
  double* lc = D(p.scr);
  double* la = lc + (size_t)SROWS * Sc;

  PackArgs<T> pa{};
  pa.act = *c.act;
  if (has_dyn) pa.dyn = *c.dyn;
  pa.sigma0 = c.sigma0 ? *c.sigma0 : LqgkMat{nullptr, 0, 0};
  pa.x = DM::X; pa.b = DM::B; pa.u = DM::U; pa.y = DM::Y; pa.nT = Tn; pa.has_dyn = has_dyn;

  for (size_t s0 = 0; s0 < (size_t)d.S; s0 += Sc) {
    const int n = (int)std::min(Sc, (size_t)d.S - s0);
    const int npad = (int)up(n, 32);
    const int nblk = (npad + BIG_TPB - 1) / BIG_TPB;
    {
      dim3 grid((npad + 127) / 128, tv ? Tn : 1);
      ProfScope ps_(PK_PACK, st);
      k_pack<T><<<grid, 128, 0, st>>>(pa, (int)s0, n, npad, D(p.cst), Sc, tstride, tv ? Tn : 1);
      LQGK_LAUNCH_CHECK();
    }
    auto store = [&](size_t off, int E, void* out) -> int {
      if (!out) return LQGK_OK;
      size_t total = (size_t)n * Tn * E;
      k_store_rows<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(D(off), Sc, n, Tn, E, (T*)out + s0 * Tn * E);
      LQGK_LAUNCH_CHECK();
      return LQGK_OK;
    };
    if (c.mode == LQGK_MODE_GAINS) {
      int rc;
      if (c.L_out) {
        ProfScope ps_(PK_LQR_FWD, st);
        kb_lqr_fwd<DM, true><<<nblk, BIG_TPB, 0, st>>>(D(p.cst), Sc, tstride, npad, Tn, c.eps, D(p.L), 0, nullptr, D(p.l), D(p.H), lc);
        LQGK_LAUNCH_CHECK();
        if ((rc = store(p.L, DM::EL, c.L_out))) return rc;
        if ((rc = store(p.l, DM::U, c.l_out))) return rc;
        if ((rc = store(p.H, DM::U * DM::U, c.H_out))) return rc;
      }
      if (c.K_out) {
        ProfScope ps_(PK_KF_FWD, st);
        kb_kf_fwd<DM><<<nblk, BIG_TPB, 0, st>>>(D(p.cst), Sc, tstride, npad, Tn, D(p.K), 0, nullptr, lc);
        LQGK_LAUNCH_CHECK();
        if ((rc = store(p.K, DM::EK, c.K_out))) return rc;
      }
      continue;
    }
    const bool vjp = c.mode == LQGK_MODE_VJP;
    int rc;
    {
      ProfScope ps_(PK_LQR_FWD, st);
      kb_lqr_fwd<DM, false><<<nblk, BIG_TPB, 0, st>>>(D(p.cst), Sc, tstride, npad, Tn, c.eps, D(p.L), vjp, vjp ? D(p.Sric) : nullptr,
                                                      nullptr, nullptr, lc);
      LQGK_LAUNCH_CHECK();
    }
    {
      ProfScope ps_(PK_KF_FWD, st);
      kb_kf_fwd<DM><<<nblk, BIG_TPB, 0, st>>>(D(p.cst), Sc, tstride, npad, Tn, D(p.K), vjp, vjp ? D(p.Pkf) : nullptr, lc);
      LQGK_LAUNCH_CHECK();
    }
    const int wblk = (npad + BW_WARPS - 1) / BW_WARPS;
    {
      if (tv) return LQGK_E_UNSUPPORTED;   // time-varying specs: gains API only on the large-system path
      size_t smem = BigW<DM>::smem_fwd();
      if ((rc = set_smem<DM>((const void*)kw_cov_fwd<DM>, smem))) return rc;
      ProfScope ps_(PK_COV_FWD, st);
      kb_load_cov_consts<DM><<<nblk, BIG_TPB, 0, st>>>(D(p.cst), Sc, npad, lc);
      LQGK_LAUNCH_CHECK();
      kw_cov_fwd<DM><<<wblk, 32 * BW_WARPS, smem, st>>>(lc, Sc, npad, Tn, D(p.L), D(p.K), vjp, vjp ? D(p.Cs) : nullptr, vjp ? D(p.FU) : nullptr,
                                                       vjp ? D(p.JS) : nullptr, vjp ? D(p.J0) : nullptr, F(p.rec));
      LQGK_LAUNCH_CHECK();
    }
    const int RT = std::min((N + 31) / 32, trial_rt_max<DM>());
    float* hist = vjp ? F(p.hist) : nullptr;
    rc = LQGK_E_UNSUPPORTED;
    static_for<1, trial_rt_max<DM>() + 1>([&](auto RTC) {
      if (RT == decltype(RTC)::value)
        rc = launch_trial_fwd<DM, decltype(RTC)::value>(st, F(p.rec), c.x_tm, (size_t)d.x_sample_stride, (int)s0, n, N, Tn, D(p.ll), hist);
    });
    if (rc) return rc;
    {
      size_t total = (size_t)n * N;
      ProfScope ps_(PK_MISC, st);
      k_store_ll<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(D(p.ll), total, (T*)c.ll_out + s0 * N);
      LQGK_LAUNCH_CHECK();
    }
    if (!vjp) continue;
    {
      size_t total = (size_t)npad * N, valid = (size_t)n * N;
      const T* lb = c.ll_bar ? (const T*)c.ll_bar + s0 * N : nullptr;
      k_load_w<T><<<(unsigned)((valid + 255) / 256), 256, 0, st>>>(lb, valid, F(p.w));
      LQGK_LAUNCH_CHECK();
      if (total > valid) {
        k_load_w<T><<<(unsigned)((total - valid + 255) / 256), 256, 0, st>>>(nullptr, total - valid, F(p.w) + valid);
        LQGK_LAUNCH_CHECK();
      }
    }
    rc = LQGK_E_UNSUPPORTED;
    static_for<1, trial_rt_max<DM>() + 1>([&](auto RTC) {
      if (RT == decltype(RTC)::value)
        rc = launch_trial_rev<DM, decltype(RTC)::value>(st, F(p.rec), c.x_tm, (size_t)d.x_sample_stride, (int)s0, hist, F(p.w), n, N, Tn, F(p.sums));
    });
    if (rc) return rc;
    if (cudaMemsetAsync(D(p.acc), 0, sizeof(double) * cl.total * Sc, st) != cudaSuccess) return LQGK_E_CUDA;
    if (npad > n) {
      if (cudaMemsetAsync(F(p.sums) + (size_t)n * Tn * DM::SUMP, 0, sizeof(float) * (size_t)(npad - n) * Tn * DM::SUMP, st) != cudaSuccess)
        return LQGK_E_CUDA;
    }
    {
      size_t smem = BigW<DM>::smem_seq();
      if ((rc = set_smem<DM>((const void*)kw_cov_seq_rev<DM>, smem))) return rc;
      ProfScope ps_(PK_COV_REV, st);
      kw_cov_seq_rev<DM><<<wblk, 32 * BW_WARPS, smem, st>>>(npad, Tn, N, F(p.w), D(p.FU), D(p.JS), D(p.J0), F(p.sums), D(p.SGB), D(p.SGBI),
                                                           D(p.SFW));
      LQGK_LAUNCH_CHECK();
    }
    {
      size_t smem = BigW<DM>::smem_con();
      if ((rc = set_smem<DM>((const void*)kw_cov_contrib<DM>, smem))) return rc;
      ProfScope ps_(PK_COV_CONTRIB, st);
      if (cudaMemsetAsync(la, 0, sizeof(double) * SROWS * Sc, st) != cudaSuccess) return LQGK_E_CUDA;
      // time ranges: enough (sample x range) warps to put ~12 on every SM when there are few samples
      int chunks = std::max(1, std::min((Tn + 3) / 4, (148 * 12 + npad - 1) / npad));
      kw_cov_contrib<DM><<<dim3(wblk, chunks), 32 * BW_WARPS, smem, st>>>(lc, Sc, npad, Tn, D(p.L), D(p.K), D(p.Cs), D(p.SGB), D(p.SGBI),
                                                                         D(p.SFW), F(p.sums), la, D(p.Lbar), D(p.Kbar));
      LQGK_LAUNCH_CHECK();
      kb_flush_cov<DM><<<nblk, BIG_TPB, 0, st>>>(Sc, npad, la, D(p.acc));
      LQGK_LAUNCH_CHECK();
    }
    {
      ProfScope ps_(PK_KF_REV, st);
      kb_kf_rev<DM><<<nblk, BIG_TPB, 0, st>>>(D(p.cst), Sc, npad, Tn, D(p.Pkf), D(p.Kbar), D(p.acc), lc, la);
      LQGK_LAUNCH_CHECK();
    }
    {
      ProfScope ps_(PK_LQR_REV, st);
      kb_lqr_rev<DM><<<nblk, BIG_TPB, 0, st>>>(D(p.cst), Sc, npad, Tn, c.eps, D(p.L), D(p.Sric), D(p.Lbar), D(p.acc), lc, la);
      LQGK_LAUNCH_CHECK();
    }
    {
      UnpackArgs<T> ua{};
      ua.act = *c.act; ua.dyn = *c.dyn; ua.sigma0 = pa.sigma0;
      if (c.gact) ua.gact = *c.gact;
      if (c.gdyn) ua.gdyn = *c.gdyn;
      if (c.gsig0) ua.gsigma0 = *c.gsig0;
      ua.x = DM::X; ua.b = DM::B; ua.u = DM::U; ua.y = DM::Y;
      ProfScope ps_(PK_UNPACK, st);
      k_unpack<T><<<(n + 63) / 64, 64, 0, st>>>(ua, (int)s0, n, D(p.acc), D(p.cst), Sc);
      LQGK_LAUNCH_CHECK();
    }
  }
  return LQGK_OK;
}

}  // namespace lqgk
