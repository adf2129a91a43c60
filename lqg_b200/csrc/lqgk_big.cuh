// lqgk_big.cuh -- launch sequence for LARGE systems (joint dim n = x + b > 12, e.g. the 24-dim delayed point-mass model
// of BASELINE config c4: TemporalDelayModel(PointMassBoundedActor, delay=2)).
//
// The small-system kernels keep every matrix of a sample in registers (one thread per sample, fully unrolled templates)
// and its constants in shared memory; neither fits at n = 24 (576-entry joint matrices, ~900 derived constants per sample),
// and these systems come in thousands, not hundreds of thousands, of samples.  The large-system path therefore maps ONE
// WARP PER PARAMETER SAMPLE with the matrices in shared memory (lqgk_bigw.cuh: kw_lqr_fwd, kw_kf_fwd, kw_cov_fwd,
// kw_cov_seq_rev, kw_cov_contrib, kw_kf_rev, kw_lqr_rev) and sample-major workspace arrays.  The per-trial FP32 kernels
// (k_trial_fwd / k_trial_rev: warp per sample, lane = trial, TMA-staged records) are shared with the small-system path;
// only their ring geometry adapts to the record size.  The object is compiled with -DLQGK_BIG (rolled FP64 loops in the
// scalar helpers the lanes still call for the u x u / y x y / d x d factorisations).
#pragma once
#include "lqgk_bigw.cuh"
#include "lqgk_run.cuh"

namespace lqgk {

template <class DM, class T>
int run_big(const Call& c) {
  const LqgkDims& d = *c.dims;
  constexpr CLayout cl = DM::CL;
  const bool has_dyn = c.dyn != nullptr;
  const bool tv = spec_time_varying(*c.act, true) || (has_dyn && spec_time_varying(*c.dyn, false));
  if (tv && c.mode == LQGK_MODE_VJP) return LQGK_E_UNSUPPORTED;
  if (c.mode == LQGK_MODE_MOMENTS) return LQGK_E_UNSUPPORTED;   // moments of the large systems stay on the host-side slow path
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

  PackArgs<T> pa{};
  pa.act = *c.act;
  if (has_dyn) pa.dyn = *c.dyn;
  pa.sigma0 = c.sigma0 ? *c.sigma0 : LqgkMat{nullptr, 0, 0};
  pa.x = DM::X; pa.b = DM::B; pa.u = DM::U; pa.y = DM::Y; pa.nT = Tn; pa.has_dyn = has_dyn;

  for (size_t s0 = 0; s0 < (size_t)d.S; s0 += Sc) {
    const int n = (int)std::min(Sc, (size_t)d.S - s0);
    const int npad = (int)up(n, 32);
    const int wblk = (npad + BW_WARPS - 1) / BW_WARPS;
    const unsigned wthr = 32 * BW_WARPS;
    const size_t sm_lqr = BigG<DM>::smem_lqr(), sm_kf = BigG<DM>::smem_kf();
    {
      dim3 grid((npad + 127) / 128, tv ? Tn : 1);
      ProfScope ps_(PK_PACK, st);
      k_pack<T><<<grid, 128, 0, st>>>(pa, (int)s0, n, npad, D(p.cst), Sc, tstride, tv ? Tn : 1);
      LQGK_LAUNCH_CHECK();
    }
    auto store = [&](size_t off, int E, void* out) -> int {   // workspace arrays are sample-major [s][t][E] = the user layout
      if (!out) return LQGK_OK;
      size_t total = (size_t)n * Tn * E;
      k_store_ll<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(D(off), total, (T*)out + s0 * Tn * E);
      LQGK_LAUNCH_CHECK();
      return LQGK_OK;
    };
    if (tv) return LQGK_E_UNSUPPORTED;   // time-varying specs are not implemented on the large-system path
    if (c.mode == LQGK_MODE_GAINS) {
      int rc;
      if (c.L_out) {
        if ((rc = set_smem<DM>((const void*)kw_lqr_fwd<DM, true>, sm_lqr))) return rc;
        ProfScope ps_(PK_LQR_FWD, st);
        kw_lqr_fwd<DM, true><<<wblk, wthr, sm_lqr, st>>>(D(p.cst), Sc, npad, Tn, c.eps, D(p.L), 0, nullptr, D(p.l), D(p.H));
        LQGK_LAUNCH_CHECK();
        if ((rc = store(p.L, DM::EL, c.L_out))) return rc;
        if ((rc = store(p.l, DM::U, c.l_out))) return rc;
        if ((rc = store(p.H, DM::U * DM::U, c.H_out))) return rc;
      }
      if (c.K_out) {
        if ((rc = set_smem<DM>((const void*)kw_kf_fwd<DM>, sm_kf))) return rc;
        ProfScope ps_(PK_KF_FWD, st);
        kw_kf_fwd<DM><<<wblk, wthr, sm_kf, st>>>(D(p.cst), Sc, npad, Tn, D(p.K), 0, nullptr);
        LQGK_LAUNCH_CHECK();
        if ((rc = store(p.K, DM::EK, c.K_out))) return rc;
      }
      continue;
    }
    const bool vjp = c.mode == LQGK_MODE_VJP;
    int rc;
    {
      if ((rc = set_smem<DM>((const void*)kw_lqr_fwd<DM, false>, sm_lqr))) return rc;
      ProfScope ps_(PK_LQR_FWD, st);
      kw_lqr_fwd<DM, false><<<wblk, wthr, sm_lqr, st>>>(D(p.cst), Sc, npad, Tn, c.eps, D(p.L), vjp, vjp ? D(p.Sric) : nullptr, nullptr, nullptr);
      LQGK_LAUNCH_CHECK();
    }
    {
      if ((rc = set_smem<DM>((const void*)kw_kf_fwd<DM>, sm_kf))) return rc;
      ProfScope ps_(PK_KF_FWD, st);
      kw_kf_fwd<DM><<<wblk, wthr, sm_kf, st>>>(D(p.cst), Sc, npad, Tn, D(p.K), vjp, vjp ? D(p.Pkf) : nullptr);
      LQGK_LAUNCH_CHECK();
    }
    {
      size_t smem = BigW<DM>::smem_fwd();
      if ((rc = set_smem<DM>((const void*)kw_cov_fwd<DM>, smem))) return rc;
      ProfScope ps_(PK_COV_FWD, st);
      kw_cov_fwd<DM><<<wblk, wthr, smem, st>>>(D(p.cst), Sc, npad, Tn, D(p.L), D(p.K), vjp, vjp ? D(p.Cs) : nullptr, vjp ? F(p.FU) : nullptr,
                                               vjp ? F(p.JS) : nullptr, vjp ? D(p.J0) : nullptr, F(p.rec));
      LQGK_LAUNCH_CHECK();
    }
    const int RT = std::min((N + 31) / 32, trial_rt_max<DM>());
    float* hist = vjp ? F(p.hist) : nullptr;
    size_t xcs = 0;
    if ((rc = launch_repack_obs(st, d, 32 * RT, c.x_tm, (int)s0, n, F(p.xc), &xcs))) return rc;
    rc = LQGK_E_UNSUPPORTED;
    static_for<1, trial_rt_max<DM>() + 1>([&](auto RTC) {
      if (RT == decltype(RTC)::value)
        rc = launch_trial_fwd<DM, decltype(RTC)::value>(st, F(p.rec), F(p.xc), xcs, n, N, Tn, D(p.ll), hist);
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
        rc = launch_trial_rev<DM, decltype(RTC)::value>(st, F(p.rec), F(p.xc), xcs, hist, F(p.w), n, N, Tn, F(p.sums), F(p.carcb));
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
      kw_cov_seq_rev<DM><<<wblk, 32 * BW_WARPS, smem, st>>>(npad, Tn, N, F(p.w), F(p.FU), F(p.JS), D(p.J0), F(p.sums), F(p.SGB), D(p.SGBI),
                                                           F(p.SFW));
      LQGK_LAUNCH_CHECK();
    }
    {
      size_t smem = BigW<DM>::smem_con();
      if ((rc = set_smem<DM>((const void*)kw_cov_contrib<DM>, smem))) return rc;
      ProfScope ps_(PK_COV_CONTRIB, st);
      // time ranges: enough (sample x range) warps to put ~12 on every SM when there are few samples
      int chunks = std::max(1, std::min((Tn + 3) / 4, (sm_count() * 12 + npad - 1) / npad));
      kw_cov_contrib<DM><<<dim3(wblk, chunks), wthr, smem, st>>>(D(p.cst), Sc, npad, Tn, D(p.L), D(p.K), D(p.Cs), F(p.SGB), D(p.SGBI),
                                                                F(p.SFW), F(p.sums), D(p.acc), D(p.Lbar), D(p.Kbar), nullptr);
      LQGK_LAUNCH_CHECK();
    }
    {
      if ((rc = set_smem<DM>((const void*)kw_kf_rev<DM>, sm_kf))) return rc;
      ProfScope ps_(PK_KF_REV, st);
      kw_kf_rev<DM><<<wblk, wthr, sm_kf, st>>>(D(p.cst), Sc, npad, Tn, D(p.Pkf), D(p.Kbar), D(p.acc));
      LQGK_LAUNCH_CHECK();
    }
    {
      if ((rc = set_smem<DM>((const void*)kw_lqr_rev<DM>, sm_lqr))) return rc;
      ProfScope ps_(PK_LQR_REV, st);
      kw_lqr_rev<DM><<<wblk, wthr, sm_lqr, st>>>(D(p.cst), Sc, npad, Tn, c.eps, D(p.L), D(p.Sric), D(p.Lbar), D(p.acc));
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
