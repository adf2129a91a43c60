// lqgk_run.cuh -- workspace planning, chunking over parameter samples and the kernel launch sequence for one
// dimension tuple.  Instantiated once per tuple by lqgk_inst.cu (compiled in parallel), called from lqgk_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "../../include/lqgk.h"
#include "lqgk_kernels.cuh"

namespace lqgk {

extern thread_local int g_launches;

// Optional per-kernel timing (bench.py): CUDA events recorded on the launching stream around every launch.
enum ProfKind { PK_PACK = 0, PK_LQR_FWD, PK_KF_FWD, PK_COV_FWD, PK_TRIAL_FWD, PK_MISC, PK_TRIAL_REV, PK_COV_REV, PK_KF_REV,
                PK_LQR_REV, PK_UNPACK, PK_COV_CONTRIB, PK_REDUCE, PK_COUNT };
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs (start, stop)
  std::vector<int> kinds;
  size_t used = 0;
};
extern thread_local Profiler g_prof;
struct ProfScope {
  cudaStream_t st;
  bool active;
  ProfScope(int kind, cudaStream_t s) : st(s), active(g_prof.on) {
    if (!active) return;
    if (g_prof.used + 2 > g_prof.ev.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      g_prof.ev.push_back(a);
      g_prof.ev.push_back(b);
    }
    g_prof.kinds.resize(g_prof.ev.size() / 2);
    g_prof.kinds[g_prof.used / 2] = kind;
    cudaEventRecord(g_prof.ev[g_prof.used], st);
  }
  ~ProfScope() {
    if (!active) return;
    cudaEventRecord(g_prof.ev[g_prof.used + 1], st);
    g_prof.used += 2;
  }
};

#define LQGK_LAUNCH_CHECK()                              \
  do {                                                   \
    ++g_launches;                                        \
    if (cudaPeekAtLastError() != cudaSuccess) return LQGK_E_CUDA; \
  } while (0)

constexpr size_t ALIGN = 256;
inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline bool spec_time_varying(const LqgkSpec& s, bool actor) {
  const LqgkMat* m[] = {&s.A, &s.B, &s.F, &s.V, &s.W, &s.Q, &s.R, &s.q, &s.r, &s.P};
  int n = actor ? 10 : 5;
  for (int i = 0; i < n; ++i)
    if (m[i]->ptr && m[i]->time_stride != 0) return true;
  return false;
}

// Workspace plan for one chunk of Sc (multiple of 32) samples.
struct Plan {
  size_t Sc = 0, bytes = 0;
  size_t cst = 0, acc = 0, L = 0, K = 0, l = 0, H = 0, Sric = 0, Pkf = 0, Cs = 0, Lbar = 0, Kbar = 0, rec = 0, ll = 0,
         sums = 0, hist = 0, w = 0, FU = 0, JS = 0, J0 = 0, SGB = 0, SGBI = 0, SFW = 0, CT = 0;
};

template <class DM>
Plan make_plan(const LqgkDims& d, int mode, bool tv, size_t Sc) {
  Plan p;
  p.Sc = Sc;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = up(off + bytes, ALIGN); return o; };
  const size_t T = d.T, N = d.N;
  constexpr CLayout cl = DM::CL;
  p.cst = take(sizeof(double) * cl.total * Sc * (tv ? T : 1));
  p.L = take(sizeof(double) * T * DM::EL * Sc);
  p.K = take(sizeof(double) * T * DM::EK * Sc);
  if (mode == LQGK_MODE_GAINS) {
    p.l = take(sizeof(double) * T * DM::U * Sc);
    p.H = take(sizeof(double) * T * DM::U * DM::U * Sc);
  } else {
    p.rec = take(sizeof(float) * Sc * T * DM::REC);
    p.ll = take(sizeof(double) * Sc * N);
  }
  if (mode == LQGK_MODE_VJP) {
    p.acc = take(sizeof(double) * cl.total * Sc);
    p.Sric = take(sizeof(double) * T * DM::ES * Sc);
    p.Pkf = take(sizeof(double) * T * DM::EP * Sc);
    p.Cs = take(sizeof(double) * T * DM::EC * Sc);
    p.Lbar = take(sizeof(double) * T * DM::EL * Sc);
    p.Kbar = take(sizeof(double) * T * DM::EK * Sc);
    p.sums = take(sizeof(float) * Sc * T * DM::SUMP);
    p.hist = take(sizeof(float) * Sc * T * DM::R * N);
    p.w = take(sizeof(float) * Sc * N);
    using SR = CovSeqRev<DM>;
    p.FU = take(sizeof(double) * T * SR::NSF * Sc);
    p.JS = take(sizeof(double) * T * SR::NJS * Sc);
    p.J0 = take(sizeof(double) * DM::R * DM::D * Sc);
    p.SGB = take(sizeof(double) * T * SR::NSGB * Sc);
    p.SGBI = take(sizeof(double) * SR::NSGB * Sc);
    p.SFW = take(sizeof(double) * T * SR::NSF * Sc);
    p.CT = take(sizeof(double) * T * CovC<DM>::n * Sc);
  }
  p.bytes = off;
  return p;
}

template <class DM>
size_t choose_chunk(const LqgkDims& d, int mode, bool tv, size_t ws_bytes, int32_t max_chunk) {
  size_t Spad = up((size_t)d.S, 32);
  size_t cap = Spad;
  if (max_chunk > 0) cap = std::min(cap, up((size_t)max_chunk, 32));
  if (ws_bytes == (size_t)-1) return cap;
  // plan size is (almost) linear in Sc: find the largest multiple of 32 that fits
  size_t lo = 0, hi = cap / 32;
  while (lo < hi) {
    size_t mid = (lo + hi + 1) / 2;
    if (make_plan<DM>(d, mode, tv, mid * 32).bytes <= ws_bytes) lo = mid; else hi = mid - 1;
  }
  return lo * 32;
}

template <class DM>
int set_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 227 * 1024) return LQGK_E_UNSUPPORTED;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return LQGK_E_CUDA;
  }
  return LQGK_OK;
}

template <class DM, int RT>
int launch_trial_fwd(cudaStream_t st, const float* rec, const float* x_tm, int n, int N, int T, double* ll, float* hist) {
  size_t smem = trial_smem_bytes<DM, RT, false>();
  int rc = set_smem<DM>((const void*)k_trial_fwd<DM, RT>, smem);
  if (rc) return rc;
  ProfScope ps_(PK_TRIAL_FWD, st);
  k_trial_fwd<DM, RT><<<(n + TRIAL_WARPS - 1) / TRIAL_WARPS, 32 * TRIAL_WARPS, smem, st>>>(rec, x_tm, n, N, T, ll, hist);
  LQGK_LAUNCH_CHECK();
  return LQGK_OK;
}
template <class DM, int RT>
int launch_trial_rev(cudaStream_t st, const float* rec, const float* x_tm, const float* hist, const float* w, int n, int N,
                     int T, float* sums) {
  size_t smem = trial_smem_bytes<DM, RT, true>();
  int rc = set_smem<DM>((const void*)k_trial_rev<DM, RT>, smem);
  if (rc) return rc;
  ProfScope ps_(PK_TRIAL_REV, st);
  k_trial_rev<DM, RT><<<(n + TRIAL_WARPS - 1) / TRIAL_WARPS, 32 * TRIAL_WARPS, smem, st>>>(rec, x_tm, hist, w, n, N, T, sums);
  LQGK_LAUNCH_CHECK();
  return LQGK_OK;
}

struct Call {
  const LqgkDims* dims;
  const LqgkSpec* act;
  const LqgkSpec* dyn;
  const LqgkMat* sigma0;
  const float* x_tm;
  const void* ll_bar;
  void* ll_out;
  const LqgkSpecGrad* gact;
  const LqgkSpecGrad* gdyn;
  const LqgkMatGrad* gsig0;
  void *L_out, *l_out, *H_out, *K_out;
  double eps;
  int mode;
  void* ws;
  size_t ws_bytes;
  cudaStream_t stream;
};

template <class DM, class T>
int run(const Call& c) {
  const LqgkDims& d = *c.dims;
  constexpr CLayout cl = DM::CL;
  const bool has_dyn = c.dyn != nullptr;
  const bool tv = spec_time_varying(*c.act, true) || (has_dyn && spec_time_varying(*c.dyn, false));
  if (tv && c.mode == LQGK_MODE_VJP) return LQGK_E_UNSUPPORTED;
  if (!c.ws || ((uintptr_t)c.ws % ALIGN) != 0) return LQGK_E_INVALID;
  const size_t Sc = choose_chunk<DM>(d, c.mode, tv, c.ws_bytes, 0);
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
    const int nblk = npad / 32;
    {
      dim3 grid((npad + 127) / 128, tv ? Tn : 1);
      ProfScope ps_(PK_PACK, st);
      k_pack<T><<<grid, 128, 0, st>>>(pa, (int)s0, n, npad, D(p.cst), Sc, tstride, tv ? Tn : 1);
      LQGK_LAUNCH_CHECK();
    }
    if (c.mode == LQGK_MODE_GAINS) {
      if (c.L_out) {
        size_t smem = sizeof(double) * 32 * LqrC<DM>::n_affine;
        ProfScope ps_(PK_LQR_FWD, st);
        k_lqr_fwd<DM, true><<<nblk, 32, smem, st>>>(D(p.cst), Sc, tstride, Tn, c.eps, D(p.L), 0, nullptr, D(p.l), D(p.H));
        LQGK_LAUNCH_CHECK();
        auto store = [&](size_t off, int E, void* out) -> int {
          if (!out) return LQGK_OK;
          size_t total = (size_t)n * Tn * E;
          k_store_rows<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(D(off), Sc, n, Tn, E, (T*)out + s0 * Tn * E);
          LQGK_LAUNCH_CHECK();
          return LQGK_OK;
        };
        int rc;
        if ((rc = store(p.L, DM::EL, c.L_out))) return rc;
        if ((rc = store(p.l, DM::U, c.l_out))) return rc;
        if ((rc = store(p.H, DM::U * DM::U, c.H_out))) return rc;
      }
      if (c.K_out) {
        size_t smem = sizeof(double) * 32 * KfC<DM>::n;
        ProfScope ps_(PK_KF_FWD, st);
        k_kf_fwd<DM><<<nblk, 32, smem, st>>>(D(p.cst), Sc, tstride, Tn, D(p.K), 0, nullptr);
        LQGK_LAUNCH_CHECK();
        size_t total = (size_t)n * Tn * DM::EK;
        k_store_rows<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(D(p.K), Sc, n, Tn, DM::EK, (T*)c.K_out + s0 * Tn * DM::EK);
        LQGK_LAUNCH_CHECK();
      }
      continue;
    }
    const bool vjp = c.mode == LQGK_MODE_VJP;
    int rc;
    {
      size_t smem = sizeof(double) * 32 * LqrC<DM>::n;
      ProfScope ps_(PK_LQR_FWD, st);
      k_lqr_fwd<DM, false><<<nblk, 32, smem, st>>>(D(p.cst), Sc, tstride, Tn, c.eps, D(p.L), vjp, vjp ? D(p.Sric) : nullptr, nullptr, nullptr);
      LQGK_LAUNCH_CHECK();
    }
    {
      size_t smem = sizeof(double) * 32 * KfC<DM>::n;
      ProfScope ps_(PK_KF_FWD, st);
      k_kf_fwd<DM><<<nblk, 32, smem, st>>>(D(p.cst), Sc, tstride, Tn, D(p.K), vjp, vjp ? D(p.Pkf) : nullptr);
      LQGK_LAUNCH_CHECK();
    }
    {
      size_t smem = smem_cov_fwd<DM>();
      if ((rc = set_smem<DM>((const void*)k_cov_fwd<DM>, smem))) return rc;
      ProfScope ps_(PK_COV_FWD, st);
      k_cov_fwd<DM><<<nblk, 32, smem, st>>>(D(p.cst), Sc, tstride, Tn, D(p.L), D(p.K), vjp, vjp ? D(p.Cs) : nullptr, vjp ? D(p.FU) : nullptr,
                                            vjp ? D(p.JS) : nullptr, vjp ? D(p.J0) : nullptr, F(p.rec));
      LQGK_LAUNCH_CHECK();
    }
    const int RT = std::min((N + 31) / 32, trial_rt_max<DM>());
    float* hist = vjp ? F(p.hist) : nullptr;
    rc = LQGK_E_UNSUPPORTED;
    static_for<1, trial_rt_max<DM>() + 1>([&](auto RTC) {
      if (RT == decltype(RTC)::value) rc = launch_trial_fwd<DM, decltype(RTC)::value>(st, F(p.rec), c.x_tm, n, N, Tn, D(p.ll), hist);
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
      size_t total = (size_t)npad * N;   // padded samples get weight of... the same trials; harmless (never unpacked)
      const T* lb = c.ll_bar ? (const T*)c.ll_bar + s0 * N : nullptr;
      size_t valid = (size_t)n * N;
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
        rc = launch_trial_rev<DM, decltype(RTC)::value>(st, F(p.rec), c.x_tm, hist, F(p.w), n, N, Tn, F(p.sums));
    });
    if (rc) return rc;
    if (cudaMemsetAsync(D(p.acc), 0, sizeof(double) * cl.total * Sc, st) != cudaSuccess) return LQGK_E_CUDA;
    if (npad > n) {   // sums of padded samples are never written by k_trial_rev: zero them so k_cov_rev reads finite data
      if (cudaMemsetAsync(F(p.sums) + (size_t)n * Tn * DM::SUMP, 0, sizeof(float) * (size_t)(npad - n) * Tn * DM::SUMP, st) != cudaSuccess)
        return LQGK_E_CUDA;
    }
    {
      size_t smem = smem_cov_seq_rev<DM>();
      if ((rc = set_smem<DM>((const void*)k_cov_seq_rev<DM>, smem))) return rc;
      ProfScope ps_(PK_COV_REV, st);
      k_cov_seq_rev<DM><<<nblk, 32, smem, st>>>(Sc, Tn, N, F(p.w), D(p.FU), D(p.JS), D(p.J0), F(p.sums), D(p.SGB), D(p.SGBI), D(p.SFW));
      LQGK_LAUNCH_CHECK();
    }
    {
      size_t smem0 = smem_cov_contrib<DM, 0>(), smem1 = smem_cov_contrib<DM, 1>();
      if ((rc = set_smem<DM>((const void*)k_cov_contrib<DM, 0>, smem0))) return rc;
      if ((rc = set_smem<DM>((const void*)k_cov_contrib<DM, 1>, smem1))) return rc;
      // (sample-group x time-range) warps: enough to occupy every SM a few times over
      int chunks = std::max(1, std::min((Tn + 7) / 8, (148 * 4 + nblk - 1) / nblk));
      dim3 grid(nblk, chunks);
      {
        ProfScope ps_(PK_COV_CONTRIB, st);
        k_cov_contrib<DM, 0><<<grid, 32, smem0, st>>>(D(p.cst), Sc, Tn, D(p.L), D(p.K), D(p.Cs), D(p.SGB), D(p.SGBI), D(p.SFW), F(p.sums),
                                                     D(p.CT), D(p.Lbar), D(p.Kbar));
        LQGK_LAUNCH_CHECK();
      }
      {
        ProfScope ps_(PK_COV_CONTRIB, st);
        k_cov_contrib<DM, 1><<<grid, 32, smem1, st>>>(D(p.cst), Sc, Tn, D(p.L), D(p.K), D(p.Cs), D(p.SGB), D(p.SGBI), D(p.SFW), F(p.sums),
                                                     D(p.CT), D(p.Lbar), D(p.Kbar));
        LQGK_LAUNCH_CHECK();
      }
      {
        ProfScope ps_(PK_REDUCE, st);
        dim3 rgrid((unsigned)((Sc + 127) / 128), CovC<DM>::n);
        k_reduce_time<CovC<DM>><<<rgrid, 128, 0, st>>>(D(p.CT), Sc, Tn, D(p.acc));
        LQGK_LAUNCH_CHECK();
      }
    }
    {
      size_t smem = smem_kf_rev<DM>();
      if ((rc = set_smem<DM>((const void*)k_kf_rev<DM>, smem))) return rc;
      ProfScope ps_(PK_KF_REV, st);
      k_kf_rev<DM><<<nblk, 32, smem, st>>>(D(p.cst), Sc, Tn, D(p.Pkf), D(p.Kbar), D(p.acc));
      LQGK_LAUNCH_CHECK();
    }
    {
      size_t smem = smem_lqr_rev<DM>();
      if ((rc = set_smem<DM>((const void*)k_lqr_rev<DM>, smem))) return rc;
      ProfScope ps_(PK_LQR_REV, st);
      k_lqr_rev<DM><<<nblk, 32, smem, st>>>(D(p.cst), Sc, Tn, c.eps, D(p.L), D(p.Sric), D(p.Lbar), D(p.acc));
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


// One specialisation per compiled dimension tuple (defined in lqgk_inst.cu).
template <int X, int B, int U, int Y, int D>
struct Runner {
  static int run_f32(const Call& c);
  static int run_f64(const Call& c);
  static size_t plan_bytes(const LqgkDims& d, int mode, int32_t max_chunk);
};

}  // namespace lqgk
