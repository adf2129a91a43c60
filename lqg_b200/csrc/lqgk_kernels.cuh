// lqgk_kernels.cuh -- __global__ kernels for sm_100a.
//
// Mapping:
//  * per-sample FP64 recursions (k_lqr_fwd, k_kf_fwd, k_cov_fwd and their adjoints): ONE THREAD PER PARAMETER
//    SAMPLE, 32 samples per single-warp CTA.  All matrices live in registers as statically indexed arrays, the
//    per-sample derived constants and cotangent accumulators in shared memory laid out [element][lane] (bank =
//    lane, conflict free).  Per-step state is exchanged through HBM workspace arrays laid out sample-minor
//    [t][element][sample] so every warp access is one coalesced 256-byte line.
//  * per-trial FP32 recursions (k_trial_fwd, k_trial_rev): ONE WARP PER SAMPLE, LANE = TRIAL (RT trials per lane).
//    The per-step operator records of the sample ([t][REC] floats, contiguous) are staged into shared memory by
//    the TMA engine with 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx) in a per-warp ring, and
//    read as warp-wide broadcasts.  Observations are time-major so lane=trial loads are coalesced.
#pragma once
#include <cuda_runtime.h>

#include "lqgk_core.h"
#include "lqgk_pack.h"
#include "lqgk_stages.h"

namespace lqgk {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared through the TMA engine (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// =========================================================================================== boundary kernels
template <class T>
__global__ void k_pack(PackArgs<T> a, int s0, int n_valid, int n_pad, double* cst, size_t Sc, size_t tstride, int nt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int t = blockIdx.y;
  if (i >= n_pad || t >= nt) return;
  int s = s0 + (i < n_valid ? i : n_valid - 1);   // padding lanes replicate the last valid sample
  WView out{cst + (size_t)t * tstride + i, Sc};
  pack_sample<T>(a, s, t, [&](int e) -> double& { return out(e); });
}

template <class T>
__global__ void k_unpack(UnpackArgs<T> a, int s0, int n_valid, const double* acc, const double* cst, size_t Sc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_valid) return;
  unpack_sample<T>(a, s0 + i, [&](int e) { return acc[(size_t)e * Sc + i]; }, [&](int e) { return cst[(size_t)e * Sc + i]; });
}

// SoA workspace [t][E][Sc] -> user layout [s][t][E]
template <class T>
__global__ void k_store_rows(const double* ws, size_t Sc, int n_valid, int Tn, int E, T* out /* already offset by s0 */) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)n_valid * Tn * E;
  if (idx >= total) return;
  int e = idx % E;
  int t = (idx / E) % Tn;
  int s = idx / ((size_t)E * Tn);
  out[idx] = (T)ws[((size_t)t * E + e) * Sc + s];
}
template <class T>
__global__ void k_store_ll(const double* ll_ws, size_t count, T* out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count) out[idx] = (T)ll_ws[idx];
}
template <class T>
__global__ void k_load_w(const T* ll_bar /* nullable, already offset */, size_t count, float* w) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count) w[idx] = ll_bar ? (float)ll_bar[idx] : 1.f;
}
template <class T>
__global__ void k_pack_obs(int N, int T1, int d, const T* x, float* x_tm) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)N * T1 * d;
  if (idx >= total) return;
  int k = idx % d;
  int i = (idx / d) % N;
  int t = idx / ((size_t)d * N);
  x_tm[idx] = (float)x[((size_t)i * T1 + t) * d + k];
}

// =========================================================================================== per-sample kernels
// All launched with blockDim = 32 and gridDim = Sc / 32 (Sc is padded to a multiple of 32 by k_pack).
template <class DM, bool AFFINE>
__global__ void __launch_bounds__(32) k_lqr_fwd(const double* cst, size_t Sc, size_t tstride, int Tn, double eps, double* L,
                                                int save_S, double* Sric, double* l, double* H) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x;
  const size_t s = (size_t)blockIdx.x * 32 + lane;
  lqr_fwd_body<DM, AFFINE>(GCst{cst + s, Sc, tstride}, WView{sm + lane, 32}, Tn, eps, WView{L + s, Sc}, save_S != 0,
                           WView{Sric + s, Sc}, WView{l + s, Sc}, WView{H + s, Sc});
}

template <class DM>
__global__ void __launch_bounds__(32) k_kf_fwd(const double* cst, size_t Sc, size_t tstride, int Tn, double* K, int save_P,
                                               double* Pkf) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x;
  const size_t s = (size_t)blockIdx.x * 32 + lane;
  kf_fwd_body<DM>(GCst{cst + s, Sc, tstride}, WView{sm + lane, 32}, Tn, WView{K + s, Sc}, save_P != 0, WView{Pkf + s, Sc});
}

// Record sink: each lane writes its sample's record into a padded shared-memory row, then the warp copies the 32
// rows to HBM ([sample][t][REC], 128-byte coalesced stores).
template <class DM>
struct SmemRecSink {
  static constexpr int RS = DM::REC + 1;   // odd row stride -> lane-strided accesses hit 32 distinct banks
  float* stage;
  float* gbase;   // rec + (first sample of this warp) * Tn * REC
  int lane, Tn;
  __device__ __forceinline__ void put(int idx, float v) { stage[lane * RS + idx] = v; }
  __device__ __forceinline__ void commit(int t) {
    __syncwarp();
    for (int j = 0; j < 32; ++j) {
      float* dst = gbase + ((size_t)j * Tn + t) * DM::REC;
      for (int i = lane; i < DM::REC; i += 32) dst[i] = stage[j * RS + i];
    }
    __syncwarp();
  }
};

template <class DM>
__global__ void __launch_bounds__(32) k_cov_fwd(const double* cst, size_t Sc, size_t tstride, int Tn, const double* L,
                                                const double* K, int save_C, double* Cs, float* rec) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x;
  const size_t s0 = (size_t)blockIdx.x * 32, s = s0 + lane;
  float* stage = reinterpret_cast<float*>(sm + CovC<DM>::n * 32);
  for (int i = lane; i < 32 * SmemRecSink<DM>::RS; i += 32) stage[i] = 0.f;
  __syncwarp();
  SmemRecSink<DM> sink{stage, rec + s0 * Tn * DM::REC, lane, Tn};
  cov_fwd_body<DM>(GCst{cst + s, Sc, tstride}, WView{sm + lane, 32}, Tn, WView{const_cast<double*>(L) + s, Sc},
                   WView{const_cast<double*>(K) + s, Sc}, save_C != 0, WView{Cs + s, Sc}, sink);
}

template <class DM>
struct SmemSumSrc {
  static constexpr int RS = DM::SUMP + 1;
  float* stage;
  const float* gbase;   // sums + (first sample of this warp) * Tn * SUMP
  int lane, Tn;
  __device__ __forceinline__ void fetch(int t) {
    __syncwarp();
    for (int j = 0; j < 32; ++j) {
      const float* src = gbase + ((size_t)j * Tn + t) * DM::SUMP;
      for (int i = lane; i < DM::SUMP; i += 32) stage[j * RS + i] = src[i];
    }
    __syncwarp();
  }
  __device__ __forceinline__ float get(int idx) const { return stage[lane * RS + idx]; }
};

template <class DM>
__global__ void __launch_bounds__(32) k_cov_rev(const double* cst, size_t Sc, int Tn, int N, const float* w, const double* L,
                                                const double* K, const double* Cs, const float* sums, double* Lbar,
                                                double* Kbar, double* acc) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x;
  const size_t s0 = (size_t)blockIdx.x * 32, s = s0 + lane;
  double* la = sm + CovC<DM>::n * 32;
  float* stage = reinterpret_cast<float*>(la + CovC<DM>::n * 32);
  double sw = 0.0;
  for (int i = 0; i < N; ++i) sw += (double)w[s * N + i];
  SmemSumSrc<DM> src{stage, sums + s0 * Tn * DM::SUMP, lane, Tn};
  auto cv = [&](const double* p) { return WView{const_cast<double*>(p) + s, Sc}; };
  cov_rev_body<DM>(GCst{cst + s, Sc, 0}, WView{sm + lane, 32}, WView{la + lane, 32}, Tn, sw, cv(L), cv(K), cv(Cs), src,
                   cv(Lbar), cv(Kbar), cv(acc));
}

template <class DM>
__global__ void __launch_bounds__(32) k_kf_rev(const double* cst, size_t Sc, int Tn, const double* Pkf, const double* Kbar,
                                               double* acc) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x;
  const size_t s = (size_t)blockIdx.x * 32 + lane;
  double* la = sm + KfC<DM>::n * 32;
  auto cv = [&](const double* p) { return WView{const_cast<double*>(p) + s, Sc}; };
  kf_rev_body<DM>(GCst{cst + s, Sc, 0}, WView{sm + lane, 32}, WView{la + lane, 32}, Tn, cv(Pkf), cv(Kbar), cv(acc));
}

template <class DM>
__global__ void __launch_bounds__(32) k_lqr_rev(const double* cst, size_t Sc, int Tn, double eps, const double* L,
                                                const double* Sric, const double* Lbar, double* acc) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x;
  const size_t s = (size_t)blockIdx.x * 32 + lane;
  double* la = sm + LqrC<DM>::n * 32;
  auto cv = [&](const double* p) { return WView{const_cast<double*>(p) + s, Sc}; };
  lqr_rev_body<DM>(GCst{cst + s, Sc, 0}, WView{sm + lane, 32}, WView{la + lane, 32}, Tn, eps, cv(L), cv(Sric), cv(Lbar),
                   cv(acc));
}

// =========================================================================================== per-trial kernels
template <int D>
__device__ __forceinline__ void load_obs(const float* __restrict__ p, float* out) {
  if constexpr (D % 4 == 0) {
    LQGK_UNROLL for (int k = 0; k < D / 4; ++k) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p) + k);
      out[4 * k] = v.x; out[4 * k + 1] = v.y; out[4 * k + 2] = v.z; out[4 * k + 3] = v.w;
    }
  } else if constexpr (D % 2 == 0) {
    LQGK_UNROLL for (int k = 0; k < D / 2; ++k) {
      float2 v = __ldg(reinterpret_cast<const float2*>(p) + k);
      out[2 * k] = v.x; out[2 * k + 1] = v.y;
    }
  } else {
    LQGK_UNROLL for (int k = 0; k < D; ++k) out[k] = __ldg(p + k);
  }
}

constexpr int TRIAL_WARPS = 4;   // warps (= samples) per CTA
constexpr int TRIAL_TB = 8;      // time steps per ring stage
constexpr int TRIAL_NST = 3;     // ring stages

template <class DM>
constexpr size_t trial_smem_bytes() {
  return (size_t)TRIAL_WARPS * TRIAL_NST * TRIAL_TB * DM::REC * sizeof(float) + TRIAL_WARPS * TRIAL_NST * sizeof(uint64_t);
}

// Per-warp ring of record chunks filled by bulk async copies.  Chunk k covers steps [k*TB, min(T,(k+1)*TB)).
// `forward` walks chunks 0..nchunk-1, otherwise nchunk-1..0.
template <class DM>
struct RecRing {
  float* buf;        // [NST][TB*REC]
  uint64_t* bars;    // [NST]
  const float* src;  // this sample's records [Tn][REC]
  int Tn, nchunk, lane;
  uint32_t phase_bits;
  __device__ __forceinline__ void init() {
    if (lane == 0) {
      for (int i = 0; i < TRIAL_NST; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
    }
    phase_bits = 0;
    __syncwarp();
  }
  // issue the copy of chunk `k` into stage `st` (lane 0 only; caller guarantees the stage is no longer read)
  __device__ __forceinline__ void issue(int k, int st) {
    if (lane == 0) {
      int t0 = k * TRIAL_TB;
      int nst = min(TRIAL_TB, Tn - t0);
      uint32_t bytes = (uint32_t)nst * DM::REC * sizeof(float);
      fence_proxy_async();
      mbar_expect_tx(&bars[st], bytes);
      bulk_g2s(buf + (size_t)st * TRIAL_TB * DM::REC, src + (size_t)t0 * DM::REC, bytes, &bars[st]);
    }
  }
  __device__ __forceinline__ const float* wait(int st) {
    mbar_wait(&bars[st], (phase_bits >> st) & 1u);
    phase_bits ^= (1u << st);
    return buf + (size_t)st * TRIAL_TB * DM::REC;
  }
};

// Forward: per-trial mean recursion + log-density.  grid = (ceil(n_samples / TRIAL_WARPS)), block = 32 * TRIAL_WARPS.
// Trials are processed in passes of 32*RT (lane owns trials base + lane + 32*j, j < RT).
template <class DM, int RT>
__global__ void __launch_bounds__(32 * TRIAL_WARPS) k_trial_fwd(const float* __restrict__ rec, const float* __restrict__ x_tm,
                                                                int n_samples, int N, int Tn, double* __restrict__ ll_ws,
                                                                float* __restrict__ hist) {
  constexpr int D = DM::D, R = DM::R;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * TRIAL_WARPS + warp;
  if (s >= n_samples) return;
  float* ring_base = reinterpret_cast<float*>(smraw);
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(smraw + (size_t)TRIAL_WARPS * TRIAL_NST * TRIAL_TB * DM::REC * sizeof(float));
  RecRing<DM> ring{ring_base + (size_t)warp * TRIAL_NST * TRIAL_TB * DM::REC, bar_base + warp * TRIAL_NST,
                   rec + (size_t)s * Tn * DM::REC, Tn, (Tn + TRIAL_TB - 1) / TRIAL_TB, lane, 0};
  ring.init();
  const int nchunk = ring.nchunk;
  for (int base = 0; base < N; base += 32 * RT) {
    int tr[RT];
    bool ok[RT];
    LQGK_UNROLL for (int j = 0; j < RT; ++j) {
      int i = base + lane + 32 * j;
      ok[j] = i < N;
      tr[j] = ok[j] ? i : N - 1;
    }
    float c[RT][R], x0[RT][D], x1[RT][D];
    double ll[RT];
    LQGK_UNROLL for (int j = 0; j < RT; ++j) {
      ll[j] = 0.0;
      LQGK_UNROLL for (int k = 0; k < R; ++k) c[j][k] = 0.f;
      load_obs<D>(x_tm + (size_t)tr[j] * D, x0[j]);
    }
    for (int k = 0; k < TRIAL_NST && k < nchunk; ++k) ring.issue(k, k);
    for (int k = 0; k < nchunk; ++k) {
      const int st = k % TRIAL_NST;
      const float* chunk = ring.wait(st);
      const int t0 = k * TRIAL_TB, nst = min(TRIAL_TB, Tn - t0);
      for (int q = 0; q < nst; ++q) {
        const int t = t0 + q;
        const float* r = chunk + q * DM::REC;
        LQGK_UNROLL for (int j = 0; j < RT; ++j) load_obs<D>(x_tm + ((size_t)(t + 1) * N + tr[j]) * D, x1[j]);
        if (hist != nullptr) {
          LQGK_UNROLL for (int j = 0; j < RT; ++j)
            if (ok[j]) {
              LQGK_UNROLL for (int m = 0; m < R; ++m) hist[(((size_t)s * Tn + t) * R + m) * N + tr[j]] = c[j][m];
            }
        }
        LQGK_UNROLL for (int j = 0; j < RT; ++j) {
          ll[j] += (double)Trial<DM>::fwd(r, x0[j], x1[j], c[j]);
          LQGK_UNROLL for (int m = 0; m < D; ++m) x0[j][m] = x1[j][m];
        }
      }
      __syncwarp();
      if (k + TRIAL_NST < nchunk) ring.issue(k + TRIAL_NST, st);
    }
    LQGK_UNROLL for (int j = 0; j < RT; ++j)
      if (ok[j]) ll_ws[(size_t)s * N + tr[j]] = ll[j];
  }
}

// Transposing butterfly: every lane holds 32 partial values; afterwards lane L holds the warp-wide total of value L.
__device__ __forceinline__ float warp_transpose_reduce(float (&val)[32], int lane) {
  LQGK_UNROLL for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
    LQGK_UNROLL for (int j = 0; j < h; ++j) {
      float send = up ? val[j] : val[j + h];
      float keep = up ? val[j + h] : val[j];
      val[j] = keep + __shfl_xor_sync(FULL, send, h);
    }
  }
  return val[0];
}

// Reverse: per-trial adjoint, t = T-1..0, plus the per-step sums over trials (DM::SUM_* layout) written to
// sums[s][t][SUMP].  Passes over trial blocks accumulate (+=) into the sums.
template <class DM, int RT>
__global__ void __launch_bounds__(32 * TRIAL_WARPS) k_trial_rev(const float* __restrict__ rec, const float* __restrict__ x_tm,
                                                                const float* __restrict__ hist, const float* __restrict__ w,
                                                                int n_samples, int N, int Tn, float* __restrict__ sums) {
  constexpr int D = DM::D, R = DM::R;
  constexpr int NG = (DM::NSUM + 31) / 32;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * TRIAL_WARPS + warp;
  if (s >= n_samples) return;
  float* ring_base = reinterpret_cast<float*>(smraw);
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(smraw + (size_t)TRIAL_WARPS * TRIAL_NST * TRIAL_TB * DM::REC * sizeof(float));
  RecRing<DM> ring{ring_base + (size_t)warp * TRIAL_NST * TRIAL_TB * DM::REC, bar_base + warp * TRIAL_NST,
                   rec + (size_t)s * Tn * DM::REC, Tn, (Tn + TRIAL_TB - 1) / TRIAL_TB, lane, 0};
  ring.init();
  const int nchunk = ring.nchunk;
  for (int base = 0; base < N; base += 32 * RT) {
    int tr[RT];
    float wt[RT];
    LQGK_UNROLL for (int j = 0; j < RT; ++j) {
      int i = base + lane + 32 * j;
      bool ok = i < N;
      tr[j] = ok ? i : N - 1;
      wt[j] = ok ? w[(size_t)s * N + tr[j]] : 0.f;   // masked trials contribute nothing (cb stays 0, w = 0)
    }
    float cb[RT][R], x1[RT][D];
    LQGK_UNROLL for (int j = 0; j < RT; ++j) {
      LQGK_UNROLL for (int k = 0; k < R; ++k) cb[j][k] = 0.f;
      load_obs<D>(x_tm + ((size_t)Tn * N + tr[j]) * D, x1[j]);
    }
    for (int k = 0; k < TRIAL_NST && k < nchunk; ++k) ring.issue(nchunk - 1 - k, k);
    for (int kk = 0; kk < nchunk; ++kk) {
      const int k = nchunk - 1 - kk;
      const int st = kk % TRIAL_NST;
      const float* chunk = ring.wait(st);
      const int t0 = k * TRIAL_TB, nst = min(TRIAL_TB, Tn - t0);
      for (int q = nst - 1; q >= 0; --q) {
        const int t = t0 + q;
        const float* r = chunk + q * DM::REC;
        float c[RT][R], x0[RT][D], e[RT][D], v[RT][D], eb[RT][D], cbn[RT][R];
        LQGK_UNROLL for (int j = 0; j < RT; ++j) {
          load_obs<D>(x_tm + ((size_t)t * N + tr[j]) * D, x0[j]);
          LQGK_UNROLL for (int m = 0; m < R; ++m) c[j][m] = __ldg(&hist[(((size_t)s * Tn + t) * R + m) * N + tr[j]]);
        }
        LQGK_UNROLL for (int j = 0; j < RT; ++j) Trial<DM>::rev(r, x0[j], x1[j], c[j], wt[j], cb[j], e[j], v[j], eb[j], cbn[j]);
        float* out = sums + ((size_t)s * Tn + t) * DM::SUMP;
        static_for<0, NG>([&](auto G) {
          float val[32];
          static_for<0, 32>([&](auto J) {
            constexpr int IDX = decltype(G)::value * 32 + decltype(J)::value;
            float a = 0.f;
            LQGK_UNROLL for (int j = 0; j < RT; ++j)
              a += Trial<DM>::template sum_term<IDX>(cb[j], eb[j], x0[j], c[j], e[j], v[j], wt[j]);
            val[decltype(J)::value] = a;
          });
          float tot = warp_transpose_reduce(val, lane);
          const int idx = decltype(G)::value * 32 + lane;
          if (idx < DM::SUMP) {
            if (base == 0) out[idx] = tot;
            else out[idx] += tot;
          }
        });
        LQGK_UNROLL for (int j = 0; j < RT; ++j) {
          LQGK_UNROLL for (int m = 0; m < R; ++m) cb[j][m] = cbn[j][m];
          LQGK_UNROLL for (int m = 0; m < D; ++m) x1[j][m] = x0[j][m];
        }
      }
      __syncwarp();
      if (kk + TRIAL_NST < nchunk) ring.issue(nchunk - 1 - (kk + TRIAL_NST), st);
    }
  }
}

}  // namespace lqgk
