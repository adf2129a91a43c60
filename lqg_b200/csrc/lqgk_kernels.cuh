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
// 1-D bulk async copy shared -> global (TMA store, bulk-group completion).
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Per-thread asynchronous global->shared copies (LDGSTS) used as a register-free software prefetch.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

template <class KC, class La>
__device__ __forceinline__ void flush_acc_atomic(double* g, size_t stride, const La& l, int nseg) {
  for (int i = 0; i < nseg; ++i) {
    int go, lo, len;
    KC::seg(i, go, lo, len);
    for (int e = 0; e < len; ++e) atomicAdd(&g[(size_t)(go + e) * stride], l(lo + e));
  }
}

// ------------------------------------------------------------------------------------------- step-input ring
// The sequential per-sample kernels are latency-bound if each step's inputs are fetched with ordinary loads (one
// HBM round trip per dependent phase, ~1 warp per scheduler).  Instead each warp keeps a ring of NST "step input"
// stages in shared memory that the TMA engine fills ahead of time: for a warp's 32 consecutive samples, row e of
// step t of a sample-minor array is one contiguous 256-byte line, copied with one cp.async.bulk issued by lane
// (e mod 32); an optional float source copies one [CNTF]-float row per sample (lane = sample) from the
// sample-major sums array.  The stage layout is [row][32 lanes] doubles (lane-strided, bank-conflict free).
struct RingSrc {
  const double* base;   // array base + first sample of the warp
  int rows;             // rows (elements) per time step
};
template <int NSRC, int NST>
struct StepRing {
  RingSrc src[NSRC];
  size_t Sc;
  const float* fsrc;    // optional: sums + s0 * Tn * SUMP + LO   (nullptr = none)
  int frow_stride;      // Tn * SUMP (floats between consecutive samples)
  int fstep_stride;     // SUMP
  int fcnt;             // floats per row copy (multiple of 4)
  double* buf;          // [NST][stage_doubles]
  uint64_t* bars;       // [NST]
  int lane, total_rows, stage_doubles;
  uint32_t phase_bits;

  __host__ __device__ __forceinline__ static constexpr int stage_size(int total_rows, int fcnt) { return total_rows * 32 + (32 * fcnt + 1) / 2; }
  __device__ __forceinline__ void init() {
    if (lane == 0) {
      for (int i = 0; i < NST; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
    }
    phase_bits = 0;
    __syncwarp();
  }
  // Caller guarantees (by __syncwarp) that no lane still reads stage `st`.
  __device__ __forceinline__ void issue(int t, int st) {
    double* dst = buf + (size_t)st * stage_doubles;
    fence_proxy_async();
    if (lane == 0) mbar_expect_tx(&bars[st], (uint32_t)(total_rows * 256 + (fsrc ? 32 * fcnt * 4 : 0)));
    int row = 0;
    LQGK_UNROLL for (int i = 0; i < NSRC; ++i) {
      for (int r = lane; r < src[i].rows; r += 32)
        bulk_g2s(dst + (size_t)(row + r) * 32, src[i].base + ((size_t)t * src[i].rows + r) * Sc, 256, &bars[st]);
      row += src[i].rows;
    }
    if (fsrc) {
      float* fdst = reinterpret_cast<float*>(dst + (size_t)total_rows * 32);
      bulk_g2s(fdst + lane * fcnt, fsrc + (size_t)lane * frow_stride + (size_t)t * fstep_stride, (uint32_t)fcnt * 4, &bars[st]);
    }
  }
  __device__ __forceinline__ const double* wait(int st) {
    mbar_wait(&bars[st], (phase_bits >> st) & 1u);
    phase_bits ^= (1u << st);
    return buf + (size_t)st * stage_doubles;
  }
  // view of source-row offset `row_off` for this lane
  __device__ __forceinline__ WView view(const double* stage, int row_off) const {
    return WView{const_cast<double*>(stage) + (size_t)row_off * 32 + lane, 32};
  }
  __device__ __forceinline__ const float* frow(const double* stage) const {
    return reinterpret_cast<const float*>(stage + (size_t)total_rows * 32) + lane * fcnt;
  }
};

// Record sink: each lane assembles its sample's record ([REC] floats, one 16-byte-aligned row) in shared memory and
// hands it to the TMA engine: one bulk store per lane and step into rec[sample][t][REC].  Two row buffers alternate so
// the store of step t overlaps the arithmetic of step t+1 (no per-element copy loop, no warp synchronisation).
template <class DM>
struct SmemRecSink {
  static constexpr int RS = DM::REC;          // floats per row
  static constexpr int FLOATS = 2 * 32 * RS;  // two buffers
  float* stage;
  float* gbase;   // rec + (this lane's sample) * Tn * REC
  int lane, buf;
  __device__ __forceinline__ void put(int idx, float v) { stage[(buf * 32 + lane) * RS + idx] = v; }
  __device__ __forceinline__ void commit(int t) {
    fence_proxy_async();                       // my generic-proxy writes -> visible to the async proxy
    bulk_s2g(gbase + (size_t)t * DM::REC, stage + (buf * 32 + lane) * RS, DM::REC * sizeof(float));
    bulk_commit();
    buf ^= 1;
    bulk_wait_read<1>();                       // the buffer about to be rewritten (store of step t-1) has been read
  }
  __device__ __forceinline__ void finish() { bulk_wait<0>(); }
};

constexpr int SEQ_NST = 3;   // ring depth of the sequential kernels

// Covariance pass (forward).  Ring inputs per step: L_t, K_t.
template <class DM>
__global__ void __launch_bounds__(32) k_cov_fwd(const double* cst, size_t Sc, size_t tstride, int Tn, const double* L,
                                                const double* K, int save_adj, double* Cs, double* FU, double* JS, double* J0,
                                                float* rec) {
  extern __shared__ __align__(128) double sm[];
  constexpr int B = DM::B, U = DM::U, Y = DM::Y, R = DM::R, D = DM::D;
  using C = CovC<DM>;
  using SR = CovSeqRev<DM>;
  using Ring = StepRing<2, SEQ_NST>;
  const int lane = threadIdx.x;
  const size_t s0 = (size_t)blockIdx.x * 32, s = s0 + lane;
  constexpr int ROWS = DM::EL + DM::EK;
  double* ring_buf = sm;
  double* lcp = ring_buf + (size_t)SEQ_NST * ROWS * 32;
  float* stage = reinterpret_cast<float*>(lcp + C::n * 32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + SmemRecSink<DM>::FLOATS);
  for (int i = lane; i < SmemRecSink<DM>::FLOATS; i += 32) stage[i] = 0.f;
  __syncwarp();
  Ring ring{{{L + s0, DM::EL}, {K + s0, DM::EK}}, Sc, nullptr, 0, 0, 0, ring_buf, bars, lane, ROWS, ROWS * 32, 0};
  ring.init();
  for (int k = 0; k < SEQ_NST && k < Tn; ++k) ring.issue(k, k);
  WView lc{lcp + lane, 32};
  GCst g{cst + s, Sc, tstride};
  load_consts<C>(g.at(0), lc, C::NSEG);
  SmemRecSink<DM> sink{stage, rec + s * Tn * DM::REC, lane, 0};
  double Cm[R * R];
  {
    double K0[B * Y], J0v[R * D];
    LQGK_UNROLL for (int i = 0; i < B * Y; ++i) K0[i] = K[(size_t)i * Sc + s];
    CovFwd<DM>::init(lc, K0, Cm, J0v);
    if (save_adj) { LQGK_UNROLL for (int i = 0; i < R * D; ++i) J0[(size_t)i * Sc + s] = J0v[i]; }
  }
  // (constants stay in shared memory here: a register copy of the 40 covariance constants made this kernel 10 % slower)
  for (int t = 0; t < Tn; ++t) {
    const int st = t % SEQ_NST;
    const double* stg = ring.wait(st);
    if (tstride && t != 0) load_consts<C>(g.at(t), lc, C::NSEG);
    double Lt[U * B], Kt[B * Y];
    {
      WView lv = ring.view(stg, 0), kv = ring.view(stg, DM::EL);
      LQGK_UNROLL for (int i = 0; i < U * B; ++i) Lt[i] = lv(i);
      LQGK_UNROLL for (int i = 0; i < B * Y; ++i) Kt[i] = kv(i);
    }
    __syncwarp();
    if (t + SEQ_NST < Tn) ring.issue(t + SEQ_NST, st);
    if (save_adj) store_sym<R>(WView{Cs + s, Sc}, (size_t)t * DM::EC, Cm);
    CovFwd<DM>::step(lc, Lt, Kt, Cm, [&](int idx, float v) { sink.put(idx, v); },
                     [&](int which, int e, double v) {
                       if (save_adj) {
                         if (which == 0) FU[((size_t)t * SR::NSF + e) * Sc + s] = v;
                         else JS[((size_t)t * SR::NJS + e) * Sc + s] = v;
                       }
                     });
    sink.commit(t);
  }
  sink.finish();
}

// Sequential covariance adjoint: lean (no constants, no accumulators); emits Sgb_t, SF_t for the parallel contraction.
// Ring inputs per step: Fu_t, (J_t, S'^-1_t) and the [SUM_J, SUMP) tail of the trial sums.
template <class DM>
__global__ void __launch_bounds__(32) k_cov_seq_rev(size_t Sc, int Tn, int N, const float* w, const double* FU, const double* JS,
                                                    const double* J0, const float* sums, double* SGB, double* SGBI, double* SFW) {
  extern __shared__ __align__(128) double sm[];
  using SR = CovSeqRev<DM>;
  using Ring = StepRing<2, SEQ_NST>;
  constexpr int R = DM::R;
  constexpr int ROWS = SR::NSF + SR::NJS, FCNT = DM::SUMP - DM::SUM_J;
  constexpr int STAGE = Ring::stage_size(ROWS, FCNT);
  const int lane = threadIdx.x;
  const size_t s0 = (size_t)blockIdx.x * 32, s = s0 + lane;
  double* scp = sm + (size_t)SEQ_NST * STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scp + SR::SC_N * 32);
  Ring ring{{{FU + s0, SR::NSF}, {JS + s0, SR::NJS}}, Sc, sums + s0 * Tn * DM::SUMP + DM::SUM_J, Tn * DM::SUMP, DM::SUMP, FCNT,
            sm, bars, lane, ROWS, STAGE, 0};
  ring.init();
  for (int k = 0; k < SEQ_NST && k < Tn; ++k) ring.issue(Tn - 1 - k, k);
  double sw = 0.0;
  for (int i = 0; i < N; ++i) sw += (double)w[s * N + i];
  WView sc{scp + lane, 32};
  double Cb[R * R];
  LQGK_UNROLL for (int i = 0; i < R * R; ++i) Cb[i] = 0.0;
  for (int kk = 0; kk < Tn; ++kk) {
    const int t = Tn - 1 - kk, st = kk % SEQ_NST;
    const double* stg = ring.wait(st);
    WView fuv = ring.view(stg, 0), jsv = ring.view(stg, SR::NSF);
    const float* fr = ring.frow(stg);
    SR::step([&](int e) { return fuv(e); }, [&](int e) { return jsv(e); }, [&](int idx) { return fr[idx - DM::SUM_J]; }, sw, sc, Cb,
             [&](int e, double v) { SGB[((size_t)t * SR::NSGB + e) * Sc + s] = v; },
             [&](int e, double v) { SFW[((size_t)t * SR::NSF + e) * Sc + s] = v; });
    __syncwarp();
    if (kk + SEQ_NST < Tn) ring.issue(Tn - 1 - (kk + SEQ_NST), st);
  }
  SR::init([&](int e) { return J0[(size_t)e * Sc + s]; }, Cb, [&](int e, double v) { SGBI[(size_t)e * Sc + s] = v; });
}

// Time-parallel contraction.  grid = (Sc / 32, chunks), block = one warp owning a contiguous range of time steps of
// its 32 samples.  Ring inputs per step: L_t, K_t, C_t, SF_t (, Sgb_t) and the trial sums F-block.  The contributions
// to the derived-constant cotangents are accumulated over the warp's time range in shared memory ([element][lane])
// and added to the global accumulators with one atomicAdd per element at the end.
// PASS 0 / 1: the two register-limited halves (run concurrently on two streams); PASS 2: both in one kernel (small systems).
template <class DM>
__host__ __device__ constexpr bool contrib_merged() { return DM::N <= 6; }
// ring depth of the contraction kernels: 2 stages when they fit in shared memory beside the constants, else 1
template <class DM, int PASS>
__host__ __device__ constexpr int contrib_nst() {
  using SR = CovSeqRev<DM>;
  constexpr int ROWS = DM::EL + DM::EK + DM::EC + SR::NSF + (PASS != 1 ? SR::NSGB : 0);
  constexpr size_t two = sizeof(double) * ((size_t)2 * (ROWS * 32 + 16 * DM::SUM_J) + 2 * CovC<DM>::n * 32);
  return two <= 200 * 1024 ? 2 : 1;
}
template <class DM, int PASS>
__global__ void __launch_bounds__(32) k_cov_contrib(const double* cst, size_t Sc, int Tn, const double* L, const double* K,
                                                    const double* Cs, const double* SGB, const double* SGBI, const double* SFW,
                                                    const float* sums, double* acc, double* Lbar, double* Kbar, double* KbarF) {
  extern __shared__ __align__(128) double sm[];
  constexpr int B = DM::B, U = DM::U, Y = DM::Y, R = DM::R;
  using SR = CovSeqRev<DM>;
  using CC = CovContrib<DM>;
  using C = CovC<DM>;
  constexpr bool P0 = PASS == 0 || PASS == 2, P1 = PASS == 1 || PASS == 2;
  constexpr int NSRC = P0 ? 5 : 4;
  constexpr int PAR_NST = contrib_nst<DM, PASS>();
  using Ring = StepRing<NSRC, PAR_NST>;
  constexpr int ROWS = DM::EL + DM::EK + DM::EC + SR::NSF + (P0 ? SR::NSGB : 0);
  constexpr int FCNT = DM::SUM_J;   // = round4(N*N)
  constexpr int STAGE = Ring::stage_size(ROWS, FCNT);
  const int lane = threadIdx.x;
  const size_t s0 = (size_t)blockIdx.x * 32, s = s0 + lane;
  double* lcp = sm + (size_t)PAR_NST * STAGE;
  double* lap = lcp + C::n * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lap + C::n * 32);
  const int nq = gridDim.y, q = blockIdx.y;
  const int per = (Tn + nq - 1) / nq;
  const int t0 = min(Tn, q * per), t1 = min(Tn, t0 + per);
  if (t0 >= t1) return;
  Ring ring;
  ring.src[0] = {L + s0, DM::EL};
  ring.src[1] = {K + s0, DM::EK};
  ring.src[2] = {Cs + s0, DM::EC};
  ring.src[3] = {SFW + s0, SR::NSF};
  if constexpr (P0) ring.src[4] = {SGB + s0, SR::NSGB};
  ring.Sc = Sc; ring.fsrc = sums + s0 * Tn * DM::SUMP; ring.frow_stride = Tn * DM::SUMP; ring.fstep_stride = DM::SUMP;
  ring.fcnt = FCNT; ring.buf = sm; ring.bars = bars; ring.lane = lane; ring.total_rows = ROWS; ring.stage_doubles = STAGE;
  ring.init();
  for (int k = 0; k < PAR_NST && t0 + k < t1; ++k) ring.issue(t0 + k, k);
  WView lc{lcp + lane, 32}, la{lap + lane, 32};
  load_consts<C>(WView{const_cast<double*>(cst) + s, Sc}, lc, C::NSEG);
  for (int e = 0; e < C::n; ++e) la(e) = 0.0;
  for (int t = t0; t < t1; ++t) {
    const int st = (t - t0) % PAR_NST;
    const double* stg = ring.wait(st);
    WView lv = ring.view(stg, 0), kv = ring.view(stg, DM::EL), cvw = ring.view(stg, DM::EL + DM::EK),
          sfv = ring.view(stg, DM::EL + DM::EK + DM::EC), sgv = ring.view(stg, DM::EL + DM::EK + DM::EC + SR::NSF);
    const float* fr = ring.frow(stg);
    double Cm[R * R], Lt[U * B], Kt[B * Y];
    LQGK_UNROLL for (int i = 0; i < U * B; ++i) Lt[i] = lv(i);
    LQGK_UNROLL for (int i = 0; i < B * Y; ++i) Kt[i] = kv(i);
    load_sym_ws<R>(cvw, 0, Cm);
    auto sf = [&](int e) { return sfv(e); };
    auto get = [&](int idx) { return fr[idx]; };
    auto out = [&](int e, double v) { la(e) += v; };
    double Kb[B * Y];
    if constexpr (P0) {
      double Lb[U * B];
      CC::pass0(lc, [&](int e) { return sgv(e); }, t == 0, [&](int e) { return SGBI[(size_t)e * Sc + s]; }, sf, get, Cm, Lt, Kt, out,
                Lb, Kb);
      LQGK_UNROLL for (int i = 0; i < U * B; ++i) Lbar[((size_t)t * DM::EL + i) * Sc + s] = Lb[i];
      if constexpr (!P1) { LQGK_UNROLL for (int i = 0; i < B * Y; ++i) Kbar[((size_t)t * DM::EK + i) * Sc + s] = Kb[i]; }
    } else {
      LQGK_UNROLL for (int i = 0; i < B * Y; ++i) Kb[i] = 0.0;
    }
    if constexpr (P1) {
      CC::pass1(lc, sf, get, Cm, Lt, Kt, out, Kb);
      // PASS 1 alone stores the transition part separately (k_kf_rev adds the two parts); PASS 2 stores the total
      double* dst = PASS == 1 ? KbarF : Kbar;
      LQGK_UNROLL for (int i = 0; i < B * Y; ++i) dst[((size_t)t * DM::EK + i) * Sc + s] = Kb[i];
      if constexpr (PASS == 2) { LQGK_UNROLL for (int i = 0; i < B * Y; ++i) KbarF[((size_t)t * DM::EK + i) * Sc + s] = 0.0; }
    }
    __syncwarp();
    if (t + PAR_NST < t1) ring.issue(t + PAR_NST, st);
  }
  flush_acc_atomic<C>(acc + s, Sc, la, C::NSEG);
}

// Kalman-gain adjoint (sequential, t descending).  Ring inputs per step: P_t, Kbar_t.
template <class DM>
__global__ void __launch_bounds__(32) k_kf_rev(const double* cst, size_t Sc, int Tn, const double* Pkf, const double* Kbar,
                                               const double* KbarF, double* acc) {
  extern __shared__ __align__(128) double sm[];
  constexpr int B = DM::B, Y = DM::Y;
  using C = KfC<DM>;
  using Ring = StepRing<3, SEQ_NST>;
  constexpr int ROWS = DM::EP + 2 * DM::EK;
  const int lane = threadIdx.x;
  const size_t s0 = (size_t)blockIdx.x * 32, s = s0 + lane;
  double* lcp = sm + (size_t)SEQ_NST * ROWS * 32;
  double* lap = lcp + C::n * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lap + C::n * 32);
  Ring ring{{{Pkf + s0, DM::EP}, {Kbar + s0, DM::EK}, {KbarF + s0, DM::EK}}, Sc, nullptr, 0, 0, 0, sm, bars, lane, ROWS, ROWS * 32, 0};
  ring.init();
  for (int k = 0; k < SEQ_NST && k < Tn; ++k) ring.issue(Tn - 1 - k, k);
  WView lc{lcp + lane, 32}, la{lap + lane, 32};
  load_consts<C>(WView{const_cast<double*>(cst) + s, Sc}, lc, C::NSEG);
  for (int e = 0; e < C::n; ++e) la(e) = 0.0;
  auto accf = [&](int e) -> double& { return la(e); };
  double Pnb[B * B];
  LQGK_UNROLL for (int i = 0; i < B * B; ++i) Pnb[i] = 0.0;
  auto sweep = [&](const auto& cv) {
    for (int kk = 0; kk < Tn; ++kk) {
      const int st = kk % SEQ_NST;
      const double* stg = ring.wait(st);
      double P[B * B], Kb[B * Y];
      load_sym_ws<B>(ring.view(stg, 0), 0, P);
      {
        WView kv = ring.view(stg, DM::EP), kv2 = ring.view(stg, DM::EP + DM::EK);
        LQGK_UNROLL for (int i = 0; i < B * Y; ++i) Kb[i] = kv(i) + kv2(i);
      }
      __syncwarp();
      if (kk + SEQ_NST < Tn) ring.issue(Tn - 1 - (kk + SEQ_NST), st);
      KfRev<DM>::step(cv, accf, P, Kb, Pnb);
    }
  };
  if constexpr (C::n <= LQGK_REG_CONSTS_MAX) {
    RegView<C::n> rc;
    LQGK_UNROLL for (int e = 0; e < C::n; ++e) rc.v[e] = lc(e);
    sweep(rc);
  } else {
    sweep(lc);
  }
  KfRev<DM>::finish(accf, Pnb);
  flush_acc_atomic<C>(acc + s, Sc, la, C::NSEG);
}

// Riccati adjoint (sequential, t ascending).  Ring inputs per step: S_{t+1}, L_t, Lbar_t.
template <class DM>
__global__ void __launch_bounds__(32) k_lqr_rev(const double* cst, size_t Sc, int Tn, double eps, const double* L,
                                                const double* Sric, const double* Lbar, double* acc) {
  extern __shared__ __align__(128) double sm[];
  constexpr int B = DM::B, U = DM::U;
  using C = LqrC<DM>;
  using Ring = StepRing<3, SEQ_NST>;
  constexpr int ROWS = DM::ES + 2 * DM::EL;
  const int lane = threadIdx.x;
  const size_t s0 = (size_t)blockIdx.x * 32, s = s0 + lane;
  double* lcp = sm + (size_t)SEQ_NST * ROWS * 32;
  double* lap = lcp + C::n * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lap + C::n * 32);
  Ring ring{{{Sric + s0, DM::ES}, {L + s0, DM::EL}, {Lbar + s0, DM::EL}}, Sc, nullptr, 0, 0, 0, sm, bars, lane, ROWS, ROWS * 32, 0};
  ring.init();
  for (int k = 0; k < SEQ_NST && k < Tn; ++k) ring.issue(k, k);
  WView lc{lcp + lane, 32}, la{lap + lane, 32};
  load_consts<C>(WView{const_cast<double*>(cst) + s, Sc}, lc, C::NSEG);
  for (int e = 0; e < C::n; ++e) la(e) = 0.0;
  auto accf = [&](int e) -> double& { return la(e); };
  double Sn[B * B];
  LQGK_UNROLL for (int i = 0; i < B * B; ++i) Sn[i] = 0.0;
  auto sweep = [&](const auto& cv) {
    for (int t = 0; t < Tn; ++t) {
      const int st = t % SEQ_NST;
      const double* stg = ring.wait(st);
      double S[B * B], Lt[U * B], Lb[U * B];
      load_sym_ws<B>(ring.view(stg, 0), 0, S);
      {
        WView lv = ring.view(stg, DM::ES), bv = ring.view(stg, DM::ES + DM::EL);
        LQGK_UNROLL for (int i = 0; i < U * B; ++i) { Lt[i] = lv(i); Lb[i] = bv(i); }
      }
      __syncwarp();
      if (t + SEQ_NST < Tn) ring.issue(t + SEQ_NST, st);
      double shift;
      {
        double Bm[B * U], SB[B * U], H[U * U];
        load_mat<B, U>(cv, C::Ba, Bm);
        mm<B, B, U>(S, Bm, SB);
        load_sym<U>(cv, C::R, H);
        mm_tn_sym<U, B, true>(Bm, SB, H);
        shift = eps - lambda_min<U>(H);
        shift = shift > 0.0 ? shift : 0.0;
      }
      LqrRev<DM>::step(cv, accf, S, Lt, Lb, shift, Sn);
    }
  };
  if constexpr (C::n <= LQGK_REG_CONSTS_MAX) {
    RegView<C::n> rc;
    LQGK_UNROLL for (int e = 0; e < C::n; ++e) rc.v[e] = lc(e);
    sweep(rc);
  } else {
    sweep(lc);
  }
  LqrRev<DM>::finish(accf, Sn);
  flush_acc_atomic<C>(acc + s, Sc, la, C::NSEG);
}

// Dynamic shared-memory sizes of the per-sample kernels (must mirror the carve-up inside each kernel).
template <class DM> constexpr size_t smem_cov_fwd() {
  return sizeof(double) * ((size_t)SEQ_NST * (DM::EL + DM::EK) * 32 + CovC<DM>::n * 32) +
         sizeof(float) * SmemRecSink<DM>::FLOATS + sizeof(uint64_t) * SEQ_NST;
}
template <class DM> constexpr size_t smem_cov_seq_rev() {
  using SR = CovSeqRev<DM>;
  return sizeof(double) * ((size_t)SEQ_NST * StepRing<2, SEQ_NST>::stage_size(SR::NSF + SR::NJS, DM::SUMP - DM::SUM_J) + SR::SC_N * 32) +
         sizeof(uint64_t) * SEQ_NST;
}
template <class DM, int PASS> constexpr size_t smem_cov_contrib() {
  using SR = CovSeqRev<DM>;
  constexpr int ROWS = DM::EL + DM::EK + DM::EC + SR::NSF + (PASS != 1 ? SR::NSGB : 0);
  constexpr int PAR_NST = contrib_nst<DM, PASS>();
  return sizeof(double) * ((size_t)PAR_NST * StepRing<5, PAR_NST>::stage_size(ROWS, DM::SUM_J) + 2 * CovC<DM>::n * 32) +
         sizeof(uint64_t) * PAR_NST;
}
template <class DM> constexpr size_t smem_kf_rev() {
  return sizeof(double) * ((size_t)SEQ_NST * (DM::EP + 2 * DM::EK) * 32 + 2 * KfC<DM>::n * 32) + sizeof(uint64_t) * SEQ_NST;
}
template <class DM> constexpr size_t smem_lqr_rev() {
  return sizeof(double) * ((size_t)SEQ_NST * (DM::ES + 2 * DM::EL) * 32 + 2 * LqrC<DM>::n * 32) + sizeof(uint64_t) * SEQ_NST;
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

// Resident CTAs per SM the forward per-trial kernel is compiled for: the small systems are latency-bound at 2 CTAs (8 warps)
// per SM, so the register budget is capped at 65536 / (3 * 128) = 170 to fit 3 (no spills: 204 -> 164 registers, -5 % time).
// The adjoint kernel needs ~240 registers; capped at 170 it spills and runs 1.6x slower (measured), so it stays at 2 CTAs.
template <class DM>
__host__ __device__ constexpr int trial_min_ctas() { return DM::N <= 6 ? 3 : 1; }
constexpr int TRIAL_WARPS = 4;   // warps (= samples) per CTA
// time steps per ring stage of the forward kernel: 8 for the small records, fewer when one record is KBs (large systems)
template <class DM>
__host__ __device__ constexpr int trial_tb() { return DM::REC <= 128 ? 8 : (DM::REC <= 256 ? 4 : 2); }
constexpr int TRIAL_NST = 3;     // ring stages of the forward kernel

constexpr int TRIAL_PF = 4;      // observation prefetch distance in time steps (cp.async groups in flight)
// Sums over trials in the adjoint: every lane holds 32 partial values; they are transposed through shared memory (lane L
// writes its 32 values as 8 x 128-bit stores into row L of a [32][36] tile -- conflict-free per quarter warp --, then sums
// column L: 32 conflict-free loads) instead of a 31-shuffle butterfly with 62 selects.
constexpr int TRIAL_RED_STRIDE = 36;

// ------------------------------------------------------------------------------------------- checkpointed state history
// The adjoint needs the carried mean c_t of every trial at every step.  Storing all of them (T x R x N floats per sample,
// 2.9 MB at the benchmark size) made the forward kernel HBM-bound and was 70 % of the step's DRAM traffic (round 1).  Now the
// forward kernel stores c_t only at the segment starts t = k * CK ("checkpoints"); the adjoint kernel processes one segment at
// a time: it re-runs the state recursion from the checkpoint (Trial::advance, CK - 1 steps, bit-identical to the forward pass)
// into a per-warp shared-memory buffer [CK][R][32 RT] and then walks the segment backwards reading c_t from that buffer.
// CK is the largest interval whose buffer + record ring fits the per-warp share of shared memory that keeps 2 CTAs (8 warps)
// resident per SM.
#ifndef LQGK_TRIAL_CK_MAX
#define LQGK_TRIAL_CK_MAX 8
#endif
constexpr int TRIAL_CK_MAX = LQGK_TRIAL_CK_MAX;
#ifndef LQGK_TRIAL_WARP_SMEM
#define LQGK_TRIAL_WARP_SMEM 28440    // bytes per warp: 2 CTAs x 4 warps x 28,440 B + 2 KB reserved < 228 KB per SM
#endif
template <class DM>
__host__ __device__ constexpr int trial_red_bytes() { return DM::NSUM >= 32 ? 32 * TRIAL_RED_STRIDE * (int)sizeof(float) : 0; }
template <class DM>
__host__ __device__ constexpr int trial_ck(int RT) {
  const int per_step = 2 * DM::REC * (int)sizeof(float) + DM::R * 32 * RT * (int)sizeof(float);   // 2 ring stages + 1 state slot
  const int ck = (LQGK_TRIAL_WARP_SMEM - trial_red_bytes<DM>() - 16) / per_step;
  return ck < 1 ? 1 : (ck > TRIAL_CK_MAX ? TRIAL_CK_MAX : ck);
}
// Largest number of trials one lane carries: small systems leave room for 8 (amortises the broadcast record loads and
// the cross-lane reductions over more FMAs), large ones for 4.
template <class DM>
__host__ __device__ constexpr int trial_rt_max() { return DM::N * DM::N <= 36 ? 8 : 4; }
__host__ __device__ __forceinline__ constexpr int trial_rt(int N, int rt_max) { return (N + 31) / 32 < rt_max ? (N + 31) / 32 : rt_max; }
// Row stride of the checkpoints [s][segment][R][hist_stride(N)]: even, so a pair's entry is 8-byte aligned.  The forward
// kernel stores 0 into the pad element of an odd N; the adjoint reads pairs whenever the first half is a real trial.
__host__ __device__ __forceinline__ constexpr int hist_stride(int N) { return (N + 1) & ~1; }
template <class DM>
__host__ __device__ constexpr int trial_nseg(int N, int T) {
  const int ck = trial_ck<DM>(trial_rt(N, trial_rt_max<DM>()));
  return (T + ck - 1) / ck;
}

template <class DM, int RT, bool REV>
constexpr size_t trial_smem_bytes() {
  if constexpr (REV) {
    constexpr int CK = trial_ck<DM>(RT);
    return (size_t)TRIAL_WARPS * (2 * CK * DM::REC * sizeof(float) + 2 * sizeof(uint64_t) + trial_red_bytes<DM>() +
                                  (size_t)CK * DM::R * 32 * RT * sizeof(float));
  } else {
    size_t ring = (size_t)TRIAL_WARPS * TRIAL_NST * trial_tb<DM>() * DM::REC * sizeof(float) + TRIAL_WARPS * TRIAL_NST * sizeof(uint64_t);
    size_t pf = (size_t)TRIAL_WARPS * (TRIAL_PF + 1) * RT * 32 * DM::D * sizeof(float);
    return ring + pf;
  }
}
// observation x[t][trial][0..D) -> this lane's slot (vector copy of D floats, D*4 in {4, 8, 16} bytes, else scalars)
template <int D>
__device__ __forceinline__ void prefetch_obs(float* slot, int lane, const float* src) {
  if constexpr (D == 1 || D == 2 || D == 4) cp_async<D * 4>(slot + lane * D, src);
  else { LQGK_UNROLL for (int k = 0; k < D; ++k) cp_async<4>(slot + k * 32 + lane, src + k); }
}
template <int D>
__device__ __forceinline__ void read_obs(const float* slot, int lane, float* out) {
  if constexpr (D == 2) {          // one 64-bit access per lane (scalar reads at stride 2 are 2-way bank conflicts)
    float2 v = *reinterpret_cast<const float2*>(slot + lane * 2);
    out[0] = v.x; out[1] = v.y;
  } else if constexpr (D == 4) {
    float4 v = *reinterpret_cast<const float4*>(slot + lane * 4);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else if constexpr (D == 1) {
    out[0] = slot[lane];
  } else { LQGK_UNROLL for (int k = 0; k < D; ++k) out[k] = slot[k * 32 + lane]; }
}

// Per-warp ring of NST record chunks of TB time steps, filled by bulk async copies.  Chunk k covers steps
// [k*TB, min(T,(k+1)*TB)).
template <class DM, int TB, int NST>
struct RecRing {
  float* buf;        // [NST][TB*REC]
  uint64_t* bars;    // [NST]
  const float* src;  // this sample's records [Tn][REC]
  int Tn, lane;
  uint32_t phase_bits;
  __device__ __forceinline__ void init() {
    if (lane == 0) {
      for (int i = 0; i < NST; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
    }
    phase_bits = 0;
    __syncwarp();
  }
  // issue the copy of chunk `k` into stage `st` (lane 0 only; caller guarantees the stage is no longer read)
  __device__ __forceinline__ void issue(int k, int st) {
    if (lane == 0) {
      int t0 = k * TB;
      int nst = min(TB, Tn - t0);
      uint32_t bytes = (uint32_t)nst * DM::REC * sizeof(float);
      fence_proxy_async();
      mbar_expect_tx(&bars[st], bytes);
      bulk_g2s(buf + (size_t)st * TB * DM::REC, src + (size_t)t0 * DM::REC, bytes, &bars[st]);
    }
  }
  __device__ __forceinline__ const float* wait(int st) {
    mbar_wait(&bars[st], (phase_bits >> st) & 1u);
    phase_bits ^= (1u << st);
    return buf + (size_t)st * TB * DM::REC;
  }
};

// Trial -> lane assignment of one pass (32*RT trials starting at `base`): pair p of a lane holds the ADJACENT trials
// base + 64p + 2*lane and +1, so one 64-bit access moves the pair's entry of a checkpoint, already in f32x2 register
// order (no packing moves, half the memory instructions); an odd RT adds the single trial base + 64*(RT/2) + lane.
// Flat slot j: 2p, 2p+1 = the two halves of pair p; RT-1 = the single.
template <int RT>
__device__ __forceinline__ int trial_of(int base, int lane, int j) {
  constexpr int NP = RT / 2;
  return j < 2 * NP ? base + 64 * (j >> 1) + 2 * lane + (j & 1) : base + 64 * NP + lane;
}

// Copy one step's record from the shared-memory ring into registers (small systems) with 128-bit broadcast loads.
template <class DM>
struct RecRegs {
  float v[DM::REC];
  __device__ __forceinline__ void load(const float* r) {
    const float4* r4 = reinterpret_cast<const float4*>(r);
    LQGK_UNROLL for (int k = 0; k < DM::REC / 4; ++k) {
      float4 q = r4[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  }
  __device__ __forceinline__ float operator[](int i) const { return v[i]; }
};
template <class DM>
__host__ __device__ constexpr bool rec_in_regs() { return DM::REC <= 48; }

// Forward: per-trial mean recursion + log-density.  grid = (ceil(n_samples / TRIAL_WARPS)), block = 32 * TRIAL_WARPS.
// Trials are processed in passes of 32*RT (lane owns the trials trial_of(base, lane, j), j < RT).  Observations are prefetched
// TRIAL_PF steps ahead with per-lane cp.async copies into shared-memory slots (no registers held, no stalls on the
// L2 round trip).  `hist` (VJP only): checkpoints of the carried state at the segment starts, [s][segment][R][hist_stride(N)].
template <class DM, int RT>
__global__ void __launch_bounds__(32 * TRIAL_WARPS, trial_min_ctas<DM>()) k_trial_fwd(const float* __restrict__ rec, const float* __restrict__ x_all,
                                                                size_t x_sample_stride, int s_first, int n_samples, int N, int Tn,
                                                                double* __restrict__ ll_ws, float* __restrict__ hist) {
  constexpr int D = DM::D, R = DM::R, NSLOT = TRIAL_PF + 1, TB = trial_tb<DM>(), CK = trial_ck<DM>(RT);
  extern __shared__ __align__(128) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * TRIAL_WARPS + warp;
  if (s >= n_samples) return;
  const float* __restrict__ x_tm = x_all + (size_t)(s_first + s) * x_sample_stride;   // this sample's observations
  constexpr size_t RING_BYTES = (size_t)TRIAL_WARPS * TRIAL_NST * TB * DM::REC * sizeof(float);
  float* ring_base = reinterpret_cast<float*>(smraw);
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(smraw + RING_BYTES);
  float* pf = reinterpret_cast<float*>(smraw + RING_BYTES + TRIAL_WARPS * TRIAL_NST * sizeof(uint64_t)) + (size_t)warp * NSLOT * RT * 32 * D;
  RecRing<DM, TB, TRIAL_NST> ring{ring_base + (size_t)warp * TRIAL_NST * TB * DM::REC, bar_base + warp * TRIAL_NST,
                                  rec + (size_t)s * Tn * DM::REC, Tn, lane, 0};
  ring.init();
  const int nchunk = (Tn + TB - 1) / TB;
  const int Nh = hist_stride(N);
  const int nseg = (Tn + CK - 1) / CK;
  for (int base = 0; base < N; base += 32 * RT) {
    int tr[RT], trD[RT];
    bool ok[RT];
    LQGK_UNROLL for (int j = 0; j < RT; ++j) {
      int i = trial_of<RT>(base, lane, j);
      ok[j] = i < N;
      tr[j] = ok[j] ? i : N - 1;
      trD[j] = tr[j] * D;            // 32-bit element offsets: one 64-bit row pointer per step + cheap lane offsets
    }
    // checkpoint row pointers of this lane (one per state component), advanced by one segment per stored checkpoint
    // (small systems: one pointer per component -> every access is pointer + immediate; large ones: one pointer + m * Nh)
    constexpr bool HPTRS = R <= 6;
    constexpr int NHP = HPTRS ? R : 1;
    float* hw[NHP];
    LQGK_UNROLL for (int m = 0; m < NHP; ++m) hw[m] = hist + ((size_t)s * nseg * R + m) * Nh + (base + 2 * lane);
    auto hrow = [&](int m) -> float* { if constexpr (HPTRS) return hw[m]; else return hw[0] + m * Nh; };
    const int hstep = R * Nh, hsingle = 64 * (RT / 2) - lane;
    int ckc = 0;                     // steps until the next checkpoint
    // trials 2p, 2p+1 are packed into one f32x2 lane-pair state (Blackwell FFMA2); an odd last trial stays scalar
    constexpr int NP = RT / 2, NS = RT % 2;
    f32x2 cP[NP > 0 ? NP : 1][R], x0P[NP > 0 ? NP : 1][D];
    float cS[R], x0S[D];
    double ll[RT];
    {
      float x0[RT][D];
      LQGK_UNROLL for (int j = 0; j < RT; ++j) {
        ll[j] = 0.0;
        load_obs<D>(x_tm + (size_t)tr[j] * D, x0[j]);
      }
      LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) {
        LQGK_UNROLL for (int k = 0; k < R; ++k) cP[p2][k] = f32x2{0.f, 0.f};
        LQGK_UNROLL for (int m = 0; m < D; ++m) x0P[p2][m] = f32x2{x0[2 * p2][m], x0[2 * p2 + 1][m]};
      }
      LQGK_UNROLL for (int k = 0; k < R; ++k) cS[k] = 0.f;
      LQGK_UNROLL for (int m = 0; m < D; ++m) x0S[m] = x0[RT - 1][m];
    }
    // prologue: x_{1..PF} in flight (one commit group per step)
    for (int p = 0; p < TRIAL_PF; ++p) {
      const float* xrow = x_tm + (size_t)min(1 + p, Tn) * N * D;
      LQGK_UNROLL for (int j = 0; j < RT; ++j) prefetch_obs<D>(pf + ((p % NSLOT) * RT + j) * 32 * D, lane, xrow + trD[j]);
      cp_async_commit();
    }
    for (int k = 0; k < TRIAL_NST && k < nchunk; ++k) ring.issue(k, k);
    for (int k = 0; k < nchunk; ++k) {
      const int st = k % TRIAL_NST;
      const float* chunk = ring.wait(st);
      const int t0 = k * TB, nst = min(TB, Tn - t0);
      f32x2 partP[NP > 0 ? NP : 1];
      float partS = 0.f;
      LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) partP[p2] = f32x2{0.f, 0.f};
#pragma unroll 1
      for (int q = 0; q < nst; ++q) {
        const int t = t0 + q;
        const float* r = chunk + q * DM::REC;
        f32x2 x1P[NP > 0 ? NP : 1][D];
        float x1S[D];
        {   // prefetch x_{t+1+PF}, then make sure x_{t+1} has landed
          const float* xrow = x_tm + (size_t)min(t + 1 + TRIAL_PF, Tn) * N * D;
          float* slot = pf + ((t + TRIAL_PF) % NSLOT) * (RT * 32 * D);
          LQGK_UNROLL for (int j = 0; j < RT; ++j) prefetch_obs<D>(slot + j * 32 * D, lane, xrow + trD[j]);
          cp_async_commit();
          cp_async_wait<TRIAL_PF>();
          const float* cur = pf + (t % NSLOT) * (RT * 32 * D);
          LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) {
            float xa[D], xb[D];
            read_obs<D>(cur + (2 * p2) * 32 * D, lane, xa);
            read_obs<D>(cur + (2 * p2 + 1) * 32 * D, lane, xb);
            LQGK_UNROLL for (int m = 0; m < D; ++m) x1P[p2][m] = f32x2{xa[m], xb[m]};
          }
          if constexpr (NS) read_obs<D>(cur + (RT - 1) * 32 * D, lane, x1S);
        }
        if (hist != nullptr) {
          if (ckc == 0) {            // t is a segment start: store the checkpoint c_t
            ckc = CK;
            LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) {
              if (ok[2 * p2 + 1]) {   // pair entry in one 64-bit store, straight from the f32x2 register pair
                LQGK_UNROLL for (int m = 0; m < R; ++m)
                  __stcs(reinterpret_cast<float2*>(hrow(m) + 64 * p2), make_float2(cP[p2][m].x, cP[p2][m].y));   // streaming: written once, read once
              } else if (ok[2 * p2]) {   // odd N: the pad element gets 0
                LQGK_UNROLL for (int m = 0; m < R; ++m) __stcs(reinterpret_cast<float2*>(hrow(m) + 64 * p2), make_float2(cP[p2][m].x, 0.f));
              }
            }
            if constexpr (NS) {
              if (ok[RT - 1]) { LQGK_UNROLL for (int m = 0; m < R; ++m) __stcs(hrow(m) + hsingle, cS[m]); }
            }
            LQGK_UNROLL for (int m = 0; m < NHP; ++m) hw[m] += hstep;
          }
          --ckc;
        }
        auto run = [&](const auto& rr) {
          LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2)
            partP[p2] = Ops<f32x2>::add(partP[p2], Trial<DM>::template fwd<f32x2>(rr, x0P[p2], x1P[p2], cP[p2]));
          if constexpr (NS) partS += Trial<DM>::template fwd<float>(rr, x0S, x1S, cS);
        };
        if constexpr (rec_in_regs<DM>()) {
          RecRegs<DM> rr;
          rr.load(r);
          run(rr);
        } else {
          run(r);
        }
        LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) LQGK_UNROLL for (int m = 0; m < D; ++m) x0P[p2][m] = x1P[p2][m];
        if constexpr (NS) { LQGK_UNROLL for (int m = 0; m < D; ++m) x0S[m] = x1S[m]; }
      }
      // FP32 partial sum over <= TB steps, FP64 across chunks (log-likelihood error ~1e-7 relative)
      LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) {
        ll[2 * p2] += (double)partP[p2].x;
        ll[2 * p2 + 1] += (double)partP[p2].y;
      }
      if constexpr (NS) ll[RT - 1] += (double)partS;
      __syncwarp();
      if (k + TRIAL_NST < nchunk) ring.issue(k + TRIAL_NST, st);
    }
    cp_async_wait<0>();
    LQGK_UNROLL for (int j = 0; j < RT; ++j)
      if (ok[j]) ll_ws[(size_t)s * N + tr[j]] = ll[j];
  }
}

// Transposing butterfly over V = 2^k <= 32 values: every lane holds V partial values; afterwards the lanes whose top
// log2(V) lane-index bits equal i hold the warp-wide total of value i (V == 32: lane L holds value L).
template <int V>
__device__ __forceinline__ float warp_transpose_reduce(float (&val)[V], int lane) {
  int bit = 16;
  LQGK_UNROLL for (int h = V / 2; h >= 1; h >>= 1) {
    const bool up = (lane & bit) != 0;
    LQGK_UNROLL for (int j = 0; j < h; ++j) {
      float send = up ? val[j] : val[j + h];
      float keep = up ? val[j + h] : val[j];
      val[j] = keep + __shfl_xor_sync(FULL, send, bit);
    }
    bit >>= 1;
  }
  float tot = val[0];
  LQGK_UNROLL for (int m = 16 / V; m >= 1; m >>= 1) tot += __shfl_xor_sync(FULL, tot, m);
  return tot;
}
__host__ __device__ constexpr int pow2_ceil(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 8 ? 8 : (v <= 16 ? 16 : 32)))); }

// Reverse: per-trial adjoint, segment by segment from the last one: (1) the checkpoint c_{k CK} arrives in the segment
// buffer by per-lane cp.async (issued one segment ahead into the slot the previous segment freed first), (2) the state
// recursion is re-run forward through the segment, c_t -> buffer slot t - k CK, (3) the adjoint walks the segment backwards
// reading c_t from the buffer, and writes the per-step sums over trials (DM::SUM_* layout) to sums[s][t][SUMP].  Every buffer
// entry is private to the lane that owns the trial, so no warp synchronisation is needed around it.  Passes over trial blocks
// accumulate (+=) into the sums.
template <class DM, int RT>
__global__ void __launch_bounds__(32 * TRIAL_WARPS) k_trial_rev(const float* __restrict__ rec, const float* __restrict__ x_all,
                                                                size_t x_sample_stride, int s_first, const float* __restrict__ hist,
                                                                const float* __restrict__ w, int n_samples, int N, int Tn,
                                                                float* __restrict__ sums) {
  constexpr int D = DM::D, R = DM::R;
  constexpr int CK = trial_ck<DM>(RT), ROW = 32 * RT, SLOT = R * ROW;
  constexpr int NFULL = DM::NSUM / 32, REM = DM::NSUM % 32, VREM = pow2_ceil(REM);
  constexpr int NP = RT / 2, NS = RT % 2, NPA = NP > 0 ? NP : 1;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * TRIAL_WARPS + warp;
  if (s >= n_samples) return;
  const float* __restrict__ x_tm = x_all + (size_t)(s_first + s) * x_sample_stride;   // this sample's observations
  constexpr size_t RING_FLOATS = (size_t)2 * CK * DM::REC;
  float* ring_base = reinterpret_cast<float*>(smraw);
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(ring_base + TRIAL_WARPS * RING_FLOATS);
  float* red_base = reinterpret_cast<float*>(bar_base + TRIAL_WARPS * 2);                     // 16-byte aligned
  float* red = red_base + (size_t)warp * (trial_red_bytes<DM>() / sizeof(float));
  float* seg = red_base + (size_t)TRIAL_WARPS * (trial_red_bytes<DM>() / sizeof(float)) + (size_t)warp * CK * SLOT;
  RecRing<DM, CK, 2> ring{ring_base + (size_t)warp * RING_FLOATS, bar_base + warp * 2, rec + (size_t)s * Tn * DM::REC, Tn, lane, 0};
  ring.init();
  const int nseg = (Tn + CK - 1) / CK;
  const int Nh = hist_stride(N);
  for (int base = 0; base < N; base += 32 * RT) {
    int trD[RT];
    float wt[RT];
    bool okP[NPA], okS = false;
    LQGK_UNROLL for (int j = 0; j < RT; ++j) {
      int i = trial_of<RT>(base, lane, j);
      bool ok = i < N;
      int tr = ok ? i : N - 1;
      trD[j] = tr * D;                            // slots beyond the last trial re-read trial N-1 (finite; masked by w = 0, cb = 0)
      wt[j] = ok ? w[(size_t)s * N + tr] : 0.f;
      if (j < 2 * NP) { if ((j & 1) == 0) okP[j >> 1] = ok; } else okS = ok;
    }
    auto load_x = [&](int t, float (&x)[RT][D]) {
      const float* xrow = x_tm + (size_t)t * N * D;
      LQGK_UNROLL for (int j = 0; j < RT; ++j) load_obs<D>(xrow + trD[j], x[j]);
    };
    // this lane's entries of one buffer slot / one checkpoint: pair p at [m][64 p + 2 lane], the single at [m][64 NP + lane]
    const int lofs = 2 * lane, sofs = 64 * NP + lane;
    auto fetch_ckpt = [&](float* slot, const float* hk) {   // hk = hist + ((s * nseg + k) * R) * Nh + base
      LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) LQGK_UNROLL for (int m = 0; m < R; ++m) {
        float* dst = slot + m * ROW + 64 * p2 + lofs;
        if (okP[p2]) cp_async<8>(dst, hk + (size_t)m * Nh + 64 * p2 + lofs);
        else *reinterpret_cast<float2*>(dst) = make_float2(0.f, 0.f);
      }
      if constexpr (NS) {
        LQGK_UNROLL for (int m = 0; m < R; ++m) {
          float* dst = slot + m * ROW + sofs;
          if (okS) cp_async<4>(dst, hk + (size_t)m * Nh + sofs);
          else *dst = 0.f;
        }
      }
      cp_async_commit();
    };
    auto load_slot = [&](const float* slot, f32x2 (&cp)[NPA][R], float (&cs)[R]) {
      LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) LQGK_UNROLL for (int m = 0; m < R; ++m) {
        float2 v = *reinterpret_cast<const float2*>(slot + m * ROW + 64 * p2 + lofs);
        cp[p2][m] = f32x2{v.x, v.y};
      }
      if constexpr (NS) { LQGK_UNROLL for (int m = 0; m < R; ++m) cs[m] = slot[m * ROW + sofs]; }
    };
    auto store_slot = [&](float* slot, const f32x2 (&cp)[NPA][R], const float (&cs)[R]) {
      LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) LQGK_UNROLL for (int m = 0; m < R; ++m)
        *reinterpret_cast<float2*>(slot + m * ROW + 64 * p2 + lofs) = make_float2(cp[p2][m].x, cp[p2][m].y);
      if constexpr (NS) { LQGK_UNROLL for (int m = 0; m < R; ++m) slot[m * ROW + sofs] = cs[m]; }
    };
    auto next_slot = [&](float* p) { p += SLOT; return p == seg + CK * SLOT ? seg : p; };
    auto prev_slot = [&](float* p) { return (p == seg ? seg + CK * SLOT : p) - SLOT; };
    // persistent state: cotangent cb of c_{t+1}, x_{t+1}, weights
    f32x2 cbP[NPA][R], x1P[NPA][D], wP[NPA];
    float cbS[R], x1S[D], wS = wt[RT - 1];
    {
      float x1[RT][D];
      load_x(Tn, x1);
      LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) {
        LQGK_UNROLL for (int k = 0; k < R; ++k) cbP[p2][k] = f32x2{0.f, 0.f};
        LQGK_UNROLL for (int m = 0; m < D; ++m) x1P[p2][m] = f32x2{x1[2 * p2][m], x1[2 * p2 + 1][m]};
        wP[p2] = f32x2{wt[2 * p2], wt[2 * p2 + 1]};
      }
      LQGK_UNROLL for (int k = 0; k < R; ++k) cbS[k] = 0.f;
      LQGK_UNROLL for (int m = 0; m < D; ++m) x1S[m] = x1[RT - 1][m];
    }
    const float* hk = hist + ((size_t)s * nseg + (nseg - 1)) * R * Nh + base;
    float* slot0 = seg;                            // buffer slot of the current segment's first step
    fetch_ckpt(slot0, hk);
    ring.issue(nseg - 1, 0);
    if (nseg > 1) ring.issue(nseg - 2, 1);
    float* out = sums + ((size_t)s * Tn + (Tn - 1)) * DM::SUMP;
    for (int kk = 0; kk < nseg; ++kk) {
      const int k = nseg - 1 - kk, st = kk & 1;
      const int t0 = k * CK, nst = min(CK, Tn - t0);
      const float* chunk = ring.wait(st);
      cp_async_wait<0>();                          // checkpoint k is in slot0 (lane-private data: no warp sync needed)
      // ---- (2) re-run the state recursion through the segment.  Fully unrolled (CK - 1 steps) so that the rotating buffer of
      // observation rows xs[row % 4] has static indices: row j + 3 is requested while step j runs (the L2 round trip is
      // longer than one of these short steps).
      float x0c[RT][D];                            // x_t of the step the reverse walk starts with (t0 + nst - 1)
      float* sp = slot0;
      {
        f32x2 cP[NPA][R];
        float cS[R];
        float xs[4][RT][D];
        load_slot(sp, cP, cS);
        load_x(t0, xs[0]);
        if (nst > 1) load_x(t0 + 1, xs[1]);
        if (nst > 2) load_x(t0 + 2, xs[2]);
        static_for<0, CK - 1>([&](auto QC) {
          constexpr int q = decltype(QC)::value;
          if (q + 1 < nst) {
            if (q + 3 < nst) load_x(t0 + q + 3, xs[(q + 3) % 4]);
            f32x2 xaP[NPA][D], xbP[NPA][D];
            float xaS[D], xbS[D];
            LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) LQGK_UNROLL for (int m = 0; m < D; ++m) {
              xaP[p2][m] = f32x2{xs[q % 4][2 * p2][m], xs[q % 4][2 * p2 + 1][m]};
              xbP[p2][m] = f32x2{xs[(q + 1) % 4][2 * p2][m], xs[(q + 1) % 4][2 * p2 + 1][m]};
            }
            LQGK_UNROLL for (int m = 0; m < D; ++m) { xaS[m] = xs[q % 4][RT - 1][m]; xbS[m] = xs[(q + 1) % 4][RT - 1][m]; }
            const float* r = chunk + q * DM::REC;
            auto run = [&](const auto& rr) {
              LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) Trial<DM>::template advance<f32x2>(rr, xaP[p2], xbP[p2], cP[p2]);
              if constexpr (NS) Trial<DM>::template advance<float>(rr, xaS, xbS, cS);
            };
            if constexpr (rec_in_regs<DM>()) {
              RecRegs<DM> rr;
              rr.load(r);
              run(rr);
            } else {
              run(r);
            }
            sp = next_slot(sp);
            store_slot(sp, cP, cS);
          }
        });
        static_for<0, CK>([&](auto JC) {             // x_{t0 + nst - 1}: static for full segments, selected for the ragged last one
          constexpr int j = decltype(JC)::value;
          if (nst - 1 == j) { LQGK_UNROLL for (int jj = 0; jj < RT; ++jj) LQGK_UNROLL for (int m = 0; m < D; ++m) x0c[jj][m] = xs[j % 4][jj][m]; }
        });
      }
      // ---- (3) walk the segment backwards; sp = slot of its last step, x0c = x_{t0 + nst - 1}
#pragma unroll 1
      for (int q = nst - 1; q >= 0; --q) {
        const int t = t0 + q;
        const float* r = chunk + q * DM::REC;
        f32x2 x0P[NPA][D], cP[NPA][R];
        float x0S[D], cS[R];
        load_slot(sp, cP, cS);
        LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) LQGK_UNROLL for (int m = 0; m < D; ++m) x0P[p2][m] = f32x2{x0c[2 * p2][m], x0c[2 * p2 + 1][m]};
        LQGK_UNROLL for (int m = 0; m < D; ++m) x0S[m] = x0c[RT - 1][m];
        if (q > 0) load_x(t - 1, x0c);              // next reverse step's x_t, one step ahead of its use
        f32x2 eP[NPA][D], vP[NPA][D], wvP[NPA][D], nebP[NPA][D], cbnP[NPA][R];
        float eS[D], vS[D], wvS[D], nebS[D], cbnS[R];
        auto run = [&](const auto& rr) {
          LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2)
            Trial<DM>::template rev<f32x2>(rr, x0P[p2], x1P[p2], cP[p2], wP[p2], cbP[p2], eP[p2], vP[p2], wvP[p2], nebP[p2], cbnP[p2]);
          if constexpr (NS) Trial<DM>::template rev<float>(rr, x0S, x1S, cS, wS, cbS, eS, vS, wvS, nebS, cbnS);
        };
        if constexpr (rec_in_regs<DM>()) {
          RecRegs<DM> rr;
          rr.load(r);
          run(rr);
        } else {
          run(r);
        }
        if (q == nst - 1 && k > 0) {
          // the slot of the segment's last step has been consumed (its values are operands of the arithmetic above):
          // the checkpoint of the next segment to process lands there while this one is walked
          hk -= (size_t)R * Nh;
          fetch_ckpt(sp, hk);
          slot0 = sp;
        }
        auto term = [&](auto IDXC) -> float {
          constexpr int IDX = decltype(IDXC)::value;
          float a = 0.f;
          if constexpr (NP > 0) {
            f32x2 acc = f32x2{0.f, 0.f};
            LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2)
              acc = Trial<DM>::template sum_acc<IDX, f32x2>(acc, cbP[p2], nebP[p2], x0P[p2], cP[p2], eP[p2], vP[p2], wvP[p2]);
            a = acc.x + acc.y;
          }
          if constexpr (NS) a = Trial<DM>::template sum_acc<IDX, float>(a, cbS, nebS, x0S, cS, eS, vS, wvS);
          return a;
        };
        static_for<0, NFULL>([&](auto G) {
          float val[32];
          static_for<0, 32>([&](auto J) {
            val[decltype(J)::value] = term(std::integral_constant<int, decltype(G)::value * 32 + decltype(J)::value>{});
          });
          float4* dst = reinterpret_cast<float4*>(red + lane * TRIAL_RED_STRIDE);
          LQGK_UNROLL for (int q4 = 0; q4 < 8; ++q4) dst[q4] = make_float4(val[4 * q4], val[4 * q4 + 1], val[4 * q4 + 2], val[4 * q4 + 3]);
          __syncwarp();
          float part[4] = {0.f, 0.f, 0.f, 0.f};
          LQGK_UNROLL for (int k2 = 0; k2 < 32; ++k2) part[k2 & 3] += red[k2 * TRIAL_RED_STRIDE + lane];
          const float tot = (part[0] + part[1]) + (part[2] + part[3]);
          __syncwarp();
          const int idx = decltype(G)::value * 32 + lane;
          if (base == 0) out[idx] = tot;
          else out[idx] += tot;
        });
        if constexpr (REM > 0) {
          float val[VREM];
          static_for<0, VREM>([&](auto J) { val[decltype(J)::value] = term(std::integral_constant<int, NFULL * 32 + decltype(J)::value>{}); });
          float tot = warp_transpose_reduce<VREM>(val, lane);
          const int vi = lane / (32 / VREM);
          const int idx = NFULL * 32 + vi;
          if ((lane % (32 / VREM)) == 0 && idx < DM::SUMP) {
            if (base == 0) out[idx] = tot;
            else out[idx] += tot;
          }
        }
        LQGK_UNROLL for (int p2 = 0; p2 < NP; ++p2) {
          LQGK_UNROLL for (int m = 0; m < R; ++m) cbP[p2][m] = cbnP[p2][m];
          LQGK_UNROLL for (int m = 0; m < D; ++m) x1P[p2][m] = x0P[p2][m];
        }
        if constexpr (NS) {
          LQGK_UNROLL for (int m = 0; m < R; ++m) cbS[m] = cbnS[m];
          LQGK_UNROLL for (int m = 0; m < D; ++m) x1S[m] = x0S[m];
        }
        out -= DM::SUMP;
        sp = prev_slot(sp);
      }
      __syncwarp();
      if (kk + 2 < nseg) ring.issue(nseg - 1 - (kk + 2), st);
    }
  }
}

}  // namespace lqgk
