// lqgk_xla_ffi.cc -- XLA FFI custom-call handlers over the C ABI of liblqgk.so (include/lqgk.h).
//
// UNTESTED IN THIS ENVIRONMENT: the image has no jax / jaxlib (hence no xla/ffi/api/ffi.h) and no network, so this
// file is not part of __graft_entry__.build().  It is the binding a maintainer adds where jaxlib is installed:
//
//   g++ -std=c++17 -O2 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include lqgk_xla_ffi.cc -o liblqgk_xla.so -L../csrc -llqgk -Wl,-rpath,'$ORIGIN/../csrc'
//
// Pinned versions of the reference's lock file: jax / jaxlib 0.10.0, numpyro 0.21.0 (uv.lock).
// Operand convention (time-invariant models, what lqg.tracking builds): 12 base matrices with a leading sample axis
// [S, r, c] (f32), observations already time-major x_tm[T+1, N, d] (f32), a scratch workspace buffer.
#include <cuda_runtime_api.h>

#include <cstdint>

#include "../../include/lqgk.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;
using U8 = ffi::Buffer<ffi::U8>;

LqgkMat mat(const F32& b) {
  auto d = b.dimensions();                       // [S, r, c]; S == 1 means shared across samples
  int64_t per = 1;
  for (size_t i = 1; i < d.size(); ++i) per *= d[i];
  return LqgkMat{b.typed_data(), d[0] > 1 ? per : 0, 0};
}
LqgkMatGrad gmat(ffi::Result<F32>& b) {
  auto d = b->dimensions();
  int64_t per = 1;
  for (size_t i = 1; i < d.size(); ++i) per *= d[i];
  return LqgkMatGrad{b->typed_data(), per};
}
ffi::Error status(int rc, const char* what) {
  if (rc == LQGK_OK) return ffi::Error::Success();
  if (rc == LQGK_E_INVALID || rc == LQGK_E_UNSUPPORTED) return ffi::Error::InvalidArgument(std::string(what) + ": " + lqgk_strerror(rc));
  return ffi::Error::Internal(std::string(what) + ": " + lqgk_strerror(rc));
}
LqgkDims dims_of(const F32& Aa, const F32& Ad, const F32& Ba, const F32& Fa, const F32& x_tm, int64_t S) {
  LqgkDims d;
  d.S = (int32_t)S;
  d.T = (int32_t)x_tm.dimensions()[0] - 1;
  d.N = (int32_t)x_tm.dimensions()[1];
  d.d = (int32_t)x_tm.dimensions()[2];
  d.x = (int32_t)Ad.dimensions()[1];
  d.b = (int32_t)Aa.dimensions()[1];
  d.u = (int32_t)Ba.dimensions()[2];
  d.y = (int32_t)Fa.dimensions()[1];
  d.x_sample_stride = 0;   // one shared data set (parameter sweeps / vmap over parameters)
  return d;
}

// ll[S, N] = System.log_likelihood(x)            (lqg/system.py:246-248)
ffi::Error LoglikFwdImpl(cudaStream_t stream, F32 Aa, F32 Ba, F32 Fa, F32 Va, F32 Wa, F32 Q, F32 R, F32 Ad, F32 Bd, F32 Fd,
                         F32 Vd, F32 Wd, F32 x_tm, ffi::Result<F32> ll, ffi::Result<U8> ws) {
  LqgkSpec act{}, dyn{};
  act.A = mat(Aa); act.B = mat(Ba); act.F = mat(Fa); act.V = mat(Va); act.W = mat(Wa); act.Q = mat(Q); act.R = mat(R);
  dyn.A = mat(Ad); dyn.B = mat(Bd); dyn.F = mat(Fd); dyn.V = mat(Vd); dyn.W = mat(Wd);
  LqgkDims d = dims_of(Aa, Ad, Ba, Fa, x_tm, ll->dimensions()[0]);
  return status(lqgk_loglik_fwd_f32(&d, &act, &dyn, nullptr, x_tm.typed_data(), ll->typed_data(), ws->untyped_data(),
                                    ws->size_bytes(), stream), "lqgk_loglik_fwd_f32");
}

// (ll, 12 cotangents) = fused forward + adjoint with cotangent ll_bar[S, N]  (jax.value_and_grad in the reference)
ffi::Error LoglikVjpImpl(cudaStream_t stream, F32 Aa, F32 Ba, F32 Fa, F32 Va, F32 Wa, F32 Q, F32 R, F32 Ad, F32 Bd, F32 Fd,
                         F32 Vd, F32 Wd, F32 x_tm, F32 ll_bar, ffi::Result<F32> ll, ffi::Result<F32> gAa, ffi::Result<F32> gBa,
                         ffi::Result<F32> gFa, ffi::Result<F32> gVa, ffi::Result<F32> gWa, ffi::Result<F32> gQ,
                         ffi::Result<F32> gR, ffi::Result<F32> gAd, ffi::Result<F32> gBd, ffi::Result<F32> gFd,
                         ffi::Result<F32> gVd, ffi::Result<F32> gWd, ffi::Result<U8> ws) {
  LqgkSpec act{}, dyn{};
  act.A = mat(Aa); act.B = mat(Ba); act.F = mat(Fa); act.V = mat(Va); act.W = mat(Wa); act.Q = mat(Q); act.R = mat(R);
  dyn.A = mat(Ad); dyn.B = mat(Bd); dyn.F = mat(Fd); dyn.V = mat(Vd); dyn.W = mat(Wd);
  LqgkSpecGrad ga{}, gd{};
  ga.A = gmat(gAa); ga.B = gmat(gBa); ga.F = gmat(gFa); ga.V = gmat(gVa); ga.W = gmat(gWa); ga.Q = gmat(gQ); ga.R = gmat(gR);
  gd.A = gmat(gAd); gd.B = gmat(gBd); gd.F = gmat(gFd); gd.V = gmat(gVd); gd.W = gmat(gWd);
  LqgkDims d = dims_of(Aa, Ad, Ba, Fa, x_tm, ll->dimensions()[0]);
  return status(lqgk_loglik_vjp_f32(&d, &act, &dyn, nullptr, x_tm.typed_data(), ll_bar.typed_data(), ll->typed_data(), &ga, &gd,
                                    nullptr, ws->untyped_data(), ws->size_bytes(), stream), "lqgk_loglik_vjp_f32");
}

}  // namespace

#define LQGK_12 .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
#define LQGK_12R .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(LqgLoglikFwd, LoglikFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() LQGK_12.Arg<F32>().Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(LqgLoglikVjp, LoglikVjpImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() LQGK_12.Arg<F32>().Arg<F32>().Ret<F32>() LQGK_12R.Ret<U8>());
