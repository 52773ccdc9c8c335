// nqmc_xla_ffi.cu -- XLA FFI custom-call handlers over the C ABI (BASELINE north_star: "a thin C-ABI exposed as
// JAX FFI custom calls").  Compiled into libnqmc_b200.so always; the handlers themselves exist only where the
// jaxlib headers are on the include path (`-I $(python -c "import jaxlib, os; print(os.path.join(
// os.path.dirname(jaxlib.__file__), 'include'))")`).  This image has no jaxlib, so here the translation unit
// reduces to nqmc_xla_ffi_available() == 0 and the build still links; with jaxlib present the four handlers below are
// what `jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.<Symbol>), platform="CUDA")` binds.
//
// Conventions: the nqmc_ctx* travels as an int64 attribute ("ctx", from ctypes); positions cross in the library's SoA
// [3N][W] layout; in/out operands (r, logp) are copied to the result buffers first, so the handlers are functional as
// XLA requires; every call is asynchronous on XLA's stream.
#include "nqmc_internal.cuh"

#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define NQMC_HAVE_XLA_FFI 1
#endif
#endif

#ifdef NQMC_HAVE_XLA_FFI
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;

namespace {

inline nqmc_ctx* ctx_of(int64_t h) { return reinterpret_cast<nqmc_ctx*>(static_cast<intptr_t>(h)); }
inline ffi::Error status(nqmc_ctx* ctx, int rc) {
  return rc == NQMC_OK ? ffi::Error::Success() : ffi::Error::Internal(nqmc_last_error(ctx));
}

// log|psi|, sign  <-  wavefunction(params, r) vmapped  (src/nqmc/wavefunction/base.py:38-48)
ffi::Error LogPsiImpl(cudaStream_t stream, int64_t h, ffi::Buffer<ffi::F32> r, ffi::ResultBuffer<ffi::F32> logabs,
                      ffi::ResultBuffer<ffi::F32> sign) {
  nqmc_ctx* ctx = ctx_of(h);
  const int64_t W = (int64_t)logabs->element_count();
  return status(ctx, nqmc_log_psi(ctx, r.typed_data(), W, logabs->typed_data(), sign->typed_data(), stream));
}

// MetropolisSampler.step  (src/nqmc/sampler/metropolis.py:99-147); normals / uniforms are the jax.random draws
ffi::Error MhStepImpl(cudaStream_t stream, int64_t h, float sigma, ffi::Buffer<ffi::F32> r, ffi::Buffer<ffi::F32> logp,
                      ffi::Buffer<ffi::F32> normals, ffi::Buffer<ffi::F32> uniforms, ffi::ResultBuffer<ffi::F32> r_out,
                      ffi::ResultBuffer<ffi::F32> logp_out, ffi::ResultBuffer<ffi::U8> accept) {
  nqmc_ctx* ctx = ctx_of(h);
  const int64_t W = (int64_t)logp.element_count();
  cudaMemcpyAsync(r_out->typed_data(), r.typed_data(), r.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(logp_out->typed_data(), logp.typed_data(), logp.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  return status(ctx, nqmc_mh_step(ctx, r_out->typed_data(), logp_out->typed_data(), normals.typed_data(),
                                  uniforms.typed_data(), 0, 0, 0, sigma, W, accept->typed_data(), nullptr, stream));
}

// local_energy_batched  (src/nqmc/hamiltonian/molecular.py:178-197)
ffi::Error LocalEnergyImpl(cudaStream_t stream, int64_t h, ffi::Buffer<ffi::F32> r, ffi::ResultBuffer<ffi::F32> e_l) {
  nqmc_ctx* ctx = ctx_of(h);
  const int64_t W = (int64_t)e_l->element_count();
  return status(ctx, nqmc_local_energy(ctx, r.typed_data(), W, e_l->typed_data(), nullptr, stream));
}

// one sample of VMCOptimizer.step: MH move + E_L + header + centred gradient sums  (src/nqmc/optimizer/vmc.py:242-253)
ffi::Error WalkerStepImpl(cudaStream_t stream, int64_t h, uint64_t seed, uint64_t step, int64_t walker_id0, float sigma,
                          ffi::Buffer<ffi::F32> r, ffi::Buffer<ffi::F32> logp, ffi::ResultBuffer<ffi::F32> r_out,
                          ffi::ResultBuffer<ffi::F32> logp_out, ffi::ResultBuffer<ffi::F32> e_l,
                          ffi::ResultBuffer<ffi::F64> hdr, ffi::ResultBuffer<ffi::F64> gsum) {
  nqmc_ctx* ctx = ctx_of(h);
  const int64_t W = (int64_t)logp.element_count();
  cudaMemcpyAsync(r_out->typed_data(), r.typed_data(), r.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(logp_out->typed_data(), logp.typed_data(), logp.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  return status(ctx, nqmc_walker_step(ctx, r_out->typed_data(), logp_out->typed_data(), nullptr, nullptr, seed, step,
                                      walker_id0, sigma, W, e_l->typed_data(), nullptr, hdr->typed_data(),
                                      gsum->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(NqmcLogPsi, LogPsiImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("ctx")
                                  .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(NqmcMhStep, MhStepImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("ctx").Attr<float>("sigma")
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(NqmcLocalEnergy, LocalEnergyImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("ctx")
                                  .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(NqmcWalkerStep, WalkerStepImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("ctx")
                                  .Attr<uint64_t>("seed").Attr<uint64_t>("step").Attr<int64_t>("walker_id0").Attr<float>("sigma")
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
#endif  // NQMC_HAVE_XLA_FFI

extern "C" int nqmc_xla_ffi_available(void) {
#ifdef NQMC_HAVE_XLA_FFI
  return 1;
#else
  return 0;
#endif
}
