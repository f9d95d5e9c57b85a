// XLA-FFI adapter: the thin custom-call layer between jax.ffi and the C ABI of include/b200nuts.h.
// Header-gated: it compiles only where jaxlib's headers are present (python -c "import jax.ffi; print(jax.ffi.include_dir())"),
// which is NOT the case in this image (jax is not installable offline), so this file is argument marshalling only and is
// kept under 100 lines.  Build, where jax exists:
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi;print(jax.ffi.include_dir())") -I include \
//       ffi/b200nuts_ffi.cc -L numpyro_b200/csrc -lb200nuts -o ffi/libb200nuts_ffi.so
// Python side: jax.ffi.register_ffi_target("b200nuts_transition",
//       jax.ffi.pycapsule(lib.B200NutsTransition), platform="CUDA") and jax.ffi.ffi_call(...).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define B200NUTS_HAVE_XLA_FFI 1
#endif
#endif

#ifdef B200NUTS_HAVE_XLA_FFI
#include <cuda_runtime.h>
#include "xla/ffi/api/ffi.h"
#include "b200nuts.h"

namespace ffi = xla::ffi;

static ffi::Error fail(B200Nuts* h, int rc) {
    return ffi::Error(rc == B200NUTS_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, b200nuts_last_error(h));
}

// HMCState -> HMCState for every chain of the handle (numpyro/infer/hmc.py:459-530 sample_kernel, vmapped :790-798).
// The chain state is device resident inside the handle (created host-side through ctypes by numpyro_b200.infer.NUTS.init);
// `token` orders successive calls, the outputs are the HMCState fields MCMC collects (z, diverging, ...).
static ffi::Error TransitionImpl(cudaStream_t stream, int64_t handle, int32_t n_iter, ffi::Buffer<ffi::S32> token,
                                 ffi::ResultBuffer<ffi::F32> z, ffi::ResultBuffer<ffi::F32> z_grad,
                                 ffi::ResultBuffer<ffi::F32> scalars, ffi::ResultBuffer<ffi::S32> token_out) {
    B200Nuts* h = reinterpret_cast<B200Nuts*>(static_cast<intptr_t>(handle));
    if (!h) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200nuts: null handle");
    if (int rc = b200nuts_transition(h, n_iter, stream)) return fail(h, rc);                       // enqueue only
    if (int rc = b200nuts_state_to_device(h, z->typed_data(), z_grad->typed_data(), scalars->typed_data(), stream)) return fail(h, rc);
    cudaMemcpyAsync(token_out->typed_data(), token.typed_data(), sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
    return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200NutsTransition, TransitionImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int32_t>("n_iter")
                                  .Arg<ffi::Buffer<ffi::S32>>()       // token
                                  .Ret<ffi::Buffer<ffi::F32>>()       // z            [C, D]
                                  .Ret<ffi::Buffer<ffi::F32>>()       // z_grad       [C, D]
                                  .Ret<ffi::Buffer<ffi::F32>>()       // scalars      [C, 8] (B200NUTS_STATE_SCALARS)
                                  .Ret<ffi::Buffer<ffi::S32>>());     // token

// jax.value_and_grad(potential_fn) for a batch of positions (hmc_util.py:242-252): the parity hook as a custom call.
static ffi::Error PotentialImpl(cudaStream_t stream, int64_t handle, ffi::Buffer<ffi::F32> z, ffi::ResultBuffer<ffi::F32> u,
                                ffi::ResultBuffer<ffi::F32> grad) {
    B200Nuts* h = reinterpret_cast<B200Nuts*>(static_cast<intptr_t>(handle));
    if (!h) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200nuts: null handle");
    if (int rc = b200nuts_potential_and_grad(h, z.typed_data(), u->typed_data(), grad->typed_data(), stream)) return fail(h, rc);
    return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200NutsPotentialAndGrad, PotentialImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Arg<ffi::Buffer<ffi::F32>>()       // z            [C, D]
                                  .Ret<ffi::Buffer<ffi::F32>>()       // U            [C]
                                  .Ret<ffi::Buffer<ffi::F32>>());     // grad         [C, D]
#else
// jaxlib headers absent: nothing to build (tests/test_capi_symbols.py checks that this translation unit stays a no-op here)
extern "C" int b200nuts_ffi_unavailable(void) { return 1; }
#endif
