// XLA FFI shim: exposes the C ABI of libb200lanczos.so as `jax.ffi` custom-call targets.
//
// NOT built in this image (JAX / jaxlib and their headers are not installed):
// compile where `python -c "import jax.ffi; print(jax.ffi.include_dir())"` works, e.g.
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi;print(jax.ffi.include_dir())") \
//       -I../../include ffi_shim.cc -L.. -lb200lanczos -o libb200lanczos_ffi.so
// The shim holds no logic: it unpacks XLA buffers into the raw pointers the C ABI takes and
// forwards the CUDA stream XLA runs the custom call on.  Operator handles (bl_operator_t*) are
// created once on the Python side through ctypes and travel as int64 attributes.
//
// Where each target plugs into the reference: INTEGRATION.md.
#include <cstdint>

#include "b200_lanczos.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

int dtype_of(ffi::DataType t) { return t == ffi::DataType::F32 ? BL_F32 : BL_F64; }

ffi::Error to_error(int rc) {
  if (rc == BL_OK) return ffi::Error::Success();
  return ffi::Error(rc == BL_EDEPTH || rc == BL_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    bl_last_error());
}

// arnoldi._forward (arnoldi.py:57-101): (v, params) -> (Qt (K, ld), H (K, K), r (n), c (), workspace)
ffi::Error ArnoldiForward(cudaStream_t stream, ffi::AnyBuffer v, ffi::AnyBuffer params, int64_t op_handle,
                          int64_t krylov_depth, int64_t second_pass, ffi::Result<ffi::AnyBuffer> Qt,
                          ffi::Result<ffi::AnyBuffer> H, ffi::Result<ffi::AnyBuffer> r,
                          ffi::Result<ffi::AnyBuffer> c, ffi::Result<ffi::AnyBuffer> workspace) {
  auto* op = reinterpret_cast<bl_operator_t*>(op_handle);
  const int dtype = dtype_of(v.element_type());
  const int64_t n = v.dimensions()[0];
  const int64_t ld = Qt->dimensions()[1];
  const void* p[1] = {params.untyped_data()};
  int rc = bl_op_set_params(op, dtype, p, 1, stream);
  if (rc != BL_OK) return to_error(rc);
  return to_error(bl_arnoldi_forward(op, dtype, n, krylov_depth, (int)second_pass, v.untyped_data(),
                                     Qt->untyped_data(), ld, H->untyped_data(), r->untyped_data(),
                                     c->untyped_data(), workspace->untyped_data(), workspace->size_bytes(), stream));
}

// arnoldi._adjoint (arnoldi.py:104-220): residuals + cotangents -> (dv, dparams, Lambda scratch, workspace)
ffi::Error ArnoldiAdjoint(cudaStream_t stream, ffi::AnyBuffer params, ffi::AnyBuffer Qt, ffi::AnyBuffer H,
                          ffi::AnyBuffer r, ffi::AnyBuffer c, ffi::AnyBuffer dQt, ffi::AnyBuffer dH,
                          ffi::AnyBuffer dr, ffi::AnyBuffer dc, int64_t op_handle, int64_t reortho_full,
                          int64_t dense_cotangents, ffi::Result<ffi::AnyBuffer> dv,
                          ffi::Result<ffi::AnyBuffer> dparams, ffi::Result<ffi::AnyBuffer> Lambda,
                          ffi::Result<ffi::AnyBuffer> workspace) {
  auto* op = reinterpret_cast<bl_operator_t*>(op_handle);
  const int dtype = dtype_of(r.element_type());
  const int64_t n = r.dimensions()[0], K = H.dimensions()[0], ld = Qt.dimensions()[1];
  const void* p[1] = {params.untyped_data()};
  int rc = bl_op_set_params(op, dtype, p, 1, stream);
  if (rc == BL_OK) rc = bl_op_grad_zero(op, dtype, stream);
  if (rc != BL_OK) return to_error(rc);
  // JAX materialises zero cotangents for a custom_vjp; the caller says whether they are dense
  const void* dQp = dense_cotangents ? dQt.untyped_data() : nullptr;
  const void* drp = dense_cotangents ? dr.untyped_data() : nullptr;
  const void* dcp = dense_cotangents ? dc.untyped_data() : nullptr;
  rc = bl_arnoldi_adjoint(op, dtype, n, K, (int)reortho_full, Qt.untyped_data(), ld, H.untyped_data(),
                          r.untyped_data(), c.untyped_data(), dQp, dH.untyped_data(), drp, dcp, dv->untyped_data(),
                          Lambda->untyped_data(), workspace->untyped_data(), workspace->size_bytes(), stream);
  if (rc != BL_OK) return to_error(rc);
  void* g[1] = {dparams->untyped_data()};
  return to_error(bl_op_grad_export(op, dtype, g, 1, stream));
}

// the user matvec itself, for callers that only want the operator (benchmark.py:64-68)
ffi::Error Matvec(cudaStream_t stream, ffi::AnyBuffer x, ffi::AnyBuffer params, int64_t op_handle,
                  ffi::Result<ffi::AnyBuffer> y) {
  auto* op = reinterpret_cast<bl_operator_t*>(op_handle);
  const int dtype = dtype_of(x.element_type());
  const void* p[1] = {params.untyped_data()};
  int rc = bl_op_set_params(op, dtype, p, 1, stream);
  if (rc != BL_OK) return to_error(rc);
  return to_error(bl_op_matvec(op, dtype, x.untyped_data(), y->untyped_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(bl_ffi_arnoldi_forward, ArnoldiForward,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Attr<int64_t>("op_handle")
                                  .Attr<int64_t>("krylov_depth")
                                  .Attr<int64_t>("second_pass")
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(bl_ffi_arnoldi_adjoint, ArnoldiAdjoint,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Attr<int64_t>("op_handle")
                                  .Attr<int64_t>("reortho_full")
                                  .Attr<int64_t>("dense_cotangents")
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(bl_ffi_matvec, Matvec,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Attr<int64_t>("op_handle")
                                  .Ret<ffi::AnyBuffer>());
