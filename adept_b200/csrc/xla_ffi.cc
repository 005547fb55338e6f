// EXPERIMENTAL -- never compiled or run (no XLA FFI headers in the build image); excluded from the default build.
// Layer 2 of the drop-in boundary: XLA FFI handlers that unwrap XLA buffers + the CUDA stream and forward
// to the C ABI of include/adept_b200.h.  This is what jax.ffi.ffi_call binds inside the reference's jitted
// diffrax loop (adept/_vlasov1d/modules.py:343-356); adept_b200/jax_ffi.py registers the symbols and pairs forward and
// backward calls in jax.custom_vjp.
//
// Built only where the XLA FFI headers exist (python -m adept_b200.build --xla, which takes the include directory from
// jax.ffi.include_dir() or $XLA_FFI_INCLUDE_DIR) into libadept_b200_xla.so.  NOT COMPILED IN THIS IMAGE: jax is not
// installed here and cannot be (no network), so this file is unverified source; everything it calls is the tested
// C ABI.  Handlers never allocate, free, retain pointers or synchronise; scratch is declared as extra results.
//
// Conventions: f buffers are [nx, nv] or [batch, nx, nv] (vmap_method="broadcast_all" prepends the batch axis);
// scalars that are static in the reference's jitted step (dt, grid constants, charge, mass) are attributes; everything
// time dependent (e, dex, pond, nu) is a device buffer.
#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"

#include "../../include/adept_b200.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;

namespace {

ffi::Error check(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(rc == ADEPT_B200_ERR_CUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument,
                    adept_b200_last_error());
}

struct Shape3 {
  int batch, nx, nv;
  bool ok;
};

Shape3 shape3(const F64& f) {
  auto d = f.dimensions();
  if (d.size() == 2) return {1, (int)d[0], (int)d[1], true};
  if (d.size() == 3) return {(int)d[0], (int)d[1], (int)d[2], true};
  return {0, 0, 0, false};
}

ffi::Error bad_rank() { return ffi::Error(ffi::ErrorCode::kInvalidArgument, "adept_b200: f must be [nx, nv] or [batch, nx, nv]"); }

// SpaceExponential.push  (adept/_vlasov1d/solvers/pushers/vlasov.py:234-251)
ffi::Error VdfdxImpl(cudaStream_t stream, F64 f, F64 v, double dt, double k1x, F64Out out) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  return check(adept_b200_vdfdx_f64(f.typed_data(), out->typed_data(), s.batch, s.nx, s.nv, v.typed_data(), dt, k1x,
                                    nullptr, stream));
}

// SpaceExponential.push + velocity sum of the result (field.py:197-208); parts is scratch [nparts, batch*nx]
ffi::Error VdfdxRhoImpl(cudaStream_t stream, F64 f, F64 v, double dt, double k1x, F64Out out, F64Out parts) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  const int nparts = (int)parts->dimensions()[0];
  cudaError_t err = cudaMemsetAsync(parts->typed_data(), 0, parts->size_bytes(), stream);
  if (err != cudaSuccess) return ffi::Error(ffi::ErrorCode::kInternal, cudaGetErrorString(err));
  return check(adept_b200_vdfdx_rho_f64(f.typed_data(), out->typed_data(), s.batch, s.nx, s.nv, v.typed_data(), dt, k1x,
                                        nullptr, parts->typed_data(), nparts, stream));
}

// second stage of the fused density: rho = base + charge * (sum_p parts * dv)
ffi::Error ReducePartsImpl(cudaStream_t stream, F64 parts, F64 base, double dv, double charge, F64Out rho) {
  const int nparts = (int)parts.dimensions()[0];
  const long long n = (long long)rho->element_count();
  return check(adept_b200_reduce_parts_f64(parts.typed_data(), nparts, n, dv, charge, base.typed_data(),
                                           rho->typed_data(), stream));
}

// VelocityExponential.push  (vlasov.py:74-91); dex and pond are always passed (zeros when unused)
ffi::Error EdfdvExpImpl(cudaStream_t stream, F64 f, F64 e, F64 dex, F64 pond, double charge, double mass, double dt,
                        double k1v, F64Out out) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  return check(adept_b200_edfdv_exp_f64(f.typed_data(), out->typed_data(), s.batch, s.nx, s.nv, e.typed_data(),
                                        dex.typed_data(), pond.typed_data(), charge, mass, dt, k1v, stream));
}

// adjoint of VelocityExponential.push w.r.t. the acceleration (custom_vjp backward)
ffi::Error EdfdvExpBwdAccelImpl(cudaStream_t stream, F64 f, F64 g, F64 e, F64 dex, F64 pond, double charge,
                                double mass, double dt, double k1v, F64Out accel_bar) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  return check(adept_b200_edfdv_exp_bwd_accel_f64(f.typed_data(), g.typed_data(), s.batch, s.nx, s.nv, e.typed_data(),
                                                  dex.typed_data(), pond.typed_data(), charge, mass, dt, k1v,
                                                  accel_bar->typed_data(), stream));
}

// VelocityCubicSpline.push  (vlasov.py:106-172)
ffi::Error EdfdvSplineImpl(cudaStream_t stream, F64 f, F64 e, F64 dex, F64 pond, double charge, double mass, double dt,
                           double dv, F64Out out) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  return check(adept_b200_edfdv_spline_f64(f.typed_data(), out->typed_data(), s.batch, s.nx, s.nv, e.typed_data(),
                                           dex.typed_data(), pond.typed_data(), charge, mass, dt, dv, stream));
}

// SpectralPoissonSolver.__call__ / BoltzmannPoissonSolver  (field.py:210-224, 282-298)
ffi::Error PoissonImpl(cudaStream_t stream, F64 rho, F64 kmul, int64_t mode, double Te, double lambda_De, F64Out e) {
  auto d = rho.dimensions();
  const int nx = (int)d.back();
  const int batch = d.size() == 2 ? (int)d[0] : 1;
  return check(adept_b200_poisson_f64(rho.typed_data(), kmul.typed_data(), 0, e->typed_data(), batch, nx, (int)mode, Te,
                                      lambda_De, stream));
}

// Collisions.__call__ + Krook  (fokker_planck.py:378-484); nu_K and f_mx are ignored when krook == 0
ffi::Error CollideImpl(cudaStream_t stream, F64 f, F64 v, F64 nu_fp, F64 nu_K, F64 f_mx, double dv, double dt,
                       int64_t model, int64_t scheme, int64_t nodrag, int64_t fp_on, int64_t krook, double sg_m,
                       double sg_ratio, F64Out out) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  return check(adept_b200_collide_f64(f.typed_data(), out->typed_data(), s.batch, s.nx, s.nv, v.typed_data(), dv, dt,
                                      fp_on ? nu_fp.typed_data() : nullptr, krook ? nu_K.typed_data() : nullptr,
                                      f_mx.typed_data(), (int)model, (int)scheme, (int)nodrag, sg_m, sg_ratio, nullptr,
                                      stream));
}

// adjoint of the Fokker-Planck step (central differencing; LB / Dougherty): cotangents of f and of nu
ffi::Error CollideBwdImpl(cudaStream_t stream, F64 f_in, F64 f_new, F64 g, F64 v, F64 nu_fp, double dv, double dt,
                          int64_t model, int64_t scheme, F64Out f_bar, F64Out nu_bar) {
  const Shape3 s = shape3(f_in);
  if (!s.ok) return bad_rank();
  return check(adept_b200_collide_bwd_f64(f_in.typed_data(), f_new.typed_data(), g.typed_data(), f_bar->typed_data(),
                                          nu_bar->typed_data(), s.batch, s.nx, s.nv, v.typed_data(), dv, dt,
                                          nu_fp.typed_data(), (int)model, (int)scheme, stream));
}

// VelocityExponential.push followed by Collisions on the same rows, one pass over f (vector_field.py:236-238)
ffi::Error VpushCollideImpl(cudaStream_t stream, F64 f, F64 e, F64 dex, F64 pond, F64 v, F64 nu_fp, double charge,
                            double mass, double dt, double k1v, double dv, int64_t model, int64_t scheme, F64Out out) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  return check(adept_b200_vpush_collide_f64(f.typed_data(), out->typed_data(), s.batch, s.nx, s.nv, e.typed_data(),
                                            dex.typed_data(), pond.typed_data(), charge, mass, dt, k1v, v.typed_data(),
                                            dv, nu_fp.typed_data(), (int)model, (int)scheme, stream));
}

// in-loop save moments (storage.py:286-327, 119-162): out [6, batch*nx].  diffrax hands the save functions the state
// it has already interpolated, so the two-state form of the C entry point is not needed here.
ffi::Error SaveMomentsImpl(cudaStream_t stream, F64 f, F64 v, double dv, F64Out out) {
  const Shape3 s = shape3(f);
  if (!s.ok) return bad_rank();
  return check(adept_b200_save_moments_f64(f.typed_data(), nullptr, 0.0, s.batch, s.nx, s.nv, v.typed_data(), dv,
                                           out->typed_data(), stream));
}

}  // namespace

#define ADEPT_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_vdfdx, VdfdxImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Attr<double>("dt").Attr<double>("k1x").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_vdfdx_rho, VdfdxRhoImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Attr<double>("dt").Attr<double>("k1x").Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_reduce_parts, ReducePartsImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Attr<double>("dv").Attr<double>("charge").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_edfdv_exp, EdfdvExpImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Attr<double>("charge")
                                  .Attr<double>("mass").Attr<double>("dt").Attr<double>("k1v").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_edfdv_exp_bwd_accel, EdfdvExpBwdAccelImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Attr<double>("charge")
                                  .Attr<double>("mass").Attr<double>("dt").Attr<double>("k1v").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_edfdv_spline, EdfdvSplineImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Attr<double>("charge")
                                  .Attr<double>("mass").Attr<double>("dt").Attr<double>("dv").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_poisson, PoissonImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Attr<int64_t>("mode").Attr<double>("Te")
                                  .Attr<double>("lambda_De").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_collide, CollideImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Attr<double>("dv")
                                  .Attr<double>("dt").Attr<int64_t>("model").Attr<int64_t>("scheme")
                                  .Attr<int64_t>("nodrag").Attr<int64_t>("fp_on").Attr<int64_t>("krook")
                                  .Attr<double>("sg_m").Attr<double>("sg_ratio").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_collide_bwd, CollideBwdImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Attr<double>("dv")
                                  .Attr<double>("dt").Attr<int64_t>("model").Attr<int64_t>("scheme").Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_vpush_collide, VpushCollideImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Attr<double>("charge").Attr<double>("mass").Attr<double>("dt").Attr<double>("k1v")
                                  .Attr<double>("dv").Attr<int64_t>("model").Attr<int64_t>("scheme").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(adept_b200_xla_save_moments, SaveMomentsImpl,
                              ADEPT_STREAM.Arg<F64>().Arg<F64>().Attr<double>("dv").Ret<F64>());
