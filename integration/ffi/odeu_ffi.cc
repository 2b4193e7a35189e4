// XLA FFI handler for the B200 EKF-RK path: lets a jitted JAX program call odeu_ekf_run as a custom
// call (`jax.ffi.ffi_call("odeu_ekf_run", ...)`) on XLA's own CUDA stream, with the jax.Arrays'
// device buffers passed straight through - the JAX-native form of the binding that
// integration/b200.py does with ctypes.
//
// GATED: this file needs jaxlib's headers (xla/ffi/api/ffi.h), which are absent from the image this
// repository is built and tested in (no jax / jaxlib); integration/ffi/Makefile compiles
// it only where `python -c "import jax.ffi; print(jax.ffi.include_dir())"` succeeds.  It is NOT part
// of libodeu.so and nothing in tests/ or bench.py depends on it.
//
// Replaces, like odeu_ekf_run itself: unroll() scripts/run_filter.py:166-224 and the scan of nll()
// scripts/run_parameter_estimation.py:771-794.
//
// Device layout contract (include/odeu.h): per-trajectory arrays are component-major, trajectory-
// minor ([n][B], [T_obs][L] ...); the Python side (integration/ffi/odeu_jax.py) transposes once.
// The small shared matrices of odeu_ekf_io are HOST data, so they travel as call ATTRIBUTES
// (ffi::Span<const double>), not as buffers.
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "../../include/odeu.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Fail(const char* what, int rc) {
  char msg[600];
  odeu_last_error(msg, sizeof(msg));
  return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + msg);
}

// One plan per distinct descriptor would be cached by the Python side and passed as an opaque
// handle (int64 attribute); the handler itself is stateless and re-entrant per (plan, stream).
ffi::Error EkfRunImpl(cudaStream_t stream,
                      ffi::Buffer<ffi::F64> x0,            // [n][B]
                      ffi::Buffer<ffi::F64> theta,         // [p][B] or size 0 = theta_shared
                      ffi::Buffer<ffi::F64> ys,            // [T_obs][L] or [T_obs][L][B]
                      ffi::Buffer<ffi::U8> correct_flags,  // [T]
                      ffi::Buffer<ffi::S64> xy_index_map,  // [T]
                      ffi::ResultBuffer<ffi::F64> xT,      // [n][B]
                      ffi::ResultBuffer<ffi::F64> PT,      // [n*n][B]
                      ffi::ResultBuffer<ffi::F64> nll,     // [B]
                      ffi::ResultBuffer<ffi::F64> out_x,   // [T_save][n][B] or size 0
                      ffi::ResultBuffer<ffi::F64> out_eps, // [T_save][n][B] or size 0
                      ffi::ResultBuffer<ffi::F64> out_P,   // [T_save][n*n][B] or size 0
                      ffi::ResultBuffer<ffi::F64> out_Ps,  // [T_save][n*n][B] or size 0 (guard reference)
                      ffi::ResultBuffer<ffi::F64> out_yhat,// [T_save][L][B] or size 0
                      ffi::ResultBuffer<ffi::F64> out_S,   // [T_save][L*L][B] or size 0
                      ffi::ResultBuffer<ffi::F64> out_t,   // [T_save] or size 0
                      int64_t plan, int64_t T, int64_t L, int64_t save_interval, int64_t guard_mode,
                      int64_t ys_per_trajectory, double t0, double gamma_sqrt,
                      ffi::Span<const double> P0_sqrt, ffi::Span<const double> Q_sqrt,
                      ffi::Span<const double> H, ffi::Span<const double> R_sqrt,
                      ffi::Span<const double> theta_shared) {
  const auto dims = x0.dimensions();
  if (dims.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "x0 must be [n][B]");
  odeu_ekf_io io = {};
  io.B = dims[1];
  io.T = T;
  io.t0 = t0;
  io.L = static_cast<int32_t>(L);
  io.x0 = x0.typed_data();
  io.P0_sqrt = P0_sqrt.size() ? P0_sqrt.begin() : nullptr;
  io.theta = theta.element_count() ? theta.typed_data() : nullptr;
  io.theta_shared = theta_shared.size() ? theta_shared.begin() : nullptr;
  io.Q_sqrt = Q_sqrt.size() ? Q_sqrt.begin() : nullptr;
  io.gamma_sqrt = gamma_sqrt;
  if (L > 0) {
    io.H = H.begin();
    io.R_sqrt = R_sqrt.begin();
    io.ys = ys.typed_data();
    io.ys_per_trajectory = static_cast<int32_t>(ys_per_trajectory);
    io.correct_flags = correct_flags.typed_data();
    io.xy_index_map = xy_index_map.typed_data();
  }
  io.save_interval = save_interval;
  io.guard_mode = static_cast<int32_t>(guard_mode);
  auto opt = [](ffi::ResultBuffer<ffi::F64>& b) { return b->element_count() ? b->typed_data() : nullptr; };
  io.xT = opt(xT); io.PT = opt(PT); io.nll = opt(nll);
  io.out_x = opt(out_x); io.out_eps = opt(out_eps); io.out_P = opt(out_P); io.out_P_sqrt = opt(out_Ps);
  io.out_yhat = opt(out_yhat); io.out_S = opt(out_S); io.out_t = opt(out_t);
  const int rc = odeu_ekf_run(reinterpret_cast<const odeu_plan*>(plan), &io, stream);
  if (rc != 0) return Fail("odeu_ekf_run", rc);
  return ffi::Error::Success();
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    OdeuEkfRun, EkfRunImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>()   // x0
        .Arg<ffi::Buffer<ffi::F64>>()   // theta
        .Arg<ffi::Buffer<ffi::F64>>()   // ys
        .Arg<ffi::Buffer<ffi::U8>>()    // correct_flags
        .Arg<ffi::Buffer<ffi::S64>>()   // xy_index_map
        .Ret<ffi::Buffer<ffi::F64>>()   // xT
        .Ret<ffi::Buffer<ffi::F64>>()   // PT
        .Ret<ffi::Buffer<ffi::F64>>()   // nll
        .Ret<ffi::Buffer<ffi::F64>>()   // out_x
        .Ret<ffi::Buffer<ffi::F64>>()   // out_eps
        .Ret<ffi::Buffer<ffi::F64>>()   // out_P
        .Ret<ffi::Buffer<ffi::F64>>()   // out_P_sqrt
        .Ret<ffi::Buffer<ffi::F64>>()   // out_yhat
        .Ret<ffi::Buffer<ffi::F64>>()   // out_S
        .Ret<ffi::Buffer<ffi::F64>>()   // out_t
        .Attr<int64_t>("plan")
        .Attr<int64_t>("T")
        .Attr<int64_t>("L")
        .Attr<int64_t>("save_interval")
        .Attr<int64_t>("guard_mode")
        .Attr<int64_t>("ys_per_trajectory")
        .Attr<double>("t0")
        .Attr<double>("gamma_sqrt")
        .Attr<ffi::Span<const double>>("P0_sqrt")
        .Attr<ffi::Span<const double>>("Q_sqrt")
        .Attr<ffi::Span<const double>>("H")
        .Attr<ffi::Span<const double>>("R_sqrt")
        .Attr<ffi::Span<const double>>("theta_shared"));
