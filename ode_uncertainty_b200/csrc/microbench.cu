// FP64-pipe micro-benchmark: the measured denominator of the EKF kernel's roofline.
// (MEASURED_PEAKS.json holds only HBM and bf16 figures; the EKF-RK path is bound by the FP64
// FMA pipe)  Each thread runs ILP independent DFMA chains; flops = 2 * ILP *
// iters per thread.
#include "plan.h"

namespace odeu {
template <int ILP>
__global__ void __launch_bounds__(256) dfma_peak_kernel(long long iters, double a, double b, double* out) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = (double)(threadIdx.x + k);
  for (long long i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}
}  // namespace odeu

extern "C" int odeu_bench_dfma(int64_t iters, int32_t blocks, int32_t threads, double* scratch_dev,
                               double* flops_out, void* cuda_stream) {
  if (iters <= 0 || blocks <= 0 || threads <= 0 || threads > 256 || !scratch_dev || !flops_out) {
    odeu::set_error("odeu_bench_dfma: invalid argument");
    return -1;
  }
  constexpr int ILP = 8;
  odeu::dfma_peak_kernel<ILP><<<blocks, threads, 0, (cudaStream_t)cuda_stream>>>(
      (long long)iters, 0.999999, 1e-9, scratch_dev);
  odeu::count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    odeu::set_error("odeu_bench_dfma: launch failed: %s", cudaGetErrorString(err));
    return (int)err;
  }
  *flops_out = 2.0 * ILP * (double)iters * (double)blocks * (double)threads;
  return 0;
}
