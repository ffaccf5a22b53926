// FP-pipe issue-rate microbenchmarks.  MEASURED_PEAKS.json carries only the HBM copy bandwidth and the
// bf16 tensor peak; the integration and coloration kernels are bound by the FP64 (and FP32) vector
// pipes, so their roofline denominators are measured here, on the box, in the same run.
#include "dmi_internal.cuh"

namespace dmi {

template <typename F>
__global__ void __launch_bounds__(kPeakThreads) fp_peak_kernel(int iters, float* sink)
{
  F a[kPeakChains];
  const F m = (F)1.0000001, c = (F)1e-7;
#pragma unroll
  for (int q = 0; q < kPeakChains; q++) a[q] = (F)(threadIdx.x + q);
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int q = 0; q < kPeakChains; q++) a[q] = fma(a[q], m, c);
  }
  F s = 0;
#pragma unroll
  for (int q = 0; q < kPeakChains; q++) s += a[q];
  if (s == (F)123456789) sink[0] = (float)s;   // never true; keeps the chains alive
}

// flops per launch = blocks * kPeakThreads * iters * 8 * kPeakChains * 2
cudaError_t launch_fp_peak(int which, int blocks, int iters, float* d_sink, cudaStream_t s)
{
  if (which == 0)
    fp_peak_kernel<double><<<blocks, kPeakThreads, 0, s>>>(iters, d_sink);
  else
    fp_peak_kernel<float><<<blocks, kPeakThreads, 0, s>>>(iters, d_sink);
  return cudaGetLastError();
}

}  // namespace dmi
