// oracle/ref_cuda_harness.cu -- TEST / BASELINE INFRASTRUCTURE ONLY.
//
// "The reference's own CUDA kernel recompiled for the same B200" (BASELINE.md B1): the device text of
// Reconstruction/CudaReconstruction.cu (extracted at build time into the git-ignored
// oracle/_ref/ref_kernel_text.inc, never committed) behind a VTK-free host loop that reproduces
// CudaInitialize (:269-298) and the per-view loop of ProcessDepthMap<T> (:343-365): per view a
// cudaDeviceSynchronize, three synchronous H2D copies from pageable memory, one launch with
// grid (1,Ny,Nz) x block (Nx,1,1).  Built twice by oracle/Makefile: -O3 (timing baseline, FMA
// contraction on) and -O3 -fmad=false (numerics of the shipped -G build).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "_ref/ref_kernel_text.inc"

extern "C" {

void ref_cuda_initialize(const double* gridMatrix, const int* gridDims, const double* gridOrig,
                         const double* gridSpacing, double thick, double rho, double eta, double delta,
                         const int* depthMapDims)
{
  cudaMemcpyToSymbol(c_gridMatrix, gridMatrix, SizeMat4x4 * sizeof(TypeCompute));
  cudaMemcpyToSymbol(c_gridDims, gridDims, SizeDim3D * sizeof(int));
  cudaMemcpyToSymbol(c_gridOrig, gridOrig, SizePoint3D * sizeof(TypeCompute));
  cudaMemcpyToSymbol(c_gridSpacing, gridSpacing, SizeDim3D * sizeof(TypeCompute));
  cudaMemcpyToSymbol(c_rayPotentialThick, &thick, sizeof(TypeCompute));
  cudaMemcpyToSymbol(c_rayPotentialRho, &rho, sizeof(TypeCompute));
  cudaMemcpyToSymbol(c_rayPotentialEta, &eta, sizeof(TypeCompute));
  cudaMemcpyToSymbol(c_rayPotentialDelta, &delta, sizeof(TypeCompute));
  cudaMemcpyToSymbol(c_depthMapDims, depthMapDims, 2 * sizeof(int));
  ch_gridDims[0] = gridDims[0]; ch_gridDims[1] = gridDims[1]; ch_gridDims[2] = gridDims[2];
}

}  // extern "C"

// Returns 0 on success.  depths are already threshold-filtered (the reference filters on the host
// before the copy, :348).  io_scalar is uploaded first and downloaded last like :323-327, :368.
// timing[0] = sum of per-launch kernel times (ms, CUDA events; only when perKernelEvents != 0)
// timing[1] = as-driven span: first H2D of view 0 .. last kernel complete (ms, CUDA events)
template <typename T>
static int run(int nViews, const double* depths, const double* K, const double* RT, T* io_scalar,
               int perKernelEvents, float* timing)
{
  const size_t nbVoxels = (size_t)(ch_gridDims[0] - 1) * (ch_gridDims[1] - 1) * (ch_gridDims[2] - 1);
  int2 dd;
  cudaMemcpyFromSymbol(&dd, c_depthMapDims, sizeof(int2));
  const size_t npix = (size_t)dd.x * dd.y;
  if (ch_gridDims[0] - 1 > 1024) return 2;  // the reference's block = Nx threads (:330)

  T* d_out; TypeCompute *d_depth, *d_K, *d_RT;
  CudaErrorCheck(cudaMalloc((void**)&d_out, nbVoxels * sizeof(T)));
  CudaErrorCheck(cudaMemcpy(d_out, io_scalar, nbVoxels * sizeof(T), cudaMemcpyHostToDevice));
  CudaErrorCheck(cudaMalloc((void**)&d_depth, npix * sizeof(TypeCompute)));
  CudaErrorCheck(cudaMalloc((void**)&d_K, SizeMat4x4 * sizeof(TypeCompute)));
  CudaErrorCheck(cudaMalloc((void**)&d_RT, SizeMat4x4 * sizeof(TypeCompute)));
  dim3 dimBlock(ch_gridDims[0] - 1, 1, 1);
  dim3 dimGrid(1, ch_gridDims[1] - 1, ch_gridDims[2] - 1);

  cudaEvent_t e0, e1, k0, k1;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&k0); cudaEventCreate(&k1);
  float ksum = 0.f;
  cudaEventRecord(e0);
  for (int i = 0; i < nViews; i++)
  {
    CudaErrorCheck(cudaDeviceSynchronize());
    if (perKernelEvents && i > 0) { float ms; cudaEventElapsedTime(&ms, k0, k1); ksum += ms; }
    CudaErrorCheck(cudaMemcpy(d_depth, depths + npix * i, npix * sizeof(TypeCompute), cudaMemcpyHostToDevice));
    CudaErrorCheck(cudaMemcpy(d_K, K + 16 * i, SizeMat4x4 * sizeof(TypeCompute), cudaMemcpyHostToDevice));
    CudaErrorCheck(cudaMemcpy(d_RT, RT + 16 * i, SizeMat4x4 * sizeof(TypeCompute), cudaMemcpyHostToDevice));
    if (perKernelEvents) cudaEventRecord(k0);
    depthMapKernel<T><<<dimGrid, dimBlock>>>(d_depth, d_K, d_RT, d_out);
    if (perKernelEvents) cudaEventRecord(k1);
  }
  cudaEventRecord(e1);
  CudaErrorCheck(cudaDeviceSynchronize());
  if (perKernelEvents && nViews > 0) { float ms; cudaEventElapsedTime(&ms, k0, k1); ksum += ms; }
  float span; cudaEventElapsedTime(&span, e0, e1);
  if (timing) { timing[0] = ksum; timing[1] = span; }
  CudaErrorCheck(cudaMemcpy(io_scalar, d_out, nbVoxels * sizeof(T), cudaMemcpyDeviceToHost));
  cudaFree(d_out); cudaFree(d_depth); cudaFree(d_K); cudaFree(d_RT);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(k0); cudaEventDestroy(k1);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" {

int ref_cuda_process(int nViews, const double* depths, const double* K, const double* RT,
                     int scalarType, void* io_scalar, int perKernelEvents, float* timing)
{
  if (scalarType == 1) return run<double>(nViews, depths, K, RT, (double*)io_scalar, perKernelEvents, timing);
  return run<float>(nViews, depths, K, RT, (float*)io_scalar, perKernelEvents, timing);
}

}  // extern "C"
