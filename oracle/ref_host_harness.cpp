// oracle/ref_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the REFERENCE's own device code (the text of Reconstruction/CudaReconstruction.cu from
// `#define SizeMat4x4 16` up to the host helpers, extracted at build time by oracle/Makefile into the
// git-ignored oracle/_ref/ref_kernel_text.inc -- never committed) on the host CPU, by giving the CUDA
// keywords it uses a host meaning.  Purpose: pin oracle/tsdf_oracle.c against the reference's actual
// statements in this GPU-less container, and serve as the `"kind": "reference"` CPU baseline.
//
// Host stand-ins and the device behaviour each one reproduces:
//   __constant__ T x      -> plain global (set by ref_host_initialize like CudaInitialize, :282-290)
//   threadIdx / blockIdx  -> thread-local structs set per voxel (launch shape of :330-331, :363)
//   round(x) in `int = round(double)` (:187-188) -> CUDA's cvt.rzi.s32.f64 of round-half-away
//                            (saturating; NaN from an f64 source -> INT_MIN, PTX ISA cvt); x86 agrees on NaN
//                            but not on saturation (it gives INT_MIN for every out-of-range value).
// Compiled with -ffp-contract=off, i.e. the numerics of the shipped `-G` build (no FMA contraction).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <stdlib.h>
#include <omp.h>

using std::abs;
using std::sqrt;

#define __constant__
#define __device__
#define __global__
#define __host__
struct int3 { int x, y, z; };
struct int2 { int x, y; };
struct ref_uint3 { unsigned x, y, z; };
static thread_local ref_uint3 threadIdx, blockIdx;
typedef int cudaError_t;
static const int cudaSuccess = 0;
static inline const char* cudaGetErrorString(cudaError_t) { return "host build"; }

static inline int ref_cuda_round_to_int(double x)
{
  double r = std::round(x);
  if (r != r) return INT_MIN;
  if (r >= 2147483647.0) return INT_MAX;
  if (r <= -2147483648.0) return INT_MIN;
  return (int)r;
}
#define round(x) ref_cuda_round_to_int(x)

#include "_ref/ref_kernel_text.inc"

#undef round

extern "C" {

// Host threads of the OpenMP loop below: n > 0 sets the count (torch.distributed.run exports OMP_NUM_THREADS=1
// to its workers, which would silently make this a single-threaded baseline); returns the count in effect.
int ref_host_threads(int n)
{
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}

// CudaInitialize, CudaReconstruction.cu:269-298, minus the vtkMatrix4x4 unpacking.
void ref_host_initialize(const double* gridMatrix, const int* gridDims, const double* gridOrig,
                         const double* gridSpacing, double thick, double rho, double eta, double delta,
                         const int* depthMapDims)
{
  memcpy(c_gridMatrix, gridMatrix, sizeof(double) * 16);
  c_gridDims.x = gridDims[0]; c_gridDims.y = gridDims[1]; c_gridDims.z = gridDims[2];
  memcpy(c_gridOrig, gridOrig, sizeof(double) * 3);
  memcpy(c_gridSpacing, gridSpacing, sizeof(double) * 3);
  c_rayPotentialThick = thick; c_rayPotentialRho = rho;
  c_rayPotentialEta = eta; c_rayPotentialDelta = delta;
  c_depthMapDims.x = depthMapDims[0]; c_depthMapDims.y = depthMapDims[1];
  ch_gridDims[0] = gridDims[0]; ch_gridDims[1] = gridDims[1]; ch_gridDims[2] = gridDims[2];
}

// The view loop of ProcessDepthMap<T> (:343-365) with views already in memory and already
// threshold-filtered; "launches" depthMapKernel<T> over grid (1,Ny,Nz) x block (Nx,1,1).
// kz0/kz1 restrict the z range so the bench can time a bounded sample of a large grid.
void ref_host_process(int nViews, const double* depths, const double* K, const double* RT,
                      int scalarType, void* io_scalar, int kz0, int kz1)
{
  const int Nx = ch_gridDims[0] - 1, Ny = ch_gridDims[1] - 1;
  const size_t npix = (size_t)c_depthMapDims.x * c_depthMapDims.y;
  for (int v = 0; v < nViews; v++)
  {
    double* d = const_cast<double*>(depths + npix * v);
    double* k4 = const_cast<double*>(K + 16 * v);
    double* rt = const_cast<double*>(RT + 16 * v);
#pragma omp parallel for schedule(static) collapse(2)
    for (int kz = kz0; kz < kz1; kz++)
      for (int jy = 0; jy < Ny; jy++)
      {
        blockIdx.x = 0; blockIdx.y = (unsigned)jy; blockIdx.z = (unsigned)kz;
        threadIdx.y = threadIdx.z = 0;
        for (int ix = 0; ix < Nx; ix++)
        {
          threadIdx.x = (unsigned)ix;
          if (scalarType == 1)
            depthMapKernel<double>(d, k4, rt, (double*)io_scalar);
          else
            depthMapKernel<float>(d, k4, rt, (float*)io_scalar);
        }
      }
  }
}

}  // extern "C"
