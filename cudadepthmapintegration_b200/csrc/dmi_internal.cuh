// Internal declarations shared by the translation units of libdmi_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace dmi {

// Per-run parameters: what the reference keeps in __constant__ symbols (CudaReconstruction.cu:55-63),
// here a by-value kernel parameter so that contexts are independent.
struct GridParams
{
  double gm[12];        // grid matrix rows 0..2, row-major (c_gridMatrix)
  double orig[3];       // c_gridOrig
  double sp[3];         // c_gridSpacing
  int Nx, Ny, Nz;       // CELLS of the whole grid = c_gridDims - 1
  int k0, k1;           // z-slab owned by this context, global cell indices
  int W, H;             // c_depthMapDims
  double thick, rho, eta, delta;
  double rho_over_thick;  // (rho / thick) in double, as CudaReconstruction.cu:119 evaluates it
  double neg_eta_rho;     // -eta * rho, CudaReconstruction.cu:115
};

// One view, reference form: rows 0..2 of matrixTR and matrixK (CudaReconstruction.cu:172,176).
struct ViewExact
{
  double RT[12];
  double K[12];
};

constexpr int kExactChunk = 32;   // views per launch of the exact kernel (6 KB of kernel parameters)
struct ExactChunk
{
  int n;
  int pad;
  ViewExact v[kExactChunk];
};

// One view, fast-path form (see tsdf_kernels.cu).
struct ViewFast
{
  double P[12];     // rows 0..2 of K * RT * G_affine: homogeneous pixel coords from VOXEL INDICES (i,j,k,1)
  double Z[4];      // camera z from voxel indices (row 2 of RT * G_affine)
};

constexpr int kFastChunk = 128;   // views per launch of the fast kernel (16 KB parameters + exact copy in global)
struct FastChunk
{
  int n;
  int pad;
  ViewFast v[kFastChunk];
};

cudaError_t launch_tsdf_exact(const GridParams& g, const ExactChunk& c, const double* d_depths,
                              void* d_vol, int scalarType, cudaStream_t s);
cudaError_t launch_depth_threshold(double* d_depths, const double* d_cost, size_t count, double thr,
                                   cudaStream_t s);

// coloration
struct ColorViews   // device SoA: m[e][v], e in 0..20 = RT rows 0..2 (12) then K 3x3 (9)
{
  const double* m;
  int nViews;
  int stride;       // padded view count
};
cudaError_t launch_colorize(size_t nPoints, const void* d_xyz, int xyzType, ColorViews views,
                            const uint8_t* d_colors, int W, int H, uint8_t* d_mean, uint8_t* d_median,
                            int32_t* d_nb, cudaStream_t s);

// microbenchmarks
cudaError_t launch_fp_peak(int which, int blocks, int iters, float* d_sink, cudaStream_t s);
constexpr int kPeakThreads = 256;
constexpr int kPeakChains = 8;

}  // namespace dmi
