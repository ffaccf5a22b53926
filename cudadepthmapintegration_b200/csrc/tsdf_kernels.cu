// TSDF depth-map integration kernels for sm_100a.
//
// What the reference does (Reconstruction/CudaReconstruction.cu:158-212, launched once per view at
// :363): one thread per voxel, read-modify-write of the voxel for every view -> 24 B of HBM traffic
// per voxel*view and 26 global loads of the camera matrices per thread.
//
// What this file does instead: VOXEL-STATIONARY integration.  A thread owns M voxels (same i,j, M
// consecutive k), keeps their accumulators in registers, and loops over a whole chunk of views whose
// matrices arrive as by-value kernel parameters (constant bank, uniform loads).  The voxel is read
// and written once per chunk, so HBM traffic falls to 16/chunk B per voxel*view plus the depth
// pixels, which are gathered through L1/L2.  CTAs are numbered so that concurrently resident CTAs
// cover a compact 3-D region ("supertiles"), which keeps the depth-map footprint of a wave small
// enough to live in L2 for every view of the chunk.
//
// Views are added in list order per voxel, like the reference's host loop (:343), so the floating
// point accumulation order is the reference's.
#include "dmi_internal.cuh"
#include "tsdf_device.cuh"

namespace dmi {

// ---- brick / supertile decomposition ---------------------------------------------------------
// brick     = 32 (i) x BJ (j) x M (k) voxels = one CTA of 32 x BJ threads
// supertile = SI x SJ x SK bricks, enumerated contiguously, i fastest
constexpr int BJ = 4;
constexpr int SI = 2, SJ = 8, SK = 8;

struct BrickCoord { int bi, bj, bk; bool valid; };

__device__ __forceinline__ BrickCoord decode_brick(unsigned b, int nbi, int nbj, int nbk)
{
  constexpr unsigned per = SI * SJ * SK;
  const unsigned st = b / per, r = b % per;
  const unsigned nsi = (nbi + SI - 1) / SI, nsj = (nbj + SJ - 1) / SJ;
  const unsigned si = st % nsi, sj = (st / nsi) % nsj, sk = st / (nsi * nsj);
  BrickCoord c;
  c.bi = si * SI + r % SI;
  c.bj = sj * SJ + (r / SI) % SJ;
  c.bk = sk * SK + r / (SI * SJ);
  c.valid = c.bi < nbi && c.bj < nbj && c.bk < nbk;
  return c;
}

static inline unsigned brick_grid_size(int nbi, int nbj, int nbk)
{
  const unsigned nsi = (nbi + SI - 1) / SI, nsj = (nbj + SJ - 1) / SJ, nsk = (nbk + SK - 1) / SK;
  return nsi * nsj * nsk * (SI * SJ * SK);
}

template <typename T, int M>
__global__ void __launch_bounds__(32 * BJ)
tsdf_exact_kernel(const __grid_constant__ GridParams g, const __grid_constant__ ExactChunk c,
                  const double* __restrict__ depths, T* __restrict__ vol, int nbi, int nbj, int nbk)
{
  const BrickCoord b = decode_brick(blockIdx.x, nbi, nbj, nbk);
  if (!b.valid) return;
  const int i = b.bi * 32 + threadIdx.x;
  const int j = b.bj * BJ + threadIdx.y;
  const int lp = b.bk * M, kb = slab_global_k(g, lp);           // layers are multiples of M planes
  if (i >= g.Nx || j >= g.Ny) return;

  double wx[M], wy[M], wz[M];
  T acc[M];
  const size_t plane = (size_t)g.Nx * g.Ny;
  // slab-local storage: cell (i,j, local plane lp) lives at (lp*Ny + j)*Nx + i
  T* p = vol + ((size_t)lp * g.Ny + j) * g.Nx + i;
#pragma unroll
  for (int m = 0; m < M; m++)
  {
    voxel_world(g, i, j, kb + m, wx[m], wy[m], wz[m]);
    acc[m] = (lp + m < g.nLocal) ? p[m * plane] : (T)0;
  }
  const size_t npix = (size_t)g.W * g.H;
  for (int v = 0; v < c.n; v++)
  {
    const double* dv = depths + npix * v;
#pragma unroll
    for (int m = 0; m < M; m++)
      if (lp + m < g.nLocal) integrate_exact<T>(g, c.v[v], dv, wx[m], wy[m], wz[m], acc[m]);
  }
#pragma unroll
  for (int m = 0; m < M; m++)
    if (lp + m < g.nLocal) p[m * plane] = acc[m];
}

cudaError_t launch_tsdf_exact(const GridParams& g, const ExactChunk& c, const double* d_depths,
                              void* d_vol, int scalarType, cudaStream_t s)
{
  constexpr int M = 4;
  const int nbi = (g.Nx + 31) / 32, nbj = (g.Ny + BJ - 1) / BJ, nbk = (g.nLocal + M - 1) / M;
  if (nbi <= 0 || nbj <= 0 || nbk <= 0 || c.n <= 0) return cudaSuccess;
  const unsigned grid = brick_grid_size(nbi, nbj, nbk);
  const dim3 block(32, BJ, 1);
  if (scalarType == 1)
    tsdf_exact_kernel<double, M><<<grid, block, 0, s>>>(g, c, d_depths, (double*)d_vol, nbi, nbj, nbk);
  else
    tsdf_exact_kernel<float, M><<<grid, block, 0, s>>>(g, c, d_depths, (float*)d_vol, nbi, nbj, nbk);
  return cudaGetLastError();
}

// ---- ReconstructionData::ApplyDepthThresholdFilter (ReconstructionData.cxx:159-166) ----------
__global__ void __launch_bounds__(256)
depth_threshold_kernel(double* __restrict__ depths, const double* __restrict__ cost, size_t count, double thr)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n2 = count / 2;
  double2* d2 = reinterpret_cast<double2*>(depths);
  const double2* c2 = reinterpret_cast<const double2*>(cost);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += stride)
  {
    const double2 cv = c2[t];
    if (cv.x > thr || cv.y > thr)
    {
      double2 dv = d2[t];
      if (cv.x > thr) dv.x = -1.0;
      if (cv.y > thr) dv.y = -1.0;
      d2[t] = dv;
    }
  }
  if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0)
    if (cost[count - 1] > thr) depths[count - 1] = -1.0;
}

__global__ void __launch_bounds__(256)
depth_threshold_scalar_kernel(double* __restrict__ depths, const double* __restrict__ cost, size_t count, double thr)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride)
    if (cost[t] > thr) depths[t] = -1.0;
}

cudaError_t launch_depth_threshold(double* d_depths, const double* d_cost, size_t count, double thr,
                                   cudaStream_t s)
{
  if (count == 0) return cudaSuccess;
  const bool aligned = ((reinterpret_cast<uintptr_t>(d_depths) | reinterpret_cast<uintptr_t>(d_cost)) & 15) == 0;
  const size_t work = aligned ? (count + 1) / 2 : count;
  size_t blocks = (work + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (aligned)
    depth_threshold_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_depths, d_cost, count, thr);
  else
    depth_threshold_scalar_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_depths, d_cost, count, thr);
  return cudaGetLastError();
}

}  // namespace dmi
