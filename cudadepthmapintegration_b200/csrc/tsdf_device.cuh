// Device functions shared by the integration kernels: the reference's arithmetic, operation for
// operation (Reconstruction/CudaReconstruction.cu:78-120, 158-212).
#pragma once
#include "dmi_internal.cuh"

namespace dmi {

// ---- the reference's arithmetic, operation for operation ---------------------------------------
// transformFrom4Matrix (CudaReconstruction.cu:88-93): m0*x + m1*y + m2*z + m3, left to right.
// The intrinsics forbid FMA contraction, i.e. the numerics of the shipped -G build.
__device__ __forceinline__ double row_point(const double* m, double x, double y, double z)
{
  return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z)), m[3]);
}

// rayPotential<T> (CudaReconstruction.cu:105-120).  `sign` there is (int)(diff/|diff|) = +-1 for any
// finite non-zero diff and 0 for diff == 0, so rho*sign is copysign(rho, diff) or 0.
template <typename T>
__device__ __forceinline__ T ray_potential(const GridParams& g, double realDistance, double depthMapDistance)
{
  const double diff = __dsub_rn(realDistance, depthMapDistance);
  const double a = fabs(diff);
  T res;
  if (a > g.delta)
    res = (T)(diff > 0 ? 0.0 : g.neg_eta_rho);
  else if (a > g.thick)
    res = (T)(diff > 0 ? g.rho : (diff < 0 ? -g.rho : __dmul_rn(g.rho, 0.0)));
  else
    res = (T)__dmul_rn(g.rho_over_thick, diff);
  return res;
}

// One voxel, one view, exactly as depthMapKernel does it from the world-space voxel centre on
// (CudaReconstruction.cu:170-211).  (wx,wy,wz) is view-invariant and is computed once per voxel.
// `cls` (optional): the float classification image of the fast path, in which -1.0f marks exactly the
// pixels that are invalid AFTER the best-cost filter (ReconstructionData.cxx:159-166); `depth` may then
// be the caller's unfiltered map, whose valid pixels the filter leaves untouched.
// Lossless split of a depth d into hi = the float of the classification image and lo = (d - hi) in units of
// 2^(e - 53), e = max(exponent(hi), -64): d - hi is exact (hi is d rounded to float, or a neighbour), at most
// 1.5 float ulps = 1.5 * 2^30 units, and a multiple of the unit whenever |d| >= 2^-64 (or d == 0).  8 bytes
// per pixel carry the classification image AND the exact double depth.  Non-finite hi: lo = 0, d = hi.
__device__ __forceinline__ int split_exponent(float hi)
{
  return max(((__float_as_int(hi) >> 23) & 0xff) - 127, -64);
}
__device__ __forceinline__ int split_encode(double d, float hi)
{
  if (!(fabsf(hi) <= 3.402823466e+38f)) return 0;
  const double scale = __hiloint2double((1023 + 53 - split_exponent(hi)) << 20, 0);   // 2^(53 - e)
  return __double2int_rn(__dmul_rn(__dsub_rn(d, (double)hi), scale));
}
__device__ __forceinline__ double split_decode(float hi, int lo)
{
  const double unit = __hiloint2double((1023 - 53 + split_exponent(hi)) << 20, 0);    // 2^(e - 53)
  return __fma_rn((double)lo, unit, (double)hi);
}

// `depth` may be null when `lo` (the split residual image, with `cls` as its hi part) is given.
template <typename T>
__device__ __forceinline__ void integrate_exact(const GridParams& g, const ViewExact& V,
                                                const double* __restrict__ depth,
                                                double wx, double wy, double wz, T& acc,
                                                const float* __restrict__ cls = nullptr,
                                                const int* __restrict__ lo = nullptr)
{
  const double cx = row_point(V.RT + 0, wx, wy, wz);
  const double cy = row_point(V.RT + 4, wx, wy, wz);
  const double cz = row_point(V.RT + 8, wx, wy, wz);
  const double hx = row_point(V.K + 0, cx, cy, cz);
  const double hy = row_point(V.K + 4, cx, cy, cz);
  const double hz = row_point(V.K + 8, cx, cy, cz);
  if (hz < 0) return;                                   // :177
  const double u = __ddiv_rn(hx, hz);                   // :183
  const double v = __ddiv_rn(hy, hz);                   // :184
  const int px = __double2int_rz(round(u));             // :187 (cvt.rzi.s32.f64: saturating, NaN -> INT_MIN)
  const int py = __double2int_rz(round(v));             // :188
  if (px < 0 || py < 0 || px >= g.W || py >= g.H) return;   // :192-197
  const size_t id = (size_t)g.W * (size_t)(g.H - 1 - py) + px;                 // :141-149
  if (cls && __ldg(cls + id) == -1.0f) return;          // filtered out, or -1 in the file
  const double d = lo ? split_decode(__ldg(cls + id), __ldg(lo + id)) : __ldg(depth + id);   // :201
  if (d == -1) return;                                  // :202
  acc += ray_potential<T>(g, cz, d);                    // :207-211
}

// Voxel centre in world space: computeVoxelCenter + transformFrom4Matrix(c_gridMatrix, ...)
// (CudaReconstruction.cu:78-83, :168), with GLOBAL indices so z-slabs reproduce the full grid.
__device__ __forceinline__ void voxel_world(const GridParams& g, int i, int j, int k,
                                            double& wx, double& wy, double& wz)
{
  const double x = __dadd_rn(g.orig[0], __dmul_rn((double)i + 0.5, g.sp[0]));
  const double y = __dadd_rn(g.orig[1], __dmul_rn((double)j + 0.5, g.sp[1]));
  const double z = __dadd_rn(g.orig[2], __dmul_rn((double)k + 0.5, g.sp[2]));
  wx = row_point(g.gm + 0, x, y, z);
  wy = row_point(g.gm + 4, x, y, z);
  wz = row_point(g.gm + 8, x, y, z);
}

}  // namespace dmi
