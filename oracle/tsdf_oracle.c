/*
 * oracle/tsdf_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, IEEE-double, contraction-free restatement of the reference's per-voxel depth-map
 * integration (the work of depthMapKernel<T>, Reconstruction/CudaReconstruction.cu:158-212 of
 * bastienjacquet/CudaDepthMapIntegration) and of the depth threshold filter
 * (Sources/ReconstructionData.cxx:138-167).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product path (cudadepthmapintegration_b200/csrc) never does.
 *
 * Pinning: the reference ships no tests or golden vectors, so this restatement is
 * pinned against the reference's own kernel text compiled by oracle/Makefile into oracle/_ref/
 * (host build with a CUDA-keyword shim, and an nvcc -fmad=false build run on the GPU box);
 * tests/test_oracle_pinning.py requires bit-identical volumes.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (see oracle/Makefile). No -ffast-math, ever.
 *
 * Semantics copied from the device code, with the line each follows:
 *   voxel centre            orig + (idx + 0.5) * spacing                       CudaReconstruction.cu:78-83
 *   4x4 * point (rows 0-2)  m0*x + m1*y + m2*z + m3, left to right, no FMA     :88-93
 *   reject                  h.z < 0 only (h.z == 0 passes -> inf/NaN pixel)    :177
 *   pixel                   (int)round(u): round-half-away, then the CUDA
 *                           double->int conversion: saturating, NaN -> INT_MIN :187-188
 *   bounds                  px<0 || py<0 || px>=W || py>=H                     :192-197
 *   depth index             W*(H-1-py) + px   (VTK images are bottom-up)       :141-149
 *   invalid                 depth == -1 exactly                                :202
 *   real depth              camera z (not the Euclidean range)                 :207
 *   potential               see ray_potential() below                          :105-120
 *   accumulate              output[(k*Ny + j)*Nx + i] += (T)potential          :126-134, :211
 *   view order              list order, one full sweep per view                :343-365
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <limits.h>

#define ORACLE_F32 0
#define ORACLE_F64 1

/* Host threads of the OpenMP loops of the oracle (n > 0 sets the count; returns the count in effect).
 * torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: a timed baseline must say what it used. */
#include <omp.h>
int oracle_threads(int n)
{
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}

/* CUDA's cvt.rzi.s32.f64 (what `int = double` compiles to on the device; SASS F2I.F64.TRUNC):
 * saturating; a NaN from an .f64 source converts to 1 << 31 = INT_MIN (PTX ISA, cvt, "NaN input":
 * zero only when the source is not .f64).  Measured on B200 with the reference kernel itself
 * (tests/test_tsdf_parity_gpu.py::test_edge_cases_reference_kernel): the 0/0 voxel is rejected. */
static inline int cuda_double_to_int(double x)
{
  if (x != x) return INT_MIN;
  if (x >= 2147483647.0) return INT_MAX;
  if (x <= -2147483648.0) return INT_MIN;
  return (int)x;
}

/* transformFrom4Matrix, CudaReconstruction.cu:88-93 */
static inline void transform4(const double* M, const double* p, double* out)
{
  out[0] = M[0] * p[0] + M[1] * p[1] + M[2] * p[2] + M[3];
  out[1] = M[4] * p[0] + M[5] * p[1] + M[6] * p[2] + M[7];
  out[2] = M[8] * p[0] + M[9] * p[1] + M[10] * p[2] + M[11];
}

/* rayPotential<T>, CudaReconstruction.cu:105-120; the result is returned in double and cast to T
 * by the caller exactly where the reference assigns to `res`. */
static inline double ray_potential(double realDistance, double depthMapDistance,
                                   double thick, double rho, double eta, double delta)
{
  double diff = realDistance - depthMapDistance;
  double absoluteDiff = fabs(diff);
  int sign = diff != 0 ? cuda_double_to_int(diff / absoluteDiff) : 0;
  if (absoluteDiff > delta)
    return diff > 0 ? 0.0 : -eta * rho;
  else if (absoluteDiff > thick)
    return rho * sign;
  else
    return (rho / thick) * diff;
}

/* Exposed for the known-answer tests. */
double oracle_ray_potential(double realDistance, double depthMapDistance,
                            double thick, double rho, double eta, double delta)
{
  return ray_potential(realDistance, depthMapDistance, thick, rho, eta, delta);
}

int oracle_round_to_pixel(double u)
{
  return cuda_double_to_int(round(u));
}

/* ReconstructionData::ApplyDepthThresholdFilter, ReconstructionData.cxx:159-166 (strict >). */
void oracle_apply_depth_threshold(double* depths, const double* bestCost, size_t n, double threshold)
{
  for (size_t i = 0; i < n; i++)
    if (bestCost[i] > threshold) depths[i] = -1;
}

/*
 * One view added to cells k in [k0, k1) of the volume.  `out` points at the FULL volume
 * ((dimsP[0]-1)*(dimsP[1]-1)*(dimsP[2]-1) elements, VTK cell order) -- slabs are expressed by
 * the k range only, with global indices, like the product's z-slab sharding.
 */
static void integrate_one_view(const double* gm, const int* dimsP, const double* orig, const double* sp,
                               double thick, double rho, double eta, double delta,
                               int W, int H, const double* depths, const double* K, const double* RT,
                               int scalarType, void* out, int k0, int k1)
{
  const int Nx = dimsP[0] - 1, Ny = dimsP[1] - 1;
#pragma omp parallel for schedule(static)
  for (int k = k0; k < k1; k++)
    for (int j = 0; j < Ny; j++)
      for (int i = 0; i < Nx; i++)
      {
        double c[3], w[3], cam[3], h[3];
        c[0] = orig[0] + (i + 0.5) * sp[0];
        c[1] = orig[1] + (j + 0.5) * sp[1];
        c[2] = orig[2] + (k + 0.5) * sp[2];
        transform4(gm, c, w);
        transform4(RT, w, cam);
        transform4(K, cam, h);
        if (h[2] < 0) continue;
        double u = h[0] / h[2];
        double v = h[1] / h[2];
        int px = cuda_double_to_int(round(u));
        int py = cuda_double_to_int(round(v));
        if (px < 0 || py < 0 || px >= W || py >= H) continue;
        double depth = depths[(size_t)W * (size_t)(H - 1 - py) + (size_t)px];
        if (depth == -1) continue;
        size_t id = ((size_t)k * Ny + j) * Nx + i;
        double r = ray_potential(cam[2], depth, thick, rho, eta, delta);
        if (scalarType == ORACLE_F64)
          ((double*)out)[id] += r;
        else
        {
          float nv = (float)r;
          ((float*)out)[id] += nv;
        }
      }
}

/*
 * The reference's whole ProcessDepthMap loop (CudaReconstruction.cu:343-365) on in-memory views:
 * depths = double[nViews][H][W] bottom-up rows, K / RT = double[nViews][16] row-major.
 * bestCost may be NULL (no filter); otherwise depths are filtered INTO A COPY the caller provides
 * via `scratch` (W*H doubles) so the input stays const.
 */
void oracle_tsdf_integrate(const double* gridMatrix, const int* pointDims, const double* orig,
                           const double* spacing, double thick, double rho, double eta, double delta,
                           int W, int H, int nViews, const double* depths, const double* bestCost,
                           double threshold, const double* K, const double* RT,
                           int scalarType, void* io_scalar, int k0, int k1, double* scratch)
{
  const size_t npix = (size_t)W * H;
  for (int v = 0; v < nViews; v++)
  {
    const double* d = depths + npix * v;
    if (bestCost)
    {
      for (size_t i = 0; i < npix; i++) scratch[i] = d[i];
      oracle_apply_depth_threshold(scratch, bestCost + npix * v, npix, threshold);
      d = scratch;
    }
    integrate_one_view(gridMatrix, pointDims, orig, spacing, thick, rho, eta, delta, W, H, d,
                       K + 16 * v, RT + 16 * v, scalarType, io_scalar, k0, k1);
  }
}

/*
 * Diagnostic twin used by the parity tests: for ONE view, the discrete decision taken for each cell
 * of [k0,k1): -3 behind camera, -2 out of image, -1 invalid depth, else the depth-map index.
 */
void oracle_tsdf_decisions(const double* gm, const int* dimsP, const double* orig, const double* sp,
                           int W, int H, const double* depths, const double* K, const double* RT,
                           int32_t* decision, int k0, int k1)
{
  const int Nx = dimsP[0] - 1, Ny = dimsP[1] - 1;
#pragma omp parallel for schedule(static)
  for (int k = k0; k < k1; k++)
    for (int j = 0; j < Ny; j++)
      for (int i = 0; i < Nx; i++)
      {
        double c[3], w[3], cam[3], h[3];
        size_t id = ((size_t)(k - k0) * Ny + j) * Nx + i;
        c[0] = orig[0] + (i + 0.5) * sp[0];
        c[1] = orig[1] + (j + 0.5) * sp[1];
        c[2] = orig[2] + (k + 0.5) * sp[2];
        transform4(gm, c, w);
        transform4(RT, w, cam);
        transform4(K, cam, h);
        if (h[2] < 0) { decision[id] = -3; continue; }
        int px = cuda_double_to_int(round(h[0] / h[2]));
        int py = cuda_double_to_int(round(h[1] / h[2]));
        if (px < 0 || py < 0 || px >= W || py >= H) { decision[id] = -2; continue; }
        int idx = W * (H - 1 - py) + px;
        decision[id] = depths[idx] == -1 ? -1 : idx;
      }
}
