/*
 * oracle/color_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's per-point mesh coloration:
 *   MeshColoration::ProcessColoration                       Coloration/MeshColoration.cxx:98-199
 *   ReconstructionData::TransformWorldToDepthMapPosition    Sources/ReconstructionData.cxx:169-182
 *   ReconstructionData::GetColorValue                       Sources/ReconstructionData.cxx:92-116
 *   help::ComputeMedian<double>                             Sources/Helper.h:174-187
 *
 * Third-party arithmetic on this path: VTK (un-vendored, version unpinned -- the reference's
 * CMakeLists.txt:8-18 asks for the pre-9 component names, i.e. VTK 6-8).  Restated here from VTK's
 * published behaviour, anchored on the reference's call sites:
 *   vtkTransform::TransformPoint  (ReconstructionData.cxx:173): m0*x + m1*y + m2*z + m3 per row, in double
 *   vtkTransform::TransformVector (ReconstructionData.cxx:175): m0*x + m1*y + m2*z per row (3x3 part only)
 *   vtkTransform::SetMatrix stores the matrix as given (ReconstructionData.cxx:211,220)
 *   vtkImageData::ComputePointId = x + y*W for a zero-based extent (ReconstructionData.cxx:112)
 *   vtkUnsignedCharArray::SetTuple3(double...) = static_cast<unsigned char> (truncation)  (MeshColoration.cxx:180,185)
 *   vtkPoints::GetPoint promotes float32 storage to double exactly (MeshColoration.cxx:148)
 * x86-64 `int = std::round(double)` is cvttsd2si: NaN / +-inf / out-of-range -> INT_MIN, which the
 * `< 0` test at MeshColoration.cxx:158 then rejects.
 *
 * PINNED: the reference ships no tests, so this file is pinned against the reference's own
 * code instead -- Coloration/MeshColoration.cxx, Sources/ReconstructionData.cxx and Sources/Helper.h compiled
 * unmodified against the VTK stand-in of oracle/vtk_shim/ (oracle/_ref/libref_coloration.so, oracle/Makefile):
 * bit-identical NbProjectedDepthMap / MedianColoration / MeanColoration on the cases of
 * tests/test_coloration_pinning.py (general 3x3 K, points behind cameras, 0/0, inf/NaN points, x.5 pixel
 * boundaries, float32 and float64 points), golden arrays in tests/golden/color_golden.npz, plus the hand-derived
 * known answers of tests/test_oracle_kats.py.  What the stand-in restates of VTK (un-vendored) is listed above.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * Build: gcc -O2 -ffp-contract=off -fopenmp.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define ORACLE_F32 0
#define ORACLE_F64 1

/* x86-64 cvttsd2si: anything not representable -> INT_MIN ("integer indefinite"). */
static inline int x86_double_to_int(double x)
{
  if (!(x > -2147483649.0 && x < 2147483648.0)) return INT_MIN;
  return (int)x;
}

/* ReconstructionData::TransformWorldToDepthMapPosition, ReconstructionData.cxx:169-182.
 * K4 / RT4 are the 4x4 row-major matrices the reference builds (:192-221). */
void oracle_world_to_pixel(const double* K4, const double* RT4, const double* p, int* pixel)
{
  double cam[3], d[3];
  cam[0] = RT4[0] * p[0] + RT4[1] * p[1] + RT4[2] * p[2] + RT4[3];
  cam[1] = RT4[4] * p[0] + RT4[5] * p[1] + RT4[6] * p[2] + RT4[7];
  cam[2] = RT4[8] * p[0] + RT4[9] * p[1] + RT4[10] * p[2] + RT4[11];
  d[0] = K4[0] * cam[0] + K4[1] * cam[1] + K4[2] * cam[2];
  d[1] = K4[4] * cam[0] + K4[5] * cam[1] + K4[6] * cam[2];
  d[2] = K4[8] * cam[0] + K4[9] * cam[1] + K4[10] * cam[2];
  d[0] = d[0] / d[2];
  d[1] = d[1] / d[2];
  pixel[0] = x86_double_to_int(round(d[0]));
  pixel[1] = x86_double_to_int(round(d[1]));
}

static int cmp_double(const void* a, const void* b)
{
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* help::ComputeMedian<double>, Helper.h:174-187 (sorts a by-value copy; n > 0 guaranteed by caller). */
double oracle_median(const double* values, size_t n, double* work)
{
  memcpy(work, values, n * sizeof(double));
  qsort(work, n, sizeof(double), cmp_double);
  size_t mid = n / 2;
  if (n % 2 == 0) return (work[mid] + work[mid - 1]) / 2;
  return work[mid];
}

/*
 * MeshColoration::ProcessColoration, MeshColoration.cxx:140-192, on in-memory views.
 *   xyz      P points, float32 or float64, tightly packed (vtkPoints storage; read as double :147-148)
 *   colors   uint8[nViews][H][W][3], bottom-up rows (the "Color" point-data array, ReconstructionData.cxx:95)
 *   K4, RT4  double[nViews][16]
 *   outputs  pre-zeroed semantics: a point seen by no view keeps 0/0/0 and count 0 (:116-118,124-126,132)
 * Points [p0, p1) only -- the product shards points by index range the same way.
 */
void oracle_colorize(size_t p0, size_t p1, const void* xyz, int xyzType, int nViews,
                     const uint8_t* colors, const double* K4, const double* RT4, int W, int H,
                     uint8_t* mean, uint8_t* median, int32_t* nbProjected)
{
  const size_t npix = (size_t)W * H;
#pragma omp parallel
  {
    double* l0 = (double*)malloc(sizeof(double) * (size_t)(nViews > 0 ? nViews : 1) * 4);
    double* l1 = l0 + nViews;
    double* l2 = l1 + nViews;
    double* work = l2 + nViews;
#pragma omp for schedule(static)
    for (long long id = (long long)p0; id < (long long)p1; id++)
    {
      double position[3];
      if (xyzType == ORACLE_F32)
      {
        const float* f = (const float*)xyz + 3 * id;
        position[0] = f[0]; position[1] = f[1]; position[2] = f[2];
      }
      else
      {
        const double* f = (const double*)xyz + 3 * id;
        position[0] = f[0]; position[1] = f[1]; position[2] = f[2];
      }
      size_t n = 0;
      for (int v = 0; v < nViews; v++)
      {
        int pix[2];
        oracle_world_to_pixel(K4 + 16 * v, RT4 + 16 * v, position, pix);
        if (pix[0] < 0 || pix[1] < 0 || pix[0] >= W || pix[1] >= H) continue;
        const uint8_t* c = colors + (npix * v + (size_t)(H - 1 - pix[1]) * W + pix[0]) * 3;
        l0[n] = c[0]; l1[n] = c[1]; l2[n] = c[2];
        n++;
      }
      mean[3 * id + 0] = mean[3 * id + 1] = mean[3 * id + 2] = 0;
      median[3 * id + 0] = median[3 * id + 1] = median[3 * id + 2] = 0;
      nbProjected[id] = 0;
      if (n != 0)
      {
        /* std::accumulate(begin, end, 0): the accumulator is an INT (MeshColoration.cxx:176-178) */
        int s0 = 0, s1 = 0, s2 = 0;
        for (size_t q = 0; q < n; q++)
        {
          s0 = (int)(s0 + l0[q]); s1 = (int)(s1 + l1[q]); s2 = (int)(s2 + l2[q]);
        }
        double sum0 = s0, sum1 = s1, sum2 = s2, nbVal = (double)n;
        mean[3 * id + 0] = (unsigned char)(sum0 / nbVal);
        mean[3 * id + 1] = (unsigned char)(sum1 / nbVal);
        mean[3 * id + 2] = (unsigned char)(sum2 / nbVal);
        median[3 * id + 0] = (unsigned char)oracle_median(l0, n, work);
        median[3 * id + 1] = (unsigned char)oracle_median(l1, n, work);
        median[3 * id + 2] = (unsigned char)oracle_median(l2, n, work);
        nbProjected[id] = (int32_t)n;
      }
    }
    free(l0);
  }
}
