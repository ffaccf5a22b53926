// Mesh coloration kernel for sm_100a: one warp per mesh point.
//
// The reference (Coloration/MeshColoration.cxx:140-192) walks points x views on one CPU thread,
// pushes the gathered r,g,b into three std::vector<double>, then copies + std::sorts each for the
// median (Sources/Helper.h:174-187).  Here a warp owns a point; its lanes stride over the views,
// project the point with the reference's exact double arithmetic, gather the colour bytes and count
// them into three 256-bin histograms in shared memory (colours are uchar, so a counting histogram
// IS the sorted multiset).  Sum, count and both middle order statistics come out of one warp-wide
// prefix scan per channel -- all integer, hence bit-exact:
//   mean   = (int sum) / n      == (unsigned char)(sum / (double)n)                 (:176-180)
//   median = odd n: v[n/2]; even n: (v[n/2] + v[n/2-1]) / 2 truncated               (Helper.h:179-186)
#include "dmi_internal.cuh"

namespace dmi {

constexpr int kColorWarps = 8;     // warps (= points in flight) per CTA
constexpr int kBins = 256;

// TransformWorldToDepthMapPosition (Sources/ReconstructionData.cxx:169-182) for one view:
//   cam = RT * p    vtkTransform::TransformPoint   m0*x + m1*y + m2*z + m3
//   d   = K3 * cam  vtkTransform::TransformVector  m0*x + m1*y + m2*z
//   px = (int)std::round(d.x / d.z), py likewise, x86 conversion (anything unrepresentable -> INT_MIN,
//   which the bounds test at MeshColoration.cxx:158-163 rejects).  No z-sign test, no depth test.
// Returns true and the pixel when it falls inside [0,W) x [0,H).
__device__ __forceinline__ bool project_exact(const double* __restrict__ m, int stride, int v,
                                              double x, double y, double z, int W, int H, int& px, int& py)
{
  double r[21];
#pragma unroll
  for (int e = 0; e < 21; e++) r[e] = __ldg(m + (size_t)e * stride + v);
  const double cx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(r[0], x), __dmul_rn(r[1], y)), __dmul_rn(r[2], z)), r[3]);
  const double cy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(r[4], x), __dmul_rn(r[5], y)), __dmul_rn(r[6], z)), r[7]);
  const double cz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(r[8], x), __dmul_rn(r[9], y)), __dmul_rn(r[10], z)), r[11]);
  const double dx = __dadd_rn(__dadd_rn(__dmul_rn(r[12], cx), __dmul_rn(r[13], cy)), __dmul_rn(r[14], cz));
  const double dy = __dadd_rn(__dadd_rn(__dmul_rn(r[15], cx), __dmul_rn(r[16], cy)), __dmul_rn(r[17], cz));
  const double dz = __dadd_rn(__dadd_rn(__dmul_rn(r[18], cx), __dmul_rn(r[19], cy)), __dmul_rn(r[20], cz));
  const double u = round(__ddiv_rn(dx, dz));
  const double w = round(__ddiv_rn(dy, dz));
  // accept iff the x86 conversion would give a value in [0,W) / [0,H): NaN and +-inf fail the compares
  if (!(u >= 0.0 && u < (double)W && w >= 0.0 && w < (double)H)) return false;
  px = (int)u;
  py = (int)w;
  return true;
}

// Order statistic helper: given this lane's 8 consecutive bins (counts c[0..7], bins 8*lane..8*lane+7)
// and the exclusive prefix `before` of the lane, return through shuffles the value of the element of
// rank `rank` (0-based) in the sorted multiset.
__device__ __forceinline__ int select_rank(const unsigned c[8], unsigned before, unsigned rank, int lane)
{
  int found = -1;
  unsigned run = before;
#pragma unroll
  for (int q = 0; q < 8; q++)
  {
    if (found < 0 && rank >= run && rank < run + c[q]) found = lane * 8 + q;
    run += c[q];
  }
  const unsigned who = __ballot_sync(0xffffffffu, found >= 0);
  const int src = __ffs(who) - 1;
  return __shfl_sync(0xffffffffu, found, src);
}

template <typename XYZ>
__global__ void __launch_bounds__(32 * kColorWarps)
colorize_kernel(size_t nPoints, const XYZ* __restrict__ xyz, ColorViews views,
                const uint8_t* __restrict__ colors, int W, int H,
                uint8_t* __restrict__ mean, uint8_t* __restrict__ median, int32_t* __restrict__ nb)
{
  __shared__ unsigned hist[kColorWarps][3][kBins];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned (*h)[kBins] = hist[warp];
  const size_t npix = (size_t)W * H;
  const size_t warpsTotal = (size_t)gridDim.x * kColorWarps;

  for (size_t p = (size_t)blockIdx.x * kColorWarps + warp; p < nPoints; p += warpsTotal)
  {
#pragma unroll
    for (int q = 0; q < 3 * kBins / 32; q++) (&h[0][0])[q * 32 + lane] = 0u;
    __syncwarp();
    // vtkPoints::GetPoint: stored type promoted to double (MeshColoration.cxx:147-148)
    const double x = (double)xyz[3 * p + 0], y = (double)xyz[3 * p + 1], z = (double)xyz[3 * p + 2];
    unsigned s0 = 0, s1 = 0, s2 = 0, n = 0;
    for (int v = lane; v < views.nViews; v += 32)
    {
      int px, py;
      if (project_exact(views.m, views.stride, v, x, y, z, W, H, px, py))
      {
        // GetColorValue: Color[(H-1-py)*W + px] (ReconstructionData.cxx:107-115)
        const uint8_t* c = colors + (npix * v + (size_t)(H - 1 - py) * W + px) * 3;
        const unsigned r = c[0], g = c[1], b = c[2];
        atomicAdd(&h[0][r], 1u); atomicAdd(&h[1][g], 1u); atomicAdd(&h[2][b], 1u);
        s0 += r; s1 += g; s2 += b; n++;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    __syncwarp();
    unsigned med[3] = {0u, 0u, 0u};
    if (n > 0)
    {
      const unsigned hiRank = n / 2, loRank = (n % 2 == 0) ? n / 2 - 1 : n / 2;
#pragma unroll
      for (int ch = 0; ch < 3; ch++)
      {
        unsigned c[8], tot = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { c[q] = h[ch][lane * 8 + q]; tot += c[q]; }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const unsigned before = incl - tot;
        const int a = select_rank(c, before, hiRank, lane);
        const int b = select_rank(c, before, loRank, lane);
        med[ch] = (unsigned)(a + b) >> 1;      // odd n: a == b
      }
    }
    if (lane == 0)
    {
      // n == 0: arrays keep their zero fill (MeshColoration.cxx:116-118,124-126,132)
      mean[3 * p + 0] = n ? (uint8_t)(s0 / n) : 0;
      mean[3 * p + 1] = n ? (uint8_t)(s1 / n) : 0;
      mean[3 * p + 2] = n ? (uint8_t)(s2 / n) : 0;
      median[3 * p + 0] = (uint8_t)med[0];
      median[3 * p + 1] = (uint8_t)med[1];
      median[3 * p + 2] = (uint8_t)med[2];
      nb[p] = (int32_t)n;
    }
    __syncwarp();
  }
}

cudaError_t launch_colorize(size_t nPoints, const void* d_xyz, int xyzType, ColorViews views,
                            const uint8_t* d_colors, int W, int H, uint8_t* d_mean, uint8_t* d_median,
                            int32_t* d_nb, cudaStream_t s)
{
  if (nPoints == 0) return cudaSuccess;
  size_t blocks = (nPoints + kColorWarps - 1) / kColorWarps;
  const size_t cap = 148 * 8;                       // persistent: CTAs stride over the points
  if (blocks > cap) blocks = cap;
  if (xyzType == 1)
    colorize_kernel<double><<<(unsigned)blocks, 32 * kColorWarps, 0, s>>>(
        nPoints, (const double*)d_xyz, views, d_colors, W, H, d_mean, d_median, d_nb);
  else
    colorize_kernel<float><<<(unsigned)blocks, 32 * kColorWarps, 0, s>>>(
        nPoints, (const float*)d_xyz, views, d_colors, W, H, d_mean, d_median, d_nb);
  return cudaGetLastError();
}

}  // namespace dmi
