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

// Projection with the view already in registers (r[0..11] = RT rows 0..2, r[12..20] = K 3x3).
__device__ __forceinline__ bool project_exact_regs(const double (&r)[21], double x, double y, double z,
                                                   int W, int H, int& px, int& py)
{
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

// One warp colours Q points at a time.  For each tile of 32 views a lane loads ITS view once (21 doubles,
// coalesced from the SoA matrix array) and projects the Q points with it, so the matrix traffic is
// amortised Q times and the Q independent projections hide each other's FP64 / gather latency.
// Histograms: 3 channels x 256 bins x 16-bit counters per point, two bins per 32-bit word (shared-memory
// atomics are 32-bit); requires nViews < 65536, which the launcher checks.
template <typename XYZ, int Q>
__global__ void __launch_bounds__(32 * kColorWarps)
colorize_kernel(size_t nPoints, const XYZ* __restrict__ xyz, ColorViews views,
                const uint8_t* __restrict__ colors, int W, int H,
                uint8_t* __restrict__ mean, uint8_t* __restrict__ median, int32_t* __restrict__ nb)
{
  __shared__ unsigned hist[kColorWarps][Q][3][kBins / 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t npix = (size_t)W * H;
  const size_t batches = (nPoints + Q - 1) / Q;
  const size_t warpsTotal = (size_t)gridDim.x * kColorWarps;

  for (size_t bt = (size_t)blockIdx.x * kColorWarps + warp; bt < batches; bt += warpsTotal)
  {
    const size_t p0 = bt * Q;
    unsigned* hw = &hist[warp][0][0][0];
#pragma unroll
    for (int q = 0; q < Q * 3 * kBins / 2 / 32; q++) hw[q * 32 + lane] = 0u;
    __syncwarp();
    // vtkPoints::GetPoint: stored type promoted to double (MeshColoration.cxx:147-148)
    double x[Q], y[Q], z[Q];
    unsigned s0[Q], s1[Q], s2[Q], n[Q];
#pragma unroll
    for (int q = 0; q < Q; q++)
    {
      const size_t p = min(p0 + q, nPoints - 1);
      x[q] = (double)xyz[3 * p + 0]; y[q] = (double)xyz[3 * p + 1]; z[q] = (double)xyz[3 * p + 2];
      s0[q] = s1[q] = s2[q] = n[q] = 0u;
    }
    for (int v = lane; v < views.nViews; v += 32)
    {
      double r[21];
#pragma unroll
      for (int e = 0; e < 21; e++) r[e] = __ldg(views.m + (size_t)e * views.stride + v);
      const uint8_t* img = colors + npix * 3 * (size_t)v;
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
        int px, py;
        if (project_exact_regs(r, x[q], y[q], z[q], W, H, px, py))
        {
          // GetColorValue: Color[(H-1-py)*W + px] (ReconstructionData.cxx:107-115)
          const uint8_t* c = img + ((size_t)(H - 1 - py) * W + px) * 3;
          const unsigned cr = c[0], cg = c[1], cb = c[2];
          atomicAdd(&hist[warp][q][0][cr >> 1], 1u << ((cr & 1) * 16));
          atomicAdd(&hist[warp][q][1][cg >> 1], 1u << ((cg & 1) * 16));
          atomicAdd(&hist[warp][q][2][cb >> 1], 1u << ((cb & 1) * 16));
          s0[q] += cr; s1[q] += cg; s2[q] += cb; n[q]++;
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < Q; q++)
    {
      unsigned a0 = s0[q], a1 = s1[q], a2 = s2[q], an = n[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        an += __shfl_xor_sync(0xffffffffu, an, o);
      }
      unsigned med[3] = {0u, 0u, 0u};
      if (an > 0)
      {
        const unsigned hiRank = an / 2, loRank = (an % 2 == 0) ? an / 2 - 1 : an / 2;
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
        {
          unsigned c[8], tot = 0;
#pragma unroll
          for (int k = 0; k < 4; k++)
          {
            const unsigned wv = hist[warp][q][ch][lane * 4 + k];
            c[2 * k] = wv & 0xffffu; c[2 * k + 1] = wv >> 16;
            tot += c[2 * k] + c[2 * k + 1];
          }
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
      const size_t p = p0 + q;
      if (lane == 0 && p < nPoints)
      {
        // n == 0: arrays keep their zero fill (MeshColoration.cxx:116-118,124-126,132)
        mean[3 * p + 0] = an ? (uint8_t)(a0 / an) : 0;
        mean[3 * p + 1] = an ? (uint8_t)(a1 / an) : 0;
        mean[3 * p + 2] = an ? (uint8_t)(a2 / an) : 0;
        median[3 * p + 0] = (uint8_t)med[0];
        median[3 * p + 1] = (uint8_t)med[1];
        median[3 * p + 2] = (uint8_t)med[2];
        nb[p] = (int32_t)an;
      }
    }
    __syncwarp();
  }
}

cudaError_t launch_colorize(size_t nPoints, const void* d_xyz, int xyzType, ColorViews views,
                            const uint8_t* d_colors, int W, int H, uint8_t* d_mean, uint8_t* d_median,
                            int32_t* d_nb, cudaStream_t s)
{
  if (nPoints == 0) return cudaSuccess;
  if (views.nViews >= 65536) return cudaErrorInvalidValue;   // 16-bit histogram counters
  constexpr int Q = 4;
  size_t blocks = ((nPoints + Q - 1) / Q + kColorWarps - 1) / kColorWarps;
  const size_t cap = 148 * 8;                       // persistent: CTAs stride over the point batches
  if (blocks > cap) blocks = cap;
  if (xyzType == 1)
    colorize_kernel<double, Q><<<(unsigned)blocks, 32 * kColorWarps, 0, s>>>(
        nPoints, (const double*)d_xyz, views, d_colors, W, H, d_mean, d_median, d_nb);
  else
    colorize_kernel<float, Q><<<(unsigned)blocks, 32 * kColorWarps, 0, s>>>(
        nPoints, (const float*)d_xyz, views, d_colors, W, H, d_mean, d_median, d_nb);
  return cudaGetLastError();
}

}  // namespace dmi
