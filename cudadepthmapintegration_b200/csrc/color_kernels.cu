// Mesh coloration kernel for sm_100a: one warp colours Q mesh points at a time.
//
// The reference (Coloration/MeshColoration.cxx:140-192) walks points x views on one CPU thread,
// pushes the gathered r,g,b into three std::vector<double>, then copies + std::sorts each for the
// median (Sources/Helper.h:174-187).  Here the lanes of a warp stride over the views; each lane keeps
// ITS view's projection in registers for a tile of 32 views and projects the warp's Q points with it,
// gathers the colour bytes and counts them into per-point 3 x 256-bin histograms in shared memory
// (colours are uchar, so a counting histogram IS the sorted multiset).  Sum, count and both middle
// order statistics come out of one warp-wide prefix scan per channel -- all integer, hence bit-exact:
//   mean   = (int sum) / n      == (unsigned char)(sum / (double)n)                 (:176-180)
//   median = odd n: v[n/2]; even n: (v[n/2] + v[n/2-1]) / 2 truncated               (Helper.h:179-186)
//
// Which pixel a point falls on (ReconstructionData::TransformWorldToDepthMapPosition, ~37 uncontracted
// FP64 operations + 2 IEEE divisions in the reference) is decided by the same three certified tiers as
// the integration kernel (DESIGN.md "certification"):
//   T1  FP32: the composed rows are evaluated in FP64 at the batch's FIRST point (per lane-view, amortised over the
//       batch) and rounded to float; each point adds its float offset from that point through the float coefficients
//       (like the integration kernel's brick bases: the large, cancelling terms stay in FP64, the FP32 error scales with
//       the small local terms).  MUFU.RCP, magic-number rounding, distance to the integer against 0.5 - (E*|r| + c0),
//       E = kE * ((|b_n| + A_n d) + U1 (|b_z| + A_z d)).  Mesh points are float32 (vtkPoints' default), hence exact inputs.
//   T2  FP64 composed rows + residual test of the candidate (margin 2^-44 of the magnitudes the
//       reference's own evaluation order goes through).
//   T3  the reference's operation sequence (project_exact), on ties / near-zero denominators / double
//       points that are not float-representable.
// There is no z-sign test and no depth test in the reference (a point behind the camera still projects):
// all tiers work with |d.z|.
#include "dmi_internal.cuh"

#include <cmath>
#include <algorithm>

namespace dmi {

constexpr int kColorWarps = 4;     // warps per CTA
constexpr int kBins = 256;
constexpr float kMagicC = 12582912.0f;
constexpr int kMagicBitsC = 0x4B400000;

__device__ __forceinline__ float rcp_approx_c(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// T3 -- TransformWorldToDepthMapPosition (Sources/ReconstructionData.cxx:169-182) for one view:
//   cam = RT * p    vtkTransform::TransformPoint   m0*x + m1*y + m2*z + m3
//   d   = K3 * cam  vtkTransform::TransformVector  m0*x + m1*y + m2*z
//   px = (int)std::round(d.x / d.z), py likewise, x86 conversion (anything unrepresentable -> INT_MIN,
//   which the bounds test at MeshColoration.cxx:158-163 rejects).
// Returns the storage index (H-1-py)*W + px of GetColorValue (ReconstructionData.cxx:107-115), or -1.
__device__ __noinline__ int project_exact(const double* __restrict__ m, int stride, int v,
                                          double x, double y, double z, int W, int H)
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
  if (!(u >= 0.0 && u < (double)W && w >= 0.0 && w < (double)H)) return -1;
  return (H - 1 - (int)w) * W + (int)u;
}

// T2, falling through to T3.  pu, pv: T1's candidate (centred integer pixel as float), or NaN for "none".
__device__ __noinline__ int project_slow(const ColorViews& views, int v, double x, double y, double z, float mf,
                                         float pu, float pv, int W, int H)
{
  const double m = (double)mf;
  const ColorViewT2& V = views.t2[v];
  const double dz = fma(x, V.dz[0], fma(y, V.dz[1], fma(z, V.dz[2], V.dz[3])));
  const double adz = fabs(dz);
  const double mz = fma(V.mza, m, V.mzb), m2 = fma(V.m2a, m, V.m2b);
  const double half = 0.5 * adz;
  double cu = (double)pu, cv = (double)pv;
  if (adz > mz && m2 < 0.01 * half && fabs(cu) < 4194304.0 && fabs(cv) < 4194304.0)
  {
    const double sg = dz < 0.0 ? -1.0 : 1.0;
    double su = sg * fma(-cu, dz, fma(x, V.nx[0], fma(y, V.nx[1], fma(z, V.nx[2], V.nx[3]))));   // (u' - cu) * |dz|
    double sv = sg * fma(-cv, dz, fma(x, V.ny[0], fma(y, V.ny[1], fma(z, V.ny[2], V.ny[3]))));
    if (su >= half + m2) { cu += 1.0; su -= adz; } else if (su <= -half - m2) { cu -= 1.0; su += adz; }
    if (sv >= half + m2) { cv += 1.0; sv -= adz; } else if (sv <= -half - m2) { cv -= 1.0; sv += adz; }
    if (fabs(su) < half - m2 && fabs(sv) < half - m2)
    {
      const int px = (int)cu + views.cxc, py = (int)cv + views.cyc;
      return ((unsigned)px < (unsigned)W && (unsigned)py < (unsigned)H) ? (H - 1 - py) * W + px : -1;
    }
  }
  return project_exact(views.m, views.stride, v, x, y, z, W, H);
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

#ifndef DMI_COLOR_Q
#define DMI_COLOR_Q 16           // points per warp batch: each lane projects them with ITS view (amortises the view's rows,
#endif                           // and neighbouring points share image sectors); 3 x 256 x 16-bit bins of shared memory each

// Histograms: 3 channels x 256 bins x 16-bit counters per point, two bins per 32-bit word (shared-memory
// atomics are 32-bit); requires nViews < 65536, which the launcher checks.  Sum and count come out of the histogram
// too (sum = sum_b b * count_b), so the lanes carry no per-point accumulators through the view loop.
// The view loop runs in three phases per lane-view, so that the Q colour gathers of a lane are in flight together:
//   1  project the Q points (FP32 tier; uncertified ones through the FP64 / exact tiers afterwards, one call site)
//   2  gather: the RGB triple = 3 bytes at byte offset 3*idx, fetched as the two aligned 32-bit words around it
//   3  three shared-memory atomics per hit
template <typename XYZ, int Q>
__global__ void __launch_bounds__(32 * kColorWarps)
colorize_kernel(size_t nPoints, const XYZ* __restrict__ xyz, const unsigned* __restrict__ perm,
                const __grid_constant__ ColorViews views, const uint8_t* __restrict__ colors, size_t colorBytes, int W, int H,
                uint8_t* __restrict__ mean, uint8_t* __restrict__ median, int32_t* __restrict__ nb)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // per warp: the reference point + the largest offset (4 doubles), Q offsets as float4, then Q histograms of 3 x 128 words
  // (+ 1 word of padding each, so that the same bin of different points falls into different banks)
  constexpr int kHist = 3 * (kBins / 2) + 1;
  constexpr int kWarpWords = 8 + Q * 4 + Q * kHist;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* wbase = reinterpret_cast<unsigned*>(smem_raw) + (size_t)warp * kWarpWords;
  double* pref = reinterpret_cast<double*>(wbase);
  float4* pw = reinterpret_cast<float4*>(wbase + 8);
  unsigned* hw = wbase + 8 + Q * 4;
  const size_t npix = (size_t)W * H;
  const size_t batches = (nPoints + Q - 1) / Q;
  const size_t warpsTotal = (size_t)gridDim.x * kColorWarps;
  const int pxoff = kMagicBitsC - views.cxc, pyoff = kMagicBitsC - views.cyc;
  const float T = views.T;
  const unsigned* words = reinterpret_cast<const unsigned*>(colors);     // cudaMalloc'ed or 4-byte aligned (checked by the launcher)
  const size_t lastWord = (colorBytes - 1) >> 2;

  for (size_t bt = (size_t)blockIdx.x * kColorWarps + warp; bt < batches; bt += warpsTotal)
  {
    const size_t p0 = bt * Q;
    for (int q = lane; q < Q * kHist; q += 32) hw[q] = 0u;
    // vtkPoints::GetPoint: stored type promoted to double (MeshColoration.cxx:147-148).  T1 needs exact float inputs:
    // a point that is not float-representable is marked by a NaN (never certified).  pw[q] = offset of point q from the
    // batch's first point (x, y, z) and, in w, 1 or that NaN.
    {
      const size_t p = perm[min(p0 + min(lane, Q - 1), nPoints - 1)];       // batches follow the spatially sorted order
      const double xd = (double)xyz[3 * p + 0], yd = (double)xyz[3 * p + 1], zd = (double)xyz[3 * p + 2];
      const float xf = (float)xd, yf = (float)yd, zf = (float)zd;
      const bool exact = (double)xf == xd && (double)yf == yd && (double)zf == zd;
      const float x0 = __shfl_sync(0xffffffffu, xf, 0), y0 = __shfl_sync(0xffffffffu, yf, 0), z0 = __shfl_sync(0xffffffffu, zf, 0);
      const bool ok0 = __shfl_sync(0xffffffffu, exact ? 1 : 0, 0) != 0;
      const float dx = xf - x0, dy = yf - y0, dz = zf - z0;
      float d = (exact && ok0) ? fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz))) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
      if (lane < Q) pw[lane] = make_float4(dx, dy, dz, (exact && ok0) ? 1.f : NAN);
      if (lane == 0) { pref[0] = (double)x0; pref[1] = (double)y0; pref[2] = (double)z0; pref[3] = (double)d; }
    }
    __syncwarp();
    const double rx0 = pref[0], ry0 = pref[1], rz0 = pref[2];
    const float dmax = (float)pref[3];
    for (int v = lane; v < views.nViews; v += 32)
    {
      // this lane's view: the composed rows in double at the reference point -> float bases; float coefficients
      const double2* tp = reinterpret_cast<const double2*>(views.t2 + v);
      const double2 n0 = __ldg(tp + 0), n1 = __ldg(tp + 1), m0 = __ldg(tp + 2), m1 = __ldg(tp + 3), z0r = __ldg(tp + 4), z1r = __ldg(tp + 5);
      const float bx = __double2float_rn(fma(rx0, n0.x, fma(ry0, n0.y, fma(rz0, n1.x, n1.y))));
      const float by = __double2float_rn(fma(rx0, m0.x, fma(ry0, m0.y, fma(rz0, m1.x, m1.y))));
      const float bz = __double2float_rn(fma(rx0, z0r.x, fma(ry0, z0r.y, fma(rz0, z1r.x, z1r.y))));
      const float4* fp = reinterpret_cast<const float4*>(views.fast + v);
      const float4 rx = __ldg(fp + 0), ry = __ldg(fp + 1), rz = __ldg(fp + 2);
      const float lz = fmaf(rz.w, dmax, fabsf(bz));
      const float Ex = views.kE * (fmaf(rx.w, dmax, fabsf(bx)) + views.U1 * lz);
      const float Ey = views.kE * (fmaf(ry.w, dmax, fabsf(by)) + views.U1 * lz);
      const float zm = views.kZ * lz;
      const size_t vbase = npix * 3 * (size_t)v;
      // ---- phase 1: pixel of each point in this view, or -1
      int idx[Q];
      unsigned slow = 0;
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
        // lane l takes the points in the order l, l + 1, ... (mod Q): at any step the lanes work on different points, so
        // their histogram updates of phase 3 do not pile up on one point's few bins (a surface point looks alike in all views)
        const float4 P = pw[(q + lane) & (Q - 1)];
        const float fz = bz + fmaf(P.x, rz.x, fmaf(P.y, rz.y, P.z * rz.z));
        const float fx = bx + fmaf(P.x, rx.x, fmaf(P.y, rx.y, P.z * rx.z));
        const float fy = by + fmaf(P.x, ry.x, fmaf(P.y, ry.y, P.z * ry.z));
        const float r = rcp_approx_c(fz), ar = fabsf(r);
        const float tu = fmaf(fx, r, kMagicC), tv = fmaf(fy, r, kMagicC);
        const float pu = tu - kMagicC, pv = tv - kMagicC;
        const float eu = fmaf(fx, r, -pu), ev = fmaf(fy, r, -pv);
        const float tx = fmaf(-Ex, ar, T) * P.w, ty = fmaf(-Ey, ar, T);      // P.w = 1, or NaN for an inexact point
        // a NaN threshold (inexact point) or a NaN projection fails every compare
        const bool cert = (fabsf(fz) > zm) && (fabsf(eu) < tx) && (fabsf(ev) < ty);
        const int px = __float_as_int(tu) - pxoff, py = __float_as_int(tv) - pyoff;
        idx[q] = (cert && (unsigned)px < (unsigned)W && (unsigned)py < (unsigned)H) ? (H - 1 - py) * W + px : -1;
        if (!cert) slow |= 1u << q;
      }
      // ---- uncertified (point, view) pairs: FP64 tier, then the reference's own operation sequence; dynamic q
      while (slow)
      {
        const int q = __ffs(slow) - 1;
        slow &= slow - 1;
        const int qp = (q + lane) & (Q - 1);
        const float4 P = pw[qp];
        const bool exact = P.w == P.w;
        const size_t p = perm[min(p0 + qp, nPoints - 1)];
        // the point as the reference sees it (a float point converts exactly)
        const double x = (double)xyz[3 * p + 0], y = (double)xyz[3 * p + 1], z = (double)xyz[3 * p + 2];
        float pu = NAN, pv = NAN;
        if (exact)
        {
          const float fz = bz + fmaf(P.x, rz.x, fmaf(P.y, rz.y, P.z * rz.z));
          const float r = rcp_approx_c(fz);
          pu = fmaf(bx + fmaf(P.x, rx.x, fmaf(P.y, rx.y, P.z * rx.z)), r, kMagicC) - kMagicC;
          pv = fmaf(by + fmaf(P.x, ry.x, fmaf(P.y, ry.y, P.z * ry.z)), r, kMagicC) - kMagicC;
        }
        const float mf = fmaxf(fabsf((float)x), fmaxf(fabsf((float)y), fabsf((float)z)));
        const int id = project_slow(views, v, x, y, z, mf, pu, pv, W, H);
#pragma unroll
        for (int qq = 0; qq < Q; qq++) if (qq == q) idx[qq] = id;
      }
      // ---- phase 2: gathers, all in flight together.  Bytes 3*idx .. 3*idx + 2 of the view = GetColorValue's tuple
      // (ReconstructionData.cxx:107-115), taken from the two aligned words that hold them.
      unsigned w0[Q], w1[Q];
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
        const size_t o = vbase + (size_t)(idx[q] < 0 ? 0 : idx[q]) * 3;
        const size_t a = o >> 2;
        w0[q] = __ldg(words + a);
        w1[q] = __ldg(words + min(a + 1, lastWord));                // the clamp only bites on the buffer's very last word,
      }                                                             // whose triple never reaches into a following word
      // ---- phase 3: histograms
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
        if (idx[q] >= 0)
        {
          const unsigned sh = (unsigned)((vbase + (size_t)idx[q] * 3) & 3) * 8;
          const unsigned rgb = __funnelshift_r(w0[q], w1[q], sh);
          const unsigned cr = rgb & 0xffu, cg = (rgb >> 8) & 0xffu, cb = (rgb >> 16) & 0xffu;
          unsigned* h = hw + ((q + lane) & (Q - 1)) * kHist;
          atomicAdd(h + (cr >> 1), 1u << ((cr & 1) * 16));
          atomicAdd(h + (kBins / 2) + (cg >> 1), 1u << ((cg & 1) * 16));
          atomicAdd(h + kBins + (cb >> 1), 1u << ((cb & 1) * 16));
        }
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int q = 0; q < Q; q++)
    {
      unsigned med[3] = {0u, 0u, 0u}, avg[3] = {0u, 0u, 0u}, an = 0;
#pragma unroll
      for (int ch = 0; ch < 3; ch++)
      {
        unsigned c[8], tot = 0, wsum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          const unsigned wv = hw[q * kHist + ch * (kBins / 2) + lane * 4 + k];
          c[2 * k] = wv & 0xffffu; c[2 * k + 1] = wv >> 16;
          tot += c[2 * k] + c[2 * k + 1];
          wsum += c[2 * k] * (unsigned)(lane * 8 + 2 * k) + c[2 * k + 1] * (unsigned)(lane * 8 + 2 * k + 1);
        }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        an = __shfl_sync(0xffffffffu, incl, 31);                  // the same for the three channels
        if (an > 0)
        {
          const unsigned hiRank = an / 2, loRank = (an % 2 == 0) ? an / 2 - 1 : an / 2;
          const unsigned before = incl - tot;
          const int a = select_rank(c, before, hiRank, lane);
          const int b = select_rank(c, before, loRank, lane);
          med[ch] = (unsigned)(a + b) >> 1;                         // odd n: a == b
          avg[ch] = wsum / an;                                       // int sum / n, truncated (MeshColoration.cxx:176-180)
        }
      }
      if (lane == 0 && p0 + q < nPoints)
      {
        const size_t p = perm[p0 + q];
        // n == 0: arrays keep their zero fill (MeshColoration.cxx:116-118,124-126,132)
        mean[3 * p + 0] = (uint8_t)avg[0]; mean[3 * p + 1] = (uint8_t)avg[1]; mean[3 * p + 2] = (uint8_t)avg[2];
        median[3 * p + 0] = (uint8_t)med[0]; median[3 * p + 1] = (uint8_t)med[1]; median[3 * p + 2] = (uint8_t)med[2];
        nb[p] = (int32_t)an;
      }
    }
    __syncwarp();
  }
}

// ---- spatial order of the points ------------------------------------------------------------------------------------
// The reference walks the mesh in file order; every point is independent (MeshColoration.cxx:140-192), so the kernel
// may take them in any order.  A batch of Q points shares its FP64 reference point and, view by view, neighbouring image
// sectors, so batches should be spatially compact whatever the file order is (a contour's vertex list alternates
// between the front and the back of the surface).  Two-level counting sort by the 30-bit Morton code of a 1024^3 lattice over
// the bounding box: the upper 18 bits (64^3 cells) through global bucket counts, offsets and a scatter; the lower 12 bits
// inside each cell by one CTA in shared memory.  Consecutive points then form small patches of the surface, which project
// onto a few image sectors in every view.  Ties and over-full cells are left in arrival order: results do not depend on it.
constexpr int kSortBits = 6, kSortBuckets = 1 << (3 * kSortBits);
constexpr int kFineBits = 4, kFineBins = 1 << (3 * kFineBits), kFineCap = 5120;

__device__ __forceinline__ unsigned ordered_bits(float f)      // monotone map float -> unsigned
{
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u)
{
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <typename XYZ>
__global__ void __launch_bounds__(256) color_bbox_kernel(size_t nPoints, const XYZ* __restrict__ xyz, unsigned* __restrict__ bbox)
{
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < nPoints; p += (size_t)gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
      const float f = (float)xyz[3 * p + a];
      if (fabsf(f) <= 3.0e38f) { const unsigned u = ordered_bits(f); lo[a] = min(lo[a], u); hi[a] = max(hi[a], u); }
    }
#pragma unroll
  for (int a = 0; a < 3; a++)
  {
    for (int o = 16; o > 0; o >>= 1)
    {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(bbox + a, lo[a]); atomicMax(bbox + 3 + a, hi[a]); }
  }
}

// FINE = false: Morton code of the point's 64^3 cell (18 bits); FINE = true: of its position inside the cell (12 bits)
template <typename XYZ, bool FINE = false>
__device__ __forceinline__ unsigned color_sort_key(const XYZ* __restrict__ xyz, size_t p, const unsigned* __restrict__ bbox)
{
  unsigned key = 0;
  constexpr int kAll = kSortBits + kFineBits;
#pragma unroll
  for (int a = 0; a < 3; a++)
  {
    const float lo = from_ordered_bits(bbox[a]), hi = from_ordered_bits(bbox[3 + a]);
    const float f = (float)xyz[3 * p + a];
    const float t = (f - lo) / fmaxf(hi - lo, 1e-30f) * (float)(1 << kAll);
    unsigned q = (t == t) ? (unsigned)fminf(fmaxf(t, 0.f), (float)((1 << kAll) - 1)) : 0u;      // NaN / inf -> cell 0
    q = FINE ? (q & ((1u << kFineBits) - 1u)) : (q >> kFineBits);
#pragma unroll
    for (int b = 0; b < (FINE ? kFineBits : kSortBits); b++) key |= ((q >> b) & 1u) << (3 * b + a);
  }
  return key;
}

template <typename XYZ>
__global__ void __launch_bounds__(256) color_count_kernel(size_t nPoints, const XYZ* __restrict__ xyz, const unsigned* __restrict__ bbox,
                                                          unsigned* __restrict__ counts)
{
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < nPoints; p += (size_t)gridDim.x * blockDim.x)
    atomicAdd(counts + color_sort_key(xyz, p, bbox), 1u);
}

// exclusive scan of the bucket counts, in place, one block
__global__ void __launch_bounds__(1024) color_scan_kernel(unsigned* __restrict__ counts)
{
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < kSortBuckets; b0 += 1024)
  {
    const unsigned v = counts[b0 + threadIdx.x];
    unsigned incl = v;
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    unsigned before = 0;
    for (int q = 0; q < w; q++) before += s_w[q];
    const unsigned base = s_base;
    counts[b0 + threadIdx.x] = base + before + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_base = base + before + incl;
    __syncthreads();
  }
}

template <typename XYZ>
__global__ void __launch_bounds__(256) color_scatter_kernel(size_t nPoints, const XYZ* __restrict__ xyz, const unsigned* __restrict__ bbox,
                                                            unsigned* __restrict__ cursor, unsigned* __restrict__ perm)
{
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < nPoints; p += (size_t)gridDim.x * blockDim.x)
    perm[atomicAdd(cursor + color_sort_key(xyz, p, bbox), 1u)] = (unsigned)p;
}

// after the scatter `ends[b]` = end of cell b's range of `perm` (its start = ends[b - 1]): order each cell by the fine key
template <typename XYZ>
__global__ void __launch_bounds__(256) color_refine_kernel(const XYZ* __restrict__ xyz, const unsigned* __restrict__ bbox,
                                                           const unsigned* __restrict__ ends, unsigned* __restrict__ perm)
{
  __shared__ unsigned s_idx[kFineCap];
  __shared__ unsigned short s_key[kFineCap];
  __shared__ unsigned s_hist[kFineBins];
  __shared__ unsigned s_w[8];
  for (int b = blockIdx.x; b < kSortBuckets; b += gridDim.x)
  {
    const unsigned start = b ? ends[b - 1] : 0u, n = ends[b] - start;
    if (n < 2 || n > (unsigned)kFineCap) continue;             // CTA-uniform
    for (int q = threadIdx.x; q < kFineBins; q += 256) s_hist[q] = 0u;
    __syncthreads();
    for (unsigned q = threadIdx.x; q < n; q += 256)
    {
      const unsigned p = perm[start + q];
      const unsigned k = color_sort_key<XYZ, true>(xyz, p, bbox);
      s_idx[q] = p; s_key[q] = (unsigned short)k;
      atomicAdd(s_hist + k, 1u);
    }
    __syncthreads();
    // exclusive scan of the 4096 bins: 16 per thread
    unsigned local[kFineBins / 256], sum = 0;
#pragma unroll
    for (int q = 0; q < kFineBins / 256; q++) { local[q] = sum; sum += s_hist[threadIdx.x * (kFineBins / 256) + q]; }
    unsigned incl = sum;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    unsigned before = 0;
    for (int q = 0; q < w; q++) before += s_w[q];
    const unsigned base = before + incl - sum;
#pragma unroll
    for (int q = 0; q < kFineBins / 256; q++) s_hist[threadIdx.x * (kFineBins / 256) + q] = base + local[q];
    __syncthreads();
    for (unsigned q = threadIdx.x; q < n; q += 256) perm[start + atomicAdd(s_hist + s_key[q], 1u)] = s_idx[q];
    __syncthreads();
  }
}

size_t colorize_scratch_bytes(size_t nPoints) { return (size_t)kSortBuckets * 4 + 64 + nPoints * 4; }

// d_scratch: colorize_scratch_bytes(nPoints) bytes
cudaError_t launch_colorize(size_t nPoints, const void* d_xyz, int xyzType, ColorViews views,
                            const uint8_t* d_colors, int W, int H, uint8_t* d_mean, uint8_t* d_median,
                            int32_t* d_nb, void* d_scratch, cudaStream_t s)
{
  if (nPoints == 0) return cudaSuccess;
  if (nPoints >= (1ull << 32)) return cudaErrorInvalidValue;  // 32-bit point permutation
  if (views.nViews >= 65536) return cudaErrorInvalidValue;   // 16-bit histogram counters
  if (reinterpret_cast<uintptr_t>(d_colors) & 3) return cudaErrorMisalignedAddress;
  constexpr int Q = DMI_COLOR_Q;
  static_assert((Q & (Q - 1)) == 0 && Q <= 32, "Q must be a power of two");
  constexpr size_t smem = (size_t)kColorWarps * (8 + Q * 4 + Q * (3 * (kBins / 2) + 1)) * 4;
  const size_t colorBytes = (size_t)views.nViews * W * H * 3;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned* counts = (unsigned*)d_scratch;
  unsigned* bbox = counts + kSortBuckets;
  unsigned* perm = bbox + 16;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)kSortBuckets * 4, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(bbox, 0xff, 12, s);                      // lows = 0xffffffff
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(bbox + 3, 0, 12, s);                     // highs = 0
  if (e != cudaSuccess) return e;
  const unsigned sblocks = (unsigned)std::min<size_t>((nPoints + 255) / 256, (size_t)sms * 8);
  const size_t perSm = std::max<size_t>(1, (size_t)(227 * 1024) / (smem + 1024));
  size_t blocks = ((nPoints + Q - 1) / Q + kColorWarps - 1) / kColorWarps;
  blocks = std::min(blocks, (size_t)sms * perSm);            // persistent: CTAs stride over the point batches
#define DMI_COLOR_LAUNCH(XYZ)                                                                                          \
  {                                                                                                                    \
    const XYZ* x = (const XYZ*)d_xyz;                                                                                  \
    color_bbox_kernel<XYZ><<<sblocks, 256, 0, s>>>(nPoints, x, bbox);                                                  \
    color_count_kernel<XYZ><<<sblocks, 256, 0, s>>>(nPoints, x, bbox, counts);                                         \
    color_scan_kernel<<<1, 1024, 0, s>>>(counts);                                                                      \
    color_scatter_kernel<XYZ><<<sblocks, 256, 0, s>>>(nPoints, x, bbox, counts, perm);                                 \
    color_refine_kernel<XYZ><<<(unsigned)sms * 4, 256, 0, s>>>(x, bbox, counts, perm);                                 \
    e = cudaFuncSetAttribute(colorize_kernel<XYZ, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    if (e != cudaSuccess) return e;                                                                                    \
    colorize_kernel<XYZ, Q><<<(unsigned)blocks, 32 * kColorWarps, smem, s>>>(nPoints, x, perm, views, d_colors, colorBytes, W, H, \
                                                                             d_mean, d_median, d_nb);                  \
  }
  if (xyzType == 1) DMI_COLOR_LAUNCH(double) else DMI_COLOR_LAUNCH(float)
#undef DMI_COLOR_LAUNCH
  return cudaGetLastError();
}

// ---- host side: composition of one view ---------------------------------------------------------

static float upf(double x)
{
  float f = (float)x;
  if ((double)f < x) f = nextafterf(f, INFINITY);
  return f;
}

static float upf(double x);

// FP32 tier of the coloration: f = fl(fl(b) + local), b = the row at the batch's reference point evaluated in double,
// local = three float FMAs over the offsets (|offset| <= d).  |f - true| <= 7 * 2^-24 * (|b| + A d): one rounding of the base,
// one of the final add, one of each offset (a difference of floats), three FMA roundings and the coefficient roundings,
// with 1 % of slack.  E = (4/3) (delta_n + U1 delta_z); |fz| > zm must imply delta_z / |fz| <= 1/4.
void color_bound_constants(int W, int H, float* kE, float* kZ, float* U1)
{
  *kE = upf((4.0 / 3.0) * 7.0 * 1.01 * std::ldexp(1.0, -24));
  *kZ = upf(4.0 * 7.0 * 1.01 * std::ldexp(1.0, -24));
  *U1 = upf(std::max(W, H) / 2.0 + 2.0);
}

float color_threshold_T(int W, int H)
{
  const double U1 = std::max(W, H) / 2.0 + 2.0;
  const double c0 = (4.0 / 3.0) * 1.05 * std::ldexp(1.0, -23) * U1 + std::ldexp(1.0, -20);
  float T = (float)(0.5 - c0);
  if ((double)T > 0.5 - c0) T = nextafterf(T, -INFINITY);
  return T;
}

// D = K3 * RT(3x4) in long double; rows nx = D0 - cxc*D2, ny = D1 - cyc*D2, dz = D2 over (x, y, z, 1).
void compose_color_view(const double* K16, const double* RT16, int cxc, int cyc, int W, int H,
                        ColorViewFast* fo, ColorViewT2* t2)
{
  typedef long double L;
  L D[3][4];
  double Aabs[3] = {0, 0, 0}, Babs[3] = {0, 0, 0};        // magnitudes the REFERENCE's evaluation goes through
  for (int r = 0; r < 3; r++)
  {
    for (int j = 0; j < 4; j++)
    {
      D[r][j] = 0;
      for (int c = 0; c < 3; c++) D[r][j] += (L)K16[4 * r + c] * (L)RT16[4 * c + j];
    }
    for (int c = 0; c < 3; c++)
    {
      Aabs[r] += std::fabs(K16[4 * r + c]) * (std::fabs(RT16[4 * c + 0]) + std::fabs(RT16[4 * c + 1]) + std::fabs(RT16[4 * c + 2]));
      Babs[r] += std::fabs(K16[4 * r + c]) * std::fabs(RT16[4 * c + 3]);
    }
  }
  const L cc[2] = {(L)cxc, (L)cyc};
  double* rows[2] = {t2->nx, t2->ny};
  for (int r = 0; r < 2; r++)
    for (int j = 0; j < 4; j++) rows[r][j] = (double)(D[r][j] - cc[r] * D[2][j]);
  for (int j = 0; j < 4; j++) t2->dz[j] = (double)D[2][j];
  auto A = [](const double* r) { return std::fabs(r[0]) + std::fabs(r[1]) + std::fabs(r[2]); };
  for (int j = 0; j < 3; j++) { fo->nx[j] = (float)t2->nx[j]; fo->ny[j] = (float)t2->ny[j]; fo->dz[j] = (float)t2->dz[j]; }
  // [3]: sum of the |coefficients|, rounded up (bounds the local terms A * d of the FP32 tier)
  fo->nx[3] = upf(A(t2->nx) * (1.0 + 1e-6)); fo->ny[3] = upf(A(t2->ny) * (1.0 + 1e-6)); fo->dz[3] = upf(A(t2->dz) * (1.0 + 1e-6));
  const double U1 = std::max(W, H) / 2.0 + 2.0;
  // T2 margins: 2^-44 of the reference's intermediate magnitudes (|K3| * |RT| sums), composed rows included
  const double k2 = std::ldexp(1.0, -44);
  const double Ax = Aabs[0] + std::fabs((double)cxc) * Aabs[2], Bx = Babs[0] + std::fabs((double)cxc) * Babs[2];
  const double Ay = Aabs[1] + std::fabs((double)cyc) * Aabs[2], By = Babs[1] + std::fabs((double)cyc) * Babs[2];
  t2->m2a = k2 * (std::max(Ax, Ay) + U1 * Aabs[2]);
  t2->m2b = k2 * (std::max(Bx, By) + U1 * Babs[2]);
  t2->mza = k2 * Aabs[2];
  t2->mzb = k2 * Babs[2] + 1e-300;
}

}  // namespace dmi
