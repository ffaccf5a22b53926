// Certified fast path of the TSDF integration for sm_100a.
//
// The reference evaluates ~78 FP64 operations per voxel*view (CudaReconstruction.cu:158-212), most
// of them to find WHICH pixel the voxel centre rounds to and WHICH branch of the ray potential
// applies.  Those answers are discrete, so they can be found cheaply and then PROVEN: this kernel
// keeps every discrete decision of the reference (behind-camera test :177, pixel rounding :187-188,
// bounds :192-197, depth == -1 :202 after the best-cost filter, the potential's branches :114-119)
// bit-for-bit; only the continuous values (camera z, the linear branch) differ from the reference's
// rounding by a few ulp of double.  Error bounds are derived in DESIGN.md ("certification").
//
// Work item: a brick of 16 x 8 x 8 voxels (4 warps of 8 x 4 lanes, 8 voxels per thread along k) against a chunk of
// up to 64 views.  A persistent grid (5 CTAs per SM) takes the bricks of the supertiles (64 x 32 x 32 voxels) that
// some view of the chunk may touch from a global counter.
//   supertile  one thread per (supertile, view): the exact box tests below on the supertile's box -> 64-bit view
//              masks, compacted into the list of active supertiles
//   pre-pass   one thread per view the supertile kept: FP64 brick base of the projective rows -> float; projected
//              bounding box of the brick; CULL the view when every voxel is behind the camera, outside the image,
//              over tiles without a valid pixel, or farther than Delta behind every valid depth of its
//              footprint (all of which contribute exactly nothing); FREE SPACE when every voxel lands on fully
//              valid tiles farther than Delta in front of every depth there (the view adds -Eta*Rho to every
//              voxel: one add each); per-brick error scales; survivors are compacted in view order into shared
//              memory.  Footprint statistics come from per-view sparse tables (4 loads, see dmi_internal.cuh).
//   phase A    T1, FP32: 3 FFMA + MUFU.RCP + 2 FFMA give the centred pixel rounded by the magic-number
//              add; 2 more FFMA its distance to the integer; certified when that distance is below
//              0.5 - (E*r + c0).  Uncertified voxels (~0.06 %) go to
//              T2, FP64: composed rows (9 DFMA) + residual test of the candidate and its neighbours, and
//              T3, the reference's own operation sequence, when T2 meets a tie within 2^-44.
//   phase B    gather of the FLOAT classification image (4 B per voxel*view)
//   phase C    FP32: invalid / farther than Delta (certified with margin) -> add 0 or -Eta*Rho; the thin
//              band around the surface re-gathers the double depth (or rebuilds it from the classification
//              float and the int32 residual image) and evaluates the potential in FP64 (T3 when |diff| is within
//              2^-44 of Delta, the potential's only discontinuity).
// Views are accumulated in list order per voxel, like the reference's host loop (:343).
#include "dmi_internal.cuh"
#include "tsdf_device.cuh"

#include <cmath>
#include <climits>
#include <algorithm>

#ifndef DMI_FAST_CTAS
#define DMI_FAST_CTAS 6          // CTAs of the integration kernel per SM (register cap 65536 / (128 * CTAs)); measured on config 5,
                                 // ms per step at (CTAs, halves) = (4,2) 233, (5,2) 224, (6,2) 257, (5,1) 223, (6,1) 216
#endif
#ifndef DMI_FAST_HALVES
#define DMI_FAST_HALVES 1        // 2: gathers of 4 voxels in flight while the other 4 are classified; 1: all 8 at once
#endif

namespace dmi {

constexpr int FM = 8;                       // voxels per thread, consecutive k
constexpr int FBI = 16, FBJ = 8;            // brick = 16 x 8 x FM voxels = 4 warps of 8 x 4 lanes
constexpr int FT = 128;
constexpr int FSI = 4, FSJ = 4, FSK = 4;    // supertile = 64 x 32 x 32 voxels, enumerated contiguously
constexpr float kMagic = 12582912.0f;       // 1.5 * 2^23: adding it rounds |x| < 2^22 to an integer
constexpr int kMagicBits = 0x4B400000;
constexpr int kTile = 8;                    // tile edge of the per-view tile statistics

constexpr int kBad = INT_MIN + 2;           // every valid gather index (px - py*W) is above this
constexpr int kReject = kBad - 1;           // certified: this voxel*view contributes nothing
constexpr int kNeedExact = kBad - 2;        // could not certify: run T3

__device__ __forceinline__ float rcp_approx(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// Phase C for one voxel: acc += ner when the voxel is valid, certainly farther than Delta from the depth
// and in front of it; sets `bit` in `near` when it is valid but not certainly far.  Written in PTX so
// that the add is PREDICATED (the compiler otherwise adds unconditionally and selects: 6 extra FSELs).
__device__ __forceinline__ void classify_far(double& acc, unsigned& near, double ner, float fc, float d32, float thr, unsigned bit)
{
  asm("{\n\t"
      ".reg .pred pv, pf, pa, pn;\n\t"
      ".reg .f32 df, adf;\n\t"
      ".reg .b32 shi, zero;\n\t"
      ".reg .f64 sel;\n\t"
      "mov.b32 zero, 0;\n\t"
      "sub.rn.f32 df, %3, %4;\n\t"
      "abs.f32 adf, df;\n\t"
      "setp.neu.f32 pv, %4, 0fBF800000;\n\t"
      "setp.gt.and.f32 pf, adf, %5, pv;\n\t"
      "setp.lt.and.f32 pa, df, 0f00000000, pf;\n\t"
      "selp.b32 shi, 0x3FF00000, 0, pa;\n\t"          // 1.0 or 0.0 as a double: one select, then one DFMA
      "mov.b64 sel, {zero, shi};\n\t"                  // (ptxas turns a predicated DADD into DADD + 2 FSEL)
      "fma.rn.f64 %0, %2, sel, %0;\n\t"
      "xor.pred pn, pv, pf;\n\t"
      "@pn or.b32 %1, %1, %6;\n\t"
      "}" : "+d"(acc), "+r"(near) : "d"(ner), "f"(fc), "f"(d32), "f"(thr), "r"(bit));
}
__device__ __forceinline__ void classify_far(float& acc, unsigned& near, float ner, float fc, float d32, float thr, unsigned bit)
{
  asm("{\n\t"
      ".reg .pred pv, pf, pa, pn;\n\t"
      ".reg .f32 df, adf;\n\t"
      "sub.rn.f32 df, %3, %4;\n\t"
      "abs.f32 adf, df;\n\t"
      "setp.neu.f32 pv, %4, 0fBF800000;\n\t"
      "setp.gt.and.f32 pf, adf, %5, pv;\n\t"
      "setp.lt.and.f32 pa, df, 0f00000000, pf;\n\t"
      "selp.f32 df, 0f3F800000, 0f00000000, pa;\n\t"
      "fma.rn.f32 %0, %2, df, %0;\n\t"
      "xor.pred pn, pv, pf;\n\t"
      "@pn or.b32 %1, %1, %6;\n\t"
      "}" : "+f"(acc), "+r"(near) : "f"(ner), "f"(fc), "f"(d32), "f"(thr), "r"(bit));
}

// Phase C of a view for which the whole brick is certainly farther than Delta in front of every valid depth it can
// meet: acc += ner when the gathered pixel is valid (anything but -1.0f), else nothing.
__device__ __forceinline__ void add_if_valid(double& acc, double ner, float d32)
{
  asm("{\n\t"
      ".reg .pred pv;\n\t"
      ".reg .b32 shi, zero;\n\t"
      ".reg .f64 sel;\n\t"
      "mov.b32 zero, 0;\n\t"
      "setp.neu.f32 pv, %2, 0fBF800000;\n\t"
      "selp.b32 shi, 0x3FF00000, 0, pv;\n\t"
      "mov.b64 sel, {zero, shi};\n\t"
      "fma.rn.f64 %0, %1, sel, %0;\n\t"
      "}" : "+d"(acc) : "d"(ner), "f"(d32));
}
__device__ __forceinline__ void add_if_valid(float& acc, float ner, float d32)
{
  asm("{\n\t"
      ".reg .pred pv;\n\t"
      ".reg .f32 sel;\n\t"
      "setp.neu.f32 pv, %2, 0fBF800000;\n\t"
      "selp.f32 sel, 0f3F800000, 0f00000000, pv;\n\t"
      "fma.rn.f32 %0, %1, sel, %0;\n\t"
      "}" : "+f"(acc) : "f"(ner), "f"(d32));
}

__device__ __forceinline__ double affine(const double* r, double di, double dj, double dk)
{
  return fma(di, r[0], fma(dj, r[1], fma(dk, r[2], r[3])));
}

// T3: the whole voxel*view update exactly as the reference does it.  By value in and out, so the
// caller's accumulators stay in registers.
template <typename T>
__device__ __noinline__ T exact_unit(const GridParams* g, const ViewExact* e, const double* depth,
                                     const float* cls, const int* lo, int i, int j, int k, T acc)
{
  double wx, wy, wz;
  voxel_world(*g, i, j, k, wx, wy, wz);
  integrate_exact<T>(*g, *e, depth, wx, wy, wz, acc, cls, lo);
  return acc;
}

// T2: FP64 certification of the T1 candidate (pu, pv: centred integer pixel as float).
// Returns the gather index px - py*W (relative to the last storage row), kReject or kNeedExact.
__device__ __noinline__ int tier2(const ViewFast& V, int cxc, int cyc, int W, int H,
                                  double di, double dj, double dk, float pu, float pv)
{
  const double hz = affine(V.hz, di, dj, dk);
  if (!(hz > V.m2z)) return (hz < -V.m2z) ? kReject : kNeedExact;            // NaN -> exact
  const double half = 0.5 * hz;
  const double m2 = V.m2;
  if (!(m2 < 0.01 * half)) return kNeedExact;
  // the candidate is garbage when T1 overflowed; beyond 2^22 let the exact tier decide
  double cu = (double)pu, cv = (double)pv;
  if (!(fabs(cu) < 4194304.0 && fabs(cv) < 4194304.0)) return kNeedExact;
  double su = fma(-cu, hz, affine(V.nx, di, dj, dk));                         // (u' - cu) * hz
  double sv = fma(-cv, hz, affine(V.ny, di, dj, dk));
  if (su >= half + m2) { cu += 1.0; su -= hz; } else if (su <= -half - m2) { cu -= 1.0; su += hz; }
  if (sv >= half + m2) { cv += 1.0; sv -= hz; } else if (sv <= -half - m2) { cv -= 1.0; sv += hz; }
  if (!(fabs(su) < half - m2 && fabs(sv) < half - m2)) return kNeedExact;     // tie, or T1 off by more than one
  const int px = (int)cu + cxc, py = (int)cv + cyc;
  if ((unsigned)px < (unsigned)W && (unsigned)py < (unsigned)H) return px - py * W;
  return kReject;
}

// Per (brick, surviving view) record in shared memory, written by the pre-pass.
struct __align__(16) ViewSm
{
  float4 base;      // brick base of nx, ny, hz (float); brick base of camera z (non-pinhole only)
  float tcert;      // FP32 tier: certified iff max(|eu|, |ev|) < tcert = min(Tx - Ex*r, Ty - Ey*r) at the brick's largest r
                    // (-1 when the box's projection cannot be trusted, mode 0: every voxel then goes to the FP64 tier)
  float thrfar;     // far threshold (Delta + margin)
  int view;
  int flags;        // bits 0-1: front (see BoxEval); bits 2-3: mode
  long long voff;   // the view's last storage row in the classification / depth / residual images: npix*v + (H-1)*W
  int rej;          // gather index (relative to voff) of the spare -1.0f slot
  int pad;
};
// Per view of the launch, filled once per CTA from the kernel parameters.
struct __align__(16) ViewConst
{
  float4 cx, cy, cz, cc;   // fnx[0..2], fny[0..2], fhz[0..2], fcz[0..2] (w unused)
  double czr[4];           // camera-z row (FP64)
  double gd;
  float zm, pad;
};

struct BrickBox { bool valid, inside; float ux, uy; int tx0, tx1, ty0, ty1; };

// Projected bounding box of a box of voxels for one view (8 corners; the projection of a box with hz > 0
// is the convex hull of its projected corners).  (ei, ej, ek) = extents - 1 in voxels; (lx, ly, lz) bound
// the local terms.  Returns false when the box cannot be trusted (too close to the camera plane).
// Margins: 1.5 px + the FP32 evaluation error (<= 0.25 px inside the image by the E*r test,
// proportional to |u| outside).
__device__ __forceinline__ bool brick_box(const ViewFast& V, const FastChunk& c, float fbx, float fby, float fbz,
                                          float ei, float ej, float ek, float lx, float ly, float lz,
                                          float zlo, int W, int H, BrickBox& o, bool& outside)
{
  outside = false;
  const float Eg = c.k3 * (fmaxf(fabsf(fbx) + 3.f * lx, fabsf(fby) + 3.f * ly) + c.umax1g * (fabsf(fbz) + 3.f * lz));
  if (!(zlo > V.zm) || !(Eg * rcp_approx(zlo) <= 0.24f)) return false;
  const float zi = V.fhz[0] * ei, zj = V.fhz[1] * ej, zk = V.fhz[2] * ek;
  const float xi = V.fnx[0] * ei, xj = V.fnx[1] * ej, xk = V.fnx[2] * ek;
  const float yi = V.fny[0] * ei, yj = V.fny[1] * ej, yk = V.fny[2] * ek;
  float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
  for (int q = 0; q < 8; q++)
  {
    const float hz = fbz + ((q & 1) ? zi : 0.f) + ((q & 2) ? zj : 0.f) + ((q & 4) ? zk : 0.f);
    const float r = rcp_approx(hz);
    const float u = (fbx + ((q & 1) ? xi : 0.f) + ((q & 2) ? xj : 0.f) + ((q & 4) ? xk : 0.f)) * r;
    const float w = (fby + ((q & 1) ? yi : 0.f) + ((q & 2) ? yj : 0.f) + ((q & 4) ? yk : 0.f)) * r;
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, w); vmax = fmaxf(vmax, w);
  }
  if (!(umin == umin && umax == umax && vmin == vmin && vmax == vmax)) return false;    // NaN
  const float kr = 0.3f / c.umax1g;
  const float ulo = umin - (1.5f + kr * fabsf(umin)), uhi = umax + (1.5f + kr * fabsf(umax));
  const float vlo = vmin - (1.5f + kr * fabsf(vmin)), vhi = vmax + (1.5f + kr * fabsf(vmax));
  o.ux = fmaxf(fabsf(ulo), fabsf(uhi));
  o.uy = fmaxf(fabsf(vlo), fabsf(vhi));
  const float xlo = ulo + (float)c.cxc, xhi = uhi + (float)c.cxc, ylo = vlo + (float)c.cyc, yhi = vhi + (float)c.cyc;
  if (!(xhi >= 0.f && xlo <= (float)(W - 1) && yhi >= 0.f && ylo <= (float)(H - 1))) { outside = true; return true; }
  o.inside = xlo >= 0.f && xhi <= (float)(W - 1) && ylo >= 0.f && yhi <= (float)(H - 1);   // every voxel lands in the image
  const int px0 = max(0, (int)floorf(xlo)), px1 = min(W - 1, (int)ceilf(xhi));
  const int py0 = max(0, (int)floorf(ylo)), py1 = min(H - 1, (int)ceilf(yhi));
  // pixel bounds in STORAGE coordinates (rows are bottom-up: row = H-1-py)
  o.tx0 = px0; o.tx1 = px1; o.ty0 = H - 1 - py1; o.ty1 = H - 1 - py0;
  o.valid = true;
  return true;
}

// Statistics of the depths over a pixel rectangle (storage coordinates) from the view's sparse tables: four
// overlapping windows of 2^l x 2^l tiles, l = floor(log2(longer side in tiles)); the four loads of a table are
// in flight together.  {max, min} of the valid depths come in one 8-byte entry.
struct FootIdx { int o00, o01, o10, o11; };
__device__ __forceinline__ FootIdx footprint_index(const TilePyramid& pyr, const BrickBox& b)
{
  const int X0 = b.tx0 >> 3, X1 = min(b.tx1 >> 3, pyr.tw - 1), Y0 = b.ty0 >> 3, Y1 = min(b.ty1 >> 3, pyr.th - 1);
  const int n = max(X1 - X0, Y1 - Y0) + 1;
  const int l = min(31 - __clz(n), pyr.nLevels - 1);          // nLevels covers every n <= max(tw, th)
  const int sdim = 1 << l;
  const int xb = max(X0, X1 - sdim + 1), yb = max(Y0, Y1 - sdim + 1);
  const int base = l * (pyr.tw * pyr.th);
  FootIdx ix;
  ix.o00 = base + Y0 * pyr.tw + X0; ix.o01 = base + Y0 * pyr.tw + xb;
  ix.o10 = base + yb * pyr.tw + X0; ix.o11 = base + yb * pyr.tw + xb;
  return ix;
}
// x = max of the valid depths (-inf: none, +inf: a NaN), y = min of the valid depths (+inf: none, -inf: a NaN)
__device__ __forceinline__ float2 footprint_minmax(const float* __restrict__ td, const FootIdx& ix)
{
  const float2* t = reinterpret_cast<const float2*>(td);
  const float2 a = __ldg(t + ix.o00), b = __ldg(t + ix.o01), c = __ldg(t + ix.o10), d = __ldg(t + ix.o11);
  return make_float2(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fminf(fminf(a.y, b.y), fminf(c.y, d.y)));
}
// 0 when every pixel of the rectangle's windows is valid and not NaN
__device__ __forceinline__ float footprint_bad(const TilePyramid& pyr, const float* __restrict__ td, const FootIdx& ix)
{
  const float* t = td + pyr.badOff;
  return fmaxf(fmaxf(__ldg(t + ix.o00), __ldg(t + ix.o01)), fmaxf(__ldg(t + ix.o10), __ldg(t + ix.o11)));
}

// Everything the pre-pass needs to know about (box of voxels, view): FP64 base of the rows at the box
// origin rounded to float, |u| bounds over the box, and whether the view can be CULLED for the whole box:
// every voxel behind the camera (:177), outside the image (:192-197), over tiles without a valid pixel
// (:202), or farther than Delta BEHIND every valid depth it can meet (rayPotential returns 0, :114-115).
// `front` = 2: the whole box lies farther than Delta IN FRONT of every valid depth it can meet: a voxel's
// contribution is -Eta*Rho when its pixel is valid and nothing otherwise (:114-115, :202), so phase C only
// looks at validity.  `front` = 1: moreover every voxel lands inside the image on a valid pixel:
// the view adds exactly -Eta*Rho to every voxel, no projection needed.
// mode 0: nothing known about the box's projection: per-voxel certification, z test and bounds test;
// mode 1: the projected box is trusted (every voxel has h.z > zm, FP32 error <= 0.24 px at the smallest h.z): one
//         brick-wide certification threshold, no z test; mode 2: moreover every voxel lands inside the image: no bounds test.
struct BoxEval { bool keep; int front, mode; float fbx, fby, fbz, fbc, Ux, Uy, czmaxabs, rmax; };

template <bool PINHOLE>
__device__ __forceinline__ BoxEval eval_box(const ViewFast& V, const FastChunk& c, const TilePyramid& pyr,
                                            const float* __restrict__ td, int i0, int j0, int k0,
                                            float ei, float ej, float ek, float lx, float ly, float lz, float lc,
                                            int W, int H, bool cull)
{
  BoxEval o;
  o.fbx = __double2float_rn(affine(V.nx, (double)i0, (double)j0, (double)k0));
  o.fby = __double2float_rn(affine(V.ny, (double)i0, (double)j0, (double)k0));
  o.fbz = __double2float_rn(affine(V.hz, (double)i0, (double)j0, (double)k0));
  o.fbc = PINHOLE ? o.fbz : __double2float_rn(affine(V.cz, (double)i0, (double)j0, (double)k0));
  o.Ux = c.umax1g; o.Uy = c.umax1g;
  // range of h.z and of camera z over the box (affine: extremes are sums of per-axis extremes)
  const float zi = V.fhz[0] * ei, zj = V.fhz[1] * ej, zk = V.fhz[2] * ek;
  const float zslack = 4e-7f * (fabsf(o.fbz) + 3.f * lz);
  const float zlo = o.fbz + fminf(zi, 0.f) + fminf(zj, 0.f) + fminf(zk, 0.f) - zslack;
  const float zhi = o.fbz + fmaxf(zi, 0.f) + fmaxf(zj, 0.f) + fmaxf(zk, 0.f) + zslack;
  float clo = zlo, chi = zhi;
  if (!PINHOLE)
  {
    const float ci = V.fcz[0] * ei, cj = V.fcz[1] * ej, ck = V.fcz[2] * ek;
    const float cslack = 4e-7f * (fabsf(o.fbc) + 3.f * lc);
    clo = o.fbc + fminf(ci, 0.f) + fminf(cj, 0.f) + fminf(ck, 0.f) - cslack;
    chi = o.fbc + fmaxf(ci, 0.f) + fmaxf(cj, 0.f) + fmaxf(ck, 0.f) + cslack;
  }
  o.czmaxabs = fmaxf(fabsf(clo), fabsf(chi));
  o.keep = true;
  o.front = 0;
  o.mode = 0;
  o.rmax = 0.f;
  BrickBox box; box.valid = false; box.inside = false;
  bool outside = false;
  const bool boxed = brick_box(V, c, o.fbx, o.fby, o.fbz, ei, ej, ek, lx, ly, lz, zlo, W, H, box, outside);
  if (boxed)
  {
    o.Ux = fminf(c.umax1g, box.ux + 1.f); o.Uy = fminf(c.umax1g, box.uy + 1.f);
    o.mode = (box.valid && box.inside) ? 2 : 1;
    o.rmax = rcp_approx(zlo) * 1.000001f;                     // >= 1 / f_z of every voxel (zlo carries the evaluation slack)
  }
  if (cull)
  {
    if (zhi < -V.zm) o.keep = false;
    else if (boxed && outside) o.keep = false;
    else if (boxed && box.valid)
    {
      const FootIdx ix = footprint_index(pyr, box);
      const float2 mm = footprint_minmax(td, ix);
      const float thr = c.delta_up + 1e-6f * o.czmaxabs;
      if (mm.x == -INFINITY || clo - mm.x > thr) o.keep = false;
      else if (mm.y - chi > thr)                               // a NaN in the footprint makes mm.y -inf
        o.front = (box.inside && V.pad[0] != 0.f && footprint_bad(pyr, td, ix) == 0.f) ? 1 : 2;
    }
  }
  return o;
}

// First culling level: one thread per (supertile of 64 x 32 x 32 voxels, view); bit v of masks[st] = view v
// may contribute to supertile st.  A brick only examines the views its supertile kept.
template <bool PINHOLE>
__global__ void __launch_bounds__(256)
supertile_cull_kernel(const __grid_constant__ GridParams g, const __grid_constant__ FastChunk c,
                      const float* __restrict__ tileDmax, const __grid_constant__ TilePyramid pyr,
                      const ViewFast* __restrict__ gviews, int nbi, int nbj, int nbk, unsigned* __restrict__ masks, int nst)
{
  static_assert(kFastChunk == 64, "two ballots per supertile");
  const int st = blockIdx.x * 4 + (threadIdx.x >> 6);
  const int v = threadIdx.x & 63;
  bool keep = false;
  if (st < nst && v < c.n)
  {
    const unsigned nsi = (nbi + FSI - 1) / FSI, nsj = (nbj + FSJ - 1) / FSJ;
    const int i0 = (st % nsi) * FSI * FBI, j0 = ((st / nsi) % nsj) * FSJ * FBJ, k0 = slab_global_k(g, (st / (nsi * nsj)) * FSK * FM);
    const ViewFast& V = gviews[v];
    const float ei = (float)(FSI * FBI - 1), ej = (float)(FSJ * FBJ - 1), ek = (float)(FSK * FM - 1);
    const float lx = 1.001f * (fabsf(V.fnx[0]) * ei + fabsf(V.fnx[1]) * ej + fabsf(V.fnx[2]) * ek);
    const float ly = 1.001f * (fabsf(V.fny[0]) * ei + fabsf(V.fny[1]) * ej + fabsf(V.fny[2]) * ek);
    const float lz = 1.001f * (fabsf(V.fhz[0]) * ei + fabsf(V.fhz[1]) * ej + fabsf(V.fhz[2]) * ek);
    const float lc = 1.001f * (fabsf(V.fcz[0]) * ei + fabsf(V.fcz[1]) * ej + fabsf(V.fcz[2]) * ek);
    keep = eval_box<PINHOLE>(V, c, pyr, tileDmax + (size_t)v * pyr.perView, i0, j0, k0, ei, ej, ek, lx, ly, lz, lc,
                             g.W, g.H, true).keep;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0 && st < nst) masks[2 * st + ((threadIdx.x >> 5) & 1)] = bal;
}

__device__ __forceinline__ int nth_set_bit64(unsigned lo, unsigned hi, int n)
{
  const int nlo = __popc(lo);
  return n < nlo ? (int)__fns(lo, 0, n + 1) : 32 + (int)__fns(hi, 0, n - nlo + 1);
}

// List of the supertiles some view of the chunk may touch (all of them when useMasks == 0), and the work counter of the
// persistent integration kernel: work[0] = number of active supertiles, work[1] = 0.  The list is ordered by the number
// of views that may touch the supertile, most first (a counting sort over 0..64): the expensive columns near the
// surface are handed out first and the launch ends on cheap ones, so its tail is short even when a launch holds only a
// few expensive columns per CTA (a z-layer share of the grid, a short view group).  The order inside a class is
// whatever the shared-memory atomics give; the result does not depend on it (work items own disjoint voxels).
__global__ void __launch_bounds__(1024)
compact_supertiles_kernel(const unsigned* __restrict__ masks, int nst, int useMasks, int* __restrict__ list, int* __restrict__ work)
{
  __shared__ int s_bin[kFastChunk + 1];
  if (threadIdx.x <= kFastChunk) s_bin[threadIdx.x] = 0;
  __syncthreads();
  for (int st = threadIdx.x; st < nst; st += 1024)
  {
    const int n = useMasks ? __popc(masks[2 * st]) + __popc(masks[2 * st + 1]) : kFastChunk;
    if (n) atomicAdd(&s_bin[n], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int pos = 0;
    for (int n = kFastChunk; n >= 1; n--) { const int h = s_bin[n]; s_bin[n] = pos; pos += h; }
    work[0] = pos; work[1] = 0;
  }
  __syncthreads();
  for (int st = threadIdx.x; st < nst; st += 1024)
  {
    const int n = useMasks ? __popc(masks[2 * st]) + __popc(masks[2 * st + 1]) : kFastChunk;
    if (n) list[atomicAdd(&s_bin[n], 1)] = st;
  }
}

// One view against the FM voxels of a thread.  INSIDE: every voxel of the brick lands inside the image (mode 2): no bounds
// test.  front2 (CTA-uniform): the whole brick is certainly farther than Delta in front of every valid depth it can meet,
// so phase C only looks at the pixel's validity.
template <typename T, bool PINHOLE, bool COUNT, bool SPLIT, bool INSIDE>
__device__ __forceinline__ void
integrate_view(const GridParams& g, const FastChunk& c, const ViewSm& S, const ViewConst& VC, const double* __restrict__ depths,
               const int* __restrict__ lo, const float* __restrict__ cls, size_t npix, float fli, float flj,
               double di, double dj, double dk0, int i, int j, int k0, T nerT, bool front2, T (&acc)[FM], unsigned long long (&cnt)[10])
{
  const int W = g.W, H = g.H;
  const float4 b = S.base, cx = VC.cx, cy = VC.cy, cz = VC.cz;
  // one rounding at base magnitude here, one in the per-voxel FFMA (delta_n = 3 * 2^-24 * ...)
  const float fnx0 = b.x + fmaf(fli, cx.x, flj * cx.y);
  const float fny0 = b.y + fmaf(fli, cy.x, flj * cy.y);
  const float fhz0 = b.z + fmaf(fli, cz.x, flj * cz.y);
  float fcz0 = 0.f, kc = 0.f;
  if (!PINHOLE) { const float4 cc = VC.cc; fcz0 = b.w + fmaf(fli, cc.x, flj * cc.y); kc = cc.z; }
  const float kx = cx.z, ky = cy.z, kz = cz.z;
  const float thrfar = S.thrfar, tcert = S.tcert;
  // storage row (H-1-py) of the bottom-up image (CudaReconstruction.cu:141-149): index = px - py*W from there
  const long long voff = S.voff;
  const float* cv = cls + voff;
  asm volatile("" : "+l"(cv));                                // keep it as a plain 64-bit register
  // rejected voxels gather the spare float behind the classification images, which holds -1.0f:
  // no predicate, no default value, one sector for the whole warp
  const int rej = S.rej;
  // px - py*W = bits(tu) - W*bits(tv) - (pxoff - W*pyoff), all modulo 2^32 (the true index fits an int)
  const int negW = -W;
  const int pxoff = kMagicBits - c.cxc, pyoff = kMagicBits - c.cyc;
  const int ioff = pxoff + pyoff * negW;

  // Two halves of FM/2 voxels, software-pipelined: A(0) B(0) A(1) B(1) C(0) C(1), so that the gathers
  // of one half are in flight while the other half is classified.
  constexpr int NH = DMI_FAST_HALVES, HM = FM / NH;
  int idx[FM];
  float d32[FM];
  unsigned need3 = 0;
  bool anyh[2] = {false, false};
#pragma unroll
  for (int h = 0; h < NH; h++)
  {
    // ---- phase A: classify HM voxels with the FP32 tier
    unsigned need = 0;
    bool anyv = false;
#pragma unroll
    for (int mm = 0; mm < HM; mm++)
    {
      const int m = h * HM + mm;
      const float fz = fmaf((float)m, kz, fhz0);
      const float fx = fmaf((float)m, kx, fnx0);
      const float fy = fmaf((float)m, ky, fny0);
      const float r = rcp_approx(fz);
      const float tu = fmaf(fx, r, kMagic), tv = fmaf(fy, r, kMagic);      // centred pixel, rounded to integer
      const float pu = tu - kMagic, pv = tv - kMagic;
      const float eu = fmaf(fx, r, -pu), ev = fmaf(fy, r, -pv);            // distance to that integer
      // a NaN (only possible where the box is not trusted: tcert = -1) fails the compare or is dropped by the max: either
      // way nothing is certified against a negative threshold
      const bool cert = fmaxf(fabsf(eu), fabsf(ev)) < tcert;
      bool ok;
      if (INSIDE)
      {
        ok = cert;
        idx[m] = ok ? (__float_as_int(tu) + __float_as_int(tv) * negW) - ioff : rej;
      }
      else
      {
        const int px = __float_as_int(tu) - pxoff;
        const int py = __float_as_int(tv) - pyoff;
        ok = cert && (unsigned)px < (unsigned)W && (unsigned)py < (unsigned)H;
        idx[m] = ok ? px + py * negW : rej;
      }
      anyv |= ok;
      if (!cert) need |= 1u << m;
      if (COUNT) cnt[0] += cert ? 1 : 0;
    }
    // ---- rare: voxels the FP32 tier could not certify (and every voxel behind the camera plane).
    // Dynamic m, so that lanes with different m run the same instructions together.
    while (need)
    {
      const int m = __ffs(need) - 1;
      need &= need - 1;
      const float fm = (float)m;
      const float fz = fmaf(fm, kz, fhz0);
      int id = kReject;                                                    // certified behind the camera
      bool okm = false;
      if (!(fz < -VC.zm))
      {
        if (COUNT) cnt[1]++;
        const float r = rcp_approx(fz);
        const float pu = fmaf(fmaf(fm, kx, fnx0), r, kMagic) - kMagic, pv = fmaf(fmaf(fm, ky, fny0), r, kMagic) - kMagic;
        id = tier2(c.v[S.view], c.cxc, c.cyc, W, H, di, dj, dk0 + (double)m, pu, pv);
        if (id == kNeedExact) need3 |= 1u << m;
        okm = id > kBad;
      }
      if (!okm) id = rej;
#pragma unroll
      for (int mm = 0; mm < HM; mm++) if (h * HM + mm == m) idx[h * HM + mm] = id;
      anyv |= okm;
    }
    anyh[h] = anyv;
    // ---- phase B: the half's gathers of the float classification image, all in flight together
    if (anyv)
    {
#pragma unroll
      for (int mm = 0; mm < HM; mm++) d32[h * HM + mm] = __ldg(cv + idx[h * HM + mm]);
    }
    else
    {
#pragma unroll
      for (int mm = 0; mm < HM; mm++) d32[h * HM + mm] = -1.0f;
    }
  }
  if (!need3 && !anyh[0] && !anyh[1]) return;

  unsigned near = 0;
  if (front2)                                                              // only validity matters for this view
  {
#pragma unroll
    for (int m = 0; m < FM; m++)
    {
      add_if_valid(acc[m], nerT, d32[m]);
      if (COUNT) { if (d32[m] == -1.0f) cnt[7]++; else { cnt[5]++; cnt[9]++; } }
    }
  }
  else
  {
    // ---- phase C: FP32 classification; -1.0f = invalid after the filter (:202).  In front and farther
    // than Delta (certified by the margin in thrfar): -Eta*Rho; behind and farther: 0 (:114-115);
    // everything else that is valid (NaN included) goes to the FP64 band below.
#pragma unroll
    for (int h = 0; h < NH; h++)
    {
      if (anyh[h])
      {
#pragma unroll
        for (int mm = 0; mm < HM; mm++)
        {
          const int m = h * HM + mm;
          const float fc = PINHOLE ? fmaf((float)m, kz, fhz0) : fmaf((float)m, kc, fcz0);   // camera z (:207)
          classify_far(acc[m], near, nerT, fc, d32[m], thrfar, 1u << m);
          if (COUNT)
          {
            const float df = fc - d32[m];
            if (d32[m] == -1.0f) cnt[7]++;
            else if (fabsf(df) > thrfar) { if (df < 0.f) cnt[5]++; else cnt[6]++; }
          }
        }
      }
    }
  }
  // ---- the band around the surface (a shell 2*Delta thick: whole bricks are in it or not): double depth, FP64
  // potential, inline and predicated
  if (near)
  {
    const double* dv = SPLIT ? nullptr : depths + voff;
    const int* lv = SPLIT ? lo + voff : nullptr;
    // explicit roundings: the value of a voxel must not depend on which unrolled copy (m) or brick decomposition
    // evaluates it, so that z-slabs concatenate bit-identically
    const double zij = __fma_rn(di, VC.czr[0], __fma_rn(dj, VC.czr[1], VC.czr[3]));
    const double czk = VC.czr[2], gd = VC.gd;
    const double delta = g.delta, thick = g.thick, rho = g.rho, rot = g.rho_over_thick;
    double dd[FM];
#pragma unroll
    for (int m = 0; m < FM; m++)
    {
      if (SPLIT) dd[m] = (near & (1u << m)) ? split_decode(d32[m], __ldg(lv + idx[m])) : 0.0;
      else dd[m] = (near & (1u << m)) ? __ldg(dv + idx[m]) : 0.0;
    }
#pragma unroll
    for (int m = 0; m < FM; m++)
    {
      if (near & (1u << m))
      {
        if (COUNT) cnt[4]++;
        const double z = __fma_rn(dk0 + (double)m, czk, zij);              // :207, from GLOBAL indices
        const double diff = __dsub_rn(z, dd[m]);
        const double ad = fabs(diff);
        const double td = __dsub_rn(ad, delta);
        if (fabs(td) < gd) need3 |= 1u << m;                                // within 2^-44 of the discontinuity: exact tier
        else if (td > 0.0)
        {
          if (!(diff > 0.0)) acc[m] = add_rn(acc[m], nerT);                // :114-115
        }
        else
        {
          const double res = (ad > thick) ? (diff > 0.0 ? rho : -rho)      // :116-117
                                          : __dmul_rn(rot, diff);          // :118-119
          acc[m] = add_rn(acc[m], (T)res);                                  // :211
        }
        if (COUNT && fabs(td) < gd) cnt[3]++;
      }
    }
  }
  // ---- T3: the reference's own operation sequence (~1e-9 of the voxels), one copy of the call
  while (need3)
  {
    const int m = __ffs(need3) - 1;
    need3 &= need3 - 1;
    if (COUNT) cnt[2]++;
    T a = (T)0;
#pragma unroll
    for (int mm = 0; mm < FM; mm++) if (mm == m) a = acc[mm];
    const int v = S.view;
    a = exact_unit<T>(&g, &c.e[v], SPLIT ? nullptr : depths + npix * v, cls + npix * v, SPLIT ? lo + npix * v : nullptr, i, j,
                      k0 + m, a);
#pragma unroll
    for (int mm = 0; mm < FM; mm++) if (mm == m) acc[mm] = a;
  }
}

constexpr int kColBricks = FSK;   // bricks of one work item: a column of the supertile along k (16 x 8 x 32 voxels)

// One work item = the FSK bricks of a supertile column against the chunk's views.  All threads of the CTA call it.
//   pre-pass   warp w examines brick w of the column: lane l takes the candidate views l and l + 32 (the views the
//              supertile kept), so that the dependent chain of the brick test (view record -> FP64 bases -> 8
//              reciprocals -> footprint tables) is paid once per FOUR bricks and by all four warps at once
//   main loop  brick after brick, all 128 threads, each view of the brick's list against 8 voxels per thread
// SPLIT: the double depths are not resident; the band and the exact tier rebuild them from the classification
// image and the residual image `lo` (split_decode), bit for bit.
template <typename T, bool PINHOLE, bool COUNT, bool SPLIT>
__device__ __forceinline__ void
fast_column(const GridParams& g, const FastChunk& c, const double* __restrict__ depths, const int* __restrict__ lo,
            const float* __restrict__ cls, const float* __restrict__ tileDmax, const TilePyramid& pyr, int cull,
            long long clsSpare, const ViewFast* __restrict__ gviews, const unsigned* __restrict__ stmasks,
            T* __restrict__ vol, int nbi, int nbj, int nbk, FastCounters* counters, unsigned st, unsigned col,
            ViewSm (*s_view)[kFastChunk], const ViewConst* s_const, int* s_cnt)
{
  // ---- column decode (CTA-uniform)
  const unsigned nsi = (nbi + FSI - 1) / FSI, nsj = (nbj + FSJ - 1) / FSJ;
  const int bi = (st % nsi) * FSI + col % FSI;
  const int bj = ((st / nsi) % nsj) * FSJ + col / FSI;
  const int bk0 = (st / (nsi * nsj)) * FSK;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int W = g.W, H = g.H;
  const bool colValid = bi < nbi && bj < nbj;
  const int i0 = bi * FBI, j0 = bj * FBJ;
  const long long npixAll = (long long)W * H;

  // ---- pre-pass: warp w <-> brick bk0 + w; thread order = view order, which the ballot compaction preserves.
  // The view is read from the GLOBAL copy of the chunk: every lane wants a different one, which a constant-bank
  // load would serialise 32 ways.
  {
    static_assert(kColBricks == FT / 32, "one warp per brick of the column");
    unsigned mlo = 0xffffffffu, mhi = 0xffffffffu;
    if (c.n < 64) { mlo = c.n >= 32 ? 0xffffffffu : ((1u << c.n) - 1u); mhi = c.n > 32 ? ((1u << (c.n - 32)) - 1u) : 0u; }
    if (cull && stmasks) { mlo &= __ldg(stmasks + 2 * st); mhi &= __ldg(stmasks + 2 * st + 1); }
    const int ncand = __popc(mlo) + __popc(mhi);
    const bool brickValid = colValid && bk0 + w < nbk;
    const int k0 = slab_global_k(g, (bk0 + w) * FM);
    int n = 0;
    for (int r0 = 0; r0 < ncand && brickValid; r0 += 32)
    {
      const int cand = r0 + lane;
      const bool have = cand < ncand;
      const int v = have ? nth_set_bit64(mlo, mhi, cand) : 0;
      const ViewFast& V = gviews[v];
      BoxEval e;
      e.keep = false;
      if (have)
        e = eval_box<PINHOLE>(V, c, pyr, tileDmax + (size_t)v * pyr.perView, i0, j0, k0, (float)(FBI - 1), (float)(FBJ - 1),
                              (float)(FM - 1), V.lx, V.ly, V.lz, V.lc, W, H, cull != 0);
      const unsigned bal = __ballot_sync(0xffffffffu, e.keep);
      if (e.keep)
      {
        ViewSm& S = s_view[w][n + __popc(bal & ((1u << lane) - 1u))];
        const float Ex = c.k3 * ((fabsf(e.fbx) + 3.f * V.lx) + e.Ux * (fabsf(e.fbz) + 3.f * V.lz));
        const float Ey = c.k3 * ((fabsf(e.fby) + 3.f * V.ly) + e.Uy * (fabsf(e.fbz) + 3.f * V.lz));
        const float Tx = 0.5f - (e.Ux * c.kq + 9.6e-7f), Ty = 0.5f - (e.Uy * c.kq + 9.6e-7f);
        S.base = make_float4(e.fbx, e.fby, e.fbz, e.fbc);
        // the per-voxel bound T - E*r (certification) at the largest r of the brick, rounded down
        S.tcert = e.mode ? fminf(fmaf(-Ex, e.rmax, Tx), fmaf(-Ey, e.rmax, Ty)) - 1e-7f : -1.f;
        S.thrfar = c.delta_up + 1e-6f * e.czmaxabs;
        S.view = v;
        S.flags = e.front | (e.mode << 2);
        S.voff = (long long)npixAll * v + (long long)(H - 1) * W;
        S.rej = (int)(clsSpare - S.voff);
        S.pad = 0;
      }
      n += __popc(bal);
    }
    if (lane == 0) s_cnt[w] = n;
  }
  __syncthreads();
  if (COUNT && counters && threadIdx.x == 0)
  {
    int nb = 0, ns = 0;
    for (int q = 0; q < kColBricks; q++) if (colValid && bk0 + q < nbk) { nb++; ns += s_cnt[q]; }
    atomicAdd(&counters->culled, (unsigned long long)(nb * c.n - ns));
    atomicAdd(&counters->brick_views, (unsigned long long)(nb * c.n));
  }
  if (!colValid) return;

  const int li = (w & 1) * 8 + (lane & 7), lj = (w >> 1) * 4 + (lane >> 3);
  const int i = i0 + li, j = j0 + lj;
  if (i >= g.Nx || j >= g.Ny) return;                         // no barrier below
  const float fli = (float)li, flj = (float)lj;
  const double di = (double)i, dj = (double)j;
  const size_t plane = (size_t)g.Nx * g.Ny, npix = (size_t)W * H;
  const T nerT = (T)g.neg_eta_rho;
  unsigned long long cnt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // t1, t2, t3, delta guard, band, front, behind, invalid, -, validity-only
  unsigned long long n_units = 0, n_uf = 0;

#pragma unroll 1
  for (int q = 0; q < kColBricks; q++)
  {
    const int nsurv = s_cnt[q];
    if (nsurv == 0 || bk0 + q >= nbk) continue;               // the brick's voxels are not even read
    const int lp0 = (bk0 + q) * FM, k0 = slab_global_k(g, lp0);
    const double dk0 = (double)k0;
    // voxels beyond the slab's last plane (m >= nk) are computed and discarded: no divergence in the loop
    const int nk = min(FM, g.nLocal - lp0);
    T acc[FM];
    T* p = vol + ((size_t)lp0 * g.Ny + j) * g.Nx + i;
#pragma unroll
    for (int m = 0; m < FM; m++) acc[m] = (m < nk) ? p[m * plane] : (T)0;

#pragma unroll 1
    for (int s = 0; s < nsurv; s++)
    {
      const ViewSm& S = s_view[q][s];
      const int flags = S.flags;                              // CTA-uniform
      if ((flags & 3) == 1)                                   // the whole brick is free space for this view
      {
#pragma unroll
        for (int m = 0; m < FM; m++) acc[m] = add_rn(acc[m], nerT);
        if (COUNT) n_uf += nk;
        continue;
      }
      if (COUNT) n_units += nk;
      const ViewConst& VC = s_const[S.view];
      const bool f2 = (flags & 3) == 2;
      if ((flags >> 2) == 2)
        integrate_view<T, PINHOLE, COUNT, SPLIT, true>(g, c, S, VC, depths, lo, cls, npix, fli, flj, di, dj, dk0, i, j, k0, nerT, f2,
                                                       acc, cnt);
      else
        integrate_view<T, PINHOLE, COUNT, SPLIT, false>(g, c, S, VC, depths, lo, cls, npix, fli, flj, di, dj, dk0, i, j, k0, nerT, f2,
                                                        acc, cnt);
    }
#pragma unroll
    for (int m = 0; m < FM; m++)
      if (m < nk) p[m * plane] = acc[m];
  }

  if (COUNT && counters)
  {
    // one atomic per warp and counter (the lanes that returned early above are simply absent)
    const unsigned act = __activemask();
    unsigned long long vals[11] = {cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], n_units, cnt[5], cnt[6], cnt[7], n_uf, cnt[9]};
#pragma unroll
    for (int q = 0; q < 11; q++)
    {
      unsigned long long x = vals[q];
      for (int o = 16; o > 0; o >>= 1)
      {
        const unsigned long long y = __shfl_xor_sync(act, x, o);
        if ((act >> ((lane ^ o) & 31)) & 1u) x += y;
      }
      vals[q] = x;
    }
    if (lane == (__ffs(act) - 1))
    {
      atomicAdd(&counters->t1_certified, vals[0]);
      atomicAdd(&counters->t2_entered, vals[1]);
      atomicAdd(&counters->t3_entered, vals[2]);
      atomicAdd(&counters->delta_guard, vals[3]);
      atomicAdd(&counters->near_band, vals[4]);
      atomicAdd(&counters->units, vals[5]);
      atomicAdd(&counters->reserved[0], vals[6]);
      atomicAdd(&counters->reserved[1], vals[7]);
      atomicAdd(&counters->reserved[2], vals[8]);
      atomicAdd(&counters->uniform_front, vals[9]);
      atomicAdd(&counters->reserved[3], vals[10]);
    }
  }
}

// Persistent kernel: as many CTAs as fit the GPU; each takes the next column of the next ACTIVE supertile from a
// global counter until none is left (a launch over all bricks would spend ~2 ns on each of the many bricks
// no view of the chunk can touch).  The next item is fetched while the current one is integrated.  A CTA retires after
// `quota` items, so that higher-priority kernels of other streams (the next views' preparation, NCCL) find a free slot
// within a fraction of a millisecond instead of waiting for the whole launch (DMI_OPT_BRICK_QUOTA, in bricks).
template <typename T, bool PINHOLE, bool COUNT, bool SPLIT>
__global__ void __launch_bounds__(FT, DMI_FAST_CTAS)
tsdf_fast_kernel(const __grid_constant__ GridParams g, const __grid_constant__ FastChunk c,
                 const double* __restrict__ depths, const int* __restrict__ lo, const float* __restrict__ cls,
                 const float* __restrict__ tileDmax, const __grid_constant__ TilePyramid pyr, int cull,
                 long long clsSpare, const ViewFast* __restrict__ gviews, const unsigned* __restrict__ stmasks,
                 const int* __restrict__ stlist, int* __restrict__ work, int quota,
                 T* __restrict__ vol, int nbi, int nbj, int nbk, FastCounters* counters)
{
  static_assert(kFastChunk <= 64, "two rounds of 32 candidate views per brick");
  __shared__ ViewSm s_view[kColBricks][kFastChunk];
  __shared__ ViewConst s_const[kFastChunk];
  __shared__ int s_cnt[kColBricks];
  __shared__ int s_item[2];
  constexpr int per = FSI * FSJ;
  const int nItems = work[0] * per;
  if ((int)threadIdx.x < c.n)
  {
    // from the GLOBAL copy: one view per thread out of the constant bank would be serialised 64 ways
    const ViewFast& V = gviews[threadIdx.x];
    ViewConst& K = s_const[threadIdx.x];
    K.cx = make_float4(V.fnx[0], V.fnx[1], V.fnx[2], 0.f);
    K.cy = make_float4(V.fny[0], V.fny[1], V.fny[2], 0.f);
    K.cz = make_float4(V.fhz[0], V.fhz[1], V.fhz[2], 0.f);
    K.cc = make_float4(V.fcz[0], V.fcz[1], V.fcz[2], 0.f);
    K.czr[0] = V.cz[0]; K.czr[1] = V.cz[1]; K.czr[2] = V.cz[2]; K.czr[3] = V.cz[3];
    K.gd = V.gd;
    K.zm = V.zm; K.pad = 0.f;
  }
  if (threadIdx.x == 0) s_item[0] = atomicAdd(work + 1, 1);
  __syncthreads();
  int item = s_item[0];
  for (int it = 0; it < quota && item < nItems; it++)
  {
    // the next item's number arrives while this one is integrated (a CTA about to retire takes none)
    if (threadIdx.x == 0) s_item[(it + 1) & 1] = (it + 1 < quota) ? atomicAdd(work + 1, 1) : nItems;
    fast_column<T, PINHOLE, COUNT, SPLIT>(g, c, depths, lo, cls, tileDmax, pyr, cull, clsSpare, gviews, stmasks, vol, nbi, nbj, nbk,
                                          counters, (unsigned)__ldg(stlist + item / per), (unsigned)(item % per), s_view, s_const, s_cnt);
    __syncthreads();                                          // this item is done with shared memory
    item = s_item[(it + 1) & 1];
  }
}

// pad[0] of each staged view = the view's "has a fully valid tile" flag from its tile statistics: views
// without one (salt-and-pepper holes) skip the free-space test of eval_box altogether.
__global__ void __launch_bounds__(256) stage_views_kernel(const __grid_constant__ FastChunk c, ViewFast* __restrict__ dst,
                                                          const float* __restrict__ tiles, int perView, int flagOff)
{
  const unsigned* src = reinterpret_cast<const unsigned*>(c.v);
  unsigned* d = reinterpret_cast<unsigned*>(dst);
  const int words = (int)(sizeof(ViewFast) / 4) * c.n;
  for (int q = threadIdx.x; q < words; q += blockDim.x) d[q] = src[q];
  __syncthreads();
  if ((int)threadIdx.x < c.n) dst[threadIdx.x].pad[0] = tiles[(size_t)threadIdx.x * perView + flagOff];
}

// Enough CTAs to take every work item at `quota` each, at least as many as are resident at once on this device
static unsigned persistent_grid(const void* kernel, unsigned bricks, unsigned quota)
{
  int dev = 0, sms = 0, perSm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, FT, 0);
  const unsigned resident = (unsigned)(std::max(sms, 1) * std::max(perSm, 1));
  return std::max(1u, std::min(bricks, std::max(resident, (bricks + quota - 1) / quota)));
}

template <typename T, bool PINHOLE>
static void launch_variant(unsigned grid, const GridParams& g, const FastChunk& c, const double* d_depths, const int* d_lo,
                           const float* d_cls, const float* d_tileDmax, const TilePyramid& pyr, bool cull,
                           long long clsSpare, const ViewFast* d_views, unsigned* d_masks, T* d_vol, int nbi, int nbj,
                           int nbk, FastCounters* d_counters, int quota, cudaStream_t s)
{
  const int nst = (int)(grid / (FSI * FSJ * FSK));
  grid /= kColBricks;                                         // work items = supertile columns of FSK bricks
  quota = std::max(1, quota / kColBricks);
  if (cull)
    supertile_cull_kernel<PINHOLE><<<(nst + 3) / 4, 256, 0, s>>>(g, c, d_tileDmax, pyr, d_views, nbi, nbj, nbk, d_masks, nst);
  const unsigned* masks = cull ? d_masks : nullptr;
  int* list = reinterpret_cast<int*>(d_masks + 2 * (size_t)nst);
  int* work = list + nst;
  compact_supertiles_kernel<<<1, 1024, 0, s>>>(d_masks, nst, cull ? 1 : 0, list, work);
#define DMI_LAUNCH_FAST(COUNT, SPLIT, CNT)                                                                         \
  tsdf_fast_kernel<T, PINHOLE, COUNT, SPLIT><<<persistent_grid((const void*)tsdf_fast_kernel<T, PINHOLE, COUNT, SPLIT>, grid, (unsigned)quota), FT, 0, s>>>( \
      g, c, d_depths, d_lo, d_cls, d_tileDmax, pyr, cull ? 1 : 0, clsSpare, d_views, masks, list, work, quota, d_vol, nbi, nbj, nbk, CNT)
  if (d_counters) { if (d_depths) DMI_LAUNCH_FAST(true, false, d_counters); else DMI_LAUNCH_FAST(true, true, d_counters); }
  else { if (d_depths) DMI_LAUNCH_FAST(false, false, nullptr); else DMI_LAUNCH_FAST(false, true, nullptr); }
#undef DMI_LAUNCH_FAST
}

size_t tsdf_fast_mask_bytes(const GridParams& g)
{
  const int nbi = (g.Nx + FBI - 1) / FBI, nbj = (g.Ny + FBJ - 1) / FBJ, nbk = (g.nLocal + FM - 1) / FM;
  const size_t nsi = (nbi + FSI - 1) / FSI, nsj = (nbj + FSJ - 1) / FSJ, nsk = (nbk + FSK - 1) / FSK;
  return nsi * nsj * nsk * 12 + 16;       // masks (2 words), the active list (1 word) per supertile + the work counters
}

cudaError_t launch_tsdf_fast(const GridParams& g, const FastChunk& c, const double* d_depths, const int* d_lo,
                             const float* d_cls, long long clsSpare, const float* d_tileDmax, bool cull,
                             ViewFast* d_viewScratch, unsigned* d_maskScratch, void* d_vol, int scalarType,
                             FastCounters* d_counters, int quota, cudaStream_t s)
{
  quota = std::max(1, quota);
  const int nbi = (g.Nx + FBI - 1) / FBI, nbj = (g.Ny + FBJ - 1) / FBJ, nbk = (g.nLocal + FM - 1) / FM;
  if (nbi <= 0 || nbj <= 0 || nbk <= 0 || c.n <= 0) return cudaSuccess;
  const unsigned nsi = (nbi + FSI - 1) / FSI, nsj = (nbj + FSJ - 1) / FSJ, nsk = (nbk + FSK - 1) / FSK;
  const unsigned grid = nsi * nsj * nsk * (FSI * FSJ * FSK);
  const TilePyramid pyr = tile_pyramid_layout(g.W, g.H);
  // stream-ordered copy of the views from the parameter space to global memory, for the pre-pass
  stage_views_kernel<<<1, 256, 0, s>>>(c, d_viewScratch, d_tileDmax, pyr.perView, pyr.flagOff);
  const ViewFast* d_views = d_viewScratch;
  if (scalarType == 1)
  {
    if (c.pinhole) launch_variant<double, true>(grid, g, c, d_depths, d_lo, d_cls, d_tileDmax, pyr, cull, clsSpare, d_views, d_maskScratch, (double*)d_vol, nbi, nbj, nbk, d_counters, quota, s);
    else launch_variant<double, false>(grid, g, c, d_depths, d_lo, d_cls, d_tileDmax, pyr, cull, clsSpare, d_views, d_maskScratch, (double*)d_vol, nbi, nbj, nbk, d_counters, quota, s);
  }
  else
  {
    if (c.pinhole) launch_variant<float, true>(grid, g, c, d_depths, d_lo, d_cls, d_tileDmax, pyr, cull, clsSpare, d_views, d_maskScratch, (float*)d_vol, nbi, nbj, nbk, d_counters, quota, s);
    else launch_variant<float, false>(grid, g, c, d_depths, d_lo, d_cls, d_tileDmax, pyr, cull, clsSpare, d_views, d_maskScratch, (float*)d_vol, nbi, nbj, nbk, d_counters, quota, s);
  }
  return cudaGetLastError();
}

// ---- view preparation: best-cost filter (ReconstructionData.cxx:159-166) folded into a float
// classification image, plus the sparse tables of tile statistics for the brick tests.  Level 0: one warp
// per 16 x 16 pixels of storage rows (2 x 2 tiles), each lane 8 consecutive pixels; a single HBM-bound pass
// over the maps.
TilePyramid tile_pyramid_layout(int W, int H)
{
  TilePyramid p;
  p.tw = (W + kTile - 1) / kTile;
  p.th = (H + kTile - 1) / kTile;
  int m = std::max(std::max(p.tw, p.th), 1), l = 0;
  while ((2 << l) <= m) l++;                                  // floor(log2(m))
  p.nLevels = l + 1;
  const int L = p.nLevels * p.tw * p.th;
  p.badOff = 2 * L;
  p.flagOff = 3 * L;
  p.perView = (p.flagOff + 4 + 3) / 4 * 4;                    // every view's statistics start 16-byte aligned
  return p;
}

// Every output is written to ALL destinations of `dst` (dst 0 is local, the others may be peer GPUs' mapped
// buffers: plain stores over NVLink, 32-byte vectors when the rows allow it) -- view preparation fused with
// its all-gather.
__global__ void __launch_bounds__(256)
prepare_views_kernel(const double* __restrict__ depths, const double* __restrict__ cost, double thr,
                     int nViews, int W, int H, int TW, int TH, int perView, int badOff,
                     const __grid_constant__ PrepareDst dst, long long clsSpare)
{
  if (blockIdx.x == 0 && threadIdx.x == 0 && clsSpare >= 0)
    for (int r = 0; r < dst.n; r++) dst.cls[r][clsSpare] = -1.0f;                     // the spare slot, see phase B
  const int lane = threadIdx.x & 31;
  const int BW = (TW + 1) / 2, BH = (TH + 1) / 2;             // blocks of 2 x 2 tiles
  const size_t blk = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t blocksPerView = (size_t)BW * BH;
  if (blk >= blocksPerView * nViews) return;
  const int v = (int)(blk / blocksPerView);
  const int t = (int)(blk % blocksPerView);
  const int by = t / BW, bx = t % BW;
  const int row = by * 16 + (lane >> 1), col = bx * 16 + (lane & 1) * 8;
  float dmax = -INFINITY, dmin = INFINITY, bad = 0.f;
  if (row < H && col < W)
  {
    const size_t base = (size_t)v * W * H + (size_t)row * W + col;
    float f8[8];
    int l8[8];
    const bool wantLo = dst.lo[0] != nullptr;
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
      f8[q] = -1.0f; l8[q] = 0;
      if (col + q < W)
      {
        const double d = __ldg(depths + base + q);
        const bool invalid = (d == -1.0) || (cost && __ldg(cost + base + q) > thr);
        if (!invalid)
        {
          float f = __double2float_rn(d);
          if (f == -1.0f) f = (d < -1.0) ? -1.00000012f : -0.99999994f;       // never -1.0f on a valid pixel
          // tile maximum, rounded up; NaN poisons the tile (+inf: never culled as "far behind")
          float up = f;
          if ((double)up < d) up = __int_as_float(__float_as_int(up) + (up >= 0.f ? 1 : -1));
          dmax = (d != d) ? INFINITY : fmaxf(dmax, up);
          // tile minimum of the valid depths, rounded down; NaN poisons it (-inf: never "in front of everything")
          float dn = f;
          if ((double)dn > d) dn = __int_as_float(__float_as_int(dn) + (dn > 0.f ? -1 : 1));
          dmin = (d != d) ? -INFINITY : fminf(dmin, dn);
          if (d != d) bad = 1.f;
          f8[q] = f;
          if (wantLo) l8[q] = split_encode(d, f);
        }
        else bad = 1.f;                                        // an invalid pixel: the tile is not fully valid
      }
    }
    const bool vec = ((W & 7) == 0) && dst.aligned;            // whole 32-byte groups, 32-byte aligned rows
    for (int r = 0; r < dst.n; r++)
    {
      float* c = dst.cls[r] + base;
      int* l = wantLo ? dst.lo[r] + base : nullptr;
      if (vec)
      {
        reinterpret_cast<float4*>(c)[0] = make_float4(f8[0], f8[1], f8[2], f8[3]);
        reinterpret_cast<float4*>(c)[1] = make_float4(f8[4], f8[5], f8[6], f8[7]);
        if (l)
        {
          reinterpret_cast<int4*>(l)[0] = make_int4(l8[0], l8[1], l8[2], l8[3]);
          reinterpret_cast<int4*>(l)[1] = make_int4(l8[4], l8[5], l8[6], l8[7]);
        }
      }
      else
      {
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (col + q < W) { c[q] = f8[q]; if (l) l[q] = l8[q]; }
      }
    }
  }
  // the 8 lanes of a tile: same column half (bit 0) and same row half (bit 4)
#pragma unroll
  for (int o = 2; o <= 8; o <<= 1)
  {
    dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    bad = fmaxf(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  }
  const int tx = bx * 2 + (lane & 1), ty = by * 2 + (lane >> 4);
  if ((lane & 14) == 0 && tx < TW && ty < TH)
  {
    const size_t o = (size_t)v * perView;
    const int q = ty * TW + tx;
    for (int r = 0; r < dst.n; r++)
    {
      reinterpret_cast<float2*>(dst.tiles[r] + o)[q] = make_float2(dmax, dmin);
      dst.tiles[r][o + badOff + q] = bad;
    }
  }
}

// level l from level l-1: windows of 2^l tiles = four windows of 2^(l-1) tiles, h = 2^(l-1) apart (clipped);
// read from the local copy, written everywhere
__global__ void __launch_bounds__(256)
tile_level_kernel(const __grid_constant__ PrepareDst dst, int nViews, int perView, int badOff, int tw, int th, int l)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)tw * th;
  if (t >= per * nViews) return;
  const int v = (int)(t / per), q = (int)(t % per);
  const int y = q / tw, x = q % tw, h = 1 << (l - 1);
  const int x1 = min(x + h, tw - 1), y1 = min(y + h, th - 1);
  const float* view = dst.tiles[0] + (size_t)v * perView;
  const float2* src = reinterpret_cast<const float2*>(view) + (size_t)(l - 1) * per;
  const float2 a = src[y * tw + x], b = src[y * tw + x1], c = src[y1 * tw + x], d = src[y1 * tw + x1];
  const float2 mm = make_float2(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fminf(fminf(a.y, b.y), fminf(c.y, d.y)));
  const float* srb = view + badOff + (size_t)(l - 1) * per;
  const float bad = fmaxf(fmaxf(srb[y * tw + x], srb[y * tw + x1]), fmaxf(srb[y1 * tw + x], srb[y1 * tw + x1]));
  const size_t o = (size_t)v * perView;
  for (int r = 0; r < dst.n; r++)
  {
    reinterpret_cast<float2*>(dst.tiles[r] + o)[(size_t)l * per + q] = mm;
    dst.tiles[r][o + badOff + (size_t)l * per + q] = bad;
  }
}

// flag = 1.0f when some tile of the view is fully valid
__global__ void __launch_bounds__(256) view_flag_kernel(const __grid_constant__ PrepareDst dst, int perView, int badOff, int flagOff, int n0)
{
  const float* t = dst.tiles[0] + (size_t)blockIdx.x * perView + badOff;
  int any = 0;
  for (int q = threadIdx.x; q < n0; q += blockDim.x) any |= (t[q] == 0.f) ? 1 : 0;
  any = __syncthreads_or(any);
  if (threadIdx.x == 0)
    for (int r = 0; r < dst.n; r++) dst.tiles[r][(size_t)blockIdx.x * perView + flagOff] = any ? 1.0f : 0.0f;
}

// Level 0 of the tile statistics from the CLASSIFICATION image alone (what a rank does for views it received over NVLink:
// the statistics themselves are not exchanged).  cls = the depth rounded to the nearest float, so the neighbouring floats
// bound the double from above and below: one ulp looser than the owner's own bounds, still exact tests (the brick tests
// only ever prove that a view contributes nothing / free space; a looser bound proves it a little less often).
// One warp per 16 x 16 pixels of storage rows (2 x 2 tiles), each lane 8 consecutive pixels, like prepare_views_kernel.
__global__ void __launch_bounds__(256)
tile_stats_from_cls_kernel(const float* __restrict__ cls, int nViews, int W, int H, int TW, int TH, int perView, int badOff,
                           float* __restrict__ tiles)
{
  const int lane = threadIdx.x & 31;
  const int BW = (TW + 1) / 2, BH = (TH + 1) / 2;
  const size_t blk = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t blocksPerView = (size_t)BW * BH;
  if (blk >= blocksPerView * nViews) return;
  const int v = (int)(blk / blocksPerView);
  const int t = (int)(blk % blocksPerView);
  const int by = t / BW, bx = t % BW;
  const int row = by * 16 + (lane >> 1), col = bx * 16 + (lane & 1) * 8;
  float dmax = -INFINITY, dmin = INFINITY, bad = 0.f;
  if (row < H && col < W)
  {
    const float* p = cls + (size_t)v * W * H + (size_t)row * W + col;
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (col + q < W)
      {
        const float f = __ldg(p + q);
        if (f == -1.0f) bad = 1.f;                               // invalid after the filter
        else if (f != f) { dmax = INFINITY; dmin = -INFINITY; bad = 1.f; }      // NaN poisons the tile
        else
        {
          // the double lies within half an ulp of f: the next floats up / down bound it (inf stays inf)
          const float up = (fabsf(f) <= 3.4e38f) ? __int_as_float(__float_as_int(f) + ((f > 0.f) ? 1 : (f < 0.f ? -1 : 0))) : f;
          const float dn = (fabsf(f) <= 3.4e38f) ? __int_as_float(__float_as_int(f) + ((f > 0.f) ? -1 : (f < 0.f ? 1 : 0))) : f;
          dmax = fmaxf(dmax, f == 0.f ? 1.4e-45f : up);
          dmin = fminf(dmin, f == 0.f ? -1.4e-45f : dn);
        }
      }
  }
#pragma unroll
  for (int o = 2; o <= 8; o <<= 1)
  {
    dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    bad = fmaxf(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  }
  const int tx = bx * 2 + (lane & 1), ty = by * 2 + (lane >> 4);
  if ((lane & 14) == 0 && tx < TW && ty < TH)
  {
    const size_t o = (size_t)v * perView;
    const int q = ty * TW + tx;
    reinterpret_cast<float2*>(tiles + o)[q] = make_float2(dmax, dmin);
    tiles[o + badOff + q] = bad;
  }
}

// All levels of the tile statistics of nViews views from their classification images (see tile_stats_from_cls_kernel)
cudaError_t launch_tile_stats_from_cls(const float* d_cls, int nViews, int W, int H, float* d_tileStats, cudaStream_t s)
{
  const TilePyramid p = tile_pyramid_layout(W, H);
  const size_t blocks2x2 = (size_t)((p.tw + 1) / 2) * ((p.th + 1) / 2) * nViews;
  if (blocks2x2 == 0) return cudaSuccess;
  tile_stats_from_cls_kernel<<<(unsigned)((blocks2x2 + 7) / 8), 256, 0, s>>>(d_cls, nViews, W, H, p.tw, p.th, p.perView, p.badOff, d_tileStats);
  PrepareDst dst = {};
  dst.n = 1;
  dst.tiles[0] = d_tileStats;
  const size_t n = (size_t)p.tw * p.th * nViews;
  for (int l = 1; l < p.nLevels; l++)
    tile_level_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, nViews, p.perView, p.badOff, p.tw, p.th, l);
  view_flag_kernel<<<nViews, 256, 0, s>>>(dst, p.perView, p.badOff, p.flagOff, p.tw * p.th);
  return cudaGetLastError();
}

// levels: false = only the single pass over the maps (classification / residual images, level 0 of the statistics)
cudaError_t launch_prepare_views(const double* d_depths, const double* d_cost, double thr, int nViews, int W, int H,
                                 const PrepareDst& dst, long long clsSpare, cudaStream_t s, bool levels)
{
  const TilePyramid p = tile_pyramid_layout(W, H);
  const size_t blocks2x2 = (size_t)((p.tw + 1) / 2) * ((p.th + 1) / 2) * nViews;
  if (blocks2x2 == 0 || dst.n <= 0) return cudaSuccess;
  prepare_views_kernel<<<(unsigned)((blocks2x2 + 7) / 8), 256, 0, s>>>(d_depths, d_cost, thr, nViews, W, H, p.tw, p.th, p.perView,
                                                                       p.badOff, dst, clsSpare);
  if (!levels) return cudaGetLastError();
  const size_t n = (size_t)p.tw * p.th * nViews;
  for (int l = 1; l < p.nLevels; l++)
    tile_level_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, nViews, p.perView, p.badOff, p.tw, p.th, l);
  view_flag_kernel<<<nViews, 256, 0, s>>>(dst, p.perView, p.badOff, p.flagOff, p.tw * p.th);
  return cudaGetLastError();
}

cudaError_t launch_prepare_views(const double* d_depths, const double* d_cost, double thr, int nViews, int W, int H,
                                 float* d_cls, int* d_lo, long long clsSpare, float* d_tileStats, cudaStream_t s, bool levels)
{
  PrepareDst dst = {};
  dst.n = 1;
  dst.cls[0] = d_cls; dst.lo[0] = d_lo; dst.tiles[0] = d_tileStats;
  dst.aligned = ((reinterpret_cast<uintptr_t>(d_cls) | reinterpret_cast<uintptr_t>(d_lo)) & 31) == 0;
  return launch_prepare_views(d_depths, d_cost, thr, nViews, W, H, dst, clsSpare, s, levels);
}

// ---- host side: composition of the per-view affine rows ----------------------------------------

static float up(double x)      // a float >= x (x >= 0)
{
  float f = (float)x;
  if ((double)f < x) f = nextafterf(f, INFINITY);
  return f;
}

void fill_fast_chunk_constants(const GridParams& g, FastChunk* c)
{
  c->cxc = g.W / 2;
  c->cyc = g.H / 2;
  // in-bounds centred pixels satisfy |u'| <= max(W, H) / 2 + 0.5
  c->umax1g = up(std::max(g.W, g.H) / 2.0 + 2.0);
  // E = (4/3) * 3 * 2^-24 * (1 + 1%) * (...)
  c->k3 = up(4.04 * std::ldexp(1.0, -24));
  // c0 = (4/3) * U1 * 2^-23 * 1.05 (MUFU.RCP: 1 ulp), + 2^-20 of slack for the float arithmetic of the test
  c->kq = up((4.0 / 3.0) * 1.05 * std::ldexp(1.0, -23));
  c->delta_up = up(g.delta * (1.0 + std::ldexp(1.0, -20)));
}

// rows as affine functions of the global voxel index, in long double
void compose_fast_view(const GridParams& g, const double* K16, const double* RT16, int cxc, int cyc, ViewFast* out)
{
  typedef long double L;
  // world = Gm * (orig + (idx + 0.5) * sp) + Gt  ->  WA (3x3) * idx + Wb
  L WA[3][3], Wb[3];
  for (int r = 0; r < 3; r++)
  {
    Wb[r] = (L)g.gm[4 * r + 3];
    for (int a = 0; a < 3; a++)
    {
      WA[r][a] = (L)g.gm[4 * r + a] * (L)g.sp[a];
      Wb[r] += (L)g.gm[4 * r + a] * ((L)g.orig[a] + (L)0.5 * (L)g.sp[a]);
    }
  }
  // cam = R * world + T
  L CA[3][3], Cb[3];
  for (int r = 0; r < 3; r++)
  {
    Cb[r] = (L)RT16[4 * r + 3];
    for (int a = 0; a < 3; a++)
    {
      CA[r][a] = 0;
      for (int q = 0; q < 3; q++) CA[r][a] += (L)RT16[4 * r + q] * WA[q][a];
      Cb[r] += (L)RT16[4 * r + a] * Wb[a];
    }
  }
  // h = K(3x4) * (cam, 1)
  L HA[3][3], Hb[3];
  for (int r = 0; r < 3; r++)
  {
    Hb[r] = (L)K16[4 * r + 3];
    for (int a = 0; a < 3; a++)
    {
      HA[r][a] = 0;
      for (int q = 0; q < 3; q++) HA[r][a] += (L)K16[4 * r + q] * CA[q][a];
      Hb[r] += (L)K16[4 * r + a] * Cb[a];
    }
  }
  const L cc[2] = {(L)cxc, (L)cyc};
  double* rows[2] = {out->nx, out->ny};
  for (int r = 0; r < 2; r++)
  {
    for (int a = 0; a < 3; a++) rows[r][a] = (double)(HA[r][a] - cc[r] * HA[2][a]);
    rows[r][3] = (double)(Hb[r] - cc[r] * Hb[2]);
  }
  for (int a = 0; a < 3; a++) { out->hz[a] = (double)HA[2][a]; out->cz[a] = (double)CA[2][a]; }
  out->hz[3] = (double)Hb[2];
  out->cz[3] = (double)Cb[2];
  for (int a = 0; a < 4; a++)
  {
    out->fnx[a] = (float)out->nx[a]; out->fny[a] = (float)out->ny[a];
    out->fhz[a] = (float)out->hz[a]; out->fcz[a] = (float)out->cz[a];
  }

  // magnitude bounds over the whole grid (global indices 0..N-1), for the FP64 margins
  const double im[3] = {(double)std::max(g.Nx - 1, 0), (double)std::max(g.Ny - 1, 0), (double)std::max(g.Nz - 1, 0)};
  auto S = [&](const double* r) { return std::fabs(r[0]) * im[0] + std::fabs(r[1]) * im[1] + std::fabs(r[2]) * im[2] + std::fabs(r[3]); };
  // the reference evaluates h.x = K00*cx + K01*cy + K02*cz + K03 etc.: its intermediate magnitudes are
  // bounded by S(nx) + cxc*S(hz) (same for y); 2^-44 of that covers both evaluation orders many times over
  const double umax = std::max(g.W, g.H) / 2.0 + 2.0;
  const double Sx = S(out->nx) + std::fabs((double)cxc) * S(out->hz);
  const double Sy = S(out->ny) + std::fabs((double)cyc) * S(out->hz);
  out->m2 = std::ldexp(1.0, -44) * (std::max(Sx, Sy) + umax * S(out->hz));
  out->m2z = std::ldexp(1.0, -44) * S(out->hz);
  out->gd = std::ldexp(1.0, -44) * (S(out->cz) + 1.0 + std::fabs(g.delta));
  // local offsets inside one brick
  auto Lb = [&](const double* r) { return up(std::fabs(r[0]) * (FBI - 1) + std::fabs(r[1]) * (FBJ - 1) + std::fabs(r[2]) * (FM - 1)); };
  out->lx = Lb(out->nx); out->ly = Lb(out->ny); out->lz = Lb(out->hz); out->lc = Lb(out->cz);
  // fz > zm must imply dz/fz <= 1/4 with dz <= 3.03 * 2^-24 * (|hz| + 3 lz): zm = 12.5 * 2^-24 * max(...)
  out->zm = up(12.5 * std::ldexp(1.0, -24) * (S(out->hz) + 3.0 * (double)out->lz));
  out->pad[0] = out->pad[1] = out->pad[2] = 0.f;
}

}  // namespace dmi
