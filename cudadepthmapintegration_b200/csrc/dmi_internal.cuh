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
  int k0, k1;           // contiguous z-slab owned by this context, global cell indices (layL == 0)
  // layered slab (layL > 0, a multiple of 32): the context owns the layers layPhase, layPhase + layStride, ... of layL
  // planes each; its volume holds them packed in that order.  nLocal = planes owned (k1 - k0 for a contiguous slab).
  int layL, layStride, layPhase, nLocal;
  int W, H;             // c_depthMapDims
  double thick, rho, eta, delta;
  double rho_over_thick;  // (rho / thick) in double, as CudaReconstruction.cu:119 evaluates it
  double neg_eta_rho;     // -eta * rho, CudaReconstruction.cu:115
};

// local plane of the slab -> global cell index k (voxel centres always use GLOBAL indices, CudaReconstruction.cu:78-83)
__host__ __device__ inline int slab_global_k(const GridParams& g, int lp)
{
  return g.layL ? ((lp / g.layL) * g.layStride + g.layPhase) * g.layL + lp % g.layL : g.k0 + lp;
}

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

// One view, fast-path form (see tsdf_fast.cu).  Every row is an affine function of the GLOBAL voxel
// index: value(i,j,k) = r[0]*i + r[1]*j + r[2]*k + r[3], composed on the host in long double from
// K * RT * gridMatrix * (orig + (idx + 0.5) * spacing).
//   nx, ny  centred numerators  h.x - cxc*h.z,  h.y - cyc*h.z   (pixel = nx/hz + cxc)
//   hz      homogeneous denominator h.z (CudaReconstruction.cu:176-184)
//   cz      camera z = the "real depth" of CudaReconstruction.cu:207
struct ViewFast
{
  double nx[4], ny[4], hz[4], cz[4];
  double m2;                      // T2 margin on s = nx - p*hz (2^-44 of the rows' magnitude bound)
  double m2z;                     // T2 margin on hz
  double gd;                      // | |diff| - Delta | below this -> exact tier
  float fnx[4], fny[4], fhz[4], fcz[4];   // the same rows rounded to float ([3] unused: bases are per brick)
  float lx, ly, lz, lc;           // bound on |local offset terms| of each row inside one brick
  float zm;                       // |fz| <= zm: too close to the camera plane for the FP32 tier
  float pad[3];
};

constexpr int kFastChunk = 64;    // views per launch: 64 * (248 + 192) B = 28 KB of kernel parameters
struct FastChunk
{
  int n;
  int cxc, cyc;                   // integer pixel offsets removed from the numerators
  int pinhole;                    // every K of the chunk has last row 0 0 1 0: h.z == camera z
  float umax1g;                   // largest centred |pixel| that can be in bounds, + 2 (fallback bound)
  float k3;                       // E = k3 * ((|b| + 3 l) + U1 * (|bz| + 3 lz))
  float kq;                       // T = 0.5 - U1 * kq - 2^-20
  float delta_up;                 // float >= Delta * (1 + 2^-20)
  ViewFast v[kFastChunk];
  ViewExact e[kFastChunk];
};

struct FastCounters               // optional diagnostics (device memory, may be null)
{
  unsigned long long t1_certified, t2_entered, t3_entered, delta_guard, units, culled, near_band, brick_views;
  unsigned long long uniform_front;     // voxel*views settled brick-wide by one add per voxel, see eval_box
  unsigned long long reserved[7];       // [0] far in front, [1] far behind, [2] invalid pixel (voxel*views, FP32 phase C),
                                        // [3] of [0]+[2]: settled by the validity-only phase C
};

// Per-view data the fast kernel gathers from, built by launch_prepare_views:
//   cls       float[n][H][W]   depth rounded to float, -1.0f exactly on the pixels that are invalid after
//                              the best-cost filter (never -1.0f on a valid pixel)
//   tileStats float[n][perView] per 8x8 tile of storage rows, as {max, min} pairs (float2): max of the valid depths
//                              (rounded up; -inf when the tile has no valid pixel, +inf when it holds a NaN) and
//                              min of the valid depths (rounded down; +inf when none, -inf with a NaN); and,
//                              badOff floats from the view's start, 1.0f when the tile holds an invalid pixel or a
//                              NaN (inside the image), else 0.0f
// The statistics form sparse tables per view: level l holds, AT EVERY TILE POSITION (x, y), the statistic of
// the window of 2^l x 2^l tiles whose corner is (x, y) (clipped by the image).  Any rectangle of tiles is
// the union of four overlapping windows of level floor(log2(longer side)): exact along the longer side,
// at most 2x over-covered along the shorter one -- tight bounds for a brick footprint of any size in 4 loads.
struct TilePyramid
{
  int nLevels;
  int tw, th;                     // tiles per row / column; every level has tw * th entries
  int perView;                    // floats per view: the {max, min} table, the bad table, the flag; multiple of 4
  int badOff;                     // the bad table starts here; level l of either table at l * tw * th entries
  int flagOff;                    // one float: 1.0f when some tile of the view is fully valid
};
TilePyramid tile_pyramid_layout(int W, int H);
// clsSpare: index (relative to d_cls) of a float that is set to -1.0f for the rejected voxels' gathers; < 0 = none
// Destinations of the view preparation: dst 0 is local; the others may be peer GPUs' buffers (same layout).
constexpr int kMaxPrepareDst = 8;
struct PrepareDst
{
  int n;
  int aligned;                    // every cls / lo pointer is 32-byte aligned
  float* cls[kMaxPrepareDst];
  int* lo[kMaxPrepareDst];        // all null or all set
  float* tiles[kMaxPrepareDst];
};
cudaError_t launch_prepare_views(const double* d_depths, const double* d_cost, double thr, int nViews, int W, int H,
                                 const PrepareDst& dst, long long clsSpare, cudaStream_t s, bool levels = true);
// d_lo (may be null): residual image of the lossless split depth = (cls, lo), see split_encode
cudaError_t launch_prepare_views(const double* d_depths, const double* d_cost, double thr, int nViews, int W, int H,
                                 float* d_cls, int* d_lo, long long clsSpare, float* d_tileDmax, cudaStream_t s, bool levels = true);
// every level of the tile statistics of views whose classification images were received from another GPU
cudaError_t launch_tile_stats_from_cls(const float* d_cls, int nViews, int W, int H, float* d_tileStats, cudaStream_t s);
// d_cls[clsSpare] (an index relative to d_cls) must hold -1.0f: launch_prepare_views(n views) writes it at n*W*H
// d_depths null: the double depths are rebuilt from (d_cls, d_lo)
cudaError_t launch_tsdf_fast(const GridParams& g, const FastChunk& c, const double* d_depths, const int* d_lo,
                             const float* d_cls, long long clsSpare, const float* d_tileDmax, bool cull,
                             ViewFast* d_viewScratch, unsigned* d_maskScratch, void* d_vol, int scalarType,
                             FastCounters* d_counters, int quota, cudaStream_t s);
size_t tsdf_fast_mask_bytes(const GridParams& g);   // size of d_maskScratch: per supertile a 64-bit view mask + a list entry, + work counters
void compose_fast_view(const GridParams& g, const double* K16, const double* RT16, int cxc, int cyc, ViewFast* out);
void fill_fast_chunk_constants(const GridParams& g, FastChunk* c);

cudaError_t launch_tsdf_exact(const GridParams& g, const ExactChunk& c, const double* d_depths,
                              void* d_vol, int scalarType, cudaStream_t s);
cudaError_t launch_depth_threshold(double* d_depths, const double* d_cost, size_t count, double thr,
                                   cudaStream_t s);

// coloration
struct alignas(16) ColorViewFast   // T1 form of one view: coefficients of the rows over (x, y, z), float
{
  float nx[4], ny[4], dz[4];       // [0..2] coefficients; [3] = sum of the |coefficients| (rounded up): bounds the local terms
};
struct ColorViewT2                 // T2 form: the same rows in double + margins
{
  double nx[4], ny[4], dz[4];
  double m2a, m2b, mza, mzb;
};
struct ColorViews
{
  const double* m;                 // T3: SoA m[e][v], e in 0..20 = RT rows 0..2 (12) then K 3x3 (9)
  const ColorViewFast* fast;
  const ColorViewT2* t2;
  int nViews;
  int stride;                      // padded view count of m
  int cxc, cyc;
  float T;                         // certified iff |e| < T - E * |r|
  float kE, kZ, U1;                // E = kE * ((|b_n| + A_n d) + U1 * (|b_z| + A_z d)), zm = kZ * (|b_z| + A_z d); b = the rows at
                                   // the batch's reference point, d = the batch's largest |coordinate offset| from it
};
float color_threshold_T(int W, int H);
void color_bound_constants(int W, int H, float* kE, float* kZ, float* U1);
void compose_color_view(const double* K16, const double* RT16, int cxc, int cyc, int W, int H,
                        ColorViewFast* fast, ColorViewT2* t2);
size_t colorize_scratch_bytes(size_t nPoints);      // d_scratch of launch_colorize: sort buckets + point permutation
cudaError_t launch_colorize(size_t nPoints, const void* d_xyz, int xyzType, ColorViews views,
                            const uint8_t* d_colors, int W, int H, uint8_t* d_mean, uint8_t* d_median,
                            int32_t* d_nb, void* d_scratch, cudaStream_t s);

// microbenchmarks
cudaError_t launch_fp_peak(int which, int blocks, int iters, float* d_sink, cudaStream_t s);
constexpr int kPeakThreads = 256;
constexpr int kPeakChains = 8;

}  // namespace dmi
