// Isosurface extraction on the GPU: the stage that FOLLOWS the integration in the reference's pipeline
// (Reconstruction/main.cxx:151-189: vtkCellDataToPointData -> vtkContourFilter(--contour, default 1.0) ->
// vtkTransformFilter(grid matrix)), so that the fused volume (8.6 GB at 1024^3) need not travel to the host for
// contouring and the mesh coloration can take its points straight from device memory.
//
//   cell -> point   a point's scalar = average of the cells sharing it, sum_j (1/n) v_j in double, cells visited in
//                   vtkStructuredData::GetPointCells' order
//   vertices        one per grid edge whose ends differ in (scalar >= value): t = (value - s0) / (s1 - s0),
//                   x = origin + (index + t) * spacing in double -> float32 (vtkPoints), grid matrix applied to the
//                   float32 point in double -> float32.  Numbered by owning point (k, j, i order), then axis.
//   triangles       per cell (k, j, i order) from a 256-case table GENERATED at first use: crossing edges joined face
//                   by face (a face with four crossings is cut so that each segment isolates one inside corner:
//                   neighbouring cells agree, the surface is watertight), loops fan-triangulated, normals from
//                   inside (>= value) to outside.  The vertex set is the one any edge-based contouring yields (VTK's
//                   synchronized templates included); the triangulation is this file's own, not VTK's.
// VTK is un-vendored and absent here: this stage is a restatement of published behaviour, checked against an independent
// numpy restatement kept with the tests (bit-identical vertices and triangles), not against VTK itself.
// Every arithmetic step uses explicit roundings (no FMA contraction), so the result does not depend on the compiler.
#include "dmi_ctx.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

constexpr int kMaxTri = 5;                    // checked when the table is generated
struct CaseTable
{
  unsigned char ntri[256];
  signed char edge[256][3 * kMaxTri];         // edge ids, 3 per triangle
};
__constant__ CaseTable c_cases;

// ---- the 256-case table ------------------------------------------------------------------------------------------
// corners: bit 0 = x, bit 1 = y, bit 2 = z.  edge id = 4 * axis + (a + 2 b), (a, b) = the other two coordinates in
// increasing axis order; the edge runs from corner lo to corner lo | (1 << axis).
void edge_ends(int e, int& lo, int& hi)
{
  const int axis = e / 4, o1 = axis == 0 ? 1 : 0, o2 = axis == 2 ? 1 : 2;
  const int a = (e % 4) & 1, b = (e % 4) >> 1;
  lo = (a << o1) | (b << o2);
  hi = lo | (1 << axis);
}

bool build_case_table(CaseTable& t)
{
  memset(&t, 0, sizeof(t));
  for (int cs = 0; cs < 256; cs++)
  {
    int nb[12][2], deg[12];
    bool cross[12];
    for (int e = 0; e < 12; e++)
    {
      int lo, hi;
      edge_ends(e, lo, hi);
      cross[e] = ((cs >> lo) & 1) != ((cs >> hi) & 1);
      deg[e] = 0;
    }
    auto link = [&](int a, int b) { nb[a][deg[a]++] = b; nb[b][deg[b]++] = a; };
    for (int axis = 0; axis < 3; axis++)
      for (int side = 0; side < 2; side++)
      {
        int es[4], n = 0;
        for (int e = 0; e < 12; e++)
        {
          int lo, hi;
          edge_ends(e, lo, hi);
          if (cross[e] && ((lo >> axis) & 1) == side && ((hi >> axis) & 1) == side) es[n++] = e;
        }
        if (n == 2) link(es[0], es[1]);
        else if (n == 4)
          for (int c = 0; c < 8; c++)             // ambiguous face: isolate each INSIDE corner of the face
            if (((c >> axis) & 1) == side && ((cs >> c) & 1))
            {
              int pr[2], m = 0;
              for (int q = 0; q < 4; q++) { int lo, hi; edge_ends(es[q], lo, hi); if (lo == c || hi == c) pr[m++] = es[q]; }
              link(pr[0], pr[1]);
            }
      }
    bool seen[12] = {false};
    int ntri = 0;
    for (int start = 0; start < 12; start++)
    {
      if (!cross[start] || seen[start]) continue;
      if (deg[start] != 2) return false;
      int loop[12], n = 0, prev = -1, cur = start;
      loop[n++] = start; seen[start] = true;
      for (;;)
      {
        const int a = nb[cur][0], b = nb[cur][1];
        int nxt = prev < 0 ? std::min(a, b) : (b == prev ? a : b);
        if (prev >= 0 && a == b) nxt = a;
        if (nxt == start) break;
        loop[n++] = nxt; seen[nxt] = true;
        prev = cur; cur = nxt;
      }
      // orientation: normals from inside to outside (edge midpoints as vertex positions)
      double p[12][3], nrm[3] = {0, 0, 0};
      for (int q = 0; q < n; q++)
      {
        int lo, hi;
        edge_ends(loop[q], lo, hi);
        for (int a = 0; a < 3; a++) p[q][a] = 0.5 * (((lo >> a) & 1) + ((hi >> a) & 1));
      }
      for (int q = 0; q < n; q++)
      {
        const double* u = p[q];
        const double* w = p[(q + 1) % n];
        nrm[0] += u[1] * w[2] - u[2] * w[1]; nrm[1] += u[2] * w[0] - u[0] * w[2]; nrm[2] += u[0] * w[1] - u[1] * w[0];
      }
      double s = 0;
      for (int q = 0; q < n; q++)
      {
        int lo, hi;
        edge_ends(loop[q], lo, hi);
        const int axis = loop[q] / 4;
        s += ((cs >> lo) & 1) ? nrm[axis] : -nrm[axis];
      }
      if (s == 0) return false;
      if (s < 0) std::reverse(loop + 1, loop + n);
      for (int q = 1; q + 1 < n; q++)
      {
        if (ntri >= kMaxTri) return false;
        t.edge[cs][3 * ntri] = (signed char)loop[0];
        t.edge[cs][3 * ntri + 1] = (signed char)loop[q];
        t.edge[cs][3 * ntri + 2] = (signed char)loop[q + 1];
        ntri++;
      }
    }
    t.ntri[cs] = (unsigned char)ntri;
  }
  return true;
}

// ---- kernels ------------------------------------------------------------------------------------------------------

struct ContourGrid
{
  int Nx, Ny, Nz;                 // cells
  double orig[3], sp[3], gm[12];
  double value;
};

// vtkStructuredData::GetPointCells' order of the (up to) 8 cells around a point
__constant__ int c_cellOff[8][3] = {{-1, 0, 0}, {-1, -1, 0}, {-1, -1, -1}, {-1, 0, -1}, {0, 0, 0}, {0, -1, 0}, {0, -1, -1}, {0, 0, -1}};

// linear index -> (i, j, k) of an nx x ny x nz lattice: two 32-bit divisions whenever the index fits (64-bit ones cost ~10x)
__device__ __forceinline__ void decode3(size_t e, unsigned nx, unsigned ny, int& i, int& j, int& k)
{
  if (e <= 0xffffffffull)
  {
    const unsigned u = (unsigned)e, r = u / nx, kk = r / ny;
    i = (int)(u - r * nx); j = (int)(r - kk * ny); k = (int)kk;
  }
  else
  {
    const size_t r = e / nx, kk = r / ny;
    i = (int)(e - r * nx); j = (int)(r - kk * ny); k = (int)kk;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
cell_to_point_kernel(const T* __restrict__ cells, double* __restrict__ pts, int Nx, int Ny, int Nz)
{
  const size_t px = (size_t)Nx + 1, py = (size_t)Ny + 1, n = px * py * ((size_t)Nz + 1);
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
  {
    int i, j, k;
    decode3(p, (unsigned)px, (unsigned)py, i, j, k);
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
      const int ci = i + c_cellOff[q][0], cj = j + c_cellOff[q][1], ck = k + c_cellOff[q][2];
      cnt += (ci >= 0 && ci < Nx && cj >= 0 && cj < Ny && ck >= 0 && ck < Nz) ? 1 : 0;
    }
    const double w = __ddiv_rn(1.0, (double)cnt);
    double c = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
      const int ci = i + c_cellOff[q][0], cj = j + c_cellOff[q][1], ck = k + c_cellOff[q][2];
      if (ci >= 0 && ci < Nx && cj >= 0 && cj < Ny && ck >= 0 && ck < Nz)
        c = __dadd_rn(c, __dmul_rn(w, (double)cells[((size_t)ck * Ny + cj) * Nx + ci]));
    }
    pts[p] = c;
  }
}

constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

// number of surface vertices owned by point p (its +x, +y, +z edges), as a 3-bit mask
__device__ __forceinline__ unsigned vertex_mask(const double* __restrict__ pts, const ContourGrid& g, size_t p)
{
  const size_t px = (size_t)g.Nx + 1, py = (size_t)g.Ny + 1;
  int i, j, k;
  decode3(p, (unsigned)px, (unsigned)py, i, j, k);
  const bool in0 = pts[p] >= g.value;
  unsigned m = 0;
  if (i < g.Nx && (pts[p + 1] >= g.value) != in0) m |= 1u;
  if (j < g.Ny && (pts[p + px] >= g.value) != in0) m |= 2u;
  if (k < g.Nz && (pts[p + px * py] >= g.value) != in0) m |= 4u;
  return m;
}

__device__ __forceinline__ unsigned cell_case(const double* __restrict__ pts, const ContourGrid& g, size_t c)
{
  const size_t px = (size_t)g.Nx + 1, py = (size_t)g.Ny + 1;
  int i, j, k;
  decode3(c, (unsigned)g.Nx, (unsigned)g.Ny, i, j, k);
  const size_t p = ((size_t)k * py + j) * px + i;
  unsigned cs = 0;
#pragma unroll
  for (int q = 0; q < 8; q++)
    cs |= (pts[p + (q & 1) + ((q >> 1) & 1) * px + ((q >> 2) & 1) * px * py] >= g.value) ? (1u << q) : 0u;
  return cs;
}

// WHAT = 0: elements are points, count = vertices owned; WHAT = 1: elements are cells, count = triangles
template <int WHAT>
__device__ __forceinline__ unsigned element_count(const double* __restrict__ pts, const ContourGrid& g, size_t e)
{
  return WHAT == 0 ? __popc(vertex_mask(pts, g, e)) : c_cases.ntri[cell_case(pts, g, e)];
}

template <int WHAT>
__global__ void __launch_bounds__(kScanThreads)
tile_sum_kernel(const double* __restrict__ pts, const __grid_constant__ ContourGrid g, size_t n, unsigned* __restrict__ tileSums)
{
  __shared__ unsigned s_w[kScanThreads / 32];
  const size_t base = (size_t)blockIdx.x * kScanTile;
  unsigned c = 0;
#pragma unroll
  for (int q = 0; q < kScanItems; q++)
  {
    const size_t e = base + (size_t)q * kScanThreads + threadIdx.x;
    if (e < n) c += element_count<WHAT>(pts, g, e);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned t = 0;
    for (int q = 0; q < kScanThreads / 32; q++) t += s_w[q];
    tileSums[blockIdx.x] = t;
  }
}

// exclusive prefix sum of the tile sums, in place, by ONE block; total -> *total
__global__ void __launch_bounds__(1024)
scan_tile_sums_kernel(unsigned* __restrict__ tileSums, size_t nTiles, unsigned long long* __restrict__ total)
{
  __shared__ unsigned s_w[32];
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (size_t b0 = 0; b0 < nTiles; b0 += 1024)
  {
    const size_t t = b0 + threadIdx.x;
    const unsigned v = t < nTiles ? tileSums[t] : 0u;
    unsigned incl = v;
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    unsigned before = 0;
    for (int q = 0; q < w; q++) before += s_w[q];
    const unsigned long long base = s_base;
    if (t < nTiles) tileSums[t] = (unsigned)(base + before + incl - v);
    __syncthreads();
    if (threadIdx.x == 1023) s_base = base + before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_base;
}

// exclusive offset of element e inside its tile (elements are laid out q-major: e = base + q * threads + tid, and the
// order of the OUTPUT must be the element order, so the scan runs over q-major positions)
template <int WHAT>
__device__ __forceinline__ unsigned tile_offsets(const double* __restrict__ pts, const ContourGrid& g, size_t n, size_t base,
                                                 unsigned (&cnt)[kScanItems], unsigned (&off)[kScanItems], unsigned* s_w)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned run = 0;
#pragma unroll
  for (int q = 0; q < kScanItems; q++)
  {
    const size_t e = base + (size_t)q * kScanThreads + threadIdx.x;
    cnt[q] = e < n ? element_count<WHAT>(pts, g, e) : 0u;
    unsigned incl = cnt[q];
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    unsigned before = 0, all = 0;
    for (int r = 0; r < kScanThreads / 32; r++) { if (r < w) before += s_w[r]; all += s_w[r]; }
    off[q] = run + before + incl - cnt[q];
    run += all;
    __syncthreads();
  }
  return run;
}

__global__ void __launch_bounds__(kScanThreads)
emit_vertices_kernel(const double* __restrict__ pts, const __grid_constant__ ContourGrid g, size_t n,
                     const unsigned* __restrict__ tileBase, unsigned* __restrict__ pointOffset, float* __restrict__ verts)
{
  __shared__ unsigned s_w[kScanThreads / 32];
  const size_t base = (size_t)blockIdx.x * kScanTile;
  unsigned cnt[kScanItems], off[kScanItems];
  tile_offsets<0>(pts, g, n, base, cnt, off, s_w);
  const unsigned tb = tileBase[blockIdx.x];
  const size_t px = (size_t)g.Nx + 1, py = (size_t)g.Ny + 1;
#pragma unroll
  for (int q = 0; q < kScanItems; q++)
  {
    const size_t p = base + (size_t)q * kScanThreads + threadIdx.x;
    if (p >= n) continue;
    unsigned o = tb + off[q];
    pointOffset[p] = o;
    if (cnt[q] == 0) continue;
    const unsigned mask = vertex_mask(pts, g, p);
    int idx[3];
    decode3(p, (unsigned)px, (unsigned)py, idx[0], idx[1], idx[2]);
    const size_t step[3] = {1, px, px * py};
    const double s0 = pts[p];
#pragma unroll
    for (int axis = 0; axis < 3; axis++)
    {
      if (!(mask & (1u << axis))) continue;
      const double s1 = pts[p + step[axis]];
      const double t = __ddiv_rn(__dsub_rn(g.value, s0), __dsub_rn(s1, s0));
      double x[3];
#pragma unroll
      for (int a = 0; a < 3; a++)
      {
        const double ia = a == axis ? __dadd_rn((double)idx[a], t) : (double)idx[a];
        x[a] = (double)__double2float_rn(__dadd_rn(g.orig[a], __dmul_rn(ia, g.sp[a])));          // vtkPoints: float32
      }
#pragma unroll
      for (int r = 0; r < 3; r++)
      {
        const double wv = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.gm[4 * r], x[0]), __dmul_rn(g.gm[4 * r + 1], x[1])),
                                              __dmul_rn(g.gm[4 * r + 2], x[2])), g.gm[4 * r + 3]);
        verts[3 * (size_t)o + r] = __double2float_rn(wv);
      }
      o++;
    }
  }
}

__global__ void __launch_bounds__(kScanThreads)
emit_triangles_kernel(const double* __restrict__ pts, const __grid_constant__ ContourGrid g, size_t n,
                      const unsigned* __restrict__ tileBase, const unsigned* __restrict__ pointOffset, int* __restrict__ tris)
{
  __shared__ unsigned s_w[kScanThreads / 32];
  const size_t base = (size_t)blockIdx.x * kScanTile;
  unsigned cnt[kScanItems], off[kScanItems];
  tile_offsets<1>(pts, g, n, base, cnt, off, s_w);
  const unsigned tb = tileBase[blockIdx.x];
  const size_t px = (size_t)g.Nx + 1, py = (size_t)g.Ny + 1;
#pragma unroll
  for (int q = 0; q < kScanItems; q++)
  {
    const size_t c = base + (size_t)q * kScanThreads + threadIdx.x;
    if (c >= n || cnt[q] == 0) continue;
    int i, j, k;
    decode3(c, (unsigned)g.Nx, (unsigned)g.Ny, i, j, k);
    const size_t p0 = ((size_t)k * py + j) * px + i;
    const unsigned cs = cell_case(pts, g, c);
    // vertex id of each of the cell's 12 edges that the case uses: owner point's offset + rank among its own edges
    size_t o = 3 * (size_t)(tb + off[q]);
    for (unsigned t = 0; t < cnt[q]; t++)
#pragma unroll
      for (int r = 0; r < 3; r++)
      {
        const int e = c_cases.edge[cs][3 * t + r];
        const int axis = e >> 2, a = e & 1, b = (e >> 1) & 1;
        const int o1 = axis == 0 ? 1 : 0, o2 = axis == 2 ? 1 : 2;
        const int lo = (a << o1) | (b << o2);
        const size_t owner = p0 + (lo & 1) + ((lo >> 1) & 1) * px + ((lo >> 2) & 1) * px * py;
        const unsigned mask = vertex_mask(pts, g, owner);
        tris[o++] = (int)(pointOffset[owner] + __popc(mask & ((1u << axis) - 1u)));
      }
  }
}

}  // namespace

struct dmi_contour_state
{
  DevBuf pts, pointOffset, tileSums, total, verts, tris;
  size_t nVerts = 0, nTris = 0;
  bool valid = false;
};

void dmi_host::contour_release(dmi_ctx* ctx)
{
  dmi_contour_state* s = ctx->contour;
  if (!s) return;
  s->pts.release(); s->pointOffset.release(); s->tileSums.release(); s->total.release(); s->verts.release(); s->tris.release();
  delete s;
  ctx->contour = nullptr;
}

namespace {

int upload_case_table(dmi_ctx* ctx)
{
  static std::mutex mu;
  static bool built = false, ok = false;
  static CaseTable table;
  static unsigned long long uploaded = 0;          // bit per device
  std::lock_guard<std::mutex> lock(mu);
  if (!built) { ok = build_case_table(table); built = true; }
  if (!ok) return ctx->fail(DMI_ERR_CUDA, "internal error: the contour case table could not be generated");
  if (ctx->device < 64 && (uploaded >> ctx->device) & 1ull) return DMI_OK;
  DMI_CK(cudaMemcpyToSymbol(c_cases, &table, sizeof(table)));
  if (ctx->device < 64) uploaded |= 1ull << ctx->device;
  return DMI_OK;
}

template <int WHAT>
int count_and_scan(dmi_ctx* ctx, dmi_contour_state* s, const ContourGrid& g, size_t n, unsigned long long* total)
{
  const size_t nTiles = (n + kScanTile - 1) / kScanTile;
  DMI_CK(s->tileSums.ensure(std::max<size_t>(4, nTiles * 4)));
  DMI_CK(s->total.ensure(8));
  tile_sum_kernel<WHAT><<<(unsigned)nTiles, kScanThreads, 0, ctx->stream>>>((const double*)s->pts.p, g, n, (unsigned*)s->tileSums.p);
  scan_tile_sums_kernel<<<1, 1024, 0, ctx->stream>>>((unsigned*)s->tileSums.p, nTiles, (unsigned long long*)s->total.p);
  DMI_CK(cudaGetLastError());
  ctx->total_launches += 2;
  DMI_CK(cudaMemcpyAsync(total, s->total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  return DMI_OK;
}

}  // namespace

extern "C" {

int dmi_contour_device(dmi_ctx* ctx, const void* d_cellScalars, int scalarType, double value, size_t* nVertices, size_t* nTriangles)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  const dmi::GridParams& gp = ctx->g;
  if (!d_cellScalars)
  {
    if (!ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
    DMI_REQUIRE(gp.layL == 0 && gp.k0 == 0 && gp.k1 == gp.Nz, "the context's own volume must cover the whole grid (gather the shards first)");
    d_cellScalars = ctx->vol.p;
    scalarType = ctx->vol_type;
  }
  DMI_REQUIRE(scalarType == DMI_F32 || scalarType == DMI_F64, "scalarType must be DMI_F32 or DMI_F64");
  DMI_REQUIRE(value == value, "the contour value is NaN");
  DMI_CK(cudaSetDevice(ctx->device));
  int rc = upload_case_table(ctx);
  if (rc != DMI_OK) return rc;
  if (!ctx->contour) ctx->contour = new dmi_contour_state();
  dmi_contour_state* s = ctx->contour;
  s->valid = false;
  ContourGrid g;
  g.Nx = gp.Nx; g.Ny = gp.Ny; g.Nz = gp.Nz;
  for (int a = 0; a < 3; a++) { g.orig[a] = gp.orig[a]; g.sp[a] = gp.sp[a]; }
  memcpy(g.gm, gp.gm, sizeof(g.gm));
  g.value = value;
  const size_t nPts = ((size_t)g.Nx + 1) * ((size_t)g.Ny + 1) * ((size_t)g.Nz + 1);
  const size_t nCells = (size_t)g.Nx * g.Ny * g.Nz;
  DMI_CK(s->pts.ensure(nPts * 8));
  DMI_CK(s->pointOffset.ensure(nPts * 4));
  const unsigned blocks = (unsigned)std::min<size_t>((nPts + 255) / 256, 148u * 32u);
  if (scalarType == DMI_F64)
    cell_to_point_kernel<double><<<blocks, 256, 0, ctx->stream>>>((const double*)d_cellScalars, (double*)s->pts.p, g.Nx, g.Ny, g.Nz);
  else
    cell_to_point_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float*)d_cellScalars, (double*)s->pts.p, g.Nx, g.Ny, g.Nz);
  DMI_CK(cudaGetLastError());
  ctx->total_launches++;

  unsigned long long nv = 0, nt = 0;
  rc = count_and_scan<0>(ctx, s, g, nPts, &nv);
  if (rc != DMI_OK) return rc;
  DMI_REQUIRE(nv < (1ull << 31), "more than 2^31 surface vertices");
  DMI_CK(s->verts.ensure(std::max<size_t>(12, (size_t)nv * 12)));
  emit_vertices_kernel<<<(unsigned)((nPts + kScanTile - 1) / kScanTile), kScanThreads, 0, ctx->stream>>>(
      (const double*)s->pts.p, g, nPts, (const unsigned*)s->tileSums.p, (unsigned*)s->pointOffset.p, (float*)s->verts.p);
  DMI_CK(cudaGetLastError());
  ctx->total_launches++;
  rc = count_and_scan<1>(ctx, s, g, nCells, &nt);
  if (rc != DMI_OK) return rc;
  DMI_REQUIRE(nt < (1ull << 31), "more than 2^31 triangles");
  DMI_CK(s->tris.ensure(std::max<size_t>(12, (size_t)nt * 12)));
  emit_triangles_kernel<<<(unsigned)((nCells + kScanTile - 1) / kScanTile), kScanThreads, 0, ctx->stream>>>(
      (const double*)s->pts.p, g, nCells, (const unsigned*)s->tileSums.p, (const unsigned*)s->pointOffset.p, (int*)s->tris.p);
  DMI_CK(cudaGetLastError());
  ctx->total_launches++;
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  s->nVerts = (size_t)nv; s->nTris = (size_t)nt; s->valid = true;
  if (nVertices) *nVertices = s->nVerts;
  if (nTriangles) *nTriangles = s->nTris;
  return DMI_OK;
}

int dmi_contour(dmi_ctx* ctx, const void* cellScalars, int scalarType, double value, size_t* nVertices, size_t* nTriangles)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  DMI_REQUIRE(cellScalars, "null argument");
  DMI_REQUIRE(scalarType == DMI_F32 || scalarType == DMI_F64, "scalarType must be DMI_F32 or DMI_F64");
  DMI_CK(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->g.Nx * ctx->g.Ny * ctx->g.Nz * (scalarType == DMI_F64 ? 8 : 4);
  DevBuf tmp;
  DMI_CK(tmp.ensure(std::max<size_t>(8, bytes)));
  cudaError_t e = cudaMemcpyAsync(tmp.p, cellScalars, bytes, cudaMemcpyHostToDevice, ctx->stream);
  int rc = e == cudaSuccess ? dmi_contour_device(ctx, tmp.p, scalarType, value, nVertices, nTriangles) : ctx->fail_cuda(e, "cudaMemcpyAsync");
  cudaStreamSynchronize(ctx->stream);
  tmp.release();
  return rc;
}

int dmi_contour_get(dmi_ctx* ctx, float* vertices, int32_t* triangles)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->contour || !ctx->contour->valid) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_contour has not been called");
  dmi_contour_state* s = ctx->contour;
  DMI_CK(cudaSetDevice(ctx->device));
  if (vertices && s->nVerts) DMI_CK(cudaMemcpyAsync(vertices, s->verts.p, s->nVerts * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (triangles && s->nTris) DMI_CK(cudaMemcpyAsync(triangles, s->tris.p, s->nTris * 12, cudaMemcpyDeviceToHost, ctx->stream));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  return DMI_OK;
}

int dmi_contour_device_ptr(dmi_ctx* ctx, const float** d_vertices, const int32_t** d_triangles, size_t* nVertices, size_t* nTriangles)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->contour || !ctx->contour->valid) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_contour has not been called");
  if (d_vertices) *d_vertices = (const float*)ctx->contour->verts.p;
  if (d_triangles) *d_triangles = (const int32_t*)ctx->contour->tris.p;
  if (nVertices) *nVertices = ctx->contour->nVerts;
  if (nTriangles) *nTriangles = ctx->contour->nTris;
  return DMI_OK;
}

}  // extern "C"
