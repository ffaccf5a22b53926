// C ABI of libdmi_b200.so (see include/dmi_b200.h).  Host-side orchestration only: contexts, device
// buffers, the double-buffered host->device view pipeline, chunking of views into kernel launches.
// Everything numeric happens in tsdf_kernels.cu / color_kernels.cu; there is no CPU fallback.
#include "dmi_ctx.cuh"

#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <thread>

namespace {

thread_local std::string g_create_error;

}  // namespace

extern "C" {

int dmi_abi_version(void) { return DMI_ABI_VERSION; }

int dmi_device_count(int* count)
{
  if (!count) return DMI_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; g_create_error = cudaGetErrorString(e); cudaGetLastError(); return DMI_ERR_CUDA; }
  return DMI_OK;
}

int dmi_create(int device, dmi_ctx** out)
{
  if (!out) return DMI_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
  {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return DMI_ERR_CUDA;
  }
  if (device < 0 || device >= n) { g_create_error = "device index out of range"; return DMI_ERR_INVALID_ARGUMENT; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return DMI_ERR_CUDA; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return DMI_ERR_CUDA; }
  if (prop.major != 10)
  {
    g_create_error = "libdmi_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return DMI_ERR_CUDA;
  }
  dmi_ctx* ctx = new dmi_ctx();
  ctx->device = device;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
  {
    g_create_error = cudaGetErrorString(e);
    delete ctx;
    return DMI_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  for (int b = 0; b < 2; b++)
  {
    cudaEventCreateWithFlags(&ctx->ev_ready[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming);
  }
  *out = ctx;
  return DMI_OK;
}

int dmi_destroy(dmi_ctx* ctx)
{
  if (!ctx) return DMI_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  ctx->vol.release();
  for (int b = 0; b < 2; b++)
  {
    ctx->stage_depth[b].release(); ctx->stage_cost[b].release();
    if (ctx->ev_ready[b]) cudaEventDestroy(ctx->ev_ready[b]);
    if (ctx->ev_free[b]) cudaEventDestroy(ctx->ev_free[b]);
  }
  ctx->filtered.release();
  ctx->counters.release();
  ctx->cls.release(); ctx->tiles.release(); ctx->viewscratch.release(); ctx->maskscratch.release();
  ctx->c_xyz.release(); ctx->c_colors.release(); ctx->c_mats.release();
  ctx->c_mean.release(); ctx->c_median.release(); ctx->c_nb.release(); ctx->c_sort.release();
  ctx->tsdf_stats.destroy(); ctx->color_stats.destroy();
  dmi_host::shard_release(ctx);
  dmi_host::contour_release(ctx);
  cudaStreamDestroy(ctx->own_stream);
  cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return DMI_OK;
}

const char* dmi_last_error(const dmi_ctx* ctx)
{
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int dmi_set_stream(dmi_ctx* ctx, void* cuda_stream)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  ctx->stream = (cudaStream_t)cuda_stream;
  return DMI_OK;
}

int dmi_use_own_stream(dmi_ctx* ctx)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  ctx->stream = ctx->own_stream;
  return DMI_OK;
}

int dmi_synchronize(dmi_ctx* ctx)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  DMI_CK(cudaSetDevice(ctx->device));
  DMI_CK(cudaStreamSynchronize(ctx->copy_stream));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  return DMI_OK;
}

int dmi_set_option(dmi_ctx* ctx, int option, long long value)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  switch (option)
  {
    case DMI_OPT_TSDF_KERNEL:
      DMI_REQUIRE(value == DMI_TSDF_KERNEL_AUTO || value == DMI_TSDF_KERNEL_EXACT, "unknown TSDF kernel id");
      ctx->opt_kernel = value; return DMI_OK;
    case DMI_OPT_CULL:
      ctx->opt_cull = value != 0; return DMI_OK;
    case DMI_OPT_BRICK_QUOTA:
      DMI_REQUIRE(value >= 1 && value <= (1 << 20), "brick quota must be between 1 and 2^20");
      ctx->opt_quota = (int)value; return DMI_OK;
    case DMI_OPT_TIER_COUNTERS:
      ctx->counters_on = value != 0;
      if (ctx->counters_on)
      {
        DMI_CK(cudaSetDevice(ctx->device));
        DMI_CK(ctx->counters.ensure(sizeof(dmi::FastCounters)));
        DMI_CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(dmi::FastCounters), ctx->stream));
      }
      return DMI_OK;
    case DMI_OPT_VIEW_CHUNK:
      DMI_REQUIRE(value >= 0, "view chunk must be >= 0");
      ctx->opt_chunk = value; return DMI_OK;
    default:
      return ctx->fail(DMI_ERR_INVALID_ARGUMENT, "unknown option");
  }
}

// ---- TSDF -------------------------------------------------------------------------------------

int dmi_initialize(dmi_ctx* ctx, const double gridMatrix[16], const int gridDims[3],
                   const double gridOrig[3], const double gridSpacing[3],
                   double thick, double rho, double eta, double delta, const int depthMapDims[2])
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  DMI_REQUIRE(gridMatrix && gridDims && gridOrig && gridSpacing && depthMapDims, "null argument");
  DMI_REQUIRE(gridDims[0] >= 2 && gridDims[1] >= 2 && gridDims[2] >= 2, "grid point dims must be >= 2 (at least one cell)");
  DMI_REQUIRE(depthMapDims[0] >= 1 && depthMapDims[1] >= 1, "depth map dims must be >= 1");
  DMI_REQUIRE((long long)depthMapDims[0] * depthMapDims[1] < (1ll << 31), "depth map too large for the reference's int pixel index");
  // the one parameter combination the reference filter rejects (vtkCudaReconstructionFilter.cxx:138-142)
  if (rho == 0 && thick == 0) return ctx->fail(DMI_ERR_BAD_PARAMETERS, "Ray potential Rho or Thickness or both have not been set");
  dmi::GridParams& g = ctx->g;
  memcpy(g.gm, gridMatrix, sizeof(double) * 12);
  for (int a = 0; a < 3; a++) { g.orig[a] = gridOrig[a]; g.sp[a] = gridSpacing[a]; }
  g.Nx = gridDims[0] - 1; g.Ny = gridDims[1] - 1; g.Nz = gridDims[2] - 1;
  g.k0 = 0; g.k1 = g.Nz;
  g.layL = 0; g.layStride = 1; g.layPhase = 0; g.nLocal = g.Nz;
  g.W = depthMapDims[0]; g.H = depthMapDims[1];
  g.thick = thick; g.rho = rho; g.eta = eta; g.delta = delta;
  g.rho_over_thick = rho / thick;
  g.neg_eta_rho = -eta * rho;
  ctx->initialized = true;
  ctx->vol_active = false;
  return DMI_OK;
}

int dmi_set_slab(dmi_ctx* ctx, int k0, int k1)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  DMI_REQUIRE(0 <= k0 && k0 <= k1 && k1 <= ctx->g.Nz, "slab must satisfy 0 <= k0 <= k1 <= Nz");
  ctx->g.k0 = k0; ctx->g.k1 = k1;
  ctx->g.layL = 0; ctx->g.layStride = 1; ctx->g.layPhase = 0; ctx->g.nLocal = k1 - k0;
  ctx->vol_active = false;
  return DMI_OK;
}

int dmi_set_slab_layers(dmi_ctx* ctx, int layerPlanes, int phase, int stride)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  DMI_REQUIRE(layerPlanes >= 32 && layerPlanes % 32 == 0, "layerPlanes must be a positive multiple of 32");
  DMI_REQUIRE(stride >= 1 && phase >= 0 && phase < stride, "need 0 <= phase < stride");
  dmi::GridParams& g = ctx->g;
  g.k0 = 0; g.k1 = g.Nz;
  g.layL = layerPlanes; g.layStride = stride; g.layPhase = phase;
  long long n = 0;
  for (long long start = (long long)phase * layerPlanes; start < g.Nz; start += (long long)stride * layerPlanes)
    n += std::min<long long>(layerPlanes, g.Nz - start);
  g.nLocal = (int)n;
  ctx->vol_active = false;
  return DMI_OK;
}

int dmi_slab_planes(dmi_ctx* ctx, int* planes)
{
  if (!ctx || !planes) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  *planes = ctx->g.nLocal;
  return DMI_OK;
}

}  // extern "C"
namespace dmi_host {
void set_create_error(const std::string& msg) { g_create_error = msg; }
size_t slab_cells(const dmi::GridParams& g)
{
  return (size_t)g.Nx * g.Ny * (size_t)g.nLocal;
}
}  // namespace dmi_host
using dmi_host::slab_cells;
extern "C" {

// True when every BIT of the host array is zero (-0.0 is not: s + t keeps its sign when nothing is added).
// The reference's filter zero-fills its output before the call (vtkCudaReconstructionFilter.cxx:133), so
// the usual io_scalar needs no upload: a parallel scan at memory speed is several times cheaper than PCIe.
static bool host_bytes_all_zero(const void* p, size_t bytes)
{
  const unsigned char* b = static_cast<const unsigned char*>(p);
  size_t head = 0;
  while (head < bytes && (reinterpret_cast<uintptr_t>(b + head) & 7)) { if (b[head]) return false; head++; }
  const unsigned long long* w = reinterpret_cast<const unsigned long long*>(b + head);
  const size_t words = (bytes - head) / 8;
  for (size_t q = head + words * 8; q < bytes; q++) if (b[q]) return false;
  const size_t block = 1u << 17;                                   // 1 MB of words between looks at the flag
  unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (words < (size_t)nt * block) nt = 1;
  std::atomic<bool> nonzero(false);
  auto scan = [&](size_t w0, size_t w1) {
    for (size_t q = w0; q < w1 && !nonzero.load(std::memory_order_relaxed); q += block)
    {
      unsigned long long acc = 0;
      const size_t e = std::min(w1, q + block);
      for (size_t r = q; r < e; r++) acc |= w[r];
      if (acc) nonzero.store(true, std::memory_order_relaxed);
    }
  };
  if (nt == 1) scan(0, words);
  else
  {
    // no exception may cross the C ABI: a thread that cannot be created leaves its share to the caller's thread
    std::vector<std::thread> th;
    const size_t per = (words + nt - 1) / nt;
    unsigned started = 0;
    try
    {
      th.reserve(nt);
      for (; started < nt; started++)
        th.emplace_back(scan, std::min(words, started * per), std::min(words, (started + 1) * per));
    }
    catch (...) {}
    for (unsigned t = started; t < nt; t++) scan(std::min(words, t * per), std::min(words, (t + 1) * per));
    for (auto& t : th) t.join();
  }
  return !nonzero.load();
}

// Cheap necessary condition for "all zero": both ends and 4096 words spread over the array.
static bool host_zero_probe(const void* p, size_t bytes)
{
  const unsigned char* b = static_cast<const unsigned char*>(p);
  const size_t edge = std::min<size_t>(bytes, 65536);
  for (size_t q = 0; q < edge; q++) if (b[q] | b[bytes - 1 - q]) return false;
  const size_t stride = std::max<size_t>(1, bytes / 4096);
  for (size_t q = 0; q < bytes; q += stride) if (b[q]) return false;
  return true;
}

enum VolumeInit { kVolScan = 0, kVolUpload = 1, kVolAssumeZero = 2 };

static int volume_begin_impl(dmi_ctx* ctx, const void* h_scalar, int scalarType, VolumeInit mode);

int dmi_volume_begin(dmi_ctx* ctx, const void* h_scalar, int scalarType)
{
  return volume_begin_impl(ctx, h_scalar, scalarType, kVolScan);
}

static int volume_begin_impl(dmi_ctx* ctx, const void* h_scalar, int scalarType, VolumeInit mode)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  DMI_REQUIRE(scalarType == DMI_F32 || scalarType == DMI_F64, "scalarType must be DMI_F32 or DMI_F64");
  DMI_CK(cudaSetDevice(ctx->device));
  const size_t bytes = slab_cells(ctx->g) * (scalarType == DMI_F64 ? 8 : 4);
  DMI_CK(ctx->vol.ensure(bytes > 0 ? bytes : 8));
  ctx->vol_bytes = bytes;
  ctx->vol_type = scalarType;
  if (bytes)
  {
    const bool upload = h_scalar && (mode == kVolUpload ||
                                     (mode == kVolScan && !(host_zero_probe(h_scalar, bytes) && host_bytes_all_zero(h_scalar, bytes))));
    if (upload)
      DMI_CK(cudaMemcpyAsync(ctx->vol.p, h_scalar, bytes, cudaMemcpyHostToDevice, ctx->stream));
    else DMI_CK(cudaMemsetAsync(ctx->vol.p, 0, bytes, ctx->stream));
  }
  ctx->vol_active = true;
  return DMI_OK;
}

}  // extern "C"
bool dmi_host::fast_path_applies(const dmi_ctx* ctx)
{
  // The certified fast path needs the regime the reference's CLI enforces (0 < Thick, finite
  // parameters; Reconstruction/main.cxx:270-271) and pixel / voxel coordinates that floats hold
  // exactly; anything else goes through the exact kernel.
  const dmi::GridParams& g = ctx->g;
  // Delta >= 0: the brick tests and phase C compare |diff| with Delta + margin, which is only the reference's
  // `a > Delta` for a non-negative Delta (for Delta < 0 the reference always takes its first branch)
  return ctx->opt_kernel == DMI_TSDF_KERNEL_AUTO && g.thick > 0 && std::isfinite(g.rho_over_thick) &&
         std::isfinite(g.delta) && g.delta >= 0 && std::isfinite(g.neg_eta_rho) && std::isfinite(g.rho) && g.W < (1 << 21) &&
         g.H < (1 << 21) && g.Nx < (1 << 23) && g.Ny < (1 << 23) && g.Nz < (1 << 23);
}
using dmi_host::fast_path_applies;
extern "C" {

// Exact kernel over views whose depth maps at d_depths are ALREADY filtered.
static int integrate_exact_resident(dmi_ctx* ctx, int nViews, const double* d_depths, const double* K, const double* RT)
{
  const dmi::GridParams& g = ctx->g;
  const size_t npix = (size_t)g.W * g.H;
  int chunk = dmi::kExactChunk;
  if (ctx->opt_chunk > 0 && ctx->opt_chunk < chunk) chunk = (int)ctx->opt_chunk;
  for (int v0 = 0; v0 < nViews; v0 += chunk)
  {
    dmi::ExactChunk c;
    c.n = std::min(chunk, nViews - v0);
    c.pad = 0;
    for (int q = 0; q < c.n; q++)
    {
      memcpy(c.v[q].RT, RT + 16 * (size_t)(v0 + q), sizeof(double) * 12);
      memcpy(c.v[q].K, K + 16 * (size_t)(v0 + q), sizeof(double) * 12);
    }
    DMI_CK(dmi::launch_tsdf_exact(g, c, d_depths + npix * v0, ctx->vol.p, ctx->vol_type, ctx->stream));
    ctx->tsdf_stats.launches++;
    ctx->total_launches++;
  }
  return DMI_OK;
}

// Fast kernel over views that are ALREADY prepared (classification images + tile pyramids, see
// launch_prepare_views): composes the per-view rows and launches chunk by chunk.
}  // extern "C"
int dmi_host::integrate_fast_prepared(dmi_ctx* ctx, int nViews, const double* d_depths, const int* d_lo, const float* d_cls, long long clsSpare,
                                      const float* d_tiles, const double* K, const double* RT)
{
  const dmi::GridParams& g = ctx->g;
  const size_t npix = (size_t)g.W * g.H;
  const size_t tilesPerView = (size_t)dmi::tile_pyramid_layout(g.W, g.H).perView;
  int chunk = dmi::kFastChunk;
  if (ctx->opt_chunk > 0 && ctx->opt_chunk < chunk) chunk = (int)ctx->opt_chunk;
  // the kernel addresses the spare -1.0f slot as a 32-bit offset from the last storage row of each view
  {
    const long long lastRow = (long long)(g.H - 1) * g.W;
    const long long hi = clsSpare - lastRow, lo = clsSpare - ((long long)npix * (nViews - 1) + lastRow);
    DMI_REQUIRE(hi <= 2147483647ll && lo >= -2147483647ll,
                "the spare classification slot must lie within 2^31 floats of every view of the call: split the call");
  }
  DMI_CK(ctx->viewscratch.ensure(sizeof(dmi::ViewFast) * dmi::kFastChunk));
  DMI_CK(ctx->maskscratch.ensure(dmi::tsdf_fast_mask_bytes(g)));
  dmi::FastChunk* c = &ctx->fast_chunk;
  dmi::fill_fast_chunk_constants(g, c);
  for (int v0 = 0; v0 < nViews; v0 += chunk)
  {
    c->n = std::min(chunk, nViews - v0);
    c->pinhole = 1;
    for (int q = 0; q < c->n; q++)
    {
      const double* k16 = K + 16 * (size_t)(v0 + q);
      const double* rt16 = RT + 16 * (size_t)(v0 + q);
      dmi::compose_fast_view(g, k16, rt16, c->cxc, c->cyc, &c->v[q]);
      memcpy(c->e[q].RT, rt16, sizeof(double) * 12);
      memcpy(c->e[q].K, k16, sizeof(double) * 12);
      if (!(k16[8] == 0.0 && k16[9] == 0.0 && k16[10] == 1.0 && k16[11] == 0.0)) c->pinhole = 0;
    }
    DMI_CK(dmi::launch_tsdf_fast(g, *c, d_depths ? d_depths + npix * v0 : nullptr, d_lo ? d_lo + npix * v0 : nullptr, d_cls + npix * v0, clsSpare - (long long)(npix * (size_t)v0),
                                 d_tiles + tilesPerView * v0, ctx->opt_cull, (dmi::ViewFast*)ctx->viewscratch.p,
                                 (unsigned*)ctx->maskscratch.p, ctx->vol.p, ctx->vol_type,
                                 ctx->counters_on ? (dmi::FastCounters*)ctx->counters.p : nullptr, ctx->opt_quota, ctx->stream));
    ctx->tsdf_stats.launches++;
    ctx->total_launches += ctx->opt_cull ? 4 : 3;         // view staging, supertile culling, compaction, integration
  }
  return DMI_OK;
}

using dmi_host::integrate_fast_prepared;
extern "C" {

// Fast kernel over views given as the caller's depth maps + optional best-cost maps (neither modified):
// the filter is folded into the float classification image built by the view-preparation kernel.
static int integrate_fast_resident(dmi_ctx* ctx, int nViews, const double* d_depths, const double* d_cost, double thr,
                                   const double* K, const double* RT)
{
  const dmi::GridParams& g = ctx->g;
  const size_t npix = (size_t)g.W * g.H;
  const size_t tilesPerView = (size_t)dmi::tile_pyramid_layout(g.W, g.H).perView;
  const int chunk = dmi::kFastChunk;
  // views prepared at a time: whole chunks, about 1 GB of classification image
  int group = (int)std::max<size_t>(1, (1ull << 30) / (npix * 4));
  group = std::max(chunk, group / chunk * chunk);
  group = std::min(group, (nViews + chunk - 1) / chunk * chunk);
  DMI_CK(ctx->cls.ensure(((size_t)group * npix + 1) * 4));   // + the spare slot holding -1.0f
  DMI_CK(ctx->tiles.ensure((size_t)group * tilesPerView * 4));
  for (int g0 = 0; g0 < nViews; g0 += group)
  {
    const int gn = std::min(group, nViews - g0);
    DMI_CK(dmi::launch_prepare_views(d_depths + npix * g0, d_cost ? d_cost + npix * g0 : nullptr, thr, gn, g.W, g.H,
                                     (float*)ctx->cls.p, nullptr, (long long)(npix * (size_t)gn), (float*)ctx->tiles.p, ctx->stream));
    ctx->total_launches += 1 + dmi::tile_pyramid_layout(ctx->g.W, ctx->g.H).nLevels;   // level 0, upper levels, flag
    int rc = integrate_fast_prepared(ctx, gn, d_depths + npix * g0, nullptr, (const float*)ctx->cls.p, (long long)(npix * (size_t)gn),
                                     (const float*)ctx->tiles.p, K + 16 * (size_t)g0, RT + 16 * (size_t)g0);
    if (rc != DMI_OK) return rc;
  }
  return DMI_OK;
}

// Integrates nViews views resident on the device.  d_cost may be null.  When `scratch_ok` the depth
// buffer belongs to the library (a staging slot) and may be filtered in place.
static int integrate_device_views(dmi_ctx* ctx, int nViews, const double* d_depths, const double* d_cost, double thr,
                                  const double* K, const double* RT, bool scratch_ok)
{
  const dmi::GridParams& g = ctx->g;
  if (slab_cells(g) == 0) return DMI_OK;
  const size_t npix = (size_t)g.W * g.H;
  EventSpan span = ctx->tsdf_stats.open();
  DMI_CK(cudaEventRecord(span.a, ctx->stream));
  int rc = DMI_OK;
  if (fast_path_applies(ctx))
    rc = integrate_fast_resident(ctx, nViews, d_depths, d_cost, thr, K, RT);
  else if (!d_cost)
    rc = integrate_exact_resident(ctx, nViews, d_depths, K, RT);
  else if (scratch_ok)
  {
    DMI_CK(dmi::launch_depth_threshold(const_cast<double*>(d_depths), d_cost, (size_t)nViews * npix, thr, ctx->stream));
    ctx->total_launches++;
    rc = integrate_exact_resident(ctx, nViews, d_depths, K, RT);
  }
  else
  {
    // the caller's depth buffer is const: filter into scratch, a bounded number of views at a time
    const int step = (int)std::max<size_t>(1, std::min<size_t>((size_t)nViews, (512ull << 20) / (npix * 8)));
    DMI_CK(ctx->filtered.ensure((size_t)step * npix * 8));
    for (int v0 = 0; v0 < nViews && rc == DMI_OK; v0 += step)
    {
      const int n = std::min(step, nViews - v0);
      DMI_CK(cudaMemcpyAsync(ctx->filtered.p, d_depths + npix * v0, (size_t)n * npix * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      DMI_CK(dmi::launch_depth_threshold((double*)ctx->filtered.p, d_cost + npix * v0, (size_t)n * npix, thr, ctx->stream));
      ctx->total_launches++;
      rc = integrate_exact_resident(ctx, n, (const double*)ctx->filtered.p, K + 16 * (size_t)v0, RT + 16 * (size_t)v0);
    }
  }
  if (rc != DMI_OK) return rc;
  DMI_CK(cudaEventRecord(span.b, ctx->stream));
  ctx->tsdf_stats.pending.push_back(span);
  return DMI_OK;
}

int dmi_volume_integrate_device(dmi_ctx* ctx, int nViews, const double* d_depths, const double* d_bestCost,
                                double thresholdBestCost, const double* K, const double* RT)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized || !ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "no depthMap or KRTD matrix have been loaded");
  DMI_REQUIRE(d_depths && K && RT, "null argument");
  DMI_CK(cudaSetDevice(ctx->device));
  return integrate_device_views(ctx, nViews, d_depths, d_bestCost, thresholdBestCost, K, RT, false);
}

int dmi_prepared_view_sizes(dmi_ctx* ctx, size_t* clsFloatsPerView, size_t* tileFloatsPerView)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  if (clsFloatsPerView) *clsFloatsPerView = (size_t)ctx->g.W * ctx->g.H;
  if (tileFloatsPerView) *tileFloatsPerView = (size_t)dmi::tile_pyramid_layout(ctx->g.W, ctx->g.H).perView;
  return DMI_OK;
}

int dmi_prepare_views_device(dmi_ctx* ctx, int nViews, const double* d_depths, const double* d_bestCost,
                             double thresholdBestCost, float* d_cls, int* d_lo, long long clsSpareIndex, float* d_tileStats)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "no depthMap or KRTD matrix have been loaded");
  DMI_REQUIRE(d_depths && d_cls && d_tileStats, "null argument");
  DMI_REQUIRE((reinterpret_cast<uintptr_t>(d_tileStats) & 15) == 0, "d_tileStats must be 16-byte aligned");
  DMI_CK(cudaSetDevice(ctx->device));
  DMI_CK(dmi::launch_prepare_views(d_depths, d_bestCost, thresholdBestCost, nViews, ctx->g.W, ctx->g.H, d_cls, d_lo, clsSpareIndex,
                                   d_tileStats, ctx->stream));
  ctx->total_launches += 1 + dmi::tile_pyramid_layout(ctx->g.W, ctx->g.H).nLevels;   // level 0, upper levels, flag
  return DMI_OK;
}

int dmi_volume_integrate_prepared(dmi_ctx* ctx, int nViews, const double* d_depths, const int* d_lo, const float* d_cls,
                                  long long clsSpareIndex, const float* d_tileStats, const double* K, const double* RT)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized || !ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "no depthMap or KRTD matrix have been loaded");
  DMI_REQUIRE((d_depths || d_lo) && d_cls && d_tileStats && K && RT, "null argument");
  DMI_REQUIRE((reinterpret_cast<uintptr_t>(d_tileStats) & 15) == 0, "d_tileStats must be 16-byte aligned");
  if (!fast_path_applies(ctx))
    return ctx->fail(DMI_ERR_BAD_PARAMETERS, "prepared views need the certified fast path (0 < Thick, finite parameters, kernel AUTO)");
  DMI_CK(cudaSetDevice(ctx->device));
  if (slab_cells(ctx->g) == 0) return DMI_OK;
  EventSpan span = ctx->tsdf_stats.open();
  DMI_CK(cudaEventRecord(span.a, ctx->stream));
  int rc = integrate_fast_prepared(ctx, nViews, d_depths, d_depths ? nullptr : d_lo, d_cls, clsSpareIndex, d_tileStats, K, RT);
  if (rc != DMI_OK) return rc;
  DMI_CK(cudaEventRecord(span.b, ctx->stream));
  ctx->tsdf_stats.pending.push_back(span);
  return DMI_OK;
}

int dmi_volume_integrate_host(dmi_ctx* ctx, int nViews, const double* depths, const double* bestCost,
                              double thresholdBestCost, const double* K, const double* RT)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized || !ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "no depthMap or KRTD matrix have been loaded");
  DMI_REQUIRE(depths && K && RT, "null argument");
  DMI_CK(cudaSetDevice(ctx->device));
  const size_t npix = (size_t)ctx->g.W * ctx->g.H;
  // staging slot = as many views as fit in 256 MB, at least 1, at most 64
  const int step = (int)std::max<size_t>(1, std::min<size_t>(64, (256ull << 20) / (npix * 8)));
  int slot = 0;
  for (int v0 = 0; v0 < nViews; v0 += step, slot ^= 1)
  {
    const int n = std::min(step, nViews - v0);
    const size_t bytes = (size_t)n * npix * 8;
    DMI_CK(ctx->stage_depth[slot].ensure((size_t)step * npix * 8));
    if (bestCost) DMI_CK(ctx->stage_cost[slot].ensure((size_t)step * npix * 8));
    // the copy engine may only overwrite the slot once the kernels that read it are done
    if (ctx->slot_used[slot]) DMI_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[slot], 0));
    DMI_CK(cudaMemcpyAsync(ctx->stage_depth[slot].p, depths + npix * v0, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (bestCost)
      DMI_CK(cudaMemcpyAsync(ctx->stage_cost[slot].p, bestCost + npix * v0, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    DMI_CK(cudaEventRecord(ctx->ev_ready[slot], ctx->copy_stream));
    DMI_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_ready[slot], 0));
    int rc = integrate_device_views(ctx, n, (const double*)ctx->stage_depth[slot].p,
                                    bestCost ? (const double*)ctx->stage_cost[slot].p : nullptr, thresholdBestCost,
                                    K + 16 * (size_t)v0, RT + 16 * (size_t)v0, true);
    if (rc != DMI_OK) return rc;
    DMI_CK(cudaEventRecord(ctx->ev_free[slot], ctx->stream));
    ctx->slot_used[slot] = true;
  }
  // host pointers: synchronous at the ABI
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  return DMI_OK;
}

int dmi_volume_end(dmi_ctx* ctx, void* h_scalar)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized || !ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
  DMI_CK(cudaSetDevice(ctx->device));
  if (h_scalar && ctx->vol_bytes)
    DMI_CK(cudaMemcpyAsync(h_scalar, ctx->vol.p, ctx->vol_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  return DMI_OK;
}

int dmi_volume_device_ptr(dmi_ctx* ctx, void** d_ptr, size_t* bytes)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized || !ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
  if (d_ptr) *d_ptr = ctx->vol.p;
  if (bytes) *bytes = ctx->vol_bytes;
  return DMI_OK;
}

int dmi_process_depth_maps(dmi_ctx* ctx, int nViews, const double* depths, const double* bestCost,
                           double thresholdBestCost, const double* K, const double* RT,
                           void* io_scalar, int scalarType)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_initialize has not been called");
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "no depthMap or KRTD matrix have been loaded");
  DMI_REQUIRE(depths && K && RT && io_scalar, "null argument");
  DMI_REQUIRE(scalarType == DMI_F32 || scalarType == DMI_F64, "scalarType must be DMI_F32 or DMI_F64");
  // A large io_scalar that looks all-zero (the filter zero-fills it, vtkCudaReconstructionFilter.cxx:133) is
  // verified by a full host scan WHILE the views stream in and integrate onto a zeroed device volume; should
  // the scan find a set bit after all, the pass is repeated from the uploaded io_scalar (inputs are still here).
  const size_t bytes = slab_cells(ctx->g) * (scalarType == DMI_F64 ? 8 : 4);
  bool speculate = bytes >= (256u << 20) && host_zero_probe(io_scalar, bytes);
  bool zero = true;
  std::thread scan;
  if (speculate)
  {
    try { scan = std::thread([&] { zero = host_bytes_all_zero(io_scalar, bytes); }); }
    catch (...) { speculate = false; }               // no thread: scan synchronously inside volume_begin
  }
  int rc = volume_begin_impl(ctx, io_scalar, scalarType, speculate ? kVolAssumeZero : kVolScan);
  if (rc == DMI_OK) rc = dmi_volume_integrate_host(ctx, nViews, depths, bestCost, thresholdBestCost, K, RT);
  if (speculate) scan.join();
  if (rc != DMI_OK) return rc;
  if (speculate && !zero)
  {
    rc = volume_begin_impl(ctx, io_scalar, scalarType, kVolUpload);
    if (rc == DMI_OK) rc = dmi_volume_integrate_host(ctx, nViews, depths, bestCost, thresholdBestCost, K, RT);
    if (rc != DMI_OK) return rc;
  }
  return dmi_volume_end(ctx, io_scalar);
}

int dmi_apply_depth_threshold_device(dmi_ctx* ctx, size_t count, double* d_depths, const double* d_bestCost,
                                     double thresholdBestCost)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  DMI_REQUIRE(count == 0 || (d_depths && d_bestCost), "null argument");
  DMI_CK(cudaSetDevice(ctx->device));
  DMI_CK(dmi::launch_depth_threshold(d_depths, d_bestCost, count, thresholdBestCost, ctx->stream));
  if (count) ctx->total_launches++;
  return DMI_OK;
}

int dmi_tsdf_kernel_stats(dmi_ctx* ctx, float* ms, long long* launches)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  DMI_CK(cudaSetDevice(ctx->device));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  if (ms) *ms = ctx->tsdf_stats.drain(); else ctx->tsdf_stats.drain();
  if (launches) *launches = ctx->tsdf_stats.launches;
  ctx->tsdf_stats.launches = 0;
  return DMI_OK;
}

int dmi_tsdf_tier_counters(dmi_ctx* ctx, unsigned long long out[16])
{
  if (!ctx || !out) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->counters_on) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "enable DMI_OPT_TIER_COUNTERS first");
  DMI_CK(cudaSetDevice(ctx->device));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  DMI_CK(cudaMemcpy(out, ctx->counters.p, sizeof(dmi::FastCounters), cudaMemcpyDeviceToHost));
  DMI_CK(cudaMemset(ctx->counters.p, 0, sizeof(dmi::FastCounters)));
  return DMI_OK;
}

// ---- coloration ---------------------------------------------------------------------------------

static int pack_color_views(dmi_ctx* ctx, int nViews, const double* K, const double* RT, int W, int H, dmi::ColorViews* out)
{
  const int stride = (nViews + 31) & ~31;
  const size_t exactBytes = (size_t)21 * stride * 8;
  const size_t fastOff = (exactBytes + 15) & ~(size_t)15;
  const size_t fastBytes = sizeof(dmi::ColorViewFast) * (size_t)nViews;
  const size_t t2Off = (fastOff + fastBytes + 15) & ~(size_t)15;
  const size_t total = t2Off + sizeof(dmi::ColorViewT2) * (size_t)nViews;
  std::vector<unsigned char> host(total, 0);
  double* m = reinterpret_cast<double*>(host.data());
  dmi::ColorViewFast* fast = reinterpret_cast<dmi::ColorViewFast*>(host.data() + fastOff);
  dmi::ColorViewT2* t2 = reinterpret_cast<dmi::ColorViewT2*>(host.data() + t2Off);
  const int cxc = W / 2, cyc = H / 2;
  for (int v = 0; v < nViews; v++)
  {
    for (int e = 0; e < 12; e++) m[(size_t)e * stride + v] = RT[16 * (size_t)v + e];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) m[(size_t)(12 + r * 3 + c) * stride + v] = K[16 * (size_t)v + r * 4 + c];
    dmi::compose_color_view(K + 16 * (size_t)v, RT + 16 * (size_t)v, cxc, cyc, W, H, &fast[v], &t2[v]);
  }
  DMI_CK(ctx->c_mats.ensure(total));
  // pageable source: the copy has consumed `host` by the time cudaMemcpyAsync returns
  DMI_CK(cudaMemcpyAsync(ctx->c_mats.p, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  const unsigned char* base = (const unsigned char*)ctx->c_mats.p;
  out->m = (const double*)base;
  out->fast = (const dmi::ColorViewFast*)(base + fastOff);
  out->t2 = (const dmi::ColorViewT2*)(base + t2Off);
  out->nViews = nViews;
  out->stride = stride;
  out->cxc = cxc;
  out->cyc = cyc;
  out->T = dmi::color_threshold_T(W, H);
  dmi::color_bound_constants(W, H, &out->kE, &out->kZ, &out->U1);
  return DMI_OK;
}

int dmi_colorize_device(dmi_ctx* ctx, size_t nPoints, const void* d_xyz, int xyzType, int nViews,
                        const uint8_t* d_colors, const double* K, const double* RT, int W, int H,
                        uint8_t* d_mean, uint8_t* d_median, int32_t* d_nb)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  // MeshColoration::ProcessColoration returns false without views (MeshColoration.cxx:102-106)
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "Error when input has been set or during reading vti/krtd file path");
  DMI_REQUIRE(xyzType == DMI_F32 || xyzType == DMI_F64, "xyzType must be DMI_F32 or DMI_F64");
  DMI_REQUIRE(W >= 1 && H >= 1, "image dims must be >= 1");
  DMI_REQUIRE(K && RT, "null argument");
  if (nPoints == 0) return DMI_OK;
  DMI_REQUIRE(d_xyz && d_colors && d_mean && d_median && d_nb, "null argument");
  DMI_CK(cudaSetDevice(ctx->device));
  dmi::ColorViews views;
  DMI_REQUIRE(W < (1 << 21) && H < (1 << 21), "image dims must be below 2^21");
  DMI_REQUIRE((long long)W * H < (1ll << 31), "image too large for the 32-bit pixel index");
  int rc = pack_color_views(ctx, nViews, K, RT, W, H, &views);
  if (rc != DMI_OK) return rc;
  EventSpan span = ctx->color_stats.open();
  DMI_CK(cudaEventRecord(span.a, ctx->stream));
  DMI_CK(ctx->c_sort.ensure(dmi::colorize_scratch_bytes(nPoints)));
  DMI_CK(dmi::launch_colorize(nPoints, d_xyz, xyzType, views, d_colors, W, H, d_mean, d_median, d_nb, ctx->c_sort.p, ctx->stream));
  DMI_CK(cudaEventRecord(span.b, ctx->stream));
  ctx->color_stats.pending.push_back(span);
  ctx->color_stats.launches++;
  ctx->total_launches += 6;                                  // bounding box, bucket counts, scan, scatter, refine, coloration
  return DMI_OK;
}

int dmi_colorize(dmi_ctx* ctx, size_t nPoints, const void* xyz, int xyzType, int nViews,
                 const uint8_t* colors, const double* K, const double* RT, int W, int H,
                 uint8_t* mean, uint8_t* median, int32_t* nbProjected)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "Error when input has been set or during reading vti/krtd file path");
  DMI_REQUIRE(xyzType == DMI_F32 || xyzType == DMI_F64, "xyzType must be DMI_F32 or DMI_F64");
  DMI_REQUIRE(W >= 1 && H >= 1, "image dims must be >= 1");
  DMI_REQUIRE(K && RT && colors, "null argument");
  DMI_REQUIRE((long long)W * H < (1ll << 31), "image too large for the 32-bit pixel index");
  if (nPoints == 0) return DMI_OK;
  DMI_REQUIRE(xyz && mean && median && nbProjected, "null argument");
  DMI_CK(cudaSetDevice(ctx->device));
  const size_t xyzBytes = nPoints * 3 * (xyzType == DMI_F64 ? 8 : 4);
  const size_t colBytes = (size_t)nViews * W * H * 3;
  DMI_CK(ctx->c_xyz.ensure(xyzBytes));
  DMI_CK(ctx->c_colors.ensure(colBytes));
  DMI_CK(ctx->c_mean.ensure(nPoints * 3));
  DMI_CK(ctx->c_median.ensure(nPoints * 3));
  DMI_CK(ctx->c_nb.ensure(nPoints * 4));
  DMI_CK(cudaMemcpyAsync(ctx->c_xyz.p, xyz, xyzBytes, cudaMemcpyHostToDevice, ctx->stream));
  DMI_CK(cudaMemcpyAsync(ctx->c_colors.p, colors, colBytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = dmi_colorize_device(ctx, nPoints, ctx->c_xyz.p, xyzType, nViews, (const uint8_t*)ctx->c_colors.p, K, RT, W, H,
                               (uint8_t*)ctx->c_mean.p, (uint8_t*)ctx->c_median.p, (int32_t*)ctx->c_nb.p);
  if (rc != DMI_OK) return rc;
  DMI_CK(cudaMemcpyAsync(mean, ctx->c_mean.p, nPoints * 3, cudaMemcpyDeviceToHost, ctx->stream));
  DMI_CK(cudaMemcpyAsync(median, ctx->c_median.p, nPoints * 3, cudaMemcpyDeviceToHost, ctx->stream));
  DMI_CK(cudaMemcpyAsync(nbProjected, ctx->c_nb.p, nPoints * 4, cudaMemcpyDeviceToHost, ctx->stream));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  return DMI_OK;
}

int dmi_color_kernel_stats(dmi_ctx* ctx, float* ms, long long* launches)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  DMI_CK(cudaSetDevice(ctx->device));
  DMI_CK(cudaStreamSynchronize(ctx->stream));
  if (ms) *ms = ctx->color_stats.drain(); else ctx->color_stats.drain();
  if (launches) *launches = ctx->color_stats.launches;
  ctx->color_stats.launches = 0;
  return DMI_OK;
}

// ---- measurement ----------------------------------------------------------------------------------

int dmi_launch_counter(dmi_ctx* ctx, long long* launches)
{
  if (!ctx || !launches) return DMI_ERR_INVALID_ARGUMENT;
  *launches = ctx->total_launches;
  return DMI_OK;
}

int dmi_measure_fp_peak(dmi_ctx* ctx, int which, double ms_target, double* tflops)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  DMI_REQUIRE(tflops && (which == 0 || which == 1) && ms_target > 0, "bad argument");
  DMI_CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  DMI_CK(cudaGetDeviceProperties(&prop, ctx->device));
  const int blocks = prop.multiProcessorCount * 8;
  const int iters = 2048;
  const double flopsPerLaunch = (double)blocks * dmi::kPeakThreads * iters * 8.0 * dmi::kPeakChains * 2.0;
  DevBuf sink;
  DMI_CK(sink.ensure(64));
  cudaEvent_t a, b;
  DMI_CK(cudaEventCreate(&a)); DMI_CK(cudaEventCreate(&b));
  for (int w = 0; w < 3; w++) DMI_CK(dmi::launch_fp_peak(which, blocks, iters, (float*)sink.p, ctx->stream));
  DMI_CK(cudaEventRecord(a, ctx->stream));
  DMI_CK(dmi::launch_fp_peak(which, blocks, iters, (float*)sink.p, ctx->stream));
  DMI_CK(cudaEventRecord(b, ctx->stream));
  DMI_CK(cudaEventSynchronize(b));
  float one = 0.f;
  DMI_CK(cudaEventElapsedTime(&one, a, b));
  int reps = (int)std::max(1.0, std::ceil(ms_target / std::max(one, 1e-3f)));
  DMI_CK(cudaEventRecord(a, ctx->stream));
  for (int r = 0; r < reps; r++) DMI_CK(dmi::launch_fp_peak(which, blocks, iters, (float*)sink.p, ctx->stream));
  DMI_CK(cudaEventRecord(b, ctx->stream));
  DMI_CK(cudaEventSynchronize(b));
  float ms = 0.f;
  DMI_CK(cudaEventElapsedTime(&ms, a, b));
  *tflops = flopsPerLaunch * reps / (ms * 1e-3) / 1e12;
  cudaEventDestroy(a); cudaEventDestroy(b);
  sink.release();
  return DMI_OK;
}

}  // extern "C"
