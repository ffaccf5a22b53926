// Multi-GPU part of the C ABI (include/dmi_b200.h, "sharding"): z-layers dealt round-robin to the GPUs, the prepared
// views all-gathered with NCCL group by group behind the integration, the finished layers gathered once.
//
// The reference is single-GPU; north_star asks for z-sharding over 1/2/4/8 GPUs with an NCCL exchange
// of the views and no cross-GPU reduction in the hot loop.  Every voxel has one owner, so there is none: the only
// collective is the all-gather of the views' PREPARED form (classification float + int32 residual = the lossless
// 8-byte split of the filtered double depth, + tile statistics), which each view's owner builds once.
//   SPMD entry points (one process or thread per GPU):   dmi_comm_* / dmi_shard_*
//   single process, all GPUs of the box:                 dmi_group_*   (one host thread per GPU inside each call)
// NCCL is loaded with dlopen at first use, so that libdmi_b200.so neither needs NCCL when it runs on one GPU nor
// brings a second copy into a process that already has one (PyTorch's).
#include "dmi_ctx.cuh"

#include <nccl.h>      // types and prototypes only
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- NCCL, loaded on demand ------------------------------------------------------------------------------------

struct NcclApi
{
  void* lib = nullptr;
  std::string err;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  // NCCL >= 2.28 only (optional): communicators that move all-gathers with the copy engines, symmetric windows
  decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr;
  decltype(&ncclMemAlloc) MemAlloc = nullptr;
  decltype(&ncclMemFree) MemFree = nullptr;
  decltype(&ncclCommWindowRegister) CommWindowRegister = nullptr;
  decltype(&ncclCommWindowDeregister) CommWindowDeregister = nullptr;
  bool copy_engines() const { return CommInitRankConfig && MemAlloc && MemFree && CommWindowRegister && CommWindowDeregister; }
};

NcclApi& nccl()
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // an already loaded libnccl.so.2 (PyTorch's) is found by its soname; else the system's
    for (const char* name : {"libnccl.so.2", "libnccl.so"})
      if ((api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.lib) { api.err = std::string("NCCL not available: ") + dlerror(); return; }
#define DMI_NCCL_SYM(field, sym)                                                          \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, #sym));              \
    if (!api.field && api.err.empty()) api.err = "NCCL symbol missing: " #sym;
    DMI_NCCL_SYM(GetUniqueId, ncclGetUniqueId) DMI_NCCL_SYM(CommInitRank, ncclCommInitRank) DMI_NCCL_SYM(CommInitAll, ncclCommInitAll)
    DMI_NCCL_SYM(CommDestroy, ncclCommDestroy) DMI_NCCL_SYM(AllGather, ncclAllGather) DMI_NCCL_SYM(Send, ncclSend)
    DMI_NCCL_SYM(Recv, ncclRecv) DMI_NCCL_SYM(GroupStart, ncclGroupStart) DMI_NCCL_SYM(GroupEnd, ncclGroupEnd)
    DMI_NCCL_SYM(GetErrorString, ncclGetErrorString) DMI_NCCL_SYM(GetVersion, ncclGetVersion)
#undef DMI_NCCL_SYM
#define DMI_NCCL_OPT(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, #sym));
    DMI_NCCL_OPT(CommInitRankConfig, ncclCommInitRankConfig) DMI_NCCL_OPT(MemAlloc, ncclMemAlloc) DMI_NCCL_OPT(MemFree, ncclMemFree)
    DMI_NCCL_OPT(CommWindowRegister, ncclCommWindowRegister) DMI_NCCL_OPT(CommWindowDeregister, ncclCommWindowDeregister)
#undef DMI_NCCL_OPT
  });
  return api;
}

static_assert(NCCL_UNIQUE_ID_BYTES == DMI_UNIQUE_ID_BYTES, "dmi_b200.h states the size of ncclUniqueId");

#ifndef NCCL_CTA_POLICY_ZERO
#define NCCL_CTA_POLICY_ZERO 0x02                          // NCCL 2.28: all-gathers of registered windows run on the copy engines
#endif

// The view exchange prefers the copy engines (no SM is taken from the integration kernel): that needs NCCL >= 2.28, a
// communicator created with the zero-CTA policy and group buffers registered as symmetric windows.  DMI_EXCHANGE=sm
// in the environment keeps NCCL's kernels instead.
bool want_copy_engines()
{
  const char* e = getenv("DMI_EXCHANGE");
  if (e && !strcmp(e, "sm")) return false;
  int v = 0;
  return nccl().err.empty() && nccl().copy_engines() && nccl().GetVersion(&v) == ncclSuccess && v >= 22800;
}

ncclResult_t comm_init(ncclComm_t* comm, int world, const ncclUniqueId& id, int rank, bool* ce)
{
  *ce = false;
  if (want_copy_engines())
  {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.CTAPolicy = NCCL_CTA_POLICY_ZERO;
    ncclResult_t r = nccl().CommInitRankConfig(comm, world, id, rank, &cfg);
    if (r == ncclSuccess) *ce = true;
    // a library that rejects the configuration does so before any rendezvous: the plain communicator still works
    if (r != ncclInvalidArgument) return r;
  }
  return nccl().CommInitRank(comm, world, id, rank);
}

constexpr int kRing = 3;          // view-group buffers: one being integrated, one being gathered, one spare
constexpr int kLayerPlanes = 32;  // z-layer thickness of the sharding = the fast kernel's supertile depth

}  // namespace

// A group buffer: plain device memory, or (copy-engine exchange) NCCL's allocation registered as a symmetric window.
// ensure() is collective in the second case: every rank calls it with the same size in the same order.
struct ExchangeBuf
{
  void* p = nullptr;
  size_t cap = 0;
  ncclWindow_t win = nullptr;
  bool sym = false;
  std::string err;
  bool ensure(size_t bytes, ncclComm_t comm, bool wantSym)
  {
    if (bytes <= cap) return true;
    release(comm);
    bytes = (bytes + 4095) / 4096 * 4096;
    if (wantSym)
    {
      ncclResult_t r = nccl().MemAlloc(&p, bytes);
      if (r != ncclSuccess) { p = nullptr; err = std::string("ncclMemAlloc: ") + nccl().GetErrorString(r); return false; }
      // best effort: without a window (no symmetric-memory support on this system) the buffer is still valid device
      // memory and the all-gather falls back to NCCL's kernels
      r = nccl().CommWindowRegister(comm, p, bytes, &win, NCCL_WIN_COLL_SYMMETRIC);
      if (r != ncclSuccess) win = nullptr;
      sym = true;
    }
    else if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; err = cudaGetErrorString(cudaGetLastError()); return false; }
    cap = bytes;
    return true;
  }
  void release(ncclComm_t comm)
  {
    if (p && sym) { if (win) nccl().CommWindowDeregister(comm, win); nccl().MemFree(p); }
    else if (p) cudaFree(p);
    p = nullptr; cap = 0; win = nullptr; sym = false;
  }
};

struct dmi_shard_state
{
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  bool copy_engines = false;      // the communicator was created with the zero-CTA policy
  // comm_stream carries nothing but the all-gathers, so that one group's exchange follows the other without a pause;
  // the owner-side preparation before it and the rebuilding of the tile statistics after it run on streams of their own
  cudaStream_t comm_stream = nullptr, prep_stream = nullptr, stats_stream = nullptr;
  cudaEvent_t prepared[kRing] = {}, gathered[kRing] = {};
  ExchangeBuf cls[kRing], lo[kRing];
  DevBuf tiles[kRing];
  cudaEvent_t ready[kRing] = {}, freed[kRing] = {}, entry = nullptr, staged[2] = {}, stage_free[2] = {};
  bool used[kRing] = {false, false, false}, stage_used[2] = {false, false};
  size_t spare_set_for = 0;       // capacity (floats) for which the spare -1.0f slots were written
  DevBuf colors;
};

#define DMI_NCCL(call)                                                                            \
  do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) {                                       \
         ctx->err = std::string(#call) + ": " + nccl().GetErrorString(r__); return DMI_ERR_CUDA; } } while (0)

void dmi_host::shard_release(dmi_ctx* ctx)
{
  dmi_shard_state* s = ctx->shard;
  if (!s) return;
  cudaSetDevice(ctx->device);
  for (cudaStream_t st : {s->prep_stream, s->comm_stream, s->stats_stream}) if (st) cudaStreamSynchronize(st);
  for (int b = 0; b < kRing; b++)
  {
    s->cls[b].release(s->comm); s->lo[b].release(s->comm); s->tiles[b].release();
    if (s->ready[b]) cudaEventDestroy(s->ready[b]);
    if (s->freed[b]) cudaEventDestroy(s->freed[b]);
    if (s->prepared[b]) cudaEventDestroy(s->prepared[b]);
    if (s->gathered[b]) cudaEventDestroy(s->gathered[b]);
  }
  for (int b = 0; b < 2; b++)
  {
    if (s->staged[b]) cudaEventDestroy(s->staged[b]);
    if (s->stage_free[b]) cudaEventDestroy(s->stage_free[b]);
  }
  if (s->entry) cudaEventDestroy(s->entry);
  s->colors.release();
  if (s->comm && nccl().CommDestroy) nccl().CommDestroy(s->comm);
  for (cudaStream_t st : {s->prep_stream, s->comm_stream, s->stats_stream}) if (st) cudaStreamDestroy(st);
  delete s;
  ctx->shard = nullptr;
}

namespace {

// ---- who owns which view ------------------------------------------------------------------------------------------
// Views go in groups of consecutive views; inside a group of n views, rank r owns the contiguous share
// [g0 + r * pg, g0 + (r + 1) * pg) clipped to the group, pg = ceil(n / world) -- so that one in-place all-gather per array
// assembles the group in list order.  A full group holds per * world views (per = 128 / world, at least 1); the first
// groups are shorter (per/4, per/4, per/2 views per rank) so that the integration starts after a short first exchange,
// and so are the last ones, so that little is left to integrate once the last views have arrived.
struct ShardPlan
{
  int V, world, per;
  std::vector<int> start;                                 // start[g] .. start[g + 1]: views of group g
  ShardPlan(int nViews, int w) : V(nViews), world(w)
  {
    per = std::max(1, 128 / w);
    const int G = per * w;
    std::vector<int> sizes;
    int left = V;
    auto take = [&](int n) { n = std::min(n, left); if (n > 0) { sizes.push_back(n); left -= n; } };
    if (w > 1 && V >= 3 * G)
    {
      const int q = std::max(1, per / 4) * w, h = std::max(1, per / 2) * w;
      take(q); take(q); take(h);
      while (left > 2 * q + h + G) take(G);
      // what is left: up to one more full group's worth, then the short tail
      const int tail = 2 * q + h;
      if (left > tail) take(left - tail);
      take(h); take(q); take(q);
    }
    while (left > 0) take(G);
    start.push_back(0);
    for (int n : sizes) start.push_back(start.back() + n);
  }
  int nGroups() const { return (int)start.size() - 1; }
  void group(int g, int& g0, int& g1, int& pg) const
  {
    g0 = start[g]; g1 = start[g + 1];
    pg = (g1 - g0 + world - 1) / world;
  }
  void share(int g, int rank, int& a, int& b) const      // rank's views of group g: [a, b)
  {
    int g0, g1, pg;
    group(g, g0, g1, pg);
    a = std::min(g1, g0 + rank * pg);
    b = std::min(g1, a + pg);
  }
  int count(int rank) const { int n = 0; for (int g = 0; g < nGroups(); g++) { int a, b; share(g, rank, a, b); n += b - a; } return n; }
  int capacity() const { return (per + 1) * world; }    // views a group buffer must hold (padding of a ragged group included)
};

int shard_ensure(dmi_ctx* ctx, const ShardPlan& plan, bool wantLo)
{
  dmi_shard_state* s = ctx->shard;
  const size_t npix = (size_t)ctx->g.W * ctx->g.H;
  const size_t tilesPerView = (size_t)dmi::tile_pyramid_layout(ctx->g.W, ctx->g.H).perView;
  const size_t cap = (size_t)plan.capacity();
  const size_t clsFloats = cap * npix + 8;               // + the spare -1.0f slot behind the views
  for (int b = 0; b < kRing; b++)
  {
    if (!s->cls[b].ensure(clsFloats * 4, s->comm, s->copy_engines)) return ctx->fail(DMI_ERR_OUT_OF_MEMORY, s->cls[b].err);
    if (wantLo && !s->lo[b].ensure(cap * npix * 4, s->comm, s->copy_engines)) return ctx->fail(DMI_ERR_OUT_OF_MEMORY, s->lo[b].err);
    DMI_CK(s->tiles[b].ensure(cap * tilesPerView * 4));
    if (!s->ready[b])
    {
      DMI_CK(cudaEventCreateWithFlags(&s->ready[b], cudaEventDisableTiming));
      DMI_CK(cudaEventCreateWithFlags(&s->freed[b], cudaEventDisableTiming));
      DMI_CK(cudaEventCreateWithFlags(&s->prepared[b], cudaEventDisableTiming));
      DMI_CK(cudaEventCreateWithFlags(&s->gathered[b], cudaEventDisableTiming));
    }
  }
  for (int b = 0; b < 2; b++)
    if (!s->staged[b])
    {
      DMI_CK(cudaEventCreateWithFlags(&s->staged[b], cudaEventDisableTiming));
      DMI_CK(cudaEventCreateWithFlags(&s->stage_free[b], cudaEventDisableTiming));
    }
  if (!s->entry) DMI_CK(cudaEventCreateWithFlags(&s->entry, cudaEventDisableTiming));
  if (s->spare_set_for != cap * npix)
  {
    const float minus1 = -1.0f;
    for (int b = 0; b < kRing; b++)
      DMI_CK(cudaMemcpyAsync((float*)s->cls[b].p + cap * npix, &minus1, 4, cudaMemcpyHostToDevice, s->prep_stream));
    DMI_CK(cudaStreamSynchronize(s->prep_stream));       // `minus1` lives on this stack frame
    s->spare_set_for = cap * npix;
  }
  return DMI_OK;
}

// One sharded integration: `mine` = this rank's views in the order of dmi_shard_view_indices; device pointers, or host
// pointers when `fromHost` (then each group's share is uploaded on the copy stream while earlier groups are gathered
// and integrated).
int shard_integrate(dmi_ctx* ctx, int nViews, const double* depths, const double* cost, double thr, const double* K,
                    const double* RT, bool fromHost)
{
  dmi_shard_state* s = ctx->shard;
  const dmi::GridParams& g = ctx->g;
  const size_t npix = (size_t)g.W * g.H;
  const size_t tilesPerView = (size_t)dmi::tile_pyramid_layout(g.W, g.H).perView;
  const ShardPlan plan(nViews, s->world);
  int rc = shard_ensure(ctx, plan, true);
  if (rc != DMI_OK) return rc;
  const size_t spare = (size_t)plan.capacity() * npix;
  const bool haveWork = dmi_host::slab_cells(g) != 0;

  // the ring buffers may still be read by integration launches of an earlier call on the context's stream
  DMI_CK(cudaEventRecord(s->entry, ctx->stream));
  DMI_CK(cudaStreamWaitEvent(s->prep_stream, s->entry, 0));

  EventSpan span = ctx->tsdf_stats.open();
  bool spanOpen = false;
  size_t done = 0;                                        // views of `mine` consumed so far
  for (int gi = 0; gi < plan.nGroups(); gi++)
  {
    int g0, g1, pg, a, b;
    plan.group(gi, g0, g1, pg);
    plan.share(gi, s->rank, a, b);
    const int slot = gi % kRing, mineN = b - a;
    float* cls = (float*)s->cls[slot].p;
    int* lo = (int*)s->lo[slot].p;
    float* tiles = (float*)s->tiles[slot].p;
    if (s->used[slot]) DMI_CK(cudaStreamWaitEvent(s->prep_stream, s->freed[slot], 0));
    if (mineN > 0)
    {
      const double* d = depths + npix * done;
      const double* c = cost ? cost + npix * done : nullptr;
      if (fromHost)
      {
        const int st = gi & 1;
        const size_t bytes = (size_t)mineN * npix * 8;
        DMI_CK(ctx->stage_depth[st].ensure((size_t)(plan.per + 1) * npix * 8));
        if (cost) DMI_CK(ctx->stage_cost[st].ensure((size_t)(plan.per + 1) * npix * 8));
        if (s->stage_used[st]) DMI_CK(cudaStreamWaitEvent(ctx->copy_stream, s->stage_free[st], 0));
        DMI_CK(cudaMemcpyAsync(ctx->stage_depth[st].p, d, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (cost) DMI_CK(cudaMemcpyAsync(ctx->stage_cost[st].p, c, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        DMI_CK(cudaEventRecord(s->staged[st], ctx->copy_stream));
        DMI_CK(cudaStreamWaitEvent(s->prep_stream, s->staged[st], 0));
        d = (const double*)ctx->stage_depth[st].p;
        c = cost ? (const double*)ctx->stage_cost[st].p : nullptr;
      }
      // the owner prepares its views ONCE, straight into its segment of the group buffer
      const size_t seg = (size_t)(a - g0);
      // (with several ranks the tile statistics are not exchanged: every rank rebuilds them from the classification images)
      DMI_CK(dmi::launch_prepare_views(d, c, thr, mineN, g.W, g.H, cls + seg * npix, lo + seg * npix, -1,
                                       tiles + seg * tilesPerView, s->prep_stream, s->world == 1));
      ctx->total_launches += s->world == 1 ? 1 + dmi::tile_pyramid_layout(g.W, g.H).nLevels : 1;
      if (fromHost)
      {
        DMI_CK(cudaEventRecord(s->stage_free[gi & 1], s->prep_stream));
        s->stage_used[gi & 1] = true;
      }
      done += (size_t)mineN;
    }
    DMI_CK(cudaEventRecord(s->prepared[slot], s->prep_stream));
    cudaEvent_t ready = s->prepared[slot];
    if (s->world > 1)
    {
      DMI_CK(cudaStreamWaitEvent(s->comm_stream, s->prepared[slot], 0));
      // in place: rank r's segment starts r * pg views into each array (the last group is padded to pg * world views)
      const size_t r = (size_t)s->rank * pg;
      DMI_NCCL(nccl().GroupStart());
      DMI_NCCL(nccl().AllGather(cls + r * npix, cls, (size_t)pg * npix, ncclFloat, s->comm, s->comm_stream));
      DMI_NCCL(nccl().AllGather(lo + r * npix, lo, (size_t)pg * npix, ncclInt32, s->comm, s->comm_stream));
      DMI_NCCL(nccl().GroupEnd());
      // 8 bytes per pixel travelled; the tile statistics (19 % more) are rebuilt here from the classification images
      DMI_CK(cudaEventRecord(s->gathered[slot], s->comm_stream));
      DMI_CK(cudaStreamWaitEvent(s->stats_stream, s->gathered[slot], 0));
      DMI_CK(dmi::launch_tile_stats_from_cls(cls, g1 - g0, g.W, g.H, tiles, s->stats_stream));
      ctx->total_launches += 1 + dmi::tile_pyramid_layout(g.W, g.H).nLevels;
      DMI_CK(cudaEventRecord(s->ready[slot], s->stats_stream));
      ready = s->ready[slot];
    }
    DMI_CK(cudaStreamWaitEvent(ctx->stream, ready, 0));
    if (haveWork)
    {
      if (!spanOpen) { DMI_CK(cudaEventRecord(span.a, ctx->stream)); spanOpen = true; }
      rc = dmi_host::integrate_fast_prepared(ctx, g1 - g0, nullptr, lo, cls, (long long)spare, tiles, K + 16 * (size_t)g0,
                                             RT + 16 * (size_t)g0);
      if (rc != DMI_OK) return rc;
    }
    DMI_CK(cudaEventRecord(s->freed[slot], ctx->stream));
    s->used[slot] = true;
  }
  if (spanOpen)
  {
    DMI_CK(cudaEventRecord(span.b, ctx->stream));
    ctx->tsdf_stats.pending.push_back(span);
  }
  else ctx->tsdf_stats.pool.push_back(span);
  return DMI_OK;
}

int shard_check(dmi_ctx* ctx, int nViews, const void* depths, const double* K, const double* RT)
{
  if (!ctx->shard) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_comm_init has not been called");
  if (!ctx->initialized || !ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "no depthMap or KRTD matrix have been loaded");
  DMI_REQUIRE(K && RT, "null argument");
  DMI_REQUIRE(depths || ShardPlan(nViews, ctx->shard->world).count(ctx->shard->rank) == 0, "null argument");
  if (!dmi_host::fast_path_applies(ctx))
    return ctx->fail(DMI_ERR_BAD_PARAMETERS, "sharded integration needs the certified fast path (0 < Thick, 0 <= Delta, finite parameters, kernel AUTO)");
  return DMI_OK;
}

}  // namespace

extern "C" {

int dmi_comm_unique_id(unsigned char id[DMI_UNIQUE_ID_BYTES])
{
  if (!id) return DMI_ERR_INVALID_ARGUMENT;
  if (!nccl().err.empty()) { dmi_host::set_create_error(nccl().err); return DMI_ERR_CUDA; }
  ncclUniqueId u;
  ncclResult_t r = nccl().GetUniqueId(&u);
  if (r != ncclSuccess) { dmi_host::set_create_error(std::string("ncclGetUniqueId: ") + nccl().GetErrorString(r)); return DMI_ERR_CUDA; }
  memcpy(id, &u, DMI_UNIQUE_ID_BYTES);
  return DMI_OK;
}

static int attach_comm(dmi_ctx* ctx, ncclComm_t comm, int rank, int world, bool copyEngines)
{
  dmi_host::shard_release(ctx);
  dmi_shard_state* s = new dmi_shard_state();
  s->comm = comm; s->rank = rank; s->world = world; s->copy_engines = copyEngines;
  ctx->shard = s;
  // CTAs of the persistent integration kernel retire sooner, so that the exchange's kernels find a free SM slot quickly
  if (world > 1 && ctx->opt_quota == 32) ctx->opt_quota = 8;
  int lo = 0, hi = 0;
  DMI_CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  // high priority: the exchange's few CTAs must not queue behind the integration kernel's
  DMI_CK(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, hi));
  DMI_CK(cudaStreamCreateWithPriority(&s->prep_stream, cudaStreamNonBlocking, hi));
  DMI_CK(cudaStreamCreateWithPriority(&s->stats_stream, cudaStreamNonBlocking, hi));
  return DMI_OK;
}

int dmi_comm_init(dmi_ctx* ctx, const unsigned char id[DMI_UNIQUE_ID_BYTES], int rank, int world)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  DMI_REQUIRE(world >= 1 && rank >= 0 && rank < world, "need 0 <= rank < world");
  DMI_REQUIRE(world == 1 || id, "null unique id");
  DMI_CK(cudaSetDevice(ctx->device));
  ncclComm_t comm = nullptr;
  bool ce = false;
  if (world > 1)
  {
    if (!nccl().err.empty()) return ctx->fail(DMI_ERR_CUDA, nccl().err);
    ncclUniqueId u;
    memcpy(&u, id, DMI_UNIQUE_ID_BYTES);
    DMI_NCCL(comm_init(&comm, world, u, rank, &ce));
  }
  return attach_comm(ctx, comm, rank, world, ce);
}

int dmi_comm_destroy(dmi_ctx* ctx)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  dmi_host::shard_release(ctx);
  return DMI_OK;
}

int dmi_comm_info(dmi_ctx* ctx, int* rank, int* world, int* ncclVersion)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->shard) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_comm_init has not been called");
  if (rank) *rank = ctx->shard->rank;
  if (world) *world = ctx->shard->world;
  if (ncclVersion) { *ncclVersion = 0; if (ctx->shard->world > 1 && nccl().GetVersion) nccl().GetVersion(ncclVersion); }
  return DMI_OK;
}

int dmi_comm_copy_engines(dmi_ctx* ctx)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->shard) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_comm_init has not been called");
  const dmi_shard_state* s = ctx->shard;
  // (once the group buffers exist: only if their windows could be registered)
  return s->copy_engines && !(s->cls[0].p && !s->cls[0].win) ? 1 : 0;
}

int dmi_shard_initialize(dmi_ctx* ctx, const double gridMatrix[16], const int gridDims[3], const double gridOrig[3],
                         const double gridSpacing[3], double thick, double rho, double eta, double delta,
                         const int depthMapDims[2])
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->shard) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_comm_init has not been called");
  int rc = dmi_initialize(ctx, gridMatrix, gridDims, gridOrig, gridSpacing, thick, rho, eta, delta, depthMapDims);
  if (rc != DMI_OK) return rc;
  return dmi_set_slab_layers(ctx, kLayerPlanes, ctx->shard->rank, ctx->shard->world);
}

int dmi_shard_view_count(int nViews, int world, int rank, int* count)
{
  if (!count || nViews < 0 || world < 1 || rank < 0 || rank >= world) return DMI_ERR_INVALID_ARGUMENT;
  *count = ShardPlan(nViews, world).count(rank);
  return DMI_OK;
}

int dmi_shard_view_indices(int nViews, int world, int rank, int* indices)
{
  if (!indices || nViews < 0 || world < 1 || rank < 0 || rank >= world) return DMI_ERR_INVALID_ARGUMENT;
  const ShardPlan plan(nViews, world);
  int n = 0;
  for (int g = 0; g < plan.nGroups(); g++)
  {
    int a, b;
    plan.share(g, rank, a, b);
    for (int v = a; v < b; v++) indices[n++] = v;
  }
  return DMI_OK;
}

int dmi_shard_group_count(int nViews, int world, int* count)
{
  if (!count || nViews < 0 || world < 1) return DMI_ERR_INVALID_ARGUMENT;
  *count = ShardPlan(nViews, world).nGroups();
  return DMI_OK;
}

int dmi_shard_group_starts(int nViews, int world, int* starts)
{
  if (!starts || nViews < 0 || world < 1) return DMI_ERR_INVALID_ARGUMENT;
  const ShardPlan plan(nViews, world);
  for (size_t g = 0; g < plan.start.size(); g++) starts[g] = plan.start[g];
  return DMI_OK;
}

int dmi_shard_integrate_device(dmi_ctx* ctx, int nViews, const double* d_myDepths, const double* d_myBestCost,
                               double thresholdBestCost, const double* K, const double* RT)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  int rc = shard_check(ctx, nViews, d_myDepths, K, RT);
  if (rc != DMI_OK) return rc;
  DMI_CK(cudaSetDevice(ctx->device));
  return shard_integrate(ctx, nViews, d_myDepths, d_myBestCost, thresholdBestCost, K, RT, false);
}

int dmi_shard_integrate_host(dmi_ctx* ctx, int nViews, const double* myDepths, const double* myBestCost,
                             double thresholdBestCost, const double* K, const double* RT)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  int rc = shard_check(ctx, nViews, myDepths, K, RT);
  if (rc != DMI_OK) return rc;
  DMI_CK(cudaSetDevice(ctx->device));
  rc = shard_integrate(ctx, nViews, myDepths, myBestCost, thresholdBestCost, K, RT, true);
  if (rc != DMI_OK) return rc;
  DMI_CK(cudaStreamSynchronize(ctx->stream));            // host pointers: synchronous at the ABI
  return DMI_OK;
}

int dmi_shard_gather_volume_device(dmi_ctx* ctx, int root, void* d_full)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->shard) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_comm_init has not been called");
  if (!ctx->initialized || !ctx->vol_active) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_volume_begin has not been called");
  dmi_shard_state* s = ctx->shard;
  const dmi::GridParams& g = ctx->g;
  DMI_REQUIRE(root >= 0 && root < s->world, "root out of range");
  DMI_REQUIRE(s->rank != root || d_full, "null argument");
  DMI_REQUIRE(g.layL > 0 && g.layStride == s->world && g.layPhase == s->rank, "the slab was not set by dmi_shard_initialize");
  DMI_CK(cudaSetDevice(ctx->device));
  const size_t esz = ctx->vol_type == DMI_F64 ? 8 : 4;
  const ncclDataType_t dt = ctx->vol_type == DMI_F64 ? ncclFloat64 : ncclFloat32;
  const size_t plane = (size_t)g.Nx * g.Ny;
  const int L = g.layL;
  if (s->world > 1) DMI_NCCL(nccl().GroupStart());
  for (int r = 0; r < s->world; r++)
    for (int q = 0;; q++)
    {
      const long long start = ((long long)q * s->world + r) * L;
      if (start >= g.Nz) break;
      const size_t count = (size_t)std::min<long long>(L, g.Nz - start) * plane;
      char* dst = (char*)d_full + (size_t)start * plane * esz;
      const char* mine = (const char*)ctx->vol.p + (size_t)q * L * plane * esz;
      if (r == s->rank && r == root) DMI_CK(cudaMemcpyAsync(dst, mine, count * esz, cudaMemcpyDeviceToDevice, ctx->stream));
      else if (s->rank == root) DMI_NCCL(nccl().Recv(dst, count, dt, r, s->comm, ctx->stream));
      else if (r == s->rank) DMI_NCCL(nccl().Send(mine, count, dt, root, s->comm, ctx->stream));
    }
  if (s->world > 1) DMI_NCCL(nccl().GroupEnd());
  return DMI_OK;
}

// ---- coloration: points sharded by contiguous index range, colour images all-gathered -------------------------------

int dmi_shard_range(size_t n, int world, int rank, size_t* first, size_t* count)
{
  if (!first || !count || world < 1 || rank < 0 || rank >= world) return DMI_ERR_INVALID_ARGUMENT;
  const size_t per = (n + (size_t)world - 1) / (size_t)world;
  *first = std::min(n, per * (size_t)rank);
  *count = std::min(n, *first + per) - *first;
  return DMI_OK;
}

int dmi_shard_colorize_device(dmi_ctx* ctx, size_t nMyPoints, const void* d_myXyz, int xyzType, int nViews,
                              const uint8_t* d_myColors, const double* K, const double* RT, int W, int H,
                              uint8_t* d_mean, uint8_t* d_median, int32_t* d_nbProjected)
{
  if (!ctx) return DMI_ERR_INVALID_ARGUMENT;
  if (!ctx->shard) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_comm_init has not been called");
  if (nViews <= 0) return ctx->fail(DMI_ERR_NO_VIEWS, "Error when input has been set or during reading vti/krtd file path");
  DMI_REQUIRE(W >= 1 && H >= 1 && (long long)W * H < (1ll << 31), "bad image dims");
  dmi_shard_state* s = ctx->shard;
  DMI_CK(cudaSetDevice(ctx->device));
  const size_t img = (size_t)W * H * 3;
  size_t first = 0, mine = 0;
  dmi_shard_range((size_t)nViews, s->world, s->rank, &first, &mine);
  const size_t per = ((size_t)nViews + s->world - 1) / s->world;
  DMI_REQUIRE(mine == 0 || d_myColors, "null argument");
  const uint8_t* all = d_myColors;
  if (s->world > 1)
  {
    DMI_CK(s->colors.ensure(per * s->world * img));
    uint8_t* buf = (uint8_t*)s->colors.p;
    // every rank needs every view: each contributes the block it loaded, one all-gather assembles them in list order
    if (mine) DMI_CK(cudaMemcpyAsync(buf + per * s->rank * img, d_myColors, mine * img, cudaMemcpyDeviceToDevice, ctx->stream));
    DMI_NCCL(nccl().AllGather(buf + per * s->rank * img, buf, per * img, ncclUint8, s->comm, ctx->stream));
    all = buf;
  }
  return dmi_colorize_device(ctx, nMyPoints, d_myXyz, xyzType, nViews, all, K, RT, W, H, d_mean, d_median, d_nbProjected);
}

}  // extern "C"

// ---- single process, several GPUs -----------------------------------------------------------------------------------

struct dmi_group
{
  std::vector<dmi_ctx*> ctx;
  std::string err;
  int fail(int code, const std::string& m) { err = m; return code; }
};

namespace {

// runs fn(rank) on one host thread per GPU; returns the first non-zero status (rank order) and its message
template <typename F>
int group_parallel(dmi_group* grp, F fn)
{
  const int n = (int)grp->ctx.size();
  std::vector<int> rc(n, DMI_OK);
  std::vector<std::thread> th;
  for (int r = 1; r < n; r++)
  {
    try { th.emplace_back([&, r] { rc[r] = fn(r); }); }
    catch (...) { rc[r] = DMI_ERR_OUT_OF_MEMORY; grp->ctx[r]->err = "could not start a host thread"; }
  }
  rc[0] = fn(0);
  for (auto& t : th) t.join();
  for (int r = 0; r < n; r++)
    if (rc[r] != DMI_OK) { grp->err = "GPU " + std::to_string(grp->ctx[r]->device) + ": " + grp->ctx[r]->err; return rc[r]; }
  return DMI_OK;
}

}  // namespace

extern "C" {

int dmi_group_create(const int* devices, int nDevices, dmi_group** out)
{
  if (!out) return DMI_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (!devices || nDevices < 1) { dmi_host::set_create_error("need at least one device"); return DMI_ERR_INVALID_ARGUMENT; }
  for (int a = 0; a < nDevices; a++)
    for (int b = 0; b < a; b++)
      if (devices[a] == devices[b]) { dmi_host::set_create_error("a device is listed twice"); return DMI_ERR_INVALID_ARGUMENT; }
  dmi_group* grp = new dmi_group();
  for (int r = 0; r < nDevices; r++)
  {
    dmi_ctx* c = nullptr;
    int rc = dmi_create(devices[r], &c);
    if (rc != DMI_OK) { for (auto* x : grp->ctx) dmi_destroy(x); delete grp; return rc; }
    grp->ctx.push_back(c);
  }
  std::vector<ncclComm_t> comms((size_t)nDevices, nullptr);
  bool ce = false;
  if (nDevices > 1)
  {
    ncclResult_t r = ncclSystemError;
    if (nccl().err.empty() && want_copy_engines())
    {
      // what ncclCommInitAll does, with a configuration: one unique id, every rank initialised inside one group call
      ncclUniqueId u;
      r = nccl().GetUniqueId(&u);
      if (r == ncclSuccess) r = nccl().GroupStart();
      for (int q = 0; q < nDevices && r == ncclSuccess; q++)
      {
        bool one = false;
        cudaSetDevice(devices[q]);
        r = comm_init(&comms[q], nDevices, u, q, &one);
      }
      if (r == ncclSuccess) { r = nccl().GroupEnd(); ce = r == ncclSuccess; } else nccl().GroupEnd();
    }
    else if (nccl().err.empty()) r = nccl().CommInitAll(comms.data(), nDevices, devices);
    if (r != ncclSuccess)
    {
      dmi_host::set_create_error(nccl().err.empty() ? std::string("ncclComm init: ") + nccl().GetErrorString(r) : nccl().err);
      for (auto* x : grp->ctx) dmi_destroy(x);
      delete grp;
      return DMI_ERR_CUDA;
    }
  }
  for (int r = 0; r < nDevices; r++)
  {
    cudaSetDevice(devices[r]);
    int rc = attach_comm(grp->ctx[r], comms[r], r, nDevices, ce);
    if (rc != DMI_OK)
    {
      dmi_host::set_create_error(grp->ctx[r]->err);
      for (auto* x : grp->ctx) dmi_destroy(x);
      delete grp;
      return rc;
    }
  }
  *out = grp;
  return DMI_OK;
}

int dmi_group_destroy(dmi_group* grp)
{
  if (!grp) return DMI_OK;
  for (auto* c : grp->ctx) dmi_destroy(c);
  delete grp;
  return DMI_OK;
}

const char* dmi_group_last_error(const dmi_group* grp) { return grp ? grp->err.c_str() : dmi_last_error(nullptr); }

int dmi_group_size(const dmi_group* grp) { return grp ? (int)grp->ctx.size() : 0; }

dmi_ctx* dmi_group_context(dmi_group* grp, int rank)
{
  return (grp && rank >= 0 && rank < (int)grp->ctx.size()) ? grp->ctx[rank] : nullptr;
}

int dmi_group_set_option(dmi_group* grp, int option, long long value)
{
  if (!grp) return DMI_ERR_INVALID_ARGUMENT;
  for (auto* c : grp->ctx)
  {
    int rc = dmi_set_option(c, option, value);
    if (rc != DMI_OK) { grp->err = c->err; return rc; }
  }
  return DMI_OK;
}

int dmi_group_initialize(dmi_group* grp, const double gridMatrix[16], const int gridDims[3], const double gridOrig[3],
                         const double gridSpacing[3], double thick, double rho, double eta, double delta,
                         const int depthMapDims[2])
{
  if (!grp) return DMI_ERR_INVALID_ARGUMENT;
  for (auto* c : grp->ctx)
  {
    int rc = dmi_shard_initialize(c, gridMatrix, gridDims, gridOrig, gridSpacing, thick, rho, eta, delta, depthMapDims);
    if (rc != DMI_OK) { grp->err = c->err; return rc; }
  }
  return DMI_OK;
}

// ProcessDepthMap<T> (CudaReconstruction.cu:302-386) over all GPUs of the group: host pointers, ALL views, io_scalar =
// the whole grid; accumulates onto io_scalar like the reference.  GPU r uploads only the views it owns and its own
// z-layers of io_scalar, and writes its finished layers straight into their places in io_scalar.
int dmi_group_process_depth_maps(dmi_group* grp, int nViews, const double* depths, const double* bestCost,
                                 double thresholdBestCost, const double* K, const double* RT, void* io_scalar, int scalarType)
{
  if (!grp) return DMI_ERR_INVALID_ARGUMENT;
  if (nViews <= 0) return grp->fail(DMI_ERR_NO_VIEWS, "no depthMap or KRTD matrix have been loaded");
  if (!depths || !K || !RT || !io_scalar) return grp->fail(DMI_ERR_INVALID_ARGUMENT, "null argument");
  if (scalarType != DMI_F32 && scalarType != DMI_F64) return grp->fail(DMI_ERR_INVALID_ARGUMENT, "scalarType must be DMI_F32 or DMI_F64");
  const int world = (int)grp->ctx.size();
  const size_t esz = scalarType == DMI_F64 ? 8 : 4;
  return group_parallel(grp, [&](int r) -> int {
    dmi_ctx* ctx = grp->ctx[r];
    if (!ctx->initialized) return ctx->fail(DMI_ERR_NOT_INITIALIZED, "dmi_group_initialize has not been called");
    DMI_CK(cudaSetDevice(ctx->device));
    const dmi::GridParams& g = ctx->g;
    const size_t npix = (size_t)g.W * g.H, plane = (size_t)g.Nx * g.Ny;
    int rc = dmi_volume_begin(ctx, nullptr, scalarType);
    if (rc != DMI_OK) return rc;
    // this GPU's layers of io_scalar (the call accumulates onto them, :323-327)
    for (int q = 0;; q++)
    {
      const long long start = ((long long)q * world + r) * g.layL;
      if (start >= g.Nz) break;
      const size_t bytes = (size_t)std::min<long long>(g.layL, g.Nz - start) * plane * esz;
      DMI_CK(cudaMemcpyAsync((char*)ctx->vol.p + (size_t)q * g.layL * plane * esz, (const char*)io_scalar + (size_t)start * plane * esz,
                             bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    // this GPU's views, gathered from the caller's arrays in the order the exchange expects
    const ShardPlan plan(nViews, world);
    const int mine = plan.count(r);
    std::vector<const double*> dsrc, csrc;
    // contiguous runs of the caller's arrays: one per group
    DevBuf dd, dc;
    DMI_CK(dd.ensure(std::max<size_t>(8, (size_t)mine * npix * 8)));
    if (bestCost) DMI_CK(dc.ensure(std::max<size_t>(8, (size_t)mine * npix * 8)));
    size_t off = 0;
    for (int gi = 0; gi < plan.nGroups(); gi++)
    {
      int a, b;
      plan.share(gi, r, a, b);
      if (b <= a) continue;
      const size_t bytes = (size_t)(b - a) * npix * 8;
      DMI_CK(cudaMemcpyAsync((char*)dd.p + off, depths + npix * a, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
      if (bestCost) DMI_CK(cudaMemcpyAsync((char*)dc.p + off, bestCost + npix * a, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
      off += bytes;
    }
    DMI_CK(cudaStreamSynchronize(ctx->copy_stream));
    rc = dmi_shard_integrate_device(ctx, nViews, (const double*)dd.p, bestCost ? (const double*)dc.p : nullptr, thresholdBestCost, K, RT);
    if (rc == DMI_OK)
    {
      for (int q = 0;; q++)
      {
        const long long start = ((long long)q * world + r) * g.layL;
        if (start >= g.Nz) break;
        const size_t bytes = (size_t)std::min<long long>(g.layL, g.Nz - start) * plane * esz;
        DMI_CK(cudaMemcpyAsync((char*)io_scalar + (size_t)start * plane * esz, (const char*)ctx->vol.p + (size_t)q * g.layL * plane * esz,
                               bytes, cudaMemcpyDeviceToHost, ctx->stream));
      }
      DMI_CK(cudaStreamSynchronize(ctx->stream));
    }
    else cudaStreamSynchronize(ctx->stream);
    dd.release(); dc.release();
    return rc;
  });
}

// MeshColoration::ProcessColoration (MeshColoration.cxx:98-199) over all GPUs: points sharded by contiguous index range
// (every point is independent, :140-192), each GPU uploads one block of the colour images and NCCL all-gathers them.
int dmi_group_colorize(dmi_group* grp, size_t nPoints, const void* xyz, int xyzType, int nViews, const uint8_t* colors,
                       const double* K, const double* RT, int W, int H, uint8_t* mean, uint8_t* median, int32_t* nbProjected)
{
  if (!grp) return DMI_ERR_INVALID_ARGUMENT;
  if (nViews <= 0) return grp->fail(DMI_ERR_NO_VIEWS, "Error when input has been set or during reading vti/krtd file path");
  if (xyzType != DMI_F32 && xyzType != DMI_F64) return grp->fail(DMI_ERR_INVALID_ARGUMENT, "xyzType must be DMI_F32 or DMI_F64");
  if (!colors || !K || !RT || W < 1 || H < 1) return grp->fail(DMI_ERR_INVALID_ARGUMENT, "bad argument");
  if (nPoints == 0) return DMI_OK;
  if (!xyz || !mean || !median || !nbProjected) return grp->fail(DMI_ERR_INVALID_ARGUMENT, "null argument");
  const int world = (int)grp->ctx.size();
  const size_t psz = 3 * (size_t)(xyzType == DMI_F64 ? 8 : 4), img = (size_t)W * H * 3;
  return group_parallel(grp, [&](int r) -> int {
    dmi_ctx* ctx = grp->ctx[r];
    DMI_CK(cudaSetDevice(ctx->device));
    size_t p0 = 0, np = 0, v0 = 0, nv = 0;
    dmi_shard_range(nPoints, world, r, &p0, &np);
    dmi_shard_range((size_t)nViews, world, r, &v0, &nv);
    DMI_CK(ctx->c_xyz.ensure(std::max<size_t>(8, np * psz)));
    DMI_CK(ctx->c_colors.ensure(std::max<size_t>(8, nv * img)));
    DMI_CK(ctx->c_mean.ensure(std::max<size_t>(8, np * 3)));
    DMI_CK(ctx->c_median.ensure(std::max<size_t>(8, np * 3)));
    DMI_CK(ctx->c_nb.ensure(std::max<size_t>(8, np * 4)));
    if (np) DMI_CK(cudaMemcpyAsync(ctx->c_xyz.p, (const char*)xyz + p0 * psz, np * psz, cudaMemcpyHostToDevice, ctx->stream));
    if (nv) DMI_CK(cudaMemcpyAsync(ctx->c_colors.p, colors + v0 * img, nv * img, cudaMemcpyHostToDevice, ctx->stream));
    // every rank takes part in the all-gather, also one without points
    int rc = dmi_shard_colorize_device(ctx, np, ctx->c_xyz.p, xyzType, nViews, (const uint8_t*)ctx->c_colors.p, K, RT, W, H,
                                       (uint8_t*)ctx->c_mean.p, (uint8_t*)ctx->c_median.p, (int32_t*)ctx->c_nb.p);
    if (rc == DMI_OK && np)
    {
      DMI_CK(cudaMemcpyAsync(mean + p0 * 3, ctx->c_mean.p, np * 3, cudaMemcpyDeviceToHost, ctx->stream));
      DMI_CK(cudaMemcpyAsync(median + p0 * 3, ctx->c_median.p, np * 3, cudaMemcpyDeviceToHost, ctx->stream));
      DMI_CK(cudaMemcpyAsync(nbProjected + p0, ctx->c_nb.p, np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == DMI_OK && e != cudaSuccess) return ctx->fail_cuda(e, "cudaStreamSynchronize");
    return rc;
  });
}

}  // extern "C"
