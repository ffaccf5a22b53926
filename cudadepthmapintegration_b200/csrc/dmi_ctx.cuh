// Host-side state of one context and the helpers shared by the translation units of the C ABI
// (dmi_api.cu: single-GPU entry points; dmi_shard.cu: multi-GPU sharding over NCCL).
#pragma once
#include "../../include/dmi_b200.h"
#include "dmi_internal.cuh"

#include <string>
#include <vector>

struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct EventSpan { cudaEvent_t a, b; };

struct KernelStats
{
  std::vector<EventSpan> pending;
  std::vector<EventSpan> pool;
  long long launches = 0;
  float carry = 0.f;                   // time of spans recycled before anybody asked for the statistics
  EventSpan open() {
    EventSpan s;
    if (pending.size() >= 1024)        // a caller that never reads the statistics must not accumulate events
    {
      s = pending.front();
      pending.erase(pending.begin());
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) carry += ms; else cudaGetLastError();
      return s;
    }
    if (!pool.empty()) { s = pool.back(); pool.pop_back(); }
    else { cudaEventCreate(&s.a); cudaEventCreate(&s.b); }
    return s;
  }
  float drain() {
    float total = carry;
    carry = 0.f;
    for (auto& s : pending) { float ms = 0.f; if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) total += ms; pool.push_back(s); }
    pending.clear();
    return total;
  }
  void destroy() {
    for (auto& s : pending) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto& s : pool) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    pending.clear(); pool.clear();
  }
};

struct dmi_shard_state;                // dmi_shard.cu: NCCL communicator, ring of view-group buffers, comm stream
struct dmi_contour_state;              // dmi_contour.cu: point scalars, scan scratch, the last surface

struct dmi_ctx
{
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  bool initialized = false;
  dmi::GridParams g{};
  // volume slab
  DevBuf vol;
  size_t vol_bytes = 0;
  int vol_type = DMI_F64;
  bool vol_active = false;
  // host-view pipeline: two staging slots
  DevBuf stage_depth[2], stage_cost[2], filtered;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  bool slot_used[2] = {false, false};
  // coloration scratch
  DevBuf c_xyz, c_colors, c_mats, c_mean, c_median, c_nb, c_sort;
  KernelStats tsdf_stats, color_stats;
  long long opt_kernel = DMI_TSDF_KERNEL_AUTO, opt_chunk = 0;
  long long total_launches = 0;
  dmi::FastChunk fast_chunk{};
  DevBuf counters, cls, tiles, viewscratch, maskscratch;
  bool counters_on = false;
  bool opt_cull = true;
  int opt_quota = 32;
  std::string err;
  dmi_shard_state* shard = nullptr;    // set by dmi_comm_init
  dmi_contour_state* contour = nullptr;

  int fail(int code, const std::string& msg) { err = msg; return code; }
  int fail_cuda(cudaError_t e, const char* what)
  {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? DMI_ERR_OUT_OF_MEMORY : DMI_ERR_CUDA;
  }
};

#define DMI_CK(call)                                                       \
  do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ctx->fail_cuda(e__, #call); } while (0)
#define DMI_REQUIRE(cond, msg)                                             \
  do { if (!(cond)) return ctx->fail(DMI_ERR_INVALID_ARGUMENT, msg); } while (0)


// shared host helpers (dmi_api.cu)
namespace dmi_host {
size_t slab_cells(const dmi::GridParams& g);
bool fast_path_applies(const dmi_ctx* ctx);
// Fast kernel over views that are ALREADY prepared; d_depths may be null when d_lo is given
int integrate_fast_prepared(dmi_ctx* ctx, int nViews, const double* d_depths, const int* d_lo, const float* d_cls, long long clsSpare,
                            const float* d_tiles, const double* K, const double* RT);
void set_create_error(const std::string& msg);
void shard_release(dmi_ctx* ctx);      // dmi_shard.cu: frees ctx->shard (called by dmi_destroy)
void contour_release(dmi_ctx* ctx);    // dmi_contour.cu: frees ctx->contour (called by dmi_destroy)
}
