/*
 * dmi_b200.h -- C ABI of the B200-native depth-map integration engine.
 *
 * This is the drop-in boundary for the one hot path of bastienjacquet/CudaDepthMapIntegration:
 * the two free functions the reference's filter forward-declares and calls
 * (Reconstruction/vtkCudaReconstructionFilter.cxx:65-71, called at :171-176),
 *
 *     void CudaInitialize(vtkMatrix4x4*, int gridDims[3], double gridOrig[3], double gridSpacing[3],
 *                         double thick, double rho, double eta, double delta, int depthMapDims[2]);
 *     template <typename T> bool ProcessDepthMap(std::vector<std::string> vtiList,
 *                         std::vector<std::string> krtdList, double thresholdBestCost,
 *                         vtkDoubleArray* io_scalar);
 *
 * (defined in Reconstruction/CudaReconstruction.cu:269-298 and :302-386), plus the per-point loop of
 * MeshColoration::ProcessColoration (Coloration/MeshColoration.cxx:98-199).  File reading (VTK XML,
 * .krtd) stays on the caller's side of the boundary: every function takes plain pointers and sizes.
 *
 * Conventions shared by all entry points
 *   - every function returns DMI_OK (0) or a DMI_ERR_* code; it never calls exit() (the reference
 *     does, CudaReconstruction.cu:68-76).  dmi_last_error() gives the message.
 *   - 4x4 matrices are 16 doubles, row-major (the order vtkMatrixToTypeComputeTable produces,
 *     CudaReconstruction.cu:220-230).  K is the 3x3 intrinsics inside an identity 4x4
 *     (Sources/ReconstructionData.cxx:199-209), RT = [R | t] with last row 0 0 0 1 (Sources/Helper.h:161-165).
 *   - images are bottom-up (VTK), row r of storage = image row H-1-r (CudaReconstruction.cu:141-149).
 *   - depth -1 means "no measurement" (CudaReconstruction.cu:202).
 *   - the volume is the reference's cell array: (dims[0]-1)*(dims[1]-1)*(dims[2]-1) scalars, id =
 *     (k*Ny + j)*Nx + i (CudaReconstruction.cu:126-134), where dims are vtkImageData POINT dims.
 *   - one context = one GPU = one caller thread.  Calls taking HOST pointers are synchronous; calls
 *     taking DEVICE pointers are asynchronous on the context's stream (see dmi_set_stream).
 *   - there is no CPU fallback anywhere behind this header.
 */
#ifndef DMI_B200_H
#define DMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMI_ABI_VERSION 2

/* status codes */
#define DMI_OK 0
#define DMI_ERR_INVALID_ARGUMENT 1   /* null pointer, bad size, bad enum */
#define DMI_ERR_NOT_INITIALIZED 2    /* dmi_initialize / volume_begin not called yet */
#define DMI_ERR_CUDA 3               /* a CUDA runtime call failed; message has the CUDA error */
#define DMI_ERR_NO_VIEWS 4           /* nViews == 0 (CudaReconstruction.cu:308-312 returns false) */
#define DMI_ERR_BAD_PARAMETERS 5     /* Rho == 0 && Thick == 0 (vtkCudaReconstructionFilter.cxx:138-142) */
#define DMI_ERR_OUT_OF_MEMORY 6

/* scalar types: TVolumetric of ProcessDepthMap<T> (CudaReconstruction.cu:390-400); mesh point storage */
#define DMI_F32 0
#define DMI_F64 1

/* integration kernel selection (dmi_set_option DMI_OPT_TSDF_KERNEL) */
#define DMI_TSDF_KERNEL_AUTO 0       /* certified fast path with exact fallback */
#define DMI_TSDF_KERNEL_EXACT 1      /* every voxel*view through the reference's exact op sequence */

#define DMI_OPT_TSDF_KERNEL 1
#define DMI_OPT_VIEW_CHUNK 2         /* views per launch (0 = auto) */
#define DMI_OPT_TIER_COUNTERS 3      /* 1: count tier decisions of the fast kernel (diagnostic build of the kernel) */
#define DMI_OPT_CULL 4               /* 1 (default): skip (brick, view) pairs that provably contribute nothing */
#define DMI_OPT_BRICK_QUOTA 5        /* tuning: bricks a CTA of the persistent integration kernel takes before it retires
                                        (default 32; the sharded entry points use 8 so that the NCCL kernels of the
                                        view exchange find a free SM slot quickly) */

typedef struct dmi_ctx dmi_ctx;

/* ---- context ------------------------------------------------------------------------------- */

int dmi_abi_version(void);
int dmi_device_count(int* count);
/* Creates a context bound to CUDA device `device`.  On failure *ctx is NULL and
 * dmi_last_error(NULL) holds the reason (thread-local). */
int dmi_create(int device, dmi_ctx** ctx);
int dmi_destroy(dmi_ctx* ctx);
const char* dmi_last_error(const dmi_ctx* ctx);
/* Run the context's work on an existing CUDA stream (a cudaStream_t passed as void*), e.g. the
 * caller's framework stream.  NULL is the legacy default stream, as in the CUDA runtime.
 * dmi_use_own_stream goes back to the non-blocking stream the context created for itself. */
int dmi_set_stream(dmi_ctx* ctx, void* cuda_stream);
int dmi_use_own_stream(dmi_ctx* ctx);
int dmi_synchronize(dmi_ctx* ctx);
int dmi_set_option(dmi_ctx* ctx, int option, long long value);

/* ---- TSDF integration ---------------------------------------------------------------------- */

/* Replaces CudaInitialize (CudaReconstruction.cu:269-298): same arguments, same meaning, with the
 * vtkMatrix4x4 flattened to 16 row-major doubles.  gridDims are POINT dims.  State lives in the
 * context, not in process-global __constant__ symbols, so contexts are independent. */
int dmi_initialize(dmi_ctx* ctx, const double gridMatrix[16], const int gridDims[3],
                   const double gridOrig[3], const double gridSpacing[3],
                   double rayPotentialThick, double rayPotentialRho, double rayPotentialEta,
                   double rayPotentialDelta, const int depthMapDims[2]);

/* z-slab sharding (new; the reference is single-GPU): this context owns cells k in [k0, k1) of the
 * grid given to dmi_initialize.  Voxel centres keep their GLOBAL index arithmetic
 * (orig + (k + 0.5) * spacing, CudaReconstruction.cu:78-83), so slabs concatenate bit-identically.
 * Default after dmi_initialize: the whole grid.  Volume pointers below then cover the slab only. */
int dmi_set_slab(dmi_ctx* ctx, int k0, int k1);
/* Layered sharding: this context owns the z-layers phase, phase + stride, phase + 2*stride, ... of `layerPlanes`
 * cells each (a multiple of 32; the grid's last layer may be shorter).  Its volume holds them packed in that order:
 * local plane lp = q * layerPlanes + r is the global plane (q * stride + phase) * layerPlanes + r.  Dealing the layers
 * round-robin gives every GPU a share of every view's work (contiguous slabs see a view's work concentrated on
 * few GPUs), at the price of every GPU needing every view.  dmi_slab_planes: planes owned (either kind of slab). */
int dmi_set_slab_layers(dmi_ctx* ctx, int layerPlanes, int phase, int stride);
int dmi_slab_planes(dmi_ctx* ctx, int* planes);

/* Replaces ProcessDepthMap<T> (CudaReconstruction.cu:302-386) for views already in host memory:
 * uploads io_scalar (the call ACCUMULATES onto its content, :323-327), applies the best-cost
 * threshold filter of ReconstructionData::ApplyDepthThresholdFilter (depth := -1 where
 * bestCost > threshold; skipped when bestCost is NULL), integrates the views in list order, and
 * writes the result back into io_scalar.
 *   depths, bestCost  double[nViews][H][W]      K, RT  double[nViews][16]
 *   io_scalar         slab cells of type scalarType (DMI_F32 / DMI_F64)                        */
int dmi_process_depth_maps(dmi_ctx* ctx, int nViews, const double* depths, const double* bestCost,
                           double thresholdBestCost, const double* K, const double* RT,
                           void* io_scalar, int scalarType);

/* Streaming form of the same thing, for callers that read files view by view like the reference's
 * loop (:343-365) or keep data on the GPU:
 *   begin      allocates the slab on the device; uploads h_scalar, or zero-fills when NULL
 *   integrate  adds nViews views (host or device pointers; K/RT are always host pointers)
 *   end        downloads the slab into h_scalar (may be NULL to keep it on the device)          */
int dmi_volume_begin(dmi_ctx* ctx, const void* h_scalar, int scalarType);
int dmi_volume_integrate_host(dmi_ctx* ctx, int nViews, const double* depths, const double* bestCost,
                              double thresholdBestCost, const double* K, const double* RT);
int dmi_volume_integrate_device(dmi_ctx* ctx, int nViews, const double* d_depths,
                                const double* d_bestCost, double thresholdBestCost,
                                const double* K, const double* RT);
int dmi_volume_end(dmi_ctx* ctx, void* h_scalar);

/* View preparation split from integration, for multi-GPU runs: the rank that loaded a view prepares it
 * ONCE and the prepared arrays are exchanged, instead of every rank preparing every view:
 *   d_cls        float[nViews][H][W]  classification image: the depth rounded to float, -1.0f = invalid
 *                                     after the best-cost filter (never -1.0f on a valid pixel)
 *   d_lo         int32[nViews][H][W]  (optional) residual of the LOSSLESS split depth = (cls, lo): with it the
 *                                     double depth maps need not be exchanged or kept, 8 bytes per pixel
 *                                     carry everything (exact for |depth| >= 2^-64 or 0; smaller magnitudes
 *                                     keep 2^-117 absolute accuracy)
 *   d_tileStats  float[nViews][tileFloatsPerView]  tile statistics used by the brick tests (16-byte aligned;
 *                                     tileFloatsPerView is a multiple of 4)
 *   dmi_prepared_view_sizes    elements per view of d_cls / d_lo (W*H) and of d_tileStats
 *   dmi_prepare_views_device   fills the arrays (d_lo may be NULL); when clsSpareIndex >= 0,
 *                              d_cls[clsSpareIndex] is set to -1.0f
 *   dmi_volume_integrate_prepared  like dmi_volume_integrate_device, on prepared views.  d_depths = the
 *                              UNFILTERED maps (only pixels valid in d_cls are read), or NULL with d_lo given.
 *                              d_cls[clsSpareIndex] must hold -1.0f and lie within 2^31 floats of every
 *                              view of the call.
 * Asynchronous on the context's stream. */
int dmi_prepared_view_sizes(dmi_ctx* ctx, size_t* clsFloatsPerView, size_t* tileFloatsPerView);
int dmi_prepare_views_device(dmi_ctx* ctx, int nViews, const double* d_depths, const double* d_bestCost,
                             double thresholdBestCost, float* d_cls, int* d_lo, long long clsSpareIndex,
                             float* d_tileStats);
int dmi_volume_integrate_prepared(dmi_ctx* ctx, int nViews, const double* d_depths, const int* d_lo,
                                  const float* d_cls, long long clsSpareIndex, const float* d_tileStats,
                                  const double* K, const double* RT);
/* Device address and size in bytes of the slab (valid between begin and the next begin/destroy). */
int dmi_volume_device_ptr(dmi_ctx* ctx, void** d_ptr, size_t* bytes);

/* ReconstructionData::ApplyDepthThresholdFilter (Sources/ReconstructionData.cxx:138-167) on device
 * buffers, in place: d_depths[i] = -1 where d_bestCost[i] > threshold. */
int dmi_apply_depth_threshold_device(dmi_ctx* ctx, size_t count, double* d_depths,
                                     const double* d_bestCost, double thresholdBestCost);

/* Device time (ms, CUDA events on the context's stream) and launch count of the integration kernels
 * issued since the last call to this function. */
int dmi_tsdf_kernel_stats(dmi_ctx* ctx, float* ms, long long* launches);

/* Diagnostics of the certified fast path (DMI_OPT_TIER_COUNTERS = 1): out[0] voxel*views certified by
 * the FP32 tier, [1] sent to the FP64 tier, [2] sent to the exact tier, [3] exact because |diff| was
 * within the guard band of Delta, [4] voxel*views evaluated one by one after culling, [5] (brick, view)
 * pairs culled and [7] seen by the bricks the kernel visited (supertiles no view can touch are not visited),
 * [6] voxel*views in the FP64 band around the surface,
 * [8] voxel*views settled brick-wide as free space in front of the surface (one add per voxel),
 * [9] voxel*views classified far in front of the depth, [10] far behind it, [11] on an invalid pixel
 * (or rejected: behind the camera / outside the image) by the FP32 phase, [12] the part of [9] + [11] settled
 * by the validity-only phase (bricks in front of every valid depth of their footprint), [13..15] reserved (0).  Reading resets the counters. */
int dmi_tsdf_tier_counters(dmi_ctx* ctx, unsigned long long out[16]);

/* ---- mesh coloration ----------------------------------------------------------------------- */

/* Replaces the loop of MeshColoration::ProcessColoration (Coloration/MeshColoration.cxx:140-192)
 * together with ReconstructionData::TransformWorldToDepthMapPosition / GetColorValue
 * (Sources/ReconstructionData.cxx:169-182, 92-116) and help::ComputeMedian (Sources/Helper.h:174-187).
 *   xyz       nPoints * 3 coordinates, DMI_F32 (vtkPoints' default storage) or DMI_F64
 *   colors    uint8[nViews][H][W][3] bottom-up ("Color" arrays); W, H = view 0's dims (:110)
 *   mean      uint8[nPoints][3]  "MeanColoration"     (integer sum / count, truncated)
 *   median    uint8[nPoints][3]  "MedianColoration"   (even count: (a+b)/2 truncated)
 *   nbProjected int32[nPoints]   "NbProjectedDepthMap"
 * Points no view sees get 0/0/0 and count 0 (:116-118,124-126,132).                             */
int dmi_colorize(dmi_ctx* ctx, size_t nPoints, const void* xyz, int xyzType, int nViews,
                 const uint8_t* colors, const double* K, const double* RT, int W, int H,
                 uint8_t* mean, uint8_t* median, int32_t* nbProjected);
/* Same with every array except K/RT already on the device (asynchronous on the context stream). */
int dmi_colorize_device(dmi_ctx* ctx, size_t nPoints, const void* d_xyz, int xyzType, int nViews,
                        const uint8_t* d_colors, const double* K, const double* RT, int W, int H,
                        uint8_t* d_mean, uint8_t* d_median, int32_t* d_nbProjected);
int dmi_color_kernel_stats(dmi_ctx* ctx, float* ms, long long* launches);

/* ---- isosurface of the fused volume (the stage after the path: Reconstruction/main.cxx:151-189) -----------
 *
 * vtkCellDataToPointData -> vtkContourFilter(--contour) -> vtkTransformFilter(grid matrix) of the reference's CLI, on
 * the GPU, so that the fused volume (8.6 GB at 1024^3) need not travel to the host and dmi_colorize_device can take its
 * points from device memory.  Grid geometry = the one given to dmi_initialize.
 *   point scalars  average of the cells sharing the point (double)
 *   vertices       float32 world coordinates, one per grid edge whose ends differ in (scalar >= value); numbered by owning
 *                  point (k, j, i order), then axis -- the vertex set any edge-based contouring yields
 *   triangles      int32 vertex ids, per cell in k, j, i order, watertight, normals from inside (>= value) to outside;
 *                  the triangulation is this library's own (csrc/dmi_contour.cu), not VTK's
 * dmi_contour_device   d_cellScalars = NULL: the context's own volume (it must cover the whole grid).  Synchronous
 *                      (the counts come back to the host); the surface stays on the device until the next call.
 * dmi_contour          the same from host memory (cell scalars of the whole grid, VTK cell order)
 * dmi_contour_get      copies the last surface to host memory (either pointer may be NULL)
 * dmi_contour_device_ptr  device addresses of the last surface (valid until the next dmi_contour* / dmi_destroy)   */
int dmi_contour_device(dmi_ctx* ctx, const void* d_cellScalars, int scalarType, double value,
                       size_t* nVertices, size_t* nTriangles);
int dmi_contour(dmi_ctx* ctx, const void* cellScalars, int scalarType, double value,
                size_t* nVertices, size_t* nTriangles);
int dmi_contour_get(dmi_ctx* ctx, float* vertices, int32_t* triangles);
int dmi_contour_device_ptr(dmi_ctx* ctx, const float** d_vertices, const int32_t** d_triangles,
                           size_t* nVertices, size_t* nTriangles);

/* ---- sharding over several GPUs (new; the reference is single-GPU) ---------------------------------
 *
 * The grid is cut into z-layers of 32 cells dealt round-robin to the GPUs (dmi_set_slab_layers): every voxel has one
 * owner, so there is no reduction; every GPU needs every view, so the views' PREPARED form (classification float +
 * int32 residual = the lossless 8-byte split of the filtered double depth, + tile statistics; built once, by the rank
 * that loaded the view) is all-gathered with NCCL in groups of views, behind the integration of the previous group.
 * Results are bit-identical to the single-GPU volume.
 *
 * SPMD form -- one process (or thread) per GPU, each with its own context:
 *   dmi_comm_unique_id        ncclGetUniqueId on one rank; hand the 128 bytes to the others by any means
 *   dmi_comm_init             ncclCommInitRank on the context's device (collective: every rank calls it)
 *   dmi_shard_initialize      dmi_initialize + this rank's layers
 *   dmi_shard_view_count /    which views this rank must supply ("the files it loads"), in the order expected:
 *   dmi_shard_view_indices    groups of consecutive views (128 in a full group, shorter groups at both ends of the list),
 *                             each split evenly over the ranks in contiguous shares
 *   dmi_shard_group_count /   the groups themselves: starts[0 .. count] (starts[count] = nViews); rank r's share of group g
 *   dmi_shard_group_starts    is [starts[g] + r * pg, starts[g] + (r + 1) * pg) clipped to the group, pg = ceil(size / world)
 *   dmi_volume_begin          as usual, on the rank's layers (packed)
 *   dmi_shard_integrate_*     collective; myDepths / myBestCost hold THIS RANK'S views only, K / RT all views
 *   dmi_shard_gather_volume_device   collective; root receives the whole grid in VTK cell order (device memory)
 *   dmi_volume_end            as usual: the rank's layers to host memory
 *   dmi_shard_range           contiguous split of n items (coloration: points; colour images)
 *   dmi_shard_colorize_device collective; colours THIS RANK'S points (dmi_shard_range) with all views, of which it
 *                             supplies its own block (dmi_shard_range of the views); the blocks are all-gathered
 * NCCL is loaded with dlopen("libnccl.so.2") at first use: a process that already holds one (PyTorch) shares it. */
#define DMI_UNIQUE_ID_BYTES 128
int dmi_comm_unique_id(unsigned char id[DMI_UNIQUE_ID_BYTES]);
int dmi_comm_init(dmi_ctx* ctx, const unsigned char id[DMI_UNIQUE_ID_BYTES], int rank, int world);
int dmi_comm_destroy(dmi_ctx* ctx);
int dmi_comm_info(dmi_ctx* ctx, int* rank, int* world, int* ncclVersion);
/* 1 when the view exchange of this communicator runs on the copy engines (NCCL >= 2.28: zero-CTA policy, group buffers
 * registered as symmetric windows), 0 when NCCL's kernels carry it (older NCCL, or DMI_EXCHANGE=sm in the environment),
 * negative error code otherwise. */
int dmi_comm_copy_engines(dmi_ctx* ctx);
int dmi_shard_initialize(dmi_ctx* ctx, const double gridMatrix[16], const int gridDims[3],
                         const double gridOrig[3], const double gridSpacing[3],
                         double rayPotentialThick, double rayPotentialRho, double rayPotentialEta,
                         double rayPotentialDelta, const int depthMapDims[2]);
int dmi_shard_view_count(int nViews, int world, int rank, int* count);
int dmi_shard_view_indices(int nViews, int world, int rank, int* indices);
int dmi_shard_group_count(int nViews, int world, int* count);
int dmi_shard_group_starts(int nViews, int world, int* starts);
int dmi_shard_integrate_device(dmi_ctx* ctx, int nViews, const double* d_myDepths, const double* d_myBestCost,
                               double thresholdBestCost, const double* K, const double* RT);
int dmi_shard_integrate_host(dmi_ctx* ctx, int nViews, const double* myDepths, const double* myBestCost,
                             double thresholdBestCost, const double* K, const double* RT);
int dmi_shard_gather_volume_device(dmi_ctx* ctx, int root, void* d_full);
int dmi_shard_range(size_t n, int world, int rank, size_t* first, size_t* count);
int dmi_shard_colorize_device(dmi_ctx* ctx, size_t nMyPoints, const void* d_myXyz, int xyzType, int nViews,
                              const uint8_t* d_myColors, const double* K, const double* RT, int W, int H,
                              uint8_t* d_mean, uint8_t* d_median, int32_t* d_nbProjected);

/* Single-process form: one object drives all the GPUs listed (one host thread per GPU inside each call, NCCL
 * communicators created in one NCCL group call, like ncclCommInitAll).  This is the multi-GPU counterpart of the reference's two entry points:
 *   dmi_group_initialize          CudaInitialize (CudaReconstruction.cu:269-298)
 *   dmi_group_process_depth_maps  ProcessDepthMap<T> (:302-386): host pointers, ALL views, io_scalar = the whole grid
 *                                 (accumulated onto); each GPU uploads only the views and the layers it owns and
 *                                 writes its finished layers straight into io_scalar
 *   dmi_group_colorize            MeshColoration::ProcessColoration (MeshColoration.cxx:98-199): points sharded by
 *                                 contiguous index range; host pointers as in dmi_colorize                          */
typedef struct dmi_group dmi_group;
int dmi_group_create(const int* devices, int nDevices, dmi_group** grp);
int dmi_group_destroy(dmi_group* grp);
const char* dmi_group_last_error(const dmi_group* grp);
int dmi_group_size(const dmi_group* grp);
dmi_ctx* dmi_group_context(dmi_group* grp, int rank);
int dmi_group_set_option(dmi_group* grp, int option, long long value);
int dmi_group_initialize(dmi_group* grp, const double gridMatrix[16], const int gridDims[3],
                         const double gridOrig[3], const double gridSpacing[3],
                         double rayPotentialThick, double rayPotentialRho, double rayPotentialEta,
                         double rayPotentialDelta, const int depthMapDims[2]);
int dmi_group_process_depth_maps(dmi_group* grp, int nViews, const double* depths, const double* bestCost,
                                 double thresholdBestCost, const double* K, const double* RT,
                                 void* io_scalar, int scalarType);
int dmi_group_colorize(dmi_group* grp, size_t nPoints, const void* xyz, int xyzType, int nViews,
                       const uint8_t* colors, const double* K, const double* RT, int W, int H,
                       uint8_t* mean, uint8_t* median, int32_t* nbProjected);

/* ---- measurement helpers ------------------------------------------------------------------- */

/* Number of kernels this library has launched on behalf of the context since dmi_create. */
int dmi_launch_counter(dmi_ctx* ctx, long long* launches);

/* Issue-rate microbenchmarks used for the roofline denominators that MEASURED_PEAKS.json lacks.
 * which: 0 = FP64 DFMA, 1 = FP32 FFMA.  Returns sustained TFLOP/s (FMA = 2 flops) over `ms_target`
 * milliseconds of back-to-back launches. */
int dmi_measure_fp_peak(dmi_ctx* ctx, int which, double ms_target, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* DMI_B200_H */
