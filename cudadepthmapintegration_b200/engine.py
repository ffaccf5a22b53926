"""Thin Python handle on one dmi context (= one GPU).  Pointers in, status codes out; numpy arrays are
passed by address, device buffers as integer addresses (e.g. ``tensor.data_ptr()``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DMI_F32, DMI_F64, DmiError  # noqa: F401  (re-exported)


def _ptr(a):
    """address of a C-contiguous numpy array, an int device address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return C.c_void_p(a.ctypes.data)
    raise TypeError(f"unsupported buffer type {type(a)}")


def _f64(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None and a.size % shape_last != 0:
        raise ValueError("bad array size")
    return a


def scalar_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float64:
        return DMI_F64
    if dt == np.float32:
        return DMI_F32
    raise ValueError("scalar type must be float32 or float64 (TVolumetric of ProcessDepthMap<T>)")


class Context:
    """One GPU.  All numeric work happens in libdmi_b200.so; a missing library raises ImportError."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.dmi_create(int(device), C.byref(h))
        if rc != _lib.DMI_OK:
            msg = self._lib.dmi_last_error(None)
            raise DmiError(rc, msg.decode() if msg else "")
        self._h = h
        self.device = int(device)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.dmi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        _lib.check(self._h, rc)

    # -- plumbing ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        """Run on the given cudaStream_t handle (0 = legacy default stream); None = the context's own stream."""
        if cuda_stream is None:
            self._ck(self._lib.dmi_use_own_stream(self._h))
        else:
            self._ck(self._lib.dmi_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def synchronize(self):
        self._ck(self._lib.dmi_synchronize(self._h))

    def set_option(self, option: int, value: int):
        self._ck(self._lib.dmi_set_option(self._h, int(option), int(value)))

    # -- TSDF ----------------------------------------------------------------------------------
    def initialize(self, grid_matrix, grid_dims, grid_orig, grid_spacing, thick, rho, eta, delta, depth_map_dims):
        """CudaInitialize (CudaReconstruction.cu:269-298).  grid_dims are POINT dims."""
        gm = _f64(grid_matrix).reshape(16)
        gd = np.ascontiguousarray(grid_dims, dtype=np.int32).reshape(3)
        go = _f64(grid_orig).reshape(3)
        gs = _f64(grid_spacing).reshape(3)
        dd = np.ascontiguousarray(depth_map_dims, dtype=np.int32).reshape(2)
        self._ck(self._lib.dmi_initialize(self._h, _ptr(gm), _ptr(gd), _ptr(go), _ptr(gs),
                                          float(thick), float(rho), float(eta), float(delta), _ptr(dd)))
        self._cells = (int(gd[0]) - 1, int(gd[1]) - 1, int(gd[2]) - 1)
        self._slab = (0, self._cells[2])
        self._dd = (int(dd[0]), int(dd[1]))

    def set_slab(self, k0: int, k1: int):
        self._ck(self._lib.dmi_set_slab(self._h, int(k0), int(k1)))
        self._slab = (int(k0), int(k1))

    def set_slab_layers(self, layer_planes: int, phase: int, stride: int):
        """Own the z-layers phase, phase + stride, ... of `layer_planes` cells each (packed in that order)."""
        self._ck(self._lib.dmi_set_slab_layers(self._h, int(layer_planes), int(phase), int(stride)))
        self._slab = (0, self.slab_planes())

    def slab_planes(self) -> int:
        n = C.c_int()
        self._ck(self._lib.dmi_slab_planes(self._h, C.byref(n)))
        return int(n.value)

    @property
    def slab_cells(self) -> int:
        return self._cells[0] * self._cells[1] * (self._slab[1] - self._slab[0])

    # -- sharding over several GPUs (SPMD: one context per process / GPU) ---------------------------
    def comm_init(self, unique_id: bytes | None, rank: int, world: int):
        """ncclCommInitRank on this context's device (collective).  unique_id: 128 bytes from comm_unique_id()."""
        self._ck(self._lib.dmi_comm_init(self._h, unique_id, int(rank), int(world)))

    def comm_info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._lib.dmi_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return int(r.value), int(w.value), int(v.value)

    def comm_copy_engines(self):
        """True when the view exchange runs on the copy engines (see dmi_comm_copy_engines)."""
        rc = self._lib.dmi_comm_copy_engines(self._h)
        if rc < 0:
            self._ck(rc)
        return bool(rc)

    def shard_initialize(self, grid_matrix, grid_dims, grid_orig, grid_spacing, thick, rho, eta, delta, depth_map_dims):
        """dmi_initialize + this rank's z-layers (32 cells each, dealt round-robin)."""
        gm = _f64(grid_matrix).reshape(16)
        gd = np.ascontiguousarray(grid_dims, dtype=np.int32).reshape(3)
        go = _f64(grid_orig).reshape(3)
        gs = _f64(grid_spacing).reshape(3)
        dd = np.ascontiguousarray(depth_map_dims, dtype=np.int32).reshape(2)
        self._ck(self._lib.dmi_shard_initialize(self._h, _ptr(gm), _ptr(gd), _ptr(go), _ptr(gs),
                                                float(thick), float(rho), float(eta), float(delta), _ptr(dd)))
        self._cells = (int(gd[0]) - 1, int(gd[1]) - 1, int(gd[2]) - 1)
        self._dd = (int(dd[0]), int(dd[1]))
        self._slab = (0, self.slab_planes())

    def shard_integrate_device(self, n_views: int, d_my_depths: int, d_my_best_cost: int | None, threshold_best_cost, K, RT):
        """Collective.  d_my_*: THIS RANK'S views (shard_view_indices order); K, RT: all n_views views."""
        K = _f64(K, 16); RT = _f64(RT, 16)
        self._ck(self._lib.dmi_shard_integrate_device(self._h, int(n_views), _ptr(int(d_my_depths)) if d_my_depths else None,
                                                      _ptr(int(d_my_best_cost)) if d_my_best_cost else None,
                                                      float(threshold_best_cost), _ptr(K), _ptr(RT)))

    def shard_integrate_host(self, n_views: int, my_depths, my_best_cost, threshold_best_cost, K, RT):
        K = _f64(K, 16); RT = _f64(RT, 16)
        self._ck(self._lib.dmi_shard_integrate_host(self._h, int(n_views), _ptr(my_depths), _ptr(my_best_cost),
                                                    float(threshold_best_cost), _ptr(K), _ptr(RT)))

    def shard_gather_volume_device(self, root: int, d_full: int | None):
        self._ck(self._lib.dmi_shard_gather_volume_device(self._h, int(root), _ptr(int(d_full)) if d_full else None))

    def shard_colorize_device(self, n_my_points: int, d_my_xyz: int, xyz_dtype, n_views: int, d_my_colors: int, K, RT,
                              width: int, height: int, d_mean: int, d_median: int, d_nb: int):
        """Collective.  Colours this rank's points with all views; d_my_colors = this rank's block of the colour images."""
        K = _f64(K, 16); RT = _f64(RT, 16)
        self._ck(self._lib.dmi_shard_colorize_device(self._h, int(n_my_points), _ptr(int(d_my_xyz)) if d_my_xyz else None,
                                                     scalar_code(xyz_dtype), int(n_views),
                                                     _ptr(int(d_my_colors)) if d_my_colors else None, _ptr(K), _ptr(RT),
                                                     int(width), int(height), _ptr(int(d_mean)) if d_mean else None,
                                                     _ptr(int(d_median)) if d_median else None, _ptr(int(d_nb)) if d_nb else None))

    def process_depth_maps(self, depths, best_cost, threshold_best_cost, K, RT, io_scalar: np.ndarray):
        """ProcessDepthMap<T> (CudaReconstruction.cu:302-386) on in-memory views; accumulates onto io_scalar."""
        K = _f64(K, 16); RT = _f64(RT, 16)
        n = K.size // 16
        if RT.size // 16 < n:
            raise ValueError("not enough RT matrices for the K matrices")
        depths = _f64(depths)
        if best_cost is not None:
            best_cost = _f64(best_cost)
        if n > 0:
            npix = self._dd[0] * self._dd[1]
            if depths.size != n * npix or (best_cost is not None and best_cost.size != n * npix):
                raise ValueError("depth / best-cost arrays must hold nViews * H * W doubles")
        if io_scalar.size != self.slab_cells or not io_scalar.flags["C_CONTIGUOUS"]:
            raise ValueError("io_scalar must be a C-contiguous array of the slab's cells")
        self._ck(self._lib.dmi_process_depth_maps(self._h, n, _ptr(depths), _ptr(best_cost), float(threshold_best_cost),
                                                  _ptr(K), _ptr(RT), _ptr(io_scalar), scalar_code(io_scalar.dtype)))
        return io_scalar

    def volume_begin(self, h_scalar: np.ndarray | None, dtype=np.float64):
        if h_scalar is not None:
            dtype = h_scalar.dtype
            if h_scalar.size != self.slab_cells:
                raise ValueError("h_scalar must hold the slab's cells")
        self._vol_dtype = np.dtype(dtype)
        self._ck(self._lib.dmi_volume_begin(self._h, _ptr(h_scalar), scalar_code(dtype)))

    def volume_integrate_host(self, depths, best_cost, threshold_best_cost, K, RT):
        K = _f64(K, 16); RT = _f64(RT, 16)
        depths = _f64(depths)
        best_cost = None if best_cost is None else _f64(best_cost)
        self._ck(self._lib.dmi_volume_integrate_host(self._h, K.size // 16, _ptr(depths), _ptr(best_cost),
                                                     float(threshold_best_cost), _ptr(K), _ptr(RT)))

    def volume_integrate_device(self, n_views: int, d_depths: int, d_best_cost: int | None, threshold_best_cost, K, RT):
        K = _f64(K, 16); RT = _f64(RT, 16)
        self._ck(self._lib.dmi_volume_integrate_device(self._h, int(n_views), _ptr(int(d_depths)),
                                                       _ptr(int(d_best_cost)) if d_best_cost else None,
                                                       float(threshold_best_cost), _ptr(K), _ptr(RT)))

    def prepared_view_sizes(self):
        a, b = C.c_size_t(), C.c_size_t()
        self._ck(self._lib.dmi_prepared_view_sizes(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def prepare_views_device(self, n_views: int, d_depths: int, d_best_cost: int | None, threshold, d_cls: int,
                             cls_spare_index: int, d_tiles: int, d_lo: int | None = None):
        """d_lo: optional int32 residual image; (cls, lo) is then a lossless 8-byte form of the filtered depth."""
        self._ck(self._lib.dmi_prepare_views_device(self._h, int(n_views), _ptr(int(d_depths)),
                                                    _ptr(int(d_best_cost)) if d_best_cost else None, float(threshold),
                                                    _ptr(int(d_cls)), _ptr(int(d_lo)) if d_lo else None,
                                                    int(cls_spare_index), _ptr(int(d_tiles))))

    def volume_integrate_prepared(self, n_views: int, d_depths: int | None, d_cls: int, cls_spare_index: int, d_tiles: int,
                                  K, RT, d_lo: int | None = None):
        """Exactly one of d_depths (double maps) / d_lo (residual image of the split depth) is needed."""
        K = _f64(K, 16); RT = _f64(RT, 16)
        self._ck(self._lib.dmi_volume_integrate_prepared(self._h, int(n_views), _ptr(int(d_depths)) if d_depths else None,
                                                         _ptr(int(d_lo)) if d_lo else None, _ptr(int(d_cls)),
                                                         int(cls_spare_index), _ptr(int(d_tiles)), _ptr(K), _ptr(RT)))

    def volume_end(self, h_scalar: np.ndarray | None = None):
        if h_scalar is not None and (h_scalar.size != self.slab_cells or h_scalar.dtype != self._vol_dtype):
            raise ValueError("h_scalar must match the slab's size and scalar type")
        self._ck(self._lib.dmi_volume_end(self._h, _ptr(h_scalar)))
        return h_scalar

    def volume_device_ptr(self):
        p = C.c_void_p()
        b = C.c_size_t()
        self._ck(self._lib.dmi_volume_device_ptr(self._h, C.byref(p), C.byref(b)))
        return int(p.value or 0), int(b.value)

    def apply_depth_threshold_device(self, count: int, d_depths: int, d_best_cost: int, threshold: float):
        self._ck(self._lib.dmi_apply_depth_threshold_device(self._h, int(count), _ptr(int(d_depths)),
                                                            _ptr(int(d_best_cost)), float(threshold)))

    def tsdf_kernel_stats(self):
        ms = C.c_float()
        n = C.c_longlong()
        self._ck(self._lib.dmi_tsdf_kernel_stats(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def tsdf_tier_counters(self):
        out = (C.c_ulonglong * 16)()
        self._ck(self._lib.dmi_tsdf_tier_counters(self._h, out))
        return dict(zip(("t1_certified", "t2_entered", "t3_entered", "delta_guard", "units", "culled_brick_views",
                         "near_band", "brick_views", "uniform_front", "far_front", "far_behind", "invalid_or_rejected",
                         "validity_only"),
                        [int(x) for x in out]))

    # -- coloration ------------------------------------------------------------------------------
    def colorize(self, xyz: np.ndarray, colors: np.ndarray, K, RT, width: int, height: int):
        """MeshColoration::ProcessColoration's loop (MeshColoration.cxx:140-192) on in-memory views."""
        K = _f64(K, 16); RT = _f64(RT, 16)
        n = K.size // 16
        xyz = np.ascontiguousarray(xyz)
        if xyz.dtype not in (np.float32, np.float64):
            xyz = xyz.astype(np.float64)
        P = xyz.size // 3
        colors = np.ascontiguousarray(colors, dtype=np.uint8)
        if n > 0 and colors.size != n * width * height * 3:
            raise ValueError("colors must hold nViews * H * W * 3 bytes")
        mean = np.zeros((P, 3), dtype=np.uint8)
        median = np.zeros((P, 3), dtype=np.uint8)
        nb = np.zeros((P,), dtype=np.int32)
        self._ck(self._lib.dmi_colorize(self._h, P, _ptr(xyz), scalar_code(xyz.dtype), n, _ptr(colors), _ptr(K), _ptr(RT),
                                        int(width), int(height), _ptr(mean), _ptr(median), _ptr(nb)))
        return mean, median, nb

    def colorize_device(self, n_points: int, d_xyz: int, xyz_dtype, n_views: int, d_colors: int, K, RT,
                        width: int, height: int, d_mean: int, d_median: int, d_nb: int):
        K = _f64(K, 16); RT = _f64(RT, 16)
        self._ck(self._lib.dmi_colorize_device(self._h, int(n_points), _ptr(int(d_xyz)), scalar_code(xyz_dtype), int(n_views),
                                               _ptr(int(d_colors)), _ptr(K), _ptr(RT), int(width), int(height),
                                               _ptr(int(d_mean)), _ptr(int(d_median)), _ptr(int(d_nb))))

    # -- isosurface of the fused volume (Reconstruction/main.cxx:151-189) ---------------------------
    def contour_device(self, d_cell_scalars: int | None, dtype, value: float):
        """Surface of the cell scalars at d_cell_scalars (None: the context's own volume); returns (nVertices, nTriangles)."""
        nv, nt = C.c_size_t(), C.c_size_t()
        self._ck(self._lib.dmi_contour_device(self._h, _ptr(int(d_cell_scalars)) if d_cell_scalars else None,
                                              scalar_code(dtype), float(value), C.byref(nv), C.byref(nt)))
        return int(nv.value), int(nt.value)

    def contour(self, cell_scalars: np.ndarray, value: float):
        """vtkCellDataToPointData + contour + grid matrix on host cell scalars; returns (vertices f32 [n,3], triangles i32 [m,3])."""
        cs = np.ascontiguousarray(cell_scalars)
        nv, nt = C.c_size_t(), C.c_size_t()
        self._ck(self._lib.dmi_contour(self._h, _ptr(cs), scalar_code(cs.dtype), float(value), C.byref(nv), C.byref(nt)))
        return self.contour_get(int(nv.value), int(nt.value))

    def contour_get(self, n_vertices: int, n_triangles: int):
        v = np.empty((n_vertices, 3), dtype=np.float32)
        t = np.empty((n_triangles, 3), dtype=np.int32)
        self._ck(self._lib.dmi_contour_get(self._h, _ptr(v), _ptr(t)))
        return v, t

    def contour_device_ptr(self):
        pv, pt = C.c_void_p(), C.c_void_p()
        nv, nt = C.c_size_t(), C.c_size_t()
        self._ck(self._lib.dmi_contour_device_ptr(self._h, C.byref(pv), C.byref(pt), C.byref(nv), C.byref(nt)))
        return int(pv.value or 0), int(pt.value or 0), int(nv.value), int(nt.value)

    def color_kernel_stats(self):
        ms = C.c_float()
        n = C.c_longlong()
        self._ck(self._lib.dmi_color_kernel_stats(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # -- shared device buffers (multi-GPU view exchange over the copy engines) --------------------
    def launch_counter(self) -> int:
        n = C.c_longlong()
        self._ck(self._lib.dmi_launch_counter(self._h, C.byref(n)))
        return int(n.value)

    def measure_fp_peak(self, which: int, ms_target: float = 200.0) -> float:
        t = C.c_double()
        self._ck(self._lib.dmi_measure_fp_peak(self._h, int(which), float(ms_target), C.byref(t)))
        return float(t.value)


def comm_unique_id() -> bytes:
    """ncclGetUniqueId: call on one rank, hand the bytes to the others (e.g. torch.distributed.broadcast_object_list)."""
    lib = _lib.load()
    buf = C.create_string_buffer(_lib.DMI_UNIQUE_ID_BYTES)
    rc = lib.dmi_comm_unique_id(buf)
    if rc != _lib.DMI_OK:
        msg = lib.dmi_last_error(None)
        raise DmiError(rc, msg.decode() if msg else "")
    return buf.raw


def shard_view_indices(n_views: int, world: int, rank: int) -> np.ndarray:
    """Global indices of the views `rank` must supply to the sharded entry points, in the order expected."""
    lib = _lib.load()
    n = C.c_int()
    if lib.dmi_shard_view_count(int(n_views), int(world), int(rank), C.byref(n)) != _lib.DMI_OK:
        raise ValueError("bad arguments")
    out = np.zeros(max(n.value, 1), dtype=np.int32)
    if lib.dmi_shard_view_indices(int(n_views), int(world), int(rank), _ptr(out)) != _lib.DMI_OK:
        raise ValueError("bad arguments")
    return out[:n.value].copy()


def shard_groups(n_views: int, world: int):
    """[(g0, g1)] the groups of consecutive views the sharded entry points exchange one after the other."""
    lib = _lib.load()
    n = C.c_int()
    if lib.dmi_shard_group_count(int(n_views), int(world), C.byref(n)) != _lib.DMI_OK:
        raise ValueError("bad arguments")
    starts = np.zeros(n.value + 1, dtype=np.int32)
    if lib.dmi_shard_group_starts(int(n_views), int(world), _ptr(starts)) != _lib.DMI_OK:
        raise ValueError("bad arguments")
    return [(int(a), int(b)) for a, b in zip(starts[:-1], starts[1:])]


def shard_range(n: int, world: int, rank: int):
    """(first, count) of the contiguous block of n items that `rank` owns (points; colour images)."""
    lib = _lib.load()
    a, b = C.c_size_t(), C.c_size_t()
    if lib.dmi_shard_range(int(n), int(world), int(rank), C.byref(a), C.byref(b)) != _lib.DMI_OK:
        raise ValueError("bad arguments")
    return int(a.value), int(b.value)


def layer_cell_ranges(n_cells_z: int, world: int, rank: int, layer_planes: int = 32):
    """[(k0, k1)] global z-ranges of the layers rank owns, in the order its packed volume holds them."""
    out, q = [], 0
    while (q * world + rank) * layer_planes < n_cells_z:
        k0 = (q * world + rank) * layer_planes
        out.append((k0, min(n_cells_z, k0 + layer_planes)))
        q += 1
    return out


class Group:
    """All the GPUs of one box behind one object (single process; dmi_group_* of include/dmi_b200.h): the multi-GPU
    counterpart of CudaInitialize / ProcessDepthMap<T> / MeshColoration::ProcessColoration."""

    def __init__(self, devices):
        self._lib = _lib.load()
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        h = C.c_void_p()
        rc = self._lib.dmi_group_create(_ptr(devs), int(devs.size), C.byref(h))
        if rc != _lib.DMI_OK:
            msg = self._lib.dmi_last_error(None)
            raise DmiError(rc, msg.decode() if msg else "")
        self._h = h
        self.size = int(devs.size)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dmi_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != _lib.DMI_OK:
            msg = self._lib.dmi_group_last_error(self._h)
            raise DmiError(rc, msg.decode() if msg else "")

    def set_option(self, option: int, value: int):
        self._ck(self._lib.dmi_group_set_option(self._h, int(option), int(value)))

    def initialize(self, grid_matrix, grid_dims, grid_orig, grid_spacing, thick, rho, eta, delta, depth_map_dims):
        gm = _f64(grid_matrix).reshape(16)
        gd = np.ascontiguousarray(grid_dims, dtype=np.int32).reshape(3)
        go = _f64(grid_orig).reshape(3)
        gs = _f64(grid_spacing).reshape(3)
        dd = np.ascontiguousarray(depth_map_dims, dtype=np.int32).reshape(2)
        self._ck(self._lib.dmi_group_initialize(self._h, _ptr(gm), _ptr(gd), _ptr(go), _ptr(gs), float(thick), float(rho),
                                                float(eta), float(delta), _ptr(dd)))
        self._cells = (int(gd[0]) - 1) * (int(gd[1]) - 1) * (int(gd[2]) - 1)

    def process_depth_maps(self, depths, best_cost, threshold_best_cost, K, RT, io_scalar: np.ndarray):
        K = _f64(K, 16); RT = _f64(RT, 16)
        depths = _f64(depths)
        best_cost = None if best_cost is None else _f64(best_cost)
        if io_scalar.size != self._cells or not io_scalar.flags["C_CONTIGUOUS"]:
            raise ValueError("io_scalar must be a C-contiguous array of the whole grid's cells")
        self._ck(self._lib.dmi_group_process_depth_maps(self._h, K.size // 16, _ptr(depths), _ptr(best_cost),
                                                        float(threshold_best_cost), _ptr(K), _ptr(RT), _ptr(io_scalar),
                                                        scalar_code(io_scalar.dtype)))
        return io_scalar

    def colorize(self, xyz: np.ndarray, colors: np.ndarray, K, RT, width: int, height: int):
        xyz = np.ascontiguousarray(xyz)
        if xyz.dtype not in (np.float32, np.float64):
            xyz = xyz.astype(np.float64)
        P = xyz.size // 3
        K = _f64(K, 16); RT = _f64(RT, 16)
        colors = np.ascontiguousarray(colors, dtype=np.uint8)
        mean = np.zeros((P, 3), dtype=np.uint8)
        median = np.zeros((P, 3), dtype=np.uint8)
        nb = np.zeros(P, dtype=np.int32)
        self._ck(self._lib.dmi_group_colorize(self._h, P, _ptr(xyz), scalar_code(xyz.dtype), K.size // 16, _ptr(colors), _ptr(K),
                                              _ptr(RT), int(width), int(height), _ptr(mean), _ptr(median), _ptr(nb)))
        return mean, median, nb
