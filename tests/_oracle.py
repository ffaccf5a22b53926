"""ctypes bindings of the CHECKERS under oracle/ (test infrastructure; never imported by the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

F32, F64 = 0, 1
_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.oracle_ray_potential.restype = C.c_double
        lib.oracle_ray_potential.argtypes = [_d] * 6
        lib.oracle_round_to_pixel.restype = C.c_int
        lib.oracle_round_to_pixel.argtypes = [_d]
        lib.oracle_apply_depth_threshold.restype = None
        lib.oracle_apply_depth_threshold.argtypes = [_vp, _vp, _sz, _d]
        lib.oracle_tsdf_integrate.restype = None
        lib.oracle_tsdf_integrate.argtypes = [_vp, _vp, _vp, _vp, _d, _d, _d, _d, _i, _i, _i, _vp, _vp, _d, _vp, _vp,
                                              _i, _vp, _i, _i, _vp]
        lib.oracle_tsdf_decisions.restype = None
        lib.oracle_tsdf_decisions.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i]
        lib.oracle_world_to_pixel.restype = None
        lib.oracle_world_to_pixel.argtypes = [_vp, _vp, _vp, _vp]
        lib.oracle_median.restype = C.c_double
        lib.oracle_median.argtypes = [_vp, _sz, _vp]
        lib.oracle_colorize.restype = None
        lib.oracle_colorize.argtypes = [_sz, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]

        lib.oracle_threads.restype = C.c_int
        lib.oracle_threads.argtypes = [_i]

    def threads(self, n=0):
        """Sets (n > 0) and returns the number of host threads the oracle's OpenMP loops use."""
        return int(self.lib.oracle_threads(int(n)))

    def ray_potential(self, real, depth, thick, rho, eta, delta):
        return self.lib.oracle_ray_potential(real, depth, thick, rho, eta, delta)

    def round_to_pixel(self, u):
        return self.lib.oracle_round_to_pixel(u)

    def tsdf_integrate(self, grid, rp, W, H, depths, best_cost, threshold, K, RT, io_scalar, k0=None, k1=None):
        """io_scalar: FULL volume array (float32/float64), accumulated in place for cells k in [k0,k1)."""
        gm = _f64(grid.matrix); pd = np.ascontiguousarray(grid.point_dims, dtype=np.int32)
        og = _f64(grid.origin); sp = _f64(grid.spacing)
        K = _f64(K); RT = _f64(RT); depths = _f64(depths)
        n = K.size // 16
        bc = None if best_cost is None else _f64(best_cost)
        scratch = np.empty(W * H, dtype=np.float64)
        k0 = 0 if k0 is None else k0
        k1 = grid.n_cells[2] if k1 is None else k1
        st = F64 if io_scalar.dtype == np.float64 else F32
        self.lib.oracle_tsdf_integrate(_p(gm), _p(pd), _p(og), _p(sp), rp.thick, rp.rho, rp.eta, rp.delta, W, H, n,
                                       _p(depths), _p(bc), float(threshold), _p(K), _p(RT), st, _p(io_scalar), k0, k1,
                                       _p(scratch))
        return io_scalar

    def tsdf_decisions(self, grid, W, H, depth, K, RT, k0=None, k1=None):
        gm = _f64(grid.matrix); pd = np.ascontiguousarray(grid.point_dims, dtype=np.int32)
        og = _f64(grid.origin); sp = _f64(grid.spacing)
        k0 = 0 if k0 is None else k0
        k1 = grid.n_cells[2] if k1 is None else k1
        out = np.empty(grid.n_cells[0] * grid.n_cells[1] * (k1 - k0), dtype=np.int32)
        self.lib.oracle_tsdf_decisions(_p(gm), _p(pd), _p(og), _p(sp), W, H, _p(_f64(depth)), _p(_f64(K)), _p(_f64(RT)),
                                       _p(out), k0, k1)
        return out

    def apply_depth_threshold(self, depths, best_cost, threshold):
        d = _f64(depths).copy()
        self.lib.oracle_apply_depth_threshold(_p(d), _p(_f64(best_cost)), d.size, float(threshold))
        return d

    def world_to_pixel(self, K4, RT4, p):
        out = np.zeros(2, dtype=np.int32)
        self.lib.oracle_world_to_pixel(_p(_f64(K4)), _p(_f64(RT4)), _p(_f64(p)), _p(out))
        return int(out[0]), int(out[1])

    def median(self, values):
        v = _f64(values)
        work = np.empty_like(v)
        return self.lib.oracle_median(_p(v), v.size, _p(work))

    def colorize(self, xyz, colors, K, RT, W, H, p0=None, p1=None):
        xyz = np.ascontiguousarray(xyz)
        assert xyz.dtype in (np.float32, np.float64)
        P = xyz.size // 3
        K = _f64(K); RT = _f64(RT)
        n = K.size // 16
        colors = np.ascontiguousarray(colors, dtype=np.uint8)
        mean = np.zeros((P, 3), dtype=np.uint8)
        median = np.zeros((P, 3), dtype=np.uint8)
        nb = np.zeros(P, dtype=np.int32)
        p0 = 0 if p0 is None else p0
        p1 = P if p1 is None else p1
        self.lib.oracle_colorize(p0, p1, _p(xyz), F64 if xyz.dtype == np.float64 else F32, n, _p(colors), _p(K), _p(RT),
                                 W, H, _p(mean), _p(median), _p(nb))
        return mean, median, nb


_oracle = None


def load_oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        _oracle = Oracle(C.CDLL(path))
    return _oracle


class RefHost:
    """The reference's own kernel text compiled for the CPU (oracle/ref_host_harness.cpp)."""

    def __init__(self, lib):
        self.lib = lib
        lib.ref_host_initialize.restype = None
        lib.ref_host_initialize.argtypes = [_vp, _vp, _vp, _vp, _d, _d, _d, _d, _vp]
        lib.ref_host_process.restype = None
        lib.ref_host_process.argtypes = [_i, _vp, _vp, _vp, _i, _vp, _i, _i]
        lib.ref_host_threads.restype = C.c_int
        lib.ref_host_threads.argtypes = [_i]

    def threads(self, n=0):
        """Sets (n > 0) and returns the number of host threads the harness' OpenMP loop uses."""
        return int(self.lib.ref_host_threads(int(n)))

    def run(self, grid, rp, W, H, depths, K, RT, io_scalar, kz0=None, kz1=None):
        gm = _f64(grid.matrix); pd = np.ascontiguousarray(grid.point_dims, dtype=np.int32)
        og = _f64(grid.origin); sp = _f64(grid.spacing)
        dd = np.array([W, H], dtype=np.int32)
        self.lib.ref_host_initialize(_p(gm), _p(pd), _p(og), _p(sp), rp.thick, rp.rho, rp.eta, rp.delta, _p(dd))
        K = _f64(K); RT = _f64(RT); depths = _f64(depths)
        kz0 = 0 if kz0 is None else kz0
        kz1 = grid.n_cells[2] if kz1 is None else kz1
        self.lib.ref_host_process(K.size // 16, _p(depths), _p(K), _p(RT), F64 if io_scalar.dtype == np.float64 else F32,
                                  _p(io_scalar), kz0, kz1)
        return io_scalar


class RefCuda:
    """The reference's own kernel compiled by nvcc for sm_100a (oracle/ref_cuda_harness.cu)."""

    def __init__(self, lib):
        self.lib = lib
        lib.ref_cuda_initialize.restype = None
        lib.ref_cuda_initialize.argtypes = [_vp, _vp, _vp, _vp, _d, _d, _d, _d, _vp]
        lib.ref_cuda_process.restype = C.c_int
        lib.ref_cuda_process.argtypes = [_i, _vp, _vp, _vp, _i, _vp, _i, _vp]

    def run(self, grid, rp, W, H, depths, K, RT, io_scalar, per_kernel_events=False):
        gm = _f64(grid.matrix); pd = np.ascontiguousarray(grid.point_dims, dtype=np.int32)
        og = _f64(grid.origin); sp = _f64(grid.spacing)
        dd = np.array([W, H], dtype=np.int32)
        self.lib.ref_cuda_initialize(_p(gm), _p(pd), _p(og), _p(sp), rp.thick, rp.rho, rp.eta, rp.delta, _p(dd))
        K = _f64(K); RT = _f64(RT); depths = _f64(depths)
        timing = np.zeros(2, dtype=np.float32)
        rc = self.lib.ref_cuda_process(K.size // 16, _p(depths), _p(K), _p(RT),
                                       F64 if io_scalar.dtype == np.float64 else F32, _p(io_scalar),
                                       1 if per_kernel_events else 0, _p(timing))
        if rc != 0:
            raise RuntimeError(f"reference CUDA harness failed ({rc})")
        return io_scalar, float(timing[0]), float(timing[1])


def _load_ref(name, cls):
    path = os.path.join(ORACLE_DIR, "_ref", name)
    if not os.path.exists(path):
        if os.path.exists("/root/reference/Reconstruction/CudaReconstruction.cu"):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(path):
            return None
    return cls(C.CDLL(path))


class RefColoration:
    """The reference's own MeshColoration.cxx / ReconstructionData.cxx / Helper.h compiled against the VTK stand-in
    (oracle/ref_coloration_harness.cpp -> oracle/_ref/libref_coloration.so)."""

    def __init__(self, lib):
        self.lib = lib
        lib.ref_coloration_run.restype = C.c_int
        lib.ref_coloration_run.argtypes = [C.c_char_p, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]
        lib.ref_world_to_pixel.restype = C.c_int
        lib.ref_world_to_pixel.argtypes = [C.c_char_p, _vp, _vp, _vp, _vp]

    def colorize(self, xyz, colors, K, RT, W, H):
        import tempfile
        xyz = np.ascontiguousarray(xyz)
        assert xyz.dtype in (np.float32, np.float64)
        P = xyz.size // 3
        K = _f64(K); RT = _f64(RT)
        n = K.size // 16
        colors = np.ascontiguousarray(colors, dtype=np.uint8)
        assert colors.size == n * W * H * 3
        mean = np.zeros((P, 3), dtype=np.uint8)
        median = np.zeros((P, 3), dtype=np.uint8)
        nb = np.zeros(P, dtype=np.int32)
        with tempfile.TemporaryDirectory(prefix="dmi_refcolor_") as d:
            rc = self.lib.ref_coloration_run(d.encode(), P, _p(xyz), F64 if xyz.dtype == np.float64 else F32, n, _p(colors),
                                             _p(K), _p(RT), W, H, _p(mean), _p(median), _p(nb))
        if rc != 0:
            raise RuntimeError(f"reference coloration harness failed ({rc})")
        return mean, median, nb

    def world_to_pixel(self, K4, RT4, p):
        import tempfile
        out = np.zeros(2, dtype=np.int32)
        with tempfile.TemporaryDirectory(prefix="dmi_refcolor_") as d:
            rc = self.lib.ref_world_to_pixel(d.encode(), _p(_f64(K4)), _p(_f64(RT4)), _p(_f64(p)), _p(out))
        if rc != 0:
            raise RuntimeError(f"reference coloration harness failed ({rc})")
        return int(out[0]), int(out[1])


class OperatorHarness:
    """oracle/adapter_harness.cpp: the reference's own vtkCudaReconstructionFilter / MeshColoration classes driven through
    their public interfaces -- over adapters/*.cxx + libdmi_b200.so (libadapter_vtk.so) or over the reference's own
    CudaReconstruction.cu / MeshColoration.cxx (libref_full.so)."""

    def __init__(self, lib):
        self.lib = lib
        lib.harness_filter_run.restype = C.c_int
        lib.harness_filter_run.argtypes = [C.c_char_p, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]
        lib.harness_coloration_run.restype = C.c_int
        lib.harness_coloration_run.argtypes = [C.c_char_p, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]

    def reconstruct(self, grid, rp, W, H, depths, best_cost, threshold, K, RT):
        import tempfile
        gm = _f64(grid.matrix); pd = np.ascontiguousarray(grid.point_dims, dtype=np.int32)
        og = _f64(grid.origin); sp = _f64(grid.spacing)
        K = _f64(K); RT = _f64(RT); depths = _f64(depths)
        bc = None if best_cost is None else _f64(best_cost)
        out = np.full(grid.n_voxels, np.nan)
        sec = C.c_double(0.0)
        with tempfile.TemporaryDirectory(prefix="dmi_harness_") as d:
            rc = self.lib.harness_filter_run(d.encode(), _p(gm), _p(pd), _p(og), _p(sp), rp.thick, rp.rho, rp.eta, rp.delta,
                                             float(threshold), K.size // 16, _p(depths), _p(bc), _p(K), _p(RT), W, H, _p(out),
                                             C.addressof(sec))
        if rc != 0:
            raise RuntimeError(f"operator harness: filter failed ({rc})")
        return out, sec.value

    def colorize(self, xyz, colors, K, RT, W, H):
        import tempfile
        xyz = np.ascontiguousarray(xyz)
        P = xyz.size // 3
        K = _f64(K); RT = _f64(RT)
        colors = np.ascontiguousarray(colors, dtype=np.uint8)
        mean = np.zeros((P, 3), dtype=np.uint8); median = np.zeros((P, 3), dtype=np.uint8); nb = np.zeros(P, dtype=np.int32)
        with tempfile.TemporaryDirectory(prefix="dmi_harness_") as d:
            rc = self.lib.harness_coloration_run(d.encode(), P, _p(xyz), F64 if xyz.dtype == np.float64 else F32, K.size // 16,
                                                 _p(colors), _p(K), _p(RT), W, H, _p(mean), _p(median), _p(nb))
        if rc != 0:
            raise RuntimeError(f"operator harness: coloration failed ({rc})")
        return mean, median, nb


def load_adapter_vtk():
    return _load_ref("libadapter_vtk.so", OperatorHarness)


def load_ref_full():
    return _load_ref("libref_full.so", OperatorHarness)


def load_ref_coloration():
    return _load_ref("libref_coloration.so", RefColoration)


def load_ref_host():
    return _load_ref("libref_tsdf_host.so", RefHost)


def load_ref_cuda(nofma=False):
    return _load_ref("libref_tsdf_cuda_nofma.so" if nofma else "libref_tsdf_cuda.so", RefCuda)
