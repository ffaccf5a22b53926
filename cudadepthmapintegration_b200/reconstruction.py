"""Host-side mirror of the reference's operator for the integration path, over the C ABI.

``CudaReconstructionFilter`` keeps the setter names, argument meaning and error behaviour of
``vtkCudaReconstructionFilter`` (Reconstruction/vtkCudaReconstructionFilter.h:57-86,
.cxx:96-179).  VTK is not in this image, so the ``vtkImageData`` input grid becomes
(point dims, origin, spacing) and the two list files become in-memory views; file reading stays on
the caller's side of the boundary.  ``cuda_initialize`` / ``process_depth_map`` mirror the two free
functions the filter calls (CudaReconstruction.cu:269-298, :302-386).
"""
from __future__ import annotations

import sys
import time

import numpy as np

from .engine import Context


def cuda_initialize(ctx: Context, grid_matrix, grid_dims, grid_orig, grid_spacing,
                    ray_p_thick, ray_p_rho, ray_p_eta, ray_p_delta, depth_map_dims):
    """CudaInitialize(i_gridMatrix, h_gridDims, h_gridOrig, h_gridSpacing, h_rayPThick, h_rayPRho,
    h_rayPEta, h_rayPDelta, h_depthMapDims) -- CudaReconstruction.cu:269-277."""
    ctx.initialize(grid_matrix, grid_dims, grid_orig, grid_spacing,
                   ray_p_thick, ray_p_rho, ray_p_eta, ray_p_delta, depth_map_dims)


def process_depth_map(ctx: Context, depths, best_cost, K, RT, threshold_best_cost, io_scalar: np.ndarray) -> bool:
    """ProcessDepthMap<T>(vtiList, krtdList, thresholdBestCost, io_scalar) -- CudaReconstruction.cu:302-306,
    with the per-view file contents passed as arrays.  T is io_scalar's dtype.  Returns False, like the
    reference (:308-312), when there are no views."""
    K = np.asarray(K, dtype=np.float64)
    if K.size == 0 or np.asarray(RT).size == 0:
        print("Error, no depthMap or KRTD matrix have been loaded", file=sys.stderr)
        return False
    ctx.process_depth_maps(depths, best_cost, threshold_best_cost, K, RT, io_scalar)
    return True


class CudaReconstructionFilter:
    """vtkCudaReconstructionFilter without VTK.  Output = the "reconstruction_scalar" cell array."""

    def __init__(self, device: int = 0, scalar_type=np.float64):
        self._ctx = Context(device)
        self._scalar_type = np.dtype(scalar_type)   # the reference instantiates double (.cxx:175)
        self.GridMatrix = None
        self.RayPotentialRho = 0.0
        self.RayPotentialThickness = 0.0
        self.RayPotentialEta = 0.0
        self.RayPotentialDelta = 0.0
        self.ThresholdBestCost = 0.0
        self.ExecutionTime = -1.0
        self._grid = None
        self._views = None
        self._output = None

    # setters of vtkCudaReconstructionFilter.h:57-86
    def SetRayPotentialThickness(self, v): self.RayPotentialThickness = float(v)
    def SetRayPotentialRho(self, v): self.RayPotentialRho = float(v)
    def SetRayPotentialEta(self, v): self.RayPotentialEta = float(v)
    def SetRayPotentialDelta(self, v): self.RayPotentialDelta = float(v)
    def SetThresholdBestCost(self, v): self.ThresholdBestCost = float(v)
    def SetGridMatrix(self, m): self.GridMatrix = np.asarray(m, dtype=np.float64).reshape(16).copy()
    def GetExecutionTime(self): return self.ExecutionTime

    def SetInputGrid(self, point_dims, origin, spacing):
        """Stands for SetInputData(vtkImageData): GetDimensions/GetOrigin/GetSpacing (.cxx:121-126)."""
        self._grid = (tuple(int(x) for x in point_dims), np.asarray(origin, dtype=np.float64),
                      np.asarray(spacing, dtype=np.float64))

    def SetViews(self, depths, best_cost, K, RT):
        """In-memory form of SetFilePathVTI / SetFilePathKRTD: the content of the listed files, in list order."""
        self._views = (depths, best_cost, np.asarray(K, dtype=np.float64), np.asarray(RT, dtype=np.float64))

    # vtkSetMacro(FilePathKRTD / FilePathVTI, std::string), vtkCudaReconstructionFilter.h:75-77: the list files are read
    # at Update() by the VTK-free readers of dataset_io (help::ExtractAllFilePath, ReadKrtdFile, the .vti point arrays)
    def SetFilePathKRTD(self, path): self.FilePathKRTD = str(path)
    def SetFilePathVTI(self, path): self.FilePathVTI = str(path)

    def Update(self) -> int:
        """RequestData (.cxx:96-151): returns 1 on success, 0 on the reference's error paths."""
        self.ExecutionTime = -1.0
        start = time.perf_counter()
        if self._views is None and getattr(self, "FilePathKRTD", "") and getattr(self, "FilePathVTI", ""):
            from . import dataset_io
            try:
                d, c, _, K, RT = dataset_io.load_dataset(self.FilePathVTI, self.FilePathKRTD)
            except (OSError, ValueError) as e:
                print(f"Error : {e}", file=sys.stderr)
                return 0
            self._views = (d, c, K, RT)
        if self._views is None or self._grid is None:
            print("Error, some inputs have not been set.", file=sys.stderr)
            return 0
        dims, orig, spacing = self._grid
        n_cells = (dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1)
        out = np.zeros(n_cells, dtype=self._scalar_type)          # FillComponent(0, 0), .cxx:133
        self._output = out
        if self.RayPotentialRho == 0 and self.RayPotentialThickness == 0:
            print("Error : Ray potential Rho or Thickness or both have not been set", file=sys.stderr)
            return 0
        self._compute(dims, orig, spacing, out)
        self.ExecutionTime = time.perf_counter() - start
        return 1

    def _compute(self, dims, orig, spacing, out) -> int:
        """Compute (.cxx:155-179)."""
        depths, best_cost, K, RT = self._views
        n_vti = np.asarray(depths).shape[0] if np.asarray(depths).ndim == 3 else K.size // 16
        if n_vti == 0 or RT.size // 16 < n_vti:
            print("Error : There is no enough vti files, please check your vtiList.txt and krtdList.txt", file=sys.stderr)
            return -1
        d0 = np.asarray(depths)
        depth_map_dims = (d0.shape[-1], d0.shape[-2])
        gm = self.GridMatrix if self.GridMatrix is not None else np.eye(4).reshape(16)
        cuda_initialize(self._ctx, gm, dims, orig, spacing, self.RayPotentialThickness, self.RayPotentialRho,
                        self.RayPotentialEta, self.RayPotentialDelta, depth_map_dims)
        process_depth_map(self._ctx, depths, best_cost, K, RT, self.ThresholdBestCost, out)
        return 0

    def GetOutput(self) -> np.ndarray:
        """The "reconstruction_scalar" cell data, shape (Nz, Ny, Nx) in VTK cell order."""
        dims = self._grid[0]
        return self._output.reshape(dims[2] - 1, dims[1] - 1, dims[0] - 1)

    def close(self):
        self._ctx.close()
