"""Host-side mirror of ``MeshColoration`` (Coloration/MeshColoration.h:42-62) over the C ABI.

The reference class takes a vtkPolyData and two list files, loads every view, and
``ProcessColoration()`` attaches three point-data arrays.  Here the mesh is its point array and the
views are in-memory arrays (the content of the .vti "Color" arrays and .krtd files, in list order).
"""
from __future__ import annotations

import sys

import numpy as np

from .engine import Context


class MeshColoration:
    def __init__(self, points=None, colors=None, K=None, RT=None, device: int = 0):
        self._ctx = Context(device)
        self._points = None
        self._views = None
        self._output = None
        if points is not None:
            self.SetInput(points)
        if colors is not None:
            K = np.asarray(K, dtype=np.float64).reshape(-1, 16)
            RT = np.asarray(RT, dtype=np.float64).reshape(-1, 16)
            colors = np.asarray(colors)
            if RT.shape[0] < colors.shape[0]:
                # MeshColoration.cxx:59-63: views stay empty, ProcessColoration then fails
                print("Error, not enough krtd file for each vti file", file=sys.stderr)
            else:
                self._views = (colors, K[:colors.shape[0]], RT[:colors.shape[0]])

    @classmethod
    def from_files(cls, points, vti_list: str, krtd_list: str, device: int = 0):
        """MeshColoration(vtkPolyData* mesh, std::string vtiList, std::string krtdList) (MeshColoration.cxx:52-72):
        every listed view is read now, by the VTK-free readers of dataset_io."""
        from . import dataset_io
        try:
            _, _, colors, K, RT = dataset_io.load_dataset(vti_list, krtd_list, need_color=True)
        except (OSError, ValueError) as e:
            print(f"Error, {e}", file=sys.stderr)
            return cls(points, device=device)
        return cls(points, colors, K, RT, device=device)

    def SetInput(self, points):
        """SetInput(vtkPolyData*): deep copy of the mesh (MeshColoration.cxx:85-91); vtkPoints keep their
        storage type (float32 by default)."""
        p = np.array(points, copy=True)
        if p.dtype not in (np.float32, np.float64):
            p = p.astype(np.float32)
        self._points = np.ascontiguousarray(p.reshape(-1, 3))

    def ProcessColoration(self) -> bool:
        if self._points is None or self._views is None or self._views[0].shape[0] == 0:
            print("Error when input has been set or during reading vti/krtd file path", file=sys.stderr)
            return False
        colors, K, RT = self._views
        H, W = colors.shape[1], colors.shape[2]     # view 0's dims are used for every view (:110)
        mean, median, nb = self._ctx.colorize(self._points, colors, K, RT, W, H)
        self._output = {"MeanColoration": mean, "MedianColoration": median, "NbProjectedDepthMap": nb}
        return True

    def GetOutput(self):
        """Point-data arrays by the names the reference gives them (MeshColoration.cxx:119,127,133)."""
        return self._output

    def close(self):
        self._ctx.close()
