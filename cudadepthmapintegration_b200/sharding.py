"""z-slab and point-range partitioning for 1/2/4/8 GPUs.

Every voxel has exactly one owner and depends only on read-only view data
(CudaReconstruction.cu:163,211), so shards never exchange voxel data: the only communication is the
view broadcast before integration and one concatenating gather after it.
"""
from __future__ import annotations


def slab_range(n_cells_z: int, rank: int, world: int):
    """Cells k in [k0, k1) owned by `rank`: contiguous in VTK cell order ((k*Ny + j)*Nx + i)."""
    return (rank * n_cells_z) // world, ((rank + 1) * n_cells_z) // world


def point_range(n_points: int, rank: int, world: int):
    return (rank * n_points) // world, ((rank + 1) * n_points) // world


def view_range(n_views: int, rank: int, world: int):
    """Views a rank loads from its own host memory before the all-gather of views."""
    return (rank * n_views) // world, ((rank + 1) * n_views) // world
