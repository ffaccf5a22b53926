"""dmi_plan_slab_tile_intervals: the tiles of each view that a z-slab can gather from (planning of cropped view
exchanges).  Pure host code: checked here against a brute-force projection of voxel centres."""
import numpy as np
import pytest

from cudadepthmapintegration_b200 import synthetic as syn
from cudadepthmapintegration_b200.engine import plan_slab_tile_intervals
from cudadepthmapintegration_b200._lib import DmiError


def _project(grid, K16, RT16, idx):
    """Pixel (px, py) of voxel centres idx [m, 3] as the reference computes it (CudaReconstruction.cu:158-197)."""
    gm = grid.matrix.reshape(4, 4)
    c = grid.origin + (idx + 0.5) * grid.spacing
    w = c @ gm[:3, :3].T + gm[:3, 3]
    RT = RT16.reshape(4, 4); K = K16.reshape(4, 4)
    cam = w @ RT[:3, :3].T + RT[:3, 3]
    h = cam @ K[:3, :3].T + K[:3, 3]
    ok = h[:, 2] > 0
    u, v = h[ok, 0] / h[ok, 2], h[ok, 1] / h[ok, 2]
    return np.floor(np.abs(u) + 0.5) * np.sign(u), np.floor(np.abs(v) + 0.5) * np.sign(v)


@pytest.mark.parametrize("rotate", [0.0, 30.0])
def test_every_voxel_of_the_slab_lands_in_its_intervals(rotate):
    n, W, H, V, world = 64, 320, 240, 24, 4
    grid = syn.make_grid(n, rotate_deg=rotate)
    K, RT = syn.make_cameras(V, W, H)
    rng = np.random.RandomState(3)
    total = 0.0
    for r in range(world):
        k0, k1 = r * n // world, (r + 1) * n // world
        first, last = plan_slab_tile_intervals(grid.matrix, grid.point_dims, grid.origin, grid.spacing, (W, H), K, RT, k0, k1)
        assert first.shape == (V, (H + 7) // 8)
        # all 8 corners' neighbourhood + random voxels of the slab
        idx = np.stack([rng.randint(0, n, 4000), rng.randint(0, n, 4000), rng.randint(k0, k1, 4000)], axis=1).astype(np.float64)
        corners = np.array([[i, j, k] for i in (0, n - 1) for j in (0, n - 1) for k in (k0, k1 - 1)], dtype=np.float64)
        idx = np.concatenate([idx, corners])
        for v in range(V):
            px, py = _project(grid, K[v], RT[v], idx)
            inside = (px >= 0) & (px < W) & (py >= 0) & (py < H)
            row = (H - 1 - py[inside]).astype(int)            # storage row of the bottom-up image
            tr, tc = row // 8, (px[inside].astype(int)) // 8
            assert np.all(first[v, tr] <= tc) and np.all(tc <= last[v, tr]), (r, v)
        width = np.maximum(last.astype(int) - first.astype(int) + 1, 0)
        total += width.sum() / (V * first.shape[1] * ((W + 7) // 8))
    # a quarter slab needs far less than the whole image on average
    assert total / world < 0.6


def test_camera_inside_the_box_needs_everything_and_bad_slabs_are_rejected():
    n, W, H = 16, 64, 48
    grid = syn.make_grid(n)
    K, RT = syn.make_cameras(2, W, H, radius=0.5)             # camera centres inside the 2.4-box
    first, last = plan_slab_tile_intervals(grid.matrix, grid.point_dims, grid.origin, grid.spacing, (W, H), K, RT, 0, n)
    assert np.all(first == 0) and np.all(last == (W + 7) // 8 - 1)
    first, last = plan_slab_tile_intervals(grid.matrix, grid.point_dims, grid.origin, grid.spacing, (W, H), K, RT, 5, 5)
    assert np.all(first > last)                               # empty slab: nothing needed
    with pytest.raises(DmiError):
        plan_slab_tile_intervals(grid.matrix, grid.point_dims, grid.origin, grid.spacing, (W, H), K, RT, 3, n + 1)
