"""GPU isosurface (dmi_contour*, csrc/dmi_contour.cu: cell -> point averaging, surface vertices, triangles, grid matrix --
the stage after the integration, Reconstruction/main.cxx:151-189) against the numpy restatement oracle/mc_oracle.py:
bit-identical float32 vertices, identical triangles.  Parity with VTK itself is unpinned (VTK is absent)."""
import numpy as np
import pytest

from oracle import mc_oracle as mc
from cudadepthmapintegration_b200 import synthetic as syn
from tests.scenes import Scene

pytestmark = pytest.mark.gpu


def run(ctx, cells, grid, value):
    ctx.initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, 0.1, 0.8, 0.03, 0.3, (8, 8))
    return ctx.contour(cells, value)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,rot", [(12, 0.0), ((13, 9, 11), 30.0)])
def test_smooth_field_matches_the_oracle(gpu_ctx, n, rot, dtype):
    grid = syn.make_grid(n, rotate_deg=rot)
    Nx, Ny, Nz = grid.n_cells
    ax = [grid.origin[a] + (np.arange(m) + 0.5) * grid.spacing[a] for a, m in enumerate((Nx, Ny, Nz))]
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    cells = ((1.0 - np.sqrt(X * X + 1.3 * Y * Y + 0.8 * Z * Z)) * 3 + 1.0).astype(dtype).reshape(-1)
    wv, wt = mc.contour(cells, grid.n_cells, grid.origin, grid.spacing, grid.matrix, 1.0)
    gv, gt = run(gpu_ctx, cells, grid, 1.0)
    assert len(wv) > 100 and len(wt) > 100
    assert gv.shape == wv.shape and np.array_equal(gv.view(np.uint32), wv.view(np.uint32))
    assert gt.shape == wt.shape and np.array_equal(gt, wt)


def test_random_field_with_ambiguous_faces_matches_the_oracle(gpu_ctx):
    rng = np.random.RandomState(11)
    grid = syn.make_grid((9, 10, 8), rotate_deg=10.0)
    cells = rng.uniform(-1, 1, size=grid.n_voxels)
    cells[rng.randint(0, cells.size, 50)] = 0.25            # values exactly on the iso level: `>= value` is inside
    wv, wt = mc.contour(cells, grid.n_cells, grid.origin, grid.spacing, grid.matrix, 0.25)
    gv, gt = run(gpu_ctx, cells, grid, 0.25)
    assert np.array_equal(gv.view(np.uint32), wv.view(np.uint32)) and np.array_equal(gt, wt)
    # no surface at all
    gv, gt = run(gpu_ctx, cells, grid, 100.0)
    assert gv.shape == (0, 3) and gt.shape == (0, 3)


def test_fused_volume_to_coloured_mesh_on_the_device(gpu_ctx, oracle):
    """integration -> contour of the context's own volume -> coloration of the surface vertices, all in device memory."""
    import torch
    s = Scene(40, 24, 96, 72, depth_noise=0.25)
    ctx = gpu_ctx
    ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing, s.rp.thick, s.rp.rho, s.rp.eta, s.rp.delta, (s.W, s.H))
    ctx.volume_begin(None, np.float64)
    ctx.volume_integrate_host(s.depths, s.best_cost, 0.14, s.K, s.RT)
    nv, nt = ctx.contour_device(None, np.float64, 1.0)
    vol = np.empty(s.grid.n_voxels)
    ctx.volume_end(vol)
    wv, wt = mc.contour(vol, s.grid.n_cells, s.grid.origin, s.grid.spacing, s.grid.matrix, 1.0)
    gv, gt = ctx.contour_get(nv, nt)
    assert nv == len(wv) > 500 and nt == len(wt)
    assert np.array_equal(gv.view(np.uint32), wv.view(np.uint32)) and np.array_equal(gt, wt)
    r = np.linalg.norm(gv.astype(np.float64), axis=1)
    assert 0.2 < r.min() and r.max() < 1.25                  # around the unit sphere (and the shells behind it)
    # colour the vertices where they lie (device pointers from the contour)
    pv, _, n1, _ = ctx.contour_device_ptr()
    assert n1 == nv
    cols = torch.from_numpy(s.colors).cuda()
    mean = torch.zeros((nv, 3), dtype=torch.uint8, device="cuda"); med = torch.zeros_like(mean)
    nb = torch.zeros(nv, dtype=torch.int32, device="cuda")
    ctx.colorize_device(nv, pv, np.float32, s.n_views, cols.data_ptr(), s.K, s.RT, s.W, s.H, mean.data_ptr(), med.data_ptr(), nb.data_ptr())
    ctx.synchronize()
    want = oracle.colorize(gv, s.colors, s.K, s.RT, s.W, s.H)
    assert np.array_equal(nb.cpu().numpy(), want[2]) and np.array_equal(med.cpu().numpy(), want[1]) and np.array_equal(mean.cpu().numpy(), want[0])


def test_surface_of_a_256_cube_is_closed(gpu_ctx):
    n = 256
    grid = syn.make_grid(n)
    g = np.linspace(-1.2 + 1.2 / n, 1.2 - 1.2 / n, n, dtype=np.float32)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    cells = ((1.0 - np.sqrt(X * X + Y * Y + Z * Z)) * 4 + 1.0).astype(np.float32).reshape(-1)
    v, t = run(gpu_ctx, cells, grid, 1.0)
    assert len(v) > 150000
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]).astype(np.int64)
    key = np.sort(e, axis=1)
    key = key[:, 0] * len(v) + key[:, 1]
    _, counts = np.unique(key, return_counts=True)
    assert (counts == 2).all()                               # watertight
    assert len(v) - len(counts) + len(t) == 2                # Euler characteristic of a sphere
    r = np.linalg.norm(v.astype(np.float64), axis=1)
    assert abs(r.mean() - 1.0) < 0.01
