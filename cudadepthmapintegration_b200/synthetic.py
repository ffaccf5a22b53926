"""Deterministic synthetic RGB-D scene: a unit sphere seen by pinhole cameras.

The reference ships no sample data (its only example command lines point at private folders,
Reconstruction/main.cxx:102-103), so every test and benchmark runs on this scene.  The generator
produces arrays in exactly the memory layouts the reference reads:

* depth map   ``double[H][W]``, bottom-up rows, camera-z of the surface hit, ``-1`` = invalid
              (the "Depths" array of the .vti, CudaReconstruction.cu:141-149,202,207)
* best cost   ``double[H][W]`` ("Best Cost Values", ReconstructionData.cxx:146)
* colour      ``uint8[H][W][3]`` bottom-up ("Color", ReconstructionData.cxx:95)
* K, RT       4x4 row-major doubles; K is the 3x3 intrinsics inside an identity 4x4
              (ReconstructionData.cxx:199-209), RT = [R | t] with last row 0 0 0 1 (Helper.h:161-165)

torch is used as the array library only (CPU in the tests, CUDA in the benchmark, same code);
pseudo-random fields come from an integer hash of (seed, view, pixel) so both devices agree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

DEFAULT_SEED = 20261017

_M31 = 0x7FFFFFFF


def _hash31(idx: torch.Tensor, salt: int) -> torch.Tensor:
    """int64 tensor -> uniform ints in [0, 2^31); wrap-around int64 arithmetic, same on CPU and CUDA."""
    x = idx * 6364136223846793005 + (salt * 1442695040888963407 + 1013904223) % (1 << 62)
    x = x ^ ((x >> 29) & 0x7FFFFFFFF)
    x = x * -4658895280553007687  # 0xBF58476D1CE4E5B9 as int64
    x = x ^ ((x >> 32) & 0xFFFFFFFF)
    x = x * -7723592293110705685  # 0x94D049BB133111EB as int64
    x = x ^ ((x >> 31) & 0x1FFFFFFFF)
    return (x >> 16) & _M31


def _uniform(idx: torch.Tensor, salt: int) -> torch.Tensor:
    return _hash31(idx, salt).to(torch.float64) / float(1 << 31)


@dataclass
class Grid:
    """Voxel grid as the reference's CLI builds it (Reconstruction/main.cxx:123-126, 345-359)."""
    n_cells: tuple          # (Nx, Ny, Nz) cells; vtkImageData point dims are n_cells + 1
    origin: np.ndarray      # double[3]
    spacing: np.ndarray     # double[3]
    matrix: np.ndarray      # double[16] row-major grid matrix (rows = gridVecX/Y/Z)

    @property
    def point_dims(self):
        return tuple(int(n) + 1 for n in self.n_cells)

    @property
    def n_voxels(self):
        return int(self.n_cells[0]) * int(self.n_cells[1]) * int(self.n_cells[2])


@dataclass
class RayPotential:
    thick: float
    rho: float
    eta: float
    delta: float


def make_grid(n, rotate_deg: float = 0.0, extent: float = 2.4) -> Grid:
    """N^3 cells (or a tuple of three counts) covering [-extent/2, extent/2]^3 in grid-local axes.

    rotate_deg != 0 gives an orthonormal rotation about z as the grid matrix (the CLI requires the
    grid vectors to be orthogonal, Reconstruction/main.cxx:363-382).
    """
    cells = (n, n, n) if isinstance(n, int) else tuple(int(x) for x in n)
    spacing = np.array([extent / c for c in cells], dtype=np.float64)
    origin = np.array([-extent / 2] * 3, dtype=np.float64)
    m = np.eye(4, dtype=np.float64)
    if rotate_deg:
        a = math.radians(rotate_deg)
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = math.cos(a), -math.sin(a), math.sin(a), math.cos(a)
    return Grid(cells, origin, spacing, m.reshape(16).copy())


def make_ray_potential(grid: Grid) -> RayPotential:
    """Thick = 3 voxels, Delta = 10 voxels, Rho 0.8, Eta 0.03 (ratios of the example at main.cxx:102)."""
    s = float(grid.spacing.max())
    return RayPotential(thick=3.0 * s, rho=0.8, eta=0.03, delta=10.0 * s)


def make_cameras(n_views: int, width: int, height: int, seed: int = DEFAULT_SEED, radius: float = 3.0):
    """K[n,16], RT[n,16] (float64 numpy).  Centres on a Fibonacci sphere, looking at the origin, seeded roll."""
    K = np.zeros((n_views, 4, 4), dtype=np.float64)
    RT = np.zeros((n_views, 4, 4), dtype=np.float64)
    rng = np.random.RandomState(seed % (2**31 - 1))
    rolls = rng.uniform(0.0, 2.0 * math.pi, size=n_views)
    golden = math.pi * (3.0 - math.sqrt(5.0))
    f = 1.0 * height
    for i in range(n_views):
        z = 1.0 - (2.0 * i + 1.0) / n_views
        r = math.sqrt(max(0.0, 1.0 - z * z))
        phi = i * golden
        c = radius * np.array([r * math.cos(phi), r * math.sin(phi), z])
        fwd = -c / np.linalg.norm(c)
        up0 = np.array([0.0, 0.0, 1.0]) if abs(fwd[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
        right = np.cross(fwd, up0)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        cr, sr = math.cos(rolls[i]), math.sin(rolls[i])
        right, down = cr * right + sr * down, -sr * right + cr * down
        R = np.stack([right, down, fwd])
        RT[i, :3, :3] = R
        RT[i, :3, 3] = -R @ c
        RT[i, 3, 3] = 1.0
        K[i] = np.eye(4)
        K[i, 0, 0] = f
        K[i, 1, 1] = f
        K[i, 0, 2] = width / 2.0
        K[i, 1, 2] = height / 2.0
    return K.reshape(n_views, 16).copy(), RT.reshape(n_views, 16).copy()


def render_views(K: np.ndarray, RT: np.ndarray, width: int, height: int, *, seed: int = DEFAULT_SEED,
                 first_view: int = 0, device="cpu", depth_noise: float = 0.0, want_color: bool = True,
                 want_best_cost: bool = True, cost_model: str = "iid", scene: str = "sphere"):
    """Ray-cast the unit sphere for views K/RT (their global indices start at ``first_view``).

    ``scene``: "sphere" = the sphere seen from outside (first intersection; rays that miss give -1), "room" = the
    same sphere seen from INSIDE (cameras within it: the far intersection, every pixel has a depth).

    Returns (depths f64 [n,H,W], best_cost f64 [n,H,W] or None, colors u8 [n,H,W,3] or None),
    all with bottom-up rows.  ``cost_model``: "iid" = every pixel's best cost is an independent uniform in
    [0, 0.2) (salt-and-pepper holes after the 0.14 filter, ~30 % of the pixels); "coherent" = smooth value
    noise on a 48-pixel lattice in [0, 0.2) (the filter removes connected regions instead).
    """
    n = K.shape[0]
    dev = torch.device(device)
    Kt = torch.from_numpy(np.ascontiguousarray(K)).to(dev).view(n, 4, 4)
    RTt = torch.from_numpy(np.ascontiguousarray(RT)).to(dev).view(n, 4, 4)
    W, H = width, height
    px = torch.arange(W, device=dev, dtype=torch.float64).view(1, 1, W)
    # storage row r holds image row py = H-1-r
    py = (H - 1 - torch.arange(H, device=dev, dtype=torch.float64)).view(1, H, 1)
    fx = Kt[:, 0, 0].view(n, 1, 1); fy = Kt[:, 1, 1].view(n, 1, 1)
    cx = Kt[:, 0, 2].view(n, 1, 1); cy = Kt[:, 1, 2].view(n, 1, 1)
    dx = ((px - cx) / fx).expand(n, H, W)
    dy = ((py - cy) / fy).expand(n, H, W)
    R = RTt[:, :3, :3]
    t = RTt[:, :3, 3]
    # camera centre, world: -R^T t, spelled out element by element so that a view's maps do not depend on how
    # many views are rendered in one call (a batched matmul may pick another kernel, and rounding, for n == 1)
    C = -(R[:, 0, :] * t[:, 0:1] + R[:, 1, :] * t[:, 1:2] + R[:, 2, :] * t[:, 2:3])
    # world direction = R^T (dx, dy, 1)
    dw = [R[:, 0, a].view(n, 1, 1) * dx + R[:, 1, a].view(n, 1, 1) * dy + R[:, 2, a].view(n, 1, 1) for a in range(3)]
    a = dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]
    b = sum(dw[k] * C[:, k].view(n, 1, 1) for k in range(3))
    c = (C * C).sum(dim=1).view(n, 1, 1) - 1.0
    disc = b * b - a * c
    hit = disc > 0
    root = torch.sqrt(torch.clamp(disc, min=0.0))
    if scene == "sphere":
        s = (-b - root) / a                                   # camera-z of the hit (ray dir has z = 1)
    elif scene == "room":
        s = (-b + root) / a                                   # the wall behind the origin, seen from inside
    else:
        raise ValueError("scene must be 'sphere' or 'room'")
    hit = hit & (s > 0)
    vidx = torch.arange(first_view, first_view + n, device=dev, dtype=torch.int64).view(n, 1, 1)
    pix = (vidx * (H * W) + torch.arange(H * W, device=dev, dtype=torch.int64).view(1, H, W))
    depth = s
    if depth_noise > 0.0:
        g = (_uniform(pix, seed + 11) + _uniform(pix, seed + 12) + _uniform(pix, seed + 13)
             + _uniform(pix, seed + 14) - 2.0) * math.sqrt(3.0)      # ~N(0,1)
        depth = depth + depth_noise * g
    depths = torch.where(hit, depth, torch.full_like(depth, -1.0)).contiguous()
    best = None
    if want_best_cost and cost_model == "iid":
        best = (_uniform(pix, seed + 1) * 0.2).contiguous()
    elif want_best_cost and cost_model == "coherent":
        L = 48
        nx, ny = W // L + 2, H // L + 2
        gx = torch.arange(W, device=dev, dtype=torch.float64).view(1, 1, W) / L
        gy = torch.arange(H, device=dev, dtype=torch.float64).view(1, H, 1) / L
        ix, iy = gx.floor().to(torch.int64), gy.floor().to(torch.int64)
        wx, wy = gx - ix, gy - iy
        wx, wy = wx * wx * (3.0 - 2.0 * wx), wy * wy * (3.0 - 2.0 * wy)

        def node(ax, ay):
            return _uniform(vidx * (nx * ny) + ay * nx + ax, seed + 1)
        best = ((node(ix, iy) * (1 - wx) + node(ix + 1, iy) * wx) * (1 - wy)
                + (node(ix, iy + 1) * (1 - wx) + node(ix + 1, iy + 1) * wx) * wy) * 0.2
        best = best.expand(n, H, W).contiguous()
    elif want_best_cost:
        raise ValueError("cost_model must be 'iid' or 'coherent'")
    colors = None
    if want_color:
        chans = []
        bg = [(_hash31(torch.tensor([seed + 100 + k], dtype=torch.int64), 7).item() % 256) for k in range(3)]
        for k in range(3):
            nrm = C[:, k].view(n, 1, 1) + s * dw[k]              # hit point = outward normal (unit sphere)
            base = torch.floor(127.5 * (nrm + 1.0))
            noise = torch.floor(_uniform(pix, seed + 20 + k) * 17.0) - 8.0
            val = torch.clamp(base + noise, 0.0, 255.0)
            val = torch.where(hit, val, torch.full_like(val, float(bg[k])))
            chans.append(val.to(torch.uint8))
        colors = torch.stack(chans, dim=-1).contiguous()
    return depths, best, colors


def fibonacci_sphere_points(n_points: int, radius: float = 1.0, dtype=np.float32) -> np.ndarray:
    """Mesh stand-in for coloration config 1: P points on the sphere, vtkPoints-style float32 xyz."""
    i = np.arange(n_points, dtype=np.float64)
    z = 1.0 - (2.0 * i + 1.0) / n_points
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = i * (math.pi * (3.0 - math.sqrt(5.0)))
    pts = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1) * radius
    return np.ascontiguousarray(pts.astype(dtype))
