"""Small seeded scenes shared by the CPU and GPU tests."""
from __future__ import annotations

import numpy as np

from cudadepthmapintegration_b200 import synthetic as syn


class Scene:
    def __init__(self, n, n_views, W, H, rotate_deg=0.0, seed=syn.DEFAULT_SEED, depth_noise=0.0, radius=3.0, cost_model="iid"):
        self.grid = syn.make_grid(n, rotate_deg=rotate_deg)
        self.rp = syn.make_ray_potential(self.grid)
        self.W, self.H = W, H
        self.K, self.RT = syn.make_cameras(n_views, W, H, seed=seed, radius=radius)
        d, b, c = syn.render_views(self.K, self.RT, W, H, seed=seed,
                                   depth_noise=depth_noise * float(self.grid.spacing.max()), cost_model=cost_model)
        self.depths = d.numpy()
        self.best_cost = b.numpy()
        self.colors = c.numpy()
        self.n_views = n_views

    def zeros(self, dtype=np.float64):
        return np.zeros(self.grid.n_voxels, dtype=dtype)
