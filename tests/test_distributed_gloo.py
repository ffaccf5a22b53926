"""N>1 host logic on CPU: world_size-2 gloo run of the view exchange + z-slab sharding + slab gather
(cudadepthmapintegration_b200/distributed.py, sharding.py).  The per-rank integrator here is the ORACLE
(this is a test of the plumbing, not of the kernel): the gathered volume must be bit-identical to the
single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, group, out_path):
    from cudadepthmapintegration_b200 import distributed as D, sharding
    from tests import _oracle
    from tests.scenes import Scene
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = _oracle.load_oracle()
        s = Scene((18, 11, 7), 7, 48, 36, rotate_deg=30.0)         # 7 views: ragged last group
        V = s.n_views
        nz = s.grid.n_cells[2]
        plane = s.grid.n_cells[0] * s.grid.n_cells[1]
        # each rank "loads" and filters only the views it owns
        mine = D.owned_views(V, group, rank, world)
        all_views = torch.full((V, s.H, s.W), float("nan"), dtype=torch.float64)
        for v in mine:
            all_views[v] = torch.from_numpy(orc.apply_depth_threshold(s.depths[v], s.best_cost[v], 0.14).reshape(s.H, s.W))
        k0, k1 = sharding.slab_range(nz, rank, world)
        vol = np.zeros(s.grid.n_voxels)
        for g0, g1 in D.view_groups(V, group, world):
            D.all_gather_group(dist, all_views, g0, g1, rank, world)
            assert not torch.isnan(all_views[g0:g1]).any()
            orc.tsdf_integrate(s.grid, s.rp, s.W, s.H, all_views[g0:g1].numpy(), None, 0.0, s.K[g0:g1], s.RT[g0:g1], vol, k0, k1)
        slab = torch.from_numpy(vol[k0 * plane:k1 * plane].copy())
        full = torch.zeros(s.grid.n_voxels, dtype=torch.float64) if rank == 0 else None
        D.gather_slabs(dist, slab, full, plane, nz, rank, world)
        if rank == 0:
            want = orc.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
            np.save(out_path, np.stack([full.numpy(), want]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("group", [2, 4])
def test_two_ranks_reproduce_the_single_process_volume(tmp_path, group):
    out = str(tmp_path / "vol.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, group, out), nprocs=2, join=True)
    got, want = np.load(out)
    assert np.count_nonzero(want) > 0
    assert np.array_equal(got, want)


def test_view_ownership_covers_every_view_once():
    from cudadepthmapintegration_b200 import distributed as D
    for V, G, world in [(1000, 128, 8), (7, 2, 2), (10, 40, 4), (5, 3, 3)]:
        seen = sorted(v for r in range(world) for v in D.owned_views(V, G, r, world))
        assert seen == list(range(V))
        for g0, g1 in D.view_groups(V, G, world):
            assert (g1 - g0) <= max(world, (G // world) * world)
