"""N>1 host logic on CPU: a world_size-2 gloo run of the sharded job's LAYOUT -- which rank supplies which views
(dmi_shard_view_indices), in which order a group is assembled (one in-place all-gather per group), which z-layers a
rank owns (dmi_set_slab_layers / engine.layer_cell_ranges) and where they land in the gathered volume.  The per-rank
integrator here is the ORACLE and the transport is gloo (this is a test of the plumbing the library implements with
NCCL in csrc/dmi_shard.cu, not of the kernel): the assembled volume must be bit-identical to the single-process one."""
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


def _worker(rank, world, port, n_views, out_path):
    from cudadepthmapintegration_b200 import engine
    from tests import _oracle
    from tests.scenes import Scene
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = _oracle.load_oracle()
        s = Scene((18, 11, 70), n_views, 48, 36, rotate_deg=30.0)       # 70 planes: layers 0..2, the last one 6 planes thick
        V = s.n_views
        nz = s.grid.n_cells[2]
        plane = s.grid.n_cells[0] * s.grid.n_cells[1]
        # each rank "loads" and filters only the views the library says it owns, in that order
        mine = engine.shard_view_indices(V, world, rank)
        my_views = [torch.from_numpy(orc.apply_depth_threshold(s.depths[v], s.best_cost[v], 0.14).reshape(s.H, s.W)) for v in mine]
        layers = engine.layer_cell_ranges(nz, world, rank)
        vol = np.zeros(s.grid.n_voxels)
        used = 0
        for g0, g1 in engine.shard_groups(V, world):
            pg = -(-(g1 - g0) // world)
            # the group buffer holds world * pg views; rank r's segment starts r * pg views in (padding at the end of a short group)
            a = min(g1, g0 + rank * pg)
            b = min(g1, a + pg)
            seg = torch.full((pg, s.H, s.W), float("nan"), dtype=torch.float64)
            for q in range(b - a):
                assert mine[used] == a + q
                seg[q] = my_views[used]
                used += 1
            parts = [torch.empty_like(seg) for _ in range(world)]
            dist.all_gather(parts, seg)
            group = torch.cat(parts)[:g1 - g0]
            assert not torch.isnan(group).any()
            for k0, k1 in layers:
                orc.tsdf_integrate(s.grid, s.rp, s.W, s.H, group.numpy(), None, 0.0, s.K[g0:g1], s.RT[g0:g1], vol, k0, k1)
        assert used == len(mine)
        # packed layers -> their places in rank 0's whole-grid volume
        packed = torch.from_numpy(np.concatenate([vol[k0 * plane:k1 * plane] for k0, k1 in layers]) if layers else np.zeros(0))
        if rank == 0:
            full = np.zeros(s.grid.n_voxels)
            for r in range(world):
                rl = engine.layer_cell_ranges(nz, world, r)
                n = sum(k1 - k0 for k0, k1 in rl) * plane
                buf = packed if r == 0 else torch.empty(n, dtype=torch.float64)
                if r != 0 and n:
                    dist.recv(buf, src=r)
                o = 0
                for k0, k1 in rl:
                    full[k0 * plane:k1 * plane] = buf[o:o + (k1 - k0) * plane].numpy()
                    o += (k1 - k0) * plane
            want = orc.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
            np.save(out_path, np.stack([full, want]))
        elif packed.numel():
            dist.send(packed, dst=0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [7, 4])
def test_two_ranks_reproduce_the_single_process_volume(tmp_path, n_views):
    out = str(tmp_path / "vol.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_views, out), nprocs=2, join=True)
    got, want = np.load(out)
    assert np.count_nonzero(want) > 0
    assert np.array_equal(got, want)
