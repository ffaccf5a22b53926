"""Multi-GPU plumbing around the C ABI (one process per GPU, torch.distributed).

The hot loop has no collective: every voxel has one owner (a z-slab) and needs every view.  What is
exchanged is (1) the views -- each rank owns ("has loaded") a share of every GROUP of consecutive views and
the group is assembled on all ranks by an in-place all-gather, group after group, so that the exchange of
group g+1 overlaps the integration of group g -- and (2) once, the finished slabs, concatenated on rank 0
in VTK cell order.  Everything here works on CPU tensors with the gloo backend too (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

from . import sharding


def view_groups(n_views: int, group: int, world: int, ramp: bool = True):
    """[(g0, g1)] consecutive groups of views; every group size is a multiple of the world size.  With
    `ramp` the first groups are smaller (G/4, G/4, G/2), so that integration starts after a short first
    exchange instead of waiting for a full group, and so are the last ones (G/2, G/4, G/4): what is left
    to integrate once the last views have arrived is a quarter group, not a whole one."""
    g = max(world, (max(group, 1) // world) * world)

    def part(f):
        return max(world, (g // f // world) * world)
    head = [q for q in (part(4), part(4), part(2)) if q < g] if ramp and world > 1 else []
    tail = [q for q in (part(2), part(4), part(4)) if q < g] if ramp and world > 1 else []
    out, g0, i = [], 0, 0
    while g0 < n_views:
        if i < len(head):
            q = head[i]
        elif tail and n_views - g0 <= sum(tail):
            q = tail.pop(0)
        else:
            q = g
        out.append((g0, min(n_views, g0 + q)))
        g0 += q
        i += 1
    return out


def owned_range(g0: int, g1: int, rank: int, world: int):
    """Views [a, b) of group [g0, g1) that `rank` owns, and the per-rank share size."""
    per = (g1 - g0 + world - 1) // world
    a = min(g1, g0 + rank * per)
    return a, min(g1, a + per), per


def owned_views(n_views: int, group: int, rank: int, world: int):
    """All view indices `rank` owns, in the order its local buffer holds them."""
    out = []
    for g0, g1 in view_groups(n_views, group, world):
        a, b, _ = owned_range(g0, g1, rank, world)
        out.extend(range(a, b))
    return out


def all_gather_group(dist, all_views, g0: int, g1: int, rank: int, world: int):
    """Assemble group [g0, g1) of `all_views` (tensor [V, ...]) on every rank.  On entry each rank has
    written its own share all_views[a:b]; full groups use one in-place all-gather, a ragged last group
    falls back to one broadcast per owner."""
    a, b, per = owned_range(g0, g1, rank, world)
    if (g1 - g0) == per * world:
        try:
            dist.all_gather_into_tensor(all_views[g0:g1].reshape(-1), all_views[a:b].reshape(-1))
        except (RuntimeError, NotImplementedError):
            parts = [all_views[owned_range(g0, g1, r, world)[0]:owned_range(g0, g1, r, world)[1]] for r in range(world)]
            dist.all_gather(parts, all_views[a:b].clone())
    else:
        for r in range(world):
            ra, rb, _ = owned_range(g0, g1, r, world)
            if rb > ra:
                dist.broadcast(all_views[ra:rb], src=r)


def gather_slabs(dist, slab, full_volume, plane_cells: int, n_cells_z: int, rank: int, world: int):
    """Concatenate the ranks' slabs into `full_volume` on rank 0 (slabs may differ by one plane, so the
    transfer is point-to-point rather than a fixed-size gather)."""
    k0, k1 = sharding.slab_range(n_cells_z, rank, world)
    ops = []
    if rank == 0:
        full_volume[k0 * plane_cells:k1 * plane_cells].copy_(slab)
        for r in range(1, world):
            a, b = sharding.slab_range(n_cells_z, r, world)
            if b > a:
                ops.append(dist.P2POp(dist.irecv, full_volume[a * plane_cells:b * plane_cells], r))
    elif k1 > k0:
        ops.append(dist.P2POp(dist.isend, slab, 0))
    if ops:
        # one batch: the seven transfers into rank 0 run concurrently instead of one after the other
        for q in dist.batch_isend_irecv(ops):
            q.wait()
