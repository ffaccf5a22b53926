"""Host-side logic that needs no GPU: sharding arithmetic, the synthetic scene, argument checks."""
import numpy as np
import pytest

from cudadepthmapintegration_b200 import engine, synthetic as syn


@pytest.mark.parametrize("n,world", [(1024, 8), (1024, 3), (100, 8), (33, 2), (7, 8), (1, 2), (100, 1)])
def test_z_layers_partition_the_grid(n, world):
    """dmi_set_slab_layers / dmi_shard_initialize: layers of 32 cells dealt round-robin; every plane has one owner and
    an owner's layers come in increasing order (its packed volume holds them in that order)."""
    owner = np.full(n, -1)
    for r in range(world):
        ranges = engine.layer_cell_ranges(n, world, r)
        assert ranges == sorted(ranges)
        for q, (k0, k1) in enumerate(ranges):
            assert k0 == (q * world + r) * 32 and 0 < k1 - k0 <= 32
            assert (owner[k0:k1] == -1).all()
            owner[k0:k1] = r
    assert (owner >= 0).all()


@pytest.mark.parametrize("n,world", [(10, 4), (1000, 8), (7, 8), (0, 3), (10_000_000, 8)])
def test_contiguous_ranges_of_points_and_colour_images(n, world):
    """dmi_shard_range (coloration: MeshColoration.cxx:140-192 has independent points): contiguous, ordered, complete."""
    pos = 0
    for r in range(world):
        first, count = engine.shard_range(n, world, r)
        assert first == min(pos, n) and count >= 0
        pos = first + count
    assert pos == n
    assert engine.shard_range(10, 4, 0) == (0, 3) and engine.shard_range(10, 4, 3) == (9, 1)


def test_grid_and_potential_follow_survey():
    g = syn.make_grid(128)
    assert g.point_dims == (129, 129, 129) and g.n_voxels == 128 ** 3
    assert np.allclose(g.spacing, 2.4 / 128) and np.allclose(g.origin, -1.2)
    rp = syn.make_ray_potential(g)
    assert rp.thick == pytest.approx(3 * 2.4 / 128) and rp.delta == pytest.approx(10 * 2.4 / 128)
    assert rp.delta > rp.thick and 0 < rp.eta < 1            # the CLI's checks, Reconstruction/main.cxx:270-271
    gr = syn.make_grid(8, rotate_deg=30.0).matrix.reshape(4, 4)
    assert np.allclose(gr[:3, :3] @ gr[:3, :3].T, np.eye(3))  # orthogonal grid vectors, main.cxx:363-382


def test_cameras_and_depth_known_answers():
    W, H = 64, 48
    K, RT = syn.make_cameras(7, W, H)
    rt = RT.reshape(-1, 4, 4)
    assert np.allclose(rt[:, 3], [0, 0, 0, 1])
    for m in rt:
        R = m[:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
        C = -R.T @ m[:3, 3]
        assert np.linalg.norm(C) == pytest.approx(3.0)
        assert np.allclose(R @ (-C / 3.0), [0, 0, 1], atol=1e-12)     # looks at the origin
    d, b, c = syn.render_views(K, RT, W, H)
    d = d.numpy()
    # the principal ray hits the unit sphere at camera-z = 3 - 1
    # pixel (W/2, H/2) -> storage row H-1-H/2
    assert d[:, H - 1 - H // 2, W // 2] == pytest.approx(2.0, abs=1e-12)
    assert (d[:, 0, 0] == -1).all()                                   # corners miss
    assert ((b.numpy() >= 0) & (b.numpy() < 0.2)).all()
    assert c.shape == (7, H, W, 3) and c.dtype.is_floating_point is False


def test_render_is_device_independent_in_its_random_fields():
    K, RT = syn.make_cameras(3, 32, 24)
    _, b1, c1 = syn.render_views(K, RT, 32, 24, first_view=0)
    _, b2, c2 = syn.render_views(K[1:], RT[1:], 32, 24, first_view=1)
    assert np.array_equal(b1.numpy()[1:], b2.numpy())                 # hash depends on the global view index
    assert np.array_equal(c1.numpy()[1:], c2.numpy())


@pytest.mark.parametrize("V,world", [(1000, 8), (1000, 2), (1000, 1), (1000, 3), (100, 4), (7, 2), (10, 3), (5, 8), (129, 4), (128, 8), (513, 8)])
def test_view_ownership_covers_all_views_in_order(V, world):
    """dmi_shard_group_starts / dmi_shard_view_indices: views go in consecutive groups (at most (128 / world) * world views,
    shorter ones at both ends of a long list); inside a group rank r owns one contiguous share of ceil(n / world) views, so
    that an in-place all-gather assembles the group in list order.  Every view has exactly one owner and owners hold their
    views in increasing order."""
    per = max(1, 128 // world)
    groups = engine.shard_groups(V, world)
    assert groups[0][0] == 0 and groups[-1][1] == V
    assert all(a1 == b0 and a1 > a0 for (a0, a1), (b0, b1) in zip(groups, groups[1:]))
    assert all(g1 - g0 <= (per + 1) * world for g0, g1 in groups)
    if world > 1 and V >= 3 * per * world:
        assert groups[0][1] - groups[0][0] < per * world and groups[-1][1] - groups[-1][0] < per * world     # short ends
    owned = [engine.shard_view_indices(V, world, r) for r in range(world)]
    assert sorted(np.concatenate(owned).tolist()) == list(range(V))
    for r, mine in enumerate(owned):
        assert list(mine) == sorted(mine)
        for g0, g1 in groups:
            pg = -(-(g1 - g0) // world)
            lo = min(g1, g0 + r * pg)
            assert [v for v in mine if g0 <= v < g1] == list(range(lo, min(g1, lo + pg)))


def test_render_views_do_not_depend_on_the_batch():
    """A view's synthetic maps are the same whether it is rendered alone or in a batch (multi-GPU ranks render the
    runs of views they own; a batch-dependent rounding would make their volumes differ from the single-GPU one)."""
    import numpy as np
    from cudadepthmapintegration_b200 import synthetic as syn
    K, RT = syn.make_cameras(9, 96, 72)
    d, c, col = syn.render_views(K, RT, 96, 72, depth_noise=0.01)
    for i in (0, 4, 8):
        d1, c1, col1 = syn.render_views(K[i:i + 1], RT[i:i + 1], 96, 72, first_view=i, depth_noise=0.01)
        assert np.array_equal(d1[0].numpy(), d[i].numpy()) and np.array_equal(c1[0].numpy(), c[i].numpy())
        assert np.array_equal(col1[0].numpy(), col[i].numpy())


def test_python_readers_round_trip_every_vti_layout(tmp_path):
    """dataset_io.load_dataset (what SetFilePathVTI / SetFilePathKRTD and MeshColoration.from_files use) reads back,
    bit for bit, datasets written in every layout VTK's XML writers produce."""
    from cudadepthmapintegration_b200 import dataset_io
    from tests.scenes import Scene
    from tests.test_host_cli import VTI_LAYOUTS
    s = Scene(8, len(VTI_LAYOUTS), 72, 50, rotate_deg=10.0)
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT, vti_options=VTI_LAYOUTS)
    d, c, col, K, RT = dataset_io.load_dataset(str(tmp_path / "vtiList.txt"), str(tmp_path / "kList.txt"), need_color=True)
    assert np.array_equal(d, s.depths) and np.array_equal(c, s.best_cost) and np.array_equal(col, s.colors)
    assert np.array_equal(K, s.K) and np.array_equal(RT, s.RT)
    # fewer .krtd than .vti entries is the reference's error case
    (tmp_path / "short.txt").write_text("view_0000.krtd\n")
    with pytest.raises(ValueError):
        dataset_io.load_dataset(str(tmp_path / "vtiList.txt"), str(tmp_path / "short.txt"))
    # list-file tokenisation: the last blank-separated token, relative to the list's directory
    (tmp_path / "odd.txt").write_text("0 a.vti\nb.vti \n\nc d  \n")
    got = dataset_io.extract_all_file_path(str(tmp_path / "odd.txt"))
    assert got == [str(tmp_path) + "/a.vti", str(tmp_path) + "/b.vti", str(tmp_path) + "/"]
