"""Host-side logic that needs no GPU: sharding arithmetic, the synthetic scene, argument checks."""
import numpy as np
import pytest

from cudadepthmapintegration_b200 import sharding, synthetic as syn


@pytest.mark.parametrize("n,world", [(1024, 8), (1024, 3), (7, 8), (1, 2), (100, 1)])
def test_slab_ranges_partition_the_grid(n, world):
    r = [sharding.slab_range(n, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n
    for a, b in zip(r, r[1:]):
        assert a[1] == b[0]
    assert all(k1 >= k0 for k0, k1 in r)
    assert max(k1 - k0 for k0, k1 in r) - min(k1 - k0 for k0, k1 in r) <= 1


def test_point_and_view_ranges():
    assert sharding.point_range(10, 0, 4) == (0, 2)
    assert sharding.point_range(10, 3, 4) == (7, 10)
    assert sharding.view_range(1000, 7, 8) == (875, 1000)


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


@pytest.mark.parametrize("V,G,world", [(1000, 128, 8), (1000, 128, 2), (100, 16, 4), (7, 4, 1), (10, 128, 2), (64, 64, 8)])
def test_view_groups_cover_all_views_in_order_with_short_ends(V, G, world):
    """Groups are consecutive, cover [0, V) once, every size but the last is a multiple of the world size, and
    with more than one rank both ends are shorter than a full group (integration starts early, the tail is short)."""
    from cudadepthmapintegration_b200 import distributed as D
    groups = D.view_groups(V, G, world)
    assert groups[0][0] == 0 and groups[-1][1] == V
    for (a0, a1), (b0, b1) in zip(groups, groups[1:]):
        assert a1 == b0 and a1 > a0
    assert all((b - a) % world == 0 for a, b in groups[:-1])
    full = max(world, G // world * world)
    assert all(b - a <= full for a, b in groups)
    if world > 1 and V >= 4 * full:
        assert groups[0][1] - groups[0][0] < full and groups[-1][1] - groups[-1][0] < full
    # every view has exactly one owner, and owners hold their views in increasing order
    owners = [v for r in range(world) for v in D.owned_views(V, G, r, world)]
    assert sorted(owners) == list(range(V))
    for r in range(world):
        mine = D.owned_views(V, G, r, world)
        assert list(mine) == sorted(mine)


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
