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
