"""GPU parity of the coloration path through the C ABI against the oracle: NbProjectedDepthMap and
MedianColoration bit-exact, MeanColoration within +-1 (BASELINE.json) -- in fact exact, all integer."""
import numpy as np
import pytest

from cudadepthmapintegration_b200 import synthetic as syn
from tests.scenes import Scene

pytestmark = pytest.mark.gpu


def check(got, want):
    gmean, gmed, gnb = got
    wmean, wmed, wnb = want
    assert np.array_equal(gnb, wnb)
    assert np.array_equal(gmed, wmed)
    assert np.abs(gmean.astype(int) - wmean.astype(int)).max() <= 1
    assert np.array_equal(gmean, wmean)


def test_config1_40k_points_10_views(gpu_ctx, oracle):
    # BASELINE.json configs[0]: 40k-point sphere mesh x 10 views 640x480
    s = Scene(8, 10, 640, 480)
    pts = syn.fibonacci_sphere_points(40000)
    want = oracle.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    gpu_ctx.color_kernel_stats()                             # drain what earlier tests left
    got = gpu_ctx.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    assert want[2].max() == 10 and want[2].min() >= 1
    check(got, want)
    ms, launches = gpu_ctx.color_kernel_stats()
    assert launches == 1 and ms > 0


@pytest.mark.parametrize("n_views", [1, 2, 31, 32, 33, 100, 257])
def test_view_counts_and_even_odd_medians(gpu_ctx, oracle, n_views):
    s = Scene(8, n_views, 96, 72, seed=7 + n_views)
    rng = np.random.RandomState(n_views)
    pts = np.concatenate([syn.fibonacci_sphere_points(500),
                          rng.uniform(-1.5, 1.5, size=(500, 3)).astype(np.float32)])
    want = oracle.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    got = gpu_ctx.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    check(got, want)


def test_double_points_behind_camera_and_degenerate(gpu_ctx, oracle):
    s = Scene(8, 9, 64, 48, radius=1.0)                 # cameras ON the sphere: many points behind them
    rng = np.random.RandomState(3)
    pts = rng.uniform(-1.2, 1.2, size=(2000, 3))
    K4 = s.K.reshape(-1, 4, 4); RT4 = s.RT.reshape(-1, 4, 4)
    centres = np.array([-m[:3, :3].T @ m[:3, 3] for m in RT4])
    pts[:9] = centres                                   # exactly at a camera centre: 0/0 -> rejected (x86 INT_MIN)
    pts[9] = [np.inf, 0, 0]; pts[10] = [np.nan, 1, 1]; pts[11] = [1e308, 1e308, 1e308]
    want = oracle.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    got = gpu_ctx.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    check(got, want)
    assert (want[2][:9] <= 8).all()


def test_no_point_seen_keeps_zero_fill(gpu_ctx, oracle):
    # points no view sees keep 0/0/0 and count 0 (MeshColoration.cxx:116-118,124-126,132,174)
    W, H = 32, 24
    K = np.eye(4); K[0, 0] = K[1, 1] = 24.0; K[0, 2] = W / 2; K[1, 2] = H / 2
    K = np.repeat(K.reshape(1, 16), 3, axis=0); RT = np.repeat(np.eye(4).reshape(1, 16), 3, axis=0)
    colors = np.full((3, H, W, 3), 200, dtype=np.uint8)
    pts = np.zeros((70, 3), dtype=np.float32)
    pts[:, 0] = 50.0 + np.arange(70)                    # u = 24 * x + 16 >= 1216: outside every image
    pts[:, 2] = 1.0
    pts[64] = [0.0, 0.0, 1.0]                           # one seen point in the middle of the batch
    mean, med, nb = gpu_ctx.colorize(pts, colors, K, RT, W, H)
    unseen = np.arange(70) != 64
    assert (nb[unseen] == 0).all() and (mean[unseen] == 0).all() and (med[unseen] == 0).all()
    assert nb[64] == 3 and mean[64].tolist() == [200, 200, 200] and med[64].tolist() == [200, 200, 200]
    want = oracle.colorize(pts, colors, K, RT, W, H)
    check((mean, med, nb), want)


@pytest.mark.parametrize("name", ["sphere10", "even_odd_random_points", "double_points_behind_cameras", "general_k",
                                  "pixel_boundaries"])
def test_gpu_matches_reference_coloration_and_golden(gpu_ctx, name):
    """Against the reference's OWN MeshColoration.cxx (oracle/_ref/libref_coloration.so, prebuilt where /root/reference
    exists) and against the golden arrays it produced; a general 3x3 K, x.5 pixel boundaries and float64 points included."""
    import os
    from tests import _oracle, test_coloration_pinning as tcp
    pts, colors, K, RT, W, H = tcp.CASES[name]()
    got = gpu_ctx.colorize(pts, colors, K, RT, W, H)
    assert np.array_equal(tcp.pack(*got), np.load(tcp.GOLDEN)[name])
    ref = _oracle.load_ref_coloration()
    if ref is not None:
        check(got, ref.colorize(pts, colors, K, RT, W, H))


def test_mesh_coloration_class(oracle):
    from cudadepthmapintegration_b200 import MeshColoration
    s = Scene(8, 6, 80, 60)
    pts = syn.fibonacci_sphere_points(3000)
    mc = MeshColoration(pts, s.colors, s.K, s.RT)
    assert mc.ProcessColoration() is True
    out = mc.GetOutput()
    want = oracle.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    check((out["MeanColoration"], out["MedianColoration"], out["NbProjectedDepthMap"]), want)
    empty = MeshColoration(pts)
    assert empty.ProcessColoration() is False           # MeshColoration.cxx:102-106
    mc.close(); empty.close()
