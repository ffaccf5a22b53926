"""GPU parity of the integration path at BASELINE.json's own sizes against the REFERENCE'S OWN kernel
(Reconstruction/CudaReconstruction.cu:47-212, compiled unmodified by nvcc for sm_100a with -fmad=false =
the numerics of the shipped -G build: oracle/_ref/libref_tsdf_cuda_nofma.so) running on the same B200:

  configs[2]  256^3 cells x 100 views 1920x1080, whole grid
  configs[3]  512^3 cells x  96 views 1920x1080, whole grid
  configs[4]  1024^3 cells: 64 views spread over the 1000-camera list, 64-plane z-slabs at the bottom, in the
              middle and at the top of the grid (the reference kernel always runs the whole grid)

plus a general (non-pinhole) 4x4 K, which selects the PINHOLE=false instantiations of the fast kernel.
Bars: exact kernel BIT-equal; certified fast path: identical support, 1e-5 relative / 1e-6 absolute.
The certification margins of the fast path scale with N, W and H: this is where they are largest."""
import numpy as np
import pytest

from cudadepthmapintegration_b200 import _lib, synthetic as syn
from tests import _oracle
from tests.test_tsdf_parity_gpu import assert_close, run_gpu, RTOL, ATOL

pytestmark = pytest.mark.gpu


class DeviceScene:
    """Seeded scene rendered on the GPU (the CPU renderer would take minutes at 1080p), then held in host memory."""

    def __init__(self, n, view_ids, total_views, W, H):
        import torch
        self.grid = syn.make_grid(n)
        self.rp = syn.make_ray_potential(self.grid)
        K, RT = syn.make_cameras(total_views, W, H)
        self.K, self.RT = np.ascontiguousarray(K[view_ids]), np.ascontiguousarray(RT[view_ids])
        self.W, self.H = W, H
        dev = torch.device("cuda", 0)
        nv = len(view_ids)
        self.depths = np.empty((nv, H, W)); self.best_cost = np.empty((nv, H, W)); self.filtered = np.empty((nv, H, W))
        for q, v in enumerate(view_ids):
            d, c, _ = syn.render_views(K[v:v + 1], RT[v:v + 1], W, H, first_view=int(v), device=dev,
                                       depth_noise=0.25 * float(self.grid.spacing.max()), want_color=False)
            self.depths[q] = d[0].cpu().numpy(); self.best_cost[q] = c[0].cpu().numpy()
            # ReconstructionData::ApplyDepthThresholdFilter (ReconstructionData.cxx:159-166): the reference filters on the host
            self.filtered[q] = torch.where(c[0] > 0.14, torch.full_like(d[0], -1.0), d[0]).cpu().numpy()

    def zeros(self, dtype=np.float64):
        return np.zeros(self.grid.n_voxels, dtype=dtype)


def reference_volume(s):
    ref = _oracle.load_ref_cuda(nofma=True)
    if ref is None:
        pytest.skip("oracle/_ref/libref_tsdf_cuda_nofma.so not present (built where /root/reference exists)")
    want, _, _ = ref.run(s.grid, s.rp, s.W, s.H, s.filtered, s.K, s.RT, s.zeros())
    assert np.count_nonzero(want) > 0
    return want


def check_fast(got, want):
    assert np.array_equal(got != 0, want != 0)
    assert_close(got, want)


@pytest.mark.parametrize("n,nv", [(256, 100), (512, 96)])
def test_configs_3_and_4_whole_grid_vs_reference_kernel(gpu_ctx, n, nv):
    s = DeviceScene(n, np.arange(nv), nv, 1920, 1080)
    want = reference_volume(s)
    got_exact = run_gpu(gpu_ctx, s, np.float64, kernel=_lib.DMI_TSDF_KERNEL_EXACT)
    assert np.array_equal(got_exact.view(np.uint64), want.view(np.uint64))
    del got_exact
    got = run_gpu(gpu_ctx, s, np.float64)
    check_fast(got, want)


def test_config5_slabs_vs_reference_kernel(gpu_ctx):
    n, total = 1024, 1000
    ids = (np.arange(64) * total) // 64 + 7                 # 64 views spread over the 1000-camera list
    s = DeviceScene(n, ids, total, 1920, 1080)
    want = reference_volume(s).reshape(n, n * n)
    for k0 in (0, 480, 960):                                 # bottom, middle, top
        got = run_gpu(gpu_ctx, s, np.float64, slab=(k0, k0 + 64))
        w = want[k0:k0 + 64].reshape(-1)
        assert np.count_nonzero(w) > 0
        check_fast(got, w)
        got_exact = run_gpu(gpu_ctx, s, np.float64, kernel=_lib.DMI_TSDF_KERNEL_EXACT, slab=(k0, k0 + 64))
        assert np.array_equal(got_exact.view(np.uint64), w.view(np.uint64))


def general_k4(K, seed):
    """Any 4x4 K: skew, a projective third row and a non-zero fourth column (transformFrom4Matrix uses all of
    rows 0-2, CudaReconstruction.cu:88-93,176).  h.z stays positive in front of the cameras."""
    rng = np.random.RandomState(seed)
    K = K.reshape(-1, 4, 4).copy()
    for k in K:
        k[0, 1] = rng.uniform(-3, 3)
        k[1, 1] *= rng.uniform(0.9, 1.1)
        k[2, 0] = rng.uniform(-2e-2, 2e-2)
        k[2, 1] = rng.uniform(-2e-2, 2e-2)
        k[2, 2] = rng.uniform(0.9, 1.1)
        k[0, 3] = rng.uniform(-5, 5)
        k[1, 3] = rng.uniform(-5, 5)
        k[2, 3] = rng.uniform(-0.05, 0.05)
    return K.reshape(-1, 16)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
def test_general_k_non_pinhole(gpu_ctx, oracle, kernel, dtype):
    from tests.scenes import Scene
    s = Scene(96, 8, 320, 240, rotate_deg=30.0, depth_noise=0.25)
    s.K = general_k4(s.K, 17)
    assert not np.array_equal(s.K.reshape(-1, 4, 4)[:, 2], np.tile([0.0, 0.0, 1.0, 0.0], (8, 1)))
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros(dtype))
    assert np.count_nonzero(want) > 1000
    ref = _oracle.load_ref_cuda(nofma=True)
    if ref is not None:                                      # and the reference kernel itself agrees with the oracle here
        filtered = oracle.apply_depth_threshold(s.depths, s.best_cost, 0.14)
        w2, _, _ = ref.run(s.grid, s.rp, s.W, s.H, filtered, s.K, s.RT, s.zeros(dtype))
        assert np.array_equal(w2.view(np.uint8), want.view(np.uint8))
    got = run_gpu(gpu_ctx, s, dtype, kernel=kernel)
    if kernel == _lib.DMI_TSDF_KERNEL_EXACT:
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    else:
        assert np.array_equal(got != 0, want != 0)
        if dtype == np.float32:
            err = np.abs(got.astype(np.float64) - want.astype(np.float64))
            assert (err <= 1e-5 + RTOL * np.abs(want)).all()          # float32 volume: one float ulp per added view
        else:
            assert_close(got, want)


def test_mixed_pinhole_and_general_views_in_one_call(gpu_ctx, oracle):
    """The pinhole flag is per launch (chunk of <= 64 views): a list that mixes both kinds must still match."""
    from tests.scenes import Scene
    s = Scene(64, 6, 160, 120, depth_noise=0.25)
    K = s.K.copy()
    K[[1, 4]] = general_k4(s.K[[1, 4]], 5)
    s.K = K
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    got = run_gpu(gpu_ctx, s, np.float64)
    assert np.array_equal(got != 0, want != 0)
    assert_close(got, want)


@pytest.mark.parametrize("delta_scale", [0.5, -1.0])
def test_delta_below_thick_and_negative_delta(gpu_ctx, oracle, delta_scale):
    """Delta < Thick and Delta < 0 are outside the regime the fast path is proven for (the reference CLI rejects
    them, main.cxx:270-271, the filter does not): the library must fall back to the exact kernel and still match."""
    from tests.scenes import Scene
    s = Scene(48, 5, 120, 90, depth_noise=0.25)
    s.rp.delta = delta_scale * s.rp.thick
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    got = run_gpu(gpu_ctx, s, np.float64)
    if delta_scale < 0:
        assert np.array_equal(got, want)                     # exact kernel: bit-equal
    else:
        assert np.array_equal(got != 0, want != 0)
        assert_close(got, want)
