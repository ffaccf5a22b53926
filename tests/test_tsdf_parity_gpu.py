"""GPU parity of the integration path, through the C ABI, against the oracle (and, where built, the
reference's own kernel recompiled for sm_100a).

Bars (BASELINE.json north_star): every discrete decision (pixel, validity, branch) identical; voxel
values within 1e-5 relative / 1e-6 absolute.  The exact kernel is held to BIT equality with the
oracle / the -fmad=false reference build."""
import numpy as np
import pytest

from cudadepthmapintegration_b200 import _lib
from tests import _oracle
from tests.scenes import Scene
from tests.test_oracle_pinning import CASES, make_case

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6      # north_star tolerance for TSDF voxel values


def run_gpu(ctx, s, dtype, best_cost=True, threshold=0.14, start=None, kernel=_lib.DMI_TSDF_KERNEL_AUTO, slab=None):
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, kernel)
    ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing,
                   s.rp.thick, s.rp.rho, s.rp.eta, s.rp.delta, (s.W, s.H))
    if slab is not None:
        ctx.set_slab(*slab)
    n = ctx.slab_cells
    out = np.zeros(n, dtype=dtype) if start is None else start.astype(dtype).copy()
    ctx.process_depth_maps(s.depths, s.best_cost if best_cost else None, threshold, s.K, s.RT, out)
    return out


def assert_close(got, want):
    got = got.astype(np.float64); want = want.astype(np.float64)
    err = np.abs(got - want)
    ok = err <= ATOL + RTOL * np.abs(want)
    assert ok.all(), f"{(~ok).sum()} voxels out of tolerance, max abs err {err.max():.3e}"


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
@pytest.mark.parametrize("name", sorted(CASES))
def test_small_cases_against_oracle(gpu_ctx, oracle, name, kernel):
    s, dtype = make_case(name)
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros(dtype))
    got = run_gpu(gpu_ctx, s, dtype, kernel=kernel)
    assert np.count_nonzero(want) > 0
    if kernel == _lib.DMI_TSDF_KERNEL_EXACT:
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    else:
        # identical decisions => identical set of touched voxels; values within tolerance
        assert np.array_equal(got != 0, want != 0)
        assert_close(got, want)


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
def test_config2_128cube_10_views_640x480(gpu_ctx, oracle, kernel):
    # BASELINE.json configs[1]
    s = Scene(128, 10, 640, 480, depth_noise=0.25)
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    got = run_gpu(gpu_ctx, s, np.float64, kernel=kernel)
    if kernel == _lib.DMI_TSDF_KERNEL_EXACT:
        assert np.array_equal(got, want)
    else:
        assert np.array_equal(got != 0, want != 0)
        assert_close(got, want)


def test_against_reference_kernel_recompiled(gpu_ctx, oracle):
    ref_nofma = _oracle.load_ref_cuda(nofma=True)
    ref_o3 = _oracle.load_ref_cuda(nofma=False)
    if ref_nofma is None or ref_o3 is None:
        pytest.skip("oracle/_ref CUDA builds not present")
    s = Scene(96, 8, 320, 240, rotate_deg=30.0, depth_noise=0.25)
    filtered = oracle.apply_depth_threshold(s.depths, s.best_cost, 0.14)
    want_nofma, _, _ = ref_nofma.run(s.grid, s.rp, s.W, s.H, filtered, s.K, s.RT, s.zeros())
    want_o3, _, _ = ref_o3.run(s.grid, s.rp, s.W, s.H, filtered, s.K, s.RT, s.zeros())
    orc = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    # the oracle IS the reference kernel without FMA contraction (the shipped -G build), bit for bit
    assert np.array_equal(orc, want_nofma)
    got_exact = run_gpu(gpu_ctx, s, np.float64, kernel=_lib.DMI_TSDF_KERNEL_EXACT)
    assert np.array_equal(got_exact, want_nofma)
    got = run_gpu(gpu_ctx, s, np.float64)
    assert_close(got, want_nofma)
    assert_close(got, want_o3)


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
def test_accumulates_onto_io_scalar_and_view_order(gpu_ctx, oracle, kernel):
    s = Scene(40, 6, 96, 72)
    start = np.linspace(-2, 2, s.grid.n_voxels)
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, None, 0.0, s.K, s.RT, start.copy())
    got = run_gpu(gpu_ctx, s, np.float64, best_cost=False, start=start, kernel=kernel)
    if kernel == _lib.DMI_TSDF_KERNEL_EXACT:
        assert np.array_equal(got, want)
    else:
        assert_close(got, want)
    # two calls of 3 views == one call of 6 views, bit for bit (same accumulation order per voxel)
    ctx = gpu_ctx
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, kernel)
    ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing,
                   s.rp.thick, s.rp.rho, s.rp.eta, s.rp.delta, (s.W, s.H))
    two = start.copy()
    ctx.process_depth_maps(s.depths[:3], None, 0.0, s.K[:3], s.RT[:3], two)
    ctx.process_depth_maps(s.depths[3:], None, 0.0, s.K[3:], s.RT[3:], two)
    assert np.array_equal(two, got)


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
def test_zero_and_negative_zero_io_scalar(gpu_ctx, oracle, kernel):
    """An all-zero io_scalar is not uploaded (the library scans it on the host); -0.0 is NOT zero bits:
    voxels no view touches must keep their sign, as in the reference."""
    s = Scene(40, 6, 96, 72)
    for fill in (0.0, -0.0):
        start = np.full(s.grid.n_voxels, fill)
        start[17] = 0.0 if fill == 0.0 else 0.25          # one odd element in the -0.0 case
        want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, None, 0.0, s.K, s.RT, start.copy())
        got = run_gpu(gpu_ctx, s, np.float64, best_cost=False, start=start, kernel=kernel)
        if kernel == _lib.DMI_TSDF_KERNEL_EXACT:
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        else:
            # the fast path proves "contributes 0" without performing the reference's `+= 0.0`, which would
            # turn a -0.0 into +0.0: values are equal, the sign of a zero may differ (certification)
            assert_close(got, want)
            nz = want != 0
            assert np.array_equal(np.signbit(got[nz]), np.signbit(want[nz]))


def test_large_io_scalar_with_one_hidden_nonzero_voxel(gpu_ctx, oracle):
    """>= 256 MB io_scalar: dmi_process_depth_maps integrates onto a zeroed device volume while a host scan
    verifies that io_scalar really is all zero; one non-zero voxel the quick probe does not see must send
    the call through the upload path (the pass is repeated) and give the reference's result."""
    s = Scene(324, 2, 96, 72)
    start = s.zeros()
    assert start.nbytes >= (256 << 20)
    start[12345677] = 0.5
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, start.copy())
    got = run_gpu(gpu_ctx, s, np.float64, start=start)
    assert got[12345677] == want[12345677] and want[12345677] != 0
    assert np.array_equal(got != 0, want != 0)
    assert_close(got, want)
    # and the all-zero case of the same size (speculation confirmed)
    want0 = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    got0 = run_gpu(gpu_ctx, s, np.float64)
    assert np.array_equal(got0 != 0, want0 != 0)
    assert_close(got0, want0)


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
def test_z_slabs_concatenate_bit_identically(gpu_ctx, kernel):
    s = Scene((37, 21, 19), 5, 80, 60, rotate_deg=30.0)
    whole = run_gpu(gpu_ctx, s, np.float64, kernel=kernel)
    nz = s.grid.n_cells[2]
    parts = [run_gpu(gpu_ctx, s, np.float64, kernel=kernel, slab=r) for r in [(0, 5), (5, 6), (6, 6), (6, nz)]]
    assert np.array_equal(np.concatenate(parts), whole)


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
def test_view_chunking_does_not_change_results(gpu_ctx, kernel):
    s = Scene(24, 70, 48, 36)          # more views than one launch chunk
    a = run_gpu(gpu_ctx, s, np.float64, kernel=kernel)
    gpu_ctx.set_option(_lib.DMI_OPT_VIEW_CHUNK, 7)
    try:
        b = run_gpu(gpu_ctx, s, np.float64, kernel=kernel)
    finally:
        gpu_ctx.set_option(_lib.DMI_OPT_VIEW_CHUNK, 0)
    assert np.array_equal(a, b)


def test_brick_culling_is_conservative(gpu_ctx, oracle):
    """Skipping (brick, view) pairs that cannot contribute must not change a single bit; checked with the
    tier counters that pairs really were culled (cameras at radius 3 see the 2.4-box well inside a 640x480
    image, whose background and thresholded tiles are invalid)."""
    s = Scene(96, 12, 640, 480, rotate_deg=30.0, depth_noise=0.25)
    ctx = gpu_ctx
    outs = []
    for cull in (1, 0):
        ctx.set_option(_lib.DMI_OPT_CULL, cull)
        ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 1)
        try:
            outs.append(run_gpu(ctx, s, np.float64))
            counters = ctx.tsdf_tier_counters()
        finally:
            ctx.set_option(_lib.DMI_OPT_CULL, 1)
            ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 0)
        if cull:
            assert counters["culled_brick_views"] > 0
        else:
            assert counters["culled_brick_views"] >= 0
    assert np.array_equal(outs[0], outs[1])
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    assert np.array_equal(outs[0] != 0, want != 0)
    assert_close(outs[0], want)


@pytest.mark.parametrize("best_cost", [True, False])
def test_free_space_bricks_are_settled_exactly(gpu_ctx, oracle, best_cost):
    """(brick, view) pairs whose voxels all lie farther than Delta in front of fully valid depth tiles get
    -Eta*Rho by one add per voxel, without projecting them: must not change a bit against the kernel with
    the brick tests disabled, and must happen (coherent best-cost maps / no filter leave fully valid tiles)."""
    # 1.5 pixels per voxel: a brick's footprint is a few 8-pixel tiles
    s = Scene(128, 12, 320, 240, rotate_deg=30.0, depth_noise=0.25, cost_model="coherent")
    ctx = gpu_ctx
    outs = []
    for cull in (1, 0):
        ctx.set_option(_lib.DMI_OPT_CULL, cull)
        ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 1)
        try:
            outs.append(run_gpu(ctx, s, np.float64, best_cost=best_cost))
            counters = ctx.tsdf_tier_counters()
        finally:
            ctx.set_option(_lib.DMI_OPT_CULL, 1)
            ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 0)
        if cull:
            assert counters["uniform_front"] > 0
        else:
            assert counters["uniform_front"] == 0
    assert np.array_equal(outs[0], outs[1])
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost if best_cost else None, 0.14, s.K, s.RT, s.zeros())
    assert np.array_equal(outs[0] != 0, want != 0)
    assert_close(outs[0], want)


def edge_scene():
    """Voxel exactly at the camera centre (0/0 -> NaN pixel), h.z == 0 (+-inf pixel), voxels behind the
    camera, projections exactly on x.5, depth == -1, an all-invalid view.  Everything is exact in binary."""
    from cudadepthmapintegration_b200 import synthetic as syn
    W, H = 16, 12
    grid = syn.Grid((8, 8, 8), np.array([-4.0, -4.0, -4.0]), np.array([1.0, 1.0, 1.0]), np.eye(4).reshape(16))
    rp = syn.RayPotential(thick=0.5, rho=0.75, eta=0.25, delta=2.0)
    K = np.eye(4); K[0, 0] = K[1, 1] = 4.0; K[0, 2] = 8.0; K[1, 2] = 6.0
    views_K, views_RT = [], []
    # camera centres exactly on voxel centres, axis-aligned
    for c in ([0.5, 0.5, 0.5], [0.5, 0.5, -3.5], [-1.5, 2.5, -0.5]):
        RT = np.eye(4); RT[:3, 3] = -np.array(c)
        views_K.append(K.reshape(16)); views_RT.append(RT.reshape(16))
    Ks = np.array(views_K); RTs = np.array(views_RT)
    rng = np.random.RandomState(5)
    depths = rng.uniform(0.5, 6.0, size=(3, H, W))
    depths[rng.uniform(size=depths.shape) < 0.3] = -1.0
    depths[2] = -1.0                                  # a view with no valid pixel at all
    depths[0, H - 1, 0] = 1.25                        # pixel (0,0): would be hit if NaN converted to 0
    return grid, rp, W, H, depths, Ks, RTs


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
def test_edge_cases_match_oracle(gpu_ctx, oracle, kernel):
    grid, rp, W, H, depths, Ks, RTs = edge_scene()
    want = oracle.tsdf_integrate(grid, rp, W, H, depths, None, 0.0, Ks, RTs, np.zeros(512))
    gpu_ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, kernel)
    gpu_ctx.initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, rp.thick, rp.rho, rp.eta, rp.delta, (W, H))
    got = np.zeros(512)
    gpu_ctx.process_depth_maps(depths, None, 0.0, Ks, RTs, got)
    assert np.count_nonzero(want) > 100
    assert np.array_equal(got, want)


def test_split_depth_edge_values(gpu_ctx, oracle):
    """The lossless split depth (float classification + int32 residual) on awkward depths: values next to -1,
    NaN, +-inf, beyond float range, tiny, negative, many-bit mantissas.  Integration from the split form must
    equal integration from the double maps, and both the oracle (same discrete decisions)."""
    import torch
    grid, rp, W, H, depths, Ks, RTs = edge_scene()
    depths = depths.copy()
    rng = np.random.RandomState(11)
    flat = depths.reshape(-1)
    special = [np.nextafter(-1.0, 0.0), np.nextafter(-1.0, -2.0), -1.0 - 2.0 ** -25, -1.0 + 2.0 ** -26, np.nan, np.inf, -np.inf,
               1e300, -1e300, 2.0 ** -70, 0.0, -0.0, 1.0 + 2.0 ** -52, 3.0 - 2.0 ** -51, 2.0 - 2.0 ** -53, np.pi, 1e-30, 5e-324]
    pos = rng.choice(2 * W * H, size=len(special), replace=False)      # views 0 and 1 (view 2 stays all-invalid)
    flat[pos] = special
    want = oracle.tsdf_integrate(grid, rp, W, H, depths, None, 0.0, Ks, RTs, np.zeros(512))
    ctx = gpu_ctx
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_AUTO)
    ctx.initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, rp.thick, rp.rho, rp.eta, rp.delta, (W, H))
    got = np.zeros(512)
    ctx.process_depth_maps(depths, None, 0.0, Ks, RTs, got)
    nv, npix = 3, W * H
    dev = torch.device("cuda", 0)
    d = torch.from_numpy(depths).to(dev)
    ncls, ntile = ctx.prepared_view_sizes()
    cls = torch.empty(nv * ncls + 1, dtype=torch.float32, device=dev)
    lo = torch.empty(nv * npix, dtype=torch.int32, device=dev)
    tiles = torch.empty(nv * ntile, dtype=torch.float32, device=dev)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ctx.prepare_views_device(nv, d.data_ptr(), None, 0.0, cls.data_ptr(), nv * npix, tiles.data_ptr(), d_lo=lo.data_ptr())
        ctx.volume_begin(None, np.float64)
        ctx.volume_integrate_prepared(nv, None, cls.data_ptr(), nv * npix, tiles.data_ptr(), Ks, RTs, d_lo=lo.data_ptr())
        split = np.empty(512)
        ctx.volume_end(split)
    finally:
        ctx.set_stream(None)
    assert np.array_equal(split, got, equal_nan=True)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok] != 0, want[ok] != 0)
    assert_close(got[ok], want[ok])
    # the split is exact wherever |d| >= 2^-64 or d == 0 (and keeps inf / NaN)
    hi32 = cls[:nv * npix].cpu().numpy()
    e = np.maximum(((hi32.view(np.int32) >> 23) & 0xff) - 127, -64)
    with np.errstate(invalid="ignore", over="ignore"):
        rebuilt = hi32.astype(np.float64) + lo.cpu().numpy().astype(np.float64) * np.exp2((e - 53).astype(np.float64))
    dn = depths.reshape(-1)
    exact = (hi32 != -1.0) & np.isfinite(dn) & ((np.abs(dn) >= 2.0 ** -64) | (dn == 0)) & (np.abs(dn) < 3e38)
    assert np.array_equal(rebuilt[exact], dn[exact])
    assert np.count_nonzero(exact) > 100


def test_edge_cases_reference_kernel(oracle):
    """Pins the oracle's device-conversion semantics (NaN, +-inf, saturation) on the reference's own
    kernel running on this GPU: both nvcc builds must agree with the oracle bit for bit."""
    grid, rp, W, H, depths, Ks, RTs = edge_scene()
    want = oracle.tsdf_integrate(grid, rp, W, H, depths, None, 0.0, Ks, RTs, np.zeros(512))
    for nofma in (True, False):
        ref = _oracle.load_ref_cuda(nofma=nofma)
        if ref is None:
            pytest.skip("oracle/_ref CUDA builds not present")
        got, _, _ = ref.run(grid, rp, W, H, depths, Ks, RTs, np.zeros(512))
        assert np.array_equal(got, want)


def test_threshold_filter_on_device(gpu_ctx, oracle):
    import torch
    rng = np.random.RandomState(1)
    d = rng.uniform(1, 3, size=10007); c = rng.uniform(0, 0.2, size=10007)
    want = oracle.apply_depth_threshold(d, c, 0.14)
    td = torch.from_numpy(d).cuda(); tc = torch.from_numpy(c).cuda()
    gpu_ctx.apply_depth_threshold_device(d.size, td.data_ptr(), tc.data_ptr(), 0.14)
    gpu_ctx.synchronize()
    assert np.array_equal(td.cpu().numpy(), want)
    # misaligned (8-byte offset) view
    td = torch.from_numpy(d).cuda(); tc = torch.from_numpy(c).cuda()
    gpu_ctx.apply_depth_threshold_device(d.size - 1, td.data_ptr() + 8, tc.data_ptr() + 8, 0.14)
    gpu_ctx.synchronize()
    assert np.array_equal(td.cpu().numpy()[1:], want[1:])


def test_device_resident_views_and_const_inputs(gpu_ctx, oracle):
    import torch
    s = Scene(32, 5, 64, 48)
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    ctx = gpu_ctx
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_EXACT)
    ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing,
                   s.rp.thick, s.rp.rho, s.rp.eta, s.rp.delta, (s.W, s.H))
    td = torch.from_numpy(s.depths).cuda(); tb = torch.from_numpy(s.best_cost).cuda()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ctx.volume_begin(None, np.float64)
        ctx.volume_integrate_device(s.n_views, td.data_ptr(), tb.data_ptr(), 0.14, s.K, s.RT)
        out = np.empty(s.grid.n_voxels)
        ctx.volume_end(out)
    finally:
        ctx.set_stream(None)
    assert np.array_equal(out, want)
    assert np.array_equal(td.cpu().numpy(), s.depths)       # the caller's depth maps are not modified
    ms, launches = ctx.tsdf_kernel_stats()
    assert launches >= 1 and ms > 0


def test_error_codes(gpu_ctx):
    from cudadepthmapintegration_b200 import DmiError
    s = Scene(8, 2, 16, 12)
    with pytest.raises(DmiError) as e:
        gpu_ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing, 0.0, 0.0, 0.1, 0.3, (16, 12))
    assert e.value.code == _lib.DMI_ERR_BAD_PARAMETERS          # vtkCudaReconstructionFilter.cxx:138-142
    gpu_ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing, 0.1, 0.8, 0.1, 0.3, (16, 12))
    with pytest.raises(DmiError) as e:
        gpu_ctx.process_depth_maps(np.zeros((0, 12, 16)), None, 0.0, np.zeros((0, 16)), np.zeros((0, 16)), np.zeros(512))
    assert e.value.code == _lib.DMI_ERR_NO_VIEWS                # CudaReconstruction.cu:308-312
    with pytest.raises(DmiError):
        gpu_ctx.set_slab(3, 2)
    # prepared views: the spare -1.0f slot is addressed by a 32-bit offset from every view of the call
    import torch
    gpu_ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_AUTO)
    gpu_ctx.volume_begin(None, np.float64)
    ncls, ntile = gpu_ctx.prepared_view_sizes()
    cls = torch.full((2 * ncls + 1,), -1.0, dtype=torch.float32, device="cuda")
    tiles = torch.zeros(2 * ntile, dtype=torch.float32, device="cuda")
    d = torch.zeros(2 * ncls, dtype=torch.float64, device="cuda")
    K = np.tile(np.eye(4).reshape(1, 16), (2, 1))
    for spare in (1 << 33, -(1 << 33)):
        with pytest.raises(DmiError) as e:
            gpu_ctx.volume_integrate_prepared(2, d.data_ptr(), cls.data_ptr(), spare, tiles.data_ptr(), K, K)
        assert e.value.code == _lib.DMI_ERR_INVALID_ARGUMENT


def test_filter_class_mirrors_reference_api(oracle):
    from cudadepthmapintegration_b200 import CudaReconstructionFilter
    s = Scene(20, 4, 48, 36)
    f = CudaReconstructionFilter()
    f.SetInputGrid(s.grid.point_dims, s.grid.origin, s.grid.spacing)
    f.SetGridMatrix(s.grid.matrix)
    f.SetViews(s.depths, s.best_cost, s.K, s.RT)
    assert f.Update() == 0                                      # Rho == Thick == 0 -> error path, returns 0
    f.SetRayPotentialThickness(s.rp.thick); f.SetRayPotentialRho(s.rp.rho)
    f.SetRayPotentialEta(s.rp.eta); f.SetRayPotentialDelta(s.rp.delta)
    f.SetThresholdBestCost(0.14)
    assert f.Update() == 1
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    got = f.GetOutput()
    assert got.shape == (20, 20, 20)
    assert_close(got.reshape(-1), want)
    assert f.GetExecutionTime() > 0
    f.close()


def test_filter_class_reads_the_reference_file_layout(tmp_path, oracle):
    """SetFilePathVTI / SetFilePathKRTD (vtkCudaReconstructionFilter.h:75-77): list files, .krtd and .vti (appended
    base64 + zlib, the XML writers' default) read by the VTK-free readers give the same volume as in-memory views."""
    from cudadepthmapintegration_b200 import CudaReconstructionFilter, dataset_io
    s = Scene(20, 4, 48, 36)
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT,
                             vti_options=[dict(encoding="base64", compress=True), dict(encoding="raw")])
    f = CudaReconstructionFilter()
    f.SetInputGrid(s.grid.point_dims, s.grid.origin, s.grid.spacing)
    f.SetGridMatrix(s.grid.matrix)
    f.SetFilePathVTI(str(tmp_path / "vtiList.txt")); f.SetFilePathKRTD(str(tmp_path / "kList.txt"))
    f.SetRayPotentialThickness(s.rp.thick); f.SetRayPotentialRho(s.rp.rho)
    f.SetRayPotentialEta(s.rp.eta); f.SetRayPotentialDelta(s.rp.delta)
    f.SetThresholdBestCost(0.14)
    assert f.Update() == 1
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    assert_close(f.GetOutput().reshape(-1), want)
    f.close()
    g = CudaReconstructionFilter()
    g.SetInputGrid(s.grid.point_dims, s.grid.origin, s.grid.spacing)
    g.SetFilePathVTI(str(tmp_path / "missing.txt")); g.SetFilePathKRTD(str(tmp_path / "kList.txt"))
    g.SetRayPotentialThickness(s.rp.thick); g.SetRayPotentialRho(s.rp.rho)
    assert g.Update() == 0                                      # unreadable list -> the reference's error path
    g.close()


# ---- BASELINE.json sizes: size-independent properties (the oracle would take hours here) --------------

def _device_scene(n, n_views, W, H):
    import torch
    from cudadepthmapintegration_b200 import synthetic as syn
    grid = syn.make_grid(n)
    rp = syn.make_ray_potential(grid)
    K, RT = syn.make_cameras(n_views, W, H)
    dev = torch.device("cuda", 0)
    ds, cs = [], []
    for v0 in range(0, n_views, 8):
        d, c, _ = syn.render_views(K[v0:v0 + 8], RT[v0:v0 + 8], W, H, first_view=v0, device=dev,
                                   depth_noise=0.25 * float(grid.spacing.max()), want_color=False)
        ds.append(d); cs.append(c)
    return grid, rp, K, RT, torch.cat(ds), torch.cat(cs)


def _integrate_device(ctx, grid, rp, K, RT, d, c, W, H, slab=None, splits=None, cull=1):
    import torch
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_AUTO)
    ctx.set_option(_lib.DMI_OPT_CULL, cull)
    ctx.initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, rp.thick, rp.rho, rp.eta, rp.delta, (W, H))
    if slab is not None:
        ctx.set_slab(*slab)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ctx.volume_begin(None, np.float64)
        n = K.shape[0]
        bounds = [0, n] if splits is None else [0] + list(splits) + [n]
        npix = W * H
        for a, b in zip(bounds, bounds[1:]):
            ctx.volume_integrate_device(b - a, d.data_ptr() + a * npix * 8, c.data_ptr() + a * npix * 8, 0.14, K[a:b], RT[a:b])
        ptr, nbytes = ctx.volume_device_ptr()

        class _H:
            pass
        h = _H()
        h.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        out = torch.as_tensor(h, device=d.device).clone()
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
        ctx.set_option(_lib.DMI_OPT_CULL, 1)
    return out


def test_full_size_properties_512_cube_1080p(gpu_ctx):
    """512^3 cells x 96 views of 1920x1080 (the grid and image size of BASELINE.json configs[3]):
    splitting the view list across calls, splitting the grid into z-slabs and switching the brick
    culling off must not change a single bit; the volume must also be physically sensible."""
    import torch
    n, nv, W, H = 512, 96, 1920, 1080
    grid, rp, K, RT, d, c = _device_scene(n, nv, W, H)
    whole = _integrate_device(gpu_ctx, grid, rp, K, RT, d, c, W, H)
    assert torch.equal(whole, _integrate_device(gpu_ctx, grid, rp, K, RT, d, c, W, H, splits=[7, 64, 65]))
    assert torch.equal(whole, _integrate_device(gpu_ctx, grid, rp, K, RT, d, c, W, H, cull=0))
    parts = [_integrate_device(gpu_ctx, grid, rp, K, RT, d, c, W, H, slab=s) for s in [(0, 100), (100, 101), (101, 317), (317, 512)]]
    assert torch.equal(whole, torch.cat(parts))
    vol = whole.view(n, n, n)
    assert torch.isfinite(vol).all()
    # far outside the sphere, seen in free space by many views: strictly negative; the centre of the sphere is
    # behind the surface for every view: exactly zero
    assert vol[n // 2, n // 2, 5].item() < 0 and vol[5, n // 2, n // 2].item() < 0
    assert vol[n // 2, n // 2, n // 2].item() == 0.0
    # just inside the surface along +x: positive (behind the surface within Thick)
    r_in = int(n / 2 + (1.0 - 1.5 * float(grid.spacing[0])) / float(grid.spacing[0]))
    assert vol[n // 2, n // 2, r_in].item() > 0


def test_prepared_views_path_is_bit_identical(gpu_ctx):
    """dmi_prepare_views_device + dmi_volume_integrate_prepared (the multi-GPU split: owners prepare, everyone
    integrates) == dmi_volume_integrate_device, bit for bit, also when the views are prepared in two pieces."""
    import torch
    n, nv, W, H = 96, 20, 320, 240
    grid, rp, K, RT, d, c = _device_scene(n, nv, W, H)
    whole = _integrate_device(gpu_ctx, grid, rp, K, RT, d, c, W, H)
    ctx = gpu_ctx
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_AUTO)
    ctx.initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, rp.thick, rp.rho, rp.eta, rp.delta, (W, H))
    ncls, ntile = ctx.prepared_view_sizes()
    assert ncls == W * H
    cls = torch.empty(nv * ncls + 1, dtype=torch.float32, device=d.device)
    tiles = torch.empty(nv * ntile, dtype=torch.float32, device=d.device)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        npix = W * H
        # two "owners": views [0,8) and [8,20); only the second call writes the spare slot
        ctx.prepare_views_device(8, d.data_ptr(), c.data_ptr(), 0.14, cls.data_ptr(), -1, tiles.data_ptr())
        ctx.prepare_views_device(12, d.data_ptr() + 8 * npix * 8, c.data_ptr() + 8 * npix * 8, 0.14,
                                 cls.data_ptr() + 8 * npix * 4, 12 * npix, tiles.data_ptr() + 8 * ntile * 4)
        ctx.volume_begin(None, np.float64)
        ctx.volume_integrate_prepared(nv, d.data_ptr(), cls.data_ptr(), nv * npix, tiles.data_ptr(), K, RT)
        out = np.empty(n ** 3)
        ctx.volume_end(out)
    finally:
        ctx.set_stream(None)
    assert cls[-1].item() == -1.0
    assert np.array_equal(out, whole.cpu().numpy())
    # the same with the LOSSLESS split depth (classification float + int32 residual) instead of the double maps
    lo = torch.empty(nv * npix, dtype=torch.int32, device=d.device)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ctx.prepare_views_device(nv, d.data_ptr(), c.data_ptr(), 0.14, cls.data_ptr(), nv * npix, tiles.data_ptr(), d_lo=lo.data_ptr())
        ctx.volume_begin(None, np.float64)
        ctx.volume_integrate_prepared(nv, None, cls.data_ptr(), nv * npix, tiles.data_ptr(), K, RT, d_lo=lo.data_ptr())
        out2 = np.empty(n ** 3)
        ctx.volume_end(out2)
    finally:
        ctx.set_stream(None)
    assert np.array_equal(out2, out)
    # and the split itself: hi + lo * 2^(e-53) == depth, bit for bit, on every valid pixel
    hi = cls[:nv * npix].cpu().numpy().astype(np.float64)
    e = np.maximum(((cls[:nv * npix].cpu().numpy().view(np.int32) >> 23) & 0xff) - 127, -64)
    rebuilt = hi + lo.cpu().numpy().astype(np.float64) * np.exp2((e - 53).astype(np.float64))
    dn = d.cpu().numpy().reshape(-1)
    valid = hi != -1.0
    assert valid.any() and np.array_equal(rebuilt[valid], dn[valid])
    # the device encoder against the host-side statement of the format (split_depth.py)
    from cudadepthmapintegration_b200 import split_depth
    hi_spec, lo_spec = split_depth.encode(dn, c.cpu().numpy().reshape(-1), 0.14)
    assert np.array_equal(cls[:nv * npix].cpu().numpy().view(np.int32), hi_spec.view(np.int32))
    assert np.array_equal(lo.cpu().numpy(), lo_spec)
