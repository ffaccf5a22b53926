"""Sharding over several GPUs (include/dmi_b200.h: dmi_set_slab_layers, dmi_comm_* / dmi_shard_*, dmi_group_*).
Every voxel and every mesh point has one owner (CudaReconstruction.cu:163,211; MeshColoration.cxx:140-192), so the
assembled results must be BIT-identical to the single-GPU ones.  Tests that need two devices skip on a one-GPU box."""
import numpy as np
import pytest

from cudadepthmapintegration_b200 import Context, Group, _lib, engine, synthetic as syn
from tests.scenes import Scene
from tests.test_tsdf_parity_gpu import assert_close, run_gpu

pytestmark = pytest.mark.gpu


def n_devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def assemble(parts, n_cells, world):
    """packed per-rank layers -> the whole grid, VTK cell order"""
    plane = n_cells[0] * n_cells[1]
    full = np.full(plane * n_cells[2], np.nan, dtype=parts[0].dtype)
    for r in range(world):
        o = 0
        for k0, k1 in engine.layer_cell_ranges(n_cells[2], world, r):
            full[k0 * plane:k1 * plane] = parts[r][o:o + (k1 - k0) * plane]
            o += (k1 - k0) * plane
        assert o == parts[r].size
    return full


@pytest.mark.parametrize("kernel", [_lib.DMI_TSDF_KERNEL_EXACT, _lib.DMI_TSDF_KERNEL_AUTO])
@pytest.mark.parametrize("world", [2, 3])
def test_z_layers_assemble_bit_identically_on_one_gpu(gpu_ctx, kernel, world):
    s = Scene((40, 37, 100), 9, 96, 72, rotate_deg=30.0, depth_noise=0.25)        # 100 planes: layers of 32, 32, 32, 4
    whole = run_gpu(gpu_ctx, s, np.float64, kernel=kernel)
    parts = []
    for r in range(world):
        gpu_ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing,
                           s.rp.thick, s.rp.rho, s.rp.eta, s.rp.delta, (s.W, s.H))
        gpu_ctx.set_slab_layers(32, r, world)
        assert gpu_ctx.slab_planes() == sum(k1 - k0 for k0, k1 in engine.layer_cell_ranges(100, world, r))
        out = np.zeros(gpu_ctx.slab_cells)
        gpu_ctx.process_depth_maps(s.depths, s.best_cost, 0.14, s.K, s.RT, out)
        parts.append(out)
    assert np.array_equal(assemble(parts, s.grid.n_cells, world).view(np.uint64), whole.view(np.uint64))


def test_spmd_entry_points_with_one_rank(oracle):
    """dmi_comm_init(world = 1) + dmi_shard_*: the SPMD path without a second GPU (no NCCL call is made)."""
    import torch
    s = Scene((33, 21, 70), 7, 96, 72, depth_noise=0.25)
    want = oracle.tsdf_integrate(s.grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, s.zeros())
    with Context(0) as ctx:
        ctx.comm_init(None, 0, 1)
        assert ctx.comm_info()[:2] == (0, 1)
        ctx.shard_initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing, s.rp.thick, s.rp.rho, s.rp.eta,
                             s.rp.delta, (s.W, s.H))
        assert list(engine.shard_view_indices(7, 1, 0)) == list(range(7))
        d = torch.from_numpy(s.depths).cuda(); c = torch.from_numpy(s.best_cost).cuda()
        full = torch.zeros(s.grid.n_voxels, dtype=torch.float64, device="cuda")
        ctx.volume_begin(None, np.float64)
        ctx.shard_integrate_device(7, d.data_ptr(), c.data_ptr(), 0.14, s.K, s.RT)
        ctx.shard_gather_volume_device(0, full.data_ptr())
        ctx.synchronize()
        got = full.cpu().numpy()
        assert np.array_equal(got != 0, want != 0)
        assert_close(got, want)
        # host-pointer variant, bit-identical
        ctx.volume_begin(None, np.float64)
        ctx.shard_integrate_host(7, s.depths, s.best_cost, 0.14, s.K, s.RT)
        out = np.empty(ctx.slab_cells)
        ctx.volume_end(out)
        assert np.array_equal(out, got)


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_group_object_matches_the_single_gpu_results(gpu_ctx, oracle, n_gpus):
    """dmi_group_*: one object, all GPUs; volume bit-identical to one context's, coloration bit-identical to the oracle's."""
    if n_devices() < n_gpus:
        pytest.skip(f"needs {n_gpus} CUDA devices")
    s = Scene((40, 37, 100), 150, 96, 72, rotate_deg=30.0, depth_noise=0.25)      # 150 views: one full group of 128 + a short one
    start = np.linspace(-1.0, 1.0, s.grid.n_voxels)
    single = run_gpu(gpu_ctx, s, np.float64, start=start)
    with Group(list(range(n_gpus))) as grp:
        grp.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing, s.rp.thick, s.rp.rho, s.rp.eta,
                       s.rp.delta, (s.W, s.H))
        got = start.copy()
        grp.process_depth_maps(s.depths, s.best_cost, 0.14, s.K, s.RT, got)       # accumulates onto io_scalar (:323-327)
        assert np.array_equal(got.view(np.uint64), single.view(np.uint64))
        f32 = np.zeros(s.grid.n_voxels, dtype=np.float32)
        grp.process_depth_maps(s.depths[:20], None, 0.0, s.K[:20], s.RT[:20], f32)
        gpu_ctx.initialize(s.grid.matrix, s.grid.point_dims, s.grid.origin, s.grid.spacing, s.rp.thick, s.rp.rho, s.rp.eta,
                           s.rp.delta, (s.W, s.H))
        one32 = np.zeros(s.grid.n_voxels, dtype=np.float32)
        gpu_ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_AUTO)
        gpu_ctx.process_depth_maps(s.depths[:20], None, 0.0, s.K[:20], s.RT[:20], one32)
        assert np.array_equal(f32.view(np.uint32), one32.view(np.uint32))
        # coloration: points sharded by index, colour images all-gathered
        rng = np.random.RandomState(4)
        pts = np.concatenate([syn.fibonacci_sphere_points(3001), rng.uniform(-1.3, 1.3, size=(1000, 3)).astype(np.float32)])
        want = oracle.colorize(pts, s.colors[:37], s.K[:37], s.RT[:37], s.W, s.H)
        mean, med, nb = grp.colorize(pts, s.colors[:37], s.K[:37], s.RT[:37], s.W, s.H)
        assert np.array_equal(nb, want[2]) and np.array_equal(med, want[1]) and np.array_equal(mean, want[0])
        from cudadepthmapintegration_b200 import DmiError
        with pytest.raises(DmiError) as e:
            grp.process_depth_maps(s.depths[:0], None, 0.0, s.K[:0], s.RT[:0], got)
        assert e.value.code == _lib.DMI_ERR_NO_VIEWS


def test_group_rejects_bad_device_lists():
    from cudadepthmapintegration_b200 import DmiError
    with pytest.raises(DmiError):
        Group([0, 0])
    with pytest.raises(DmiError):
        Group([n_devices() + 3])
