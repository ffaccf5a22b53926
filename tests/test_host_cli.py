"""The C++ host mirror (csrc/host): list-file / .krtd / .vti parsing without a GPU, and the CLI end to end
on the GPU against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from cudadepthmapintegration_b200 import dataset_io
from tests.scenes import Scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cudadepthmapintegration_b200", "dmi_cli")


def _need_cli():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "cudadepthmapintegration_b200", "csrc")], check=True, capture_output=True)


def test_parsers_read_what_was_written(tmp_path):
    _need_cli()
    s = Scene(8, 3, 40, 30, rotate_deg=10.0)
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT, ascii_views=(1,))
    r = subprocess.run([CLI, "inspect", "--vti", str(tmp_path / "vtiList.txt"), "--krtd", str(tmp_path / "kList.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "views 3 krtd 3"
    for v, line in enumerate(lines[1:]):
        t = line.split()
        assert int(t[3]) == 40 and int(t[5]) == 30
        assert float(t[7]) == pytest.approx(float(s.depths[v].sum()), rel=1e-13)
        assert float(t[9]) == pytest.approx(float(s.best_cost[v].sum()), rel=1e-13)
        assert int(t[11]) == int(s.colors[v].astype(np.int64).sum())
        assert int(t[13]) == int((s.depths[v] == -1).sum())
        K = np.array([float(x) for x in t[15:31]])
        RT = np.array([float(x) for x in t[32:48]])
        assert np.array_equal(K, s.K[v]) and np.array_equal(RT, s.RT[v])      # repr() round-trips doubles


VTI_LAYOUTS = [dict(encoding="raw"), dict(encoding="ascii"), dict(encoding="base64"), dict(encoding="base64", compress=True),
               dict(encoding="base64", compress=True, header_type="UInt64"), dict(encoding="binary"),
               dict(encoding="binary", compress=True), dict(encoding="raw", compress=True, header_type="UInt64"),
               dict(encoding="base64", header_with_data=True), dict(encoding="binary", header_with_data=True, header_type="UInt64")]


def test_vti_reader_handles_the_xml_writers_layouts(tmp_path):
    """ascii, inline base64, appended raw / base64, zlib blocks, 32- and 64-bit block headers, the byte count encoded
    on its own or with the data: every view of the dataset uses another layout and must read back identically."""
    _need_cli()
    n = len(VTI_LAYOUTS)
    s = Scene(8, n, 72, 50, rotate_deg=10.0)          # 72 x 50 x 8 B = 28.8 kB: less than one 32 kB zlib block ...
    big = Scene(8, 1, 160, 120)                       # ... and 153.6 kB: several blocks with a ragged last one
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT, vti_options=VTI_LAYOUTS)
    r = subprocess.run([CLI, "inspect", "--vti", str(tmp_path / "vtiList.txt"), "--krtd", str(tmp_path / "kList.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == f"views {n} krtd {n}"
    for v, line in enumerate(lines[1:]):
        t = line.split()
        assert float(t[7]) == pytest.approx(float(s.depths[v].sum()), rel=1e-13), VTI_LAYOUTS[v]
        assert float(t[9]) == pytest.approx(float(s.best_cost[v].sum()), rel=1e-13), VTI_LAYOUTS[v]
        assert int(t[11]) == int(s.colors[v].astype(np.int64).sum()), VTI_LAYOUTS[v]
        assert int(t[13]) == int((s.depths[v] == -1).sum()), VTI_LAYOUTS[v]
    sub = tmp_path / "big"
    dataset_io.write_dataset(str(sub), big.depths, big.best_cost, big.colors, big.K, big.RT,
                             vti_options=[dict(encoding="base64", compress=True)])
    r = subprocess.run([CLI, "inspect", "--vti", str(sub / "vtiList.txt"), "--krtd", str(sub / "kList.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = r.stdout.strip().splitlines()[1].split()
    assert float(t[7]) == pytest.approx(float(big.depths[0].sum()), rel=1e-13)
    assert int(t[11]) == int(big.colors[0].astype(np.int64).sum())


def test_vti_reader_reports_a_corrupt_block(tmp_path):
    _need_cli()
    s = Scene(8, 1, 40, 30)
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT, vti_options=[dict(encoding="base64", compress=True)])
    f = tmp_path / "view_0000.vti"
    data = f.read_bytes()
    cut = data.index(b"_") + 40
    f.write_bytes(data[:cut] + b"AAAA" + data[cut + 4:])          # damage the first compressed stream
    r = subprocess.run([CLI, "inspect", "--vti", str(tmp_path / "vtiList.txt"), "--krtd", str(tmp_path / "kList.txt")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Depths" in r.stderr


def test_readers_on_a_vtk_style_file_assembled_independently():
    """tests/golden/vtk_style_view.vti was assembled by tests/golden/make_vti_fixture.py from the VTK file-format description
    (version 0.1, UInt32 headers, vtkZLibDataCompressor, appended base64 with separately encoded block header) without
    dataset_io's writer: both readers (C++ DmiVti.h through `dmi_cli inspect`, Python dataset_io) must return its arrays."""
    _need_cli()
    import importlib.util
    g = os.path.join(ROOT, "tests", "golden")
    spec = importlib.util.spec_from_file_location("make_vti_fixture", os.path.join(g, "make_vti_fixture.py"))
    fx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fx)
    depths, cost, color, K, R, T = fx.arrays()
    d, c, col, Ks, RTs = dataset_io.load_dataset(os.path.join(g, "vtk_style_vtiList.txt"), os.path.join(g, "vtk_style_kList.txt"))
    assert np.array_equal(d[0], depths) and np.array_equal(c[0], cost) and np.array_equal(col[0], color)
    K4 = Ks.reshape(-1, 4, 4)[0]; RT4 = RTs.reshape(-1, 4, 4)[0]
    assert np.array_equal(K4[:3, :3], K) and np.array_equal(RT4[:3, :3], R) and np.array_equal(RT4[:3, 3], T)
    r = subprocess.run([CLI, "inspect", "--vti", os.path.join(g, "vtk_style_vtiList.txt"), "--krtd", os.path.join(g, "vtk_style_kList.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = r.stdout.strip().splitlines()[1].split()
    assert int(t[3]) == fx.W and int(t[5]) == fx.H
    assert float(t[7]) == pytest.approx(float(depths.sum()), rel=1e-13) and float(t[9]) == pytest.approx(float(cost.sum()), rel=1e-13)
    assert int(t[11]) == int(color.astype(np.int64).sum()) and int(t[13]) == int((depths == -1).sum())
    assert np.array_equal(np.array([float(x) for x in t[15:31]]).reshape(4, 4)[:3, :3], K)
    assert np.array_equal(np.array([float(x) for x in t[32:48]]).reshape(4, 4)[:3, 3], T)


def test_vti_reader_rejects_a_truncated_ascii_array(tmp_path):
    _need_cli()
    s = Scene(8, 1, 40, 30)
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT, vti_options=[dict(encoding="ascii")])
    f = tmp_path / "view_0000.vti"
    text = f.read_text()
    a = text.index(">", text.index('Name="Depths"')) + 1
    b = text.index("</DataArray>", a)
    vals = text[a:b].split()
    f.write_text(text[:a] + " ".join(vals[:len(vals) // 2]) + " garbage " + text[b:])       # half the values, then junk
    r = subprocess.run([CLI, "inspect", "--vti", str(tmp_path / "vtiList.txt"), "--krtd", str(tmp_path / "kList.txt")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Depths" in r.stderr


def test_cli_rejects_the_reference_cli_error_cases(tmp_path):
    _need_cli()
    base = [CLI, "reconstruction", "--gridOrigin", "-1", "-1", "-1", "--gridEnd", "1", "1", "1", "--gridDims", "9",
            "--dataFolder", str(tmp_path), "--outputGridFilename", str(tmp_path / "o.mhd")]
    # Delta < Thick and Eta outside [0,1] are argument errors (Reconstruction/main.cxx:270-271)
    assert subprocess.run(base + ["--rayThick", "0.5", "--rayDelta", "0.3"], capture_output=True).returncode != 0
    assert subprocess.run(base + ["--rayThick", "0.1", "--rayDelta", "0.3", "--rayEta", "1.5"], capture_output=True).returncode != 0
    # non-orthogonal grid vectors (:363-382)
    assert subprocess.run(base + ["--rayThick", "0.1", "--rayDelta", "0.3", "--gridVecX", "1", "1", "0"], capture_output=True).returncode != 0


@pytest.mark.gpu
def test_cli_end_to_end_against_oracle(tmp_path, oracle):
    _need_cli()
    from cudadepthmapintegration_b200 import synthetic as syn
    s = Scene(24, 5, 64, 48, rotate_deg=30.0, depth_noise=0.25)
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT)
    gm = s.grid.matrix.reshape(4, 4)
    out = tmp_path / "vol.mhd"
    cmd = [CLI, "reconstruction", "--gridDims", "25", "25", "25", "--gridOrigin", *[repr(float(x)) for x in s.grid.origin],
           "--gridEnd", *[repr(float(o + 25 * sp)) for o, sp in zip(s.grid.origin, s.grid.spacing)],
           "--gridVecX", *[repr(float(x)) for x in gm[0, :3]], "--gridVecY", *[repr(float(x)) for x in gm[1, :3]],
           "--gridVecZ", *[repr(float(x)) for x in gm[2, :3]], "--dataFolder", str(tmp_path),
           "--rayThick", repr(s.rp.thick), "--rayRho", repr(s.rp.rho), "--rayEta", repr(s.rp.eta), "--rayDelta", repr(s.rp.delta),
           "--threshBestCost", "0.14", "--outputGridFilename", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(str(tmp_path / "vol.raw"), dtype=np.float64)
    # the CLI derives spacing = (end - origin) / dims like the reference (main.cxx:318-323): rebuild the same grid
    spacing = (np.array([float(o + 25 * sp) for o, sp in zip(s.grid.origin, s.grid.spacing)]) - s.grid.origin) / 25.0
    grid = syn.Grid((24, 24, 24), s.grid.origin, spacing, s.grid.matrix)
    want = oracle.tsdf_integrate(grid, s.rp, s.W, s.H, s.depths, s.best_cost, 0.14, s.K, s.RT, np.zeros(24 ** 3))
    assert np.array_equal(got != 0, want != 0)
    assert np.abs(got - want).max() <= 1e-6
    # a corrupt depth map must not leave a valid-looking all-zero volume behind with exit code 0
    bad = tmp_path / "bad"
    dataset_io.write_dataset(str(bad), s.depths, s.best_cost, s.colors, s.K, s.RT, vti_options=[dict(encoding="base64", compress=True)] * 5)
    f = bad / "view_0003.vti"
    data = f.read_bytes()
    cut = data.index(b"_") + 40
    f.write_bytes(data[:cut] + b"AAAA" + data[cut + 4:])
    cmd_bad = [str(bad) if x == str(tmp_path) else x for x in cmd[:-1]] + [str(bad / "vol.mhd")]
    r = subprocess.run(cmd_bad, capture_output=True, text=True)
    assert r.returncode != 0
    assert not (bad / "vol.raw").exists() and not (bad / "vol.mhd").exists()
    # coloration through the CLI
    pts = syn.fibonacci_sphere_points(2000)
    pts.tofile(str(tmp_path / "pts.f32"))
    r = subprocess.run([CLI, "coloration", "--input", str(tmp_path / "pts.f32"), "--output", str(tmp_path / "col"),
                        "--krtd", str(tmp_path / "kList.txt"), "--vti", str(tmp_path / "vtiList.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    wmean, wmed, wnb = oracle.colorize(pts, s.colors, s.K, s.RT, s.W, s.H)
    assert np.array_equal(np.fromfile(str(tmp_path / "col.nb.i32"), dtype=np.int32), wnb)
    assert np.array_equal(np.fromfile(str(tmp_path / "col.median.u8"), dtype=np.uint8).reshape(-1, 3), wmed)
    assert np.array_equal(np.fromfile(str(tmp_path / "col.mean.u8"), dtype=np.uint8).reshape(-1, 3), wmean)


@pytest.mark.gpu
def test_cli_on_two_gpus_is_bit_identical(tmp_path):
    """dmi_cli --gpus 2 (dmihost::CudaReconstructionFilter::SetDevices / MeshColoration::SetDevices over dmi_group_*)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    _need_cli()
    from cudadepthmapintegration_b200 import synthetic as syn
    s = Scene((24, 20, 70), 9, 64, 48, depth_noise=0.25)
    dataset_io.write_dataset(str(tmp_path), s.depths, s.best_cost, s.colors, s.K, s.RT)
    end = [repr(float(o + n * sp)) for o, sp, n in zip(s.grid.origin, s.grid.spacing, s.grid.point_dims)]
    vols = []
    for gpus in (1, 2):
        out = tmp_path / f"vol{gpus}.mhd"
        cmd = [CLI, "reconstruction", "--gridDims", *[str(d) for d in s.grid.point_dims], "--gridOrigin", *[repr(float(x)) for x in s.grid.origin],
               "--gridEnd", *end, "--dataFolder", str(tmp_path), "--rayThick", repr(s.rp.thick), "--rayRho", repr(s.rp.rho),
               "--rayEta", repr(s.rp.eta), "--rayDelta", repr(s.rp.delta), "--outputGridFilename", str(out), "--gpus", str(gpus)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        vols.append(np.fromfile(str(tmp_path / f"vol{gpus}.raw"), dtype=np.float64))
    assert np.count_nonzero(vols[0]) > 0 and np.array_equal(vols[0].view(np.uint64), vols[1].view(np.uint64))
    pts = syn.fibonacci_sphere_points(2500)
    pts.tofile(str(tmp_path / "pts.f32"))
    outs = []
    for gpus in (1, 2):
        r = subprocess.run([CLI, "coloration", "--input", str(tmp_path / "pts.f32"), "--output", str(tmp_path / f"col{gpus}"),
                            "--krtd", str(tmp_path / "kList.txt"), "--vti", str(tmp_path / "vtiList.txt"), "--gpus", str(gpus)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append([np.fromfile(str(tmp_path / f"col{gpus}{sfx}"), dtype=np.uint8) for sfx in (".mean.u8", ".median.u8", ".nb.i32")])
    assert all(np.array_equal(a, b) for a, b in zip(*outs)) and outs[0][2].any()
