"""Known-answer tests of the oracle.  The reference ships no tests, so these
hand-derived values (from the statements cited next to each) are the first pin of the restatement."""
import numpy as np
import pytest

T, R, E, D = 0.08, 0.8, 0.03, 0.3      # the example command line, Reconstruction/main.cxx:102


@pytest.mark.parametrize("diff,want", [
    (0.0, 0.0), (0.04, 0.4), (-0.04, -0.4),            # a <= Thick: (Rho/Thick)*diff   CudaReconstruction.cu:119
    (0.08, 0.8), (-0.08, -0.8),                        # a == Thick is still the linear branch (:116 is strict >)
    (0.2, 0.8), (-0.2, -0.8),                          # Thick < a <= Delta: Rho*sign   :117
    (0.3, 0.8),                                        # a == Delta: strict > at :114
    (0.31, 0.0), (-0.31, -0.024),                      # beyond Delta: 0 behind, -Eta*Rho in front  :115
])
def test_ray_potential(oracle, diff, want):
    depth = 2.0
    got = oracle.ray_potential(depth + diff, depth, T, R, E, D)
    assert got == pytest.approx(want, abs=1e-15)


def test_ray_potential_exact_values(oracle):
    # exactly representable inputs -> exact outputs
    assert oracle.ray_potential(2.5, 2.0, 1.0, 0.5, 0.25, 2.0) == 0.25          # linear: (0.5/1)*0.5
    assert oracle.ray_potential(5.0, 2.0, 1.0, 0.5, 0.25, 2.0) == 0.0           # far behind
    assert oracle.ray_potential(-1.0, 2.0, 1.0, 0.5, 0.25, 2.0) == -0.125       # far in front: -Eta*Rho
    assert oracle.ray_potential(0.5, 2.0, 1.0, 0.5, 0.25, 2.0) == -0.5          # band: -Rho


@pytest.mark.parametrize("u,want", [
    (2.5, 3), (-0.5, -1), (-0.4, 0), (0.49999999999999994, 0), (1919.5, 1920), (1919.4999, 1919),
    (float("nan"), -2**31),                  # cvt.rzi.s32.f64 of NaN = 1<<31 (PTX ISA cvt; measured on B200)
    (float("inf"), 2**31 - 1), (float("-inf"), -2**31), (1e300, 2**31 - 1),
])
def test_round_to_pixel_device_semantics(oracle, u, want):
    assert oracle.round_to_pixel(u) == want


def test_threshold_is_strict(oracle):
    d = np.array([1.0, 2.0, 3.0, 4.0])
    c = np.array([0.13, 0.14, 0.140001, 1.0])
    out = oracle.apply_depth_threshold(d, c, 0.14)        # ReconstructionData.cxx:162: bestCost > threshold
    assert out.tolist() == [1.0, 2.0, -1.0, -1.0]


def test_median(oracle):
    assert oracle.median([10, 21]) == 15.5                 # Helper.h:181; truncated to 15 by the uchar store
    assert oracle.median([255, 255, 0, 0]) == 127.5
    assert oracle.median([3, 1, 2]) == 2
    assert oracle.median([7]) == 7


def _one_view_identity(W, H, f=100.0):
    K = np.eye(4); K[0, 0] = K[1, 1] = f; K[0, 2] = W / 2; K[1, 2] = H / 2
    RT = np.eye(4)
    return K.reshape(1, 16), RT.reshape(1, 16)


def test_world_to_pixel_no_z_test_and_x86_conversion(oracle):
    W, H = 64, 48
    K, RT = _one_view_identity(W, H)
    assert oracle.world_to_pixel(K[0], RT[0], [0, 0, 1]) == (32, 24)
    # behind the camera still projects (no z-sign test, ReconstructionData.cxx:169-182)
    assert oracle.world_to_pixel(K[0], RT[0], [0.01, 0.02, -1]) == (31, 22)
    # d.z == 0 -> division by zero -> x86 conversion gives INT_MIN
    px, py = oracle.world_to_pixel(K[0], RT[0], [1, 1, 0])
    assert px == -2**31 and py == -2**31
    px, py = oracle.world_to_pixel(K[0], RT[0], [0, 0, 0])      # 0/0 = NaN
    assert px == -2**31 and py == -2**31


def test_colorize_statistics(oracle):
    # 4 views looking at the same pixel with colours chosen so that mean != median and truncation shows
    W, H = 8, 6
    K, RT = _one_view_identity(W, H, f=10.0)
    K = np.repeat(K, 4, axis=0); RT = np.repeat(RT, 4, axis=0)
    colors = np.zeros((4, H, W, 3), dtype=np.uint8)
    vals = [(10, 255, 0), (21, 255, 7), (0, 0, 200), (0, 0, 201)]
    # point (0,0,1) -> pixel (4,3) -> storage row H-1-3 = 2
    for v, rgb in enumerate(vals):
        colors[v, 2, 4] = rgb
    pts = np.array([[0, 0, 1], [100, 100, 1]], dtype=np.float32)    # second point: outside every image
    mean, median, nb = oracle.colorize(pts, colors, K, RT, W, H)
    assert nb.tolist() == [4, 0]
    assert mean[0].tolist() == [31 // 4, 510 // 4, 408 // 4]        # int accumulate, truncated (MeshColoration.cxx:176-180)
    assert median[0].tolist() == [5, 127, 103]                      # (0+10)/2, (0+255)/2 -> 127.5 -> 127, (7+200)/2 -> 103.5
    assert mean[1].tolist() == [0, 0, 0] and median[1].tolist() == [0, 0, 0]
