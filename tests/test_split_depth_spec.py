"""The interchange format of prepared views (float classification image + int32 residual image), as stated in
cudadepthmapintegration_b200/split_depth.py: lossless round trip on the CPU.  tests/test_tsdf_parity_gpu.py checks
the device encoder against the same statement."""
import numpy as np
from hypothesis import given, settings, strategies as st

from cudadepthmapintegration_b200 import split_depth


def test_round_trip_on_awkward_values():
    d = np.array([np.nextafter(-1.0, 0.0), np.nextafter(-1.0, -2.0), -1.0 - 2.0 ** -25, -1.0 + 2.0 ** -26, 1.0 + 2.0 ** -52,
                  3.0 - 2.0 ** -51, 2.0 - 2.0 ** -53, np.pi, 2.0 ** -64, 0.0, -0.0, 3.4e38, 65504.123456789, 1e-3, -7.25,
                  2.0 ** -126 * 1.5, 123456789.123456789])
    hi, lo = split_depth.encode(d)
    assert not np.any(hi == np.float32(-1.0))                 # all valid: -1.0f must not appear
    back = split_depth.decode(hi, lo)
    assert np.array_equal(back, d)                            # (-0.0 comes back as +0.0: the same number)


def test_invalid_pixels_and_the_strict_filter():
    d = np.array([2.0, -1.0, 2.5, 3.0])
    cost = np.array([0.14, 0.0, 0.14000001, 0.2])
    hi, lo = split_depth.encode(d, cost, 0.14)
    assert list(hi == np.float32(-1.0)) == [False, True, True, True]      # cost == threshold stays valid (strict >)
    assert lo[1] == 0 and lo[2] == 0 and lo[3] == 0
    assert split_depth.decode(hi, lo)[0] == 2.0


def test_non_finite_depths_are_kept_by_the_float():
    d = np.array([np.inf, -np.inf, np.nan, 1e300, -1e300])
    hi, lo = split_depth.encode(d)
    assert np.all(lo == 0)
    back = split_depth.decode(hi, lo)
    assert back[0] == np.inf and back[1] == -np.inf and np.isnan(back[2])
    assert back[3] == np.inf and back[4] == -np.inf           # beyond float range: classified "far" either way


def test_tiny_magnitudes_keep_absolute_accuracy():
    d = np.array([1e-30, -3e-25, 5e-324])
    hi, lo = split_depth.encode(d)
    assert np.all(np.abs(split_depth.decode(hi, lo) - d) < 2.0 ** -117)


@settings(max_examples=300, deadline=None)
@given(st.lists(st.floats(min_value=-1e30, max_value=1e30, allow_nan=False, allow_infinity=False), min_size=1, max_size=64))
def test_round_trip_property(values):
    d = np.array(values)
    d = d[(np.abs(d) >= 2.0 ** -64) | (d == 0)]
    d = d[d != -1.0]
    hi, lo = split_depth.encode(d)
    assert np.array_equal(split_depth.decode(hi, lo), d)
    assert not np.any(hi == np.float32(-1.0))
