"""Host-side statement of the lossless split form of a filtered depth map that prepared views are exchanged in
(include/dmi_b200.h, dmi_prepare_views_device: d_cls + d_lo; device code: csrc/tsdf_device.cuh, split_encode /
split_decode).  Not on any product path -- the library builds these arrays on the GPU -- but an integrator who
stores or ships prepared views can produce and check them with numpy.

    hi  float32   the depth rounded to nearest float; -1.0f EXACTLY on pixels that are invalid after the best-cost
                  filter (depth == -1, or best cost > threshold); a valid depth that rounds to -1.0f takes the
                  neighbouring float instead, so -1.0f never appears on a valid pixel
    lo  int32     (depth - hi) in units of 2^(e - 53), e = max(unbiased exponent of hi, -64); 0 on invalid pixels
                  and where hi is not finite

depth == hi + lo * 2^(e - 53) exactly whenever |depth| >= 2^-64 or depth == 0 (inf and NaN are kept by hi; a depth
of -0.0 comes back as +0.0, the same number)."""
from __future__ import annotations

import numpy as np

_NEG1_BELOW = np.float32(-1.00000012)     # the floats next to -1.0f
_NEG1_ABOVE = np.float32(-0.99999994)


def _unit_exponent(hi: np.ndarray) -> np.ndarray:
    e = ((hi.view(np.int32) >> 23) & 0xFF) - 127
    return np.maximum(e, -64)


def encode(depth: np.ndarray, best_cost: np.ndarray | None = None, threshold: float = 0.0):
    """(hi, lo) of a double depth map (any shape); the best-cost filter is strict, like the reference's
    (ReconstructionData.cxx:162)."""
    d = np.asarray(depth, dtype=np.float64)
    invalid = d == -1.0
    if best_cost is not None:
        invalid = invalid | (np.asarray(best_cost, dtype=np.float64) > threshold)
    with np.errstate(over="ignore", invalid="ignore"):
        hi = d.astype(np.float32)
        clash = (hi == np.float32(-1.0)) & ~invalid
        hi = np.where(clash, np.where(d < -1.0, _NEG1_BELOW, _NEG1_ABOVE), hi).astype(np.float32)
        hi = np.where(invalid, np.float32(-1.0), hi).astype(np.float32)
        finite = np.isfinite(hi) & ~invalid
        scale = np.exp2((53 - _unit_exponent(hi)).astype(np.float64))
        resid = np.where(finite, (d - hi.astype(np.float64)) * scale, 0.0)
        lo = np.rint(resid).astype(np.int64)
    assert np.all(np.abs(lo) < 2 ** 31)
    return hi, lo.astype(np.int32)


def decode(hi: np.ndarray, lo: np.ndarray) -> np.ndarray:
    """The double depth of every pixel (meaningless where hi == -1.0f: those pixels are invalid)."""
    hi = np.asarray(hi, dtype=np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        unit = np.exp2((_unit_exponent(hi) - 53).astype(np.float64))
        return hi.astype(np.float64) + np.asarray(lo, dtype=np.float64) * unit
