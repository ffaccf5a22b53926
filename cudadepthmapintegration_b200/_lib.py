"""ctypes binding of libdmi_b200.so (the C ABI declared in include/dmi_b200.h).

The shared object is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``).  There is no
fallback of any kind: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DMI_B200_LIBRARY: another build of the same library (tuning experiments); the default is the in-tree one
LIB_PATH = os.environ.get("DMI_B200_LIBRARY") or os.path.join(_HERE, "libdmi_b200.so")

DMI_OK = 0
DMI_ERR_INVALID_ARGUMENT = 1
DMI_ERR_NOT_INITIALIZED = 2
DMI_ERR_CUDA = 3
DMI_ERR_NO_VIEWS = 4
DMI_ERR_BAD_PARAMETERS = 5
DMI_ERR_OUT_OF_MEMORY = 6
DMI_F32 = 0
DMI_F64 = 1
DMI_TSDF_KERNEL_AUTO = 0
DMI_TSDF_KERNEL_EXACT = 1
DMI_OPT_TSDF_KERNEL = 1
DMI_OPT_VIEW_CHUNK = 2
DMI_OPT_TIER_COUNTERS = 3
DMI_OPT_CULL = 4
DMI_OPT_BRICK_QUOTA = 5
DMI_UNIQUE_ID_BYTES = 128


class DmiError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"dmi error {code}: {message}")
        self.code = code
        self.message = message


_lib = None

_vp, _i, _d, _sz, _ll = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_longlong
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)

_PROTOTYPES = {
    "dmi_abi_version": (C.c_int, []),
    "dmi_device_count": (C.c_int, [_pi]),
    "dmi_create": (C.c_int, [_i, C.POINTER(_vp)]),
    "dmi_destroy": (C.c_int, [_vp]),
    "dmi_last_error": (C.c_char_p, [_vp]),
    "dmi_set_stream": (C.c_int, [_vp, _vp]),
    "dmi_use_own_stream": (C.c_int, [_vp]),
    "dmi_synchronize": (C.c_int, [_vp]),
    "dmi_set_option": (C.c_int, [_vp, _i, _ll]),
    "dmi_initialize": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _vp]),
    "dmi_set_slab": (C.c_int, [_vp, _i, _i]),
    "dmi_set_slab_layers": (C.c_int, [_vp, _i, _i, _i]),
    "dmi_slab_planes": (C.c_int, [_vp, _pi]),
    "dmi_process_depth_maps": (C.c_int, [_vp, _i, _vp, _vp, _d, _vp, _vp, _vp, _i]),
    "dmi_volume_begin": (C.c_int, [_vp, _vp, _i]),
    "dmi_volume_integrate_host": (C.c_int, [_vp, _i, _vp, _vp, _d, _vp, _vp]),
    "dmi_volume_integrate_device": (C.c_int, [_vp, _i, _vp, _vp, _d, _vp, _vp]),
    "dmi_volume_end": (C.c_int, [_vp, _vp]),
    "dmi_prepared_view_sizes": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "dmi_prepare_views_device": (C.c_int, [_vp, _i, _vp, _vp, _d, _vp, _vp, _ll, _vp]),
    "dmi_volume_integrate_prepared": (C.c_int, [_vp, _i, _vp, _vp, _vp, _ll, _vp, _vp, _vp]),
    "dmi_volume_device_ptr": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_sz)]),
    "dmi_apply_depth_threshold_device": (C.c_int, [_vp, _sz, _vp, _vp, _d]),
    "dmi_tsdf_kernel_stats": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(_ll)]),
    "dmi_tsdf_tier_counters": (C.c_int, [_vp, C.POINTER(C.c_ulonglong)]),
    "dmi_colorize": (C.c_int, [_vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "dmi_colorize_device": (C.c_int, [_vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "dmi_color_kernel_stats": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(_ll)]),
    "dmi_measure_fp_peak": (C.c_int, [_vp, _i, _d, _pd]),
    "dmi_launch_counter": (C.c_int, [_vp, C.POINTER(_ll)]),
    "dmi_contour_device": (C.c_int, [_vp, _vp, _i, _d, C.POINTER(_sz), C.POINTER(_sz)]),
    "dmi_contour": (C.c_int, [_vp, _vp, _i, _d, C.POINTER(_sz), C.POINTER(_sz)]),
    "dmi_contour_get": (C.c_int, [_vp, _vp, _vp]),
    "dmi_contour_device_ptr": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_sz)]),
    "dmi_comm_unique_id": (C.c_int, [C.c_char_p]),
    "dmi_comm_init": (C.c_int, [_vp, C.c_char_p, _i, _i]),
    "dmi_comm_destroy": (C.c_int, [_vp]),
    "dmi_comm_info": (C.c_int, [_vp, _pi, _pi, _pi]),
    "dmi_comm_copy_engines": (C.c_int, [_vp]),
    "dmi_shard_initialize": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _vp]),
    "dmi_shard_view_count": (C.c_int, [_i, _i, _i, _pi]),
    "dmi_shard_view_indices": (C.c_int, [_i, _i, _i, _vp]),
    "dmi_shard_group_count": (C.c_int, [_i, _i, _pi]),
    "dmi_shard_group_starts": (C.c_int, [_i, _i, _vp]),
    "dmi_shard_integrate_device": (C.c_int, [_vp, _i, _vp, _vp, _d, _vp, _vp]),
    "dmi_shard_integrate_host": (C.c_int, [_vp, _i, _vp, _vp, _d, _vp, _vp]),
    "dmi_shard_gather_volume_device": (C.c_int, [_vp, _i, _vp]),
    "dmi_shard_range": (C.c_int, [_sz, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "dmi_shard_colorize_device": (C.c_int, [_vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "dmi_group_create": (C.c_int, [_vp, _i, C.POINTER(_vp)]),
    "dmi_group_destroy": (C.c_int, [_vp]),
    "dmi_group_last_error": (C.c_char_p, [_vp]),
    "dmi_group_size": (C.c_int, [_vp]),
    "dmi_group_context": (_vp, [_vp, _i]),
    "dmi_group_set_option": (C.c_int, [_vp, _i, _ll]),
    "dmi_group_initialize": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _vp]),
    "dmi_group_process_depth_maps": (C.c_int, [_vp, _i, _vp, _vp, _d, _vp, _vp, _vp, _i]),
    "dmi_group_colorize": (C.c_int, [_vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
}


def load():
    """Load libdmi_b200.so; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C cudadepthmapintegration_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def check(ctx_handle, code: int):
    if code != DMI_OK:
        msg = load().dmi_last_error(ctx_handle)
        raise DmiError(code, msg.decode() if msg else "")
