"""The C-ABI library loads without a GPU and exports every symbol include/dmi_b200.h declares.
No compute is called here; without a device the library must FAIL LOUDLY, not fall back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dmi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmi_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path_entry_points():
    syms = declared_symbols()
    for must in ("dmi_initialize", "dmi_process_depth_maps", "dmi_colorize", "dmi_set_slab",
                 "dmi_volume_integrate_device", "dmi_apply_depth_threshold_device"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from cudadepthmapintegration_b200 import _lib
    lib = _lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), f"libdmi_b200.so does not export {s}"
    assert sorted(_lib.exported_symbols()) == declared_symbols()
    assert lib.dmi_abi_version() == 2


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cudadepthmapintegration_b200 import Context, DmiError
    with pytest.raises(DmiError):
        Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cudadepthmapintegration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cxx", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle/" not in text and "_oracle" not in text, f
