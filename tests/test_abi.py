"""The C-ABI library loads on a machine without a GPU, exports every symbol that
include/bmc.h declares, and FAILS LOUDLY (no CPU fallback) when asked to compute."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bmc.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(bmc_\w+)\s*\(", src, re.M)))


def test_header_symbols_exported(bmc):
    lib = bmc.load_library()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(bmc.ABI_SYMBOLS) == names


def test_struct_layouts_match_header(bmc):
    assert ctypes.sizeof(bmc.BmcLeavingFlow) == 24
    assert ctypes.sizeof(bmc.BmcCounters) == 8 * (6 + 15)
    assert ctypes.sizeof(bmc.BmcConfig) == 16 + 8 * 4 + 8 + 8 * 4 + 8 + 8


def test_no_cpu_fallback(bmc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is exercised on CPU-only machines")
    with pytest.raises(bmc.BmcError):
        bmc.ParticleLoop("monod", 1, 500)


def test_invalid_arguments_return_codes(bmc):
    lib = bmc.load_library()
    h = ctypes.c_void_p()
    assert lib.bmc_create(ctypes.byref(h), None) == -1            # BMC_ERR_INVALID
    cfg = bmc.BmcConfig(model=99, n_species=1, n_compartments=1)
    assert lib.bmc_create(ctypes.byref(h), ctypes.byref(cfg)) != 0 and not h.value
    cfg = bmc.BmcConfig(model=0, n_species=0, n_compartments=1)
    assert lib.bmc_create(ctypes.byref(h), ctypes.byref(cfg)) == -1
    assert lib.bmc_destroy(ctypes.byref(h)) == -1                # null handle
    assert lib.bmc_cycle(None, 0.1) == -1
    assert lib.bmc_last_error(None) == b"null context"


def test_product_does_not_reference_oracle():
    """the product path must never import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "biocma-mcst_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "bmc_oracle" not in txt and "import oracle" not in txt and "orc_" not in txt, f
