"""The C-ABI shared library loads and exports every symbol include/pixelflow_gpu.h declares.
No compute calls (this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pixelflow_b200 import build
    return ctypes.CDLL(build.build_library())


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pixelflow_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(pf_[a-z_0-9]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_expected_entry_points():
    from pixelflow_b200 import EXPORTS
    assert declared_functions() == sorted(EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_abi_version_and_config_defaults(lib):
    from pixelflow_b200.api import PfConfig
    lib.pf_abi_version.restype = ctypes.c_int
    assert lib.pf_abi_version() == 2
    cfg = PfConfig()
    lib.pf_config_init(ctypes.byref(cfg))
    assert cfg.struct_size == ctypes.sizeof(PfConfig)
    assert (cfg.iter_max, cfg.nranks, cfg.device, cfg.use_graph, cfg.halo_transport) == (100, 1, -1, 1, 0)
    assert cfg.relux_factor == 1.7 and cfg.thickness == 1.5
    assert list(cfg.wall) == [1, 0, 0, 0, 2, 0]


def test_create_rejects_bad_configs_without_touching_a_gpu(lib):
    """validation happens before any CUDA call, so these paths are testable on CPU"""
    from pixelflow_b200.api import PfConfig
    lib.pf_last_error.restype = ctypes.c_char_p
    lib.pf_last_error.argtypes = [ctypes.c_void_p]
    cfg = PfConfig()
    lib.pf_config_init(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    cfg.m, cfg.n, cfg.l = 1, 8, 8
    assert lib.pf_create(ctypes.byref(h), ctypes.byref(cfg)) != 0
    assert b"m and n" in lib.pf_last_error(None)
    cfg.m = 8
    cfg.struct_size = 4
    assert lib.pf_create(ctypes.byref(h), ctypes.byref(cfg)) != 0
    assert b"struct_size" in lib.pf_last_error(None)
    lib.pf_config_init(ctypes.byref(cfg))
    cfg.m, cfg.n, cfg.l = 8, 8, 8
    cfg.solver_case = 0  # 2D
    cfg.nranks, cfg.rank = 2, 0
    assert lib.pf_create(ctypes.byref(h), ctypes.byref(cfg)) != 0
    assert b"2D" in lib.pf_last_error(None)


def test_product_package_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under pixelflow_b200/ may reference it"""
    pkg = os.path.join(ROOT, "pixelflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f in (), f"{f} mentions the oracle"
