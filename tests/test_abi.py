"""The C-ABI library loads without a GPU and exports every symbol include/abm_b200.h declares; the ctypes
binding covers each of them; and the product fails loudly (no fallback) when the library is missing."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "abm_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abm_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    import hipsc_abm_b200
    return hipsc_abm_b200


def test_header_declares_the_whole_path():
    names = declared_symbols()
    for must in ("abm_bin_sort", "abm_neighbor_search", "abm_jkr_forces", "abm_apply_forces", "abm_jkr_substep",
                 "abm_diffusion", "abm_sample_concentrations", "abm_deposit_morphogens", "abm_graph_edge_list"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in abm_b200.h but not exported"


def test_ctypes_binding_covers_every_symbol(built):
    from hipsc_abm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.abm_abi_version() == _lib.ABI_VERSION == 2 and lib.abm_compiled_arch() == 100


def test_no_process_wide_options(built):
    """no global mutable state.  Launch options are arguments (ABM_LOOP_CHAINED, max_fused_substeps)."""
    assert "abm_set_option" not in declared_symbols()
    lib = ctypes.CDLL(built.LIB_PATH)
    assert not hasattr(lib, "abm_set_option")
    from hipsc_abm_b200 import _lib
    assert _lib.load().abm_diffusion_launches(10, 0) == 2 and _lib.load().abm_diffusion_launches(10, 1) == 10


def test_host_helpers_need_no_gpu(built):
    from hipsc_abm_b200 import _lib
    lib = _lib.load()
    dims = (ctypes.c_int32 * 3)()
    assert lib.abm_bin_dims_host(_lib.d3([2000, 2000, 0]), 15.0, _lib.SORT_REFERENCE, dims) == 0
    assert list(dims) == list(np.ceil(np.array([2000, 2000, 0]) / 15).astype(int) + 3)   # simulation.py:202
    assert lib.abm_bin_dims_host(_lib.d3([2000, 2000, 0]), 15.0, 0, dims) == 0
    assert list(dims) == [137, 137, 1]             # a flat world needs one z layer, not the reference's three
    # bond reach for two r=5 cells = the break distance probed in SURVEY 3.3 (10.31001 um)
    reach = lib.abm_jkr_reach(10.0, 5.0, 0.5, 1000.0, 0.000107)
    assert abs(reach - 10.31001) < 1e-4
    assert lib.abm_jkr_reach(15.0, 5.0, 0.5, 1000.0, 0.000107) >= 15.0
    # invalid arguments come back as negative codes with a message, never as exceptions from C
    assert lib.abm_bin_dims_host(_lib.d3([1, 1, 1]), -1.0, 0, dims) == -1
    assert b"distance" in lib.abm_last_error()
    assert lib.abm_bin_sort_workspace_bytes(1000, 5000) > 3 * 4000


def test_missing_library_fails_loudly(built, monkeypatch):
    from hipsc_abm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libabm_b200.so")
    with pytest.raises(_lib.AbmError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_engine_refuses_to_run_without_cuda(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hipsc_abm_b200.engine import ColonyEngine
    with pytest.raises(built.AbmError, match="no CPU path"):
        ColonyEngine([10, 10, 0])


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (or the import shim) may reference it."""
    pkg = os.path.join(ROOT, "python-hipsc-abm_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "libabm_oracle" in text:
                    offenders.append(os.path.join(base, f))
    assert not offenders, offenders
    assert "oracle" not in open(os.path.join(ROOT, "hipsc_abm_b200.py")).read()


def test_flag_constants_of_the_binding_equal_the_header():
    """The ctypes side passes flags as plain ints: every ABM_SORT_* / ABM_LOOP_* / ABM_SUBSTEP_* it defines must have
    the header's value, and the bits of one family must not collide."""
    import re
    from hipsc_abm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "abm_b200.h")).read()
    defined = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+ABM_((?:SORT|LOOP|SUBSTEP)_[A-Z0-9_]+)\s+(\d+)",
                                                                 header)}
    bound = {name: getattr(_lib, name) for name in dir(_lib) if re.match(r"(SORT|LOOP|SUBSTEP)_[A-Z0-9_]+$", name)}
    assert bound, "no flag constants found in _lib"
    for name, value in bound.items():
        assert defined.get(name) == value, f"{name}: binding {value}, header {defined.get(name)}"
    for family in ("SORT", "LOOP", "SUBSTEP"):
        bits = [v for k, v in defined.items() if k.startswith(family + "_")]
        assert all(v and v & (v - 1) == 0 for v in bits), f"ABM_{family}_* must be single bits: {bits}"
        assert len(set(bits)) == len(bits), f"ABM_{family}_* bits collide: {bits}"
