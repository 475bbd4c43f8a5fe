"""ctypes binding of libabm_b200.so (the C-ABI declared in include/abm_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABM_B200_LIB") or os.path.join(_HERE, "libabm_b200.so")   # env: kernel A/B builds

P = C.c_void_p          # device pointer / stream
I64 = C.c_int64
I32 = C.c_int
F64 = C.c_double
SZ = C.c_size_t
D3 = C.POINTER(C.c_double)    # host double[3]
I3 = C.POINTER(C.c_int32)     # host int32[3]
L3 = C.POINTER(C.c_int64)     # host int64[3]

# name -> (restype, argtypes); must list every symbol of include/abm_b200.h (tests/test_abi.py checks)
class ColonyArrays(C.Structure):
    """abm_colony_arrays: device pointers of every per-cell array of a device-resident colony."""
    _fields_ = [(name, C.c_void_p) for name in (
        "locations", "radii", "motility_forces", "gid", "states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK",
        "death_counters", "diff_counters", "div_counters", "fds_counters", "bond_len", "bond_rows", "nbr_start",
        "nbr_len")]


class RuleParams(C.Structure):
    """abm_rule_params (cell_simulation.py:33-57, templates/experimental.yaml)."""
    _fields_ = [("seed", C.c_int64), ("step", C.c_int64), ("field", C.c_int32), ("dox_step", C.c_int32),
                ("lonely_thresh", C.c_int32), ("guye_move", C.c_int32), ("pluri_div_thresh", C.c_int32),
                ("diff_div_thresh", C.c_int32), ("pluri_to_diff", C.c_int32), ("death_thresh", C.c_int32),
                ("fds_thresh", C.c_int32), ("three_d", C.c_int32), ("max_radius", C.c_double),
                ("min_radius", C.c_double), ("motility_force", C.c_double)]


CA, RP = C.POINTER(ColonyArrays), C.POINTER(RuleParams)

SIGNATURES = {
    "abm_abi_version": (I32, []),
    "abm_compiled_arch": (I32, []),
    "abm_last_error": (C.c_char_p, []),
    "abm_device_info": (I32, [C.POINTER(I32), C.POINTER(I32), C.POINTER(I32)]),
    "abm_bin_dims_host": (I32, [D3, F64, I32, I3]),
    "abm_bin_sort_workspace_bytes": (SZ, [I64, I64]),
    "abm_bin_sort": (I32, [P, P, I64, D3, D3, F64, I32, P, P, P, P, P, P, SZ, P, P]),
    "abm_bins_dense": (I32, [P, I64, D3, F64, P, P, I64, P, P, P, P]),
    "abm_neighbor_search": (I32, [P, P, P, P, I64, I3, F64, P, P, I64, P, P, P, P, P, I64, P, P, P]),
    "abm_read_i32_sync": (I32, [P, C.POINTER(C.c_int32), P]),
    "abm_graph_edge_list_workspace_bytes": (SZ, [I64]),
    "abm_graph_edge_list": (I32, [P, I64, P, P, I64, P, SZ, P, I64, P, P]),
    "abm_jkr_forces": (I32, [P, I64, P, P, P, P, P, I64, F64, F64, F64, P]),
    "abm_apply_forces": (I32, [P, P, P, P, I64, F64, D3, F64, P, P]),
    "abm_jkr_reach": (F64, [F64, F64, F64, F64, F64]),
    "abm_jkr_substep": (I32, [P, P, P, P, I64, I3, F64, F64, P, I64, P, P, P, P, I64, P, I32, P, P,
                              F64, F64, F64, F64, D3, F64, P, F64, D3, D3, P, P, P, P, P]),
    "abm_mechanics_loop": (I32, [P, P, P, I64, D3, D3, D3, F64, F64, I32, C.POINTER(C.c_double),
                                 P, P, P, P, P, P, P, SZ, P, I64, P, P, P, P, P, P, I64, P, P,
                                 F64, F64, F64, F64, F64, I32, P, P, P]),
    "abm_mechanics_loop_timed": (I32, [P, P, P, I64, D3, D3, D3, F64, F64, I32, C.POINTER(C.c_double),
                                       P, P, P, P, P, P, P, SZ, P, I64, P, P, P, P, P, P, I64, P, P,
                                       F64, F64, F64, F64, F64, I32, P, P, C.POINTER(C.c_void_p), P]),
    "abm_diffusion": (I32, [P, P, I64, I64, I64, F64, F64, F64, F64, F64, F64, I32, I32, C.POINTER(I32), P]),
    "abm_diffusion_launches": (I64, [I64, I32]),
    "abm_diffusion3d": (I32, [P, P, I64, I64, I64, I64, F64, F64, F64, F64, F64, F64, I32, C.POINTER(I32), P]),
    "abm_deposit_morphogens3d": (I32, [P, P, F64, I64, P, L3, F64, I64, I64, P, P]),
    "abm_sample_concentrations": (I32, [P, I64, P, L3, F64, I64, I64, P, P, P]),
    "abm_rules_workspace_bytes": (SZ, [I64]),
    "abm_rule_division": (I32, [CA, I64, I64, P, P, RP, P, P, SZ, P]),
    "abm_rule_death": (I32, [CA, CA, I64, I64, P, P, RP, P, P, SZ, P]),
    "abm_rule_update": (I32, [CA, I64, P, P, P, RP, P, P]),
    "abm_slab_workspace_bytes": (SZ, [I64]),
    "abm_slab_pack": (I32, [P, P, P, P, P, P, I64, P, I64, F64, F64, I32, I32, P, I64, P, I64, P, P, SZ, P]),
    "abm_slab_merge": (I32, [P, P, P, P, P, P, I64, P, I64, P, I64, I64, F64, F64, I32, I32, P, P, P, P, P, P, P, P, SZ,
                             P, P, P, I64, P]),
    "abm_slab_exchange_pack": (I32, [P, P, P, P, P, P, I64, P, I64, F64, F64, I32, I32, P, P, I64, P, P, P]),
    "abm_slab_exchange_merge": (I32, [P, P, P, P, P, P, I64, P, P, P, I64, I64, F64, F64, I32, I32, P, P, P, P, P, P, P,
                                      I64, P, P, P, P, P]),
    "abm_ipc_open": (I32, [C.c_char_p, I64, C.POINTER(C.c_void_p)]),
    "abm_peer_exchange": (I32, [P, P, P, P, P, P, P, P, I64, I64, I32, P, P, I64, P, P]),
    "abm_peer_flags": (I32, [P, P, P, P, I32, P, I64, P, P]),
    "abm_deposit_morphogens": (I32, [P, P, F64, I64, P, L3, F64, I64, I64, P, P]),
}


ABI_VERSION = 2


class AbmError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AbmError(
            f"{LIB_PATH} is missing: build it with `python python-hipsc-abm_b200/build.py` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.abm_abi_version() != ABI_VERSION:
        raise AbmError(f"libabm_b200.so ABI version {lib.abm_abi_version()} != {ABI_VERSION}: rebuild it "
                       "(python python-hipsc-abm_b200/build.py)")
    _lib = lib
    return lib


def check(rc: int, what: str) -> int:
    """Raise on negative return codes; positive ones (capacity overflow) are returned to the caller."""
    if rc < 0:
        msg = load().abm_last_error().decode("utf-8", "replace")
        raise AbmError(f"{what} failed with code {rc}: {msg}")
    return rc


SORT_STABLE, SORT_REFERENCE, SORT_PREBINNED, SORT_CHAINED = 1, 2, 4, 8
LOOP_CARRY_OUT, LOOP_CARRY_IN, LOOP_CHAINED, LOOP_FP32, LOOP_NO_PLANE, LOOP_FLAT_Z = 1, 2, 4, 8, 16, 32
SUBSTEP_SORT_ROWS, SUBSTEP_FLAT_Z = 1, 2


def d3(values):
    return (C.c_double * 3)(*[float(v) for v in values])


def i3(values):
    return (C.c_int32 * 3)(*[int(v) for v in values])


def l3(values):
    return (C.c_int64 * 3)(*[int(v) for v in values])
