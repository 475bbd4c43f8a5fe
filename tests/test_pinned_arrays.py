"""Host logic of engine.PinnedArrays (in-place page-locking of long-lived caller arrays) with a fake CUDA runtime:
when an array is registered, when it is not, and that registrations are released."""
import gc

import numpy as np
import pytest

from hipsc_abm_b200 import engine


class FakeRuntime:
    def __init__(self):
        self.live = {}          # ptr -> nbytes
        self.calls = []
        self.fail_next = False

    def cudaHostRegister(self, ptr, nbytes, flags):
        self.calls.append(("reg", ptr, nbytes))
        if self.fail_next:
            self.fail_next = False
            return 1
        for p, nb in self.live.items():
            assert not (p < ptr + nbytes and ptr < p + nb), "registered an overlapping range"
        self.live[ptr] = nbytes
        return 0

    def cudaHostUnregister(self, ptr):
        self.calls.append(("unreg", ptr))
        del self.live[ptr]
        return 0


@pytest.fixture
def rt(monkeypatch):
    fake = FakeRuntime()
    monkeypatch.setattr(engine.torch.cuda, "cudart", lambda: fake)
    return fake


def big(n=1 << 18):
    return np.zeros(n, dtype=np.float64)       # 2 MiB


def test_registered_on_the_second_sighting_only(rt):
    pins = engine.PinnedArrays()
    a = big()
    assert not pins.consider(a) and not rt.live                 # first sighting: staging path
    assert pins.consider(a) and rt.live == {a.ctypes.data: a.nbytes}
    assert pins.consider(a) and len(rt.calls) == 1              # stays registered, no second call
    assert pins.is_registered(a) and pins.bytes == a.nbytes


def test_small_readonly_strided_and_foreign_arrays_are_left_alone(rt):
    pins = engine.PinnedArrays()
    small = np.zeros(100)
    ro = big()
    ro.flags.writeable = False
    strided = big()[::2]
    foreign = np.frombuffer(bytearray(4 << 20), dtype=np.float64)   # memory owned by a non-numpy object
    for arr in (small, ro, strided, foreign):
        assert not pins.consider(arr) and not pins.consider(arr)
    assert not rt.live


def test_view_of_a_field_plane_shares_the_registration(rt):
    pins = engine.PinnedArrays()
    field = np.zeros((600, 600, 1))
    plane = np.ascontiguousarray(field[:, :, 0])
    assert plane.ctypes.data == field.ctypes.data               # no copy was made
    pins.consider(field)
    assert pins.consider(plane)                                 # second sighting of the same (address, size)
    assert pins.consider(field) and len(rt.live) == 1


def test_released_when_the_owner_dies_and_replaced_on_address_reuse(rt):
    pins = engine.PinnedArrays()
    a = big()
    pins.consider(a)
    pins.consider(a)
    ptr = a.ctypes.data
    del a
    gc.collect()
    assert not rt.live and ("unreg", ptr) in rt.calls and pins.bytes == 0
    # a stale registration (owner still alive elsewhere) that a new array overlaps is dropped first
    keep = big(1 << 19)
    pins.consider(keep)
    pins.consider(keep)
    half = keep[: 1 << 18]                                     # same start address, half the size
    pins.consider(half)
    assert pins.consider(half)
    assert rt.live == {half.ctypes.data: half.nbytes}


def test_a_refused_registration_is_not_retried(rt):
    pins = engine.PinnedArrays()
    a = big()
    pins.consider(a)
    rt.fail_next = True
    assert not pins.consider(a)
    assert not pins.consider(a) and len([c for c in rt.calls if c[0] == "reg"]) == 1
