"""Host <-> device coherence of the agent arrays without a copy per public call (SURVEY 7 "hard parts", 8 f-2).

The reference keeps its state in numpy arrays that the rule methods edit freely between the hot-path calls --
in place (`self.motility_forces[index] += ...`, cell_methods.py:330-340) or by replacing the whole array
(`np.delete` / `np.concatenate` on death and division, cell_methods.py:39-44, 86-94).  Re-uploading every array at the
start of every public call is always correct and was what round 1 did (`locations` travelled to the device four
times per step).  This module makes the upload conditional instead:

  * `TrackedArray` is an ndarray VIEW of the caller's array (same memory, no copy) that carries a generation counter
    and bumps it on every write numpy lets a subclass see: item / slice assignment, in-place operators and ufuncs
    with `out=`, `np.copyto` / `np.put` / `np.place` / `np.putmask` / `np.fill_diagonal` / `np.put_along_axis`, and
    the mutating methods (`fill`, `sort`, `put`, `partition`, `itemset`, `setfield`, `byteswap(True)`, `resize`).
    Views of a TrackedArray (`a[i]`, `a[:, :2]`) share its counter.
  * the hot-path methods remember, per array name, which counter and generation the device copy was made from; the
    next call uploads only if the attribute is a different object, the generation moved, or the shape changed.

What numpy does NOT show a subclass: writes through the buffer protocol / `ctypes` / a Numba or C kernel holding
the raw pointer, writes through a plain-ndarray alias of the same memory that the caller kept, `a.flat[...] = `.
Code that does that calls `sim.b200_mark_dirty("locations")` afterwards, or sets `sim.b200_coherence = "always"`
to get the round-1 behaviour (upload at every call) back.
"""
from __future__ import annotations

import numpy as np


class _State:
    """Generation counter shared by a TrackedArray and all its views."""
    __slots__ = ("gen",)

    def __init__(self):
        self.gen = 0


_MUTATING_FUNCTIONS = {np.copyto, np.put, np.place, np.putmask, np.fill_diagonal, np.put_along_axis}


def _plain(x):
    return x.view(np.ndarray) if isinstance(x, TrackedArray) else x


def _plain_all(seq):
    if isinstance(seq, (list, tuple)):
        return type(seq)(_plain_all(v) for v in seq)
    return _plain(seq)


class TrackedArray(np.ndarray):
    _b200 = None

    def __new__(cls, array):
        obj = np.asarray(array).view(cls)
        obj._b200 = _State()
        return obj

    def __array_finalize__(self, obj):
        # a view of a tracked array (a[i], a[:, :2], a.T, a.reshape(..)) shares its counter; anything that owns new
        # memory (copies, arithmetic results) is untracked
        st = getattr(obj, "_b200", None)
        base = self.base
        if st is not None and base is not None and (base is obj or base is getattr(obj, "base", None)):
            self._b200 = st
        else:
            self._b200 = None

    # ---- generation -------------------------------------------------------------------------------------------
    def _touch(self):
        st = self._b200
        if st is not None:
            st.gen += 1

    @property
    def b200_generation(self):
        return None if self._b200 is None else self._b200.gen

    # ---- writes numpy routes through Python ------------------------------------------------------------------------
    def __setitem__(self, key, value):
        self._touch()
        np.ndarray.__setitem__(self, key, _plain(value))

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if out is not None:
            for o in out:
                if isinstance(o, TrackedArray):
                    o._touch()
            kwargs["out"] = tuple(_plain(o) for o in out)
        if method == "at" and inputs and isinstance(inputs[0], TrackedArray):
            inputs[0]._touch()                       # np.add.at(a, idx, v) writes into a
        result = getattr(ufunc, method)(*[_plain(x) for x in inputs], **kwargs)
        if out is not None and result is not NotImplemented:
            # in-place operators must hand back the very object they were applied to (a += 1 rebinds a)
            if isinstance(result, tuple):
                return tuple(o if isinstance(o, TrackedArray) else r for o, r in zip(out, result))
            return out[0] if isinstance(out[0], TrackedArray) else result
        return result

    def __array_function__(self, func, types, args, kwargs):
        if func in _MUTATING_FUNCTIONS and args and isinstance(args[0], TrackedArray):
            args[0]._touch()
        out = kwargs.get("out")
        if out is not None:
            for o in (out if isinstance(out, tuple) else (out,)):
                if isinstance(o, TrackedArray):
                    o._touch()
            kwargs = dict(kwargs, out=_plain_all(out))
        return func(*[_plain_all(a) for a in args], **kwargs)

    def fill(self, value):
        self._touch()
        return np.ndarray.fill(self, value)

    def sort(self, *a, **k):
        self._touch()
        return np.ndarray.sort(self, *a, **k)

    def partition(self, *a, **k):
        self._touch()
        return np.ndarray.partition(self, *a, **k)

    def put(self, *a, **k):
        self._touch()
        return np.ndarray.put(self, *a, **k)

    def itemset(self, *a):                  # removed in numpy 2, kept for the reference's numpy<=1.19 pin
        self._touch()
        return np.ndarray.itemset(self, *a)

    def setfield(self, *a, **k):
        self._touch()
        return np.ndarray.setfield(self, *a, **k)

    def byteswap(self, inplace=False):
        if inplace:
            self._touch()
        return np.ndarray.byteswap(self, inplace)

    def resize(self, *a, **k):
        self._touch()
        return np.ndarray.resize(self, *a, **k)

    # ---- pickling: a checkpoint (Simulation.temp, reference simulation.py:290-298) holds plain arrays ---------------
    def __reduce__(self):
        return np.ndarray.__reduce__(self.view(np.ndarray))

    def __deepcopy__(self, memo):
        return np.array(self, copy=True).view(np.ndarray)


def track(array) -> TrackedArray:
    """`array` itself if it already carries a counter, else a tracked view of the same memory."""
    if isinstance(array, TrackedArray) and array._b200 is not None:
        return array
    return TrackedArray(array)


class DeviceCopy:
    """What the device copy of one named host array was made from."""
    __slots__ = ("state", "gen", "shape", "tensor")

    def __init__(self, tracked: TrackedArray, tensor):
        self.state, self.gen, self.shape, self.tensor = tracked._b200, tracked._b200.gen, tracked.shape, tensor

    def current_for(self, array) -> bool:
        return (isinstance(array, TrackedArray) and array._b200 is self.state and self.state is not None and
                self.state.gen == self.gen and array.shape == self.shape)

    def refresh(self):
        """The host array was just rewritten FROM the device copy (a download into it): still coherent."""
        self.gen = self.state.gen
