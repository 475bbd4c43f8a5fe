"""arrays.TrackedArray: every write numpy routes through Python moves the generation counter, reads do not, views share
the counter, results of arithmetic are plain arrays; and B200Simulation uploads an array only when it changed."""
import copy
import pickle

import numpy as np
import pytest

import hipsc_abm_b200  # noqa: F401
from hipsc_abm_b200.arrays import DeviceCopy, TrackedArray, track


def gen(a):
    return a.b200_generation


def test_reads_do_not_touch_and_results_are_plain():
    base = np.arange(12.0).reshape(4, 3)
    a = TrackedArray(base)
    assert a.base is base and np.shares_memory(a, base)          # a view: no copy
    g0 = gen(a)
    _ = a[1], a[:, 1], a.sum(), a + 1, np.linalg.norm(a, axis=1), a * a, np.delete(a, 1, axis=0), a.copy(), a.T @ a
    _ = np.concatenate([a, a]), a[a[:, 0] > 3], float(a[2, 1]), a.astype(np.float32), np.asarray(a).max()
    assert gen(a) == g0
    assert type(a + 1) is np.ndarray and type(np.delete(a, 0, axis=0)) is np.ndarray
    assert type(np.concatenate([a, a])) is np.ndarray and type(a.copy().view(np.ndarray)) is np.ndarray
    assert a.copy().b200_generation is None                       # owns new memory: untracked


@pytest.mark.parametrize("write", [
    lambda a: a.__setitem__((1, 2), 5.0),
    lambda a: a.__setitem__(slice(None), 0),
    lambda a: a.__setitem__(a[:, 0] > 1, -1.0),
    lambda a: a.__iadd__(1.0),
    lambda a: a.__imul__(2.0),
    lambda a: np.add(a, 1, out=a),
    lambda a: np.multiply(a, 2, out=(a,)),
    lambda a: np.add.at(a, ([0, 0], [1, 1]), 1.0),
    lambda a: np.copyto(a, 3.0),
    lambda a: np.put(a, [0, 1], [9.0, 9.0]),
    lambda a: np.place(a, a > 2, [0.5]),
    lambda a: np.putmask(a, a > 2, 0.25),
    lambda a: a.fill(7.0),
    lambda a: a.sort(axis=0),
    lambda a: a.put([0], [1.0]),
    lambda a: np.clip(a, 0, 1, out=a),
    lambda a: a[2].__setitem__(0, 4.0),                 # through a row view
    lambda a: a[:, :2].__iadd__(1.0),                   # through a column-block view
    lambda a: a.reshape(-1).__setitem__(3, 1.0),        # through a reshaped view
    lambda a: a[1:].__getitem__(0).fill(2.0),           # a view of a view
])
def test_every_python_visible_write_moves_the_generation(write):
    base = np.arange(12.0).reshape(4, 3)
    a = TrackedArray(base)
    g0 = gen(a)
    write(a)
    assert gen(a) > g0


def test_inplace_operators_return_the_same_object_and_the_right_values():
    a = TrackedArray(np.ones((3, 3)))
    b = a
    a += 2
    a *= 3
    a[1] -= 1
    assert a is b and isinstance(a, TrackedArray)
    assert np.array_equal(np.asarray(a), [[9, 9, 9], [8, 8, 8], [9, 9, 9]])


def test_pickle_and_deepcopy_give_plain_arrays():
    a = TrackedArray(np.arange(6.0).reshape(2, 3))
    for b in (pickle.loads(pickle.dumps(a, protocol=-1)), copy.deepcopy(a)):
        assert type(b) is np.ndarray and np.array_equal(b, np.asarray(a))


def test_device_copy_currency():
    a = TrackedArray(np.zeros((5, 3)))
    dc = DeviceCopy(a, tensor="T")
    assert dc.current_for(a) and dc.current_for(a[:])          # same counter, same shape
    assert not dc.current_for(np.zeros((5, 3)))                # another (plain) object
    assert not dc.current_for(TrackedArray(np.zeros((5, 3))))  # another tracked object
    a[0, 0] = 1
    assert not dc.current_for(a)
    dc.refresh()
    assert dc.current_for(a)
    assert track(a) is a and track(np.zeros(3)).b200_generation == 0


class FakeTensor:
    def __init__(self, name, shape):
        self.name, self.shape = name, tuple(shape)


class FakeEngine:
    """Stands in for ColonyEngine: counts what travels."""

    def __init__(self):
        self.sent, self.n = [], 0
        self.loc = self.radii = self.mot = self.jkr = None

    def use_current_stream(self):
        pass

    def to_device(self, array, name=None):
        self.sent.append((name, np.array(array, copy=True)))
        return FakeTensor(name, np.shape(array))

    def adopt(self, n, loc=None, radii=None, radii_host=None, mot=None, jkr=None):
        self.n = n
        for k, v in (("loc", loc), ("radii", radii), ("mot", mot), ("jkr", jkr)):
            if v is not None:
                setattr(self, k, v)

    def _zeros(self, shape, dtype):
        return FakeTensor("zeros", shape)


def make_sim(n=6):
    from hipsc_abm_b200.simulation import ColonySimulation
    sim = ColonySimulation([100.0, 100.0, 0.0])
    sim.add_agents(n)
    sim.agent_array("locations", override=np.random.default_rng(0).random((n, 3)))
    sim.agent_array("radii", func=lambda: 5.0)
    sim.agent_array("motility_forces", vector=3)
    sim.__dict__["_b200_engine"] = FakeEngine()
    return sim


def names(sim):
    out = [s[0] for s in sim.b200_engine().sent]
    sim.b200_engine().sent.clear()
    return out


def test_unchanged_arrays_are_not_uploaded_again():
    sim = make_sim()
    sim.b200_upload()
    assert names(sim) == ["locations", "radii"]
    sim.b200_upload()
    assert names(sim) == []                                     # nothing changed: nothing travels
    sim.b200_upload(with_forces="motility")
    assert names(sim) == ["motility_forces"]
    sim.motility_forces[2] += 1e-9                              # an in-place edit by a rule method
    sim.locations[0, 0] = 3.0
    sim.b200_upload(with_forces="motility")
    assert sorted(names(sim)) == ["locations", "motility_forces"]
    assert sim.b200_engine().sent == [] and sim.b200_engine().loc.name == "locations"


def test_replaced_arrays_and_dirty_marks_are_uploaded():
    sim = make_sim()
    sim.b200_upload()
    names(sim)
    # division / death replace the arrays wholesale (reference cell_methods.py:39-44, 86-94)
    sim.number_agents += 1
    sim.locations = np.concatenate([sim.locations, [[1.0, 2.0, 0.0]]])
    sim.radii = np.concatenate([sim.radii, [5.0]])
    sim.motility_forces = np.concatenate([sim.motility_forces, np.zeros((1, 3))])
    sim.b200_upload()
    assert names(sim) == ["locations", "radii"] and sim.b200_engine().n == 7
    sim.b200_mark_dirty("radii")                                # e.g. after a raw-pointer write
    sim.b200_upload()
    assert names(sim) == ["radii"]
    sim.b200_coherence = "always"                               # the round-1 behaviour on request
    sim.b200_upload()
    assert names(sim) == ["locations", "radii"]


def test_uploaded_values_are_the_current_host_values():
    sim = make_sim()
    sim.b200_upload()
    sim.locations[3] = [7.0, 8.0, 0.0]
    sim.b200_upload()
    sent = dict(sim.b200_engine().sent)
    assert np.array_equal(sent["locations"][3], [7.0, 8.0, 0.0])


def test_wrong_length_array_raises():
    sim = make_sim()
    sim.number_agents = 9
    with pytest.raises(ValueError):
        sim.b200_upload()
