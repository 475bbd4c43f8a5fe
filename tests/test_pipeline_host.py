"""Host logic of pipeline.HotPathStep that needs no GPU: which motility forces and bond-row buffers each step uses
(a stand-in engine records the calls)."""
import pytest
import torch

from hipsc_abm_b200.engine import GraphTensors
from hipsc_abm_b200.pipeline import HotPathStep, StepConfig


class RecordingEngine:
    """The slice of ColonyEngine that HotPathStep.enqueue() touches, on CPU tensors."""

    def __init__(self, n):
        self.n = n
        self.device = torch.device("cpu")
        self.max_radius, self.uniform_radius = 5.0, 5.0
        self.launches = 0
        self.loc = torch.zeros((n, 3), dtype=torch.float64)
        self.mot = torch.zeros((n, 3), dtype=torch.float64)
        self._carry = None
        self._bufs = None
        self.calls = []

    def use_current_stream(self):
        return 0

    def _bond_buffers(self, n, stride):
        if self._bufs is None:
            self._bufs = dict(len=[torch.zeros(n, dtype=torch.int32) for _ in range(3)],
                              cols=[torch.zeros(n * stride, dtype=torch.int32) for _ in range(3)],
                              max_len=torch.zeros(2, dtype=torch.int32))
        return self._bufs

    def sort(self, distance, loc=None):
        self.launches += 4
        return ("sort", distance)

    def enqueue_search(self, srt, distance, old, row_start, row_len, cols, cursor, loc=None):
        self.launches += 1
        return GraphTensors(self.n, row_start, 0, row_len, cols, True)

    def enqueue_mechanics(self, old, distance, dts, stride, stokes, pair=None, zero_flag=True, carry=False):
        b = self._bond_buffers(self.n, stride)
        src = [i for i in range(3) if b["cols"][i] is old.cols][0]
        self.calls.append(dict(mot=self.mot, src=src, pair=tuple(pair), carry=carry))
        self.launches += 3 * len(dts)
        last = pair[(len(dts) - 1) & 1]
        return GraphTensors(self.n, None, stride, b["len"][last], b["cols"][last], True)


def test_steps_alternate_bond_buffers_and_motility_forces():
    n = 10
    eng = RecordingEngine(n)
    a = torch.ones((n, 3), dtype=torch.float64)
    pipe = HotPathStep(eng, StepConfig(), motility_by_parity=(a, -a))
    for _ in range(4):
        pipe.step()
    assert [c["mot"] is a for c in eng.calls] == [True, False, True, False]
    assert [c["src"] for c in eng.calls] == [0, 1, 0, 1]            # the bonds ping-pong between row buffers 0 and 1
    assert all(c["pair"] == (1 - c["src"], 2) and c["carry"] for c in eng.calls)
    assert pipe.parity == 0 and pipe.launches_per_step == 5 + 33


def test_without_a_pair_the_engine_forces_are_used():
    eng = RecordingEngine(4)
    pipe = HotPathStep(eng, StepConfig())
    pipe.step()
    pipe.step()
    assert all(c["mot"] is eng.mot for c in eng.calls)


def test_motility_pair_must_match_the_colony():
    eng = RecordingEngine(4)
    with pytest.raises(ValueError):
        HotPathStep(eng, StepConfig(), motility_by_parity=(torch.zeros((3, 3)), torch.zeros((3, 3))))
