"""The peer-memory cell exchange primitives (abm_peer_exchange / abm_peer_flags, csrc/slab.cu) on ONE GPU: two emulated
ranks on two streams run the step protocol of slab_device.PeerSlab against each other's buffers (plain device memory
here, CUDA-IPC mappings of a neighbour's memory in production): only the valid records and the header travel, a
buffer is not overwritten before its acknowledgement, and a poll whose flag never comes gives up instead of hanging."""
import numpy as np
import pytest
import torch

from hipsc_abm_b200 import _lib

pytestmark = pytest.mark.gpu

CAP, REC = 1000, 128
DATA, ACK = 0, 2


class Rank:
    def __init__(self, dev):
        u8, i32 = torch.uint8, torch.int32
        self.send = [torch.zeros((CAP + 1, REC), dtype=u8, device=dev) for _ in range(2)]
        self.recv = [torch.zeros((CAP + 1, REC), dtype=u8, device=dev) for _ in range(2)]
        self.flags = torch.zeros(8, dtype=i32, device=dev)
        self.counters = torch.zeros(2, dtype=i32, device=dev)
        self.timed_out = torch.zeros(1, dtype=i32, device=dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.got = []

    def flag(self, k):
        return self.flags[k:].data_ptr()


def message(rng, count, dev):
    m = torch.from_numpy(rng.integers(0, 255, (CAP + 1, REC), dtype=np.uint8)).to(dev)
    m[CAP].zero_()
    m[CAP, :4] = torch.from_numpy(np.frombuffer(np.int32(count).tobytes(), dtype=np.uint8).copy()).to(dev)
    return m


def test_two_emulated_ranks_exchange_over_flag_words():
    lib, dev = _lib.load(), torch.device("cuda:0")
    rng = np.random.default_rng(3)
    a, b = Rank(dev), Rank(dev)            # a: left rank (neighbour on its right = side 1), b: right rank (side 0)
    wait_ns = int(5e9)
    sent = []
    for k in range(1, 4):
        ma, mb = message(rng, 100 * k, dev), message(rng, CAP if k == 2 else 7 * k, dev)
        a.send[1].copy_(ma)
        b.send[0].copy_(mb)
        sent.append((ma.cpu().numpy(), mb.cpu().numpy()))
        torch.cuda.synchronize()
        for me, other, side in ((a, b, 1), (b, a, 0)):
            there = 1 - side
            s = me.stream.cuda_stream
            args = [0, 0, 0, 0, 0, 0, 0, 0]
            args[4 * side:4 * side + 4] = [other.recv[there].data_ptr(), me.send[side].data_ptr(),
                                           other.flag(DATA + there), me.flag(ACK + side)]
            _lib.check(lib.abm_peer_exchange(*args, CAP, REC, k, 0, me.counters.data_ptr(), wait_ns,
                                             me.timed_out.data_ptr(), s), "abm_peer_exchange")
            waits = [0, 0]
            waits[side] = me.flag(DATA + side)
            _lib.check(lib.abm_peer_flags(waits[0], waits[1], 0, 0, k, 0, wait_ns, me.timed_out.data_ptr(), s),
                       "abm_peer_flags")
            with torch.cuda.stream(me.stream):
                me.got.append(me.recv[side].clone())             # "merge": read the buffer ...
            rel = [0, 0]
            rel[side] = other.flag(ACK + there)
            _lib.check(lib.abm_peer_flags(0, 0, rel[0], rel[1], k, 0, wait_ns, me.timed_out.data_ptr(), s),
                       "abm_peer_flags")                          # ... then let the sender overwrite it
    torch.cuda.synchronize()
    assert int(a.timed_out.item()) == 0 and int(b.timed_out.item()) == 0
    for k, (ma, mb) in enumerate(sent):
        for got, want in ((a.got[k].cpu().numpy(), mb), (b.got[k].cpu().numpy(), ma)):
            count = int(np.frombuffer(want[CAP, :4].tobytes(), dtype=np.int32)[0])
            assert np.array_equal(got[:count], want[:count])     # the valid records
            assert np.array_equal(got[CAP], want[CAP])           # the header row (the count)
    assert a.flags[DATA + 1].item() == 3 and b.flags[DATA + 0].item() == 3
    assert a.flags[ACK + 1].item() == 3 and b.flags[ACK + 0].item() == 3


def test_a_poll_gives_up_instead_of_hanging():
    lib, dev = _lib.load(), torch.device("cuda:0")
    r = Rank(dev)
    _lib.check(lib.abm_peer_flags(r.flag(DATA), 0, 0, 0, 5, 0, int(2e6), r.timed_out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), "abm_peer_flags")
    torch.cuda.synchronize()
    assert int(r.timed_out.item()) == 5
