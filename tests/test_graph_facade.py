"""DeviceGraph (the igraph replacement) on its host side, against the oracle's independent HostGraph."""
import pickle

import numpy as np

from hipsc_abm_b200.graph import DeviceGraph
from oracle.oracle import HostGraph


def same(a: DeviceGraph, b: HostGraph):
    assert a.vcount() == b.vcount() and a.ecount() == b.ecount()
    assert a.get_edgelist() == b.get_edgelist()
    for i in range(a.vcount()):
        assert a.neighbors(i) == b.neighbors(i)
        assert a.num_neighbors(i) == b.num_neighbors(i)


def test_random_operation_sequences():
    rng = np.random.default_rng(0)
    for trial in range(20):
        n = int(rng.integers(1, 40))
        a, b = DeviceGraph(n), HostGraph(n)
        for op in range(30):
            kind = rng.integers(0, 6)
            if kind <= 1 and a.vcount() > 0:
                es = rng.integers(0, a.vcount(), (int(rng.integers(0, 25)), 2))
                a.add_edges(es)
                a.simplify()
                b.add_edges(es)
                b.simplify()
            elif kind == 2 and a.ecount() > 0:
                idx = np.unique(rng.integers(0, a.ecount(), int(rng.integers(1, 5))))
                a.delete_edges(idx)
                b.delete_edges(idx)
            elif kind == 3 and a.vcount() > 1:
                idx = np.unique(rng.integers(0, a.vcount(), int(rng.integers(1, 4))))
                a.delete_vertices(idx)
                b.delete_vertices(idx)
            elif kind == 4:
                a.add_vertex()
                b.add_vertex()
            else:
                if rng.random() < 0.1:
                    a.delete_edges(None)
                    b.delete_edges(None)
            same(a, b)


def test_sticky_capacities_and_pickle():
    g = DeviceGraph(5)
    assert g.max_neighbors == 1 and g.max_agents == 1          # reference backend.py:22-23
    g.add_edges([[0, 1], [1, 2], [4, 0]])
    g.max_agents = 12
    h = pickle.loads(pickle.dumps(g, -1))
    assert h.get_edgelist() == [(0, 1), (0, 4), (1, 2)] and h.max_agents == 12 and h.vcount() == 5
    assert h.neighbors(0) == [1, 4]


def test_empty_and_ragged():
    g = DeviceGraph(0)
    assert g.get_edgelist() == [] and g.ecount() == 0
    g.add_vertex()
    assert g.neighbors(0) == [] and g.num_neighbors(0) == 0
    g.add_edges(np.zeros((0, 2), dtype=int))
    g.add_edges([[0, 0]])                                      # self loop is dropped (simplify)
    assert g.ecount() == 0
    g.delete_vertices([])
    assert g.vcount() == 1
