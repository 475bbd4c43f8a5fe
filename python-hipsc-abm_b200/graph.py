"""DeviceGraph -- drop-in for the reference's igraph-backed Graph (reference backend.py:13-28).

It keeps the 8 methods the reference calls: add_edges, delete_edges, simplify,
get_edgelist, neighbors, num_neighbors, add_vertex, delete_vertices -- plus the sticky public
capacities max_neighbors / max_agents.  The authoritative copy is either the device row CSR produced
by the search / mechanics kernels, or a host CSR (ptr, cols) when the rule methods have mutated it;
conversions happen lazily and only at those phase boundaries.  The graph is always simple
(undirected, no loops, no multi-edges), i.e. what igraph holds after simplify().
"""
from __future__ import annotations

import numpy as np


class DeviceGraph:
    def __init__(self, n=0, engine=None):
        self.n = int(n)
        self.max_neighbors = 1      # reference backend.py:22
        self.max_agents = 1         # reference backend.py:23
        self._engine = engine
        self._dev = None            # engine.GraphTensors or None
        self._ptr = np.zeros(self.n + 1, dtype=np.int64)   # host CSR, ascending rows
        self._cols = np.zeros(0, dtype=np.int32)
        self._host_valid = True
        self.row_stride_hint = 16

    # ---- device side ---------------------------------------------------------------------------------
    def set_device(self, tensors, engine=None):
        """Adopt a kernel result as the authoritative copy."""
        if engine is not None:
            self._engine = engine
        self._dev = tensors
        self.n = tensors.n
        self._host_valid = False

    def device_tensors(self, engine):
        """Row CSR on the device (uploading the host copy if that is the newer one)."""
        import torch
        from .engine import GraphTensors
        self._engine = engine
        if self._dev is None or self._dev.n != self.n:
            ptr, cols = self._host()
            dev = engine.device
            row_start = engine.to_device(ptr[:-1].astype(np.int32))
            row_len = engine.to_device(np.diff(ptr).astype(np.int32))
            if self.n == 0:
                row_start = torch.zeros(1, dtype=torch.int32, device=dev)
                row_len = torch.zeros(1, dtype=torch.int32, device=dev)
            cols_t = engine.to_device(np.ascontiguousarray(cols, dtype=np.int32))
            if cols_t.numel() == 0:
                cols_t = torch.zeros(1, dtype=torch.int32, device=dev)
            self._dev = GraphTensors(self.n, row_start, 0, row_len, cols_t, True)
        return self._dev

    def _host(self):
        """Host CSR (ptr int64[n+1], cols int32), downloading the device copy if needed."""
        if not self._host_valid:
            g = self._dev
            dl = self._engine.host if self._engine is not None else (lambda t: t.cpu().numpy())
            row_len = dl(g.row_len[:g.n]).astype(np.int64)
            ptr = np.zeros(g.n + 1, dtype=np.int64)
            np.cumsum(row_len, out=ptr[1:])
            total = int(ptr[-1])
            if g.row_start is None:
                start = np.arange(g.n, dtype=np.int64) * g.row_stride
                flat = g.cols.reshape(-1)
            else:
                start = dl(g.row_start[:g.n]).astype(np.int64)
                flat = g.cols
            if total:
                # position of every entry inside its row + the row's start in the device layout
                src = np.repeat(start - ptr[:-1], row_len) + np.arange(total, dtype=np.int64)
                if g.row_start is None:
                    cols = dl(flat)[src]
                else:
                    hi = int(src.max()) + 1
                    cols = dl(flat[:hi])[src]
            else:
                cols = np.zeros(0, dtype=np.int32)
            if not g.sorted_rows and total:
                rows = np.repeat(np.arange(g.n, dtype=np.int64), row_len)
                order = np.lexsort((cols, rows))
                cols = cols[order]
            self._ptr, self._cols = ptr, np.ascontiguousarray(cols, dtype=np.int32)
            self._host_valid = True
        return self._ptr, self._cols

    def _set_host(self, ptr, cols):
        self._ptr, self._cols = ptr, np.ascontiguousarray(cols, dtype=np.int32)
        self._host_valid = True
        self._dev = None

    @staticmethod
    def _csr_from_edges(n, edges):
        """Symmetric ascending CSR of the simple graph with the given (i,j) rows."""
        e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        e = e[e[:, 0] != e[:, 1]]
        if len(e) == 0:
            return np.zeros(n + 1, dtype=np.int64), np.zeros(0, dtype=np.int32)
        a = np.concatenate([e[:, 0], e[:, 1]])
        b = np.concatenate([e[:, 1], e[:, 0]])
        key = np.unique(a * n + b)          # sorts by (row, col) and removes duplicates
        rows, cols = key // n, key % n
        ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n), out=ptr[1:])
        return ptr, cols.astype(np.int32)

    # ---- the igraph surface ----------------------------------------------------------------------------
    def vcount(self):
        return self.n

    def ecount(self):
        return int(self._host()[0][-1] // 2)

    def edge_array(self) -> np.ndarray:
        """(E,2) int64 rows (min,max), sorted -- the order of get_edgelist()."""
        ptr, cols = self._host()
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(ptr))
        keep = cols > rows
        return np.stack([rows[keep], cols[keep].astype(np.int64)], axis=1)

    def get_edgelist(self):
        return [tuple(r) for r in self.edge_array().tolist()]

    def neighbors(self, index):
        ptr, cols = self._host()
        return cols[ptr[index]:ptr[index + 1]].tolist()

    def num_neighbors(self, index):
        ptr, _ = self._host()
        return int(ptr[index + 1] - ptr[index])

    def degrees(self) -> np.ndarray:
        return np.diff(self._host()[0])

    def add_edges(self, edges):
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        if len(edges):
            self._set_host(*self._csr_from_edges(self.n, np.concatenate([self.edge_array(), edges])))

    def delete_edges(self, which=None):
        if which is None:
            self._set_host(np.zeros(self.n + 1, dtype=np.int64), np.zeros(0, dtype=np.int32))
            return
        which = np.asarray(which, dtype=np.int64).reshape(-1)
        if len(which):
            e = self.edge_array()
            keep = np.ones(len(e), dtype=bool)
            keep[which] = False
            self._set_host(*self._csr_from_edges(self.n, e[keep]))

    def simplify(self):
        pass    # always simple

    def add_vertex(self):
        self.add_vertices(1)

    def add_vertices(self, k):
        ptr, cols = self._host()
        self.n += int(k)
        self._set_host(np.concatenate([ptr, np.full(int(k), ptr[-1], dtype=np.int64)]), cols)

    def delete_vertices(self, indices):
        """Remove vertices and renumber the rest (igraph semantics; reference cell_methods.py:47-48)."""
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        if len(idx) == 0:
            return
        ptr, cols = self._host()
        dead = np.zeros(self.n, dtype=bool)
        dead[idx] = True
        renumber = np.cumsum(~dead) - 1
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(ptr))
        keep = ~(dead[rows] | dead[cols])
        new_n = self.n - int(dead.sum())
        new_rows = renumber[rows[keep]]
        new_ptr = np.zeros(new_n + 1, dtype=np.int64)
        np.cumsum(np.bincount(new_rows, minlength=new_n), out=new_ptr[1:])
        self.n = new_n
        self._set_host(new_ptr, renumber[cols[keep]].astype(np.int32))   # monotone renumbering keeps rows ascending

    # ---- pickling (Simulation.temp pickles the whole object, reference simulation.py:290-298) -------------
    def __getstate__(self):
        ptr, cols = self._host()
        return dict(n=self.n, max_neighbors=self.max_neighbors, max_agents=self.max_agents, ptr=ptr, cols=cols,
                    row_stride_hint=self.row_stride_hint)

    def __setstate__(self, st):
        self.n = st["n"]
        self.max_neighbors, self.max_agents = st["max_neighbors"], st["max_agents"]
        self.row_stride_hint = st.get("row_stride_hint", 16)
        self._engine = None
        self._dev = None
        self._ptr, self._cols = st["ptr"], st["cols"]
        self._host_valid = True
