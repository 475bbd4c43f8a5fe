"""Device-resident state and stage launches of the cell-update hot path.

`ColonyEngine` owns the torch CUDA tensors (torch is used for memory and streams only) and calls the
kernels of libabm_b200.so through the C-ABI.  The reference-facing classes in simulation.py /
cell_methods.py sit on top of it; bench.py drives it directly for the HBM-resident measurement.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib

# physical constants hard-coded in the reference methods
JKR_ADHESION = 0.000107    # reference cell_methods.py:347
JKR_POISSON = 0.5          # reference cell_methods.py:348
JKR_YOUNGS = 1000.0        # reference cell_methods.py:349
STOKES = 10000.0           # reference cell_methods.py:392


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


class PinnedArrays:
    """Page-locks long-lived numpy arrays of the caller IN PLACE (cudaHostRegister), so that the copies between
    them and the device are plain DMA transfers instead of going through a pinned staging buffer and a host memcpy.

    The reference keeps its agent arrays and morphogen fields as attributes that are edited in place step after
    step (`self.motility_forces[:, :] = 0`, `self.fgf4_values[:, :, 0] = ...`): an array that shows up at the same
    address with the same size twice in a row is taken to be such an array and registered.  Arrays that the rule
    methods replace every step (np.concatenate / np.delete on division and death) never reach the second sighting
    and keep using the staging path.  A registration is dropped when the owning numpy object dies (weakref) or
    when another array turns up at an overlapping address."""
    MIN_BYTES = 1 << 20

    def __init__(self):
        self.sightings = {}     # (ptr, nbytes) -> count
        self.registered = {}    # ptr -> (nbytes, finalizer)
        self.refused = set()
        self.bytes = 0

    @staticmethod
    def _owner(a: np.ndarray):
        while isinstance(a.base, np.ndarray):
            a = a.base
        return a if a.base is None else None            # memory owned by something we cannot watch: leave alone

    def _unregister(self, ptr):
        entry = self.registered.pop(ptr, None)
        if entry is not None:
            entry[1].detach()
            torch.cuda.cudart().cudaHostUnregister(ptr)
            self.bytes -= entry[0]

    def is_registered(self, a: np.ndarray) -> bool:
        e = self.registered.get(a.ctypes.data)
        return e is not None and e[0] == a.nbytes

    def consider(self, a: np.ndarray) -> bool:
        """True if `a` is (now) page-locked in place."""
        if a.nbytes < self.MIN_BYTES or not a.flags.c_contiguous or not a.flags.writeable:
            return False
        ptr, key = a.ctypes.data, (a.ctypes.data, a.nbytes)
        if self.is_registered(a):
            return True
        if key in self.refused:
            return False
        seen = self.sightings.get(key, 0) + 1
        if len(self.sightings) > 256:
            self.sightings.clear()
        self.sightings[key] = seen
        if seen < 2:
            return False
        owner = self._owner(a)
        if owner is None:
            self.refused.add(key)
            return False
        for p, (nb, _) in list(self.registered.items()):      # a stale registration at an overlapping address
            if p < ptr + a.nbytes and ptr < p + nb:
                self._unregister(p)
        import weakref
        err = torch.cuda.cudart().cudaHostRegister(ptr, a.nbytes, 0)
        if int(err) != 0:
            self.refused.add(key)
            return False
        self.registered[ptr] = (a.nbytes, weakref.finalize(owner, self._owner_died, ptr))
        self.bytes += a.nbytes
        return True

    def _owner_died(self, ptr):
        entry = self.registered.pop(ptr, None)
        if entry is not None:
            try:
                torch.cuda.cudart().cudaHostUnregister(ptr)
            except Exception:
                pass
            self.bytes -= entry[0]


@dataclass
class SortState:
    """Result of one counting sort (abm_bin_sort)."""
    distance: float
    dims: tuple
    nbins: int
    n: int
    bin_start: torch.Tensor
    sorted_id: torch.Tensor
    sorted_pos: torch.Tensor
    sorted_bin: torch.Tensor
    max_count: torch.Tensor


@dataclass
class GraphTensors:
    """A row-CSR graph on the device (include/abm_b200.h, 'row CSR')."""
    n: int
    row_start: Optional[torch.Tensor]   # None -> fixed-stride rows
    row_stride: int
    row_len: torch.Tensor
    cols: torch.Tensor
    sorted_rows: bool = True

    def clone(self) -> "GraphTensors":
        return GraphTensors(self.n, None if self.row_start is None else self.row_start.clone(), self.row_stride,
                            self.row_len.clone(), self.cols.clone(), self.sorted_rows)


class _LibOnDevice:
    """The C-ABI library with every call issued while `device` is the current CUDA device: the kernels are launched
    with <<<>>> on the engine's stream, which is only valid on the device the stream belongs to.  (An engine on
    cuda:1 in a process whose current device is cuda:0 would otherwise launch into the wrong context.)"""

    def __init__(self, lib, device):
        self._lib, self._index = lib, device.index if device.index is not None else torch.cuda.current_device()
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw, index = getattr(self._lib, name), self._index

            def fn(*args):
                if torch.cuda.current_device() == index:
                    return raw(*args)
                with torch.cuda.device(index):
                    return raw(*args)
            self._cache[name] = fn
        return fn


class ColonyEngine:
    def __init__(self, size, device=None):
        if not torch.cuda.is_available():
            raise _lib.AbmError("ColonyEngine needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _LibOnDevice(_lib.load(), self.device)
        self.size = np.asarray(size, dtype=np.float64).reshape(3)     # the [0,size] box positions are clamped to
        self._size_c = _lib.d3(self.size)
        self.set_bin_lattice(None, None)
        self.n = 0
        self.loc = self.radii = self.mot = self.jkr = None
        self._max_radius = 0.0      # None: not evaluated yet for the current radii (done on the device when needed)
        self._uniform_radius = 0.0  # common radius when all radii are equal (contact constants evaluated once)
        self._radii_host = None     # the numpy array the radii were last uploaded from (None: device-resident state)
        self.global_id = None   # int32 (n,) when bond rows are kept by global id (slab decomposition)
        self.n_device = None    # int32 (1,) tensor: the cell count lives on the device, self.n is only the capacity
                                # the launches and buffers are sized for (include/abm_b200.h, `n_device`)
        self._sort_wss = {}     # sort scratch per bin count (sorts of different lattices may run on different streams)
        self._carry = None      # (nbins, n, pointers..., table index) when the last mechanics loop handed its cells over
        self._sort_bufs = {}
        self._edge_ws = None
        self._bond_bufs = None
        self.cols_capacity_hint = {}
        self.launches = 0      # number of kernels of this library enqueued (bench.py's gpu_launches)
        self.stream = 0
        # per-engine launch options (arguments of the calls they affect; the library has no process-wide switches)
        self.chained_launch = False     # abm_mechanics_loop(ABM_LOOP_CHAINED): programmatic dependent launches
        self.move_window = None         # device (2,) float64: only cells with x outside it count for max_substep_move()
        self.mechanics_dtype = "f64"    # "f32": abm_mechanics_loop(ABM_LOOP_FP32), bin-relative single precision
        self.substep_events = None      # list of 2 x (substeps) recorded torch.cuda.Event(enable_timing=True): the next
                                        # mechanics loop records them around its k_jkr_substep launches (bench.py)
        self.flat_z_vouched = False     # abm_mechanics_loop(ABM_LOOP_FLAT_Z): every z is known to be 0 at the next loop's entry
        self.plane_kernel = True        # False: abm_mechanics_loop(ABM_LOOP_NO_PLANE), the three-axis kernel in a flat world too
        self.diffusion_fuse = int(os.environ.get("ABM_B200_DIFFUSION_FUSE", "0"))   # substeps per launch, 0 = 5
        self.use_current_stream()
        self.h2d_bytes = 0     # bytes moved by upload()/to_device() and host()/locations_host()
        self.d2h_bytes = 0
        self._pinned = {}
        self._handed_out = weakref.WeakValueDictionary()   # data pointer -> pinned tensor backing a numpy array
        self.pin_in_place = PinnedArrays()                 # long-lived caller arrays, page-locked where they are

    # ------------------------------------------------------------------------------------------------
    def use_current_stream(self):
        """Bind the engine to torch's current stream (looked up once per step / capture, not per launch)."""
        self.stream = torch.cuda.current_stream(self.device).cuda_stream
        return self.stream

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    # ---- state -------------------------------------------------------------------------------------
    def to_device(self, array, name=None) -> torch.Tensor:
        """numpy -> device.  The host arrays of the reference API are pageable numpy, so the copy is staged
        through a reusable pinned buffer (asynchronous DMA at full PCIe rate); arrays this engine handed out
        itself (host()) already live in pinned memory and are sent as they are."""
        a = np.ascontiguousarray(array)
        self.h2d_bytes += a.nbytes
        if a.size == 0:
            return torch.from_numpy(a.copy()).to(self.device)
        backing = self._handed_out.get(a.ctypes.data)
        if backing is not None and backing.numel() * backing.element_size() == a.nbytes and \
                backing.dtype == torch.from_numpy(a[:0].copy()).dtype:
            return backing.view(a.shape).to(self.device, non_blocking=True)
        if self.pin_in_place.consider(a):
            return torch.from_numpy(a).to(self.device, non_blocking=True)      # DMA straight out of the caller's array
        t = torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())
        if name is None:
            return t.to(self.device)
        stage, busy = self._pinned.get(name, (None, None))
        if stage is None or stage.dtype != t.dtype or stage.numel() < t.numel():
            stage = torch.empty(max(int(1.1 * t.numel()), 1), dtype=t.dtype).pin_memory()
            busy = None
        if busy is not None:
            busy.synchronize()                # the previous DMA out of this staging buffer has finished
        view = stage[:t.numel()].view(t.shape)
        view.copy_(t)
        out = view.to(self.device, non_blocking=True)
        busy = torch.cuda.Event()
        busy.record(torch.cuda.current_stream(self.device))
        self._pinned[name] = (stage, busy)
        return out

    def host(self, t: torch.Tensor, sync=True) -> np.ndarray:
        """device -> numpy (counted).  The array is backed by pinned memory, so handing it back to to_device()
        later needs no staging copy.  sync=False: the copy is only enqueued -- the caller synchronises the stream
        before anyone reads the array."""
        self.d2h_bytes += t.numel() * t.element_size()
        if t.numel() == 0:
            return t.cpu().numpy()
        dst = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        dst.copy_(t, non_blocking=True)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        arr = dst.numpy()
        self._handed_out[arr.ctypes.data] = dst
        return arr

    def host_into(self, t: torch.Tensor, dst: np.ndarray):
        """device -> an existing numpy array of the caller (counted).  Straight DMA if the array is page-locked in
        place, else through a pinned buffer and one host copy."""
        if dst.shape != tuple(t.shape):
            raise ValueError("host_into: shape mismatch")
        if dst.flags.c_contiguous and dst.dtype == np.dtype(str(t.dtype).replace("torch.", "")) and \
                self.pin_in_place.consider(dst):
            self.d2h_bytes += t.numel() * t.element_size()
            torch.from_numpy(dst).copy_(t, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        else:
            dst[...] = self.host(t)

    def upload(self, locations, radii, motility=None, jkr=None):
        """Host -> device copy of the agent arrays the path reads (reference dtypes: float64)."""
        self.use_current_stream()
        loc = np.ascontiguousarray(locations, dtype=np.float64)
        n = loc.shape[0]
        rad = np.asarray(radii, dtype=np.float64)
        if rad.shape != (n,) or not rad.flags.c_contiguous:      # a scalar or a strided view: materialise it
            rad = np.ascontiguousarray(np.broadcast_to(rad, (n,)))
        self.n = n
        self._carry = None
        self.loc = self.to_device(loc, "loc")
        self.radii = self.to_device(rad, "radii")
        self._radii_host = rad                # radius statistics come from here: no read-back that waits for the DMA
        self._max_radius = self._uniform_radius = None if n else 0.0    # only the mechanics need them: lazily
        self.mot = (self.to_device(np.asarray(motility, dtype=np.float64), "mot")
                    if motility is not None else self._zeros((n, 3), torch.float64))
        self.jkr = (self.to_device(np.asarray(jkr, dtype=np.float64), "jkr")
                    if jkr is not None else self._zeros((n, 3), torch.float64))

    def adopt(self, n, loc=None, radii=None, radii_host=None, mot=None, jkr=None):
        """Replace individual device arrays (tensors that already live on the device) and keep the others: what
        the reference-facing methods use when only some of the host arrays changed since the last call
        (arrays.py).  Radius statistics are re-evaluated only when the radii change."""
        self.use_current_stream()
        if n != self.n or loc is not None or radii is not None:
            self._carry = None
        self.n = int(n)
        if loc is not None:
            self.loc = loc
        if radii is not None:
            self.radii = radii
            self._radii_host = radii_host
            self._max_radius = self._uniform_radius = None if n else 0.0
        if mot is not None:
            self.mot = mot
        if jkr is not None:
            self.jkr = jkr

    def set_device_state(self, loc, radii, mot=None, jkr=None, max_radius=None, uniform_radius=0.0):
        """Adopt tensors that already live on the device (no copies)."""
        self.use_current_stream()
        self._carry = None
        self.n = int(loc.shape[0])
        self.loc, self.radii = loc, radii
        self._radii_host = None
        self.mot = mot if mot is not None else self._zeros((self.n, 3), torch.float64)
        self.jkr = jkr if jkr is not None else self._zeros((self.n, 3), torch.float64)
        self._max_radius = float(max_radius) if max_radius is not None else float(radii.max().item())
        self._uniform_radius = float(uniform_radius)

    def _radius_stats(self):
        src = self._radii_host
        if src is not None and src.shape == (self.n,):
            # the host array the radii were uploaded from, within the same public call: evaluating it here keeps
            # the host from waiting for the upload before it can enqueue the mechanics (0.6 ms at 500 k cells)
            lo, hi = float(src.min()), float(src.max())
        else:
            lo, hi = (float(v) for v in torch.aminmax(self.radii[:self.n]))   # one small kernel + one read-back
        self._max_radius = hi
        self._uniform_radius = hi if lo == hi else 0.0

    @property
    def max_radius(self) -> float:
        if self._max_radius is None:
            self._radius_stats()
        return self._max_radius

    @max_radius.setter
    def max_radius(self, v):
        self._max_radius = v

    @property
    def uniform_radius(self) -> float:
        if self._uniform_radius is None:
            self._radius_stats()
        return self._uniform_radius

    @uniform_radius.setter
    def uniform_radius(self, v):
        self._uniform_radius = v

    def locations_host(self) -> np.ndarray:
        return self.host(self.loc)

    # ---- a1/a2: binning ------------------------------------------------------------------------------
    def set_bin_lattice(self, origin, extent):
        """Bin only the sub-box [origin, origin+extent) (a slab and its halo); None = the whole [0,size] box."""
        self.bin_origin = np.zeros(3) if origin is None else np.asarray(origin, dtype=np.float64).reshape(3)
        self.bin_extent = self.size.copy() if extent is None else np.asarray(extent, dtype=np.float64).reshape(3)
        self._origin_c = _lib.d3(self.bin_origin)
        self._extent_c = _lib.d3(self.bin_extent)

    def bin_dims(self, distance, flags=0):
        dims = (C.c_int32 * 3)()
        _lib.check(self.lib.abm_bin_dims_host(self._extent_c, float(distance), int(flags), dims),
                   "abm_bin_dims_host")
        return tuple(int(d) for d in dims)

    def _sort_buffers(self, distance, flags=0):
        n = self.n
        dims = self.bin_dims(distance, flags)
        nbins = dims[0] * dims[1] * dims[2]
        b = self._sort_bufs.get(nbins)
        if b is None or b["cap"] < n:
            cap = max(int(1.1 * n) + 1024, 1)
            b = dict(cap=cap, bin_start=self._empty(nbins + 1, torch.int32),
                     bin_start_b=self._empty(nbins + 1, torch.int32), sorted_id=self._empty(cap, torch.int32),
                     sorted_pos=self._empty((cap, 4), torch.float64), sorted_bin=self._empty(cap, torch.int32),
                     max_count=self._empty(1, torch.int32))
            if len(self._sort_bufs) > 4:
                self._sort_bufs.clear()
            self._sort_bufs[nbins] = b
        need = self.lib.abm_bin_sort_workspace_bytes(n, nbins)
        ws = self._sort_wss.get(nbins)
        if ws is None or ws.numel() < need:
            if len(self._sort_wss) > 4:
                self._sort_wss.clear()
            ws = self._sort_wss[nbins] = self._empty(int(1.1 * need) + 4096, torch.uint8)
        b["ws"] = ws
        return dims, nbins, b

    def sort(self, distance, flags=0, loc=None) -> SortState:
        """Counting sort of the current locations (or `loc`) into bins of edge `distance` (abm_bin_sort).
        flags: _lib.SORT_STABLE | _lib.SORT_REFERENCE for the reference-shaped assign_bins() output."""
        n = self.n
        dims, nbins, b = self._sort_buffers(distance, flags)
        if self._carry is not None and self._carry[0] == nbins:
            self._carry = None                               # this sort reuses the workspace the hand-over lives in
        _lib.check(self.lib.abm_bin_sort(_ptr(self.loc if loc is None else loc), _ptr(self.radii), n,
                                         self._extent_c, self._origin_c,
                                         float(distance), int(flags), _ptr(b["bin_start"]), _ptr(b["sorted_id"]),
                                         _ptr(b["sorted_pos"]), _ptr(b["sorted_bin"]), _ptr(b["max_count"]),
                                         _ptr(b["ws"]), b["ws"].numel(), _ptr(self.n_device), self.stream),
                   "abm_bin_sort")
        self.launches += (5 if flags & _lib.SORT_STABLE else 4) if n else 0
        return SortState(float(distance), dims, nbins, n, b["bin_start"], b["sorted_id"], b["sorted_pos"],
                         b["sorted_bin"], b["max_count"])

    def bins_dense(self, sort: SortState, max_agents: int):
        """Reference-shaped (bins, bins_help, bin_locations) of Simulation.assign_bins."""
        dims, n = sort.dims, sort.n
        bins = self._zeros(dims + (int(max_agents),), torch.int64)
        bins_help = self._zeros(dims, torch.int64)
        bin_locations = self._zeros((n, 3), torch.int64)
        _lib.check(self.lib.abm_bins_dense(_ptr(self.loc), n, self._size_c, sort.distance, _ptr(sort.bin_start),
                                           _ptr(sort.sorted_id), int(max_agents), _ptr(bins), _ptr(bins_help),
                                           _ptr(bin_locations), self.stream), "abm_bins_dense")
        self.launches += 1 if n else 0
        return bins, bins_help, bin_locations

    # ---- a3/a4: search -> row CSR --------------------------------------------------------------------
    def read_i32(self, t: torch.Tensor) -> int:
        out = C.c_int32(0)
        _lib.check(self.lib.abm_read_i32_sync(_ptr(t), C.byref(out), self.stream), "abm_read_i32_sync")
        return int(out.value)

    def search(self, sort: SortState, distance, old: Optional[GraphTensors] = None,
               capacity: Optional[int] = None) -> GraphTensors:
        """Fixed-radius search (abm_neighbor_search); grows the column buffer and retries on overflow,
        like the reference's max_neighbors loop (simulation.py:277-281)."""
        n = sort.n
        row_start = self._empty(max(n, 1), torch.int32)
        row_len = self._zeros(max(n, 1), torch.int32)
        cursor = self._empty(1, torch.int32)
        hint_key = (round(float(distance), 9), old is not None)
        cap = int(capacity if capacity is not None else self.cols_capacity_hint.get(hint_key, 8 * max(n, 1)))
        while True:
            cols = self._empty(max(cap, 1), torch.int32)
            self.enqueue_search(sort, distance, old, row_start, row_len, cols, cursor)
            used = self.read_i32(cursor)
            if used <= cap:
                break
            cap = int(used * 1.25) + 1024
        self.cols_capacity_hint[hint_key] = max(int(used * 1.25) + 1024, 1024)
        g = GraphTensors(n, row_start, 0, row_len, cols, True)
        g.entries = used
        return g

    def enqueue_search(self, sort: SortState, distance, old, row_start, row_len, cols, cursor, loc=None):
        """One abm_neighbor_search launch into caller-provided buffers; no host sync (the caller checks
        cursor <= cols.numel() later)."""
        if distance > sort.distance * (1 + 1e-12):
            raise _lib.AbmError("search radius exceeds the bin edge of the sort")
        n = sort.n
        _lib.check(self.lib.abm_neighbor_search(
            _ptr(sort.bin_start), _ptr(sort.sorted_id), _ptr(sort.sorted_pos), _ptr(sort.sorted_bin), n,
            _lib.i3(sort.dims), float(distance), _ptr(self.loc if loc is None else loc),
            _ptr(old.row_start) if old is not None else 0, old.row_stride if old is not None else 0,
            _ptr(old.row_len) if old is not None else 0, _ptr(old.cols) if old is not None else 0,
            _ptr(row_start), _ptr(row_len), _ptr(cols), cols.numel(), _ptr(cursor), _ptr(self.n_device), self.stream),
            "abm_neighbor_search")
        self.launches += 1 if n else 0
        return GraphTensors(n, row_start, 0, row_len, cols, True)

    def edge_list(self, g: GraphTensors) -> np.ndarray:
        """Canonical (i<j) int64 edge array of a graph (abm_graph_edge_list), downloaded."""
        n = g.n
        if n == 0:
            return np.zeros((0, 2), dtype=np.int64)
        if not g.sorted_rows:
            raise _lib.AbmError("edge_list needs ascending rows")
        need = self.lib.abm_graph_edge_list_workspace_bytes(n)
        ws = self._empty(int(need), torch.uint8)
        n_edges = self._empty(1, torch.int32)
        cap = int(g.cols.numel() // 2 + 1)
        edges = self._empty((cap, 2), torch.int64)
        _lib.check(self.lib.abm_graph_edge_list(_ptr(g.row_start), g.row_stride, _ptr(g.row_len), _ptr(g.cols), n,
                                                _ptr(ws), ws.numel(), _ptr(edges), cap, _ptr(n_edges), self.stream),
                   "abm_graph_edge_list")
        self.launches += 4
        e = self.read_i32(n_edges)
        if e > cap:
            raise _lib.AbmError("edge list capacity underestimated")
        return self.host(edges[:e])

    # ---- a6/a7/a9: unfused mechanics -------------------------------------------------------------------
    def jkr_forces(self, g: GraphTensors):
        """calculate_jkr on the device: accumulates into self.jkr, removes broken bonds from g in place."""
        _lib.check(self.lib.abm_jkr_forces(_ptr(g.row_start), g.row_stride, _ptr(g.row_len), _ptr(g.cols),
                                           _ptr(self.loc), _ptr(self.radii), _ptr(self.jkr), g.n, JKR_POISSON,
                                           JKR_YOUNGS, JKR_ADHESION, self.stream), "abm_jkr_forces")
        self.launches += 1 if g.n else 0

    def integrate(self, move_dt, zero_jkr=True, stokes=STOKES):
        self._carry = None
        _lib.check(self.lib.abm_apply_forces(_ptr(self.jkr), _ptr(self.mot), _ptr(self.loc), _ptr(self.radii),
                                             self.n, float(stokes), self._size_c, float(move_dt),
                                             _ptr(self.jkr) if zero_jkr else 0, self.stream), "abm_apply_forces")
        self.launches += 1 if self.n else 0

    # ---- a8: fused mechanics loop ------------------------------------------------------------------------
    def jkr_reach(self, distance) -> float:
        return float(self.lib.abm_jkr_reach(float(distance), float(self.max_radius), JKR_POISSON, JKR_YOUNGS,
                                            JKR_ADHESION))

    def _bond_buffers(self, n, stride):
        """Three rotating fixed-stride row buffers (+ the overflow flag).  Allocated with some headroom so a
        slab whose cell count drifts from step to step does not reallocate."""
        b = self._bond_bufs
        if b is None or b["cap"] < n or b["stride"] != stride:
            cap = max(int(1.1 * n) + 1024, 1)
            b = dict(cap=cap, stride=stride,
                     len_full=[self._zeros(cap, torch.int32) for _ in range(3)],
                     cols_full=[self._empty((cap, stride), torch.int32) for _ in range(3)],
                     max_len=self._zeros(2, torch.int32))      # [0] longest row, [1] largest substep move (float bits)
            self._bond_bufs = b
        if b.get("n") != n:
            b["n"] = n
            b["len"] = [t[:max(n, 1)] for t in b["len_full"]]
            b["cols"] = [t[:max(n, 1)] for t in b["cols_full"]]
        return b

    def enqueue_mechanics(self, old: Optional[GraphTensors], distance, dts, stride, stokes=STOKES, pair=None,
                          zero_flag=True, carry=False):
        """Enqueue len(dts) fused substeps (sort + abm_jkr_substep each).  No host sync.  Returns the
        fixed-stride graph holding the final bonds; overflow is reported through bufs['max_len'].
        carry: the caller runs such loops back to back on device-resident cells (pipeline.HotPathStep): the last
        substep then hands its cells to the next call's first sort, which skips its histogram pass.  The hand-over
        is dropped whenever the cells or the sort workspace are touched in between (upload, set_device_state,
        integrate, a stand-alone sort on the same lattice, a loop without `carry`)."""
        n = self.n
        b = self._bond_buffers(n, stride)
        if zero_flag:
            b["max_len"].zero_()
        if len(dts) == 0:
            return old
        reach = self.jkr_reach(distance)
        # three row buffers: the two that do not hold `old` ping-pong, so `old` survives a retry
        held = [i for i in range(3) if old is not None and old.cols.data_ptr() == b["cols"][i].data_ptr()]
        if pair is None:
            pair = [i for i in range(3) if i not in held][:2]
        dims, nbins, sb = self._sort_buffers(reach)
        dts_c = (C.c_double * len(dts))(*[float(d) for d in dts])
        tables = [sb["bin_start"], sb["bin_start_b"]]
        flags, first = 0, 0
        # everything the hand-over depends on: lattice, cells, radii (they travel with the positions), the buffers
        key = (nbins, n, self.loc.data_ptr(), self.radii.data_ptr(), reach, sb["ws"].data_ptr(),
               sb["bin_start"].data_ptr(), sb["bin_start_b"].data_ptr(), self.mechanics_dtype, self.plane_kernel,
               float(self.uniform_radius))     # (the last two decide the layout of the handed-over records)
        if self.chained_launch:
            flags |= _lib.LOOP_CHAINED
        if not self.plane_kernel:
            flags |= _lib.LOOP_NO_PLANE
        if self.flat_z_vouched and float(self._size_c[2]) == 0.0:
            flags |= _lib.LOOP_FLAT_Z
        if self.mechanics_dtype == "f32":
            flags |= _lib.LOOP_FP32
        elif self.mechanics_dtype != "f64":
            raise _lib.AbmError(f"mechanics_dtype must be 'f64' or 'f32', not {self.mechanics_dtype!r}")
        if carry and n:
            flags |= _lib.LOOP_CARRY_OUT
            if self._carry is not None and self._carry[:-1] == key:
                flags |= _lib.LOOP_CARRY_IN
                first = self._carry[-1]                      # the table the previous call left the histogram in
        events = self.substep_events
        timed = ()
        call = self.lib.abm_mechanics_loop
        if events is not None:
            if len(events) != 2 * len(dts):
                raise _lib.AbmError(f"substep_events: {len(events)} events for {len(dts)} substeps")
            timed = ((C.c_void_p * len(events))(*[int(e.cuda_event) for e in events]),)
            call = self.lib.abm_mechanics_loop_timed
        _lib.check(call(
            _ptr(self.radii), _ptr(self.loc), _ptr(self.mot), n, self._extent_c, self._origin_c, self._size_c,
            float(distance), reach, len(dts), dts_c,
            _ptr(tables[first]), _ptr(tables[1 - first]), _ptr(sb["sorted_id"]), _ptr(sb["sorted_pos"]),
            _ptr(sb["sorted_bin"]), _ptr(sb["max_count"]), _ptr(sb["ws"]), sb["ws"].numel(),
            _ptr(old.row_start) if old is not None else 0, old.row_stride if old is not None else 0,
            _ptr(old.row_len) if old is not None else 0, _ptr(old.cols) if old is not None else 0,
            _ptr(b["len"][pair[0]]), _ptr(b["cols"][pair[0]]), _ptr(b["len"][pair[1]]), _ptr(b["cols"][pair[1]]),
            stride, _ptr(b["max_len"]), _ptr(self.global_id), float(self.uniform_radius), JKR_POISSON, JKR_YOUNGS,
            JKR_ADHESION, float(stokes), flags, _ptr(self.n_device), _ptr(self.move_window), *timed, self.stream),
            "abm_mechanics_loop")
        # the last substep k = K-1 sorted with table (first + K-1) % 2 and left the next histogram in the other one
        self._carry = key + ((first + len(dts)) % 2,) if flags & _lib.LOOP_CARRY_OUT else None
        # first substep: histogram, scan x2, scatter, substep (scan, rescatter, substep when the cells were handed
        # over by the previous call); every later one: scan, rescatter, substep
        self.launches += (3 * len(dts) + (0 if flags & _lib.LOOP_CARRY_IN else 2)) if n else 0
        last = pair[(len(dts) - 1) & 1]
        return GraphTensors(n, None, stride, b["len"][last], b["cols"][last], True)

    def max_substep_move(self) -> float:
        """Host synchronisation: the largest distance (um) a cell moved in ONE substep since the flag was last zeroed."""
        if self._bond_bufs is None:
            return 0.0
        bits = int(self._bond_bufs["max_len"][1].item())
        return float(np.sqrt(np.array([bits], dtype=np.int32).view(np.float32)[0]))

    def mechanics(self, old: Optional[GraphTensors], distance, dts, stride=16, stokes=STOKES, meanwhile=None,
                  fetch_locations=False):
        """The whole substep loop of CellMethods.apply_forces on the device, with the reference's
        grow-and-retry on row overflow.  Returns (bond graph, stride used).  meanwhile: host work to do once while
        the device runs the loop (called after the launches are enqueued, before the first wait).
        fetch_locations: also return the moved positions as a host array; their copy is enqueued right behind the
        loop, so it runs while the host is still busy with `meanwhile`."""
        if self.n == 0 or len(dts) == 0:
            if meanwhile is not None:
                meanwhile()
            return (old, stride, self.locations_host()) if fetch_locations else (old, stride)
        start = self.loc.clone()
        while True:
            g = self.enqueue_mechanics(old, distance, dts, stride, stokes)
            moved = self.host(self.loc, sync=False) if fetch_locations else None
            if meanwhile is not None:
                meanwhile()
                meanwhile = None
            worst = self.read_i32(self._bond_bufs["max_len"])      # synchronises the stream: `moved` is complete
            if worst <= stride:
                return (g, stride, moved) if fetch_locations else (g, stride)
            stride = 2 * worst          # reference simulation.py:281 doubles the needed capacity
            self.loc.copy_(start)

    # ---- a10-a13: morphogen field ---------------------------------------------------------------------------
    def diffuse(self, field: torch.Tensor, steps, diffuse_dt, last_dt, diffuse_const, spat_res2, max_concentration,
                degradation, use_tma=None, scratch=None) -> torch.Tensor:
        """update_diffusion on a (gx,gy) float64 device plane; returns the plane holding the result
        (either `field` or the scratch plane -- no copy back)."""
        gx, gy = int(field.shape[0]), int(field.shape[1])
        if use_tma is None:
            use_tma = (gy % 2 == 0)
        if scratch is None:
            scratch = torch.empty_like(field)
        where = C.c_int(0)
        _lib.check(self.lib.abm_diffusion(_ptr(field), _ptr(scratch), gx, gy, int(steps), float(diffuse_dt),
                                          float(last_dt), float(diffuse_const), float(spat_res2),
                                          float(max_concentration), float(degradation), 1 if use_tma else 0,
                                          self.diffusion_fuse, C.byref(where), self.stream), "abm_diffusion")
        self.launches += int(self.lib.abm_diffusion_launches(int(steps), self.diffusion_fuse))
        return scratch if where.value else field

    def diffuse3d(self, field: torch.Tensor, steps, diffuse_dt, last_dt, diffuse_const, spat_res2, max_concentration,
                  degradation, use_tma=None, scratch=None) -> torch.Tensor:
        """7-point diffusion of a (gx,gy,gz) float64 device field (3-D extension, abm_diffusion3d); returns the
        tensor holding the result (either `field` or the scratch field -- no copy back)."""
        gx, gy, gz = (int(v) for v in field.shape)
        if use_tma is None:
            use_tma = (gz % 2 == 0)
        if scratch is None:
            scratch = torch.empty_like(field)
        where = C.c_int(0)
        _lib.check(self.lib.abm_diffusion3d(_ptr(field), _ptr(scratch), gx, gy, gz, int(steps), float(diffuse_dt),
                                            float(last_dt), float(diffuse_const), float(spat_res2),
                                            float(max_concentration), float(degradation), 1 if use_tma else 0,
                                            C.byref(where), self.stream), "abm_diffusion3d")
        self.launches += int(steps)
        return scratch if where.value else field

    def sample(self, field: torch.Tensor, spat_res, rows=None, grid_rows=None, loc=None, n=None,
               out=None, use_count=True) -> torch.Tensor:
        """get_concentration for every cell.  `field` is the whole (gx,gy[,gz]) grid, or -- with rows=(lo,hi) and
        grid_rows=gx -- the slab of global rows [lo,hi) of it."""
        gs = list(tuple(field.shape) + (1,) * (3 - field.dim()))
        lo, hi = (0, gs[0]) if rows is None else rows
        if grid_rows is not None:
            gs[0] = int(grid_rows)
        n = self.n if n is None else int(n)
        if out is None:                  # (a caller that replays captured graphs passes its own persistent buffer)
            out = self._empty(max(n, 1), torch.float64)
        _lib.check(self.lib.abm_sample_concentrations(_ptr(self.loc if loc is None else loc), n, _ptr(field),
                                                      _lib.l3(gs),
                                                      float(spat_res), int(lo), int(hi), _ptr(out),
                                                      _ptr(self.n_device) if use_count else 0, self.stream),
                   "abm_sample_concentrations")
        self.launches += 1 if n else 0
        return out[:n]

    def deposit(self, field: torch.Tensor, amounts, spat_res, rows=None, grid_rows=None, loc=None, n=None,
                use_count=True, cube=False):
        """adjust_morphogens for every cell; amounts: a device tensor (one per cell) or a python float (all cells).
        cube: the 3-D extension -- the eight corners of the cell's grid cube instead of the reference's four corner
        points on the z = 0 plane."""
        gs = list(tuple(field.shape) + (1,) * (3 - field.dim()))
        lo, hi = (0, gs[0]) if rows is None else rows
        if grid_rows is not None:
            gs[0] = int(grid_rows)
        n = self.n if n is None else int(n)
        scalar = not isinstance(amounts, torch.Tensor)
        fn = self.lib.abm_deposit_morphogens3d if cube else self.lib.abm_deposit_morphogens
        _lib.check(fn(_ptr(self.loc if loc is None else loc),
                      0 if scalar else _ptr(amounts), float(amounts) if scalar else 0.0, n, _ptr(field), _lib.l3(gs),
                      float(spat_res), int(lo), int(hi), _ptr(self.n_device) if use_count else 0, self.stream),
                   "abm_deposit_morphogens")
        self.launches += 1 if n else 0
