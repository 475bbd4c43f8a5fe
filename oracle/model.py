"""Host rule model used by the 100-step state-identity tests (TEST INFRASTRUCTURE ONLY).

The hot path only makes sense inside the reference's step loop, whose other half -- division, death, the
FGF/ERK Boolean network, differentiation and motility -- is stochastic per-cell Python that stays on the host
(out of scope for kernels).  The reference cannot travel to the GPU box, so the rules are
restated here, consuming Python's `random` / `numpy.random` streams in exactly the reference's order, and are
pinned by tests/test_model_states.py: with the CPU oracle as backend they must reproduce, step for step, the
discrete states recorded from the reference's own CellSimulation.steps() (oracle/make_golden.py).  The same
class then drives the B200 backend in the GPU test.

Reference being followed: cell_simulation.py:14-57 (constants), :77-126 (step order), :128-157 (initial arrays);
cell_methods.py:11-52 (death), :54-117 (division), :176-228 (pathway), :230-244 (differentiation),
:246-340 (motility); simulation.py:473-487 (random_vector); backend.py:186-196 (normal_vector).
"""
from __future__ import annotations

import math
import random as r

import numpy as np

TWO_PI = 2 * math.pi


class ColonyRules:
    """Mixin: needs the backend to provide get_neighbors(), apply_forces(), agent_graph() and the graphs'
    neighbors()/num_neighbors()/add_vertex()/delete_vertices()."""

    # ---- constants (cell_simulation.py:33-57, templates/experimental.yaml as used for the golden runs) ----------
    def init_constants(self, size, dox_step=5, lonely_thresh=2, guye_move=True, num_gata6=0):
        self.size = np.array(size, dtype=float)
        self.number_agents, self.current_step = 0, 0
        self.agent_array_names, self.graph_names, self.method_times = [], [], {}
        self.step_dt, self.move_dt = 1800, 180
        self.field = 2
        self.pluri_div_thresh, self.diff_div_thresh, self.pluri_to_diff = 36, 72, 36
        self.death_thresh, self.fds_thresh = 144, 1
        self.max_radius = 5
        self.min_radius = self.max_radius / 2 ** 0.5
        self.dox_step, self.lonely_thresh, self.guye_move, self.num_gata6 = dox_step, lonely_thresh, guye_move, num_gata6

    def _new_array(self, name, values):
        setattr(self, name, values)
        self.agent_array_names.append(name)

    def _drawn(self, n, hi):
        """n successive r.randrange(0, hi) draws (what agent_array(func=...) does, simulation.py:176-179)."""
        return np.array([r.randrange(0, hi) for _ in range(n)], dtype=int)

    def seed_colony(self, num_to_start):
        """cell_simulation.py:128-157, same RNG consumption order."""
        n = num_to_start + self.num_gata6
        self.number_agents = n
        self._new_array("locations", np.random.rand(n, 3) * self.size)
        self._new_array("radii", np.full(n, float(self.max_radius)))
        for name in ("FGF4", "FGFR", "ERK"):
            self._new_array(name, self._drawn(n, self.field))
        self._new_array("GATA6", np.zeros(n, dtype=int))
        self._new_array("NANOG", self._drawn(n, self.field))
        self._new_array("states", np.zeros(n, dtype=int))
        self._new_array("death_counters", self._drawn(n, self.death_thresh))
        self._new_array("diff_counters", self._drawn(n, self.pluri_to_diff))
        self._new_array("div_counters", self._drawn(n, self.pluri_div_thresh))
        self._new_array("fds_counters", self._drawn(n, self.fds_thresh))
        self._new_array("motility_forces", np.zeros((n, 3)))
        self._new_array("jkr_forces", np.zeros((n, 3)))
        for i in range(num_to_start, n):                      # the "GATA6_high" agent type
            self.GATA6[i] = r.randrange(1, self.field)
        self.NANOG[num_to_start:n] = 0
        self.agent_graph("neighbor_graph")
        self.agent_graph("jkr_graph")

    # ---- helpers -----------------------------------------------------------------------------------------------
    def random_vector(self):
        theta = r.random() * TWO_PI
        if self.size[2] == 0:
            return np.array([math.cos(theta), math.sin(theta), 0])
        phi = r.random() * TWO_PI
        rad = math.cos(phi)
        return np.array([rad * math.cos(theta), rad * math.sin(theta), math.sin(phi)])

    @staticmethod
    def unit(v):
        mag = np.linalg.norm(v)
        return np.zeros(3) if mag == 0 else v / mag

    # ---- rules ---------------------------------------------------------------------------------------------------
    def cell_division(self):
        n = self.number_agents
        dividing = []
        for i in range(n):
            self.div_counters[i] += r.randint(0, 1)
            if self.states[i] == 0:
                ready = self.div_counters[i] >= self.pluri_div_thresh
            else:     # contact inhibition for differentiated cells
                ready = self.div_counters[i] >= self.diff_div_thresh and self.neighbor_graph.num_neighbors(i) < 6
            if ready:
                dividing.append(i)
        mothers = np.array(dividing, dtype=int)
        for name in self.agent_array_names:
            arr = getattr(self, name)
            setattr(self, name, np.concatenate((arr, arr[mothers]), axis=0))
        for k, mother in enumerate(mothers):
            daughter = n + k
            shift = self.random_vector() * (self.max_radius - self.min_radius)
            self.locations[mother] += shift
            self.locations[daughter] -= shift
            self.div_counters[mother] = self.div_counters[daughter] = 0
            for g in self.graph_names:
                getattr(self, g).add_vertex()
        self.number_agents = n + len(mothers)

    def cell_death(self):
        doomed = []
        for i in range(self.number_agents):
            if self.states[i] == 0:
                if self.neighbor_graph.num_neighbors(i) < self.lonely_thresh:
                    self.death_counters[i] += 1
                else:
                    self.death_counters[i] = 0
                if self.death_counters[i] >= self.death_thresh:
                    doomed.append(i)
        doomed = np.array(doomed, dtype=int)
        for name in self.agent_array_names:
            setattr(self, name, np.delete(getattr(self, name), doomed, axis=0))
        for g in self.graph_names:
            getattr(self, g).delete_vertices(doomed)
        self.number_agents -= len(doomed)

    def cell_pathway(self):
        if self.current_step < self.dox_step:
            return
        top = self.field - 1
        for i in range(self.number_agents):
            group = self.neighbor_graph.neighbors(i)
            group.append(i)
            share = len(group)
            sensed = 0
            for j in group:                                    # sequential: earlier cells were already updated
                sensed += (1 + r.gauss(0, 1)) * (self.FGF4[j] / share)
            x1 = min(max(int(sensed), 0), top)
            if self.fds_counters[i] % self.fds_thresh == 0:
                x2, x3, x4, x5 = self.FGFR[i], self.ERK[i], self.GATA6[i], self.NANOG[i]
                if self.field == 2:
                    self.FGF4[i] = x5
                    self.FGFR[i] = (1 + x5 + x4 * x5) % 2
                    self.ERK[i] = (x1 * x2) % 2
                    self.GATA6[i] = (x3 + x4 + x3 * x4 + x3 * x5 + x4 * x5 + x3 * x4 * x5) % 2
                    self.NANOG[i] = (x5 + x3 * x5 + x4 * x5 + x3 * x4 * x5) % 2
                else:
                    self.FGF4[i] = x5
                    self.FGFR[i] = (x1 * x4 * ((2 * x1 + 1) * (2 * x4 + 1) + x1 * x4)) % 3
                    self.ERK[i] = x2 % 3
                    self.GATA6[i] = ((x4 ** 2) * (x5 + 1) + (x5 ** 2) * (x4 + 1) + 2 * x5 + 1) % 3
                    self.NANOG[i] = (x5 ** 2 + x5 * (x5 + 1) * (x3 * (2 * x4 ** 2 + 2 * x3 + 1) +
                                     x4 * (2 * x3 ** 2 + 2 * x4 + 1)) + (2 * x3 ** 2 + 1) * (2 * x4 ** 2 + 1)) % 3
            self.fds_counters[i] += 1

    def cell_differentiate(self):
        for i in range(self.number_agents):
            if self.GATA6[i] > self.NANOG[i] and self.states[i] == 0:
                self.diff_counters[i] += r.randint(0, 1)
                if self.diff_counters[i] >= self.pluri_to_diff:
                    self.states[i] = 1
                    self.NANOG[i] = 0

    def _pull(self, i, nbrs, toward):
        """Motility of cell i relative to its NANOG-high neighbours: +0.8 toward / -0.8 away, 0.2 random."""
        high = [j for j in nbrs if self.NANOG[j] > self.GATA6[j]]
        if not high:
            return self.random_vector()
        total = np.array([0.0, 0.0, 0.0])
        for j in high:
            total += self.locations[j] - self.locations[i]
        heading = self.unit(total)
        return heading * (0.8 if toward else -0.8) + self.random_vector() * 0.2

    def cell_motility(self):
        force = 0.000000002
        for i in range(self.number_agents):
            nbrs = self.neighbor_graph.neighbors(i)
            if len(nbrs) >= 6:
                continue
            if self.states[i] == 1:
                move = self._pull(i, nbrs, toward=False)
            elif self.GATA6[i] > self.NANOG[i]:
                # Guye mode tests the cell's own (pluripotent) state instead of the neighbour's, so it never
                # finds a target and moves randomly (reference cell_methods.py:287; kept for state parity)
                move = self.random_vector() if self.guye_move else self._pull(i, nbrs, toward=False)
            elif self.GATA6[i] < self.NANOG[i]:
                move = self._pull(i, nbrs, toward=True)
            else:
                move = self.random_vector()
            self.motility_forces[i] += move * force

    # ---- step loop (cell_simulation.py:77-126 without the outputs) ---------------------------------------------------
    def run_one(self, step):
        self.current_step = step
        self.get_neighbors("neighbor_graph", 15)
        self.cell_division()
        self.cell_death()
        self.cell_pathway()
        self.cell_differentiate()
        self.cell_motility()
        self.apply_forces()

    def run(self, end_step, on_step=None):
        for step in range(1, end_step + 1):
            self.run_one(step)
            if on_step is not None:
                on_step(self)
