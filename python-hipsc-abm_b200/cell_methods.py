"""Host-side mirror of the hot-path methods of the reference `CellMethods`
(reference cell_methods.py:342-521), backed by libabm_b200.so.

Method names, arguments and side effects follow the reference; `jkr_neighbor_forces` and
`update_concentrations` are the aliases / batched form named in BASELINE.json's north star.
"""
from __future__ import annotations

import numpy as np

from .engine import STOKES
from .simulation import record_time


class B200CellMethods:
    # ---- a6/a7 ---------------------------------------------------------------------------------------------
    def calculate_jkr(self):
        """reference cell_methods.py:342-384: JKR force of every bond of jkr_graph into self.jkr_forces;
        bonds with d <= -0.360562 are deleted from the graph."""
        graph = self.jkr_graph
        if graph._dev is None and graph.ecount() == 0:       # (device copy first: ecount() would download it)
            return
        if graph.n != self.number_agents:
            raise ValueError(f"jkr_graph has {graph.n} vertices for {self.number_agents} agents")
        eng = self.b200_upload(with_forces=True)
        g = graph.device_tensors(eng)
        eng.jkr_forces(g)
        graph.set_device(g, eng)
        self.jkr_forces = eng.host(eng.jkr)

    jkr_neighbor_forces = calculate_jkr

    # ---- a8/a9 ---------------------------------------------------------------------------------------------
    @record_time
    def apply_forces(self):
        """reference cell_methods.py:386-439: int(step_dt // move_dt) + 1 substeps of
        (contact search at 2*max_radius merged into jkr_graph, JKR forces, Stokes integration).

        When the class still uses this package's get_neighbors / calculate_jkr the whole loop runs as fused
        device substeps (one sort + one abm_jkr_substep launch each, nothing copied in between); a subclass
        that overrides either method gets the reference's literal stage-by-stage loop instead."""
        steps, last_dt = divmod(self.step_dt, self.move_dt)
        total_steps = int(steps) + 1
        dts = [self.move_dt] * (total_steps - 1) + [last_dt]
        distance = 2 * self.max_radius
        from .simulation import B200Simulation
        fused = (type(self).calculate_jkr is B200CellMethods.calculate_jkr and
                 getattr(type(self).get_neighbors, "__wrapped__", None) is
                 getattr(B200Simulation.get_neighbors, "__wrapped__", None))
        if fused:
            import torch
            eng = self.b200_upload(with_forces="motility")
            uploaded = torch.cuda.Event()
            uploaded.record(torch.cuda.current_stream(eng.device))
            graph = self.jkr_graph
            if graph.n != self.number_agents:
                # a graph left un-resized after division / death would be read out of bounds on the device
                raise ValueError(f"jkr_graph has {graph.n} vertices for {self.number_agents} agents")
            old = graph.device_tensors(eng)

            def reset_forces():
                # the reference leaves both force arrays zeroed (cell_methods.py:436,439); do it while the device
                # runs the loop -- jkr_forces is not read by the fused loop, motility_forces once its upload is done
                self.jkr_forces[:, :] = 0
                uploaded.synchronize()
                self.motility_forces[:, :] = 0
            bonds, graph.row_stride_hint, moved = eng.mechanics(old, distance, dts, graph.row_stride_hint, STOKES,
                                                                meanwhile=reset_forces, fetch_locations=True)
            graph.set_device(bonds, eng)
            self.locations = moved
            return
        else:
            for step in range(total_steps):
                self.get_neighbors("jkr_graph", distance, clear=False)
                self.calculate_jkr()
                eng = self.b200_upload(with_forces=True)
                eng.integrate(dts[step], zero_jkr=False, stokes=STOKES)
                self.locations = eng.locations_host()
                self.jkr_forces[:, :] = 0
        self.motility_forces[:, :] = 0

    # ---- a10/a11 ---------------------------------------------------------------------------------------------
    @record_time
    def update_diffusion(self, gradient_name):
        """reference cell_methods.py:441-468 for the (gx,gy,1) float64 field `gradient_name`."""
        import torch
        steps, last_dt = divmod(self.step_dt, self.diffuse_dt)
        steps = int(steps) + 1
        eng = self.b200_engine()
        field = self.__dict__[gradient_name]
        plane = eng.to_device(np.ascontiguousarray(field[:, :, 0], dtype=np.float64), "field")
        out = eng.diffuse(plane, steps, self.diffuse_dt, last_dt, self.diffuse_const, self.spat_res2,
                          self.max_concentration, self.degradation)
        eng.host_into(out, field[:, :, 0])

    # ---- a12/a13 ---------------------------------------------------------------------------------------------
    def get_concentration(self, gradient_name, index):
        """reference cell_methods.py:470-483 (one cell; use update_concentrations for all of them)."""
        return float(self.update_concentrations(gradient_name, None, cells=slice(index, index + 1))[0])

    def adjust_morphogens(self, gradient_name, index, amount):
        """reference cell_methods.py:485-521 (one cell)."""
        self.update_concentrations(gradient_name, np.array([amount], dtype=np.float64),
                                   cells=slice(index, index + 1), sample=False)

    def update_concentrations(self, gradient_name, amounts=None, cells=slice(None), sample=True):
        """Batched get_concentration + adjust_morphogens over `cells` (default: all): every cell first reads
        its nearest grid point, then -- if `amounts` is given -- deposits amounts[i] on its in-range corner
        points.  Returns the sampled concentrations (or None if sample=False)."""
        eng = self.b200_engine()
        field = self.__dict__[gradient_name]
        locations = np.ascontiguousarray(self.locations[cells], dtype=np.float64)
        saved = (eng.n, eng.loc, eng.radii, eng.mot, eng.jkr, eng._max_radius)
        eng.n = len(locations)
        eng.loc = eng.to_device(locations, "loc")
        try:
            f = eng.to_device(np.ascontiguousarray(field, dtype=np.float64), "field")
            sampled = eng.host(eng.sample(f, self.spat_res)) if sample else None
            if amounts is not None:
                if np.ndim(amounts) == 0:                     # one amount for every cell: no array travels
                    eng.deposit(f, float(amounts), self.spat_res)
                else:
                    a = np.ascontiguousarray(np.broadcast_to(np.asarray(amounts, dtype=np.float64), (eng.n,)))
                    eng.deposit(f, eng.to_device(a, "amounts"), self.spat_res)
                eng.host_into(f.view(field.shape), field)
        finally:
            eng.n, eng.loc, eng.radii, eng.mot, eng.jkr, eng._max_radius = saved
        return sampled
