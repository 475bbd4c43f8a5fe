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
        self.b200_adopt_host("jkr_forces", eng.host(eng.jkr), eng.jkr)

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
            # Forces already sitting in jkr_forces at entry: the reference's first calculate_jkr accumulates on top
            # of them and the first integration consumes them (cell_backend.py:108-109, cell_methods.py:435-436).
            # The array is known to be zero if nobody wrote to it since this method zeroed it last time.
            pre = self._b200_pending_jkr()
            if pre is not None:
                first_mot = eng.mot + eng.to_device(pre, "jkr_forces")
                keep = eng.mot
                eng.mot = first_mot
                try:
                    old, graph.row_stride_hint = eng.mechanics(old, distance, dts[:1], graph.row_stride_hint, STOKES)
                finally:
                    eng.mot = keep
                dts = dts[1:]

            def reset_forces():
                # the reference leaves both force arrays zeroed (cell_methods.py:436,439); do it while the device
                # runs the loop -- jkr_forces is not read by the fused loop, motility_forces once its upload is done
                self._b200_zero("jkr_forces")
                uploaded.synchronize()
                self._b200_zero("motility_forces")
            bonds, graph.row_stride_hint, moved = eng.mechanics(old, distance, dts, graph.row_stride_hint, STOKES,
                                                                meanwhile=reset_forces, fetch_locations=True)
            # only now (a row overflow makes eng.mechanics run the loop again, and the loop reads the device forces):
            # the device copy follows the host array to zero, so an untouched array is not uploaded next time
            self._b200_zero_device("motility_forces")
            graph.set_device(bonds, eng)
            self.b200_adopt_host("locations", moved, eng.loc)
            return
        else:
            for step in range(total_steps):
                self.get_neighbors("jkr_graph", distance, clear=False)
                self.calculate_jkr()
                eng = self.b200_upload(with_forces=True)
                eng.integrate(dts[step], zero_jkr=False, stokes=STOKES)
                self.b200_adopt_host("locations", eng.locations_host(), eng.loc)
                self.jkr_forces[:, :] = 0
        self.motility_forces[:, :] = 0

    def _b200_zero(self, name):
        """array[:, :] = 0 on the host, remembered: an array nobody has written to since is known to be zero."""
        a = self.__dict__[name]
        zeros = self.__dict__.setdefault("_b200_zeroed", {})
        known = zeros.get(name)
        from .arrays import TrackedArray, track
        if known is not None and isinstance(a, TrackedArray) and a._b200 is known[0] and a._b200.gen == known[1]:
            return                                          # still the zeros written last time
        t = track(a)
        if t is not a:
            self.__dict__[name] = t
        t.view(np.ndarray)[...] = 0
        t._touch()
        zeros[name] = (t._b200, t._b200.gen)

    def _b200_zero_device(self, name):
        """The host array `name` is known to be all zeros: make its device copy (if any) the same and current."""
        a = self.__dict__[name]
        known = self.__dict__.get("_b200_zeroed", {}).get(name)
        entry = self.__dict__.get("_b200_cache", {}).get(name)
        from .arrays import TrackedArray
        if (known is not None and entry is not None and isinstance(a, TrackedArray) and a._b200 is known[0] and
                a._b200.gen == known[1] and entry.state is a._b200 and entry.shape == a.shape):
            entry.tensor.zero_()
            entry.refresh()

    def _b200_pending_jkr(self):
        """jkr_forces if it may hold non-zero values at entry of apply_forces, else None."""
        a = self.__dict__.get("jkr_forces")
        if a is None:
            return None
        known = self.__dict__.get("_b200_zeroed", {}).get("jkr_forces")
        from .arrays import TrackedArray
        if known is not None and isinstance(a, TrackedArray) and a._b200 is known[0] and a._b200.gen == known[1]:
            return None
        plain = np.asarray(a).view(np.ndarray)
        return np.ascontiguousarray(plain, dtype=np.float64) if plain.any() else None

    # ---- a10/a11 ---------------------------------------------------------------------------------------------
    def _b200_field(self, gradient_name):
        """(host field, device tensor of the same shape).  Uploaded only if the host array changed since the device
        copy was made or downloaded (arrays.py); the three morphogen calls of a step then share one device copy."""
        f, _ = self.b200_device_array(gradient_name)
        return self.__dict__[gradient_name], f

    def _b200_field_to_host(self, gradient_name, device_field):
        """The device field is the newer one: bring the host array up to date (in place, it keeps its identity)."""
        eng = self.b200_engine()
        field = self.__dict__[gradient_name]
        plain = np.asarray(field).view(np.ndarray)
        eng.host_into(device_field.view(plain.shape), plain)
        entry = self.__dict__.get("_b200_cache", {}).get(gradient_name)
        if entry is not None:
            entry.tensor = device_field
            entry.refresh()

    @record_time
    def update_diffusion(self, gradient_name):
        """reference cell_methods.py:441-468 for the (gx,gy,1) float64 field `gradient_name`."""
        steps, last_dt = divmod(self.step_dt, self.diffuse_dt)
        steps = int(steps) + 1
        eng = self.b200_engine()
        field, f = self._b200_field(gradient_name)
        out = self._b200_diffuse_plane(eng, f, steps, last_dt)
        self._b200_field_to_host(gradient_name, out)

    # The reference's update_diffusion only ever touches the z = 0 plane ("currently 2D", cell_backend.py:184), also
    # of a (gx,gy,gz) field.  b200_diffusion_3d = True selects the 3-D extension for fields with gz > 1: 7-point
    # stencil over the whole field (abm_diffusion3d) and deposits over the eight corners of the cell's grid cube.
    b200_diffusion_3d = False

    def _b200_volume(self, f):
        return bool(self.b200_diffusion_3d) and f.dim() == 3 and f.shape[2] > 1

    def _b200_diffuse_plane(self, eng, f, steps, last_dt):
        """`steps` diffusion substeps on the device field f (gx,gy,gz); returns the device field holding the result
        (f itself, or its ping-pong partner)."""
        args = (steps, self.diffuse_dt, last_dt, self.diffuse_const, self.spat_res2, self.max_concentration,
                self.degradation)
        if self._b200_volume(f):
            return eng.diffuse3d(f, *args)
        if f.dim() == 3 and f.shape[2] != 1:                      # reference semantics: the z = 0 plane of a volume
            plane = f[:, :, 0].contiguous()
            f[:, :, 0] = eng.diffuse(plane, *args)
            return f
        plane = f.view(f.shape[0], f.shape[1])
        return eng.diffuse(plane, *args).view(f.shape)

    # ---- a12/a13 ---------------------------------------------------------------------------------------------
    def get_concentration(self, gradient_name, index):
        """reference cell_methods.py:470-483 (one cell; use update_concentrations for all of them)."""
        return float(self.update_concentrations(gradient_name, None, cells=slice(index, index + 1))[0])

    def adjust_morphogens(self, gradient_name, index, amount):
        """reference cell_methods.py:485-521 (one cell)."""
        self.update_concentrations(gradient_name, np.array([amount], dtype=np.float64),
                                   cells=slice(index, index + 1), sample=False)

    def update_concentrations(self, gradient_name, amounts=None, cells=slice(None), sample=True, diffuse=False):
        """Batched get_concentration + adjust_morphogens over `cells` (default: all).  Order of the stages:

          diffuse=False (default)  every cell first reads its nearest grid point (if sample), then -- if `amounts` is
                                   given -- deposits amounts[i] on its in-range corner points;
          diffuse=True             deposit -> update_diffusion -> sample as ONE pass over a device-resident field
                                   (the north star's "diffusion fused with the per-cell gradient sampling"): what three
                                   separate calls update_concentrations(amounts, sample=False); update_diffusion();
                                   update_concentrations(None) compute, with one download of the field instead of two
                                   and none of the intermediate uploads.

        Returns the sampled concentrations (or None if sample=False)."""
        eng = self.b200_engine()
        whole = isinstance(cells, slice) and cells == slice(None)
        n = int(self.number_agents)
        if whole:
            loc, _ = self.b200_device_array("locations", (n, 3))
            count = n
        else:
            part = np.ascontiguousarray(np.asarray(self.locations).view(np.ndarray)[cells], dtype=np.float64)
            loc, count = eng.to_device(part, "loc_part"), len(part)
        field, f = self._b200_field(gradient_name)
        sampled = None
        changed = False
        if sample and not diffuse:
            sampled = eng.host(eng.sample(f, self.spat_res, loc=loc, n=count))
        if amounts is not None:
            cube = self._b200_volume(f)
            if np.ndim(amounts) == 0:                     # one amount for every cell: no array travels
                eng.deposit(f, float(amounts), self.spat_res, loc=loc, n=count, cube=cube)
            else:
                a = np.ascontiguousarray(np.broadcast_to(np.asarray(amounts, dtype=np.float64), (count,)))
                eng.deposit(f, eng.to_device(a, "amounts"), self.spat_res, loc=loc, n=count, cube=cube)
            changed = True
        if diffuse:
            steps, last_dt = divmod(self.step_dt, self.diffuse_dt)
            f = self._b200_diffuse_plane(eng, f, int(steps) + 1, last_dt)
            changed = True
            if sample:
                sampled = eng.host(eng.sample(f, self.spat_res, loc=loc, n=count), sync=False)
        if changed:
            self._b200_field_to_host(gradient_name, f)     # synchronises the stream: `sampled` is complete too
        return sampled
