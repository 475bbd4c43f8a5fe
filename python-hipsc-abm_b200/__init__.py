"""hipsc_abm_b200 -- B200-native per-step cell-update hot path of Python-hiPSC-ABM.

Import as `hipsc_abm_b200` (the repo root holds a shim module of that name, because the directory name
`python-hipsc-abm_b200` is not a valid Python identifier).
"""
from . import _lib
from ._lib import AbmError, LIB_PATH
from .graph import DeviceGraph


def __getattr__(name):
    # torch-dependent pieces are imported lazily so that `import hipsc_abm_b200` stays cheap
    if name in ("ColonyEngine", "GraphTensors", "SortState"):
        from . import engine
        return getattr(engine, name)
    if name in ("B200Simulation", "ColonySimulation", "record_time"):
        from . import simulation
        return getattr(simulation, name)
    if name == "B200CellMethods":
        from . import cell_methods
        return cell_methods.B200CellMethods
    raise AttributeError(name)


__all__ = ["AbmError", "LIB_PATH", "DeviceGraph", "ColonyEngine", "GraphTensors", "SortState", "B200Simulation",
           "ColonySimulation", "B200CellMethods", "record_time"]
