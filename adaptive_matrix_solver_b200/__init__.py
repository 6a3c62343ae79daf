"""Importable alias of the package directory ``adaptive-matrix-solver_b200/`` (a hyphen cannot be imported).

All code lives in ``adaptive-matrix-solver_b200/``; this shim only points ``__path__`` there.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "adaptive-matrix-solver_b200"))

from ._abi import load_library, library_path, MausError          # noqa: E402,F401
from .engine import MausEngine                                   # noqa: E402,F401
from .solver import GpuInverseIterateSolver                      # noqa: E402,F401
from .population import step_population, install_dropin          # noqa: E402,F401
from .candidates import Candidate, ProblemType                   # noqa: E402,F401
