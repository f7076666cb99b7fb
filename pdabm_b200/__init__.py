"""Import alias for the package directory ``flamegpu2-prisoners-dilemma-abm_b200/``.

The directory name the project layout prescribes contains hyphens and cannot be an
import name, so this shim makes ``import pdabm_b200.<module>`` resolve into it.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg = _os.path.normpath(_os.path.join(_here, "..", "flamegpu2-prisoners-dilemma-abm_b200"))
__path__.insert(0, _pkg)  # noqa: F821  (package attribute)

from .sim import Simulation, SimConfig, PdabmError  # noqa: E402,F401
from . import model  # noqa: E402,F401
from .decomp import DecomposedWorld, LocalComm, DistComm, PeerComm, best_comm  # noqa: E402,F401
