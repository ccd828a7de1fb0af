"""turbulent_lbm_multigpu_b200 -- B200-native D3Q19 alpha/beta lattice-Boltzmann time step.

Host-side mirror (Python) of the reference's solver/communication surface
(CLbmSolver, CController, CManager, CComm, CDomain, CConfiguration) over the C ABI of
``lib/liblbm_b200.so`` (hand-written sm_100a CUDA, see csrc/).  Importing the package does
not load the CUDA library; creating a solver does, and fails loudly when it is missing.
"""
from .domain import CComm, CDomain  # noqa: F401
from .skeleton import (FLAG_FLUID, FLAG_GHOST_LAYER, FLAG_OBSTACLE,  # noqa: F401
                       FLAG_VELOCITY_INJECTION, LBM_UNITS, compute_parameters)

__version__ = "0.1.0"
