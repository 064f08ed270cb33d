"""B200-native drop-in for the time-integration hot path of dolfin_navier_scipy
(same module names as `dolfin_navier_scipy/__init__.py`)"""
from . import dolfin_to_sparrays
from . import data_output_utils
from . import problem_setups
from . import stokes_navier_utils
from . import time_int_utils
from . import residual_checks

__all__ = ["dolfin_to_sparrays",
           "data_output_utils",
           "stokes_navier_utils",
           "problem_setups",
           "time_int_utils",
           "residual_checks"]
