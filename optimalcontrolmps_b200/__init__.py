"""optimalcontrolmps_b200 -- B200-native (sm_100a) engine for the hot path of fskovbo/OptimalControlMPS.

The package holds the CUDA sources of ``libocmps.so`` (``csrc/``), its ctypes binding (``_lib``) and a
host-side mirror of the reference's C++ interface (``api``).  It never falls back to the CPU.
"""
from .api import (Args, BH_tDMRG, BoseHubbard, Context, ControlBasis, ControlBasisFactory, DeviceMPS, IQMPS,
                  OptimalControl, SeedGenerator, SliceStore, batch_cost_gradient, overlapC, overlapC_K)
from ._lib import OcmpsError, LIB_PATH

__all__ = ["Args", "BH_tDMRG", "BoseHubbard", "Context", "ControlBasis", "ControlBasisFactory", "DeviceMPS", "IQMPS",
           "OptimalControl", "SeedGenerator", "batch_cost_gradient", "SliceStore", "overlapC", "overlapC_K", "OcmpsError", "LIB_PATH"]
