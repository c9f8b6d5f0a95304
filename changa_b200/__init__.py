"""changa_b200 -- B200-native gravity force evaluation behind ChaNGa's GPU entry points.

  lib        ctypes view of the C ABI (include/changa_b200_api.h)
  hostcuda   Python mirror of the reference's HostCUDA.h / EwaldCUDA.h interface
  workloads  synthetic inputs in the serialized shape the entry points consume
  csrc/      CUDA kernels (sm_100a) + the host runtime, built in-tree
"""
from . import lib  # noqa: F401
