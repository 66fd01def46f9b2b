"""B200-native FCIQMC walker-propagation engine (drop-in hot path for NECI).

  capi    ctypes binding of the C ABI (include/neci_gpu.h) -> libneci_gpu.so
  host    host-side mirror of the Fortran host's setup (tables, config)
  driver  the FciMCPar outer loop: iterate, reduce statistics, update shift
"""
from . import capi, host, driver  # noqa: F401
from .capi import Engine, EngineError, ST, ST_COUNT  # noqa: F401

__all__ = ["capi", "host", "driver", "Engine", "EngineError", "ST", "ST_COUNT"]
