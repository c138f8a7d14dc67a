"""mars_ode_physics_b200 -- B200-native batched implementation of the hot path behind
WorldPhysics::stepTheWorld of mars_ode_physics (reference src/WorldPhysics.cpp:349-421).

The product is the CUDA shared library ``lib/libmars_ode_b200.so`` (C ABI in
``include/mars_ode_b200.h``) plus the C++ facade in ``facade/`` that mirrors the reference's
WorldPhysics / Frame / Joint classes.  This Python package only wraps the C ABI for tests and
benchmarks.  There is no CPU fallback.
"""
from . import capi  # noqa: F401
from .arena import Arena  # noqa: F401
from .capi import ALL_WORLDS, B2PError  # noqa: F401

__all__ = ["Arena", "capi", "ALL_WORLDS", "B2PError"]
