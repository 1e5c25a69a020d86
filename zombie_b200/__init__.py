"""zombie_b200: B200-native implementation of Zombie's random-walk estimator hot path.

The CUDA extension (zombie_b200/csrc -> zombie_b200/libzombie_b200.so) is loaded lazily by
zombie_b200.native.lib(); there is no CPU fallback and loading fails loudly when it is missing.
"""
from . import abi, procedural  # noqa: F401

__all__ = ["abi", "procedural"]
