"""B200-native SHOT/FPFH recognition hot path.

The directory name is not a Python identifier; load it with
    importlib.import_module("3d-object-detection-of-industrial-joints_b200")
`synth` is pure numpy; `binding` is the ctypes mirror of include/b200reg.h and raises loudly when
libb200reg.so (the CUDA library) is missing — there is no CPU fallback.
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
