"""fmwr_b200: B200-native factorization-machine engine behind FMwR's R API.

The compute path is the CUDA library fmwr_b200/libfmwr_b200.so (C ABI: include/fmwr_b200.h).
`fmwr_b200.api` mirrors the reference's R interface (fm.train, predict.FM, fm.update, fm.track,
*.solver, *.control) on top of that ABI for hosts without R.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
