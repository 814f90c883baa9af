"""nek5000_b200 -- B200-native (sm_100a, FP64, hand-written CUDA) drop-in for Nek5000's Helmholtz /
gather-scatter / PCG hot path.  See DESIGN.md, INTEGRATION.md and include/nekb200.h.

    nek5000_b200.nek   host-side mirror of the reference call signatures (axhelm, cggo, dssum, ...)
    nek5000_b200.bp5   BP5 benchmark driver (examples/bp5/bp5.usr)
    nek5000_b200.build nvcc build of libnekb200.so
"""
from . import build  # noqa: F401
from ._lib import NekbError, declared_symbols, lib  # noqa: F401
from . import nek  # noqa: F401
from . import bp5  # noqa: F401
