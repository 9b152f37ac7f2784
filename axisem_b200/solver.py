"""Host-side mirror of the reference's solver seam for the B200 library.

`time_loop(problem)` is the stand-in for `call time_loop` (SOLVER/main.f90:92): it hands
the pre-computed module arrays to libaxisem_b200.so (CUDA, sm_100a) through the C ABI of
include/axisem_b200.h and returns a :class:`axisem_b200.capi.TimeLoop`.

There is no CPU path: if the CUDA library is missing, or no GPU is visible, this module
raises — it never falls back to the oracle or to numpy.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

from .capi import AxbError, Library, TimeLoop, connect_local, run_group  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_FAST = os.path.join(HERE, "libaxisem_b200.so")
LIB_STRICT = os.path.join(HERE, "libaxisem_b200_strict.so")

_libs = {}


def load_library(strict: bool = False) -> Library:
    """Load the product library (`strict=True`: the -fmad=false build whose results are
    bit-identical to the oracle; used by the parity tests)."""
    path = LIB_STRICT if strict else LIB_FAST
    if not strict and os.environ.get("AXB_LIBRARY"):
        path = os.environ["AXB_LIBRARY"]       # developer override: a differently tuned build
    if path not in _libs:
        if not os.path.exists(path):
            raise AxbError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  axisem_b200 has no CPU fallback.")
        _libs[path] = Library(path, "axb_")
    return _libs[path]


def time_loop(problem, device: int = 0, strict: bool = False) -> TimeLoop:
    """Create the device-resident time loop for one rank's `Problem`."""
    return TimeLoop(load_library(strict), problem, device=device)


def time_loop_group(problems: Sequence, devices: Optional[Sequence[int]] = None,
                    strict: bool = False):
    """All ranks of a domain-decomposed run inside one process (single-GPU loop-back or
    one process driving several GPUs); peers are wired with direct device pointers."""
    lib = load_library(strict)
    devices = devices or [0] * len(problems)
    loops = [TimeLoop(lib, p, device=d) for p, d in zip(problems, devices)]
    connect_local(lib, loops)
    return lib, loops
