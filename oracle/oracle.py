"""Loader for the CPU oracle (oracle/libaxisem_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the axisem_b200 package."""
from __future__ import annotations

import os
import subprocess

from axisem_b200.capi import Library, TimeLoop, connect_local, run_group  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libaxisem_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "axisem_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "axisem_b200.h")
    stale = (not os.path.exists(LIB)
             or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", HERE, "-B", "libaxisem_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def load() -> Library:
    global _lib
    if _lib is None:
        _lib = Library(build(), "axo_")
    return _lib


def make_loop(prob) -> TimeLoop:
    return TimeLoop(load(), prob)
