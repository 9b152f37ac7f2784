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


LIB_FAST = os.path.join(HERE, "libaxisem_oracle_fast.so")


def build_fast(force: bool = False) -> str:
    """-O3 / FMA build of the same source: the CPU baseline of bench.py (timing only)."""
    src = os.path.join(HERE, "axisem_oracle.c")
    if force or not os.path.exists(LIB_FAST) or os.path.getmtime(LIB_FAST) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libaxisem_oracle_fast.so"], stdout=subprocess.DEVNULL)
    return LIB_FAST


_lib = None
_lib_fast = None


def load_fast() -> Library:
    global _lib_fast
    if _lib_fast is None:
        _lib_fast = Library(build_fast(), "axo_")
    return _lib_fast


def load() -> Library:
    global _lib
    if _lib is None:
        _lib = Library(build(), "axo_")
    return _lib


def make_loop(prob) -> TimeLoop:
    return TimeLoop(load(), prob)


HOST_EXE = os.path.join(HERE, "axisem_host_oracle")


def build_host(force: bool = False) -> str:
    """The native (C++) host of the time-loop seam, axisem_b200/hostcxx/, compiled against this
    oracle's implementation of the header (prefix axo_): lets the CPU tests drive the host
    logic without a GPU."""
    build()
    src_dir = os.path.join(HERE, "..", "axisem_b200", "hostcxx")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cpp", ".hpp"))]
    stale = (not os.path.exists(HOST_EXE)
             or os.path.getmtime(HOST_EXE) < max([os.path.getmtime(f) for f in srcs] + [os.path.getmtime(LIB)]))
    if force or stale:
        cpp = [os.path.join(src_dir, f) for f in ("main.cpp", "time_loop.cpp", "modules.cpp", "meshdb.cpp", "precomp.cpp", "mapping.cpp",
                                                        "background_models.cpp", "receivers.cpp", "rundir.cpp")]
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(HERE, "..", "include"),
                               "-DAXB_PREFIX=axo_", "-o", HOST_EXE] + cpp
                              + ["-L" + HERE, "-laxisem_oracle", "-Wl,-rpath," + HERE])
    return HOST_EXE
