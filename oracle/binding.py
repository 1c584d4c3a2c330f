"""ctypes loader of the CPU oracle (oracle/cph_oracle.cpp).

TEST INFRASTRUCTURE.  Importing this module registers the oracle with the generic call table of
constant_ph_b200.capi, so that `capi.Engine("orc")` drives it with the same method names as the
CUDA engine.  Only tests/ (through conftest.py), __graft_entry__.smoke() and bench.py's CPU legs
import it; the product package never does (tests/test_abi.py checks that).
"""
import ctypes
import os
import subprocess

from constant_ph_b200 import capi

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))


def build(native=False):
    """Compile the oracle with its own Makefile (gcc only).  native: -march=native build for timing runs."""
    subprocess.run(["make", "-C", ORACLE_DIR, "native" if native else "all"], check=True, capture_output=True)
    return os.path.join(ORACLE_DIR, "libcph_oracle_native.so" if native else "libcph_oracle.so")


def load(native=False):
    path = os.path.join(ORACLE_DIR, "libcph_oracle_native.so" if native else "libcph_oracle.so")
    src = os.path.join(ORACLE_DIR, "cph_oracle.cpp")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        path = build(native)
    return ctypes.CDLL(path)


capi.register_library("orc", load)
