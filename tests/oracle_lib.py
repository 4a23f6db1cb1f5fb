"""Shim: the ctypes wrapper of the CPU oracle lives in oracle/cpu.py (TEST INFRASTRUCTURE)."""
from oracle.cpu import *  # noqa: F401,F403
from oracle.cpu import FQ, FR, Oracle, build_oracle, ints_to_limbs, limbs_to_ints, rand_fr  # noqa: F401
