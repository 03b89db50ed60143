"""Loads the host EMULATION build of the product sources (tests/emu/libb200rate_emu.so).

TEST INFRASTRUCTURE: same C++ orchestration and the same CTA programs as libb200rate.so, compiled with
-DB200RATE_EMU so a kernel launch is a serial loop. Lets the CPU-only test tier check plan, bookkeeping and
every index computation of the CUDA path against the oracle. It is never used by the product package."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from foo_dsp_resampler_b200 import _capi  # noqa: E402

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libb200rate_emu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", EMU_DIR], stdout=subprocess.DEVNULL)
        _lib = _capi.bind(EMU_SO)
    return _lib
