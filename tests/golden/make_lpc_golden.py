"""Generates tests/golden/lpc_reference.npz from the COMPILED REFERENCE (oracle/_ref/libref_lpc.so = the unmodified
/root/reference/lpc/lpc.cpp and util.h, see oracle/Makefile). Run in the build container:

    python tests/golden/make_lpc_golden.py

The reference ships no vectors for its track-edge extrapolation, so these pin the restatement (oracle/lpc_oracle.c) and
the CUDA kernels (csrc/lpc.cu) to outputs of the reference itself: for every (signal kind, shape) of tests/lpclib.py
the extrapolated frames, and the edge lengths of a grid of rate pairs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import lpclib  # noqa: E402

RATES = (8000, 11025, 16000, 22050, 32000, 44100, 48000, 48001, 88200, 96000, 176400, 192000, 352800, 384000, 768000,
         44101, 12345)


def main():
    out = {}
    for kind in lpclib.KINDS:
        for si, (n, nch, order, bk, fw) in enumerate(lpclib.SHAPES):
            if n > 5000 and kind not in (0, 5):
                continue                                     # keep the fixture small: the long base for two kinds only
            buf = np.zeros((bk + n + fw, nch), np.float32)
            buf[bk:bk + n] = lpclib.signal(kind, n, nch)
            lpclib.ref_extrapolate2(buf, bk, n, bk, fw, order)
            out["k%d_s%d_bkwd" % (kind, si)] = buf[:bk].copy()
            out["k%d_s%d_fwd" % (kind, si)] = buf[bk + n:].copy()
    edges = np.array([[a, b, *lpclib.ref_edge_lengths(a, b)] for a in RATES for b in RATES], dtype=np.int64)
    out["edge_lengths"] = edges
    path = os.path.join(HERE, "lpc_reference.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
