#!/usr/bin/env python
"""Writes tests/golden/ref_golden.npz: inputs and outputs of the hot path as computed by the REFERENCE's own Fortran
statements (oracle/_ref, transpiled from /root/reference by oracle/ref_build.py -- needs /root/reference or a prebuilt
oracle/_ref).  Cases and the reference routines they run: tests/refcases.py.

    python tests/golden/gen_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refcases  # noqa: E402

if __name__ == "__main__":
    flat = refcases.reference_all()
    np.savez_compressed(refcases.GOLDEN, **flat)
    print(f"wrote {refcases.GOLDEN}: {len(flat)} arrays, {os.path.getsize(refcases.GOLDEN) / 1e6:.2f} MB")
    if "--full" in sys.argv:         # the 1536-element turbChannel mesh: needs `python oracle/ref_build.py --lelt 1536`
        full = refcases.ref_channel_full()
        np.savez_compressed(refcases.GOLDEN_CHANNEL_FULL, **full)
        print(f"wrote {refcases.GOLDEN_CHANNEL_FULL}: {len(full)} arrays, {os.path.getsize(refcases.GOLDEN_CHANNEL_FULL) / 1e6:.2f} MB")
