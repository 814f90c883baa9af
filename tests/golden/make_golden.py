"""Generates tests/golden/bp5_fixture.npz from the reference's own BP5 mesh fixtures.

Run in the authoring container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Sources (binary data files, not source code):
  examples/bp5/bp5.re2  -- genbox output, '#v002' header (80 B) + endian test float + 25 float64 per element
                           (group, x(8), y(8), z(8) in preprocessor corner order), format per
                           core/reader_re2.f:430-471,543-639
  examples/bp5/bp5.ma2  -- genmap output, 132 B header + endian test float + 9 int32 per element
                           (RSB leaf, 8 vertex ids in symmetric corner order), format per core/map2.f:755-830
"""
import os
import struct

import numpy as np

REF = "/root/reference/examples/bp5"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bp5_fixture.npz")


def main():
    b = open(os.path.join(REF, "bp5.re2"), "rb").read()
    hdr = b[:80].decode()
    assert hdr.startswith("#v002"), hdr
    nel = int(hdr[5:14])
    assert abs(struct.unpack("<f", b[80:84])[0] - 6.54321) < 1e-5
    rec = np.frombuffer(b, dtype="<f8", count=25 * nel, offset=84).reshape(nel, 25)
    xc, yc, zc = rec[:, 1:9].copy(), rec[:, 9:17].copy(), rec[:, 17:25].copy()
    off = 84 + 200 * nel
    ncurve = int(np.frombuffer(b, dtype="<f8", count=1, offset=off)[0])
    assert ncurve == 0
    nbc = int(np.frombuffer(b, dtype="<f8", count=1, offset=off + 8)[0])
    bc = np.frombuffer(b, dtype="<f8", count=8 * nbc, offset=off + 16).reshape(nbc, 8)
    bc_elem = bc[:, 0].astype(np.int32)
    bc_face = bc[:, 1].astype(np.int32)
    bc_type = np.array([bc[i, 7].tobytes()[:3].decode() for i in range(nbc)])

    m = open(os.path.join(REF, "bp5.ma2"), "rb").read()
    mh = m[:132].decode().split()
    assert mh[0] == "#v001" and int(mh[1]) == nel
    assert abs(struct.unpack("<f", m[132:136])[0] - 6.54321) < 1e-5
    v = np.frombuffer(m, dtype="<i4", offset=136).reshape(nel, 9)
    np.savez_compressed(OUT, xc=xc, yc=yc, zc=zc, leaf=v[:, 0].copy(), vertex=v[:, 1:].copy(),
                        ma2_header=np.array([int(x) for x in mh[1:8]], dtype=np.int64),
                        bc_elem=bc_elem, bc_face=bc_face, bc_type=bc_type)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
