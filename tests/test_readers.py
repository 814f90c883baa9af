"""Host-side readers of the reference's binary mesh files: .re2 (core/reader_re2.f), .ma2
(core/map2.f:712-941) and assign_gllnid (core/map2.f:943-1026).  No GPU involved: these C-ABI entry points are plain host
code, so they are exercised in the CPU suite.

Pins: (i) tests/golden/bp5_fixture.npz -- the contents of the reference's own examples/bp5/bp5.{re2,ma2} as parsed by the
committed generator tests/golden/make_golden.py; the test writes files in the same on-disk format (both word sizes, both
byte orders) and reads them back through the library; where /root/reference exists the real files are read too.
(ii) assign_gllnid against the reference's own routine (tests/golden/ref_golden.npz, map/*)."""
import os
import struct

import numpy as np
import pytest

import refcases

FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bp5_fixture.npz"))
REF_BP5 = "/root/reference/examples/bp5"


@pytest.fixture(scope="module")
def nek():
    from nek5000_b200 import nek as N
    return N


def write_re2(path, xc, yc, zc, bc_elem, bc_face, bc_type, version=2, big_endian=False, curves=()):
    """The on-disk layout of core/reader_re2.f: 80-byte header, endian tag, mesh records, curve count, one BC section."""
    nel = len(xc)
    e = ">" if big_endian else "<"
    wd = "f4" if version == 1 else "f8"
    with open(path, "wb") as f:
        f.write(f"#v00{version}{nel:9d}{3:3d}{nel:9d} this is the hdr".ljust(80).encode())
        f.write(struct.pack(e + "f", 6.54321))
        rec = np.zeros((nel, 25), dtype=e + wd)
        rec[:, 0] = 0.0                        # group
        rec[:, 1:9], rec[:, 9:17], rec[:, 17:25] = xc, yc, zc
        f.write(rec.tobytes())
        cnt = (lambda n: struct.pack(e + "i", n)) if version == 1 else (lambda n: np.array([n], dtype=e + "f8").tobytes())
        f.write(cnt(len(curves)))              # curved sides: element, side, curve(5), ccurve
        for (ce, cs, cv, cc) in curves:
            if version == 1:
                f.write(struct.pack(e + "ii", int(ce), int(cs)))
                f.write(np.asarray(cv, dtype=e + "f4").tobytes())
                f.write(cc.ljust(4).encode())
            else:
                f.write(np.array([ce, cs, *cv], dtype=e + "f8").tobytes())
                f.write(cc.ljust(8).encode())
        f.write(cnt(len(bc_elem)))
        for k in range(len(bc_elem)):
            if version == 1:
                f.write(struct.pack(e + "ii", int(bc_elem[k]), int(bc_face[k])))
                f.write(np.arange(5, dtype=e + "f4").tobytes())
                f.write(str(bc_type[k]).ljust(4).encode())
            else:
                f.write(np.array([bc_elem[k], bc_face[k], 0.5, 1.5, 2.5, 3.5, 4.5], dtype=e + "f8").tobytes())
                f.write(str(bc_type[k]).ljust(8).encode())


def write_ma2(path, hdr, leaf, vertex, big_endian=False):
    e = ">" if big_endian else "<"
    with open(path, "wb") as f:
        f.write(("#v001" + "".join(f"{int(v):12d}" for v in hdr)).ljust(132).encode())
        f.write(struct.pack(e + "f", 6.54321))
        rec = np.concatenate([leaf[:, None], vertex], axis=1).astype(e + "i4")
        f.write(rec.tobytes())


@pytest.mark.parametrize("version,big", [(2, False), (2, True), (1, False), (1, True)])
def test_re2_round_trip_of_the_bp5_fixture(nek, tmp_path, version, big):
    p = str(tmp_path / "a.re2")
    write_re2(p, FIX["xc"], FIX["yc"], FIX["zc"], FIX["bc_elem"], FIX["bc_face"], FIX["bc_type"], version, big)
    info = nek.re2_info(p)
    assert info == dict(nelgt=1000, ldim=3, nelgv=1000, wdsize=4 if version == 1 else 8, ncurve=0, nbc=[600])
    xc, yc, zc, grp = nek.re2_read_mesh(p)
    if version == 1:      # 4-byte words: the file holds float32 roundings
        for a, b in ((xc, FIX["xc"]), (yc, FIX["yc"]), (zc, FIX["zc"])):
            assert np.array_equal(a, b.astype(np.float32).astype(np.float64))
    else:
        assert np.array_equal(xc, FIX["xc"]) and np.array_equal(yc, FIX["yc"]) and np.array_equal(zc, FIX["zc"])
    assert not grp.any()
    # a rank reads only its own slice (ragged: 37 elements from element 411)
    xs, ys, zs, _ = nek.re2_read_mesh(p, 411, 37)
    assert np.array_equal(xs, xc[411:448]) and np.array_equal(zs, zc[411:448])
    xe, _, _, _ = nek.re2_read_mesh(p, 1000, 0)
    assert xe.shape == (0, 8)
    cbc, bc = nek.re2_read_bc(p, 0)
    want = np.full((1000, 6), b"   ", dtype="S3")
    want[FIX["bc_elem"] - 1, FIX["bc_face"] - 1] = [s.encode() for s in FIX["bc_type"]]
    assert np.array_equal(cbc, want)
    assert (cbc == b"v  ").sum() == 600                                  # the six sides of the 10^3 box
    k = 17
    ref_bl = np.arange(5.0) if version == 1 else np.array([0.5, 1.5, 2.5, 3.5, 4.5])
    assert np.array_equal(bc[FIX["bc_elem"][k] - 1, FIX["bc_face"][k] - 1], ref_bl)


@pytest.mark.parametrize("version,big", [(2, False), (2, True), (1, False), (1, True)])
def test_re2_curved_sides(nek, tmp_path, version, big):
    """readp_re2_curve / buf_to_curve (reader_re2.f:160-290,473-497): curve(5,12,nelgt), ccurve(12,nelgt); the sections after
    the curve records (boundary conditions) are still found."""
    p = str(tmp_path / "c.re2")
    curves = [(3, 1, (0.5, 0.25, 0.0, 0.0, 0.0), "C"), (3, 12, (1.0, 2.0, 3.0, 0.0, 0.0), "m"), (1000, 5, (0.0, 0.0, 0.0, 1.5, 0.0), "s"),
              (412, 7, (-0.75, 0.0, 0.0, 0.0, 0.0), "C")]
    write_re2(p, FIX["xc"], FIX["yc"], FIX["zc"], FIX["bc_elem"], FIX["bc_face"], FIX["bc_type"], version, big, curves)
    info = nek.re2_info(p)
    assert info["ncurve"] == 4 and info["nbc"] == [600]
    cc, cv = nek.re2_read_curves(p)
    assert cc.shape == (1000, 12) and cv.shape == (1000, 12, 5)
    assert (cc != b" ").sum() == 4
    for (ce, cs, v, t) in curves:
        assert cc[ce - 1, cs - 1] == t.encode() and np.array_equal(cv[ce - 1, cs - 1], np.asarray(v))   # exact in f4 too
    assert not cv[cc == b" "].any()
    cbc, _ = nek.re2_read_bc(p, 0)
    assert (cbc == b"v  ").sum() == 600
    # a record naming a side outside 1..12 is refused
    from nek5000_b200.nek import NekbError
    write_re2(p, FIX["xc"], FIX["yc"], FIX["zc"], FIX["bc_elem"], FIX["bc_face"], FIX["bc_type"], version, big, [(3, 13, (0,) * 5, "C")])
    with pytest.raises(NekbError):
        nek.re2_read_curves(p)


@pytest.mark.parametrize("big", [False, True])
def test_ma2_round_trip_and_numbering(nek, tmp_path, big):
    p = str(tmp_path / "a.ma2")
    write_ma2(p, FIX["ma2_header"], FIX["leaf"], FIX["vertex"], big)
    hdr, leaf, vertex = nek.ma2_read(p)
    assert np.array_equal(hdr, FIX["ma2_header"]) and np.array_equal(leaf, FIX["leaf"])
    assert vertex.dtype == np.int64 and np.array_equal(vertex, FIX["vertex"])
    _, l2, v2 = nek.ma2_read(p, 8, 990, 10)
    assert np.array_equal(l2, leaf[990:]) and np.array_equal(v2, vertex[990:])
    # the vertex ids feed setvert3d directly: 1331 distinct corners -> (10*7+1)^3 global nodes at lx1 = 8, of which the
    # 6^3 interior nodes of every element are not numbered (id 0)
    glo, ngv = nek.setvert3d(8, 1000, vertex)
    assert len(np.unique(vertex)) == 1331 and len(np.unique(glo)) == 71 ** 3 - 1000 * 6 ** 3 + 1


@pytest.mark.parametrize("version,big", [(1, False), (1, True), (2, False)])
def test_co2_round_trip(nek, tmp_path, version, big):
    """read_con (map2.f:338-473): header '#v001' + 3 x i12 or '#v002' list-directed, endian tag, (element id, vertex ids)
    int32 records; the vertex ids feed setvert3d exactly like those of a .ma2."""
    e = ">" if big else "<"
    p = str(tmp_path / "a.co2")
    nel, vertex = 1000, FIX["vertex"]
    with open(p, "wb") as f:
        hdr = f"#v001{nel:12d}{nel:12d}{8:12d}" if version == 1 else f"#v002 {nel} {nel} 8"
        f.write(hdr.ljust(132).encode())
        f.write(struct.pack(e + "f", 6.54321))
        f.write(np.concatenate([np.arange(1, nel + 1)[:, None], vertex], axis=1).astype(e + "i4").tobytes())
    ngt, ngv, eid, v = nek.co2_read(p)
    assert (ngt, ngv) == (1000, 1000) and np.array_equal(eid, np.arange(1, 1001)) and np.array_equal(v, vertex) and v.dtype == np.int64
    _, _, e2, v2 = nek.co2_read(p, 8, 411, 37)
    assert np.array_equal(e2, eid[411:448]) and np.array_equal(v2, vertex[411:448])
    glo, _ = nek.setvert3d(8, 1000, v)
    glo_ma2, _ = nek.setvert3d(8, 1000, FIX["vertex"])
    assert np.array_equal(glo, glo_ma2)
    from nek5000_b200.nek import NekbError
    with pytest.raises(NekbError):
        nek.co2_read(p, 4)                      # 'Number of vertices do not match!' (map2.f:443)
    with pytest.raises(NekbError):
        nek.co2_read(p, 8, 990, 20)


def test_reader_errors_are_loud(nek, tmp_path):
    from nek5000_b200.nek import NekbError
    with pytest.raises(NekbError):
        nek.re2_info(str(tmp_path / "missing.re2"))
    bad = tmp_path / "bad.re2"
    bad.write_bytes(b"#v009" + b" " * 200)
    with pytest.raises(NekbError):
        nek.re2_info(str(bad))
    p = str(tmp_path / "t.re2")
    write_re2(p, FIX["xc"], FIX["yc"], FIX["zc"], FIX["bc_elem"], FIX["bc_face"], FIX["bc_type"])
    with open(p, "r+b") as f:
        f.truncate(84 + 200 * 500)
    with pytest.raises(NekbError):
        nek.re2_read_mesh(p, 0, 1000)
    with pytest.raises(NekbError):
        nek.re2_read_mesh(p, 990, 20)


def test_assign_gllnid_matches_the_reference(nek):
    g = refcases.load_golden()["map"]
    for npr in refcases.MAP_NP:
        got = nek.assign_gllnid(FIX["leaf"], None, npr)
        assert np.array_equal(got, g[f"gllnid_np{npr}"]), npr          # integer: bit-exact
    # local element order = ascending global id on each rank (map2.f:233-236): the partition of 8 ranks is balanced
    got = nek.assign_gllnid(FIX["leaf"], None, 8)
    assert np.bincount(got).tolist() == [125] * 8


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BP5, "bp5.re2")), reason="reference tree not present")
def test_the_reference_files_themselves(nek):
    xc, yc, zc, _ = nek.re2_read_mesh(os.path.join(REF_BP5, "bp5.re2"))
    assert np.array_equal(xc, FIX["xc"]) and np.array_equal(yc, FIX["yc"]) and np.array_equal(zc, FIX["zc"])
    cbc, _ = nek.re2_read_bc(os.path.join(REF_BP5, "bp5.re2"), 0)
    assert (cbc == b"v  ").sum() == 600
    hdr, leaf, vertex = nek.ma2_read(os.path.join(REF_BP5, "bp5.ma2"))
    assert np.array_equal(leaf, FIX["leaf"]) and np.array_equal(vertex, FIX["vertex"]) and np.array_equal(hdr, FIX["ma2_header"])
    # examples/turbChannel (BASELINE config 5): the box the oracle generates for tests/refcases.channel_case IS the file's mesh
    # before usrdat -- same vertices, and the same identification of corners as the genmap vertex ids in turbChannel.ma2
    import oracle
    tc = "/root/reference/examples/turbChannel/turbChannel"
    if os.path.exists(tc + ".re2"):
        x, y, z, _ = nek.re2_read_mesh(tc + ".re2")
        c = oracle.Case(16, 12, 8, nx=4, periodic=(1, 0, 1), rescale=False)
        E = c.nel
        for a, b in ((c.xc, x), (c.yc, y), (c.zc, z)):
            assert np.abs(a.reshape(8, E, order="F").T - b).max() <= 2e-16
        cbc, _ = nek.re2_read_bc(tc + ".re2", 0)
        assert (cbc == b"W  ").sum() == 2 * 16 * 8 and (cbc == b"P  ").sum() == 2 * (12 * 8 + 16 * 12)
        _, leaf, vertex = nek.ma2_read(tc + ".ma2")
        for npr in (4, 8):                          # turbChannel.par: minNumProcesses = 4; the 8-GPU box of BASELINE config 5
            part = nek.assign_gllnid(leaf, None, npr)
            assert np.bincount(part, minlength=npr).tolist() == [1536 // npr] * npr
        mine = c.vertex.reshape(E, 8)
        assert len(np.unique(vertex)) == len(np.unique(mine)) == 16 * 13 * 8
        fwd = dict(zip(vertex.ravel().tolist(), mine.ravel().tolist()))
        assert all(fwd[a] == b for a, b in zip(vertex.ravel().tolist(), mine.ravel().tolist()))
    # short_tests/ethier (a curved ball mesh from exo2nek): 32 elements, 96 curved-side records
    et = "/root/reference/short_tests/ethier/ethier.re2"
    if os.path.exists(et):
        cc, cv = nek.re2_read_curves(et)
        info = nek.re2_info(et)
        assert info["ncurve"] == (cc != b" ").sum() == 96
        raw = open(et, "rb").read()
        off = 84 + 32 * 25 * 8 + 8                                      # header, endian tag, mesh records, curve count
        rec = np.frombuffer(raw[off:off + 96 * 64], dtype="<f8").reshape(96, 8)
        for r, row in enumerate(rec):
            e_, s_ = int(row[0]), int(row[1])
            assert cc[e_ - 1, s_ - 1] == raw[off + r * 64 + 56:off + r * 64 + 57] and np.array_equal(cv[e_ - 1, s_ - 1], row[2:7])
    for name in ("short_tests/ethier/ethier", "examples/turbChannel/turbChannel"):
        p = os.path.join("/root/reference", name + ".re2")
        if os.path.exists(p):
            info = nek.re2_info(p)
            x, y, z, _ = nek.re2_read_mesh(p)
            assert info["ldim"] == 3 and x.shape == (info["nelgt"], 8) and np.isfinite(x).all()
            assert sum(info["nbc"]) > 0
