"""Writes tests/golden/channel_partition.npz: BASELINE config 5's element partition, from the reference's own files.

    python tests/golden/gen_channel_partition.py          (needs /root/reference and oracle/_ref; run in the authoring image)

Contents: `leaf` (1536 RSB leaves) and `vertex` (1536 x 8 genmap vertex ids) exactly as core/map2.f:712-941 reads them from
/root/reference/examples/turbChannel/turbChannel.ma2 (parsed by the library's host reader, which tests/test_readers.py pins
against the file format), and `gllnid_np{2,4,8}` = the rank of every global element as the reference's own assign_gllnid
(core/map2.f:943-1026, executed from oracle/_ref) computes it.  /root/reference does not exist on the GPU box, so the
multi-GPU channel run (tests/_mgpu_channel_worker.py) reads this fixture instead of the .ma2 file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
MA2 = "/root/reference/examples/turbChannel/turbChannel.ma2"


def main():
    from nek5000_b200 import nek
    from oracle.ref import Ref
    hdr, leaf, vertex = nek.ma2_read(MA2)
    R = Ref(8, 8, 64, fresh=True)
    out = dict(ma2_header=hdr, leaf=leaf.astype(np.int32), vertex=vertex.astype(np.int64))
    for npr in (2, 4, 8):
        g, scratch = leaf.astype(np.int32).copy(), np.zeros(len(leaf), dtype=np.int32)
        R.call("assign_gllnid", g, scratch, len(g), len(g), npr)
        assert np.array_equal(g, nek.assign_gllnid(leaf, None, npr))      # the library's routine agrees (tests/test_readers.py)
        out[f"gllnid_np{npr}"] = g
        print(npr, np.bincount(g))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "channel_partition.npz"), **out)


if __name__ == "__main__":
    main()
