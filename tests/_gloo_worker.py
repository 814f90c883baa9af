"""Worker of tests/test_host_logic.py: one rank of a world_size-N gloo job exercising the multi-rank HOST logic
of libnekb200 (no GPU): distributed setvert3d (gbtuple_rank8 with real tuple exchange) and the shared-id
rendezvous of gs_setup.  Writes its results to an .npz for the parent to check against the oracle."""
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nek5000_b200 import nek  # noqa: E402


def main():
    out = sys.argv[1]
    nelx, nely, nelz, nx = (int(a) for a in sys.argv[2:6])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nek.set_transport_torch()
    # slab partition in z of the global box, ascending global element id inside a rank (map2.f:233-236)
    import oracle
    case = oracle.Case(nelx, nely, nelz, nx=nx, np_ranks=world)
    nel = case.nel
    per = nel // world
    lo, hi = rank * per, (nel if rank == world - 1 else (rank + 1) * per)
    vertex = case.vertex.reshape(nel, 8)[lo:hi].copy()
    glo, ngv = nek.setvert3d(nx, hi - lo, vertex, world)
    uniq = np.unique(glo[glo != 0])
    peers, off, ids = nek.gs_discover(uniq)
    np.savez(out + f".{rank}.npz", glo=glo, ngv=ngv, lo=lo, hi=hi, peers=peers, off=off, ids=ids)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
