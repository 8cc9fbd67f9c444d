"""CPU, world_size 2 over gloo: the host-side multi-block logic (neighbour maps, block boundary types,
message routing/ordering of ParallelContext.exchange) reproduces the periodic / interior halos of a
global array.  Pack/unpack are emulated with NumPy slicing here (the CUDA pack/unpack kernels are
covered by tests/test_gpu_parity.py::test_pack_unpack_faces_roundtrip and tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import ROOT

FACES = ("east", "west", "north", "south", "top", "bottom")
AX = {"east": 0, "west": 0, "north": 1, "south": 1, "top": 2, "bottom": 2}


def _worker(rank, world, split, cells, bc_kind, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jaxfluids_b200.domain_information import DomainInformation
    from jaxfluids_b200.parallel import ParallelContext
    nh = 3
    di = DomainInformation(cells, ((0.0, 1.0),) * 3, split, nh)
    ctx = ParallelContext(di, rank, world)
    bc = {f: (bc_kind if cells[AX[f]] > 1 else "INACTIVE") for f in FACES}
    types = ctx.block_boundary_types(bc)
    nbrs = ctx.neighbors(bc)
    # global field with a unique value per cell; local block with halos
    glob = np.arange(np.prod(cells), dtype=np.float64).reshape(cells)
    sl = di.block_slices(rank)
    n = di.device_number_of_cells
    loc = np.full(tuple(m + 2 * nh if N > 1 else 1 for m, N in zip(n, cells)), -1.0)
    inter = tuple(slice(nh, -nh) if N > 1 else slice(None) for N in cells)
    loc[inter] = glob[sl]
    send, recv = {}, {}
    for f in nbrs:
        ax = AX[f]
        idx = list(inter)
        idx[ax] = slice(-2 * nh, -nh) if f in ("east", "north", "top") else slice(nh, 2 * nh)
        send[f] = torch.from_numpy(np.ascontiguousarray(loc[tuple(idx)]).ravel().copy())
        recv[f] = torch.empty_like(send[f])
    for r in ctx.exchange(nbrs, send, recv):
        r.wait()
    ok = True
    for f in nbrs:
        ax = AX[f]
        # expected: the neighbour's interior layers adjacent to the shared face, from the global array
        lo, hi = sl[ax].start, sl[ax].stop
        N = cells[ax]
        if f in ("east", "north", "top"):
            rng = [(hi + k) % N for k in range(nh)]
        else:
            rng = [(lo - nh + k) % N for k in range(nh)]
        idx = [np.arange(s.start, s.stop) for s in sl]
        idx[ax] = np.array(rng)
        expect = glob[np.ix_(*idx)].ravel()
        ok = ok and np.array_equal(recv[f].numpy(), expect)
    q.put((rank, ok, types, sorted(nbrs.items())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("split,cells,bc", [((2, 1, 1), (12, 8, 6), "PERIODIC"), ((1, 2, 1), (6, 12, 8), "SYMMETRY"),
                                            ((1, 1, 2), (6, 8, 12), "PERIODIC"), ((2, 1, 1), (16, 1, 1), "ZEROGRADIENT")])
def test_two_rank_exchange_routing(split, cells, bc):
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = 29600 + (os.getpid() + sum(cells)) % 300
    procs = [ctxmp.Process(target=_worker, args=(r, 2, split, cells, bc, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    ax = [i for i in range(3) if split[i] == 2][0]
    hi_face, lo_face = FACES[2 * ax], FACES[2 * ax + 1]
    for rank, ok, types, nbrs in res:
        assert ok, f"rank {rank}: wrong halo payload"
        if bc == "PERIODIC":
            assert types[hi_face] == types[lo_face] == "NEIGHBOR" and len(nbrs) == 2
        else:
            # physical boundary on the outer side, neighbour on the inner side
            inner = hi_face if rank == 0 else lo_face
            outer = lo_face if rank == 0 else hi_face
            assert types[inner] == "NEIGHBOR" and types[outer] == bc and len(nbrs) == 1


def test_edge_widening_masks_of_shared_faces():
    """Which transverse halos a shared face's slab is widened over on the dissipative path (runtime.
    BlockRuntime._edge_ext_mask): all halos of the axes exchanged earlier, the PHYSICAL halos of the later ones.
    bit 0/1 = low/high side of the slower transverse axis, bit 2/3 = of the faster one."""
    from types import SimpleNamespace
    from jaxfluids_b200.runtime import BlockRuntime
    # block (0,0,0) of a (2,2,1) split, SYMMETRY walls: east + north shared, west/south/top/bottom physical
    rt = SimpleNamespace(cfg=SimpleNamespace(cells=(16, 16, 16)),
                         bc_block={"east": "NEIGHBOR", "west": "SYMMETRY", "north": "NEIGHBOR", "south": "SYMMETRY",
                                   "top": "SYMMETRY", "bottom": "SYMMETRY"})
    m = lambda f: BlockRuntime._edge_ext_mask(rt, f)
    # x face: transverse (y, z) are exchanged later -> only their physical sides: south (y low), bottom, top
    assert m("east") == (1 << 0) | (1 << 2) | (1 << 3)
    # y face: transverse (x, z): x was exchanged earlier -> both sides; z physical on both sides
    assert m("north") == 0b1111
    # 2-D block (z inactive): y face widened over both x sides only
    rt2 = SimpleNamespace(cfg=SimpleNamespace(cells=(16, 16, 1)),
                          bc_block={"east": "ZEROGRADIENT", "west": "NEIGHBOR", "north": "NEIGHBOR", "south": "NEIGHBOR",
                                    "top": "INACTIVE", "bottom": "INACTIVE"})
    assert BlockRuntime._edge_ext_mask(rt2, "north") == 0b0011
    assert BlockRuntime._edge_ext_mask(rt2, "west") == 0           # y sides are shared and exchanged later
