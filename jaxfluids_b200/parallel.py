"""Block decomposition across GPUs: one process per GPU, torch.distributed (NCCL)
for the plumbing.  Replaces the reference's jax.pmap domain decomposition
(simulation_manager.py:1206-1228) and its ppermute / pmin collectives
(halos/inner/material.py:74, time_step_size.py:152, positivity_handler.py:253-254).

Per RK stage every block exchanges the `nh` primitive layers next to each face it
shares with another block (send/recv pairs, grouped into one NCCL group); the
receiver recomputes the conservatives in its halo (halos/inner/material.py:83-88).
Per step one MAX all-reduce carries {max sum(|u_i|+c), -min rho, -min p}.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .domain_information import DomainInformation, FACES

OPPOSITE = {"east": "west", "west": "east", "north": "south", "south": "north", "top": "bottom", "bottom": "top"}
FACE_ID = {f: i for i, f in enumerate(FACES)}


class ParallelContext:
    def __init__(self, domain_information: DomainInformation, rank: int = 0, world_size: int = 1, group=None):
        self.domain_information = domain_information
        self.rank = int(rank)
        self.world_size = int(world_size)
        self.group = group
        need = domain_information.no_subdomains
        if need != self.world_size:
            raise RuntimeError(
                f"case file decomposition needs {need} blocks (split_x*split_y*split_z) but the job has "
                f"{self.world_size} rank(s); launch one process per GPU with torchrun --nproc-per-node {need}")

    @classmethod
    def from_environment(cls, domain_information: DomainInformation) -> "ParallelContext":
        if dist.is_available() and dist.is_initialized():
            return cls(domain_information, dist.get_rank(), dist.get_world_size())
        if domain_information.no_subdomains > 1:
            ws = int(os.environ.get("WORLD_SIZE", "1"))
            if ws > 1:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
                if torch.cuda.is_available():
                    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
                opts = None
                if backend == "nccl":
                    try:              # NCCL kernels overlap sweeps that fill the SMs: high-priority stream
                        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                    except Exception:
                        opts = None
                dist.init_process_group(backend=backend, pg_options=opts)
                return cls(domain_information, dist.get_rank(), dist.get_world_size())
        return cls(domain_information, 0, 1)

    @property
    def is_parallel(self) -> bool:
        return self.world_size > 1

    # ------------------------------------------------------------------
    def block_boundary_types(self, bc: Dict[str, str]) -> Dict[str, str]:
        """Per-face type of THIS block: the physical type on faces at the domain boundary,
        NEIGHBOR on faces shared with another block (incl. periodic wrap across a split axis);
        cf. the device masks of halos/inner/halo_communication.py:103-150."""
        di = self.domain_information
        out = {}
        for f in FACES:
            t = bc[f]
            if t == "INACTIVE":
                out[f] = t
                continue
            nb = di.neighbor(self.rank, f, periodic=(t == "PERIODIC"))
            out[f] = "NEIGHBOR" if nb is not None else t
        return out

    def neighbors(self, bc: Dict[str, str]) -> Dict[str, int]:
        di = self.domain_information
        out = {}
        for f in FACES:
            if bc[f] == "INACTIVE":
                continue
            nb = di.neighbor(self.rank, f, periodic=(bc[f] == "PERIODIC"))
            if nb is not None:
                out[f] = nb
        return out

    # ------------------------------------------------------------------
    def exchange(self, neighbors: Dict[str, int], send: Dict[str, torch.Tensor], recv: Dict[str, torch.Tensor]):
        """Post all face messages as one batch.  send[f] = slab of my interior layers next to face f,
        destined for the opposite-face halo of neighbors[f]; recv[f] = slab for my halo at face f.
        Between one pair of ranks NCCL matches sends and receives in posting order, so sends go out
        in face order and receives are posted in the order of the SENDER's faces."""
        if not neighbors:
            return []
        ops: List[dist.P2POp] = []
        for f in FACES:                                   # sender-face order
            if f in neighbors:
                ops.append(dist.P2POp(dist.isend, send[f], neighbors[f], group=self.group, tag=FACE_ID[f]))
        for sender_face in FACES:
            f = OPPOSITE[sender_face]                     # my halo face fed by the peer's `sender_face` message
            if f in neighbors:
                ops.append(dist.P2POp(dist.irecv, recv[f], neighbors[f], group=self.group, tag=FACE_ID[sender_face]))
        return dist.batch_isend_irecv(ops)

    def allreduce_max(self, t: torch.Tensor):
        if self.is_parallel:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self):
        if self.is_parallel:
            dist.barrier(group=self.group)
