"""One process per rank (= per GPU): wiring of the halo exchange between processes.

The reference opens its neighbour channels through MPI (SOLVER/commpi.F90:134 `ppinit`,
:408-449 ISEND/IRECV); here every rank exports an opaque blob describing its receive slabs
(`axb_ipc_export`: CUDA IPC handles for the product, a POSIX shared-memory name for the
CPU oracle), the blobs travel once over `torch.distributed` (any backend — gloo on CPU,
the gloo side-group of an NCCL job on GPUs), and each rank imports the blobs of the
ranks its halo lists name (`axb_ipc_import`).  After that the time loop never talks to
the host again: messages are peer stores + flags inside `axb_run`.
"""
from __future__ import annotations

import ctypes as C

BLOB_BYTES = 1024          # AXB_IPC_BLOB_BYTES of include/axisem_b200.h (checked against axb_ipc_blob_bytes())


def neighbours(prob):
    """Ranks this rank exchanges halo sums with (solid and fluid lists, data_comm.f90)."""
    m = prob.mesh
    peers = set()
    for hs in (m.halo_solid, m.halo_fluid):
        for k in range(hs.nmsg):
            peers.add(int(hs.list_peer[k]))
    return sorted(peers)


def connect_ranks(loop, rank: int, world: int, group=None):
    """Collective over `group`: export my blob, all-gather, import my neighbours' blobs."""
    import torch.distributed as dist
    need = int(getattr(loop.lib.lib, loop.lib.prefix + "ipc_blob_bytes")())
    if need != BLOB_BYTES:
        raise RuntimeError(f"library wants {need}-byte IPC blobs, this module sends {BLOB_BYTES}")
    blob = bytearray(BLOB_BYTES)
    buf = (C.c_char * BLOB_BYTES).from_buffer(blob)
    loop.lib.check(loop.lib.fn["ipc_export"](loop.h, buf, C.c_int32(BLOB_BYTES)))
    blobs = [None] * world
    dist.all_gather_object(blobs, bytes(blob), group=group)
    for peer in neighbours(loop.prob):
        if not 0 <= peer < world or peer == rank:
            raise RuntimeError(f"rank {rank}: halo list names rank {peer} outside the job")
        pb = (C.c_char * BLOB_BYTES).from_buffer_copy(blobs[peer])
        loop.lib.check(loop.lib.fn["ipc_import"](loop.h, C.c_int32(peer), pb, C.c_int32(BLOB_BYTES)))
    dist.barrier(group=group)
