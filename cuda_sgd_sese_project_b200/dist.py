"""Multi-GPU plumbing: one process per GPU, torch.distributed for the rendezvous only.

The gradient exchange itself does not go through NCCL: every rank exports an IPC handle of its
exchange buffer, the handles are all-gathered here, every rank maps its peers' buffers and the step
kernel then writes its slice sums straight into them over NVLink (csrc/sgd_kernels.cuh, SCATTER
path).  This module is the only place where the package touches torch.
"""
from __future__ import annotations


def connect_peers(engine) -> None:
    """All-gather the peer handles over the default process group and enable the fused exchange."""
    import torch.distributed as dist

    if engine.nranks == 1:
        return
    assert dist.is_initialized() and dist.get_world_size() == engine.nranks
    handles = [None] * engine.nranks
    dist.all_gather_object(handles, engine.peer_export())
    for r, h in enumerate(handles):
        engine.peer_import(r, h)
    engine.peer_ready()
    dist.barrier()  # nobody starts stepping before every rank has mapped every inbox
