"""Host-side logic of the multi-GPU (one process per GPU) mode.

The reference's only parallel pattern is replicated data + target partition +
a loop of MPI_BCAST (src/MPISetup.f90:132-146, src/SphereBVESolver.f90:422-429).
Here torch.distributed is the plumbing that replaces MPI for bootstrapping
(rank/world, broadcasting the NCCL unique id, barriers, max-over-ranks timing);
the data exchange itself is lpm_comm_allgather_slices_dev inside liblpmgpu.so.
`exchange_slices` is the same exchange written with torch.distributed
collectives: it runs on gloo, so the slice arithmetic is testable without GPUs.
"""
import torch
import torch.distributed as dist


def slice_of(n, world, rank):
    """0-based half-open LoadBalance slice of `rank` (MPISetup.f90:138-144)."""
    chunk = n // world
    beg = rank * chunk
    end = n if rank == world - 1 else (rank + 1) * chunk
    return beg, end


def broadcast_unique_id(uid, src=0):
    """Rank `src` passes the 128-byte NCCL id from lpm_comm_unique_id; everyone gets it."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(uid), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def exchange_slices(tensors, group=None):
    """for r in ranks: broadcast(slice_r, root=r) on each array, in place."""
    world = dist.get_world_size(group)
    for t in tensors:
        n = t.numel()
        for r in range(world):
            b, e = slice_of(n, world, r)
            if e > b:
                dist.broadcast(t[b:e], src=r, group=group)


def max_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
