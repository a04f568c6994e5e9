"""Multi-GPU sharding of an image batch: one process per GPU, image i -> rank i mod world (the reference
parallelises over files, src/main.rs:223; images are independent units, so there is NO data-path collective).
The only collective is the broadcast of the quantiser / rate tables from rank 0 (BASELINE.json north_star)."""
import numpy as np


def shard_indices(n_images, rank, world):
    """Indices of the images owned by `rank` (static round-robin: sizes are uniform per config)."""
    return list(range(rank, n_images, world))


def broadcast_tables(tables_u16, rank, device=None):
    """Broadcast the RDO rate tables (uint16 numpy array on rank 0) to every rank with torch.distributed
    (NCCL on GPUs, gloo in the CPU tests).  Returns the numpy array every rank must use."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(tables_u16.astype(np.int32)) if rank == 0 else torch.zeros(tables_u16.shape[0], dtype=torch.int32)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy().astype(np.uint16)


def gather_stats(values, device=None):
    """all_gather of per-rank float stats (pixels, seconds, bytes) for the report."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().tolist() for o in out]
