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


def broadcast_and_install_rate_tables(qidxs, rank, device):
    """The one collective of a multi-GPU batch (BASELINE.json north_star): rank 0 builds the quantiser / rate tables of every
    base_q_idx the batch uses, ONE broadcast (NCCL over NVLink when `device` is a CUDA device) carries them to every rank,
    and every rank installs the copy it received - straight from device memory - into the encoder
    (b200avif_install_rate_tables), which then never builds tables of its own for those quantisers.
    Returns the bytes broadcast."""
    import ctypes
    import torch
    import torch.distributed as dist
    from . import encoder
    L = encoder.lib()
    nb = L.b200avif_rate_tables_bytes()
    host = torch.zeros(len(qidxs) * nb, dtype=torch.uint8)
    if rank == 0:
        for k, q in enumerate(qidxs):
            buf = (ctypes.c_uint8 * nb)()
            assert L.b200avif_build_rate_tables(int(q), buf) == 0
            host[k * nb:(k + 1) * nb] = torch.frombuffer(buf, dtype=torch.uint8)
    t = host.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=0)
    if t.is_cuda:
        torch.cuda.synchronize(t.device)
        for k, q in enumerate(qidxs):
            rc = L.b200avif_install_rate_tables(t.device.index, int(q), ctypes.c_void_p(t.data_ptr() + k * nb), nb)
            assert rc == 0, L.b200avif_last_error_message()
    return t.numel()
