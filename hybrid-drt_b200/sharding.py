"""Sharding of a batch of independent spectra over the ranks of one node, and the final gather.

The fit has no coupling between spectra (reference: one ``fit_observation`` per observation,
hybdrt/mapping/drtmd.py:303-319), so the only multi-GPU machinery is: who fits which spectra, and how the
results come back.  One process per GPU (``torchrun``); NCCL when the tensors live on GPUs, gloo in the CPU
tests.  No collective runs during the fit.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items, world_size, rank, interleave=False):
    """Indices of the items rank ``rank`` fits.

    contiguous (default): ranks get consecutive blocks whose sizes differ by at most one.
    interleave: item i goes to rank i % world_size -- for maps whose difficulty varies smoothly with position
    (SURVEY.md section 8e), so that every rank sees the same mix of easy and hard spectra.
    """
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    if interleave:
        return np.arange(rank, n_items, world_size)
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return np.arange(start, start + base + (1 if rank < extra else 0))


def gather_results(local, n_items, interleave=False, dst=None):
    """Reassemble per-rank result arrays into full-batch arrays.

    ``local``: dict name -> array / tensor whose first axis is this rank's shard (in shard_indices order).
    Returns a dict of numpy arrays of leading size ``n_items`` on every rank (``dst=None``, all-gather) or on
    rank ``dst`` only (others get None).  Shards may have different sizes: they are padded to the largest one
    for the collective and trimmed afterwards.
    """
    rank, ws = world()
    if ws == 1:
        return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in local.items()}
    backend = dist.get_backend()
    counts = [len(shard_indices(n_items, ws, r, interleave)) for r in range(ws)]
    cmax = max(counts)
    out = {}
    for name in sorted(local):
        v = local[name]
        t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(v))
        if backend == 'nccl' and not t.is_cuda:
            t = t.cuda()
        if backend != 'nccl' and t.is_cuda:
            t = t.cpu()
        pad = torch.zeros((cmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        if dst is None:
            bufs = [torch.empty_like(pad) for _ in range(ws)]
            dist.all_gather(bufs, pad)
        else:
            bufs = [torch.empty_like(pad) for _ in range(ws)] if rank == dst else None
            dist.gather(pad, bufs, dst=dst)
            if rank != dst:
                out[name] = None
                continue
        full = np.empty((n_items,) + tuple(t.shape[1:]), dtype=pad.cpu().numpy().dtype)
        for r in range(ws):
            full[shard_indices(n_items, ws, r, interleave)] = bufs[r][:counts[r]].cpu().numpy()
        out[name] = full
    return out
