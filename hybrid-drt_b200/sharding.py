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


_PINNED = {}          # (name, shape, dtype) -> page-locked staging tensor of the device -> host copies (a handful of them)


def _staging(name, shape, dtype):
    key = (name, tuple(shape), dtype)        # per result name: two results of one gather never share a buffer
    buf = _PINNED.get(key)
    if buf is None:
        if len(_PINNED) >= 32:
            _PINNED.pop(next(iter(_PINNED)))
        buf = _PINNED[key] = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
    return buf


def gather_results(local, n_items, interleave=False, dst=None, copy=True):
    """Reassemble per-rank result arrays into full-batch arrays.

    ``local``: dict name -> array / tensor whose first axis is this rank's shard (in shard_indices order).
    Returns a dict of numpy arrays of leading size ``n_items`` on every rank (``dst=None``, all-gather) or on
    rank ``dst`` only (others get None).  Shards may have different sizes: they are padded to the largest one
    for the collective and trimmed afterwards.

    One collective per dtype: the arrays of a dtype travel as the columns of one packed matrix; the rows are put back
    into item order on the device (one index_select) and every result leaves it as one contiguous device -> host copy
    through a cached page-locked buffer.  ``copy=False`` returns views of those buffers: valid until the next gather of
    the same shapes -- for callers that consume the arrays at once (the map scatter); the default copies them out.
    """
    rank, ws = world()
    if ws == 1:
        return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in local.items()}
    nccl = dist.get_backend() == 'nccl'
    dev = torch.device('cuda', torch.cuda.current_device()) if nccl else torch.device('cpu')
    idx = [shard_indices(n_items, ws, r, interleave) for r in range(ws)]
    counts = [len(i) for i in idx]
    cmax = max(counts)
    # row of item i in the stacked (ws * cmax) layout of the collective
    in_order = (not interleave) and all(c == cmax for c in counts)
    pos = None
    if not in_order:
        pos_np = np.empty(n_items, dtype=np.int64)
        for r in range(ws):
            pos_np[idx[r]] = r * cmax + np.arange(counts[r])
        pos = torch.from_numpy(pos_np).to(dev)
    tensors, groups = {}, {}
    for name in sorted(local):
        v = local[name]
        t = v.detach() if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
        t = t.to(dev)
        if t.shape[0] != counts[rank]:
            raise ValueError(f'{name}: {t.shape[0]} rows for a shard of {counts[rank]}')
        tensors[name] = t
        groups.setdefault(t.dtype, []).append(name)
    receives = dst is None or rank == dst
    out = {name: None for name in tensors}
    for dtype, names in groups.items():
        cols = [tensors[nm].reshape(counts[rank], int(np.prod(tensors[nm].shape[1:], dtype=np.int64))) for nm in names]
        widths = [c.shape[1] for c in cols]
        k = sum(widths)
        if len(cols) == 1 and counts[rank] == cmax:
            pad = cols[0].contiguous()
        else:
            pad = torch.zeros((cmax, k), dtype=dtype, device=dev)
            off = 0
            for c, w in zip(cols, widths):
                pad[:counts[rank], off:off + w] = c
                off += w
        big = torch.empty((ws * cmax, k), dtype=dtype, device=dev) if receives else None
        if dst is None:
            if nccl:
                dist.all_gather_into_tensor(big, pad)
            else:
                dist.all_gather(list(big.view(ws, cmax, k).unbind(0)), pad)
        else:
            dist.gather(pad, list(big.view(ws, cmax, k).unbind(0)) if receives else None, dst=dst)
        if not receives:
            continue
        ordered = big[:n_items] if in_order else big.index_select(0, pos)
        off = 0
        for nm, w in zip(names, widths):
            part = ordered if len(names) == 1 else ordered[:, off:off + w].contiguous()
            if part.is_cuda:
                host = _staging(nm, part.shape, part.dtype)
                host.copy_(part)
                arr = host.numpy().copy() if copy else host.numpy()
            else:
                arr = part.numpy()
            out[nm] = arr.reshape((n_items,) + tuple(tensors[nm].shape[1:]))
            off += w
    return out
