"""Host-side multi-rank logic on CPU: shard assignment and the final gather (gloo, world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hybdrt_b200 import sharding


@pytest.mark.parametrize('n,ws', [(0, 2), (1, 2), (7, 2), (10, 3), (65536, 8), (5, 8)])
@pytest.mark.parametrize('interleave', [False, True])
def test_shards_partition_the_batch(n, ws, interleave):
    parts = [sharding.shard_indices(n, ws, r, interleave) for r in range(ws)]
    allidx = np.concatenate(parts) if parts else np.zeros(0, int)
    assert sorted(allidx.tolist()) == list(range(n))                      # every item exactly once
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1                                    # balanced
    if not interleave:
        assert all(np.all(np.diff(p) == 1) for p in parts if len(p) > 1)   # contiguous
    with pytest.raises(ValueError):
        sharding.shard_indices(n, ws, ws, interleave)


def test_single_process_gather_is_identity():
    out = sharding.gather_results({'x': np.arange(6.0).reshape(3, 2), 't': torch.arange(3)}, 3)
    assert np.array_equal(out['x'], np.arange(6.0).reshape(3, 2)) and np.array_equal(out['t'], [0, 1, 2])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n, interleave, dst, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        idx = sharding.shard_indices(n, ws, rank, interleave)
        # every "fit result" is a function of the global index, so the gathered arrays are checkable
        local = {'x': np.stack([np.full(4, float(i)) for i in idx]) if len(idx) else np.zeros((0, 4)),
                 'status': torch.as_tensor(idx.astype(np.int32) * 3),
                 # per-factor results of a PFRT map ([B, F, n_tau]) travel through the same gather
                 'pfrt_x': np.stack([np.full((3, 2), float(i)) for i in idx]) if len(idx) else np.zeros((0, 3, 2))}
        out = sharding.gather_results(local, n, interleave=interleave, dst=dst)
        if dst is None or rank == dst:
            ok = np.array_equal(out['x'], np.repeat(np.arange(n, dtype=float)[:, None], 4, 1)) and \
                np.array_equal(out['status'], np.arange(n) * 3) and \
                np.array_equal(out['pfrt_x'], np.broadcast_to(np.arange(n, dtype=float)[:, None, None], (n, 3, 2)))
        else:
            ok = out['x'] is None and out['status'] is None and out['pfrt_x'] is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('ws,n,interleave,dst', [(2, 9, False, None), (2, 9, True, None), (3, 10, True, 0), (2, 1, False, None)])
def test_gather_reassembles_the_batch_gloo(ws, n, interleave, dst):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n, interleave, dst, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(ws))
    assert all(ok for _, ok in res)
