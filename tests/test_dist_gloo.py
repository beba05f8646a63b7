"""not-gpu: the N>1 host path (shard -> run -> gather on rank 0 -> metric) with world_size 2 over gloo on the CPU.
The chain kernel itself cannot run here, so each rank fills its shard with a deterministic function of the GLOBAL
chain id -- which is exactly the invariant the real path relies on (Philox subsequence = global chain id)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    import psgla_b200 as P
    r, w, _ = P.dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, ws)
    a, b = P.dist.shard_range(n_total, rank, ws)
    ids = torch.arange(a, b, dtype=torch.float64)
    local = torch.stack([torch.sin(ids), torch.cos(ids)], 1)  # "final samples" of chains a..b-1
    full = P.dist.gather_to_rank0(local, n_total)
    pooled = P.dist.reduce_mean_to_rank0(local.sum(0), b - a)
    t = P.dist.max_over_ranks(10.0 + rank)
    assert t == 10.0 + ws - 1
    if rank == 0:
        np.save(out_path, np.concatenate([full.numpy().reshape(-1), pooled.numpy()]))
    else:
        assert full is None and pooled is None
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_gather_world2(tmp_path):
    n_total = 1001  # ragged: 501 + 500
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    got = np.load(out)
    ids = np.arange(n_total, dtype=np.float64)
    want = np.stack([np.sin(ids), np.cos(ids)], 1)
    assert np.allclose(got[:-2].reshape(n_total, 2), want)
    assert np.allclose(got[-2:], want.mean(0))
