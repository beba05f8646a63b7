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


def _fake_sample(index, im):
    """Stands in for the CUDA sampler: a row that depends only on the image and its GLOBAL index (what the real path
    guarantees through chain_id0 = index * n_chains)."""
    import psgla_b200 as P
    row = torch.zeros(len(P.image_set.ROW_FIELDS), dtype=torch.float64)
    row[0] = index
    row[1] = float(im.double().mean()) + index
    row[2] = float(im.double().std())
    row[-2], row[-1] = im.shape[-2], im.shape[-1]
    return row, None


def _set_worker(rank, ws, port, n_images, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    import psgla_b200 as P
    P.dist.init_from_env(backend="gloo")
    g = torch.Generator().manual_seed(0)
    images = [torch.rand(3, 8 + i, 10, generator=g) for i in range(n_images)]  # ragged sizes, same list on every rank
    assert P.image_set.deal_round_robin(n_images, rank, ws) == list(range(rank, n_images, ws))
    res = P.run_image_set(images, _sample=_fake_sample)
    if rank == 0:
        np.save(out_path, np.array([[d["index"], d["psnr_mmse"], d["ssim_mmse"], d["H"], d["W"]] for d in res]))
    else:
        assert res is None
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_image_set_round_robin_is_world_size_invariant(tmp_path):
    """run_image_set's host logic (deal -> per-image rows -> gather -> order) at world size 2 and 3 over gloo against the single
    process: 5 images (ragged deal 3 + 2, and 2 + 2 + 1), more ranks than images is covered by n_images = 1."""
    import psgla_b200 as P
    for n_images in (5, 1):
        g = torch.Generator().manual_seed(0)
        images = [torch.rand(3, 8 + i, 10, generator=g) for i in range(n_images)]
        solo = P.run_image_set(images, _sample=_fake_sample)
        want = np.array([[d["index"], d["psnr_mmse"], d["ssim_mmse"], d["H"], d["W"]] for d in solo])
        assert [d["index"] for d in solo] == list(range(n_images))
        for ws in (2, 3):
            out = str(tmp_path / ("s%d_%d.npy" % (n_images, ws)))
            mp.spawn(_set_worker, args=(ws, _free_port(), n_images, out), nprocs=ws, join=True)
            assert np.array_equal(np.load(out), want)
