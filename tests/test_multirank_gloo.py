"""N>1 host logic on CPU: world_size-2 gloo processes shard the samples and reduce exact int64
accumulators to rank 0 (the same code path bench.py uses with NCCL)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_samples_partitions_exactly():
    dist = importlib.import_module("raytracing-practice_b200.dist")
    for spp in (1, 7, 10, 64, 1250, 10000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                b, n = dist.shard_samples(spp, r, world)
                assert n >= 0
                seen += list(range(b, b + n))
            assert seen == list(range(spp))
            sizes = [dist.shard_samples(spp, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_samples(10, 2, 2)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch

    dist = importlib.import_module("raytracing-practice_b200.dist")
    r, _, w = dist.init_process_group("gloo")
    assert (r, w) == (rank, world)
    # a deterministic stand-in for the per-rank accumulator: sample s contributes f(pixel, s);
    # each rank sums its own shard exactly like the kernel's int64 fixed-point adds
    spp, npix = 37, 1000
    begin, count = dist.shard_samples(spp, rank, world)
    pix = np.arange(npix * 3, dtype=np.int64)
    acc = np.zeros(npix * 3, np.int64)
    for s in range(begin, begin + count):
        acc += (pix * 2654435761 + s * 40503) % (1 << 40)
    t = torch.from_numpy(acc)
    dist.reduce_accum_to_rank0(t)
    if rank == 0:
        np.save(out_path, t.numpy())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_reduce_is_exact(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "acc.npy")
    mp.start_processes(_worker, args=(2, port, out), nprocs=2, join=True, start_method="spawn")
    got = np.load(out)
    pix = np.arange(3000, dtype=np.int64)
    want = np.zeros(3000, np.int64)
    for s in range(37):
        want += (pix * 2654435761 + s * 40503) % (1 << 40)
    assert np.array_equal(got, want)
