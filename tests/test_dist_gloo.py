"""Host-side sharding logic on CPU: world_size-2 gloo processes exchange synthetic row bands,
sample-batch partial frames and PPM {S, n} arrays exactly as bench.py does over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from ceng795_b200 import dist as rdist


def test_splits_partition_exactly():
    for total in (1, 7, 1080, 4096, 10_000_001):
        for world in (1, 2, 3, 8):
            bands = [rdist.split_rows(total, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == total
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in bands]
            assert max(sizes) - min(sizes) <= 1
            ph = [rdist.split_photons(total, r, world) for r in range(world)]
            assert sum(c for _, c in ph) == total and ph[0][0] == 0
            st = [rdist.split_samples(total, r, world) for r in range(world)]
            assert st[-1]["sample_end"] == total
    rows = sorted(j for r in range(4) for j in range(r, 37, 4))      # the reference's interleave
    assert rows == list(range(37))
    assert rdist.split_interleaved(37, 1, 4) == dict(row_begin=1, row_end=37, row_stride=4)
    assert rdist.cameras_of(5, 1, 2) == [1, 3]


def _worker(rank, world, port, H, W):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        frame = torch.rand((H, W, 4), generator=g)                    # the frame every rank would agree on
        # (1) row bands gathered with one all-gather
        b, e = rdist.split_rows(H, rank, world)
        out = rdist.gather_rows(frame[b:e].clone(), H, rank, world)
        assert torch.equal(out, frame)
        # (2) interleaved rows: disjoint partial frames, one sum-reduce
        part = torch.zeros_like(frame)
        part[rank::world] = frame[rank::world]
        assert torch.equal(rdist.reduce_interleaved(part), frame)
        # (3) sample batches: Pixel{sum c*w, sum w} adds up
        ns = 16
        samples = torch.rand((ns, H, W, 4), generator=g)
        o = rdist.split_samples(ns, rank, world)
        mine = samples[o["sample_begin"]:o["sample_end"]].sum(0)
        tot = rdist.allreduce_sum(mine)
        assert torch.allclose(tot, samples.sum(0), atol=1e-5)
        # (4) PPM: every rank holds the same hit points; {S.rgb, n} of its photon share are summed
        n_hp = 1000
        first, cnt = rdist.split_photons(12345, rank, world)
        S_n = torch.zeros((n_hp, 4))
        idx = torch.arange(first, first + cnt) % n_hp
        S_n.index_add_(0, idx, torch.ones((cnt, 4)))
        tot = rdist.allreduce_sum(S_n)
        want = torch.zeros((n_hp, 4)).index_add_(0, torch.arange(12345) % n_hp, torch.ones((12345, 4)))
        assert torch.equal(tot, want)
    finally:
        tdist.destroy_process_group()


def test_world_size_2_gloo_collectives():
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 37, 16), nprocs=2, join=True)
