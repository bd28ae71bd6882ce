"""Host-side sharding logic on CPU: the row / strata / photon splits tile their ranges exactly, and
world_size-2 gloo processes exchange what bench.py exchanges over NCCL — the frame's handle (plain bytes), sample-batch
partial frames and the PPM count deltas — with synthetic data."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from ceng795_b200 import _lib
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


def test_row_bands_deal_every_row_once_and_evenly():
    """b200rt_multi_share's split: 8-row bands round-robin (the reference's `j = i; j += thread_count` in 8-row units)."""
    for height in (1, 7, 8, 9, 72, 800, 1080):
        for parts in (1, 2, 3, 4, 8, 16):
            got = [rdist.row_bands(height, p, parts) for p in range(parts)]
            assert sorted(r for rows in got for r in rows) == list(range(height))
            sizes = [len(r) for r in got]
            assert max(sizes) - min(sizes) <= 8
            for rows in got:   # whole bands: k_generate's 8 x 4 pixel tiles stay inside one share
                assert all(r % 8 == 0 or r - 1 in rows for r in rows)
    assert len(rdist.row_bands(1080, 0, 8)) == 136 and len(rdist.row_bands(1080, 7, 8)) == 128


def test_shared_frame_handle_is_plain_bytes():
    f = _lib.SharedFrame()
    f.device_ptr = 0x7f0012345600; f.pid = 4242; f.tail_offset = 1920 * 1080 * 16; f.device = 3
    f.width, f.height, f.direct = 1920, 1080, 1
    for i in range(64):
        f.ipc_handle[i] = i
    raw = bytes(f)
    assert len(raw) == 64 + 8 + 8 + 8 + 16
    g = _lib.SharedFrame.from_buffer_copy(raw)
    assert (g.device_ptr, g.pid, g.tail_offset, g.device, g.width, g.height, g.direct) == (
        0x7f0012345600, 4242, 1920 * 1080 * 16, 3, 1920, 1080, 1) and bytes(g.ipc_handle) == bytes(range(64))


def _worker(rank, world, port, H, W):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        frame = torch.rand((H, W, 4), generator=g)                    # the frame every rank would agree on
        # (1) the frame's handle travels as an object; every rank ends up with rank 0's bytes
        box = [bytes(range(160)) if rank == 0 else None]
        tdist.broadcast_object_list(box, src=0)
        assert box[0] == bytes(range(160))
        # (2) row bands: disjoint shares of one frame, each written once (here: summed, they are disjoint)
        part = torch.zeros_like(frame)
        mine = rdist.row_bands(H, rank, world)
        part[mine] = frame[mine]
        assert torch.equal(rdist.allreduce_sum(part), frame)
        # (3) sample batches: Pixel{sum c*w, sum w} adds up
        ns = 16
        samples = torch.rand((ns, H, W, 4), generator=g)
        o = rdist.split_samples(ns, rank, world)
        mine = samples[o["sample_begin"]:o["sample_end"]].sum(0)
        tot = rdist.allreduce_sum(mine)
        assert torch.allclose(tot, samples.sum(0), atol=1e-5)
        # (4) PPM counts: every rank holds the same hit points and the same n_base; after a pass each has accepted
        # some photons locally; n = n_base + sum of the deltas on every rank
        n_hp = 1000
        n_base = torch.arange(n_hp, dtype=torch.int32) % 7
        first, cnt = rdist.split_photons(12345, rank, world)
        delta = torch.zeros(n_hp, dtype=torch.int32)
        delta.index_add_(0, torch.arange(first, first + cnt) % n_hp, torch.ones(cnt, dtype=torch.int32))
        n = rdist.sum_counts(delta, n_base)
        want = n_base + torch.zeros(n_hp, dtype=torch.int32).index_add_(0, torch.arange(12345) % n_hp,
                                                                        torch.ones(12345, dtype=torch.int32))
        assert torch.equal(n, want)
        # (5) PPM flux: S of the photon shares is additive
        S = torch.zeros((n_hp, 4))
        S.index_add_(0, torch.arange(first, first + cnt) % n_hp, torch.ones((cnt, 4)))
        tot = rdist.allreduce_sum(S)
        assert torch.equal(tot, torch.zeros((n_hp, 4)).index_add_(0, torch.arange(12345) % n_hp, torch.ones((12345, 4))))
    finally:
        tdist.destroy_process_group()


def test_world_size_2_gloo_collectives():
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 37, 16), nprocs=2, join=True)
