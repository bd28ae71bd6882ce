"""The in-tile counting sort of the HW7 shade kernels (csrc/rt_shade7.cuh, tile_order7), restated step by step.

The device code groups the vertices of a tile by class with per-warp-round `__match_any_sync` ranks, a per-class
prefix over the warp-rounds and an exclusive scan over the classes.  This walks the same five steps with the same
tables (cnt[class][segment], first[class], perm[position]) for full, ragged and tiny tiles and checks what the kernels
rely on: perm[0 .. n_valid) is a permutation of the tile's vertices, nothing is written beyond n_valid, and the order is
the STABLE sort by class (so the walk is the same from run to run).  The device code itself is covered by the parity
tests of every HW5 / HW6 / HW7 scene (`-m gpu`): a vertex is shaded from its own slot whatever thread takes it.
"""
import re
from pathlib import Path

import numpy as np
import pytest

SRC = (Path(__file__).resolve().parents[1] / "ceng795_b200" / "csrc" / "rt_shade7.cuh").read_text()


def _macro(name):
    m = re.search(rf"#define {name} (\d+)", SRC)
    assert m, name
    return int(m.group(1))


BLOCK, ROUNDS, CLASSES = _macro("RT_SHADE7_BLOCK"), _macro("RT_SHADE7_ROUNDS"), _macro("RT_SHADE7_CLASSES")
WARPS = BLOCK // 32
SEGS = WARPS * ROUNDS
TILE = BLOCK * ROUNDS


def tile_order(cls, n_valid):
    cnt = np.zeros((CLASSES, SEGS), np.int64)                       # step 1: cleared
    c = np.full((ROUNDS, BLOCK), CLASSES, np.int64)
    rank = np.zeros((ROUNDS, BLOCK), np.int64)
    for j in range(ROUNDS):                                         # step 2: rank inside the warp-round, leader counts
        for w in range(WARPS):
            tid = w * 32 + np.arange(32)
            i = j * BLOCK + tid
            cj = np.where(i < n_valid, cls[np.minimum(i, len(cls) - 1)] & (CLASSES - 1), CLASSES)
            for lane in range(32):
                same = np.flatnonzero(cj == cj[lane])               # __match_any_sync
                r = int((same < lane).sum())
                c[j, tid[lane]], rank[j, tid[lane]] = cj[lane], r
                if r == 0 and cj[lane] < CLASSES:
                    cnt[cj[lane], j * WARPS + w] = len(same)
    first = np.zeros(CLASSES, np.int64)
    for t in range(CLASSES):                                        # step 3: one thread per class, over the segments
        run = 0
        for s in range(SEGS):
            v = cnt[t, s]
            cnt[t, s] = run
            run += v
        first[t] = run
    per = CLASSES // 32                                             # step 4: warp 0, `per` consecutive classes per lane
    mine = first.reshape(32, per).sum(1)
    incl = np.cumsum(mine)                                          # the shuffle-up scan
    at = incl - mine
    sizes = first.copy()
    for lane in range(32):
        a = at[lane]
        for q in range(per):
            first[per * lane + q] = a
            a += sizes[per * lane + q]
    perm = np.full(TILE, -1, np.int64)
    for j in range(ROUNDS):                                         # step 5: scatter
        for tid in range(BLOCK):
            if c[j, tid] < CLASSES:
                pos = first[c[j, tid]] + cnt[c[j, tid], j * WARPS + tid // 32] + rank[j, tid]
                assert perm[pos] == -1, "two vertices at one position"
                perm[pos] = j * BLOCK + tid
    return perm


def test_macros_are_what_the_tables_assume():
    assert CLASSES % 32 == 0 and TILE <= 65535 and BLOCK % 32 == 0   # unsigned short tables, whole warps


@pytest.mark.parametrize("n_valid", [TILE, TILE - 1, BLOCK + 1, BLOCK, 33, 1])
def test_tile_walk_is_the_stable_sort_by_class(n_valid):
    rng = np.random.default_rng(n_valid)
    for spread in (3, 20, 200):                                      # few classes (a Cornell box), many, with collisions
        cls = rng.integers(0, spread, TILE)
        perm = tile_order(cls, n_valid)
        assert (perm[n_valid:] == -1).all()
        want = np.argsort(cls[:n_valid] & (CLASSES - 1), kind="stable")
        assert np.array_equal(perm[:n_valid], want)
