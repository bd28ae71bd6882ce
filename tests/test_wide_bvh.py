"""The default traversal's per-lane source (csrc/rt_wide.cuh) and the 8-wide tree builder (csrc/wide_bvh.cpp), compiled
for the host by g++ (tests/wide_host.cpp), against the oracle: hit ids and t have to be the same BITS on camera rays and
on random rays (axis-parallel and zero directions included) of every BVH tree of the reference.  No GPU needed; the
device runs the same functions under its warp scheduling (tests/test_gpu_parity.py holds that side)."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from ceng795_b200 import scene, scenes
from oracle import cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _twin():
    src = os.path.join(ROOT, "tests", "wide_host.cpp")
    inc = "/usr/local/cuda/include"
    if not (shutil.which("g++") and os.path.exists(os.path.join(inc, "cuda_runtime.h"))):
        pytest.skip("g++ / CUDA headers not available")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    so = os.path.join(ROOT, "build", "libwide_host.so")
    deps = [src] + [os.path.join(ROOT, "ceng795_b200", "csrc", f)
                    for f in ("rt_wide.cuh", "rt_traverse.cuh", "rt_common.cuh", "wide_bvh.cpp", "wide_bvh.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-I", inc,
                        "-I", os.path.join(ROOT, "include"), src, "-o", so], check=True)
    H = C.CDLL(so)
    H.widehost_trace.restype = C.c_int
    H.widehost_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return H


def _trace(H, hs, rays, any_hit=False):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    n = rays.shape[0]
    ids = np.zeros((n, 2), np.int32)
    t = np.zeros(n, np.float32)
    stats = np.zeros(6, np.int64)
    rc = H.widehost_trace(C.cast(hs.desc_ptr, C.c_void_p), rays.ctypes.data, n, 1 if any_hit else 0, ids.ctypes.data,
                          t.ctypes.data, stats.ctypes.data)
    assert rc == 0, rc
    return ids, t, stats


def _camera_rays(hs, n=4000, seed=0):
    """Rays from the camera position through random points of its image plane (the oracle's own camera is not needed:
    any ray will do, these just look at the scene)."""
    c = hs.desc.cameras[0]
    rng = np.random.default_rng(seed)
    e = np.array(c.e[:], np.float32)
    tl, su, sv = (np.array(x[:], np.float32) for x in (c.top_left, c.su, c.sv))
    x = rng.uniform(0, c.width, n).astype(np.float32)[:, None]
    y = rng.uniform(0, c.height, n).astype(np.float32)[:, None]
    s = tl + su * x - sv * y
    d = s - e
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = e; rays[:, 4:7] = d
    rays[:, 3] = rng.uniform(0, 1, n)
    return rays


def _random_rays(n, seed, span=5.0):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-span, span, (n, 3))
    d = rng.normal(size=(n, 3))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3] = rng.uniform(0, 1, n)
    # axis-parallel directions (the reference's box test skips |d| < 1e-6 axes), tiny components, zero direction
    rays[:32, 4:7] = 0; rays[:32, 5] = -1
    rays[32:64, 4:7] = [1, 0, 0]
    rays[64:96, 4:7] = [0, 0, -1]
    rays[96:100, 4:7] = 0
    rays[100:132, 4] = 1e-7
    rays[132:164, 6] = -1e-20
    return rays


def _check(H, hs, rays):
    ids, t, stats = _trace(H, hs, rays)
    o_ids, o_t, _ = cpu.trace_rays(hs, rays)
    same = np.all(ids == o_ids, -1)
    assert same.all(), (float(same.mean()), np.nonzero(~same)[0][:8], ids[~same][:4], o_ids[~same][:4])
    assert np.array_equal(t.view(np.uint32), o_t.view(np.uint32))
    return stats, o_t


@pytest.mark.parametrize("tree", ["hw3", "hw4", "hw7"])
def test_wide_tree_gives_the_oracles_hits_on_cfg2(tree):
    H = _twin()
    hs = scene.HostScene(scene.parse_xml(scenes.cfg2_hw3(n=64, width=32, height=32, seed=5), tree))
    stats, o_t = _check(H, hs, _camera_rays(hs, 6000))
    assert np.isfinite(o_t).mean() > 0.5
    assert stats[2] <= 16 and stats[4] <= 16, stats        # RT_WSTACK: the shared-memory stack holds the whole walk
    _check(H, hs, _random_rays(20_000, 7))


@pytest.mark.parametrize("tree", ["hw2", "hw3", "hw4", "hw5", "hw6", "hw7"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_wide_tree_on_fuzz_scenes(tree, seed):
    """scenes.fuzz_scene: mirroring scales, transform chains, degenerate triangles, camera inside a sphere, moving
    instances and spheres (per-ray time), both the closed and the open variant."""
    H = _twin()
    for radius in (2.5, 40.0):
        hs = scene.HostScene(scene.parse_xml(scenes.fuzz_scene(seed, camera_sphere_radius=radius), tree))
        _check(H, hs, _camera_rays(hs, 3000, seed))
        _check(H, hs, _random_rays(6000, seed + 11, span=2.5))


def test_wide_tree_on_ties_and_ppm_and_hw7_scene():
    H = _twin()
    for tree in ("hw3", "hw7"):
        for spacing in (0.0, 6e-7):
            hs = scene.HostScene(scene.parse_xml(scenes.hw1_near_ties(spacing=spacing, width=64, height=48), tree))
            _check(H, hs, _camera_rays(hs, 4000))
            _check(H, hs, _random_rays(8000, 3, span=2.5))
    hs = scene.HostScene(scene.parse_xml(scenes.cfg5_ppm(width=64, height=48, mesh_slices=(24, 20)), "ppm"))
    _check(H, hs, _camera_rays(hs, 4000))
    _check(H, hs, _random_rays(8000, 5, span=3.0))
    d = tempfile.mkdtemp()
    hs = scene.HostScene(scene.parse_xml(scenes.hw7_scene(d, width=64, height=48), "hw7", base_dir=d))
    _check(H, hs, _camera_rays(hs, 4000))
    _check(H, hs, _random_rays(8000, 6, span=4.0))


def test_any_hit_agrees_with_closest_hit():
    """The shadow predicate `something in (0, limit)` through the any-hit walk == closest-hit t < limit."""
    H = _twin()
    hs = scene.HostScene(scene.parse_xml(scenes.cfg2_hw3(n=48, width=32, height=32, seed=2), "hw3"))
    rays = _camera_rays(hs, 4000, 3)
    _, o_t, _ = cpu.trace_rays(hs, rays)
    rng = np.random.default_rng(1)
    lim = np.where(np.isfinite(o_t), o_t * rng.choice([0.5, 0.999, 1.001, 2.0], len(o_t)), 10.0).astype(np.float32)
    rays[:, 7] = lim
    occ, _, _ = _trace(H, hs, rays, any_hit=True)
    assert np.array_equal(occ[:, 0] == 1, o_t < lim)
