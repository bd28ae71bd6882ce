"""The reference-side half of the drop-in boundary: integration/flatten_hw3.cpp, compiled against the UNTOUCHED objects
of the reference's HW3 tree (oracle/build_ref.sh flat3), turns the `Scene` the reference's own loader built into a
B200SceneDesc.  The other producer, b200rt_builder_* (what the Python host uses, and what every other test goes
through), restates the loader's arithmetic; here the two are held equal on the same XML — every array bit for bit, the
top-level objects up to the labelling (the reference keeps no object list after BVH::create_bvh, so the flatten numbers
instances / spheres / triangles in the order its walk meets them; `rank`, the depth-first position the tie-breaks use, is
the same number in both).

Live runs need oracle/_ref/flat3_flatten (built in the authoring container, where /root/reference exists); the
committed tests/golden/flat3_fuzz0.b200rt is that binary's output for fuzz_scene(0) and is checked everywhere."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ceng795_b200 import _lib, scene, scenes
from oracle import cpu
from tests.test_scene_cache import _arrays

HERE = os.path.dirname(os.path.abspath(__file__))
FLATTEN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "flat3_flatten")
GOLDEN = os.path.join(HERE, "golden", "flat3_fuzz0.b200rt")

# arrays that do not depend on how top-level objects are numbered
PLAIN = ("tri_geom", "tri_vidx", "tri_info", "tri_gate", "tri_uv", "tri_tex", "vert_pos", "vert_normal", "materials",
         "lights", "cameras", "textures", "brdfs", "texels", "light_cdf", "light_tri")


def _canonical(hs):
    """(objects by rank, nodes with top-level child refs rewritten to (kind, rank), root likewise)."""
    d = hs.desc
    by_rank, rank_of = {}, {}
    for kind, arr, n, ty in ((4, d.instances, d.n_instances, _lib.Instance), (2, d.spheres, d.n_spheres, _lib.Sphere),
                             (3, d.triprims, d.n_triprims, _lib.TriPrim)):
        for i in range(n):
            assert arr[i].rank not in by_rank, "two objects share a depth-first position"
            by_rank[arr[i].rank] = (kind, C.string_at(C.addressof(arr[i]), C.sizeof(ty)))
            rank_of[(kind, i)] = arr[i].rank
    nodes = np.frombuffer(C.string_at(d.nodes, d.n_nodes * 64), np.uint32).reshape(-1, 16).copy()

    def relabel(r):
        k, i = int(r) >> 28, int(r) & 0x0FFFFFFF
        return (k << 28) | rank_of[(k, i)] if (k, i) in rank_of else int(r)

    for n in range(d.n_nodes):
        nodes[n, 12], nodes[n, 13] = relabel(nodes[n, 12]), relabel(nodes[n, 13])
    return by_rank, nodes, relabel(d.root)


def _assert_same_description(ref, own):
    a, b = _arrays(ref.desc), _arrays(own.desc)
    for k in PLAIN:
        assert a[k] == b[k], k
    for f in ("semantics", "max_depth", "shadow_eps", "isect_eps", "n_tris", "n_nodes", "n_verts", "n_meshes",
              "n_instances", "n_spheres", "n_triprims", "background_texture", "env_light", "integrator"):
        assert getattr(ref.desc, f) == getattr(own.desc, f), f
    for f in ("background", "ambient", "root_box"):
        assert list(getattr(ref.desc, f)) == list(getattr(own.desc, f)), f
    for i in range(ref.desc.n_meshes):       # everything but fast_root: the flatten leaves the SAH tree to the library
        x, y = ref.desc.meshes[i], own.desc.meshes[i]
        assert (x.root, list(x.root_box), x.tri_begin, x.tri_count, x.node_begin, x.node_count) == \
               (y.root, list(y.root_box), y.tri_begin, y.tri_count, y.node_begin, y.node_count), i
        assert x.fast_root == 0xF0000000
    assert ref.desc.n_fast_tris == 0 and ref.desc.n_fast_nodes == 0
    ra, na, roota = _canonical(ref)
    rb, nb, rootb = _canonical(own)
    assert sorted(ra) == sorted(rb) == list(range(len(ra)))
    for r in ra:
        assert ra[r] == rb[r], f"top-level object at depth-first position {r}"
    assert np.array_equal(na, nb) and roota == rootb


def _flatten(xml, tmp_path):
    p = str(tmp_path / "scene.xml")
    with open(p, "w") as f:
        f.write(xml)
    out = str(tmp_path / "flat.b200rt")
    r = subprocess.run([FLATTEN, p, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return out


needs_ref = pytest.mark.skipif(not os.path.exists(FLATTEN), reason="oracle/_ref/flat3_flatten not built on this box")


@needs_ref
@pytest.mark.parametrize("case", ["fuzz0", "fuzz1", "fuzz2", "fuzz3_stochastic", "fuzz4_open", "cfg2_n48_smooth",
                                  "cfg2_glass", "cfg3_n24"])
def test_flattened_reference_scene_equals_builder(tmp_path, case):
    xml = {
        "fuzz0": lambda: scenes.fuzz_scene(0),
        "fuzz1": lambda: scenes.fuzz_scene(1),
        "fuzz2": lambda: scenes.fuzz_scene(2),
        "fuzz3_stochastic": lambda: scenes.fuzz_scene(3, spp=9, stochastic=True),   # area light, DOF, moving instance + sphere
        "fuzz4_open": lambda: scenes.fuzz_scene(4, camera_sphere_radius=40.0),
        "cfg2_n48_smooth": lambda: scenes.cfg2_hw3(n=48, width=128, height=96, smooth=True, depth=3),
        "cfg2_glass": lambda: scenes.cfg2_hw3(n=32, width=128, height=96, n_instances=3, depth=4, glass=True),
        "cfg3_n24": lambda: scenes.cfg3_hw3(n=24, width=64, height=48, spp=4),
    }[case]()
    ref = scene.CachedHostScene(_flatten(xml, tmp_path))
    own = scene.HostScene(scene.parse_xml(xml, "hw3"))
    _assert_same_description(ref, own)


@needs_ref
def test_golden_is_what_the_flatten_writes_today(tmp_path):
    out = _flatten(scenes.fuzz_scene(0), tmp_path)
    assert open(out, "rb").read() == open(GOLDEN, "rb").read()


def test_golden_flattened_scene_equals_builder_and_renders_the_same():
    """Runs everywhere (no reference binary needed): the committed description the reference's own objects produced
    against the builder's, then the oracle on both — same ids, same t, same image, bit for bit."""
    xml = scenes.fuzz_scene(0)
    ref = scene.CachedHostScene(GOLDEN)
    own = scene.HostScene(scene.parse_xml(xml, "hw3"))
    _assert_same_description(ref, own)
    img_r, ids_r, t_r, st_r = cpu.render(ref, 0, want_ids=True)
    img_o, ids_o, t_o, st_o = cpu.render(own, 0, want_ids=True)
    assert np.array_equal(img_r.view(np.uint32), img_o.view(np.uint32))
    assert np.array_equal(t_r.view(np.uint32), t_o.view(np.uint32))
    assert np.array_equal(ids_r < 0, ids_o < 0)      # misses agree; hit labels differ only by the object numbering
