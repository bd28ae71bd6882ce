"""Binary scene cache: a flattened description written to disk and read back is the same
description — every array byte for byte — and a corrupt or foreign file is refused loudly."""
import ctypes as C
import os

import numpy as np
import pytest

from ceng795_b200 import _lib, image, scene, scenes
from oracle import cpu
from tests.golden.make_golden import ASSETS, scene_xml


def _arrays(d):
    """(name, bytes) of every array a B200SceneDesc points at."""
    out = {}
    for name, count, size in (("nodes", d.n_nodes, 64), ("tri_geom", d.n_tris * 12, 4), ("tri_vidx", d.n_tris * 3, 4),
                              ("tri_info", d.n_tris * 2, 4), ("fast_nodes", d.n_fast_nodes, 64),
                              ("fast_tri_geom", d.n_fast_tris * 12, 4), ("fast_tri_ref", d.n_fast_tris, 4),
                              ("tri_gate", d.n_tris * 6, 4), ("vert_pos", d.n_verts * 3, 4),
                              ("vert_normal", d.n_verts * 3, 4), ("tri_uv", d.n_tris * 6, 4), ("tri_tex", d.n_tris, 4),
                              ("texels", d.n_texels, 4), ("light_cdf", d.n_light_cdf, 4), ("light_tri", d.n_light_cdf, 4)):
        p = getattr(d, name)
        out[name] = C.string_at(p, int(count) * size) if p and count else b""
    for name, count in (("meshes", d.n_meshes), ("instances", d.n_instances), ("spheres", d.n_spheres),
                        ("triprims", d.n_triprims), ("materials", d.n_materials), ("lights", d.n_lights),
                        ("cameras", d.n_cameras), ("textures", d.n_textures), ("brdfs", d.n_brdfs)):
        p = getattr(d, name)
        out[name] = C.string_at(p, int(count) * C.sizeof(p._type_)) if p and count else b""
    return out


@pytest.mark.parametrize("name", ["cfg2_hw3_n48_smooth", "hw7_whitted_depth1", "cfg1_hw1_depth0"])
def test_round_trip_is_byte_identical(tmp_path, name):
    tree, xml = scene_xml(name)
    hs = scene.HostScene(scene.parse_xml(xml, tree, base_dir=ASSETS))
    path = str(tmp_path / "scene.b200rt")
    hs.save(path)
    back = scene.CachedHostScene(path)
    a, b = _arrays(hs.desc), _arrays(back.desc)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k] == b[k], k
    for f in ("semantics", "max_depth", "shadow_eps", "isect_eps", "root", "n_tris", "n_fast_tris", "n_lights",
              "background_texture", "env_light", "integrator", "importance_sampling"):
        assert getattr(hs.desc, f) == getattr(back.desc, f), f
    assert list(hs.desc.root_box) == list(back.desc.root_box)
    # written again from the loaded copy: the same file (pointer values never reach the disk)
    path2 = str(tmp_path / "again.b200rt")
    back.save(path2)
    assert open(path, "rb").read() == open(path2, "rb").read()
    # and the oracle renders the same image from it
    r0, _, _, _ = cpu.render(hs, 0)
    r1, _, _, _ = cpu.render(back, 0)
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32))


def test_corrupt_and_foreign_files_are_refused(tmp_path):
    tree, xml = scene_xml("cfg2_hw3_n32")
    hs = scene.HostScene(scene.parse_xml(xml, tree))
    path = str(tmp_path / "scene.b200rt")
    hs.save(path)
    blob = bytearray(open(path, "rb").read())
    blob[len(blob) // 2] ^= 0x40
    bad = str(tmp_path / "bad.b200rt")
    open(bad, "wb").write(blob)
    with pytest.raises(_lib.B200RTError, match="checksum"):
        scene.CachedHostScene(bad)
    open(bad, "wb").write(bytes(blob[: len(blob) // 3]))
    with pytest.raises(_lib.B200RTError, match="truncated"):
        scene.CachedHostScene(bad)
    open(bad, "wb").write(b"<Scene></Scene>" * 20)
    with pytest.raises(_lib.B200RTError, match="not a b200rt scene cache"):
        scene.CachedHostScene(bad)
    with pytest.raises(_lib.B200RTError, match="cannot open"):
        scene.CachedHostScene(str(tmp_path / "missing.b200rt"))


@pytest.mark.gpu
def test_cached_scene_renders_identically(gpu, tmp_path):
    tree, xml = scene_xml("hw7_whitted_160")
    sc = scene.Scene(xml, tree, base_dir=ASSETS)
    path = str(tmp_path / "scene.b200rt")
    sc.host.save(path)
    sc2 = scene.Scene.from_cache(path)
    a, sa = sc.render_image(0)
    b, sb = sc2.render_image(0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert (sa.rays_primary, sa.rays_shadow, sa.rays_secondary) == (sb.rays_primary, sb.rays_shadow, sb.rays_secondary)
    ia, ta = sc.primary_ids(0)
    ib, tb = sc2.primary_ids(0)
    assert np.array_equal(ia, ib) and np.array_equal(ta.view(np.uint32), tb.view(np.uint32))


@pytest.mark.parametrize("threads", ["2", "3", "8", "64"])
def test_parallel_builder_is_byte_identical_to_sequential(tmp_path, monkeypatch, threads):
    """The host builder hands the two halves of a large range to different threads (both trees: the reference-topology
    BVH whose depth-first order ids, gates and ties hang on, and the SAH tree the device walks).  Depth-first slots are
    fixed before the halves are built, so the flattened scene must not depend on the thread count: every array byte for
    byte (80 000 triangles — above the 16 384-object threshold where the split begins, so several levels fork)."""
    spec = scene.parse_xml(scenes.cfg2_hw3(n=200, width=64, height=36), "hw3")
    monkeypatch.setenv("B200RT_BUILD_THREADS", "1")
    seq = scene.HostScene(spec)
    assert seq.desc.n_tris >= 4 * 16384
    monkeypatch.setenv("B200RT_BUILD_THREADS", threads)
    par = scene.HostScene(spec)
    a, b = _arrays(seq.desc), _arrays(par.desc)
    for k in a:
        assert a[k] == b[k], k
    assert seq.desc.root == par.desc.root and list(seq.desc.root_box) == list(par.desc.root_box)
    p0, p1 = str(tmp_path / "seq.b200rt"), str(tmp_path / "par.b200rt")
    seq.save(p0)
    par.save(p1)
    assert open(p0, "rb").read() == open(p1, "rb").read()


@pytest.mark.gpu
def test_reference_flattened_scene_renders_like_the_oracle(gpu):
    """The drop-in path end to end: tests/golden/flat3_fuzz0.b200rt is the description integration/flatten_hw3.cpp
    made from the `Scene` the reference's own loader built (tests/test_flatten_reference.py).  It carries no SAH tree —
    b200rt_scene_create builds it — and the device's ids, t and ray counts are the oracle's on the same description, the
    image to 1 LSB (the bar of test_gpu_parity.py::test_fuzz_scenes_match_the_oracle, which renders this same scene
    from the builder's description)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flat3_fuzz0.b200rt")
    sc = scene.Scene.from_cache(path)
    assert sc.host.desc.n_fast_tris == 0
    o_rgbw, o_ids, o_t, cnt = cpu.render(sc.host, 0, want_ids=True)
    for exact in (0, 1):
        ids, t = sc.primary_ids(0, exact_traversal=exact)
        assert np.array_equal(ids, o_ids) and np.array_equal(t.view(np.uint32), o_t.view(np.uint32))
    rgbw, st = sc.render_image(0)
    assert st.nodes_visited > 0
    assert (st.rays_primary, st.rays_shadow, st.rays_secondary) == (cnt.rays_primary, cnt.rays_shadow, cnt.rays_secondary)
    cmp = image.compare_8bit(image.quantise(rgbw, sc.semantics), image.quantise(o_rgbw, sc.semantics))
    assert cmp["within1"] >= 0.999, cmp
