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
FLATTEN7 = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "flat7_flatten")
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


def _lights(hs):
    """lights[] as bytes, with the `object` of mesh / sphere lights (an instance / sphere LABEL) rewritten to the
    depth-first rank of that object."""
    d = hs.desc
    out = []
    for i in range(d.n_lights):
        l = _lib.Light.from_buffer_copy(C.string_at(C.addressof(d.lights[i]), C.sizeof(_lib.Light)))
        if l.kind == _lib.LIGHT_MESH:
            l.object = d.instances[l.object].rank
        elif l.kind == _lib.LIGHT_SPHERE:
            l.object = d.spheres[l.object].rank
        out.append(bytes(l))
    return out


def _assert_same_description(ref, own, skip=()):
    a, b = _arrays(ref.desc), _arrays(own.desc)
    for k in PLAIN:
        if k in skip:
            continue
        if k == "materials":      # `degamma` is an INPUT flag of the builder (values already converted); the objects keep no trace of it
            x, y = (np.frombuffer(v, np.uint32).reshape(-1, 24).copy() for v in (a[k], b[k]))
            x[:, 19] = y[:, 19] = 0
            assert np.array_equal(x, y), k
        elif k == "lights":
            assert _lights(ref) == _lights(own), k
        elif k == "textures":     # Perlin tables: shuffled by a clock-seeded engine in every run of the reference
            x, y = (np.frombuffer(v, np.uint32).reshape(-1, C.sizeof(_lib.Texture) // 4).copy() for v in (a[k], b[k]))
            o = _lib.Texture.perm.offset // 4
            for row in x:
                if row[0] == _lib.TEX_PERLIN:
                    assert sorted(row[o:o + 16].tolist()) == list(range(16))
            x[:, o:o + 16] = y[:, o:o + 16] = 0
            assert np.array_equal(x, y), k
        else:
            assert a[k] == b[k], k
    for f in ("semantics", "max_depth", "shadow_eps", "isect_eps", "n_tris", "n_nodes", "n_verts", "n_meshes",
              "n_instances", "n_spheres", "n_triprims", "n_lights", "n_textures", "n_texels", "n_brdfs", "n_light_cdf",
              "background_texture", "env_light", "integrator", "importance_sampling"):
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


def _flatten(xml, tmp_path, exe=FLATTEN):
    p = str(tmp_path / "scene.xml")
    with open(p, "w") as f:
        f.write(xml)
    out = str(tmp_path / "flat.b200rt")
    r = subprocess.run([exe, p, out], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))   # textures: relative paths
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


needs_ref7 = pytest.mark.skipif(not os.path.exists(FLATTEN7), reason="oracle/_ref/flat7_flatten not built on this box")


@needs_ref7
@pytest.mark.parametrize("case", ["textured", "lights_pathtraced", "perlin", "env_light", "background", "camera_left",
                                  "camera_simple", "fuzz_open", "fuzz_stochastic"])
def test_flattened_hw7_scene_equals_builder(tmp_path, case):
    """integration/flatten_hw7.cpp on the HW7 tree's own objects: everything flatten_hw3 covers plus image textures (texels
    as `Texture` holds them after degamma), Perlin textures, texture coordinates per triangle, the five BRDF classes,
    the polymorphic light list in the order the shading loop walks it — point, area, spot, directional, environment
    (its EXR texels), mesh lights with their area CDF over the BVH-reordered triangles, sphere lights — glass instances'
    culling flag, the left-handed and the `simple` camera, the integrator switch.  Only two things are not compared
    bit for bit, because the objects do not hold them: the builder's `degamma` input flag, and the Perlin table (the
    reference shuffles a fresh one from the clock in every run; it must be a permutation of 0..15)."""
    import re
    d = str(tmp_path)
    kw = dict(width=48, height=36, n=6, depth=2, mesh_slices=(6, 7))
    if case == "textured":
        xml = scenes.hw7_scene(d, width=64, height=48, light_scale=3.0)
    elif case == "lights_pathtraced":
        xml = scenes.hw7_scene(d, spp=4, stochastic_lights=True, path_tracing=True, importance=True, light_scale=0.5,
                               **{**kw, "depth": 3})
    elif case == "perlin":
        xml = scenes.hw7_scene(d, perlin=True, **kw)
    elif case == "env_light":
        from ceng795_b200 import image
        rng = np.random.default_rng(3)
        image.write_exr_f32(str(tmp_path / "env.exr"), (rng.uniform(0, 1, (16, 32, 3)) ** 3 * 40).astype(np.float32))
        xml = scenes.hw7_scene(d, **kw).replace(
            "  </Lights>", '    <SphericalDirectionalLight id="1"><EnvMapName>env.exr</EnvMapName>'
                           "</SphericalDirectionalLight>\n  </Lights>")
    elif case == "background":
        xml = scenes.hw7_scene(d, **kw).replace(
            "<Scene>\n", "<Scene>\n  <BackgroundTexture>tex_checker.png</BackgroundTexture>\n", 1)
    elif case == "camera_left":
        xml = scenes.hw7_scene(d, **kw).replace('<Camera id="1">', '<Camera id="1" handedness="left">')
    elif case == "camera_simple":
        xml = scenes.hw7_scene(d, **kw)
        xml = re.sub(r"      <Gaze>.*?</Gaze>\n", "      <GazePoint>0.2 0.6 -1.5</GazePoint>\n", xml)
        xml = re.sub(r"      <NearPlane>.*?</NearPlane>\n", "      <FovY>52</FovY>\n", xml)
        xml = xml.replace('<Camera id="1">', '<Camera id="1" type="simple">')
    elif case == "fuzz_open":
        xml = scenes.fuzz_scene(2, camera_sphere_radius=40.0)
    else:
        xml = scenes.fuzz_scene(1, spp=9, stochastic=True, camera_sphere_radius=40.0)
    ref = scene.CachedHostScene(_flatten(xml, tmp_path, FLATTEN7))
    own = scene.HostScene(scene.parse_xml(xml, "hw7", base_dir=d))
    assert ref.desc.semantics == own.desc.semantics == 7
    _assert_same_description(ref, own)


FLATTEN_PPM = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "flatppm_flatten")


@pytest.mark.skipif(not os.path.exists(FLATTEN_PPM), reason="oracle/_ref/flatppm_flatten not built on this box")
@pytest.mark.parametrize("kw", [dict(width=64, height=48, mesh_slices=(12, 10)), dict(width=40, height=30, mesh_slices=(31, 29), seed=5)])
def test_flattened_ppm_scene_equals_builder(tmp_path, kw):
    """integration/flatten_ppm.cpp on the PPM tree's own objects (cfg5's scene at small size: glass and mirror spheres,
    mesh instances with the culling-off flag, point light; spheres ahead of the meshes in Scene::bvh).  It is also how
    the builder learned that this tree normalises the normal of EVERY vertex, so that a vertex no face uses (a sphere
    centre) holds 0/0 — bit-equal now, NaN included."""
    kw = dict(kw)
    seed = kw.pop("seed", None)
    xml = scenes.cfg5_ppm(photons_per_iteration=1000, iterations=2, **({"seed": seed} if seed is not None else {}), **kw)
    ref = scene.CachedHostScene(_flatten(xml, tmp_path, FLATTEN_PPM))
    own = scene.HostScene(scene.parse_xml(xml, "ppm"))
    assert ref.desc.semantics == own.desc.semantics == 8
    _assert_same_description(ref, own)
    assert np.isnan(np.frombuffer(_arrays(own.desc)["vert_normal"], np.float32)).any()


@needs_ref
@needs_ref7
def test_flatten_equals_builder_on_twenty_more_random_scenes(tmp_path):
    """A sweep rather than a parametrisation: 20 more fuzz seeds (open and closed, deterministic and stochastic
    features alternating) through both trees' flattens."""
    for seed in range(10, 30):
        xml = scenes.fuzz_scene(seed, stochastic=bool(seed % 2), spp=4 if seed % 2 else 1,
                                camera_sphere_radius=40.0 if seed % 3 else 2.5)
        for tree, exe in (("hw3", FLATTEN), ("hw7", FLATTEN7)):
            d = tmp_path / f"{tree}_{seed}"
            d.mkdir()
            ref = scene.CachedHostScene(_flatten(xml, d, exe))
            own = scene.HostScene(scene.parse_xml(xml, tree))
            _assert_same_description(ref, own)


FLATTEN1 = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "flat1_flatten")


@pytest.mark.skipif(not os.path.exists(FLATTEN1), reason="oracle/_ref/flat1_flatten not built on this box")
@pytest.mark.parametrize("case", ["cfg1", "near_ties", "fuzz_open"])
def test_flattened_hw1_scene_equals_builder(tmp_path, case):
    """integration/flatten_hw1.cpp on the HW1 tree's objects (cfg1's dialect: no BVH, no transforms; the scan order
    meshes -> triangles -> spheres that HW1's order-dependent hit rule depends on).  Every array byte for byte,
    including the object structs (nothing is renumbered here); the only difference is su / sv of the camera, which
    the builder fills for uniformity and HW1's Camera does not have (the device does not read them for this tree)."""
    xml = {"cfg1": lambda: scenes.cfg1_hw1(width=96, height=96),
           "near_ties": lambda: scenes.hw1_near_ties(),
           "fuzz_open": lambda: scenes.fuzz_scene(1, camera_sphere_radius=40.0)}[case]()
    ref = scene.CachedHostScene(_flatten(xml, tmp_path, FLATTEN1))
    own = scene.HostScene(scene.parse_xml(xml, "hw1"))
    a, b = _arrays(ref.desc), _arrays(own.desc)
    for k in a:
        if k == "cameras":
            x, y = (np.frombuffer(v, np.uint32).reshape(-1, C.sizeof(_lib.Camera) // 4).copy() for v in (a[k], b[k]))
            o = _lib.Camera.su.offset // 4
            assert not x[:, o:o + 6].any()
            x[:, o:o + 6] = y[:, o:o + 6] = 0
            assert np.array_equal(x, y)
        else:
            assert a[k] == b[k], k
    for f in ("semantics", "max_depth", "shadow_eps", "isect_eps", "root", "n_tris", "n_nodes", "n_meshes", "n_spheres",
              "n_instances", "n_triprims", "n_lights"):
        assert getattr(ref.desc, f) == getattr(own.desc, f), f
    for f in ("background", "ambient", "root_box"):
        assert list(getattr(ref.desc, f)) == list(getattr(own.desc, f)), f
    # and the oracle sees the same scene: ids, t and image bits
    img_r, ids_r, t_r, _ = cpu.render(ref, 0, want_ids=True)
    img_o, ids_o, t_o, _ = cpu.render(own, 0, want_ids=True)
    assert np.array_equal(ids_r, ids_o) and np.array_equal(t_r.view(np.uint32), t_o.view(np.uint32))
    assert np.array_equal(img_r.view(np.uint32), img_o.view(np.uint32))


@pytest.mark.parametrize("tree", ["hw4", "hw5", "hw6"])
@pytest.mark.parametrize("case", ["textured", "perlin", "fuzz_open", "fuzz_stochastic"])
def test_flattened_hw4_hw5_hw6_scene_equals_builder(tmp_path, tree, case):
    """The same flatten source compiled against the HW4, HW5 and HW6 trees (-DB200RT_TREE=4 / 5 / 6: earlier states of
    HW7's classes — bytes instead of floats in `Texture`, no BRDF objects before HW6, HW3's plain light vectors in HW4,
    no environment / mesh / sphere lights) on
    the scenes whose loaders differ from HW7's: HW5 / HW6 keep only the FIRST spot and the first directional light,
    convert `degamma` materials with a double pow, and ignore what they do not know."""
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", f"flat{tree[-1]}_flatten")
    if not os.path.exists(exe):
        pytest.skip(f"oracle/_ref/flat{tree[-1]}_flatten not built on this box")
    d = str(tmp_path)
    xml = {"textured": lambda: scenes.hw7_scene(d, width=64, height=48, light_scale=3.0),
           "perlin": lambda: scenes.hw7_scene(d, width=48, height=36, n=6, depth=2, perlin=True, mesh_slices=(6, 7)),
           "fuzz_open": lambda: scenes.fuzz_scene(2, camera_sphere_radius=40.0),
           "fuzz_stochastic": lambda: scenes.fuzz_scene(1, spp=9, stochastic=True, camera_sphere_radius=40.0)}[case]()
    ref = scene.CachedHostScene(_flatten(xml, tmp_path, exe))
    own = scene.HostScene(scene.parse_xml(xml, tree, base_dir=d))
    assert ref.desc.semantics == own.desc.semantics == int(tree[-1])
    _assert_same_description(ref, own)


FLATTEN2 = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "flat2_flatten")


@pytest.mark.skipif(not os.path.exists(FLATTEN2), reason="oracle/_ref/flat2_flatten not built on this box")
@pytest.mark.parametrize("case", ["one_mesh_glass", "one_mesh_small", "three_meshes"])
def test_flattened_hw2_scene_equals_builder(tmp_path, case):
    """integration/flatten_hw2.cpp on the HW2 tree's objects (a `Mesh` sits in Scene::bvh itself; no transforms, no
    instancing, no vertex normals).  `Scene` keeps no mesh list there, so the flatten numbers meshes in the order its
    walk meets them: with one mesh the description equals the builder's array for array (vertex normals aside — the
    builder accumulates them for every tree, HW2 has none and never reads them); with several, triangle and node
    arrays are the same content in another mesh order, and what is held equal is what the oracle makes of both:
    every t and every float of the image."""
    xml = {"one_mesh_glass": lambda: scenes.cfg2_hw3(n=32, width=128, height=96, n_instances=0, depth=4, glass=True),
           "one_mesh_small": lambda: scenes.cfg2_hw3(n=9, width=40, height=30, n_instances=0, depth=2),
           "three_meshes": lambda: scenes.fuzz_scene(1, camera_sphere_radius=40.0)}[case]()
    ref = scene.CachedHostScene(_flatten(xml, tmp_path, FLATTEN2))
    own = scene.HostScene(scene.parse_xml(xml, "hw2"))
    assert ref.desc.semantics == own.desc.semantics == 2
    if case != "three_meshes":
        assert ref.desc.n_meshes == 1
        _assert_same_description(ref, own, skip=("vert_normal",))
    else:
        assert ref.desc.n_meshes == own.desc.n_meshes == 3 and ref.desc.n_tris == own.desc.n_tris
        assert ref.desc.n_nodes == own.desc.n_nodes
    img_r, ids_r, t_r, _ = cpu.render(ref, 0, want_ids=True)
    img_o, ids_o, t_o, _ = cpu.render(own, 0, want_ids=True)
    assert np.array_equal(t_r.view(np.uint32), t_o.view(np.uint32))
    assert np.array_equal(img_r.view(np.uint32), img_o.view(np.uint32))
    assert np.array_equal(ids_r[..., 0], ids_o[..., 0])       # the top-level rank of every hit (mesh-local ids follow the mesh numbering)
