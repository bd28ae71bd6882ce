"""Refit for geometry that moves between frames (include/b200rt.h "refit").

The reference has one answer to a moved object — parse and build everything again — so the bar is: after
move_instance / move_sphere + refit, the flattened description equals a FRESH build of the moved scene byte for byte
(CPU), and the resident device scene renders the fresh scene's frame bit for bit (GPU), while only the top-level half
(kilobytes) is rebuilt and copied."""
import copy
import ctypes as C
import time

import numpy as np
import pytest

from ceng795_b200 import _lib, scene, scenes
from tests.test_scene_cache import _arrays


def _moved_spec(spec, k, delta):
    """The XML scene with MeshInstance k translated by delta (appended to its transform list) and sphere 0 too."""
    s2 = copy.deepcopy(spec)
    s2.translations = list(s2.translations) + [np.asarray(delta, np.float32)]
    t_idx = len(s2.translations) - 1
    s2.instances[k].transforms = list(s2.instances[k].transforms) + [("t", t_idx)]
    s2.spheres[0]["transforms"] = list(s2.spheres[0]["transforms"]) + [("t", t_idx)]
    return s2


def _refit_like(hs, spec2, k):
    """Applies spec2's transforms of instance k and sphere 0 to the already-built host scene hs."""
    inst = spec2.instances[k]
    tmp = scene.HostScene.__new__(scene.HostScene)       # _compose needs only the library and the spec's tables
    tmp._L, tmp.spec = hs._L, spec2
    base = None if inst.reset_transform else tmp._compose(spec2.meshes[inst.base_mesh].transforms, None)
    hs.move_instance(len(spec2.meshes) + k, tmp._compose(inst.transforms, base), inst.velocity)
    hs.move_sphere(0, tmp._compose(spec2.spheres[0]["transforms"], None), spec2.spheres[0]["velocity"])
    tmp._b = None


@pytest.mark.parametrize("tree,xml", [("hw3", scenes.cfg2_hw3(n=48, width=96, height=60)),
                                      ("hw3", scenes.cfg2_hw3(n=20, width=64, height=40, n_instances=3))])
def test_refit_equals_a_fresh_build_of_the_moved_scene(tree, xml):
    spec = scene.parse_xml(xml, tree)
    hs = scene.HostScene(spec)
    before = _arrays(hs.desc)
    for step, delta in enumerate([(0.35, 0.1, -0.2), (-1.5, 0.0, 2.25)]):   # two frames of an animation
        k = step % max(1, len(spec.instances))
        spec2 = _moved_spec(spec, k, delta) if spec.instances else None
        if spec2 is None:
            pytest.skip("scene without instances")
        fresh = scene.HostScene(spec2)
        _refit_like(hs, spec2, k)
        t0 = time.perf_counter()
        hs.refit()
        dt = time.perf_counter() - t0
        a, b = _arrays(hs.desc), _arrays(fresh.desc)
        for name in a:
            assert a[name] == b[name], f"{name} differs after refit (frame {step})"
        assert (hs.desc.root, list(hs.desc.root_box)) == (fresh.desc.root, list(fresh.desc.root_box))
        # the mesh half was not rebuilt: same bytes as before the move
        for name in ("tri_geom", "tri_gate", "fast_nodes", "fast_tri_geom", "fast_tri_ref", "vert_normal", "meshes"):
            assert a[name] == before[name]
        assert a["instances"] != before["instances"] and a["nodes"] != before["nodes"]
        assert dt < 0.05, dt                                 # milliseconds: the top level has a handful of objects
        # undo for the next frame: the fresh scene of the next step starts from the original spec
        hs.close()
        hs = scene.HostScene(spec)
        fresh.close()


def test_refit_of_an_hw7_scene_rebuilds_the_light_objects_too(tmp_path):
    """HW7: LightMesh / LightSphere are top-level objects that add a light each (with the mesh light's area CDF); the
    refit has to drop and rebuild exactly those.  Moving an ordinary instance and a sphere, then refitting, equals a
    fresh build — lights, CDF and all."""
    xml = scenes.hw7_scene(str(tmp_path), width=48, height=36, stochastic_lights=True, perlin=True)
    spec = scene.parse_xml(xml, "hw7", base_dir=str(tmp_path))
    assert spec.light_meshes and spec.light_spheres and spec.instances
    hs = scene.HostScene(spec)
    n_lights = hs.desc.n_lights
    spec2 = _moved_spec(spec, 0, (0.2, -0.1, 0.3))
    fresh = scene.HostScene(spec2)
    _refit_like(hs, spec2, 0)
    hs.refit()
    hs.refit()                                               # idempotent: nothing accumulates from call to call
    a, b = _arrays(hs.desc), _arrays(fresh.desc)
    for name in a:
        assert a[name] == b[name], f"{name} differs after refit"
    assert hs.desc.n_lights == fresh.desc.n_lights == n_lights and hs.desc.n_light_cdf == fresh.desc.n_light_cdf > 0


def test_refit_refuses_nonsense():
    hs = scene.HostScene(scene.parse_xml(scenes.cfg2_hw3(n=16, width=32, height=20), "hw3"))
    ident = np.eye(4, dtype=np.float32).reshape(16)
    with pytest.raises(_lib.B200RTError, match="bad instance index"):
        hs.move_instance(999, ident)
    with pytest.raises(_lib.B200RTError, match="bad sphere index"):
        hs.move_sphere(-1, ident)
    hs.move_instance(1, np.zeros(16, np.float32))            # singular: the reference prints NOT INVERTIBLE MATRIX
    with pytest.raises(_lib.B200RTError, match="NOT INVERTIBLE"):
        hs.refit()


@pytest.mark.gpu
def test_refit_on_the_device_renders_the_fresh_scene_bit_for_bit():
    L = _lib.load()
    if L.b200rt_device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200")
    xml = scenes.cfg2_hw3(n=128, width=320, height=200)
    sc = scene.Scene(xml, "hw3")
    first, _ = sc.render_image(0)
    spec = sc.spec
    ms_all = []
    for step, delta in enumerate([(0.35, 0.1, -0.2), (-0.4, 0.3, 0.5), (0.0, -0.25, 0.1)]):
        k = step % len(spec.instances)
        spec2 = _moved_spec(spec, k, delta)
        _refit_like(sc.host, spec2, k)
        ms_all.append(sc.refit())
        got, st = sc.render_image(0)
        ids, t = sc.primary_ids(0)
        fresh = scene.Scene.__new__(scene.Scene)
        fresh.spec, fresh.host, fresh.semantics, fresh.device, fresh._scene, fresh._L = (
            spec2, scene.HostScene(spec2), spec2.semantics, 0, None, L)
        fresh.upload()
        want, st2 = fresh.render_image(0)
        ids2, t2 = fresh.primary_ids(0)
        assert np.array_equal(ids, ids2) and np.array_equal(t.view(np.uint32), t2.view(np.uint32))
        assert st.rays_total == st2.rays_total
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert not np.array_equal(got, first)                # the object did move
        spec = spec2
        sc.spec = spec2
        fresh.close()
    print(f"refit: device copy {ms_all} ms per frame")
