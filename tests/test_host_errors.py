"""Error behaviour of the host side, next to the reference's where the reference has one: a singular object transform
ends the reference's loader with `NOT INVERTIBLE MATRIX!` and exit(-1) (HW3/src/Transformation.cpp:76-82); the builder
refuses the same scene with the same words instead of handing the device a matrix of infinities.  Indices outside the
vertex array are undefined behaviour in the reference (vector::operator[]); here they are an error with a name."""
import os
import subprocess

import numpy as np
import pytest

from ceng795_b200 import _lib, scene, scenes
from oracle import cpu


def _singular(xml: str) -> str:
    # flatten every Scaling of the scene to the plane z = 0: M has no inverse
    import re
    out, n = re.subn(r"(<Scaling id=\"\d+\">)([^<]*)(</Scaling>)", lambda m: m.group(1) + "1 1 0" + m.group(3), xml)
    assert n >= 1
    return out


@pytest.mark.parametrize("tree", ["hw3", "hw7"])
def test_singular_transform_is_refused_like_the_reference(tmp_path, tree):
    xml = _singular(scenes.fuzz_scene(0))
    with pytest.raises(_lib.B200RTError, match="NOT INVERTIBLE MATRIX"):
        scene.HostScene(scene.parse_xml(xml, tree))
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    p = tmp_path / "scene.xml"
    p.write_text(xml)
    exe = os.path.join(cpu.REF_DIR, f"{tree}_raytracer")
    r = subprocess.run([exe, str(p)], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "NOT INVERTIBLE MATRIX" in r.stderr
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".png")]


def test_indices_outside_the_vertex_array_have_a_name():
    xml = scenes.cfg2_hw3(n=8, width=16, height=12)
    spec = scene.parse_xml(xml, "hw3")
    spec.meshes[0].faces = spec.meshes[0].faces.copy()
    spec.meshes[0].faces[3, 1] = len(spec.vertices) + 7
    with pytest.raises(_lib.B200RTError, match="mesh face index out of range"):
        scene.HostScene(spec)
    spec = scene.parse_xml(xml, "hw3")
    spec.spheres[0]["center"] = -5
    with pytest.raises(_lib.B200RTError, match="sphere centre index out of range"):
        scene.HostScene(spec)


def test_builder_rejects_what_a_tree_does_not_have():
    """HW1 / HW2 have no instancing: MeshInstance elements are ignored by their loaders (the parser drops them), and
    the C ABI itself answers UNSUPPORTED rather than building something those trees could not have produced."""
    import ctypes as C
    L = _lib.load()
    b = L.b200rt_builder_create(2)      # B200RT_SEM_HW2
    assert b
    try:
        v = np.zeros((3, 3), np.float32)
        assert L.b200rt_builder_set_vertices(b, v.ctypes.data, 3) == 0
        f = np.array([[0, 1, 2]], np.int32)
        m = L.b200rt_builder_add_mesh(b, 0, 0, f.ctypes.data, 1, None, None)
        assert m == 0
        ident = np.eye(4, dtype=np.float32).ravel()
        rc = L.b200rt_builder_add_instance(b, 0, 0, ident.ctypes.data_as(C.POINTER(C.c_float)), None)
        assert rc == _lib.E_UNSUPPORTED if hasattr(_lib, "E_UNSUPPORTED") else rc < 0
    finally:
        L.b200rt_builder_destroy(b)


def test_multi_gpu_entry_points_refuse_bad_arguments_without_touching_a_device():
    """The multi-GPU entry points of the C ABI check their arguments before anything else: error codes and a message,
    no crash, on a box without a GPU too."""
    import ctypes as C
    L = _lib.load()
    st = _lib.RenderStats()
    out = np.zeros(4, np.float32)
    null_scenes = (C.c_void_p * 2)(None, None)
    assert L.b200rt_render_multi(null_scenes, 2, 0, None, out.ctypes.data, C.byref(st)) < 0
    assert b"null scene" in L.b200rt_last_error()
    assert L.b200rt_render_multi(null_scenes, 0, 0, None, out.ctypes.data, C.byref(st)) < 0
    assert L.b200rt_render_multi(null_scenes, 17, 0, None, out.ctypes.data, C.byref(st)) < 0
    assert b"1..16 scenes" in L.b200rt_last_error()
    f = _lib.SharedFrame()
    for fn, args in (("b200rt_frame_export", (None, 0, C.byref(f))), ("b200rt_frame_attach", (None, C.byref(f))),
                     ("b200rt_frame_detach", (None,)), ("b200rt_frame_begin", (None,)),
                     ("b200rt_render_shared", (None, 0, None, C.byref(st))),
                     ("b200rt_frame_collect", (None, 1, None, None)), ("b200rt_frame_collect_async", (None, 1, out.ctypes.data)),
                     ("b200rt_render_async", (None, 0, None, out.ctypes.data, C.byref(st))), ("b200rt_render_wait", (None,)),
                     ("b200rt_ppm_set_share", (None, 0, 1)), ("b200rt_ppm_counts_apply", (None,)),
                     ("b200rt_ppm_eye_pass_multi", (null_scenes, 2, 0, 0, None)),
                     ("b200rt_ppm_trace_photons_multi", (null_scenes, 2, None, None)),
                     ("b200rt_ppm_density_multi", (null_scenes, 2, 1.0, out.ctypes.data)),
                     ("b200rt_scene_refit", (None, None, None))):
        assert getattr(L, fn)(*args) < 0, fn
    o = _lib.RenderOpts()
    assert L.b200rt_multi_share(None, 0, 0, 1, C.byref(o)) < 0
