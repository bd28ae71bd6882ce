"""CPU: the C-ABI library loads, exports every symbol include/b200rt.h declares,
mirrors the header's struct sizes, and fails loudly without a GPU."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from ceng795_b200 import _build, _lib, scene, scenes


def test_exports_every_declared_symbol():
    L = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in b200rt.h but not exported"


def test_struct_sizes_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "b200rt.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(b200rt_bvh_node),sizeof(b200rt_mesh),sizeof(b200rt_instance),sizeof(b200rt_sphere),"
                   "sizeof(b200rt_triprim),sizeof(b200rt_material),sizeof(b200rt_light),sizeof(b200rt_camera),"
                   "sizeof(B200SceneDesc),sizeof(B200RenderOpts),sizeof(B200RenderStats));"
                   'printf("%zu %zu %zu %zu %zu\\n",sizeof(b200rt_texture),sizeof(b200rt_brdf),sizeof(b200rt_camera_in),'
                   "sizeof(B200PpmOpts),sizeof(B200PpmStats));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", _build.ROOT + "/include", str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(t) for t in (_lib.BvhNode, _lib.Mesh, _lib.Instance, _lib.Sphere, _lib.TriPrim, _lib.Material,
                                  _lib.Light, _lib.Camera, _lib.SceneDesc, _lib.RenderOpts, _lib.RenderStats,
                                  _lib.Texture, _lib.Brdf, _lib.CameraIn, _lib.PpmOpts, _lib.PpmStats)]
    assert got == want
    assert got[0] == 64 and got[5] == 96


def test_transform_helpers_follow_reference_arithmetic():
    L = _lib.load()
    fp = _lib.fptr
    a = np.zeros(16, np.float32); b = np.zeros(16, np.float32); o = np.zeros(16, np.float32)
    L.b200rt_xf_rotation(37.0, 1.0, 2.0, 3.0, fp(a))
    L.b200rt_xf_rotation(-37.0, 1.0, 2.0, 3.0, fp(b))
    L.b200rt_xf_mul(fp(a), fp(b), fp(o))
    assert np.allclose(o.reshape(4, 4), np.eye(4), atol=1e-6)
    L.b200rt_xf_scaling(2.0, 4.0, 8.0, fp(a))
    assert L.b200rt_xf_invert(fp(a), fp(b)) == 0
    assert np.array_equal(b.reshape(4, 4).diagonal(), np.array([0.5, 0.25, 0.125, 1.0], np.float32))
    a[:] = 0
    assert L.b200rt_xf_invert(fp(a), fp(b)) < 0  # reference: "NOT INVERTIBLE MATRIX" + exit(-1)


def test_bvh_topology_invariants():
    xml = scenes.cfg2_hw3(n=16, width=32, height=18)
    hs = scene.HostScene(scene.parse_xml(xml, "hw3"))
    d = hs.desc
    assert d.n_tris == 2 * 16 * 16
    # N triangles -> N-1 inner nodes for the mesh, M objects -> M-1 for the top level
    n_top = d.n_instances + d.n_spheres + d.n_triprims
    assert d.n_nodes == (d.n_tris - 1) + (n_top - 1)
    nodes = hs.nodes()
    refs = np.concatenate([nodes[:, 12], nodes[:, 13]])
    kinds = refs >> 28
    tri_refs = np.sort(refs[kinds == _lib.KIND_TRI] & 0x0FFFFFFF)
    assert np.array_equal(tri_refs, np.arange(d.n_tris))          # every triangle exactly once
    ranks = sorted([d.instances[i].rank for i in range(d.n_instances)] +
                   [d.spheres[i].rank for i in range(d.n_spheres)])
    assert ranks == list(range(n_top))                             # DFS ranks are a permutation


def test_empty_and_single_object_scenes():
    # a scene with one sphere: Scene::bvh is the object itself (create_bvh size==1)
    xml = ("<Scene><MaxRecursionDepth>1</MaxRecursionDepth><Cameras><Camera id='1'><Position>0 0 5</Position>"
           "<Gaze>0 0 -1</Gaze><Up>0 1 0</Up><NearPlane>-1 1 -1 1</NearPlane><NearDistance>1</NearDistance>"
           "<ImageResolution>8 8</ImageResolution><ImageName>a.png</ImageName></Camera></Cameras>"
           "<Lights><AmbientLight>1 1 1</AmbientLight></Lights><Materials><Material id='1'>"
           "<AmbientReflectance>1 1 1</AmbientReflectance></Material></Materials>"
           "<VertexData>0 0 0</VertexData><Objects><Sphere id='1'><Material>1</Material><Center>1</Center>"
           "<Radius>1</Radius></Sphere></Objects></Scene>")
    hs = scene.HostScene(scene.parse_xml(xml, "hw3"))
    assert hs.desc.n_nodes == 0 and (hs.desc.root >> 28) == _lib.KIND_SPHERE
    xml0 = xml.replace("<Sphere id='1'><Material>1</Material><Center>1</Center><Radius>1</Radius></Sphere>", "")
    hs0 = scene.HostScene(scene.parse_xml(xml0, "hw3"))
    assert (hs0.desc.root >> 28) == _lib.KIND_NONE


def test_no_cpu_fallback_without_gpu():
    L = _lib.load()
    if L.b200rt_device_count() > 0:
        pytest.skip("GPU present")
    xml = scenes.cfg1_hw1(width=8, height=8, n_spheres=2)
    with pytest.raises(_lib.B200RTError):
        scene.Scene(xml, "hw1")


def test_lbvh_hierarchy_rule_on_the_host():
    """csrc/lbvh.cuh: Karras' range / split rule, run on the CPU over sorted keys including
    long runs of equal Morton codes (told apart by position).  The n-1 internal nodes must form a binary tree over the n leaves: node 0 covers
    everything, every node's children partition its range, every leaf and every other node has exactly one parent."""
    import ctypes as C
    import numpy as np
    from ceng795_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(5)
    for n, spread in ((2, 10), (3, 1), (17, 4), (1000, 50), (4096, 1 << 30), (5000, 3)):
        morton = np.sort(rng.integers(0, spread, n).astype(np.uint64))          # many duplicates when spread is small
        keys = np.ascontiguousarray(morton)
        ranges = np.zeros((n - 1, 3), np.int32)
        assert L.b200rt_lbvh_ranges(keys.ctypes.data, n, ranges.ctypes.data) == 0
        first, last, split = ranges[:, 0], ranges[:, 1], ranges[:, 2]
        assert first[0] == 0 and last[0] == n - 1
        assert np.all((first <= split) & (split < last))
        leaf_parents = np.zeros(n, np.int32)
        node_parents = np.zeros(n - 1, np.int32)
        for i in range(n - 1):
            for lo, hi in ((first[i], split[i]), (split[i] + 1, last[i])):
                if lo == hi:
                    leaf_parents[lo] += 1
                else:
                    child = split[i] if lo == first[i] else split[i] + 1      # Karras: the child node sits at the split
                    assert first[child] == lo and last[child] == hi, (n, i, child)
                    node_parents[child] += 1
        assert np.all(leaf_parents == 1) and node_parents[0] == 0 and np.all(node_parents[1:] == 1)
    assert L.b200rt_lbvh_morton63(0.0, 0.0, 0.0) == 0
    assert L.b200rt_lbvh_morton63(1.0, 1.0, 1.0) == (1 << 63) - 1
    assert L.b200rt_lbvh_morton63(0.999, 0.0, 0.0) > L.b200rt_lbvh_morton63(0.0, 0.999, 0.999)   # x is the top bit
