"""Edge cases of the hot path (GPU vs oracle, bit-exact where the arithmetic is deterministic): empty scene,
a scene whose root is a single object (create_bvh size==1, including the lone-sphere negative-t quirk of
SURVEY Appendix B.10), image sizes that do not tile into 8x4 blocks, no lights, depth 0, a camera that sees
nothing, axis-parallel primary rays."""
import numpy as np
import pytest

from ceng795_b200 import image, scene, scenes
from oracle import cpu

pytestmark = pytest.mark.gpu

CAM = ("<Cameras><Camera id='1'><Position>{pos}</Position><Gaze>{gaze}</Gaze><Up>0 1 0</Up>"
       "<NearPlane>-1 1 -1 1</NearPlane><NearDistance>1</NearDistance><ImageResolution>{w} {h}</ImageResolution>"
       "<ImageName>a.png</ImageName></Camera></Cameras>")
MATS = ("<Materials><Material id='1'><AmbientReflectance>1 1 1</AmbientReflectance>"
        "<DiffuseReflectance>0.6 0.5 0.4</DiffuseReflectance><SpecularReflectance>0.3 0.3 0.3</SpecularReflectance>"
        "<PhongExponent>10</PhongExponent></Material></Materials>")


def _xml(objects, lights="<PointLight id='1'><Position>3 4 5</Position><Intensity>900 900 900</Intensity></PointLight>",
         pos="0 0 5", gaze="0 0 -1", w=24, h=16, depth=1, verts="0 0 0\n-1 -1 0\n1 -1 0\n0 1 0"):
    return ("<Scene><BackgroundColor>7 8 9</BackgroundColor><MaxRecursionDepth>%d</MaxRecursionDepth>" % depth
            + CAM.format(pos=pos, gaze=gaze, w=w, h=h)
            + "<Lights><AmbientLight>20 20 20</AmbientLight>" + lights + "</Lights>" + MATS
            + "<VertexData>" + verts + "</VertexData><Objects>" + objects + "</Objects></Scene>")


SPHERE = "<Sphere id='1'><Material>1</Material><Center>1</Center><Radius>1</Radius></Sphere>"
MESH1 = "<Mesh id='1'><Material>1</Material><Faces>2 3 4</Faces></Mesh>"


def _check(xml, tree, exact_bits=True):
    sc = scene.Scene(xml, tree)
    rgbw, st = sc.render_image(0)
    o_rgbw, o_ids, o_t, cnt = cpu.render(sc.host, 0, want_ids=True)
    ids, t = sc.primary_ids(0)
    assert np.array_equal(ids, o_ids)
    assert np.array_equal(t.view(np.uint32), o_t.view(np.uint32))
    assert (st.rays_primary, st.rays_shadow, st.rays_secondary) == (cnt.rays_primary, cnt.rays_shadow, cnt.rays_secondary)
    cmp = image.compare_8bit(image.quantise(rgbw, sc.semantics), image.quantise(o_rgbw, sc.semantics))
    assert cmp["within1"] == 1.0, cmp
    if exact_bits:
        assert np.mean(rgbw.view(np.uint32) == o_rgbw.view(np.uint32)) >= 0.995
    return rgbw, ids


@pytest.mark.parametrize("tree", ["hw1", "hw3", "hw7"])
def test_empty_scene_is_background(gpu, tree):
    rgbw, ids = _check(_xml(""), tree)
    assert np.all(ids == -1) and np.all(rgbw[..., :3] == np.array([7, 8, 9], np.float32))


@pytest.mark.parametrize("tree", ["hw3", "hw7"])
def test_single_object_scenes(gpu, tree):
    _check(_xml(SPHERE), tree)                       # root is the sphere
    _check(_xml(MESH1), tree)                        # root is an instance of a one-triangle mesh
    # camera inside / behind: rays whose line meets the lone sphere at t < 0 still "hit" it in the reference
    _, ids = _check(_xml(SPHERE, pos="0 0 3", gaze="0 0 1"), tree)
    assert (ids[..., 0] == 0).any()


@pytest.mark.parametrize("tree,size", [("hw1", (37, 23)), ("hw3", (37, 23)), ("hw7", (41, 19)), ("hw3", (8, 3))])
def test_sizes_that_do_not_tile(gpu, tree, size):
    if tree == "hw1":
        xml = scenes.cfg1_hw1(width=size[0], height=size[1], n_spheres=6, sphere_mesh=(6, 6))
    else:
        xml = scenes.cfg2_hw3(n=10, width=size[0], height=size[1])
    _check(xml, tree if tree != "hw7" else "hw3")


@pytest.mark.parametrize("tree", ["hw1", "hw3", "hw7"])
def test_no_lights_depth0_and_blind_camera(gpu, tree):
    _check(_xml(SPHERE + MESH1, lights="", depth=0), tree)
    _, ids = _check(_xml(SPHERE + MESH1, gaze="0 0 1"), tree)      # looking away: every primary ray misses
    assert np.all(ids == -1)


def test_axis_parallel_primary_rays(gpu):
    """Rays with direction components under the reference's axis-skip threshold (|d| < 1e-6): the centre column
    and row of an odd-sized image looking straight down an axis."""
    xml = scenes.cfg2_hw3(n=12, width=33, height=33).replace("<Gaze>0 -0.62 -1</Gaze>", "<Gaze>0 -1 0</Gaze>") \
        .replace("<Up>0 1 0</Up>", "<Up>0 0 -1</Up>")
    _check(xml, "hw3")

