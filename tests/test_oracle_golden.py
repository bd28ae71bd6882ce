"""CPU: the C oracle must reproduce the reference binaries' images byte for byte
(golden PNGs rendered by oracle/_ref/*, see tests/golden/make_golden.py), and —
when the compiled reference is present — live renders too."""
import os

import numpy as np
import pytest

from ceng795_b200 import image, scene
from oracle import cpu
from tests.golden.make_golden import ASSETS, CASES, FLOAT_CASES, float_scene_xml, scene_xml

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden_png(name):
    tree, xml = scene_xml(name)
    spec = scene.parse_xml(xml, tree, base_dir=ASSETS)
    hs = scene.HostScene(spec)
    rgbw, _, _, cnt = cpu.render(hs, 0)
    got = image.quantise(rgbw, spec.semantics)
    want = image.read_png(os.path.join(GOLDEN, name + ".png"))
    assert got.shape == want.shape
    cmp = image.compare_8bit(got, want)
    if tree == "hw4":   # splat order = the reference's thread interleaving (tests/golden/make_golden.py)
        assert cmp["within1"] == 1.0 and cmp["exact"] >= 0.995, cmp
    else:
        assert cmp["exact"] == 1.0, cmp  # bit-exact: same compiler, same IEEE arithmetic
    assert cnt.rays_primary == got.shape[0] * got.shape[1]


@pytest.mark.parametrize("name", sorted(FLOAT_CASES))
def test_oracle_matches_float_golden(name):
    """HW5 / HW6 (and HW7 again) against float32 dumps of Pixel::get_color() from the tree's own objects
    (oracle/ref_driver_hdr.cpp): the restatement reproduces every accumulator to the last bit."""
    sem, _, xml = float_scene_xml(name)
    spec = scene.parse_xml(xml, sem, base_dir=ASSETS)
    rgbw, _, _, cnt = cpu.render(scene.HostScene(spec), 0)
    got = rgbw[..., :3] / rgbw[..., 3:4]
    want = np.load(os.path.join(GOLDEN, name + ".npy"))
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw5", "hw6"])
def test_oracle_hw56_multisample_statistics_vs_reference(tmp_path, tree):
    """HW5 / HW6 multi-sample path (area light, glossy mirror, jitter, sigma-1/2 Gaussian splat): the reference
    seeds from random_device, so equal mean and an RMSE within the oracle's own seed-to-seed noise."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    import re
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=40, height=30, spp=100, depth=2, stochastic_lights=True, light_scale=0.25)
    xml = re.sub(r"    <Light(Mesh|Sphere).*?</Light(Mesh|Sphere)>\n", "", xml, flags=re.S)   # HW7-only objects
    ref = list(cpu.run_reference(tree, xml, workdir=str(tmp_path))["hdr"].values())[0].astype(np.float64)
    hs = scene.HostScene(scene.parse_xml(xml, tree, base_dir=str(tmp_path)))
    imgs = []
    for seed in (7, 8):
        rgbw, _, _, _ = cpu.render(hs, 0, seed=seed)
        imgs.append((rgbw[..., :3] / rgbw[..., 3:4]).astype(np.float64))
    assert abs(imgs[0].mean() - ref.mean()) <= 0.02 * ref.mean(), (imgs[0].mean(), ref.mean())
    noise = np.sqrt(((imgs[0] - imgs[1]) ** 2).mean())
    rmse = np.sqrt(((imgs[0] - ref) ** 2).mean())
    assert rmse <= 1.35 * noise + 0.5, (rmse, noise)


@pytest.mark.parametrize("tree,gen_kw", [("hw1", dict(width=80, height=60, seed=11)),
                                         ("hw3", dict(n=24, width=96, height=64, seed=12, depth=4))])
def test_oracle_matches_live_reference(tree, gen_kw):
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.cfg1_hw1(**gen_kw) if tree == "hw1" else scenes.cfg2_hw3(**gen_kw)
    ref = cpu.run_reference(tree, xml)
    want = list(ref["images"].values())[0]
    spec = scene.parse_xml(xml, tree)
    hs = scene.HostScene(spec)
    rgbw, _, _, _ = cpu.render(hs, 0)
    cmp = image.compare_8bit(image.quantise(rgbw, spec.semantics), want)
    assert cmp["exact"] == 1.0, cmp


@pytest.mark.parametrize("tree", ["hw2", "hw3", "hw7"])
@pytest.mark.parametrize("spacing", [0.0, 6e-7])
def test_oracle_tie_breaking_matches_live_reference(tree, spacing):
    """Exact ties (duplicate coplanar quads, shared edges) and near ties across top-level objects: the BVH trees keep
    the first hit in DFS order under `child.t < hit.t`; the restatement has to pick the same one everywhere."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw1_near_ties(spacing=spacing, width=128, height=96)
    want = list(cpu.run_reference(tree, xml)["images"].values())[0]
    hs = scene.HostScene(scene.parse_xml(xml, tree))
    rgbw, _, _, _ = cpu.render(hs, 0)
    assert image.compare_8bit(image.quantise(rgbw, hs.spec.semantics), want)["exact"] == 1.0


@pytest.mark.parametrize("tree", ["hw1", "hw2", "hw3", "hw4", "hw5", "hw6", "hw7"])
@pytest.mark.parametrize("seed,radius", [(0, 2.5), (1, 2.5), (2, 2.5), (0, 40.0), (1, 40.0), (3, 40.0)])
def test_oracle_fuzz_scenes_match_live_reference(tree, seed, radius):
    """Differential test on scenes built to be awkward (scenes.fuzz_scene: mirroring scales, transform chains,
    resetTransform on and off, zero-area triangles, a sphere around the camera, an axis-aligned centre ray, mirror and
    glass): PNG trees byte for byte, HW5 / HW6 float for float.  Radius 2.5 is the closed room (nearly every primary
    ray ends on the inside of the camera's sphere); radius 40 opens it, and 10-14 different objects, their shadows,
    mirror images and refractions fill the frame."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.fuzz_scene(seed, camera_sphere_radius=radius)
    res = cpu.run_reference(tree, xml)
    hs = scene.HostScene(scene.parse_xml(xml, tree))
    rgbw, _, _, _ = cpu.render(hs, 0)
    if res["images"]:
        cmp = image.compare_8bit(image.quantise(rgbw, hs.spec.semantics), list(res["images"].values())[0])
        assert cmp["within1"] == 1.0 and (tree == "hw4" or cmp["exact"] == 1.0), cmp
    else:
        want = list(res["hdr"].values())[0]
        got = rgbw[..., :3] / rgbw[..., 3:4]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("seed", [0, 3, 5])
def test_oracle_textured_scenes_match_live_reference_floats(tmp_path, seed):
    """Randomised variants of the textured / BRDF scene (image textures in every mode, image bump, five BRDFs, spot and
    directional lights, mirror, glass, rotated instances) against float dumps of four trees' own objects: HW5, HW6 and
    HW7 to the last bit, HW4 to the ulp its racy splat order allows."""
    for tree in ("hw4f", "hw5", "hw6", "hw7f"):
        if not cpu.have_ref(tree):
            pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=75, height=57, n=5 + seed % 5, depth=1 + seed % 4, seed=900 + seed,
                           light_scale=1.0 + seed % 3, mesh_slices=(6 + seed % 4, 7 + seed % 3), image_name="o.png")
    for tree in ("hw4f", "hw5", "hw6", "hw7f"):
        sem = tree.rstrip("f")
        want = list(cpu.run_reference(tree, xml, workdir=str(tmp_path))["hdr"].values())[0]
        hs = scene.HostScene(scene.parse_xml(xml, sem, base_dir=str(tmp_path)))
        rgbw, _, _, _ = cpu.render(hs, 0)
        got = rgbw[..., :3] / rgbw[..., 3:4]
        if sem == "hw4":
            got = np.clip(got, 0, 255)          # HW4's Pixel::get_color clamps
            assert np.max(np.abs(got - want) / (np.abs(want) + 1)) < 1e-6
        else:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), tree


def _replay(tree, xml, workdir):
    """Reference twin on ONE thread with std::random_device AND std::chrono::system_clock::now interposed to 795
    (REF_SEED) against the oracle replaying the same engines (oracle_hw7_mt_mode, oracle_clock_mode):
    returns (oracle floats, reference floats)."""
    import ctypes as C
    sem = tree.rstrip("f")
    os.environ["REF_SEED"] = "795"
    try:
        want = list(cpu.run_reference(tree, xml, workdir=workdir)["hdr"].values())[0]
    finally:
        del os.environ["REF_SEED"]
    hs = scene.HostScene(scene.parse_xml(xml, sem, base_dir=workdir))
    L = cpu.load()
    for f in (L.oracle_hw7_mt_mode, L.oracle_clock_mode):
        f.restype = None
        f.argtypes = [C.c_int, C.c_uint32]
    L.oracle_hw7_mt_mode(1, 795)      # engines seeded by random_device
    L.oracle_clock_mode(1, 795)       # engines seeded by the clock (HW3, HW4; glossy in HW5 / HW6)
    try:
        rgbw, _, _, _ = cpu.render(hs, 0)
    finally:
        L.oracle_hw7_mt_mode(0, 0)
        L.oracle_clock_mode(0, 0)
    got = rgbw[..., :3] / rgbw[..., 3:4]
    if sem in ("hw3", "hw4"):
        got = np.clip(got, 0, 255)    # their Pixel::get_color clamps (the HW2 twin dumps color / weight itself)
    return got, want


@pytest.mark.parametrize("tree,pt,spp", [("hw7f", False, 9), ("hw7f", True, 9), ("hw7f", True, 16), ("hw6", False, 9),
                                         ("hw5", False, 4), ("hw4f", False, 9)])
def test_oracle_replays_the_stochastic_paths_draw_for_draw(tmp_path, tree, pt, spp):
    """The stochastic rows (a2 jitter, a10 area / mesh / sphere light sampling, a11 glossy, a14 path tracing) pinned
    EXACTLY in the one setting where the reference is a function: one thread, known seeds.  Every engine of the tree
    is `mt19937 generator(rd())`; the float-dump driver defines random_device::_M_getval itself, and the oracle draws
    from the same per-site streams in the same program order (libstdc++'s uniform_real_distribution<float> restated).
    All floats of the frame are the same bits — light-mesh CDF pick, cone sampling with its re-intersection, the
    double-counted direct light of trace_path, the sigma-1/2 splat sums of HW5/HW6.  (It also showed that the tree, as
    built here, traces the TRANSMISSION child of a dielectric before the reflection child: unspecified in C++.)"""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    import re
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=48, height=36, n=6, spp=spp, depth=3, stochastic_lights=True,
                           path_tracing=pt, importance=True, light_scale=0.5, mesh_slices=(6, 7), image_name="o.png")
    if tree != "hw7f":
        xml = re.sub(r"    <Light(Mesh|Sphere).*?</Light(Mesh|Sphere)>\n", "", xml, flags=re.S)   # HW7-only objects
    got, want = _replay(tree, xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw3f", "hw4f", "hw5", "hw6", "hw7f"])
def test_oracle_replays_dof_and_motion_blur_draw_for_draw(tmp_path, tree):
    """cfg3's features (depth of field, motion blur, area lights, glossy mirror, glass, the Gaussian splat) through the
    same replay — including cfg3's own tree, HW3, whose engines are minstd_rand0 re-seeded from the clock at every
    pixel, area-light sample and glossy bounce (so, under a constant clock, the same few numbers over and over: a
    degenerate sampler, but the whole code path, and every float of the frame agrees)."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.cfg3_hw3(n=12, width=40, height=30, spp=9, depth=3, image_name="o.png")
    got, want = _replay(tree, xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw3f", "hw4f", "hw5", "hw6", "hw7f"])
@pytest.mark.parametrize("seed,radius", [(0, 2.5), (1, 2.5), (0, 40.0), (2, 40.0)])
def test_oracle_replays_stochastic_fuzz_scenes(tmp_path, tree, seed, radius):
    """The awkward scenes of scenes.fuzz_scene with an area light, a glossy mirror, depth of field, a motion-blurred
    instance and sphere, 4 samples per pixel: every float of the frame, on five trees, through the replay — the closed
    room (radius 2.5) and the open variant where the moving instance, the glossy mirror and the glass are in view."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.fuzz_scene(seed, width=41, height=33, spp=4, stochastic=True, camera_sphere_radius=radius)
    got, want = _replay(tree, xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw4f", "hw6", "hw7f"])
def test_oracle_perlin_textures_against_the_trees(tmp_path, tree):
    """Perlin noise as texture and as BUMP map at image level.  The trees shuffle the permutation table with an mt19937
    seeded from the clock, so no two runs of a binary agree — but with the clock interposed (REF_SEED) the table is
    std::shuffle(0..15, mt19937(795)), which oracle/cpu.py restates (libstdc++'s shuffle and uniform_int_distribution);
    the host uploads that table and the frame has to come out the same float for float: Perlin_noise::get_value_at at
    the object-space hit point, Texture::bump_normal's finite differences, blend_kd."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    import re
    from ceng795_b200 import scenes
    assert cpu.perlin_table_for_clock(795) == [0, 7, 12, 2, 15, 11, 9, 6, 4, 5, 1, 3, 13, 14, 8, 10]
    xml = scenes.hw7_scene(str(tmp_path), width=80, height=60, n=6, depth=2, perlin=True, mesh_slices=(6, 7),
                           image_name="o.png")
    table = " ".join(str(v) for v in cpu.perlin_table_for_clock(795))
    xml, n = re.subn(r"<Permutation>.*?</Permutation>", f"<Permutation>{table}</Permutation>", xml)
    assert n >= 1
    got, want = _replay(tree, xml, str(tmp_path))
    if tree == "hw4f":
        assert np.max(np.abs(got - want) / (np.abs(want) + 1)) < 1e-6      # its 1-spp splat sums in thread order... one thread here
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw6", "hw7f"])
@pytest.mark.parametrize("variant", ["left", "simple"])
def test_oracle_camera_variants_against_the_trees(tmp_path, tree, variant):
    """Row a1: the left-handed camera (an HW7 attribute; HW6 ignores it) and the `type="simple"` camera (GazePoint +
    FovY -> near plane through tanf, HW7/src/Scene.cpp:580-601) on the textured scene, float for float."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    import re
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=90, height=60, n=6, depth=2, mesh_slices=(6, 7), image_name="o.png")
    if variant == "left":
        xml = xml.replace('<Camera id="1">', '<Camera id="1" handedness="left">')
    else:
        xml = re.sub(r"      <Gaze>.*?</Gaze>\n", "      <GazePoint>0.2 0.6 -1.5</GazePoint>\n", xml)
        xml = re.sub(r"      <NearPlane>.*?</NearPlane>\n", "      <FovY>52</FovY>\n", xml)
        xml = xml.replace('<Camera id="1">', '<Camera id="1" type="simple">')
    got, want = _replay(tree, xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("pt,spp", [(False, 1), (True, 4)])
def test_oracle_environment_light_against_the_tree(tmp_path, pt, spp):
    """SphericalDirectionalLight (row a10): the tree reads its environment map with tinyexr, samples the hemisphere
    uniformly, looks the lat-long texel up with nearest sampling and divides by the pdf; rays that leave the scene
    pick the map up too.  Both sides get the same uncompressed float32 EXR (image.write_exr_f32); through the replay
    the frames agree float for float in ray-tracing and in path-tracing mode."""
    if not cpu.have_ref("hw7f"):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    rng = np.random.default_rng(3)
    env = (rng.uniform(0, 1, (16, 32, 3)) ** 3 * 40).astype(np.float32)
    image.write_exr_f32(str(tmp_path / "env.exr"), env)
    assert np.array_equal(image.read_exr_f32(str(tmp_path / "env.exr")), env)
    xml = scenes.hw7_scene(str(tmp_path), width=64, height=48, n=6, depth=2, mesh_slices=(6, 7), image_name="o.png",
                           path_tracing=pt, spp=spp)
    xml = xml.replace("  </Lights>", '    <SphericalDirectionalLight id="1"><EnvMapName>env.exr</EnvMapName>'
                                     "</SphericalDirectionalLight>\n  </Lights>")
    got, want = _replay("hw7f", xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw4f", "hw6", "hw7f"])
def test_oracle_background_texture_against_the_trees(tmp_path, tree):
    """BackgroundTexture (rows a1 / a9): primary rays that miss look the image up at (bg_u, bg_v) of the camera ray."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=72, height=54, n=6, depth=2, mesh_slices=(6, 7), image_name="o.png")
    xml = xml.replace("<Scene>\n", "<Scene>\n  <BackgroundTexture>tex_checker.png</BackgroundTexture>\n", 1)
    # open the room so that primary rays can leave: drop the stand-alone triangles of the back wall
    xml = xml.replace('<Gaze>0 -0.32 -1</Gaze>', '<Gaze>0 0.55 1</Gaze>')
    got, want = _replay(tree, xml, str(tmp_path))
    assert (np.abs(want - np.array([12, 16, 30], np.float32)) > 1e-3).any()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("pt", [False, True])
def test_oracle_replays_transformed_light_shapes(tmp_path, pt):
    """LightMesh and LightSphere under transformations (the light mesh keeps its OBJECT-space box in the tree's BVH,
    HW7/include/Light_mesh.h:52-68; the sphere light is re-intersected to get its distance) and path tracing with the
    UNIFORM hemisphere sample instead of importance sampling: float for float through the replay."""
    if not cpu.have_ref("hw7f"):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=64, height=48, n=6, spp=4, depth=3, stochastic_lights=True,
                           path_tracing=pt, importance=not pt, light_scale=0.5, mesh_slices=(6, 7), image_name="o.png")
    xml = xml.replace('<LightMesh id="1">\n      <Material>1</Material>',
                      '<LightMesh id="1">\n      <Material>1</Material>\n      <Transformations>r2 t3</Transformations>')
    xml = xml.replace('<LightSphere id="1">\n      <Material>1</Material>',
                      '<LightSphere id="1">\n      <Material>1</Material>\n      <Transformations>s2 t4</Transformations>')
    assert "r2 t3" in xml and "s2 t4" in xml
    got, want = _replay("hw7f", xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw5", "hw6", "hw7f"])
def test_oracle_second_spot_and_directional_light(tmp_path, tree):
    """A loader quirk that changes pixels: HW5 / HW6 walk SpotLight and DirectionalLight elements with
    FirstChildElement instead of NextSiblingElement and so keep only the FIRST of each (HW6/src/Scene.cpp:661,678);
    HW7 reads them all.  Two of each (and two point lights, read by every tree) — float for float per tree."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=64, height=48, n=6, depth=2, mesh_slices=(6, 7), image_name="o.png")
    extra = ('    <SpotLight id="2"><Position>-3 5 2</Position><Direction>0.4 -1 -0.3</Direction>'
             '<Intensity>2500 2000 1500</Intensity><CoverageAngle>40</CoverageAngle><FalloffAngle>20</FalloffAngle></SpotLight>\n'
             '    <DirectionalLight id="2"><Direction>-0.4 -1 0.2</Direction><Radiance>15 20 25</Radiance></DirectionalLight>\n'
             '    <PointLight id="2"><Position>2 6 -1</Position><Intensity>900 800 700</Intensity></PointLight>\n')
    xml = xml.replace("  </Lights>", extra + "  </Lights>")
    hs = scene.HostScene(scene.parse_xml(xml, tree.rstrip("f"), base_dir=str(tmp_path)))
    assert hs.desc.n_lights == (6 if tree == "hw7f" else 4)
    got, want = _replay(tree, xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw4f", "hw6", "hw7f"])
def test_oracle_transformed_textured_spheres(tmp_path, tree):
    """Spheres under scale / rotate / translate chains with an image texture, an image BUMP map and a Perlin bump map
    (uv from the object-space hit point, tangents of Sphere::intersect, the always-applied normal matrix), one of
    them motion-blurred (per-ray inverse), 4 samples per pixel — float for float through the replay."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    import re
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=80, height=60, n=6, depth=2, mesh_slices=(6, 7), image_name="o.png",
                           perlin=True)
    table = " ".join(str(v) for v in cpu.perlin_table_for_clock(795))
    xml = re.sub(r"<Permutation>.*?</Permutation>", f"<Permutation>{table}</Permutation>", xml)
    n0 = xml.count("<Transformations>")
    for sid, tr in ((1, "s2 r3 t2"), (4, "r1 s3"), (5, "s4 r2")):
        xml = re.sub(rf'(<Sphere id="{sid}">\n      <Material>\d+</Material>\n)',
                     rf'\1      <Transformations>{tr}</Transformations>\n', xml)
    assert xml.count("<Transformations>") == n0 + 3
    xml = re.sub(r'(<Sphere id="2">\n      <Material>\d+</Material>\n)', r'\1      <MotionBlur>0.2 0.1 0</MotionBlur>\n', xml)
    xml = xml.replace("<NumSamples>1</NumSamples>", "<NumSamples>4</NumSamples>")
    got, want = _replay(tree, xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


@pytest.mark.parametrize("tree", ["hw4f", "hw6", "hw7f"])
def test_oracle_moving_bumped_mesh_and_textured_instances(tmp_path, tree):
    """A motion-blurred, rotated mesh with an image BUMP map (dp/du, dp/dv from the texture coordinates, the normal
    matrix of the per-ray inverse) and a base mesh WITHOUT texture coordinates given an image texture, which its
    rotated / scaled instances inherit (constant uv: the degenerate-Jacobian guard) — float for float, 4 spp."""
    if not cpu.have_ref(tree):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=80, height=60, n=6, depth=2, mesh_slices=(6, 7), image_name="o.png", spp=4)
    a = '<Mesh id="1" shadingMode="smooth">\n      <Material>4</Material>\n      <Texture>1</Texture>\n'
    b = '<Mesh id="2" shadingMode="smooth">\n      <Material>5</Material>\n'
    assert a in xml and b in xml
    xml = xml.replace(a, '<Mesh id="1" shadingMode="smooth">\n      <Material>4</Material>\n      <Texture>3</Texture>\n'
                         '      <MotionBlur>0.15 0 0.1</MotionBlur>\n      <Transformations>r1 t1</Transformations>\n')
    xml = xml.replace(b, b + '      <Texture>1</Texture>\n')
    got, want = _replay(tree, xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


def test_oracle_replays_hw2_multisample_draw_for_draw(tmp_path):
    """HW2's multi-sample loop (one minstd_rand0 per pixel seeded from the clock, 3x3 Gaussian splat) through the float
    twin of the HW2 tree.  HW2's Pixel::get_color truncates to ints, so its PNGs hide the accumulators; the twin dumps
    color / weight itself.  This is where HW2's filter weight turned out to differ from HW3's by an ulp (its divisor
    2 * M_PI * sigma is not cast to float, HW2/Scene.cpp:12-14)."""
    if not cpu.have_ref("hw2f"):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.cfg2_hw3(n=16, width=48, height=36, n_instances=0, depth=3, glass=True, image_name="o.png")
    assert "<NumSamples>1</NumSamples>" in xml
    xml = xml.replace("<NumSamples>1</NumSamples>", "<NumSamples>9</NumSamples>")
    got, want = _replay("hw2f", xml, str(tmp_path))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.mean(got == want))


def test_parse_time():
    assert cpu.parse_time_ms("4 seconds 790 milliseconds 898 microseconds") == pytest.approx(4790.898)
    assert cpu.parse_time_ms("1 minute 2 seconds 0 milliseconds 5 microseconds") == pytest.approx(62000.005)


@pytest.mark.parametrize("pt", [False, True])
def test_oracle_hw7_stochastic_statistics_vs_reference(tmp_path, pt):
    """Stochastic HW7 (area / mesh / sphere lights, glossy mirror, path tracing): every reference RNG is seeded
    from random_device, so the oracle (counter-based RNG) is validated statistically against the untouched HW7
    binary at equal sample count: equal image mean (2 %) and an RMSE no larger than 1.35x the oracle's own
    seed-to-seed noise."""
    if not cpu.have_ref("hw7"):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=40, height=30, spp=144, depth=2, stochastic_lights=True,
                           path_tracing=pt, importance=True, light_scale=0.25)
    ref = list(cpu.run_reference("hw7", xml, workdir=str(tmp_path))["images"].values())[0].astype(np.float64)
    hs = scene.HostScene(scene.parse_xml(xml, "hw7", base_dir=str(tmp_path)))
    imgs = []
    for seed in (7, 8):
        rgbw, _, _, _ = cpu.render(hs, 0, seed=seed)
        imgs.append(image.quantise(rgbw, hs.spec.semantics).astype(np.float64))
    assert abs(imgs[0].mean() - ref.mean()) <= 0.02 * ref.mean(), (imgs[0].mean(), ref.mean())
    noise = np.sqrt(((imgs[0] - imgs[1]) ** 2).mean())
    rmse = np.sqrt(((imgs[0] - ref) ** 2).mean())
    assert rmse <= 1.35 * noise + 0.5, (rmse, noise)


def test_oracle_hw3_distribution_statistics_vs_reference():
    """cfg 3 (area lights, DOF, glossy mirror, motion blur, 3x3 Gaussian splat): clock-seeded in the reference,
    so validated statistically against the untouched HW3 binary at equal sample count."""
    if not cpu.have_ref("hw3"):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.cfg3_hw3(n=24, width=48, height=36, spp=64, depth=3)
    ref = list(cpu.run_reference("hw3", xml)["images"].values())[0].astype(np.float64)
    hs = scene.HostScene(scene.parse_xml(xml, "hw3"))
    q = []
    for seed in (1, 2):
        rgbw, _, _, _ = cpu.render(hs, 0, seed=seed)
        q.append(image.quantise(rgbw, hs.spec.semantics).astype(np.float64))
    assert abs(q[0].mean() - ref.mean()) <= 0.01 * ref.mean()
    noise = np.sqrt(((q[0] - q[1]) ** 2).mean())
    assert np.sqrt(((q[0] - ref) ** 2).mean()) <= 1.35 * noise + 0.5


def test_host_tonemap_matches_reference_png(tmp_path):
    """Photographic_tmo + sRGB (HW7/src/Photographic_tmo.cpp, main.cpp:59-91): the host restatement against the HW7
    binary's tonemapped PNG of a deterministic scene."""
    if not cpu.have_ref("hw7"):
        pytest.skip("oracle/_ref not built on this box")
    from ceng795_b200 import scenes
    xml = scenes.hw7_scene(str(tmp_path), width=96, height=72, light_scale=3.0, tonemap=True)
    want = list(cpu.run_reference("hw7", xml, workdir=str(tmp_path))["images"].values())[0]
    spec = scene.parse_xml(xml, "hw7", base_dir=str(tmp_path))
    rgbw, _, _, _ = cpu.render(scene.HostScene(spec), 0)
    tm = spec.cameras[0]["tonemap"]
    got = image.tonemap_photographic(rgbw, tm["key"], tm["burn"], tm["saturation"])
    assert image.compare_8bit(got, want)["exact"] == 1.0
