"""Every BASELINE.json configuration at its FULL size on the device against the oracle (and, for config 4, against the
unmodified HW7 tree's own high-spp render), with the tolerance of north_star written into each assert:

  integer outputs  hit primitive ids equal on >= 99.99 % of the primary rays (here: all of them), t bit-equal
  deterministic    8-bit images within 1 LSB on >= 99.9 % of the pixels
  stochastic       matched counter-based sampling against the oracle on a row window (the oracle cannot render a
                   whole 100-spp / 4096-spp 1080p frame in test time), RMSE against the reference's own high-spp frame
                   within a stated bound, and a COUNTED decision-flip rate for the path tracer

The oracle runs on the GPU box's host cores; sizes are chosen so that each test's CPU side stays under ~30 s there.
"""
import json
import os

import numpy as np
import pytest

from ceng795_b200 import image, scene, scenes
from oracle import cpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu():
    from ceng795_b200 import _lib
    L = _lib.load()
    if L.b200rt_device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200")
    return L


def _rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))


# ------------------------------------------------------------------------------------------------ config 1
@pytest.mark.gpu
def test_cfg1_hw1_800x800_whole_frame_ids_and_image_match_the_oracle(gpu):
    """cfg1: 64 spheres + 2 meshes + 10 triangles, 3 point lights, mirror depth 6, 800 x 800.  The oracle tests every
    primitive like HW1 (no BVH); the device walks a SAH tree and replays HW1's order-dependent rule."""
    sc = scene.Scene(scenes.cfg1_hw1(), "hw1")
    assert sc.image_size(0) == (800, 800)
    o_rgbw, o_ids, o_t, cnt = cpu.render(sc.host, 0, want_ids=True)
    ids, t = sc.primary_ids(0)
    same = np.all(ids == o_ids, -1)
    assert same.mean() >= 0.9999, float(same.mean())                       # north_star: >= 99.99 % of primary rays
    assert np.array_equal(t[same].view(np.uint32), o_t[same].view(np.uint32))
    rgbw, st = sc.render_image(0)
    assert (st.rays_primary, st.rays_shadow, st.rays_secondary) == (cnt.rays_primary, cnt.rays_shadow, cnt.rays_secondary)
    cmp = image.compare_8bit(image.quantise(rgbw, sc.semantics), image.quantise(o_rgbw, sc.semantics))
    assert cmp["within1"] >= 0.999, cmp                                    # north_star: 1 LSB on >= 99.9 % of pixels
    print(f"cfg1 full size: ids equal {same.mean():.6f}, image {cmp}, rays {st.rays_total}")


# ------------------------------------------------------------------------------------------------ config 2
@pytest.mark.gpu
def test_cfg2_hw3_1080p_whole_frame_ids_and_image_match_the_oracle(gpu):
    """cfg2: 1 048 352-triangle heightfield + 8 instances + 4 spheres, 2 point lights, mirror depth 2, 1920 x 1080.
    Whole frame: every primary id, every t bit, every ray count, the 8-bit image."""
    sc = scene.Scene(scenes.cfg2_hw3(), "hw3")
    assert sc.image_size(0) == (1920, 1080) and sc.host.desc.n_tris >= 1_048_352
    o_rgbw, o_ids, o_t, cnt = cpu.render(sc.host, 0, want_ids=True)
    ids, t = sc.primary_ids(0)
    same = np.all(ids == o_ids, -1)
    assert same.mean() >= 0.9999, float(same.mean())
    assert np.array_equal(t[same].view(np.uint32), o_t[same].view(np.uint32))
    for k in (1, 3):                                                       # one launch train, and three row bands on three streams
        rgbw, st = sc.render_image(0, contexts=k)
        assert st.contexts == k
        assert (st.rays_primary, st.rays_shadow, st.rays_secondary) == (cnt.rays_primary, cnt.rays_shadow, cnt.rays_secondary)
        cmp = image.compare_8bit(image.quantise(rgbw, sc.semantics), image.quantise(o_rgbw, sc.semantics))
        assert cmp["within1"] >= 0.999, cmp
    print(f"cfg2 full size: ids equal {same.mean():.6f}, image {cmp}, rays {st.rays_total}")


# ------------------------------------------------------------------------------------------------ config 3
@pytest.mark.gpu
def test_cfg3_hw3_1080p_100spp_row_window_matches_the_oracle(gpu):
    """cfg3: 262 k triangles, 2 area lights, depth of field, glossy mirror, glass, a motion-blurred sphere, 1920 x 1080,
    100 samples per pixel (10 x 10 strata), 3 x 3 Gaussian splat.  Rows 536..544 of the full frame with the same
    counter-based samples on both sides (1.5 M camera paths): identical ray bookkeeping, colours to 1e-4 relative."""
    sc = scene.Scene(scenes.cfg3_hw3(), "hw3")
    assert sc.image_size(0) == (1920, 1080) and sc.cameras[0]["num_samples"] == 100
    rows = (536, 544)
    rgbw, st = sc.render_image(0, seed=77, row_begin=rows[0], row_end=rows[1])
    o_rgbw, _, _, cnt = cpu.render(sc.host, 0, rows=rows, seed=77)
    assert (st.rays_primary, st.rays_shadow, st.rays_secondary) == (cnt.rays_primary, cnt.rays_shadow, cnt.rays_secondary)
    assert st.rays_primary == 8 * 1920 * 100
    sl = slice(rows[0] - 1, rows[1] + 1)                                   # the splat reaches one row beyond the window
    assert not rgbw[: rows[0] - 1].any() and not rgbw[rows[1] + 1:].any()
    assert np.allclose(rgbw[sl, :, 3], o_rgbw[sl, :, 3], rtol=1e-5)        # sums of filter weights
    a = rgbw[sl, :, :3] / rgbw[sl, :, 3:]
    b = o_rgbw[sl, :, :3] / o_rgbw[sl, :, 3:]
    rel = np.abs(a - b) / (np.abs(b) + 1.0)
    assert rel.max() < 1e-4, float(rel.max())                              # tolerance: 1e-4 relative (+1 absolute floor)
    inner = slice(rows[0] + 1, rows[1] - 1)                                # rows that received all of their samples
    cmp = image.compare_8bit(image.quantise(rgbw[inner], sc.semantics), image.quantise(o_rgbw[inner], sc.semantics))
    assert cmp["within1"] >= 0.999, cmp
    print(f"cfg3 window: max rel {rel.max():.2e}, image {cmp}, rays {st.rays_total}")


# ------------------------------------------------------------------------------------------------ config 4
def _cfg4(tmp, **kw):
    info = json.load(open(os.path.join(GOLDEN, "cfg4_ref.json")))
    xml = scenes.cfg4_hw7(str(tmp), **kw)
    table = " ".join(str(v) for v in info["perlin_table"])   # the table the HW7 tree shuffles with its clock interposed
    a, b = xml.index("<Permutation>"), xml.index("</Permutation>")
    return scene.Scene(xml[:a] + "<Permutation>" + table + xml[b:], "hw7", base_dir=str(tmp)), info


@pytest.mark.gpu
def test_cfg4_hw7_1080p_4096spp_row_matches_the_oracle_with_matched_sampling(gpu, tmp_path):
    """cfg4 exactly as SURVEY 8(d) words it (Cornell box, LightMesh + LightSphere, 3 instances of a 100 k-triangle
    textured mesh, Perlin bump, 5 BRDFs, mirror, glass; path tracing + importance sampling, depth 4), 1920 x 1080,
    NumSamples 4096 = 64 x 64 strata.  One row of the full frame — 7.9 M camera paths — with the same samples on both
    sides.  Per-sample decisions can differ only where the reference itself is a rounding-level coin flip (its light
    shapes sit at t = dist - eps of their own shadow ray, HW7/src/Scene.cpp:283-290, and CUDA's sinf / cosf / asinf
    differ from glibc's by an ulp in a bounce direction); at 4096 samples those flips average out:
      * primary ray count equal, total ray count within 0.2 %
      * per-pixel relative RMSE <= 2 % (the Monte-Carlo noise of this frame at 4096 spp is ~6 %, cfg4_ref.json)
      * row mean within 0.5 %."""
    sc, info = _cfg4(tmp_path)
    assert sc.image_size(0) == (1920, 1080) and sc.cameras[0]["num_samples"] == 4096
    row = 700
    rgbw, st = sc.render_image(0, seed=4, row_begin=row, row_end=row + 1)
    o_rgbw, _, _, cnt = cpu.render(sc.host, 0, rows=(row, row + 1), seed=4)
    assert st.rays_primary == cnt.rays_primary == 1920 * 4096
    assert abs(st.rays_total - cnt.rays_total) <= 0.002 * cnt.rays_total, (st.rays_total, cnt.rays_total)
    a = rgbw[row, :, :3] / rgbw[row, :, 3:]
    b = o_rgbw[row, :, :3] / o_rgbw[row, :, 3:]
    fin = np.isfinite(a).all(-1) & np.isfinite(b).all(-1)
    assert fin.mean() > 0.999
    rr = _rel_rmse(a[fin], b[fin])
    assert rr <= 0.02, rr
    assert abs(a[fin].mean() - b[fin].mean()) <= 0.005 * b[fin].mean(), (float(a[fin].mean()), float(b[fin].mean()))
    print(f"cfg4 row {row} at 4096 spp: rel RMSE {rr:.4f}, mean {a[fin].mean():.4f} vs {b[fin].mean():.4f}, rays {st.rays_total}")


@pytest.mark.gpu
def test_cfg4_hw7_against_the_reference_trees_own_1024spp_frame(gpu, tmp_path):
    """SURVEY 8(c) plan 2(i): the cfg4 scene at 240 x 135 and 1024 spp rendered by the UNMODIFIED HW7 tree
    (tests/golden/make_cfg4_reference.py -> cfg4_ref_240x135_1024spp.npy; its random_device / clock interposed, so the
    file is reproducible) against the device at the same size and sample count with its own counter-based samples.
    Two independent 1024-spp estimates of one image: their RMSE is the Monte-Carlo noise, which the reference tree
    measured against ITSELF (a second run, other seed) is `ref_vs_ref_*` in cfg4_ref.json.  Stated bounds:
      * per-pixel relative RMSE (clipped at the reference's 99.5th percentile: fireflies) <= 1.15 x ref-vs-ref
      * 15 x 15 block means (noise / 15): relative RMSE <= 1.5 x ref-vs-ref, i.e. no regional bias above ~2 %
      * image mean within 1 %."""
    sc, info = _cfg4(tmp_path, width=info_w(), height=info_h(), spp=1024)
    ref = np.load(os.path.join(GOLDEN, "cfg4_ref_240x135_1024spp.npy")).astype(np.float32)
    assert ref.shape == (info["height"], info["width"], 3) and "second_run_block_means" in info
    rgbw, st = sc.render_image(0, seed=2026)
    a = rgbw[..., :3] / np.maximum(rgbw[..., 3:], 1)
    fin = np.isfinite(a).all(-1) & np.isfinite(ref).all(-1)
    assert fin.mean() > 0.999
    cap = info["clip_at_p99_5"]
    ac, rc = np.clip(np.where(fin[..., None], a, 0), 0, cap), np.clip(np.where(fin[..., None], ref, 0), 0, cap)
    rr = _rel_rmse(ac, rc)
    assert rr <= 1.15 * info["ref_vs_ref_rel_rmse_clipped"], (rr, info["ref_vs_ref_rel_rmse_clipped"])
    blk = lambda x: x.reshape(info["height"] // 15, 15, info["width"] // 15, 15, 3).mean((1, 3))
    br = _rel_rmse(blk(ac), blk(rc))
    assert br <= 1.5 * info["ref_vs_ref_block_rel_rmse"], (br, info["ref_vs_ref_block_rel_rmse"])
    assert abs(ac.mean() - rc.mean()) <= 0.01 * rc.mean(), (float(ac.mean()), float(rc.mean()))
    print(f"cfg4 vs HW7 tree at 1024 spp: rel RMSE {rr:.4f} (ref-vs-ref {info['ref_vs_ref_rel_rmse_clipped']:.4f}), "
          f"block {br:.4f} (ref-vs-ref {info['ref_vs_ref_block_rel_rmse']:.4f}), mean {ac.mean():.4f} vs {rc.mean():.4f}")


def info_w():
    return json.load(open(os.path.join(GOLDEN, "cfg4_ref.json")))["width"]


def info_h():
    return json.load(open(os.path.join(GOLDEN, "cfg4_ref.json")))["height"]


@pytest.mark.gpu
@pytest.mark.parametrize("importance", [True, False])
def test_path_tracer_decisions_differ_from_the_oracle_only_on_light_shapes(gpu, tmp_path, importance):
    """Row a14's bar, counted.  Both sides count, per pixel and light, the shadow tests that came out LIT
    (b200rt_debug_lit_counts / oracle_set_lit_counts) on a path-traced frame with every light class of HW7 in it.
    With matched samples the two tables must be EQUAL for point, area, spot and directional lights (<= 1e-3 of the
    entries and of the lit tests may differ: an ulp of libm in a bounce direction moving a vertex), and may differ
    only for the LightMesh / LightSphere, whose own surface lies at t = dist - eps of their shadow ray up to rounding
    (HW7/src/Scene.cpp:283-290): there the flip rate is counted and bounded, and because a flip is a fair coin the image
    stays unbiased: per-pixel RMSE and mean are bounded too."""
    from ceng795_b200 import _lib
    xml = scenes.hw7_scene(str(tmp_path), width=96, height=72, spp=16, depth=3, stochastic_lights=True, perlin=True,
                           path_tracing=True, importance=importance, light_scale=2.0)
    sc = scene.Scene(xml, "hw7", base_dir=str(tmp_path))
    sc.count_lit(True)
    rgbw, st = sc.render_image(0, seed=1234)
    lit = sc.lit_counts(0)
    o_lit = np.zeros_like(lit)
    o_rgbw, _, _, cnt = cpu.render(sc.host, 0, seed=1234, lit_counts=o_lit)
    kinds = [sc.host.desc.lights[i].kind for i in range(sc.host.desc.n_lights)]
    shapes = [i for i, k in enumerate(kinds) if k in (_lib.LIGHT_MESH, _lib.LIGHT_SPHERE)]
    others = [i for i, k in enumerate(kinds) if k not in (_lib.LIGHT_MESH, _lib.LIGHT_SPHERE)]
    assert len(shapes) == 2 and len(others) >= 4
    assert st.rays_primary == cnt.rays_primary and abs(st.rays_total - cnt.rays_total) <= 0.002 * cnt.rays_total
    diff_other = float(np.mean(lit[..., others] != o_lit[..., others]))
    diff_other_tests = float(np.abs(lit[..., others].astype(np.int64) - o_lit[..., others]).sum() / o_lit[..., others].sum())
    flips = np.abs(lit[..., shapes].astype(np.int64) - o_lit[..., shapes]).sum(axis=(0, 1))
    flip_rate = flips / np.maximum(1.0, o_lit[..., shapes].sum(axis=(0, 1)))
    net = (lit[..., shapes].sum(axis=(0, 1)) - o_lit[..., shapes].sum(axis=(0, 1))) / np.maximum(1, o_lit[..., shapes].sum(axis=(0, 1)))
    a = rgbw[..., :3] / np.maximum(rgbw[..., 3:], 1)
    b = o_rgbw[..., :3] / np.maximum(o_rgbw[..., 3:], 1)
    rr = _rel_rmse(a, b)
    print(f"path tracer decisions (importance={importance}): other lights differ in {diff_other:.2e} of entries / "
          f"{diff_other_tests:.2e} of lit tests; light-shape flip rate {flip_rate}, net {net}; rel RMSE {rr:.4f}; "
          f"mean {a.mean():.5f} vs {b.mean():.5f}; rays {st.rays_total} vs {cnt.rays_total}")
    # point / area / spot / directional lights: a decision differs only after an ulp of libm in a bounce direction moved a
    # vertex (then the whole subtree below it differs).  Measured 1.1e-4 .. 4.0e-4 of the pixel x light entries.
    assert diff_other <= 1e-3, diff_other
    assert diff_other_tests <= 1e-3, diff_other_tests
    # light shapes: the flip needs an ulp of difference upstream, i.e. a bounce vertex, and is then a fair coin
    assert (flip_rate <= 0.10).all(), flip_rate
    assert (np.abs(net) <= 0.01).all(), net                                         # the coin is fair: no net gain or loss of light
    assert rr <= 0.05, rr                                                           # per-pixel RMSE / mean at 16 spp
    assert abs(a.mean() - b.mean()) <= 0.005 * b.mean(), (float(a.mean()), float(b.mean()))


# ------------------------------------------------------------------------------------------------ config 5
@pytest.mark.gpu
def test_cfg5_ppm_1080p_eye_pass_and_a_million_photons_match_the_oracle(gpu):
    """cfg5: closed box + mirror sphere + glass sphere + 50 k-triangle mesh, 1 point light, depth 8, 1920 x 1080 eye
    pass.  Eye pass + hash grid: hit-point count, initial radius and every hit point of a 64-row window bit for bit.
    Photon pass: 1 000 000 photons of the same counter-based stream on both sides — photon rays equal to 0.1 %, cell
    scans to 0.2 %, accepted photons to 2 % (the order hit points shrink in is free, in the reference too), image mean
    to 2 %, and 15 x 15 block means of the density estimate within 1.25 x what the oracle itself differs by when it traces the same
    photons in the opposite order."""
    sc = scene.Scene(scenes.cfg5_ppm(), "ppm")
    assert sc.image_size(0) == (1920, 1080)
    n_photons = 1_000_000
    rgbw, s0, s1 = sc.render_ppm(0, photons=n_photons, seed=5)
    o_rgbw, o_hp, o_r0, cnt = cpu.ppm(sc.host, 0, seed=5, n_photons=n_photons)
    assert s0.hit_points == len(o_hp) > 2_000_000
    assert np.float32(s0.initial_radius) == np.float32(o_r0)
    hp = sc.hit_points()
    win = lambda h: h[(h[:, 13].astype(np.int64) // 1920 >= 500) & (h[:, 13].astype(np.int64) // 1920 < 564)]
    key = lambda h: h[np.lexsort([h[:, c] for c in (14, 2, 1, 0, 13)])]
    a, b = key(win(hp)), key(win(o_hp))
    assert a.shape == b.shape and len(a) > 100_000
    assert np.array_equal(a[:, :15].view(np.uint32), b[:, :15].view(np.uint32))
    assert abs(int(s1.photon_rays) - int(cnt.rays_shadow)) <= 0.001 * cnt.rays_shadow
    assert abs(int(s1.cell_scans) - int(cnt.spheres_tested)) <= 0.002 * cnt.spheres_tested
    assert abs(int(s1.deposits) - int(cnt.instances_entered)) <= 0.02 * cnt.instances_entered
    ga, gb = rgbw[..., :3] / np.maximum(rgbw[..., 3:], 1e-30), o_rgbw[..., :3] / np.maximum(o_rgbw[..., 3:], 1e-30)
    assert abs(ga.mean() - gb.mean()) <= 0.02 * gb.mean(), (float(ga.mean()), float(gb.mean()))
    # Yardstick for the density estimate: the oracle on the SAME photons traced last to first.  At ~8 photons per hit
    # point the radius a photon is tested against depends strongly on its rank; the reference's threads leave that
    # order free.  Measured here: order-vs-order 0.038, device-vs-oracle 0.037, another seed 0.355 (block rel RMSE).
    r_rgbw, _, _, _ = cpu.ppm(sc.host, 0, seed=5, n_photons=n_photons, want_hit_points=False, reverse_order=True)
    gr = r_rgbw[..., :3] / np.maximum(r_rgbw[..., 3:], 1e-30)
    blk = lambda x: x.reshape(72, 15, 128, 15, 3).mean((1, 3))
    br, yard = _rel_rmse(blk(ga), blk(gb)), _rel_rmse(blk(gr), blk(gb))
    print(f"cfg5 full size: {s0.hit_points} hit points, photon rays {s1.photon_rays} vs {cnt.rays_shadow}, deposits "
          f"{s1.deposits} vs {cnt.instances_entered}, block rel RMSE {br:.4f} (same photons in reverse order: {yard:.4f})")
    assert br <= 1.25 * yard, (br, yard)
