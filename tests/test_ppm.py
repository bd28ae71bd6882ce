"""PPM: the oracle against the untouched PPM binary (CPU, statistical — the
reference seeds from random_device and interleaves threads), and the CUDA path against the oracle (GPU):
eye pass and hash grid bit-exact, photon pass statistical (the per-photon radius/flux update is
order-dependent in the reference itself)."""
import os

import numpy as np
import pytest

from ceng795_b200 import image, scene, scenes
from oracle import cpu


def _scene_xml(**kw):
    args = dict(width=96, height=72, mesh_slices=(16, 14), photons_per_iteration=400, iterations=500, depth=8)
    args.update(kw)
    return scenes.cfg5_ppm(**args)


def test_oracle_ppm_statistics_vs_reference():
    if not cpu.have_ref("ppm"):
        pytest.skip("oracle/_ref not built on this box")
    xml = _scene_xml()
    ref = cpu.run_reference("ppm", xml)
    T = ref["threads"]
    img = list(ref["images"].values())[0].astype(np.float64)
    hs = scene.HostScene(scene.parse_xml(xml, "ppm"))
    per_thread = 400 // T
    n_photons = per_thread * T * 500                  # what trace_n_photons actually emits (PPM/src/main.cpp:74-88)
    total = 400 * per_thread * T                      # what main passes to density_estimation (:97-98, Appendix B.7)
    q = []
    for seed in (3, 4):
        rgbw, hp, r0, cnt = cpu.ppm(hs, 0, seed=seed, n_photons=n_photons, total_num_of_photons=total)
        q.append(image.quantise(rgbw, hs.spec.semantics).astype(np.float64))
    assert abs(q[0].mean() - img.mean()) <= 0.03 * img.mean(), (q[0].mean(), img.mean())
    noise = np.sqrt(((q[0] - q[1]) ** 2).mean())
    rmse = np.sqrt(((q[0] - img) ** 2).mean())
    assert rmse <= 1.35 * noise + 0.5, (rmse, noise)


def test_hash_wraps_like_the_reference():
    # (ix*73856093) ^ (iy*19349663) ^ (iz*83492791) overflows int for ix > 29: uint32 wraparound must match
    ix, iy, iz = 1234, 777, 4321
    want = ((ix * 73856093) ^ (iy * 19349663) ^ (iz * 83492791)) & 0xffffffff
    got = (np.uint32(ix) * np.uint32(73856093)) ^ (np.uint32(iy) * np.uint32(19349663)) ^ (np.uint32(iz) * np.uint32(83492791))
    assert int(got) == want


def _sorted_hp(hp):
    key = np.lexsort((hp[:, 2], hp[:, 1], hp[:, 0], hp[:, 13]))
    return hp[key]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(spp=4, width=64, height=48, depth=6)])
def test_eye_pass_and_grid_bit_exact(gpu, kw):
    xml = _scene_xml(**kw)
    sc = scene.Scene(xml, "ppm")
    st = sc.eye_trace_lines(0, seed=11)
    _, o_hp, o_r0, cnt = cpu.ppm(sc.host, 0, seed=11, n_photons=0)
    hp = sc.hit_points()
    assert st.hit_points == len(o_hp) > 0
    assert np.float32(st.initial_radius) == np.float32(o_r0)
    a, b = _sorted_hp(hp), _sorted_hp(o_hp)
    # position, normal, w_o, attenuation, material, pixel, weight: the same floats bit for bit
    assert np.array_equal(a[:, :15].view(np.uint32), b[:, :15].view(np.uint32))


@pytest.mark.gpu
def test_photon_pass_matches_oracle_statistically(gpu):
    xml = _scene_xml()
    sc = scene.Scene(xml, "ppm")
    n_photons = 200_000
    rgbw, s0, s1 = sc.render_ppm(0, photons=n_photons, seed=5)
    q = [image.quantise(rgbw * np.array([40, 40, 40, 1], np.float32), sc.semantics).astype(np.float64)]
    for seed in (5, 6):
        o, _, _, cnt = cpu.ppm(sc.host, 0, seed=seed, n_photons=n_photons, want_hit_points=False)
        q.append(image.quantise(o * np.array([40, 40, 40, 1], np.float32), sc.semantics).astype(np.float64))
        if seed == 5:
            # photon paths depend on the RNG stream and the geometry only: same rays (up to libm ulps in a bounce)
            assert abs(int(s1.photon_rays) - int(cnt.rays_shadow)) <= 0.001 * cnt.rays_shadow
            assert abs(int(s1.cell_scans) - int(cnt.spheres_tested)) <= 0.002 * cnt.spheres_tested
            # accepted photons depend on the order hit points shrink in (racy in the reference too)
            assert abs(int(s1.deposits) - int(cnt.instances_entered)) <= 0.02 * cnt.instances_entered
    assert abs(q[0].mean() - q[1].mean()) <= 0.02 * q[1].mean(), (q[0].mean(), q[1].mean())
    noise = np.sqrt(((q[1] - q[2]) ** 2).mean())
    rmse = np.sqrt(((q[0] - q[1]) ** 2).mean())   # same photon stream, different update order
    assert rmse <= 1.0 * noise + 0.5, (rmse, noise)
    # the additive export (what the multi-GPU all-reduce sums) reproduces density_estimation exactly
    p, n = sc.ppm_export()
    again = sc.density_estimation_from(p, n_photons)
    assert n == s0.hit_points and np.array_equal(again.view(np.uint32), rgbw.view(np.uint32))


@pytest.mark.gpu
def test_photon_calls_accumulate(gpu):
    """trace_n_photons is called once per thread in the reference; calls with disjoint photon ranges add up."""
    xml = _scene_xml()
    sc = scene.Scene(xml, "ppm")
    sc.eye_trace_lines(0, seed=1)
    a = sc.trace_n_photons(60_000, first_photon=0, seed=9)
    b = sc.trace_n_photons(40_000, first_photon=60_000, seed=9)
    two = sc.density_estimation(100_000)
    sc.eye_trace_lines(0, seed=1)
    c = sc.trace_n_photons(100_000, first_photon=0, seed=9)
    one = sc.density_estimation(100_000)
    assert a.photon_rays + b.photon_rays == c.photon_rays      # identical photon paths
    assert abs(float(two[..., :3].mean()) - float(one[..., :3].mean())) <= 0.02 * float(one[..., :3].mean())


@pytest.mark.parametrize("kw", [dict(width=80, height=60, depth=6), dict(width=57, height=91, depth=3, mesh_slices=(9, 7)),
                                dict(width=120, height=40, depth=9, mesh_slices=(24, 20))])
def test_oracle_eye_pass_and_hash_grid_match_the_ppm_tree(tmp_path, kw):
    """Rows a16 / a17 pinned directly: the PPM tree's own eye_trace_lines + build_hash_grid (oracle/ref_driver_ppm.cpp,
    linked with the tree's unmodified objects) against the oracle — every hit point float for float (position, normal,
    w_o, attenuation, pixel, weight), the initial radius, hash scale, bounding-box corner, and the size of EVERY cell
    of the hash grid (the signed-overflow hash, the abs(int()) cell coordinates, the +-r0 box insertion)."""
    import ctypes as C
    import subprocess
    exe = os.path.join(os.path.dirname(cpu.__file__), "_ref", "ppm_eye")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built on this box")
    xml = _scene_xml(**kw)
    (tmp_path / "s.xml").write_text(xml)
    subprocess.run([exe, str(tmp_path / "s.xml"), str(tmp_path / "o.bin")], check=True, capture_output=True,
                   stdin=subprocess.DEVNULL, cwd=str(tmp_path))
    blob = open(tmp_path / "o.bin", "rb").read()
    n_hp, num_hash = np.frombuffer(blob, np.int32, 2)
    hdr = np.frombuffer(blob, np.float32, 5, 8)
    ref_hp = np.frombuffer(blob, np.float32, n_hp * 16, 28).reshape(n_hp, 16)
    ref_cells = np.frombuffer(blob, np.int32, num_hash, 28 + n_hp * 64)
    hs = scene.HostScene(scene.parse_xml(xml, "ppm"))
    _, o_hp, o_r0, _ = cpu.ppm(hs, 0, seed=1, n_photons=0)
    assert len(o_hp) == n_hp > 500 and num_hash == n_hp
    # the reference appends hit points in traversal order of ONE thread here; so does the oracle: compare row by row
    cols = list(range(12)) + [13, 14]          # position, normal, w_o, attenuation, pixel, pixel_weight
    assert np.array_equal(o_hp[:, cols].view(np.uint32), ref_hp[:, cols].view(np.uint32))
    L = cpu.load()
    L.oracle_ppm_grid.restype = C.c_int
    L.oracle_ppm_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    out5 = np.zeros(5, np.float32)
    cells = np.zeros(num_hash, np.int32)
    nh = C.c_int64()
    assert L.oracle_ppm_grid(C.cast(hs.desc_ptr, C.c_void_p), 0, out5.ctypes.data, cells.ctypes.data, int(num_hash),
                             C.byref(nh)) == 0
    assert nh.value == num_hash
    assert np.array_equal(out5.view(np.uint32), hdr.view(np.uint32)), (out5, hdr)
    assert np.float32(o_r0) * np.float32(o_r0) == hdr[0]
    assert np.array_equal(cells, ref_cells), int((cells != ref_cells).sum())
    assert ref_cells.max() > 8 and (ref_cells == 0).any()


@pytest.mark.parametrize("n_photons,spp", [(50_000, 1), (20_000, 4)])
def test_oracle_photon_pass_replays_the_ppm_tree_draw_for_draw(tmp_path, n_photons, spp):
    """Rows a18-a20 pinned EXACTLY.  The PPM tree draws from three mt19937 engines seeded by std::random_device;
    oracle/ref_driver_ppm.cpp defines random_device::_M_getval itself (constant seed) and runs the tree's own
    trace_n_photons + density_estimation on one thread.  With `oracle_ppm_mt_mode` the restatement draws from the same
    three streams (libstdc++'s uniform_real_distribution<float> restated) in the same program order: after 50 000
    photons every hit point's photon count, radius^2 and flux and every pixel of the image are the same bits — the
    generate_photon quirk, the cell lookup, the radius / flux update, the cosine sample used as survival probability,
    the Fresnel roulette, the density estimate."""
    import ctypes as C
    import subprocess
    exe = os.path.join(os.path.dirname(cpu.__file__), "_ref", "ppm_eye")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built on this box")
    w, h = 80, 60
    xml = _scene_xml(width=w, height=h, depth=6, spp=spp)     # spp = 4: the jittered 2 x 2 eye pass is replayed too
    (tmp_path / "s.xml").write_text(xml)
    subprocess.run([exe, str(tmp_path / "s.xml"), str(tmp_path / "o.bin"), str(n_photons)], check=True,
                   capture_output=True, stdin=subprocess.DEVNULL, cwd=str(tmp_path))
    blob = open(tmp_path / "o.bin", "rb").read()
    n_hp, num_hash = np.frombuffer(blob, np.int32, 2)
    off = 28 + n_hp * 64 + num_hash * 4
    ref_state = np.frombuffer(blob, np.float32, n_hp * 5, off).reshape(n_hp, 5)
    ref_img = np.frombuffer(blob, np.float32, w * h * 3, off + n_hp * 20).reshape(h, w, 3)
    hs = scene.HostScene(scene.parse_xml(xml, "ppm"))
    L = cpu.load()
    L.oracle_ppm_mt_mode.restype = None
    L.oracle_ppm_mt_mode.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_int64]
    state = np.zeros((n_hp, 5), np.float32)
    L.oracle_ppm_mt_mode(1, 795, state.ctypes.data, int(n_hp))
    try:
        rgbw, _, _, cnt = cpu.ppm(hs, 0, seed=1, n_photons=n_photons, total_num_of_photons=n_photons)
    finally:
        L.oracle_ppm_mt_mode(0, 0, None, 0)
    assert ref_state[:, 4].sum() > 100_000 and ref_state[:, 4].max() > 20          # a real workout
    assert np.array_equal(state[:, 4], ref_state[:, 4])                              # photon counts n
    assert np.array_equal(state[:, 3].view(np.uint32), ref_state[:, 3].view(np.uint32))     # radius^2
    assert np.array_equal(state[:, :3].view(np.uint32), ref_state[:, :3].view(np.uint32))   # flux
    img = np.where(rgbw[..., 3:4] == 0, np.float32(0), rgbw[..., :3] / np.maximum(rgbw[..., 3:4], np.float32(1e-30)))
    assert np.array_equal(img.astype(np.float32).view(np.uint32), ref_img.view(np.uint32))
    assert int(cnt.instances_entered) == int(ref_state[:, 4].sum())
