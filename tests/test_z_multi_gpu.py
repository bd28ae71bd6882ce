"""Multi-GPU entry points of the C ABI (include/b200rt.h "multi-GPU") on hardware.

The shared-frame protocol (export / attach / begin / render_shared / collect) and the PPM count synchronisation do not
care whether two shares come from two GPUs, two processes or two scenes on ONE GPU, so most of this file runs on a
single B200: two scene objects on device 0 play the two contributors, and two PROCESSES on device 0 exercise the cudaIpc
path.  Tests that need peer memory between different devices skip on a one-GPU box and run under `gpurun --gpus 2`."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from ceng795_b200 import _lib, scene, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHARE_KEYS = ("row_begin", "row_end", "row_stride", "row_group", "sample_begin", "sample_end")


@pytest.fixture(scope="module")
def gpu():
    L = _lib.load()
    if L.b200rt_device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200")
    return L


def _share(sc, part, parts):
    return {k: v for k, v in sc.multi_share(part, parts, 0).items() if k in SHARE_KEYS}


def _shared_render(contributors, seed, frames=2):
    """contributors[0] owns the frame; every contributor renders its share of every frame."""
    owner = contributors[0]
    h = owner.frame_export(0)
    for c in contributors:
        c.frame_attach(h)
    W, H = owner.image_size(0)
    out = np.zeros((H, W, 4), np.float32)
    rays = 0
    for _ in range(frames):
        owner.frame_begin()
        rays = 0
        for k, c in enumerate(contributors):
            rays += c.render_shared(0, seed=seed, **_share(c, k, len(contributors))).rays_total
        owner.frame_collect(len(contributors), out)
    for c in contributors:
        c.frame_detach()
    return out, rays


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [2, 3])
def test_shared_frame_of_a_one_sample_camera_is_bit_equal_to_one_render(gpu, parts):
    """cfg2's dialect (HW3, mirror depth 2, 1 spp): 8-row bands dealt round-robin, each share stored in place."""
    xml = scenes.cfg2_hw3(n=96, width=320, height=200)
    cs = [scene.Scene(xml, "hw3") for _ in range(parts)]
    assert _share(cs[1], 1, parts) == dict(row_begin=8, row_end=200, row_stride=8 * parts, row_group=8, sample_begin=0,
                                           sample_end=0)
    out, rays = _shared_render(cs, seed=3)
    ref, st = cs[0].render_image(0, seed=3)
    assert rays == st.rays_total
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.gpu
def test_shared_frame_of_splatting_cameras_adds_up(gpu, tmp_path):
    """Multi-sample cameras are split by sample strata and ADDED into the owner's frame (a row split would
    lose the 3x3 Gaussian splat across band edges); HW4 splats even at 1 spp and is split by rows, its halo rows added."""
    xml = scenes.cfg3_hw3(n=48, width=160, height=96, spp=16)
    cs = [scene.Scene(xml, "hw3") for _ in range(2)]
    assert _share(cs[1], 1, 2) == dict(row_begin=0, row_end=0, row_stride=1, row_group=0, sample_begin=8, sample_end=16)
    out, rays = _shared_render(cs, seed=9)
    ref, st = cs[0].render_image(0, seed=9)
    assert rays == st.rays_total
    assert np.allclose(out, ref, rtol=2e-5, atol=1e-6)          # float sums in another order, as the reference's threads
    xml4 = scenes.hw7_scene(str(tmp_path), width=96, height=72)
    c4 = [scene.Scene(xml4, "hw4", base_dir=str(tmp_path)) for _ in range(2)]
    assert _share(c4[1], 1, 2)["row_group"] == 8
    out4, rays4 = _shared_render(c4, seed=1)
    ref4, st4 = c4[0].render_image(0, seed=1)
    assert rays4 == st4.rays_total and np.allclose(out4, ref4, rtol=2e-5, atol=1e-6)
    assert (out4[..., 3] > 0).all()                              # every pixel, band edges included, got its weights


@pytest.mark.gpu
def test_shared_frame_through_the_hw7_kernels(gpu, tmp_path):
    """The HW5-HW7 combine kernel (k_combine7) behind the same protocol: a 1-sample Whitted frame with textures, BRDFs
    and every light class is bit-equal across three shares; a 16-spp path-traced frame (strata batches, box filter
    adds) agrees to float summation order."""
    xml = scenes.hw7_scene(str(tmp_path), width=96, height=72, light_scale=3.0)
    cs = [scene.Scene(xml, "hw7", base_dir=str(tmp_path)) for _ in range(3)]
    out, rays = _shared_render(cs, seed=5)
    ref, st = cs[0].render_image(0, seed=5)
    assert rays == st.rays_total and np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    xml_pt = scenes.hw7_scene(str(tmp_path), width=64, height=48, spp=16, depth=3, stochastic_lights=True,
                              path_tracing=True, importance=True, light_scale=2.0)
    cp = [scene.Scene(xml_pt, "hw7", base_dir=str(tmp_path)) for _ in range(2)]
    assert _share(cp[0], 0, 2)["sample_end"] == 8
    out, rays = _shared_render(cp, seed=6)
    ref, st = cp[0].render_image(0, seed=6)
    assert rays == st.rays_total
    fin = np.isfinite(ref).all(-1) & np.isfinite(out).all(-1)
    assert fin.mean() > 0.99 and np.allclose(out[fin], ref[fin], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_render_shared_without_begin_times_out_cleanly_is_not_needed_when_owner_opens(gpu):
    """Protocol errors are codes, not hangs: rendering into a frame that was never exported is refused."""
    sc = scene.Scene(scenes.cfg2_hw3(n=24, width=64, height=40), "hw3")
    with pytest.raises(_lib.B200RTError, match="frame_attach"):
        sc.render_shared(0)
    with pytest.raises(_lib.B200RTError, match="frame_export"):
        sc.frame_begin()


@pytest.mark.gpu
def test_ppm_hit_points_are_numbered_canonically_and_counts_synchronise(gpu):
    """Two scenes (two 'GPUs'): the same hit point has the same index in both — the eye pass appends with atomics, the
    canonical sort undoes that — and with every share claiming every second acceptance index and the counts summed after
    every pass, the two-share run follows the one-share run's radius schedule: accepted photons within 3 %, image mean
    within 2 %.  (Each share counting only its own photons between syncs accepts 7 % more: its radii lag.)"""
    import torch
    from ceng795_b200 import dist as rdist
    xml = scenes.cfg5_ppm(width=160, height=120, mesh_slices=(24, 22))
    a, b, one = (scene.Scene(xml, "ppm") for _ in range(3))
    a.ppm_set_share(0, 2); b.ppm_set_share(1, 2)   # each claims every second acceptance index between two syncs
    for s in (a, b, one):
        s.eye_trace_lines(0, seed=5)
    ha, hb = a.hit_points(), b.hit_points()
    assert len(ha) > 10_000 and np.array_equal(ha[:, :15].view(np.uint32), hb[:, :15].view(np.uint32))
    passes, photons = 4, 200_000
    for k in range(passes):
        a.trace_n_photons(photons // 2, first_photon=k * photons, seed=5)
        b.trace_n_photons(photons - photons // 2, first_photon=k * photons + photons // 2, seed=5)
        (pa, n), (pb, _) = a.ppm_counts_delta(), b.ppm_counts_delta()
        ta, tb = rdist.tensor_from_ptr(pa, (n,), 0, "<i4"), rdist.tensor_from_ptr(pb, (n,), 0, "<i4")
        tot = ta + tb                                    # what the all-reduce leaves on every rank
        ta.copy_(tot); tb.copy_(tot)
        torch.cuda.synchronize()
        a.ppm_counts_apply(); b.ppm_counts_apply()
        one.trace_n_photons(photons, first_photon=k * photons, seed=5)
    na, nb, n1 = a.hit_points()[:, 15], b.hit_points()[:, 15], one.hit_points()[:, 15]
    assert np.array_equal(na, nb)
    assert abs(na.sum() - n1.sum()) <= 0.03 * n1.sum(), (na.sum(), n1.sum())
    (sa, n), (sb, _) = a.ppm_export(), b.ppm_export()
    ta, tb = rdist.tensor_from_ptr(sa, (n, 4), 0), rdist.tensor_from_ptr(sb, (n, 4), 0)
    ta += tb
    torch.cuda.synchronize()
    img = a.density_estimation_from(sa, photons * passes)
    ref = one.density_estimation(photons * passes)
    ga, gb = img[..., :3] / np.maximum(img[..., 3:], 1e-30), ref[..., :3] / np.maximum(ref[..., 3:], 1e-30)
    assert abs(ga.mean() - gb.mean()) <= 0.02 * gb.mean(), (float(ga.mean()), float(gb.mean()))


def _torchrun(n, backend, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multi_rank_check.py"), "--backend", backend]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-1500:], r.stderr[-1500:])
    return json.loads(lines[-1])


@pytest.mark.gpu
def test_two_processes_share_a_frame_and_ppm_counts_over_cuda_ipc(gpu):
    """One process per 'GPU' as bench.py runs under torchrun: the frame handle crosses the process boundary
    (cudaIpcOpenMemHandle), shares are stored / added by the other process's kernels, arrival is awaited on the
    device.  On a one-GPU box both ranks use device 0 and gloo carries the handle; with >= 2 GPUs it is NCCL + NVLink."""
    two = gpu.b200rt_device_count() >= 2
    res = _torchrun(2, "nccl" if two else "gloo", 29600 + os.getpid() % 300)
    assert res["ok"] and res["hw3_1spp"]["bit_equal"] and res["hw3_16spp"]["close"] and res["ppm"]["same_numbering"], res
    print(res)


@pytest.mark.gpu
def test_render_multi_on_two_devices_equals_one_device(gpu):
    """b200rt_render_multi / b200rt_ppm_*_multi: ONE host thread, a scene per device, peer stores over NVLink."""
    if gpu.b200rt_device_count() < 2:
        pytest.skip("needs two devices (gpurun --gpus 2)")
    xml = scenes.cfg2_hw3(n=128, width=640, height=360)
    ms = scene.MultiScene(xml, "hw3", devices=[0, 1])
    out, st = ms.render_image(0, seed=4)
    ref, st1 = ms.parts[0].render_image(0, seed=4)
    assert st.rays_total == st1.rays_total and st.contexts == 2
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    out2, _ = ms.render_image(0, seed=4)                         # attached state is reused
    assert np.array_equal(out2.view(np.uint32), ref.view(np.uint32))
    ms.close()
    m3 = scene.MultiScene(scenes.cfg3_hw3(n=48, width=160, height=96, spp=16), "hw3", devices=[0, 1])
    o3, s3 = m3.render_image(0, seed=2)
    r3, s31 = m3.parts[0].render_image(0, seed=2)
    assert s3.rays_total == s31.rays_total and np.allclose(o3, r3, rtol=2e-5, atol=1e-6)
    m3.close()
    mp = scene.MultiScene(scenes.cfg5_ppm(width=160, height=120, mesh_slices=(24, 22)), "ppm", devices=[0, 1])
    img, s0, s1 = mp.render_ppm(0, photons=200_000, passes=4, seed=5)
    one = mp.parts[0]
    one.eye_trace_lines(0, seed=5)
    for k in range(4):
        one.trace_n_photons(200_000, first_photon=k * 200_000, seed=5)
    ref = one.density_estimation(800_000)
    ga, gb = img[..., :3] / np.maximum(img[..., 3:], 1e-30), ref[..., :3] / np.maximum(ref[..., 3:], 1e-30)
    assert abs(ga.mean() - gb.mean()) <= 0.02 * gb.mean(), (float(ga.mean()), float(gb.mean()))
    mp.close()
