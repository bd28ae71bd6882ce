"""Regenerates tests/golden/*.png by running the UNTOUCHED reference binaries
(oracle/_ref/, built by oracle/build_ref.sh from /root/reference) on the seeded
scenes below.  Run in the build container; the fixtures are committed so that
the GPU box (no /root/reference) can still pin both the oracle and the CUDA path.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ceng795_b200 import image, scenes  # noqa: E402
from oracle import cpu  # noqa: E402

#: name -> (tree, generator, kwargs)
CASES = {
    "cfg1_hw1_96": ("hw1", "cfg1_hw1", dict(width=96, height=96)),
    "cfg1_hw1_depth0": ("hw1", "cfg1_hw1", dict(width=64, height=48, depth=0, n_spheres=12)),
    # HW1's `t < best - 1e-6` makes the ORDER of near-tied primitives matter: stacks of coplanar quads 6e-7 apart,
    # duplicates, tangent spheres (the device replays the order on the near-tie set, csrc/rt_hw1_fast.cuh)
    "hw1_near_ties": ("hw1", "hw1_near_ties", dict(width=96, height=72)),
    "cfg2_hw3_n32": ("hw3", "cfg2_hw3", dict(n=32, width=160, height=90)),
    "cfg2_hw3_n48_smooth": ("hw3", "cfg2_hw3", dict(n=48, width=128, height=96, smooth=True, depth=3)),
    # HW2 tree: no instancing (MeshInstance elements are ignored by its loader), Mesh sits in the top BVH itself,
    # dielectric with the flipped Beer sign (HW2/Scene.cpp:166-168)
    "cfg2_hw2_glass": ("hw2", "cfg2_hw3", dict(n=32, width=128, height=96, n_instances=0, depth=4, glass=True)),
    "cfg2_hw3_glass": ("hw3", "cfg2_hw3", dict(n=32, width=128, height=96, n_instances=3, depth=4, glass=True)),
    # HW4 tree: HW3's shading loop + textures / bump (band triangle test, no culling), 3x3 Gaussian splat even at
    # 1 spp.  The splat makes the reference's float sums depend on its thread interleaving: 1-LSB tolerance.
    "hw4_textured": ("hw4", "hw7_scene", dict(width=128, height=96, light_scale=3.0, depth=2)),
    # HW7 Whitted, every pixel deterministic: textures (nearest/bilinear, repeat/clamp, replace_kd/blend_kd,
    # image bump), 5 BRDFs, point/spot/directional lights, mirror + glass, culling, instancing
    "hw7_whitted_160": ("hw7", "hw7_scene", dict(width=160, height=120, light_scale=3.0)),
    "hw7_whitted_depth1": ("hw7", "hw7_scene", dict(width=96, height=80, depth=1, n=8, light_scale=2.0, seed=801)),
}

#: PNGs of the unmodified HW3 binary that the DROP-IN binary (integration/main_hw3_b200.cpp: the reference's host with
#: the render fan-out replaced by libb200rt.so) has to reproduce — tests/test_z_dropin_binary.py.  Kept apart from CASES:
#: those parametrise the early GPU parity tests, these were added after the round's GPU budget was spent.
DROPIN_CASES = {
    # the open fuzz scene: meshes under mirroring transform chains, instances, transformed triangles and spheres,
    # mirror and glass all in view
    "fuzz_open_hw3": ("hw3", "fuzz_scene", dict(seed=1, width=96, height=72, camera_sphere_radius=40.0)),
    # HW7 with a Photographic tonemapping operator on the camera: the PNG is apply_tmo + the sRGB loop of main
    "hw7_tonemapped": ("hw7", "hw7_scene", dict(width=96, height=72, light_scale=3.0, tonemap=True)),
    # one run of the PPM binary (random_device seeds: a SAMPLE, compared statistically; two runs of the binary itself
    # differ by an RMSE of 9.1 grey levels at this photon count)
    "ppm_cfg5_small": ("ppm", "cfg5_ppm", dict(width=96, height=72, mesh_slices=(16, 14), photons_per_iteration=400,
                                               iterations=500, depth=8)),
}


#: float-level pins: name -> (semantics, reference binary, generator, kwargs).  The binary is the tree's own objects
#: linked with oracle/ref_driver_hdr.cpp (HW5/HW6: their main.cpp needs OpenCV; hw7f: the same twin for HW7), which
#: dumps Pixel::get_color() as float32 — so these fixtures pin the oracle to the last bit, not to 8 bits.
FLOAT_CASES = {
    # HW6: HW7's ray-tracing mode minus culling / the n.w_i gate, lights added straight onto the colour; 5 BRDFs
    "hw6_brdf_160": ("hw6", "hw6", "hw7_scene", dict(width=160, height=120, light_scale=3.0)),
    # HW5: the same tree without BRDF objects (the BRDF attributes of the scene are ignored by its loader)
    "hw5_lights_128": ("hw5", "hw5", "hw7_scene", dict(width=128, height=96, depth=2, n=9, seed=802)),
    "hw7_whitted_f32": ("hw7", "hw7f", "hw7_scene", dict(width=160, height=120, light_scale=3.0)),
}


ASSETS = os.path.join(HERE, "assets")   # texture images the HW7 scenes refer to (committed)


def scene_xml(name):
    tree, gen, kw = CASES[name] if name in CASES else DROPIN_CASES[name]
    if gen == "hw7_scene":
        os.makedirs(ASSETS, exist_ok=True)
        return tree, getattr(scenes, gen)(ASSETS, **kw)
    return tree, getattr(scenes, gen)(**kw)


def float_scene_xml(name):
    sem, binary, gen, kw = FLOAT_CASES[name]
    os.makedirs(ASSETS, exist_ok=True)
    return sem, binary, getattr(scenes, gen)(ASSETS, **kw)


if __name__ == "__main__":
    import numpy as np
    for name in FLOAT_CASES:
        sem, binary, xml = float_scene_xml(name)
        res = cpu.run_reference(binary, xml, workdir=ASSETS)
        hdr = list(res["hdr"].values())[0]
        np.save(os.path.join(HERE, name + ".npy"), hdr.astype(np.float32))
        print(name, hdr.shape, float(hdr.mean()))
    for f in os.listdir(ASSETS):
        if f.endswith(".f32") or f == "scene.xml":
            os.remove(os.path.join(ASSETS, f))
    meta = {}
    for name in list(CASES) + list(DROPIN_CASES):
        tree, xml = scene_xml(name)
        gen = (CASES[name] if name in CASES else DROPIN_CASES[name])[1]
        res = cpu.run_reference(tree, xml, workdir=ASSETS if gen == "hw7_scene" else None)
        pngs = [k for k in res["images"] if k.endswith(".png")]      # a tonemapped camera also leaves an .exr
        img = res["images"][pngs[0]] if pngs else list(res["images"].values())[0]
        image.write_png(os.path.join(HERE, name + ".png"), img)
        meta[name] = {"tree": tree, "shape": list(img.shape), "threads": res["threads"]}
        print(name, img.shape)
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
