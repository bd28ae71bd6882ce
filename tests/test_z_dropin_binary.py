"""The drop-in, end to end: oracle/_ref/hw3_b200_raytracer is the reference's HW3 host — its XML loader, its `Scene`,
its BVH construction, its lodepng — linked from the tree's untouched objects, with the one block this repo replaces
(the std::thread fan-out over Scene::render_image, HW3/src/main.cpp:27-39) going through integration/flatten_hw3.cpp
and libb200rt.so instead (integration/main_hw3_b200.cpp; built by `oracle/build_ref.sh b200main3` in the authoring
container, shipped to the GPU box like the library itself).  It has to write the PNG the unmodified binary wrote.

Sorts last on purpose (see tests/test_z_open_fuzz.py): added after the round's GPU budget was spent."""
import os
import subprocess

import numpy as np
import pytest

from ceng795_b200 import image
from tests.golden.make_golden import scene_xml

HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "hw3_b200_raytracer")
needs_binary = pytest.mark.skipif(not os.path.exists(BINARY), reason="oracle/_ref/hw3_b200_raytracer not built")


def _run(xml, tmp_path):
    p = str(tmp_path / "scene.xml")
    with open(p, "w") as f:
        f.write(xml)
    return subprocess.run([BINARY, p], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)


@needs_binary
def test_dropin_binary_refuses_to_render_without_a_device(tmp_path):
    """No CPU fallback behind the reference's host either: on a box without a GPU the program parses the scene with
    the reference's loader, says why it cannot go on and exits non-zero without writing an image."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    _, xml = scene_xml("cfg2_hw3_n32")
    r = _run(xml, tmp_path)
    assert r.returncode == 2 and "no CUDA device" in r.stderr and "Scene is parsed" in r.stdout
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".png")]


@needs_binary
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2_hw3_n32", "cfg2_hw3_n48_smooth", "cfg2_hw3_glass", "fuzz_open_hw3"])
def test_dropin_binary_writes_the_reference_picture(gpu, tmp_path, name):
    """Same XML, same host code, the device instead of the thread pool: the PNG against the one the unmodified HW3
    binary wrote (tests/golden/<name>.png) — the 1-LSB bar every device image is held to (DESIGN.md §11 lists the
    two arithmetic deviations that can move a value across a rounding boundary)."""
    _, xml = scene_xml(name)
    r = _run(xml, tmp_path)
    assert r.returncode == 0, r.stderr
    pngs = [f for f in os.listdir(tmp_path) if f.endswith(".png")]
    assert len(pngs) == 1
    got = image.read_png(str(tmp_path / pngs[0]))[..., :3]
    want = image.read_png(os.path.join(HERE, "golden", name + ".png"))[..., :3]
    cmp = image.compare_8bit(got, want)
    assert cmp["within1"] >= 0.999 and cmp["max"] <= 2, cmp
    assert "kernel launches" in r.stdout
