"""Regenerates tests/golden/leaf_vectors.bin: known-answer vectors of the reference's own leaf functions
(Bounding_box::intersect, the five BRDFs, Texture::get_color_at in its four mode combinations, Perlin_noise with a
fixed table, Mesh_triangle::intersect), produced by oracle/_ref/ref_probe = HW7's unmodified objects + oracle/ref_probe.cpp.

    bash oracle/build_ref.sh probe && python tests/golden/make_leaf_vectors.py
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

if __name__ == "__main__":
    out = os.path.join(HERE, "leaf_vectors.bin")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_probe"), out, os.path.join(HERE, "assets")], check=True)
    print(out, os.path.getsize(out), "bytes")
