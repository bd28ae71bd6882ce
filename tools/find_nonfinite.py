#!/usr/bin/env python
"""cfg4 at its full 4096 spp: which pixels hold a non-finite sample?  Prints them (GPU box), so that the oracle can be
asked about exactly those pixels with the same counter-based samples (tools/check_nonfinite_oracle.py, CPU).

    gpurun -- python tools/find_nonfinite.py [spp]
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceng795_b200 import scene, scenes  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
assets = tempfile.mkdtemp()
xml = scenes.hw7_scene(assets, width=1920, height=1080, n=224, mesh_slices=(224, 225), spp=spp, depth=4,
                       path_tracing=True, importance=True, stochastic_lights=True, perlin=True, light_scale=0.25)
sc = scene.Scene(xml, "hw7", base_dir=assets)
a, st = sc.render_image(0, seed=spp)
bad = np.argwhere(~np.isfinite(a).all(-1))
out = dict(spp=spp, seed=spp, pixels=[[int(y), int(x)] for y, x in bad], values=[[float(v) for v in a[y, x]] for y, x in bad])
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "nonfinite_cfg4.json"), "w"))
