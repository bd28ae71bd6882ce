#!/usr/bin/env python
"""Re-renders, through the CPU oracle and with the same counter-based samples, the rows of cfg4 (4096 spp) in which
tools/find_nonfinite.py found non-finite pixels on the GPU: are they non-finite in the restatement of the reference too?

    python tools/check_nonfinite_oracle.py gpurun_out/nonfinite_cfg4.json
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceng795_b200 import scene, scenes  # noqa: E402
from oracle import cpu  # noqa: E402

info = json.load(open(sys.argv[1]))
spp = info["spp"]
assets = tempfile.mkdtemp()
xml = scenes.hw7_scene(assets, width=1920, height=1080, n=224, mesh_slices=(224, 225), spp=spp, depth=4,
                       path_tracing=True, importance=True, stochastic_lights=True, perlin=True, light_scale=0.25)
hs = scene.HostScene(scene.parse_xml(xml, "hw7", base_dir=assets))
for y, x in info["pixels"]:
    os.environ["ORACLE_COLS"] = f"{max(0, x - 2)}:{x + 3}"          # the pixel and its two neighbours on either side
    rgbw, _, _, _ = cpu.render(hs, 0, rows=(y, y + 1), seed=info["seed"])
    row = rgbw[y]
    bad = np.argwhere(~np.isfinite(row).all(-1)).ravel().tolist()
    print(f"row {y}: GPU non-finite at x={x}; oracle (columns {os.environ['ORACLE_COLS']}) non-finite at x={bad}", flush=True)
