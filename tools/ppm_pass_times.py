#!/usr/bin/env python
"""Wall and device time of every photon pass of cfg5 (tuning aid: where the scan-grid rebuilds fall and what they cost)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ceng795_b200 import scene, scenes
sc = scene.Scene(scenes.cfg5_ppm(), "ppm")
t0 = time.perf_counter(); eye = sc.eye_trace_lines(0, seed=795); print("eye pass", round((time.perf_counter() - t0) * 1e3, 1), "ms wall", eye.hit_points, "hit points")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for k in range(n):
    t0 = time.perf_counter()
    st = sc.trace_n_photons(10_000_000, first_photon=k * 10_000_000, seed=795)
    w = (time.perf_counter() - t0) * 1e3
    print(f"pass {k:2d} wall {w:7.2f} ms  device {st.ms_photon:7.2f}  extend {st.ms_extend:6.2f} deposit {st.ms_deposit:6.2f}  rebuilds {st.grid_rebuilds} cell {st.scan_cell:.4f} scanned/ray {st.entries_scanned / max(1, st.photon_rays):.1f}")
