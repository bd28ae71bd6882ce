#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (Mrays/s, ms/frame) on config 2:
HW2-feature BVH scene in the HW3 dialect, 1 048 352-triangle mesh + 8 instances
+ 4 spheres, 2 point lights, mirror depth 2, 1920x1080, 1 spp.

  python bench.py --gpus N --steps K --warmup W            # the CUDA core (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU renderer on the host cores

A step = one pass of the hot path (generate -> extend -> shade -> shadow ->
combine) over one frame per rank.  At N > 1 the frames are independent units
(rank r renders camera r of an orbit; the scene is replicated): no data-path
collective, "scaling": "weak".  `--split rows` instead splits ONE frame into row
bands and gathers them over NCCL (strong scaling; used for the heavy configs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/sec"
UNIT = "Mrays/s"
WORKLOAD = "cfg2: HW3-dialect BVH scene, 1048352-tri heightfield + 8 instances + 4 spheres, 2 point lights, mirror depth 2, 1920x1080, 1 spp"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def orbit_xml(rank: int):
    """Camera r of the orbit: the canonical cfg2 camera rotated by r*45 degrees about +y."""
    from ceng795_b200 import scenes
    return scenes.cfg2_hw3(cam_angle_deg=45.0 * rank)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ceng795_b200 import scene, scenes
    from oracle import cpu
    xml = scenes.cfg2_hw3()
    spec = scene.parse_xml(xml, "hw3")
    hs = scene.HostScene(spec)
    have = cpu.have_ref("hw3")
    cores = os.cpu_count() or 1
    H = spec.cameras[0]["height"]
    if have:
        kind = "reference"
        # one full frame through the untouched HW3 binary tells us how many frames fit in the budget
        r = cpu.run_reference("hw3", xml)
        first_ms = list(r["ms"].values())[0]
        cores = r["threads"]
        # ray count of a frame: the oracle restatement's counters (the binary reports only time)
        _, _, _, cnt = cpu.render(hs, 0)
        rays = cnt.rays_total
        budget_s = 240.0
        full = (args.steps + args.warmup) * first_ms / 1e3 <= budget_s
    else:
        kind = "port"
        full = False
        first_ms = None
        rays = None
    times = []
    if have and full:
        sample = "full frame per step through oracle/_ref/hw3_raytracer (its own 'is saved in' timer)"
        for k in range(args.warmup + args.steps):
            r = cpu.run_reference("hw3", xml)
            if k >= args.warmup:
                times.append(list(r["ms"].values())[0])
        ms = sum(times) / len(times)
        rays_step = rays
    else:
        # bounded sample: a band of rows through the OpenMP oracle port, all host threads
        band = (H // 2 - 60, H // 2 + 60)
        sample = f"rows {band[0]}..{band[1]} of the frame per step through the C oracle port (OpenMP, {cores} threads)"
        kind = "port"
        rays_step = None
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _, _, _, cnt = cpu.render(hs, 0, rows=band)
            dt = (time.perf_counter() - t0) * 1e3
            if k >= args.warmup:
                times.append(dt)
            rays_step = cnt.rays_total
        ms = sum(times) / len(times)
    value = rays_step / ms / 1e3
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step": rays_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def cpu_baseline(xml, host_scene, rays_per_frame):
    """Reference CPU path on this box's host cores, one bounded sample (rank 0, N=1 only)."""
    from oracle import cpu
    try:
        if cpu.have_ref("hw3"):
            r = cpu.run_reference("hw3", xml)
            ms = list(r["ms"].values())[0]
            return {"value": rays_per_frame / ms / 1e3, "unit": UNIT, "cores": r["threads"], "kind": "reference",
                    "ms_per_frame": ms,
                    "sample": "1 full frame of the same XML through the untouched HW3 binary (oracle/_ref), "
                              "timed by its own 'is saved in' timer; ray count = this frame's (pinned to the "
                              "oracle's count by tests/test_gpu_parity.py)"}
        H = host_scene.spec.cameras[0]["height"]
        band = (H // 2 - 60, H // 2 + 60)
        t0 = time.perf_counter()
        _, _, _, cnt = cpu.render(host_scene, 0, rows=band)
        ms = (time.perf_counter() - t0) * 1e3
        return {"value": cnt.rays_total / ms / 1e3, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": f"rows {band[0]}..{band[1]} through the C oracle port (OpenMP)"}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}


def run_b200(args):
    import numpy as np
    import torch
    from ceng795_b200 import dist as rdist
    from ceng795_b200 import scene

    rank, world, local = rdist.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 core has no CPU fallback")
    torch.cuda.set_device(local)
    import torch.distributed as tdist

    split = args.split
    xml = orbit_xml(rank if split == "cameras" else 0)
    sc = scene.Scene(xml, "hw3", device=local)
    W, H = sc.image_size(0)
    opts = {}
    if split == "rows" and world > 1:
        b, e = rdist.split_rows(H, rank, world)
        opts = dict(row_begin=b, row_end=e)
    stream = torch.cuda.ExternalStream(sc._L.b200rt_scene_stream(sc._scene), device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def one_step():
        p, st = sc.render_device(0, **opts)
        if split == "rows" and world > 1:
            # tiles gathered over NVLink: one all-gather of the row bands
            frame = rdist.tensor_from_ptr(p, (H, W, 4), local)
            rdist.gather_rows(frame[opts["row_begin"]:opts["row_end"]], H, rank, world)
        return st

    # ---- device-resident timing ------------------------------------------------
    for _ in range(args.warmup):
        one_step()
    agg = {k: 0 for k in ("rays_primary", "rays_shadow", "rays_secondary", "nodes_visited", "tris_tested",
                          "spheres_tested", "instances_entered", "sh_nodes_visited", "sh_tris_tested",
                          "sh_spheres_tested", "sh_instances_entered", "launches")}
    ms_stage = {"ms_generate": 0.0, "ms_extend": 0.0, "ms_shade": 0.0, "ms_shadow": 0.0}
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        st = one_step()
        for k in agg:
            agg[k] += getattr(st, k)
        for k in ms_stage:
            ms_stage[k] += getattr(st, k)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    rays = agg["rays_primary"] + agg["rays_shadow"] + agg["rays_secondary"]

    # ---- end to end through the C ABI with HOST buffers ---------------------------
    pinned = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
    out = pinned.numpy()
    for _ in range(max(1, min(3, args.warmup))):
        sc.render_image(0, out=out, **opts)
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for _ in range(args.steps):
        _, st = sc.render_image(0, out=out, **opts)
        e2e_rays += st.rays_total
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
        r = torch.tensor([rays, e2e_rays, agg["launches"]], dtype=torch.float64, device="cuda")
        tdist.all_reduce(r, op=tdist.ReduceOp.SUM)
        rays, e2e_rays, launches_all = float(r[0]), float(r[1]), int(r[2])
    else:
        launches_all = agg["launches"]
    if rank != 0:
        return 0

    hbm, hbm_src = peaks()
    n_closest = agg["rays_primary"] + agg["rays_secondary"]
    ext_bytes = (64 * agg["nodes_visited"] + 48 * agg["tris_tested"] + 20 * agg["spheres_tested"]
                 + 112 * agg["instances_entered"] + (32 + 24) * n_closest)
    sh_bytes = (64 * agg["sh_nodes_visited"] + 48 * agg["sh_tris_tested"] + 20 * agg["sh_spheres_tested"]
                + 112 * agg["sh_instances_entered"] + (36 + 1) * agg["rays_shadow"])
    ext_gbs = ext_bytes / (ms_stage["ms_extend"] * 1e-3) / 1e9 if ms_stage["ms_extend"] > 0 else None
    sh_gbs = sh_bytes / (ms_stage["ms_shadow"] * 1e-3) / 1e9 if ms_stage["ms_shadow"] > 0 else None
    n_ext_launch = args.steps * 3
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_extend_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": rays / ms_total / 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if split == "cameras" else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "split": split, "frames_per_step": world if split == "cameras" else 1,
                   "rays_per_step": rays / args.steps,
                   "l2": "working set (112 MB scene + ~1 GB wavefront pool per frame) exceeds the 126 MB L2; no explicit flush",
                   "scene_bytes": int(sc._L.b200rt_scene_bytes(sc._scene))},
        "ms_per_frame": ms_total / args.steps,
        "stage_ms_per_frame": {k: v / args.steps for k, v in ms_stage.items()},
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": UNIT, "ms_per_frame": e2e_s * 1e3 / args.steps,
                "h2d_bytes_per_step": 176 + 64, "d2h_bytes_per_step": H * W * 16 * (world if split == "cameras" else 1)},
        "gpu_launches": launches_all,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_trace_fast<false,2> (closest-hit traversal, SAH tree + reference gate)", "achieved": ext_gbs,
                     "peak": hbm, "peak_source": hbm_src, "unit": "GB/s",
                     "frac": (ext_gbs / hbm) if ext_gbs else None, "traffic": traffic,
                     "algorithmic_bytes_per_launch": ext_bytes / n_ext_launch,
                     "avg_launch_ms": ms_stage["ms_extend"] / n_ext_launch,
                     "per_ray": {"nodes": agg["nodes_visited"] / max(1, n_closest),
                                 "tris": agg["tris_tested"] / max(1, n_closest)},
                     "shadow_kernel": {"achieved": sh_gbs, "frac": (sh_gbs / hbm) if sh_gbs else None,
                                       "nodes_per_ray": agg["sh_nodes_visited"] / max(1, agg["rays_shadow"])}},
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(xml, sc.host, rays / args.steps)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--split", default="cameras", choices=["cameras", "rows"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
