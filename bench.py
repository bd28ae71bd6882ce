#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (Mrays/s, ms/frame) on the B200 rendering core.

  python bench.py --gpus N --steps K --warmup W                 # config 2 (the default, see below)
  python bench.py --config cfg4 --spp 64 ...                    # any other BASELINE config
  python bench.py --impl reference --gpus N --steps K ...       # the reference's own CPU renderer on the host cores

Default workload = cfg2, the configuration the metric is quoted on that fits one GPU: HW2-feature BVH scene in
the HW3 dialect, 1 048 352-triangle mesh + 8 instances + 4 spheres, 2 point lights, mirror depth 2, 1920x1080, 1 spp.

A step = one pass of the hot path over one batch of work per rank:
  cfg1..cfg4  one frame: generate -> (extend -> shade_a -> shadow -> shade_b)* -> combine
  cfg5        one photon pass (PhotonCountPerIteration photons): emit -> (extend -> deposit/scatter)*
Multi-GPU (one process per GPU, scene replicated, SURVEY.md §8e, include/b200rt.h "multi-GPU"):
  --split frame    ONE frame per step, shared by all ranks ("strong", the default): the Pixel array lives on rank 0's
                   GPU; 1-sample cameras are cut into 8-row bands dealt round-robin and every rank's last combine kernel
                   stores its pixels into rank 0's array over NVLink peer memory; multi-sample cameras are cut into
                   batches of sample strata, accumulated locally and added with vector atomics over NVLink.  A
                   system-scope counter at the end of the array tells rank 0's stream when all shares are in.
                   ms_per_frame is that frame's latency (max over ranks).
  --split cameras  rank r renders frame r of a camera orbit; no data-path exchange ("weak")
  cfg5             photons of the pass split by index; after each pass the accepted-photon counts are all-reduced
                   (NCCL on the scene's stream) so that every rank tests the next pass against the same radii; the
                   flux sums S are added once per image ("strong")
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/sec"
UNIT = "Mrays/s"


def workloads(args):
    """name -> dict(tree, xml(rank), base_dir, desc).  Scenes are generated (seeded) on the fly: data = synthetic."""
    from ceng795_b200 import scenes
    assets = os.path.join(tempfile.gettempdir(), "b200rt_bench_assets")
    os.makedirs(assets, exist_ok=True)
    spp4 = args.spp or 4096
    return {
        "cfg1": dict(tree="hw1", xml=lambda r: scenes.cfg1_hw1(), base_dir=None,
                     desc="cfg1: HW1 scene, 64 spheres + 2 meshes (2000 tris) + 10 triangles, 3 point lights, "
                          "Blinn-Phong + mirror depth 6, 800x800, 1 spp (HW1 tests every primitive; here a SAH tree + its order-dependent rule replayed)"),
        "cfg2": dict(tree="hw3", xml=lambda r: scenes.cfg2_hw3(cam_angle_deg=45.0 * r), base_dir=None,
                     desc="cfg2: HW3-dialect BVH scene, 1048352-tri heightfield + 8 instances + 4 spheres, "
                          "2 point lights, mirror depth 2, 1920x1080, 1 spp"),
        "cfg3": dict(tree="hw3", xml=lambda r: scenes.cfg3_hw3(spp=args.spp or 100), base_dir=None,
                     desc=f"cfg3: HW3 distribution ray tracing, 262088-tri mesh, 2 area lights, DOF, glossy mirror, "
                          f"glass, motion blur, 1920x1080, {args.spp or 100} spp"),
        "cfg4": dict(tree="hw7", base_dir=assets,
                     xml=lambda r: scenes.cfg4_hw7(assets, spp=spp4),
                     desc=f"cfg4: HW7 path tracing (importance sampling, depth 4), Cornell box + LightMesh + LightSphere + "
                          f"3 instances of a 100k-tri textured mesh, Perlin bump, 5 BRDFs, mirror, glass, 1920x1080, "
                          f"{spp4} spp camera; a step = {args.strata_per_step or 64} strata of it"),
        "cfg5": dict(tree="ppm", xml=lambda r: scenes.cfg5_ppm(photons_per_iteration=args.photons or 10_000_000),
                     base_dir=None,
                     desc=f"cfg5: PPM, closed box + mirror sphere + glass sphere + 50240-tri mesh, 1 point light, depth 8, "
                          f"1920x1080 eye pass, {args.photons or 10_000_000} photons per pass (a step = one pass)"),
    }


def config_dict(args, wl):
    """The workload both arms are measured on — identical in the two JSON lines."""
    split = "photons" if wl["tree"] == "ppm" else (args.split or "frame")
    return {"workload": wl["desc"], "split": split,
            "l2": "working set (scene + ~1 GB wavefront pool per step) exceeds the 126 MB L2; no explicit flush"}


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


# ------------------------------------------------------------------ reference CPU path
def reference_sample(args, wl, rays_per_step=None, sample_rays=None):
    """One bounded sample of the reference's own CPU implementation of the workload on this box's host cores.
    Returns dict(ms, rays, cores, kind, sample).  rays_per_step: ray count of the FULL step measured on the
    device (the binaries print only time); scaled to the sample."""
    from ceng795_b200 import scene, scenes
    from oracle import cpu
    tree = wl["tree"]
    name = args.config
    if name == "cfg3":
        spp, full = 4, (args.spp or 100)
        xml = scenes.cfg3_hw3(spp=spp)
        frac, what = spp / full, f"{spp} of {full} spp of the same frame"
    elif name == "cfg4":
        spp, full, per_step = 4, (args.spp or 4096), (args.strata_per_step or 64)
        xml = wl["xml"](0).replace(f"<NumSamples>{full}</NumSamples>", f"<NumSamples>{spp}</NumSamples>")
        assert f"<NumSamples>{spp}</NumSamples>" in xml
        frac, what = spp / min(full, per_step), f"{spp} spp of the same frame (a step is {min(full, per_step)} of its {full} strata)"
    elif name == "cfg5":
        full = args.photons or 10_000_000
        n = 400_000
        xml = scenes.cfg5_ppm(photons_per_iteration=n, iterations=1)
        frac, what = n / full, f"{n} of the pass's {full} photons (PPM binary built with -DNDEBUG, all threads)"
    else:
        xml = wl["xml"](0)
        frac, what = 1.0, "1 full frame of the same XML"
    if not cpu.have_ref(tree):
        # the oracle port (OpenMP) on a band of rows
        spec = scene.parse_xml(xml, tree, base_dir=wl["base_dir"])
        hs = scene.HostScene(spec)
        H = spec.cameras[0]["height"]
        band = (H // 2 - 30, H // 2 + 30)
        t0 = time.perf_counter()
        if tree == "ppm":
            _, _, _, cnt = cpu.ppm(hs, 0, n_photons=100_000, want_hit_points=False)
            rays = cnt.rays_shadow
        else:
            _, _, _, cnt = cpu.render(hs, 0, rows=band)
            rays = cnt.rays_total
        ms = (time.perf_counter() - t0) * 1e3
        return dict(ms=ms, rays=rays, cores=os.cpu_count(), kind="port",
                    sample=f"rows {band[0]}..{band[1]} through the C oracle port (OpenMP)")
    r = cpu.run_reference(tree, xml, workdir=wl["base_dir"])
    ms = r["ms"]["photon"] if tree == "ppm" else list(r["ms"].values())[0]
    if sample_rays is not None:
        rays = sample_rays
    elif rays_per_step is None:
        spec = scene.parse_xml(xml, tree, base_dir=wl["base_dir"])
        hs = scene.HostScene(spec)
        if tree == "ppm":
            _, _, _, cnt = cpu.ppm(hs, 0, n_photons=int(frac * (args.photons or 10_000_000)), want_hit_points=False)
            rays = cnt.rays_shadow
        else:
            _, _, _, cnt = cpu.render(hs, 0)   # the restatement's counters (pinned to the device's in tests/)
            rays = cnt.rays_total
    else:
        rays = rays_per_step * frac
    return dict(ms=ms, rays=rays, cores=r["threads"], kind="reference",
                sample=f"{what} through the untouched {tree.upper()} binary (oracle/_ref), timed by its own timer; "
                       "ray count = the device's for the same work (pinned to the oracle's counters by tests/)")


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl = workloads(args)[args.config]
    times, rays = [], None
    s = None
    budget_s = 240.0
    t_all = time.perf_counter()
    for k in range(args.warmup + args.steps):
        s = reference_sample(args, wl, sample_rays=rays)   # ray count: computed once (it is the same frame)
        if k >= args.warmup:
            times.append(s["ms"])
        rays = s["rays"]
        if time.perf_counter() - t_all > budget_s and times:
            break
    ms = sum(times) / len(times)
    value = rays / ms / 1e3
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if args.split == "cameras" else "strong",   # what the own arm reports for this split
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl), "detail": {"rays_per_step": rays},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": s["cores"], "kind": s["kind"], "sample": s["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ the CUDA core
def run_b200(args):
    import torch
    from ceng795_b200 import dist as rdist
    from ceng795_b200 import scene

    rank, world, local = rdist.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 core has no CPU fallback")
    torch.cuda.set_device(local)
    import torch.distributed as tdist

    wl = workloads(args)[args.config]
    ppm = wl["tree"] == "ppm"
    split = args.split or "frame"
    if ppm:
        split = "photons"
    xml = wl["xml"](rank if split == "cameras" else 0)
    sc = scene.Scene(xml, wl["tree"], device=local, base_dir=wl["base_dir"])
    W, H = sc.image_size(0)
    strata = sc.host.desc.cameras[0].samples_sqrt ** 2
    # cfg4 at its BASELINE size is a 4096-spp camera; a step renders one batch of `strata_per_step` strata of it
    strata_per_step = min(strata, args.strata_per_step or (64 if args.config == "cfg4" else strata))
    batches = max(1, strata // strata_per_step)
    step_no = [0]
    shared = world > 1 and split == "frame"
    if shared:
        rdist.share_frame(sc, rank, world, 0)
    stream = torch.cuda.ExternalStream(sc._L.b200rt_scene_stream(sc._scene), device=local)
    photons = args.photons or (sc.spec.ppm["photon_count_per_iteration"] if ppm else 0)
    ph_first, ph_cnt = rdist.split_photons(photons, rank, world) if ppm else (0, 0)
    pass_no = [0]
    if ppm and world > 1:
        sc.ppm_set_share(rank, world)
    eye = sc.eye_trace_lines(0, seed=args.seed) if ppm else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def step_opts():
        """The strata batch of this step (cfg4: 64 of the 4096), then this rank's share of it."""
        o = {}
        if strata_per_step < strata:
            b = (step_no[0] % batches) * strata_per_step
            o = dict(sample_begin=b, sample_end=b + strata_per_step)
        step_no[0] += 1
        if shared:
            if o:   # a strata batch is cut further by strata
                b, e = rdist.split_rows(strata_per_step, rank, world)
                o = dict(sample_begin=o["sample_begin"] + b, sample_end=o["sample_begin"] + e)
            else:
                o = {k: v for k, v in sc.multi_share(rank, world, 0).items() if k in
                     ("row_begin", "row_end", "row_stride", "row_group", "sample_begin", "sample_end")}
        return o

    def one_step(out=None, pipelined=False):
        if ppm:
            st = sc.trace_n_photons(ph_cnt, first_photon=pass_no[0] * photons + ph_first, seed=args.seed)
            pass_no[0] += 1
            if world > 1:
                rdist.ppm_sync_counts(sc, local, stream)   # counts over NVLink (NCCL), on the scene's stream
            return st
        o = step_opts()
        if shared:
            if rank == 0:
                sc.frame_begin()
            st = sc.render_shared(0, seed=args.seed, **o)
            if rank == 0:
                if pipelined and out is not None:
                    sc.frame_collect_async(world, out)     # the host copy overlaps the next frame (render_wait below)
                else:
                    sc.frame_collect(world, out)           # waits on the device for every rank's share
            return st
        if out is not None and pipelined:
            return sc.render_image_async(0, out, seed=args.seed, **o)
        if out is not None:
            return sc.render_image(0, out=out, seed=args.seed, **o)[1]
        return sc.render_device(0, seed=args.seed, **o)[1]

    for _ in range(args.warmup):
        one_step()
    if ppm:
        keys = ("photon_rays", "cell_scans", "entries_scanned", "deposits", "nodes_visited", "tris_tested", "launches",
                "grid_rebuilds")
        ms_keys = ("ms_extend", "ms_deposit")
    else:
        keys = ("rays_primary", "rays_shadow", "rays_secondary", "nodes_visited", "tris_tested", "spheres_tested",
                "instances_entered", "sh_nodes_visited", "sh_tris_tested", "sh_spheres_tested", "sh_instances_entered",
                "launches", "waves")
        ms_keys = ("ms_generate", "ms_extend", "ms_shade", "ms_shadow")
    agg = {k: 0 for k in keys}
    ms_stage = {k: 0.0 for k in ms_keys}
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
    rays = agg["photon_rays"] if ppm else agg["rays_primary"] + agg["rays_shadow"] + agg["rays_secondary"]

    # ---- end to end through the C ABI with HOST buffers ---------------------------
    # two page-locked host frames: step k's frame is copied into buffer k % 2 while step k+1 renders
    # (b200rt_render_async / b200rt_frame_collect_async); every step's frame reaches host memory inside the timed region
    pinned = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    outs = [p.numpy() for p in pinned]
    out = outs[0]
    e2e_no = [0]

    def e2e_step():
        if ppm:
            st = one_step()
            if world > 1:
                p, _ = rdist.ppm_sum_state(sc, local, stream)     # flux sums of all ranks, then the image on rank 0
                if rank == 0:
                    sc.density_estimation_from(p, photons * pass_no[0], out=out)
            else:
                sc.density_estimation(photons * pass_no[0], out=out)
            return st.photon_rays
        e2e_no[0] += 1
        st = one_step(outs[e2e_no[0] % 2], pipelined=not args.no_pipeline)
        return st.rays_total

    for _ in range(max(1, min(3, args.warmup))):
        e2e_step()
    if not ppm:
        sc.render_wait()
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for _ in range(args.steps):
        e2e_rays += e2e_step()
    if not ppm:
        sc.render_wait()            # the last frames' host copies
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
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return 0

    hbm, hbm_src = peaks()
    # second ceiling: scene fetches are served by L1/L2, so the traversal kernels are quoted against the L2
    # read bandwidth measured here (48 MB working set, after the timed region); HBM is the ceiling only where ncu
    # shows DRAM traffic of the order of the algorithmic bytes (the PPM deposit)
    l2 = None
    import ctypes as C
    g = C.c_double(0.0)
    if sc._L.b200rt_measure_read_bandwidth(local, 48 << 20, 50, C.byref(g)) == 0:
        l2 = g.value
    traffic_all = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic_all = json.load(open(tp))
        except Exception:
            traffic_all = {}
    traffic_cfg = traffic_all.get(args.config, {})
    nb = int(st.node_bytes) or 64

    def kernel_line(name, bytes_, ms, launches, bound, traffic_key):
        """achieved = algorithmic bytes / CUDA-event time of the stage; frac against the ceiling the counters name."""
        gbs = bytes_ / (ms * 1e-3) / 1e9 if ms > 0 else None
        peak = l2 if bound == "l2" else hbm
        tr = traffic_cfg.get(traffic_key, {}).get("dram_bytes_per_launch") if isinstance(traffic_cfg.get(traffic_key), dict) else None
        avg_ms = ms / max(1, launches)
        return {"kernel": name, "bound": bound, "achieved": gbs, "peak": peak, "unit": "GB/s",
                "frac": (gbs / peak) if (gbs and peak) else None,
                "hbm_frac_algorithmic": (gbs / hbm) if gbs else None,
                "traffic": tr, "dram_frac": (tr / (avg_ms * 1e-3) / 1e9 / hbm) if (tr and avg_ms > 0) else None,
                "algorithmic_bytes_per_launch": bytes_ / max(1, launches), "avg_launch_ms": avg_ms,
                "share_of_step": ms / ms_total if ms_total > 0 else None}

    if ppm:
        n_closest = agg["photon_rays"]
        top_tests = int(st.top_objects) * n_closest
        ext_bytes = (nb * (agg["nodes_visited"] - top_tests) + 32 * top_tests + 48 * agg["tris_tested"]
                     + (32 + 24) * n_closest)
        n_launch = max(1, args.steps * (sc.spec.max_depth - 1) * ((photons + (8 << 20) - 1) // (8 << 20)))   # per 8 Mi-photon batch
        per_ray = {"nodes": (agg["nodes_visited"] - top_tests) / max(1, n_closest),
                   "top_level_gates": int(st.top_objects), "tris": agg["tris_tested"] / max(1, n_closest),
                   "hit_points_scanned_per_photon_ray": agg["entries_scanned"] / max(1, n_closest),
                   "reference_loop_visits_per_photon_ray": agg["cell_scans"] / max(1, n_closest),
                   "scan_cell_over_reference_cell": float(st.scan_cell) / (2.0 * float(eye.initial_radius)) if eye.initial_radius else None,
                   "grid_rebuilds": agg["grid_rebuilds"],
                   "deposits_per_photon_ray": agg["deposits"] / max(1, n_closest)}
        # deposit: per photon ray 52 B in + 56 B out of the pool, 16 B per scanned grid entry, 64 B record + 4 B count
        # + 16 B flux per deposit ("32 k + 16 a", with this layout's record sizes)
        dep_bytes = 108 * n_closest + 16 * agg["entries_scanned"] + 84 * agg["deposits"]
        kernels = [kernel_line("k_ppm_photon_shade (scatter + cell scan + hit-point radius / flux update)", dep_bytes,
                               ms_stage["ms_deposit"], n_launch, "hbm", "deposit"),
                   kernel_line(f"k_trace_{'wide' if nb != 64 else 'fast'}<false,7> (photon closest-hit traversal)",
                               ext_bytes, ms_stage["ms_extend"], n_launch, "l2", "extend")]
    else:
        n_closest = agg["rays_primary"] + agg["rays_secondary"]
        n_any = agg["rays_shadow"]
        top_c, top_a = int(st.top_objects) * n_closest, int(st.top_objects) * n_any
        ext_bytes = (nb * (agg["nodes_visited"] - top_c) + 32 * top_c + 48 * agg["tris_tested"]
                     + 20 * agg["spheres_tested"] + 112 * agg["instances_entered"] + (32 + 24) * n_closest)
        sh_bytes = (nb * (agg["sh_nodes_visited"] - top_a) + 32 * top_a + 48 * agg["sh_tris_tested"]
                    + 20 * agg["sh_spheres_tested"] + 112 * agg["sh_instances_entered"] + (36 + 1) * n_any)
        n_launch = max(1, agg["waves"])   # one closest-hit and one shadow launch per wave of every row / strata batch
        mode = 7 if wl["tree"] in ("hw4", "hw5", "hw6", "hw7", "ppm") else 2
        if wl["tree"] == "hw1":
            kn = "k_trace_hw1_fast<{}> (SAH tree over all triangles, HW1's order-dependent rule on the near-tie set)"
        elif nb != 64:
            kn = "k_trace_wide<{}," + str(mode) + "> (8-wide quantised tree, shared-memory stack, reference gate per candidate)"
        else:
            kn = "k_trace_fast<{}," + str(mode) + "> (binary SAH tree + reference gate)"
        per_ray = {"nodes": (agg["nodes_visited"] - top_c) / max(1, n_closest), "top_level_gates": int(st.top_objects),
                   "tris": agg["tris_tested"] / max(1, n_closest),
                   "shadow_nodes": (agg["sh_nodes_visited"] - top_a) / max(1, n_any),
                   "shadow_tris": agg["sh_tris_tested"] / max(1, n_any), "node_bytes": nb}
        kernels = [kernel_line(kn.format("false") + " closest hit", ext_bytes, ms_stage["ms_extend"], n_launch, "l2", "extend"),
                   kernel_line(kn.format("true") + " any hit (shadow rays)", sh_bytes, ms_stage["ms_shadow"], n_launch, "l2", "shadow")]
    kernels.sort(key=lambda k: -(k["share_of_step"] or 0.0))
    dom = kernels[0]
    frames = (world if split == "cameras" else 1) * strata_per_step / strata
    line = {
        "metric": METRIC, "value": rays / ms_total / 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if split == "cameras" else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(args, wl),
        "detail": {"rays_per_step": rays / args.steps, "frames_per_step": frames,
                   "scene_bytes": int(sc._L.b200rt_scene_bytes(sc._scene))},
        "ms_per_frame": ms_total / args.steps * strata / strata_per_step,
        "stage_ms_per_step": {k: v / args.steps for k, v in ms_stage.items()},
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": UNIT, "ms_per_step": e2e_s * 1e3 / args.steps,
                "h2d_bytes_per_step": 176 + 64,
                "d2h_bytes_per_step": H * W * 16 * (world if split == "cameras" else 1)},
        "gpu_launches": launches_all,
        "clocks": clocks,
        # the dominant kernel BY TIME, against the ceiling its counters name; the other kernels of the step follow
        "roofline": {"bound": dom["bound"], "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"],
                     "peak_source": ("measured here: b200rt_measure_read_bandwidth, 48 MB x 50 passes (L2 read)"
                                     if dom["bound"] == "l2" else hbm_src),
                     "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"], "dram_frac": dom["dram_frac"],
                     "traffic_source": traffic_all.get("_source"),
                     "hbm": {"peak": hbm, "peak_source": hbm_src, "frac_algorithmic": dom["hbm_frac_algorithmic"]},
                     "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                     "avg_launch_ms": dom["avg_launch_ms"], "share_of_step": dom["share_of_step"],
                     "per_ray": per_ray, "other_kernels": kernels[1:]},
    }
    if ppm:
        line["detail"]["photons_per_step"] = photons
        line["detail"]["hit_points"] = int(eye.hit_points)
        line["photons_per_s"] = photons * args.steps / (ms_total * 1e-3)
    if world == 1 and not args.no_cpu_baseline:
        try:
            s = reference_sample(args, wl, rays_per_step=rays / args.steps)
            line["cpu_baseline"] = {"value": s["rays"] / s["ms"] / 1e3, "unit": UNIT, "cores": s["cores"],
                                    "kind": s["kind"], "ms": s["ms"], "sample": s["sample"]}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable",
                                    "sample": repr(e)}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--split", default=None, choices=["frame", "cameras"])
    ap.add_argument("--strata-per-step", type=int, default=None,
                    help="strata of the camera rendered per step (cfg4 default 64 of its 4096)")
    ap.add_argument("--spp", type=int, default=None, help="NumSamples override for cfg3 / cfg4")
    ap.add_argument("--photons", type=int, default=None, help="photons per pass override for cfg5")
    ap.add_argument("--seed", type=int, default=795)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="e2e: wait for every frame's host copy before the next frame starts (b200rt_render)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
