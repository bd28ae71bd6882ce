#!/usr/bin/env python
"""One process per GPU sharing ONE frame and ONE set of PPM counts through the C ABI's export / attach primitives —
the path bench.py takes under torchrun — checked on rank 0 against the same work done by a single scene.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tools/multi_rank_check.py [--backend nccl|gloo]

With fewer devices than ranks the ranks share devices (cudaIpc works between processes on one GPU too); NCCL refuses
two ranks on one GPU, so that case needs --backend gloo (the handle and the count deltas then travel over gloo).
Prints one JSON line from rank 0; exit code 0 = every check held."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    args = ap.parse_args()
    import torch
    import torch.distributed as tdist
    from ceng795_b200 import dist as rdist
    from ceng795_b200 import scene, scenes

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank))) % max(1, torch.cuda.device_count())
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    kw = {"device_id": torch.device("cuda", local)} if args.backend == "nccl" else {}
    tdist.init_process_group(backend=args.backend, rank=rank, world_size=world, **kw)
    res = {"world": world, "backend": args.backend}
    ok = True
    try:
        # ---- (1) 1-sample camera: row bands stored in place over peer memory, bit-equal to one device ----
        # ---- (2) multi-sample camera: strata batches added with vector atomics, equal up to summation order ----
        for name, xml, tree in (("hw3_1spp", scenes.cfg2_hw3(n=96, width=320, height=200), "hw3"),
                                ("hw3_16spp", scenes.cfg3_hw3(n=48, width=160, height=96, spp=16), "hw3")):
            sc = scene.Scene(xml, tree, device=local)
            W, H = sc.image_size(0)
            rdist.share_frame(sc, rank, world, 0)
            share = {k: v for k, v in sc.multi_share(rank, world, 0).items()
                     if k in ("row_begin", "row_end", "row_stride", "row_group", "sample_begin", "sample_end")}
            out = np.zeros((H, W, 4), np.float32)
            rays = []
            for frame in range(3):                       # frames follow each other without a host barrier
                if rank == 0:
                    sc.frame_begin()
                st = sc.render_shared(0, seed=11, **share)
                rays.append(st.rays_total)
                if rank == 0:
                    sc.frame_collect(world, out)
            t = torch.tensor([float(rays[-1])], dtype=torch.float64)
            if args.backend == "nccl":
                t = t.cuda()
            tdist.all_reduce(t)
            tdist.barrier()
            sc.frame_detach()
            if rank == 0:
                ref, st1 = sc.render_image(0, seed=11)
                same = bool(np.array_equal(out.view(np.uint32), ref.view(np.uint32)))
                close = bool(np.allclose(out, ref, rtol=2e-5, atol=1e-6))
                res[name] = {"bit_equal": same, "close": close, "rays": float(t[0]), "rays_single": int(st1.rays_total)}
                ok &= close and float(t[0]) == st1.rays_total and (same or name != "hw3_1spp")
            tdist.barrier()
            sc.close()
        # ---- (3) PPM: photons split by index, counts synchronised after every pass ----
        sc = scene.Scene(scenes.cfg5_ppm(width=160, height=120, mesh_slices=(24, 22)), "ppm", device=local)
        sc.ppm_set_share(rank, world)
        s0 = sc.eye_trace_lines(0, seed=5)
        hp = sc.hit_points()
        passes, photons = 4, 200_000
        for k in range(passes):
            first, cnt = rdist.split_photons(photons, rank, world)
            sc.trace_n_photons(cnt, first_photon=k * photons + first, seed=5)
            rdist.ppm_sync_counts(sc, local)
        n_multi = sc.hit_points()[:, 15].copy()
        p, n = rdist.ppm_sum_state(sc, local)
        if rank == 0:
            img = sc.density_estimation_from(p, photons * passes)
            one = scene.Scene(scenes.cfg5_ppm(width=160, height=120, mesh_slices=(24, 22)), "ppm", device=local)
            one.eye_trace_lines(0, seed=5)
            hp1 = one.hit_points()
            for k in range(passes):
                one.trace_n_photons(photons, first_photon=k * photons, seed=5)
            ref = one.density_estimation(photons * passes)
            n_one = one.hit_points()[:, 15]
            a = img[..., :3] / np.maximum(img[..., 3:], 1e-30)
            b = ref[..., :3] / np.maximum(ref[..., 3:], 1e-30)
            blk = lambda x: x.reshape(120 // 8, 8, 160 // 8, 8, 3).mean((1, 3))
            rr = float(np.sqrt(np.mean((blk(a) - blk(b)) ** 2)) / np.mean(blk(b)))
            res["ppm"] = {"hit_points": int(s0.hit_points),
                          "same_numbering": bool(np.array_equal(hp[:, :15].view(np.uint32), hp1[:, :15].view(np.uint32))),
                          "accepted_multi": float(n_multi.sum()), "accepted_single": float(n_one.sum()),
                          "mean_multi": float(a.mean()), "mean_single": float(b.mean()), "block_rel_rmse": rr}
            ok &= res["ppm"]["same_numbering"]
            ok &= abs(n_multi.sum() - n_one.sum()) <= 0.03 * n_one.sum()      # counts are shared: same radii schedule
            ok &= abs(a.mean() - b.mean()) <= 0.02 * b.mean() and rr <= 0.08
            one.close()
        tdist.barrier()
        sc.close()
    finally:
        tdist.destroy_process_group()
    if rank == 0:
        res["ok"] = bool(ok)
        print(json.dumps(res))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
