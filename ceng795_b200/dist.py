"""Multi-GPU sharding of the hot path, one process per GPU, the scene replicated.

The data path is the C ABI's (include/b200rt.h "multi-GPU"): the frame lives on rank 0's GPU, every rank's final
combine kernel stores (1-sample cameras) or adds (multi-sample cameras) its share into it over NVLink peer memory, and
a counter at the end of the frame tells rank 0's stream when all shares have arrived.  torch.distributed is plumbing:
it carries the frame's handle (plain bytes) to the other ranks once, and all-reduces the PPM accepted-photon counts
after each pass (NCCL on the scene's own stream).  Nothing here touches pixel values.

* ``share_frame``    — rank 0 exports its Pixel array, everybody attaches.
* ``row_bands``      — Python twin of b200rt_multi_share's row split (8-row bands dealt round-robin), for tests.
* ``ppm_sync_counts``— delta of n since the last sync, all-reduce(sum), apply: every rank then tests the next pass
  against the same radii r2(n), as the reference's threads do on their shared Hit_point array
  (PPM/src/Scene.cpp:135-169).
* cameras are independent units: rank r renders cameras r, r+world, ... with no data-path collective (weak scaling).
"""
from __future__ import annotations

import os


def split_rows(height: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous band [begin, end) of rank; bands differ by at most one row."""
    base, rem = divmod(height, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def split_interleaved(height: int, rank: int, world: int) -> dict:
    """RenderOpts for the reference's `starting_row=rank, height_increase=world`."""
    return dict(row_begin=rank, row_end=height, row_stride=world)


def split_samples(n_strata: int, rank: int, world: int) -> dict:
    """RenderOpts for a sample-batch split (cfg 4): rank renders strata [begin, end) of samples_sqrt^2.
    Pixel{sum c*w, sum w} is additive over samples, so one sum-reduce assembles the frame."""
    b, e = split_rows(n_strata, rank, world)
    return dict(sample_begin=b, sample_end=e)


def split_photons(n_photons: int, rank: int, world: int) -> tuple[int, int]:
    """(first_photon, count) of rank: the reference hands photon_count/thread_count photons to each thread
    (PPM/src/main.cpp:74-88); here every photon index is its own RNG stream, so ranges just partition."""
    b, e = split_rows(n_photons, rank, world)
    return b, e - b


def allreduce_sum(t):
    """Sum over ranks (sample-batch frames; PPM {S.rgb, n} per hit point)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def cameras_of(n_cameras: int, rank: int, world: int) -> list[int]:
    return list(range(rank, n_cameras, world))


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world, local_rank); world==1 when not launched distributed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1:
        import torch
        import torch.distributed as dist
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


class _DevPtr:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def tensor_from_ptr(ptr: int, shape, device: int = 0, typestr: str = "<f4"):
    """Zero-copy torch view of a device buffer owned by the C-ABI scene ("<f4" float32, "<i4" int32)."""
    import torch
    return torch.as_tensor(_DevPtr(ptr, shape, typestr), device=f"cuda:{device}")


def row_bands(height: int, part: int, parts: int, band: int = 8) -> list[int]:
    """Rows of share `part`: bands of `band` rows starting at part*band, every parts*band rows — the reference's
    `j = i; j += thread_count` (HW7/src/main.cpp:39-44) in 8-row units (b200rt_multi_share)."""
    rows = []
    for start in range(part * band, height, parts * band):
        rows.extend(range(start, min(start + band, height)))
    return rows


def share_frame(sc, rank: int, world: int, camera_index: int = 0):
    """Rank 0 exports the camera's Pixel array; every rank (rank 0 included) attaches to it.  The handle is 160 plain
    bytes; torch.distributed only carries them."""
    import torch.distributed as dist
    box = [sc.frame_export(camera_index) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    sc.frame_attach(box[0])
    if world > 1:
        dist.barrier()   # nobody renders into the frame before everybody holds it


def sum_counts(delta, n_base):
    """The counts rule on tensors: n = n_at_last_sync + sum over ranks of (photons accepted locally since)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    return n_base + delta


def _allreduce_device(t, stream=None):
    """Sum over ranks of a device tensor, in place: NCCL on `stream`; with a gloo group (tests: several ranks on one
    GPU, which NCCL refuses) staged through host memory."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return
    if dist.get_backend() == "gloo":
        c = t.cpu()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        t.copy_(c)
        torch.cuda.synchronize()
    elif stream is not None:
        with torch.cuda.stream(stream):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        stream.synchronize()
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()


def ppm_sync_counts(sc, device: int, stream=None):
    """After a photon pass: all-reduce(sum) of the per-hit-point count deltas in place on the device (NCCL over
    NVLink, on `stream`), then n = n_base + sum on every rank."""
    p, n = sc.ppm_counts_delta()
    if n:
        _allreduce_device(tensor_from_ptr(p, (n,), device, "<i4"), stream)
    sc.ppm_counts_apply()


def ppm_sum_state(sc, device: int, stream=None):
    """{S.rgb, n} per hit point summed over the ranks in place (once per image): returns (device pointer, count).
    density = S / (pi r0^2 N) needs S only, so the summed n column is not used."""
    p, n = sc.ppm_export()
    if n:
        _allreduce_device(tensor_from_ptr(p, (n, 4), device), stream)
    return p, n
