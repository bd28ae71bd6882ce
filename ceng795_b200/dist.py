"""Multi-GPU sharding of the hot path: one process per GPU, the
scene replicated, image rows or cameras split across ranks.

* ``split_rows``     — contiguous row bands per rank (strong scaling of one frame);
  ``gather_rows`` assembles the frame on rank 0 with one collective
  (`all_gather_into_tensor` over NCCL/NVLink on GPU tensors, gloo on CPU).
* ``split_interleaved`` — the reference's own scheme, `j = rank; j += world`
  (HW7/src/main.cpp:39-44); the partial frames have disjoint rows so a
  sum-reduce assembles them.
* cameras are independent units: rank r renders cameras r, r+world, ... with no
  data-path collective (weak scaling).

torch.distributed is plumbing only; nothing here touches pixel values.
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


def gather_rows(local_band, height: int, rank: int, world: int):
    """local_band: tensor [rows_of_rank, W, 4] (GPU for nccl, CPU for gloo).
    Returns the assembled [H, W, 4] frame on every rank (all-gather of padded bands)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_band
    max_rows = split_rows(height, 0, world)[1] - split_rows(height, 0, world)[0]
    pad = torch.zeros((max_rows,) + tuple(local_band.shape[1:]), dtype=local_band.dtype, device=local_band.device)
    pad[: local_band.shape[0]] = local_band
    out = torch.empty((world * max_rows,) + tuple(local_band.shape[1:]), dtype=local_band.dtype,
                      device=local_band.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(world):
        b, e = split_rows(height, r, world)
        parts.append(out[r * max_rows: r * max_rows + (e - b)])
    return torch.cat(parts, 0)


class _DevPtr:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


def tensor_from_ptr(ptr: int, shape, device: int = 0):
    """Zero-copy torch view of a float32 device buffer owned by the C-ABI scene."""
    import torch
    return torch.as_tensor(_DevPtr(ptr, shape), device=f"cuda:{device}")


def reduce_interleaved(partial_frame):
    """Sum of frames whose rows are disjoint (interleaved split) == the frame."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(partial_frame, op=dist.ReduceOp.SUM)
    return partial_frame
