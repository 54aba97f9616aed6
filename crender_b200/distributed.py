"""Multi-GPU plumbing for a Python host with one process per GPU (torchrun).

The multi-GPU path itself lives in the library (csrc/multi.cu): replicas, partition, NCCL merge on a side stream
with the resolve fused behind it. This module only (a) carries the 128-byte NCCL id from rank 0 to the other ranks
over torch.distributed and creates the rank handle (`rank_renderer`), and (b) keeps the pure host-side partition
arithmetic + a bring-your-own-collective merge that the world_size-2 gloo tests on CPU exercise.

The path shards with NO data-path exchange: every pixel-sample is independent
(_sample_pixel only touches its own pixel, src/render/renderer.cpp:362-383) and the scene is read-only
during a pass. Scene + BVH are replicated (each rank builds its own, the build is deterministic); the
work is partitioned either by sample index (spp partition) or by row bands (tile partition). The only
collective is the final merge of the float4 accumulation buffers: one all-reduce (sum) of W*H*4 floats
— 33 MB at 1080p — after which every rank re-resolves the display buffer. The sampler is keyed by the
GLOBAL pixel and sample index, so the union over ranks is the same set of paths as a single-GPU render;
only the float summation order across ranks differs.
"""
from __future__ import annotations

from typing import Tuple


def sample_range(rank: int, world: int, n_spp: int, first_sample: int = 0) -> Tuple[int, int]:
    """Contiguous global sample indices [lo, hi) rendered by `rank` (spp partition, BASELINE config 4)."""
    base, rem = divmod(n_spp, world)
    lo = first_sample + rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def row_bands(rank: int, world: int, height: int, band: int = 8):
    """Interleaved row bands [(y0, y1), ...] owned by `rank` (tile partition, BASELINE config 5)."""
    out = []
    for i, y0 in enumerate(range(0, height, band)):
        if i % world == rank:
            out.append((y0, min(height, y0 + band)))
    return out


def merge_partial_sums(raw_sum, passes_local: int, partition: str = "spp"):
    """All-reduce(sum) of the per-rank accumulation buffers, in place. `raw_sum` is a torch tensor (CUDA
    tensor aliasing the renderer's accumulation buffer with the nccl backend; a CPU tensor with gloo in the
    host-logic tests). Returns the merged pass count: summed for the spp partition, unchanged for tiles
    (each rank rendered every pass of its own rows; foreign rows are zero)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(raw_sum, op=dist.ReduceOp.SUM)
        if partition == "spp":
            import torch

            n = torch.tensor([passes_local], dtype=torch.int64, device=raw_sum.device)
            dist.all_reduce(n, op=dist.ReduceOp.SUM)
            return int(n.item())
    return int(passes_local)


class _DevicePtr:
    """Minimal __cuda_array_interface__ carrier so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None}


def accum_tensor(rend):
    """Zero-copy torch view of a crender_b200.api.renderer's float4 accumulation buffer."""
    import torch

    ptr, n = rend.accum_ptr()
    return torch.as_tensor(_DevicePtr(ptr, n), device="cuda")


def merge_renderer(rend, partition: str = "spp", passes_local=None) -> int:
    """Merge the accumulation buffers of all ranks into every rank's renderer and re-resolve.
    passes_local: passes this rank rendered per pixel. Defaults to the renderer's own counter, which is
    right for the spp partition; with the tile partition a rank issues one render call per row band, so the
    caller states the per-pixel pass count."""
    import torch

    rend.sync()
    t = accum_tensor(rend)
    torch.cuda.synchronize()
    if passes_local is None:
        if partition != "spp":
            raise ValueError("tile partition: pass the per-pixel pass count (passes_local)")
        passes_local = rend.current_stats().passes
    passes = merge_partial_sums(t, passes_local, partition)
    torch.cuda.synchronize()
    rend.set_pass_count(passes)
    rend.resolve()
    rend.sync()
    return passes


def comm_unique_id(lib_path=None) -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls this; the bytes go to every rank by any means)."""
    import ctypes as C

    from . import _capi

    lib = _capi.load(lib_path)
    buf = C.create_string_buffer(128)
    _capi.check(lib, lib.crb_comm_unique_id(buf))
    return buf.raw


def rank_renderer(res_x, res_y, bounces, scn, partition="spp", seed=0, **kw):
    """One process per GPU: every rank passes its own committed copy of the scene; rank 0's NCCL id is broadcast over
    torch.distributed (any backend), then the library owns the communicator (crb_render_create_rank) and
    renderer.render / current_progress / raw_sum work on the MERGED image like on one GPU."""
    import torch.distributed as dist

    from . import api

    rank, world = dist.get_rank(), dist.get_world_size()
    box = [comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    part = api.PARTITION_TILE if partition in ("tile", api.PARTITION_TILE) else api.PARTITION_SPP
    return api.renderer(res_x, res_y, bounces, scn, seed=seed, partition=part, comm=(box[0], rank, world), **kw)
