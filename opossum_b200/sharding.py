"""Multi-GPU layer of the ray path: rays shard by source index, one process per GPU, and the only exchange is the
all-reduce of detector results.

    rank g owns source positions [g*N/G, (g+1)*N/G)        -- whole wavelength groups stay on one GPU (rays.rs:284-289)
    fluence map   : all-reduce(min/max) of the data-dependent map range (rays_hit_map.rs:522-562), every rank bins its
                    hits into the SAME grid, all-reduce(sum) of the maps (hit_map/mod.rs:363-379 merges hit maps the same way)
    spot diagram  : all-reduce of the phase-1 sums, then of the phase-2 radius accumulators (rays.rs:599-701)
    tallies       : all-reduce(sum) of per-bounce-order counts (int64) and energies (f64)

Everything here is host logic over `torch.distributed` (NCCL on the GPUs, gloo in the CPU tests); the per-ray work is
done by the caller's kernels.  Reductions of a handful of scalars travel as one small tensor.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when no process group is initialised"""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_positions: int, rank: int, world_size: int) -> tuple[int, int]:
    """contiguous index range [g*N/G, (g+1)*N/G) of the source POSITIONS owned by `rank`"""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return (rank * n_positions) // world_size, ((rank + 1) * n_positions) // world_size


def _reduce(values, op, device, dtype=torch.float64):
    t = torch.tensor(list(values), dtype=dtype, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=op)
    return t.tolist()


def global_map_range(lrtb, device="cpu"):
    """(left, right, top, bottom) over all ranks from each rank's own bounding box; a rank without hits passes None.
    Raises like the reference (`could not calculate bounding box`) when no rank has a hit."""
    inf = math.inf
    l, r, t, b = lrtb if lrtb is not None else (inf, -inf, -inf, inf)
    nl, r, t, nb = _reduce([-l, r, t, -b], dist.ReduceOp.MAX, device)
    if not (math.isfinite(nl) and math.isfinite(r) and math.isfinite(t) and math.isfinite(nb)):
        raise ValueError("could not calculate bounding box")
    return (-nl, r, t, -nb)


def allreduce_map(fluence_map: torch.Tensor) -> torch.Tensor:
    """sum of the per-rank maps (all binned on the same global grid), in place"""
    if world()[1] > 1:
        dist.all_reduce(fluence_map, op=dist.ReduceOp.SUM)
    return fluence_map


def allreduce_tally(counts, energies, device="cpu"):
    """per-bounce-order valid-ray counts (exact, int64) and energies (f64) over all ranks"""
    c = _reduce([int(x) for x in counts], dist.ReduceOp.SUM, device, torch.int64)
    e = _reduce([float(x) for x in energies], dist.ReduceOp.SUM, device)
    return c, e


def reduce_spot_sums(p, device="cpu"):
    """phase 1 -> global sums (fields of an OpmSpotPartial-like object are updated in place)"""
    sums = _reduce(list(p.sum), dist.ReduceOp.SUM, device)
    n = _reduce([int(p.n_valid), int(p.n_total)], dist.ReduceOp.SUM, device, torch.int64)
    # bounce level = bounce count of the first valid ray (rays.rs:183-194): ids are local to a shard and shards are
    # ordered by rank, so the global first valid ray is the first one of the lowest rank that has any
    rank = world()[0]
    big = (1 << 62)
    key = ((rank << 32) | (int(p.first_key) & 0xFFFFFFFF)) if int(p.n_valid) > 0 else big
    key = -_reduce([-key], dist.ReduceOp.MAX, device, torch.int64)[0]
    for j in range(7):
        p.sum[j] = sums[j]
    p.n_valid, p.n_total = n
    p.first_key = 0xFFFFFFFFFFFFFFFF if key >= big else key
    return p


def reduce_spot_radii(p, device="cpu"):
    """phase 2 -> global radius accumulators"""
    p.max_d2_mm2 = _reduce([p.max_d2_mm2], dist.ReduceOp.MAX, device)[0]
    p.sum_d2_mm2, p.sum_ed2 = _reduce([p.sum_d2_mm2, p.sum_ed2], dist.ReduceOp.SUM, device)
    return p


# ---------------------------------------------------------------------------------------------------------------
# drivers over the product API (GPU): used by bench.py and by applications that shard a scene over the GPUs of a box
# ---------------------------------------------------------------------------------------------------------------
def fluence_map(ctx, hit_maps, nx: int, ny: int, map_tensor: torch.Tensor):
    """HitMap::calc_fluence_map(Binning) of hit maps sharded over the ranks, result replicated on every rank.
    `map_tensor`: nx*ny float64 CUDA tensor that receives the (all-reduced) map.  Returns FluenceData (map on the device)."""
    from . import api
    dev = map_tensor.device
    try:
        lrtb, _ = api.hitmaps_bounding_box(hit_maps)
    except api.OpossumError as e:
        if e.code != api.F.OPM_ERR_EMPTY_HITMAP or world()[1] == 1:
            raise
        lrtb = None
    if world()[1] > 1:
        lrtb = global_map_range(lrtb, dev)
    fd = api.fluence_binning(ctx, hit_maps, nx, ny, lrtb=lrtb, map_dev=map_tensor.data_ptr(), download=False)
    if world()[1] > 1:
        allreduce_map(map_tensor)
        fd = api.fluence_peak_total(ctx, map_tensor.data_ptr(), nx, ny, lrtb)
    return fd


def spot_stats(scene, node: int, device):
    """SpotDiagram::node_report of a bundle sharded over the ranks (exact two-phase reduction)"""
    from . import _ffi as F
    import ctypes as C
    p = scene.node_spot_partial(node)
    if world()[1] > 1:
        reduce_spot_sums(p, device)
    scene.node_spot_partial(node, reduced=p)
    if world()[1] > 1:
        reduce_spot_radii(p, device)
    out = F.OpmSpotStats()
    F.lib().opm_spot_stats_from_partial(C.byref(p), C.byref(out))
    return out
