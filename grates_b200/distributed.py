"""Multi-GPU partitioning for the spherical-harmonic hot path: one process per GPU
(torchrun), ``torch.distributed`` for the plumbing.

The path shards without any data-path collective: epochs are independent for synthesis,
analysis and filtering; grid rows (parallels) are independent for covariance propagation.
Collectives appear only around the kernels: an optional broadcast of the covariance matrix
from the rank that holds it, and an optional gather of the result shards.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(count, world_size, rank):
    """Contiguous shard [start, stop) of ``count`` independent units; the first
    ``count % world_size`` ranks get one extra unit."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("invalid rank {0} for world size {1}".format(rank, world_size))
    base, extra = divmod(count, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_counts(count, world_size):
    return [shard_range(count, world_size, r)[1] - shard_range(count, world_size, r)[0] for r in range(world_size)]


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_shards(local, counts, group=None):
    """All-gather shards of unequal length along dim 0.  ``local``: tensor [counts[rank], ...] on
    the device matching the backend (CUDA for nccl, CPU for gloo).  Returns the concatenation
    [sum(counts), ...] on every rank."""
    world, rank = _world(group)
    if world == 1:
        return local
    if local.shape[0] != counts[rank]:
        raise ValueError("local shard has {0} rows, expected {1}".format(local.shape[0], counts[rank]))
    width = max(counts)
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    buf = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded.contiguous(), group=group)
    parts = [buf[r * width:r * width + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def synthesis_sharded(anm, grid, kernel='ewh', GM=3.9860044150e+14, R=6.3781363000e+06, gather=False, group=None):
    """Epoch-sharded batched synthesis.  ``anm``: the FULL [E, L, L] coefficient batch (numpy or
    tensor, identical on every rank); each rank synthesises its contiguous epoch shard on its own
    GPU.  Returns (values, (start, stop)): the local shard [stop-start, nlat, nlon] as a CUDA
    tensor, or the full [E, nlat, nlon] tensor on every rank with ``gather=True`` (NCCL
    all-gather over NVLink; 62 MB per rank at config 2)."""
    from .gravityfield import to_grid_batch
    world, rank = _world(group)
    E = anm.shape[0]
    start, stop = shard_range(E, world, rank)
    local = anm[start:stop]
    if not isinstance(local, torch.Tensor):
        local = torch.as_tensor(np.ascontiguousarray(local, dtype=float)).cuda()
    values = to_grid_batch(local.contiguous(), grid, kernel, GM, R)
    if gather:
        return gather_shards(values, shard_counts(E, world), group), (0, E)
    return values, (start, stop)


def covariance_propagation_sharded(covariance_matrix, grid, min_degree, max_degree, kernel='potential',
                                   GM=3.9860044150e+14, R=6.3781363000e+06, src=0, gather=True, group=None):
    """Row-block-sharded covariance propagation.  The covariance matrix needs to be valid on
    rank ``src`` only (pass None elsewhere); it is broadcast once (708 MB at degree 96), every rank
    propagates its block of parallels, and the standard deviations (2 MB in total) are gathered.
    Returns the [nlat*nlon] standard deviations (gather=True) or the local [rows, nlon] block."""
    from . import plan as _plan
    world, rank = _world(group)
    p = _plan.get_plan(grid, max_degree, kernel, GM, R)
    dev = torch.device("cuda", p.device)
    kp = (max_degree + 1) ** 2 - min_degree ** 2
    if rank == src:
        sigma = torch.as_tensor(np.ascontiguousarray(covariance_matrix, dtype=float)) if not isinstance(
            covariance_matrix, torch.Tensor) else covariance_matrix
        sigma = sigma.to(dev).contiguous()
    else:
        sigma = torch.empty((kp, kp), dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(sigma, src=src, group=group)
    rows = covariance_row_blocks(p.nlat, world, rank)
    parts = []
    for kind, start, count in rows:
        if count:
            parts.append(p.covariance_propagation(sigma, min_degree, start, count, mirrored=(kind == "mirrored")))
    local = torch.cat(parts) if parts else torch.empty((0, p.nlon), dtype=torch.float64, device=dev)
    if not gather:
        return local
    gathered = gather_shards(local, covariance_row_counts(p.nlat, world), group)
    return reorder_row_blocks(gathered, p.nlat, world).reshape(-1)


def covariance_row_counts(nlat, world):
    """Output rows every rank contributes under covariance_row_blocks."""
    return [sum((2 * c if k == "mirrored" else c) for k, _, c in covariance_row_blocks(nlat, world, r)) for r in range(world)]


def reorder_row_blocks(gathered, nlat, world):
    """Rank-major concatenation of the ranks' (northern block, mirrored block[, equator]) outputs -> rows in the order
    of the parallels."""
    out = torch.empty_like(gathered)
    at = 0
    for r in range(world):
        for kind, start, count in covariance_row_blocks(nlat, world, r):
            out[start:start + count] = gathered[at:at + count]
            at += count
            if kind == "mirrored":
                out[nlat - start - count:nlat - start] = gathered[at:at + count]
                at += count
    return out


def covariance_row_blocks(nlat, world, rank):
    """Row blocks of rank `rank`: the northern parallels are cut into contiguous shards and every rank takes its shard
    together with the mirror images (GB_COV_MIRRORED: both halves share the first contraction on equator-symmetric
    grids); with an odd number of parallels the equator goes to the last rank as a plain block.
    Returns [(kind, first parallel, count)], kind "mirrored" or "plain"."""
    half = nlat // 2
    start, stop = shard_range(half, world, rank)
    blocks = [("mirrored", start, stop - start)]
    if nlat % 2 and rank == world - 1:
        blocks.append(("plain", half, 1))
    return blocks
