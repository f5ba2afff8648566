"""Multi-GPU plumbing: chains are independent, so they are partitioned contiguously over
ranks with NO collective in the data path; the only exchange is a gather of per-chain diagnostics / samples
(NCCL over NVLink on the GPU box, gloo in the CPU tests) so that rank 0 can write one posterior file.

RNG streams are keyed by the GLOBAL chain id (``engine.chain_offset`` + local index), so a chain's trajectory does not
depend on how many ranks the job runs on.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_chains(num_chains: int, world_size: int, rank: int):
    """Contiguous partition: returns (global id of this rank's first chain, number of local chains).
    The first ``num_chains % world_size`` ranks hold one extra chain."""
    base, extra = divmod(int(num_chains), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def _gather_tensor(t: torch.Tensor, chain_dim: int, counts, group=None):
    """all-gather ``t`` along ``chain_dim`` when ranks hold different numbers of chains (pad to the maximum)."""
    world = dist.get_world_size(group)
    cmax = max(counts)
    moved = t.movedim(chain_dim, 0).contiguous()
    as_bool = moved.dtype == torch.bool
    if as_bool:
        moved = moved.to(torch.uint8)
    if moved.shape[0] < cmax:
        pad = torch.zeros((cmax - moved.shape[0],) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
        moved = torch.cat([moved, pad], dim=0)
    out = torch.empty((world * cmax,) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
    dist.all_gather_into_tensor(out, moved, group=group)
    parts = [out[r * cmax: r * cmax + counts[r]] for r in range(world)]
    full = torch.cat(parts, dim=0)
    if as_bool:
        full = full.to(torch.bool)
    return full.movedim(0, chain_dim)


def gather_chains(tree, num_chains: int, chain_dim: int = 1, group=None):
    """Gather a (nested dict / list of) per-rank tensors with a chain axis into the global chain order on every rank.
    ``chain_dim`` = 1 for draw-major traces ``[n, B_local, ...]``, 0 for per-chain vectors ``[B_local, ...]``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tree
    world = dist.get_world_size(group)
    counts = [shard_chains(num_chains, world, r)[1] for r in range(world)]

    def rec(node):
        if isinstance(node, dict):
            return {k: rec(v) for k, v in node.items()}
        if isinstance(node, (list, tuple)):
            return type(node)(rec(v) for v in node)
        return _gather_tensor(node, chain_dim, counts, group)

    return rec(tree)
