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


# ---- the product path: per-burst gather to the rank that writes the posterior file -----------------------------------------
def world():
    """(rank, world_size) of the default process group, (0, 1) outside torch.distributed."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(device_index=None):
    """One process per GPU under ``torchrun``: initialise NCCL (or gloo without CUDA) from RANK / WORLD_SIZE / LOCAL_RANK.
    Returns (rank, world_size, local_rank); a no-op (0, 1, 0) when WORLD_SIZE is unset or 1."""
    import os

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1:
        return 0, 1, 0
    local = int(os.environ.get("LOCAL_RANK", "0")) if device_index is None else int(device_index)
    if not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    return dist.get_rank(), dist.get_world_size(), local


_RAW = {torch.uint16: torch.uint8, torch.int16: torch.uint8}  # dtypes NCCL / gloo do not take: moved as raw bytes (last axis x 2)


def _gather_to(t: torch.Tensor, chain_dim: int, counts, dst: int, group=None, max_bytes=1 << 30):
    """Gather ``t`` (chain axis ``chain_dim``, ``counts[r]`` chains on rank r) to rank ``dst`` in global chain order; other
    ranks get None.  Large tensors (the compact event draws) travel in slices of the leading axis of at most ``max_bytes``
    per rank, so that the receive buffers stay bounded."""
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    cmax = max(counts)
    raw = _RAW.get(t.dtype)
    x = t.contiguous().view(raw) if raw is not None else (t.to(torch.uint8) if t.dtype == torch.bool else t)
    x = x.movedim(chain_dim, 0).contiguous()  # [B_local, ...]
    if x.shape[0] < cmax:
        x = torch.cat([x, torch.zeros((cmax - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)], dim=0)
    rest = tuple(x.shape[1:])
    lead = rest[0] if rest else 1
    per_lead = x[0].numel() * x.element_size() * cmax // max(lead, 1)
    step = max(1, min(lead, max_bytes // max(per_lead, 1))) if rest else 1
    pieces = []
    for a in range(0, lead, step):
        part = x[:, a:a + step].contiguous() if rest else x
        bufs = [torch.empty_like(part) for _ in range(ws)] if rank == dst else None
        dist.gather(part, bufs, dst=dst, group=group)
        if rank == dst:
            pieces.append(torch.cat([bufs[r][:counts[r]] for r in range(ws)], dim=0))
        if not rest:
            break
    if rank != dst:
        return None
    full = torch.cat(pieces, dim=1) if rest else pieces[0]
    full = full.movedim(0, chain_dim)
    if t.dtype == torch.bool:
        return full.to(torch.bool)
    return full.contiguous().view(t.dtype) if raw is not None else full


def gather_to_rank0(tree, num_chains: int, chain_dim: int = 1, group=None, dst: int = 0, max_bytes=1 << 30):
    """The per-burst exchange of the product path: every rank's slice of a (nested dict / list of) tensors with
    a chain axis -> rank ``dst`` in global chain order (NCCL gather over NVLink on the GPU box, gloo in the CPU tests).
    Returns the gathered tree on ``dst`` and None elsewhere; the identity outside torch.distributed.  ``None`` leaves pass.

    All leaves travel in ONE message per rank -- each leaf's chains as rows of raw bytes, the leaves side by side (a burst's
    trace is ~20 small tensors: one collective of latency instead of twenty) -- except leaves above ``max_bytes`` per rank
    (the compact event draws), which go on their own in bounded slices of the draw axis."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tree
    ws = dist.get_world_size(group)
    counts = [shard_chains(num_chains, ws, r)[1] for r in range(ws)]
    me = dist.get_rank(group)
    cmax = max(counts)
    leaves = []

    def collect(node):
        if node is None:
            return
        if isinstance(node, dict):
            for v in node.values():
                collect(v)
        elif isinstance(node, (list, tuple)):
            for v in node:
                collect(v)
        else:
            leaves.append(node)

    collect(tree)
    small, packed_rows = [], []
    for t in leaves:
        x = t.movedim(chain_dim, 0).contiguous()  # [B_local, ...]
        nbytes = (x[0].numel() if x.shape[0] else 0) * x.element_size()
        if nbytes * cmax > max_bytes or x.shape[0] == 0:
            small.append(None)
            continue
        flat = (x.to(torch.uint8) if x.dtype == torch.bool else x).reshape(x.shape[0], -1)
        if flat.stride() != (flat.shape[1], 1):  # (a size-1 axis keeps whatever stride it had: "contiguous", but not viewable as bytes)
            flat = torch.empty(flat.shape, dtype=flat.dtype, device=flat.device).copy_(flat)
        rows = flat.view(torch.uint8)  # [B_local, nbytes]
        small.append((tuple(x.shape[1:]), t.dtype, rows.shape[1]))
        packed_rows.append(rows)
    gathered = None
    if packed_rows:
        buf = torch.cat(packed_rows, dim=1)
        if buf.shape[0] < cmax:
            buf = torch.cat([buf, torch.zeros((cmax - buf.shape[0], buf.shape[1]), dtype=torch.uint8, device=buf.device)], dim=0)
        bufs = [torch.empty_like(buf) for _ in range(ws)] if me == dst else None
        dist.gather(buf.contiguous(), bufs, dst=dst, group=group)
        if me == dst:
            gathered = torch.cat([bufs[r][:counts[r]] for r in range(ws)], dim=0)  # [B_total, total bytes]
    out_leaves, col = [], 0
    for t, meta in zip(leaves, small):
        if meta is None:
            out_leaves.append(_gather_to(t, chain_dim, counts, dst, group, max_bytes))
            continue
        rest, dtype, nb = meta
        if me == dst:
            raw = gathered[:, col:col + nb].contiguous()
            x = raw.view(torch.uint8 if dtype == torch.bool else dtype).reshape((raw.shape[0],) + rest)
            if dtype == torch.bool:
                x = x.to(torch.bool)
            out_leaves.append(x.movedim(0, chain_dim))
        else:
            out_leaves.append(None)
        col += nb
    if me != dst:
        return None
    it = iter(out_leaves)

    def rebuild(node):
        if node is None:
            return None
        if isinstance(node, dict):
            return {k: rebuild(v) for k, v in node.items()}
        if isinstance(node, (list, tuple)):
            return type(node)(rebuild(v) for v in node)
        return next(it)

    return rebuild(tree)
