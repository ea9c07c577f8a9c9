"""Sharding of query windows over the GPUs of one box.

Query windows are independent (every batch row of the reference forward is its own episode,
model.py:59-148), so the batch is split contiguously over ranks with no data-path collective.
The only exchanges are (1) one broadcast of the support-set operands per support-set change and
(2) one all-gather of `[logits | is_true]` per batch.  One process per GPU; `torch.distributed`
(NCCL over NVLink on the box, gloo in the CPU tests) owns the communicator -- only device pointers
cross the C ABI.

The scorer argument is anything with `export_support() -> Tensor`, `support_blob_numel(way) -> int`,
`import_support(blob, way)`, and `score(query) -> (logits (B,W), is_true (B,1)|None)`; the product
passes `isbfsar_b200.TRXOS`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous split of n windows: the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_support(scorer, way: int, src: int = 0, group=None, device=None) -> None:
    """Rank `src` has called set_support for `way` classes; every other rank receives the operands.
    `way` must be passed identically on every rank (no metadata exchange, no host synchronisation)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    rank = dist.get_rank(group)
    if rank == src:
        blob = scorer.export_support()
    else:
        blob = torch.empty((scorer.support_blob_numel(way),), dtype=torch.float32, device=device)
    dist.broadcast(blob, src=src, group=group)
    if rank != src:
        scorer.import_support(blob, way)


def gather_scores(logits: torch.Tensor, is_true, n_total: int, group=None):
    """All-gather the per-rank `[logits | is_true]` rows: one collective per batch.  Equal shards are gathered
    straight into the result; ragged shards are padded to the largest and trimmed."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return logits, is_true
    world = dist.get_world_size(group)
    way = logits.shape[1]
    cols = way + (1 if is_true is not None else 0)
    mx = shard_bounds(n_total, world, 0)[1]
    even = n_total % world == 0
    packed = torch.empty((mx, cols), dtype=torch.float32, device=logits.device)
    if not even:
        packed.zero_()
    packed[: logits.shape[0], :way] = logits
    if is_true is not None:
        packed[: logits.shape[0], way:] = is_true
    out = torch.empty((world * mx, cols), dtype=torch.float32, device=logits.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    if even:
        full = out
    else:
        out3 = out.view(world, mx, cols)
        full = torch.cat([out3[r, : shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0]] for r in range(world)])
    return full[:, :way], (full[:, way:] if is_true is not None else None)


def score_sharded(scorer, query_full_or_local, n_total: int, local: bool = False, group=None):
    """Score this rank's contiguous shard and gather everything.
    `query_full_or_local`: the full (B,T,3J) batch (every rank slices its shard) or, with local=True,
    only this rank's rows."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    s, e = shard_bounds(n_total, world, rank)
    q = query_full_or_local if local else query_full_or_local[s:e]
    assert q.shape[0] == e - s
    logits, is_true = scorer.score(q)
    return gather_scores(logits, is_true, n_total, group)
