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


_comm_streams = {}
_blobs = {}


def broadcast_support(scorer, way: int, src: int = 0, group=None, device=None) -> None:
    """Rank `src` has called set_support for `way` classes; every other rank receives the support-set tuple
    embeddings.  `way` must be passed identically on every rank (no metadata exchange, no host synchronisation).
    On CUDA the export, the collective and the import run on a dedicated communication stream, so the query-side
    kernels of the next `score` overlap them; the scorer joins on its own event before it reads the operands."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    rank = dist.get_rank(group)
    n = scorer.support_blob_numel(way)
    on_cuda = device is not None and torch.device(device).type == "cuda"
    if not on_cuda:
        blob = scorer.export_support() if rank == src else torch.empty((n,), dtype=torch.float32, device=device)
        dist.broadcast(blob, src=src, group=group)
        if rank != src:
            scorer.import_support(blob, way)
        return
    key = (torch.device(device), n)
    if key not in _blobs:                     # persistent buffer: it is touched by another stream than the allocating one
        _blobs[key] = torch.empty((n,), dtype=torch.float32, device=device)
        _comm_streams[torch.device(device)] = torch.cuda.Stream(device=device)
    blob, comm = _blobs[key], _comm_streams[torch.device(device)]
    comm.wait_stream(torch.cuda.current_stream(device))      # set_support was issued on the current stream
    with torch.cuda.stream(comm):
        if rank == src:
            scorer.export_support(out=blob)
            exported = torch.cuda.Event()
            exported.record()
        dist.broadcast(blob, src=src, group=group)
        if rank != src:
            scorer.import_support(blob, way)
    if rank == src:
        # the next set_support (issued on the current stream) rewrites the tensors the export is reading: order it
        # behind the export copy (not behind the collective)
        torch.cuda.current_stream(device).wait_event(exported)


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


class ScoreGatherer:
    """Preallocated all-gather of equal shards: every rank scores `n_local` windows straight into its slot of a
    flat `[logits (n_local*way) | is_true (n_local)]` buffer; one `all_gather_into_tensor` per batch, no copies.

    `depth` > 1 gives that many buffer sets and a communication stream: `gather_async()` enqueues the collective
    of the batch just scored behind an event, `wait(ticket)` makes the current stream wait for it.  A caller that
    waits for batch k-1 after scoring batch k overlaps every collective (and the rank skew it absorbs) with the next
    batch's kernels."""

    def __init__(self, n_local: int, way: int, has_is_true: bool, device, group=None, depth: int = 1):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_local, self.way, self.has_is_true = n_local, way, has_is_true
        self.per = n_local * (way + (1 if has_is_true else 0))
        self.depth = max(1, depth)
        self._local = [torch.empty((self.per,), dtype=torch.float32, device=device) for _ in range(self.depth)]
        self._all = [torch.empty((self.world * self.per,), dtype=torch.float32, device=device) if self.world > 1 else self._local[i]
                     for i in range(self.depth)]
        self._cur = 0
        self._comm = torch.cuda.Stream(device=device) if (self.depth > 1 and self.world > 1 and torch.device(device).type == "cuda") else None
        self._done = [None] * self.depth

    @property
    def local(self):
        return self._local[self._cur]

    @property
    def all(self):
        return self._all[self._cur]

    def _views(self, flat):
        lo = flat[: self.n_local * self.way].view(self.n_local, self.way)
        it = flat[self.n_local * self.way:].view(self.n_local, 1) if self.has_is_true else None
        return lo, it

    def out(self):
        """(logits, is_true) views of this rank's slot: pass as `out=` to `scorer.score`."""
        return self._views(self.local)

    def _result(self, i):
        return [self._views(self._all[i][r * self.per:(r + 1) * self.per]) for r in range(self.world)]

    def gather(self):
        """-> list over ranks of (logits (n_local,way), is_true (n_local,1)) views of the gathered buffer."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.all, self.local, group=self.group)
        return self._result(self._cur)

    def gather_async(self):
        """Enqueue the all-gather of the batch just scored into the current buffer set on the communication stream and
        advance to the next set.  Returns a ticket for `wait`."""
        i = self._cur
        if self._comm is None:
            if self.world > 1:
                dist.all_gather_into_tensor(self._all[i], self._local[i], group=self.group)
        else:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(ready)
                if self.world > 1:
                    dist.all_gather_into_tensor(self._all[i], self._local[i], group=self.group)
                done = torch.cuda.Event()
                done.record()
            self._done[i] = done
        self._cur = (i + 1) % self.depth
        return i

    def wait(self, ticket):
        """Make the current stream wait for the collective of `ticket`; -> the per-rank views (valid until the
        buffer set comes round again, `depth` batches later)."""
        if self._done[ticket] is not None:
            torch.cuda.current_stream().wait_event(self._done[ticket])
        return self._result(ticket)


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
