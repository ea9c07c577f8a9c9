"""world_size-2 gloo test of the multi-GPU plumbing on CPU: contiguous sharding of the query windows,
one broadcast of the support operands, one all-gather of [logits | is_true].  The scorer injected here is
the CPU oracle (test infrastructure) standing in for the CUDA scorer, so only the host logic is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.synth import Cfg, make_episode, make_state_dict
from oracle.trx_oracle import TrxOracle


class OracleScorer:
    """Same small interface isbfsar_b200.dist expects from TRXOS."""

    def __init__(self, cfg, sd):
        self.o = TrxOracle(cfg, sd)
        self.cfg = cfg
        self.feats = None

    def set_support(self, poses):
        self.feats = self.o.embed(torch.from_numpy(poses))

    def export_support(self):
        return self.feats.reshape(-1).clone()

    def support_blob_numel(self, way):
        return way * self.cfg.seq_len * self.cfg.trans_linear_in_dim

    def import_support(self, blob, way):
        self.feats = blob.reshape(way, self.cfg.seq_len, self.cfg.trans_linear_in_dim).clone()

    def score(self, q):
        way = self.feats.shape[0]
        if q.shape[0] == 0:
            return torch.zeros((0, way)), torch.zeros((0, 1))
        lo, it = self.o.score(None, np.arange(way)[None], q.numpy(), ss_features=self.feats[None])
        return torch.from_numpy(lo), torch.from_numpy(it)


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isbfsar_b200.dist import broadcast_support, score_sharded, shard_bounds
        cfg = Cfg()
        sd = make_state_dict(cfg, 0)
        support, labels, query, _ = make_episode(cfg, B, 1, "structured")
        sc = OracleScorer(cfg, sd)
        if rank == 0:
            sc.set_support(support[0])            # only rank 0 sees the support poses
        broadcast_support(sc, way=5, src=0)
        assert sc.feats is not None and sc.feats.shape == (5, 16, 256)
        logits, is_true = score_sharded(sc, torch.from_numpy(query), B)
        s, e = shard_bounds(B, world, rank)
        q.put((rank, logits.numpy(), is_true.numpy(), (s, e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 16])
def test_sharded_scoring_world2(B):
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = Cfg()
    sd = make_state_dict(cfg, 0)
    support, labels, query, _ = make_episode(cfg, B, 1, "structured")
    lo, it = TrxOracle(cfg, sd).score(support, labels, query)
    for rank, logits, is_true, _ in res:
        assert logits.shape == (B, 5) and is_true.shape == (B, 1)
        np.testing.assert_allclose(logits, lo, rtol=1e-5, atol=1e-6)      # every rank holds the full result
        np.testing.assert_allclose(is_true, it, rtol=1e-5, atol=1e-6)


def _gatherer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isbfsar_b200.dist import ScoreGatherer
        n_local, way = 6, 5
        g = ScoreGatherer(n_local, way, True, "cpu", depth=2)
        tickets, got = [], []
        for k in range(4):                      # pipelined: the collective of batch k is joined after batch k+1
            lo, it = g.out()
            lo.copy_(torch.full((n_local, way), 100.0 * k + rank))
            it.copy_(torch.full((n_local, 1), 100.0 * k + rank + 0.5))
            tickets.append(g.gather_async())
            if len(tickets) > 1:
                got.append([(a.clone(), b.clone()) for a, b in g.wait(tickets.pop(0))])
        got.append([(a.clone(), b.clone()) for a, b in g.wait(tickets.pop(0))])
        ok = len(got) == 4
        for k, per_rank in enumerate(got):
            for r, (lo, it) in enumerate(per_rank):
                ok = ok and bool((lo == 100.0 * k + r).all()) and bool((it == 100.0 * k + r + 0.5).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_pipelined_score_gatherer_world2():
    """ScoreGatherer with two buffer sets: batch k's all-gather is joined one batch later; every rank sees every shard."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gatherer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
