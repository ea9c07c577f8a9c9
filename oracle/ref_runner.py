"""Runs the UNMODIFIED reference ``TRXOS`` staged under ``oracle/_ref`` (see oracle/build_ref.py) on the CPU.

TEST / BASELINE INFRASTRUCTURE ONLY: used by ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline``
leg as the timed CPU arm, and by tests to cross-check the oracle port.  Never imported by ``isbfsar_b200``.
The reference's call shape for B windows against one support set is SURVEY.md 3.3:
``model(None, labels[:1], {"sk": q}, ss_features=ssf.expand(B, -1, -1, -1))``.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json")) and os.path.exists(os.path.join(REF_DIR, "modules/ar/utils/model.py"))


def verify() -> None:
    """The staged files must be the ones the recipe copied (guards against edits of the reference arm)."""
    from .build_ref import sha256
    man = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))["sha256"]
    for rel, digest in man.items():
        if sha256(os.path.join(REF_DIR, rel)) != digest:
            raise RuntimeError(f"oracle/_ref/{rel} does not match its manifest: re-run python -m oracle.build_ref")


def _import_reference():
    """Import the staged reference modules (their top-level package names are `modules` and `utils`)."""
    sys.dont_write_bytecode = True
    for name in ("utils", "utils.params", "modules", "modules.ar", "modules.ar.utils", "modules.ar.utils.model"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "") or getattr(mod, "__path__", "")).count(REF_DIR):
            del sys.modules[name]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with contextlib.redirect_stdout(io.StringIO()):          # utils/params.py:11 prints at import
        from utils.params import TRXConfig
        from modules.ar.utils.model import TRXOS
    return TRXConfig, TRXOS


class ReferenceScorer:
    """The reference model itself (`modules/ar/utils/model.py:219-328`), CPU, fp32, with the synthetic weights."""

    def __init__(self, cfg, state_dict, threads: int | None = None):
        import torch
        verify()
        TRXConfig, TRXOS = _import_reference()
        if threads:
            torch.set_num_threads(threads)
        a = TRXConfig()
        a.device = "cpu"                                    # model.py:54 places the tuple tensors on args.device
        a.way, a.seq_len, a.temp_set = cfg.way, cfg.seq_len, list(cfg.temp_set)
        a.model = cfg.model
        self.torch = torch
        self.model = TRXOS(a).eval()
        full = {k: v.clone() for k, v in self.model.state_dict().items()}
        for k, v in state_dict.items():
            full[k] = torch.from_numpy(np.asarray(v))
        self.model.load_state_dict(full)

    def embed(self, support):
        with self.torch.no_grad():
            return self.model.features_extractor["sk"](self.torch.from_numpy(np.asarray(support)))

    def score(self, support, labels, query, chunk=512, ss_features=None):
        """-> (logits (B,W), is_true (B,1)) numpy, chunked like SURVEY 8d's CPU baseline."""
        torch = self.torch
        ssf = ss_features if ss_features is not None else self.embed(support)
        lab = torch.from_numpy(np.asarray(labels))
        lo, it = [], []
        with torch.no_grad():
            for s in range(0, query.shape[0], chunk):
                q = torch.from_numpy(query[s:s + chunk])
                r = self.model(None, lab, {"sk": q}, ss_features=ssf.expand(q.shape[0], -1, -1, -1))
                lo.append(r["logits"].numpy())
                it.append(r["is_true"].numpy())
        return np.concatenate(lo), np.concatenate(it)
