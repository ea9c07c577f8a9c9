"""CPU restatement of the reference TRX-OS scoring path (torch CPU ops, fp32 or fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never imported by the
product package.  Each function cites the reference lines it restates
(paths relative to /root/reference).  Arithmetic uses the same torch CPU
primitives the reference dispatches to (linear, layer_norm eps=1e-5, softmax,
matmul), so in fp32 it agrees with the reference to summation-order noise;
``dtype=torch.float64`` gives the error-budget yardstick.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as Fn

from .synth import Cfg, tuple_table


def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.detach().to("cpu", dtype)
    return torch.as_tensor(np.asarray(x)).to(dtype)


class TrxOracle:
    """Functional restatement of TRXOS (modules/ar/utils/model.py:219-328)."""

    def __init__(self, cfg: Cfg, state_dict: dict, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.w = {k: _t(v, dtype) for k, v in state_dict.items()}
        T = cfg.seq_len
        # model.py:52-55 -- lexicographic combinations, int64
        self.tuples = [torch.from_numpy(tuple_table(T, c)) for c in cfg.temp_set]

    # -- model.py:164-180: relu(fc2(relu(fc1(x)))) -- ReLU on the output too
    def embed(self, x):
        w = self.w
        h = torch.relu(Fn.linear(x, w["features_extractor.sk.fc1.weight"], w["features_extractor.sk.fc1.bias"]))
        return torch.relu(Fn.linear(h, w["features_extractor.sk.fc2.weight"], w["features_extractor.sk.fc2.bias"]))

    # -- model.py:26-28,65-72: x + pe[:, :T]; tuple feature = concat of frame features in tuple order
    def tuple_features(self, feats, ti=0):
        pe = self.w[f"transformers.{ti}.pe.pe"]
        x = feats + pe[:, : feats.shape[-2]]
        idx = self.tuples[ti]                                   # (N, c)
        g = x[..., idx, :]                                      # (..., N, c, F)
        return g.reshape(*g.shape[:-2], -1)                     # (..., N, c*F)

    # -- model.py:75-84: k_linear / v_linear on both sides, LayerNorm on K only
    def project(self, tup, ti=0):
        p = f"transformers.{ti}."
        w = self.w
        k = Fn.linear(tup, w[p + "k_linear.weight"], w[p + "k_linear.bias"])
        v = Fn.linear(tup, w[p + "v_linear.weight"], w[p + "v_linear.bias"])
        D = k.shape[-1]
        k = Fn.layer_norm(k, (D,), w[p + "norm_k.weight"], w[p + "norm_k.bias"], 1e-5)
        return k, v

    # -- model.py:59-148
    def cross_transformer(self, ss_feats, ss_labels, q_feats, ti=0, want=("logits",)):
        """ss_feats (b,W,T,F), ss_labels (b,W) (row 0 read, model.py:95), q_feats (b,1,T,F)."""
        D = self.cfg.trans_linear_out_dim
        sk, sv = self.project(self.tuple_features(ss_feats, ti), ti)     # (b,W,N,D)
        qk, qv = self.project(self.tuple_features(q_feats, ti), ti)      # (b,1,N,D)
        N = qk.shape[-2]
        logits, diffs, protos, probs = [], [], [], []
        for c in ss_labels[0].tolist():
            ck = sk[:, c:c + 1]                                          # index_select(-3, c)
            cv = sv[:, c:c + 1]
            s = torch.matmul(qk, ck.transpose(-2, -1)) / math.sqrt(D)    # (b,1,Nq,Ns)
            p = torch.softmax(s, dim=-2)                                 # over QUERY tuples (model.py:49,109)
            proto = torch.matmul(p, cv)                                  # (b,1,Nq,D)
            diff = qv - proto
            dist = torch.norm(diff, dim=[-2, -1]) ** 2 / N               # model.py:131-132
            logits.append(-dist)
            if "diffs" in want:
                diffs.append(diff)
            if "prototypes" in want:
                protos.append(proto)
            if "probs" in want:
                probs.append(p)
        out = {"logits": torch.cat(logits, dim=1)}
        if diffs:
            out["diffs"] = torch.cat(diffs, dim=1)
        if protos:
            out["prototypes"] = protos
        if probs:
            out["probs"] = probs
        return out

    # -- model.py:183-204
    def discriminator(self, feature):
        w = self.w
        p = "discriminator."
        b = feature.shape[0]
        y = Fn.linear(feature, w[p + "dimensionality_reduction.weight"], w[p + "dimensionality_reduction.bias"])
        y = y.reshape(b, -1)
        y = torch.relu(Fn.linear(y, w[p + "fc1.weight"], w[p + "fc1.bias"]))
        y = torch.relu(Fn.linear(y, w[p + "fc2.weight"], w[p + "fc2.bias"]))
        return torch.sigmoid(Fn.linear(y, w[p + "fc3.weight"], w[p + "fc3.bias"]))

    # -- model.py:291-328
    def forward(self, ss_data, ss_labels, query_data, ss_features=None, want=()):
        q = _t(query_data["sk"], self.dtype)
        b = q.shape[0]
        qf = self.embed(q).unsqueeze(1)
        if ss_features is None:
            ss_features = self.embed(_t(ss_data["sk"], self.dtype))
        else:
            ss_features = _t(ss_features, self.dtype)
        labels = torch.as_tensor(np.asarray(ss_labels)).long() if not isinstance(ss_labels, torch.Tensor) else ss_labels.long().cpu()
        out = self.cross_transformer(ss_features, labels, qf, 0, want=("logits", "diffs") + tuple(want))
        logits = out["logits"]
        res = {"logits": logits, "support_features": ss_features}
        if self.cfg.model == "DISC":
            chosen = torch.argmax(logits, dim=1)                          # first max on ties
            feature = out["diffs"][torch.arange(b), chosen]               # model.py:323-324
            res["is_true"] = self.discriminator(feature)
            res["chosen"] = chosen
        if "prototypes" in want:
            res["prototypes"] = out["prototypes"]
        if "probs" in want:
            res["probs"] = out["probs"]
        return res

    # -- batched scoring of B windows against ONE support set
    def score(self, support, labels, query, chunk=1024, ss_features=None):
        """support (1,W,T,90) or ss_features (1,W,T,F); query (B,T,90).
        Returns logits (B,W), is_true (B,1) as numpy."""
        q = _t(query, self.dtype)
        if ss_features is None:
            ss_features = self.embed(_t(support, self.dtype))
        else:
            ss_features = _t(ss_features, self.dtype)
        lab = torch.as_tensor(np.asarray(labels)).long()[:1]
        lo, it = [], []
        with torch.no_grad():
            for s in range(0, q.shape[0], chunk):
                qq = q[s:s + chunk]
                r = self.forward(None, lab, {"sk": qq}, ss_features=ss_features.expand(qq.shape[0], -1, -1, -1))
                lo.append(r["logits"])
                if "is_true" in r:
                    it.append(r["is_true"])
        logits = torch.cat(lo).numpy()
        is_true = torch.cat(it).numpy() if it else None
        return logits, is_true


class ActionRecognizerOracle:
    """Restatement of the stateful wrapper (modules/ar/ar.py:10-96) around TrxOracle."""

    def __init__(self, cfg: Cfg, state_dict: dict):
        self.model = TrxOracle(cfg, state_dict)
        self.support_set = OrderedDict()
        self.requires_focus = {}
        self.previous_frames = []
        self.seq_len = cfg.seq_len
        self.way = cfg.way

    def inference(self, data):                                  # ar.py:30-84
        if data is None or len(data) == 0:
            return {}, 0, {}
        if len(self.support_set) == 0:
            return {}, 0, {}
        frame = torch.as_tensor(np.asarray(data["sk"], dtype=np.float32))
        self.previous_frames.append(frame)
        if len(self.previous_frames) < self.seq_len:
            return {}, 0, {}
        elif len(self.previous_frames) == self.seq_len + 1:
            self.previous_frames = self.previous_frames[1:]
        q = torch.stack(self.previous_frames).unsqueeze(0)
        n = len(self.support_set)
        labels = torch.arange(n).unsqueeze(0)
        ss, ss_f = None, None
        if all("features" in v for v in self.support_set.values()):
            ss_f = torch.stack([v["features"] for v in self.support_set.values()])
            pad = torch.zeros_like(ss_f[0]).unsqueeze(0)
            while ss_f.shape[0] < self.way:
                ss_f = torch.cat((ss_f, pad), dim=0)
            ss_f = ss_f.unsqueeze(0)
        else:
            ss = {"sk": torch.stack([v["poses"] for v in self.support_set.values()]).unsqueeze(0)}
        with torch.no_grad():
            out = self.model.forward(ss, labels, {"sk": q}, ss_features=ss_f)
        if ss_f is None:
            for i, s in enumerate(self.support_set.keys()):
                self.support_set[s]["features"] = out["support_features"][0][i]
        fs = torch.softmax(out["logits"].squeeze(0), dim=0).numpy()
        os_ = out["is_true"].squeeze(0).numpy()
        results = {k: fs[i] for i, k in enumerate(self.support_set.keys())}
        return results, os_, self.requires_focus

    def remove(self, flag):                                     # ar.py:86-92
        if flag in self.support_set:
            self.support_set.pop(flag)
            self.requires_focus.pop(flag)
            return True
        return False

    def train(self, inp):                                       # ar.py:94-96
        self.support_set[inp["flag"]] = {c: torch.as_tensor(np.asarray(inp["data"][c], dtype=np.float32))
                                         for c in inp["data"].keys()}
        self.requires_focus[inp["flag"]] = inp["requires_focus"]
