"""Deterministic synthetic weights and inputs (numpy PCG64, platform independent).

Test infrastructure (see oracle/__init__.py).  The reference has no golden
vectors and its checkpoints are not in the repository, so
weights are random-init.  They are produced here with numpy rather than with
``torch.manual_seed`` so that the golden generator (which loads them into the
unmodified reference with ``load_state_dict``), the oracle, the tests and
``bench.py`` all see bit-identical values on any machine.

The distribution mirrors PyTorch's default ``nn.Linear`` init (Kaiming uniform
with a=sqrt(5): U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias) and
``nn.LayerNorm`` (gamma=1, beta=0) that the reference's modules use
(/root/reference/modules/ar/utils/model.py:41-46,170-172,186-191).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from itertools import combinations

import numpy as np


@dataclass
class Cfg:
    """Mirror of the TRXConfig fields the path reads (utils/params.py:50-95)."""
    way: int = 5
    seq_len: int = 16
    n_joints: int = 30
    trans_linear_in_dim: int = 256
    trans_linear_out_dim: int = 128
    temp_set: list = field(default_factory=lambda: [2])
    model: str = "DISC"
    input_type: str = "skeleton"
    trans_dropout: float = 0.0
    device: str = "cpu"
    num_gpus: int = 1
    shot: int = 1


def _linear(rng, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)
    w = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return w, b


def positional_encoding(max_len: int, d_model: int, scale: float = 0.1) -> np.ndarray:
    """model.py:18-23, evaluated with torch's fp32 ops order (exp/sin/cos in fp32)."""
    import torch
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term) * scale
    pe[:, 1::2] = torch.cos(position * div_term) * scale
    return pe.unsqueeze(0).numpy()


def make_state_dict(cfg: Cfg, seed: int = 0, affine_ln: bool = False,
                    include_post_resnet: bool = False) -> dict:
    """Reference state_dict schema as numpy arrays."""
    rng = np.random.default_rng(seed)
    J3 = cfg.n_joints * 3
    F, D, T = cfg.trans_linear_in_dim, cfg.trans_linear_out_dim, cfg.seq_len
    sd = {}
    sd["features_extractor.sk.fc1.weight"], sd["features_extractor.sk.fc1.bias"] = _linear(rng, 2 * J3, J3)
    sd["features_extractor.sk.fc2.weight"], sd["features_extractor.sk.fc2.bias"] = _linear(rng, 256, 2 * J3)
    for i, c in enumerate(cfg.temp_set):
        p = f"transformers.{i}."
        sd[p + "pe.pe"] = positional_encoding(int(T * 1.5), F)
        sd[p + "k_linear.weight"], sd[p + "k_linear.bias"] = _linear(rng, D, F * c)
        sd[p + "v_linear.weight"], sd[p + "v_linear.bias"] = _linear(rng, D, F * c)
        if affine_ln:
            sd[p + "norm_k.weight"] = rng.uniform(0.5, 1.5, size=(D,)).astype(np.float32)
            sd[p + "norm_k.bias"] = rng.uniform(-0.2, 0.2, size=(D,)).astype(np.float32)
        else:
            sd[p + "norm_k.weight"] = np.ones((D,), np.float32)
            sd[p + "norm_k.bias"] = np.zeros((D,), np.float32)
    if cfg.model == "DISC":
        n2 = (T - 1) * T // 2
        p = "discriminator."
        sd[p + "dimensionality_reduction.weight"], sd[p + "dimensionality_reduction.bias"] = _linear(rng, T, D)
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = _linear(rng, 256, n2 * T)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = _linear(rng, 64, 256)
        sd[p + "fc3.weight"], sd[p + "fc3.bias"] = _linear(rng, 1, 64)
    if include_post_resnet:
        sd["post_resnet.l1.weight"], sd["post_resnet.l1.bias"] = _linear(rng, 256, 2048)
    return sd


def make_episode(cfg: Cfg, B: int, seed: int = 1, kind: str = "structured",
                 way: int | None = None):
    """Synthetic inputs of SURVEY.md section 8(d).

    structured: support = 0.17*N(0,1) (1,W,T,90); planted class per window;
                query = support[cls] + 0.05*N(0,1).
    iid:        query = 0.17*N(0,1).
    Returns (support (1,W,T,90) f32, labels (1,W) i32, query (B,T,90) f32, planted (B,) i64).
    """
    W = cfg.way if way is None else way
    T, J3 = cfg.seq_len, cfg.n_joints * 3
    rng = np.random.default_rng(seed)
    support = (0.17 * rng.standard_normal((1, W, T, J3))).astype(np.float32)
    planted = rng.integers(0, W, size=(B,))
    noise = rng.standard_normal((B, T, J3)).astype(np.float32)
    if kind == "structured":
        query = support[0, planted] + np.float32(0.05) * noise
    elif kind == "iid":
        query = np.float32(0.17) * noise
    else:
        raise ValueError(kind)
    labels = np.arange(W, dtype=np.int32)[None]
    return support, labels, query.astype(np.float32), planted.astype(np.int64)


def tuple_table(T: int, c: int) -> np.ndarray:
    """itertools.combinations(range(T), c), lexicographic (model.py:52-54)."""
    return np.array(list(combinations(range(T), c)), dtype=np.int64).reshape(-1, c)


def make_heatmaps(B: int, seed: int = 2, n_joints: int = 32, depth: int = 8,
                  spike: float = 12.0) -> np.ndarray:
    """MetrABS-style head output (B,8,8,32+8*32) f32 (hpe.py:109-112): 3*N(0,1)
    plus a +spike per joint inside the FOV band so most joints pass is_within_fov."""
    rng = np.random.default_rng(seed)
    x = (3.0 * rng.standard_normal((B, 8, 8, n_joints * (1 + depth)))).astype(np.float32)
    h = rng.integers(1, 7, size=(B, n_joints))
    w = rng.integers(1, 7, size=(B, n_joints))
    d = rng.integers(0, depth, size=(B, n_joints))
    b = np.arange(B)[:, None]
    j = np.arange(n_joints)[None, :]
    x[b, h, w, j] += np.float32(spike)                                  # 2-D head, channel j
    x[b, h, w, n_joints + d * n_joints + j] += np.float32(spike)        # 3-D head, channel (d j)
    return x
