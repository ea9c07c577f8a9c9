"""The CPU oracle against the golden vectors frozen from the unmodified reference
(oracle/gen_golden.py).  fp32 restatement vs fp32 reference: summation-order noise only."""
import os
from itertools import combinations

import numpy as np
import pytest
import torch

from oracle.synth import Cfg, make_episode, make_state_dict, make_heatmaps, tuple_table
from oracle.trx_oracle import ActionRecognizerOracle, TrxOracle
from oracle import decode_oracle as D

TRX_CASES = [
    ("cfg1_w5_t16_structured", Cfg()),
    ("cfg1_w5_t16_iid", Cfg()),
    ("cfg1_w5_t16_affine", Cfg()),
    ("w3_t16_structured", Cfg()),
    ("cfg3_w60_t16", Cfg(way=60)),
    ("cfg4_w20_t32_pairs", Cfg(way=20, seq_len=32, temp_set=[2, 3])),
]


def load_case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    way, T, B, wseed, iseed, affine = [int(x) for x in g["meta"][:6]]
    return g, way, T, B, wseed, iseed, bool(affine), str(g["kind"])


@pytest.mark.parametrize("name,cfg", TRX_CASES)
def test_trx_oracle_matches_reference(golden_dir, name, cfg):
    g, way, T, B, wseed, iseed, affine, kind = load_case(golden_dir, name)
    sd = make_state_dict(cfg, wseed, affine_ln=affine)
    support, labels, query, planted = make_episode(cfg, B, iseed, kind, way=way)
    o = TrxOracle(cfg, sd)
    logits, is_true = o.score(support, labels, query, chunk=64)
    np.testing.assert_allclose(logits, g["logits"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(is_true, g["is_true"], rtol=2e-5, atol=1e-6)
    assert np.array_equal(logits.argmax(1), g["logits"].argmax(1))
    assert np.array_equal(np.stack([t.numpy() for t in [o.tuples[0]]])[0], g["tuples"])
    ssf = o.embed(torch.from_numpy(support)).numpy()
    np.testing.assert_allclose(ssf[0, :2], g["support_features"], rtol=1e-5, atol=1e-6)
    # uncached path (support poses given)
    r = o.forward({"sk": support}, labels, {"sk": query[:1]})
    np.testing.assert_allclose(r["logits"].numpy(), g["logits_uncached"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(r["is_true"].numpy(), g["is_true_uncached"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("name,cfg", [("w5_t16_triples", Cfg(way=5, seq_len=16, temp_set=[2, 3])),
                                      ("cfg4_w20_t32_triples", Cfg(way=20, seq_len=32, temp_set=[2, 3]))])
def test_triple_transformer_matches_reference(golden_dir, name, cfg):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    way, T, B, wseed, iseed, ti = [int(x) for x in g["meta"][:6]]
    if T == 32:
        way_run = 2          # full 20-way N=4960 is ~98 MB/score matrix; check the first classes only
    else:
        way_run = way
    sd = make_state_dict(cfg, wseed)
    support, labels, query, _ = make_episode(cfg, B, iseed, "structured")
    o = TrxOracle(cfg, sd)
    with torch.no_grad():
        ssf = o.embed(torch.from_numpy(support))
        qf = o.embed(torch.from_numpy(query)).unsqueeze(1)
        out = o.cross_transformer(ssf.expand(B, -1, -1, -1), torch.from_numpy(labels[:, :way_run]).long(), qf, ti)
    np.testing.assert_allclose(out["logits"].numpy(), g["logits"][:, :way_run], rtol=2e-5, atol=1e-6)
    assert np.array_equal(o.tuples[ti].numpy(), g["tuples"])


@pytest.mark.parametrize("T,c", [(4, 2), (8, 2), (16, 2), (32, 2), (8, 3), (16, 3), (32, 3)])
def test_tuple_table_is_itertools(T, c):
    assert np.array_equal(tuple_table(T, c), np.array(list(combinations(range(T), c))))


def test_action_recognizer_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ar_stream.npz"))
    cfg = Cfg()
    ar = ActionRecognizerOracle(cfg, make_state_dict(cfg, 0))
    rng = np.random.default_rng(7)
    poses = (0.17 * rng.standard_normal((3, 16, 90))).astype(np.float32)
    frames = (0.17 * rng.standard_normal((23, 90))).astype(np.float32)
    frames[5:21] = poses[1] + 0.05 * rng.standard_normal((16, 90)).astype(np.float32)
    assert ar.inference(None) == ({}, 0, {})
    assert ar.inference({"sk": frames[0]}) == ({}, 0, {})      # empty support set; frame NOT recorded (ar.py:37-38)
    for i, n in enumerate(["wave", "clap", "kick"]):
        ar.train({"flag": n, "data": {"poses": poses[i]}, "requires_focus": bool(i % 2)})
    probs, os_, empties = [], [], 0
    for f in range(20):
        res, o, rf = ar.inference({"sk": frames[f]})
        if len(res) == 0:
            empties += 1
            continue
        assert list(res.keys()) == ["wave", "clap", "kick"]
        probs.append([res[k] for k in res])
        os_.append(float(np.asarray(o).reshape(-1)[0]))
    assert empties == int(g["empties"])
    np.testing.assert_allclose(np.array(probs), g["probs"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(np.array(os_), g["open_set"], rtol=1e-5, atol=1e-6)
    assert ar.remove("clap") and not ar.remove("nope")
    probs2 = []
    for f in range(20, 23):
        res, o, rf = ar.inference({"sk": frames[f]})
        probs2.append([res[k] for k in ["wave", "kick"]])
    np.testing.assert_allclose(np.array(probs2), g["probs_after_remove"], rtol=1e-4, atol=1e-6)


def test_decode_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_64.npz"))
    hm = make_heatmaps(64, seed=2)
    p2, p3 = D.soft_argmax(hm)
    np.testing.assert_allclose(p2, g["pred2d"], rtol=1e-12)
    np.testing.assert_allclose(p3, g["pred3d"], rtol=1e-12)
    K = D.realsense_K()
    nk, R = D.homography(100, 300, 50, 450, K, 256)
    assert np.array_equal(nk, g["new_K"]) and np.array_equal(R, g["homo_inv"])
    # column-selected (32,30) remap == full (32,122) remap then select
    E30 = g["expand30"]
    poses, valid = D.decode_frames(hm, E30, np.arange(30), nk, R)
    assert valid.all()
    np.testing.assert_allclose(poses, g["poses"], rtol=1e-10, atol=1e-12)
