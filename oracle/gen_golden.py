"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run here only (needs /root/reference, read-only):
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden
The GPU box has no /root/reference; tests there read the committed fixtures.

What is pinned:
  * TRXOS.forward  (modules/ar/utils/model.py:291-328) on synth weights/inputs
  * TemporalCrossTransformer(args, 3).forward (model.py:59-148) for triples
  * ActionRecognizer.inference/train/remove (modules/ar/ar.py:30-96), run on CPU
    by neutralising `.cuda()` (the only CUDA dependence of that file)
  * modules/hpe/utils/misc.py helpers used by the decode (is_within_fov,
    reconstruct_absolute, homography)
Inputs/weights are NOT stored: they are regenerated from oracle/synth.py seeds.
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import numpy as np

sys.dont_write_bytecode = True
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_model(cfg, sd):
    import torch
    sys.path.insert(0, REF)
    from utils.params import TRXConfig
    from modules.ar.utils.model import TRXOS
    a = TRXConfig()
    a.device = "cpu"
    a.way, a.seq_len, a.temp_set = cfg.way, cfg.seq_len, list(cfg.temp_set)
    m = TRXOS(a).eval()
    full = {k: v.clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        assert tuple(full[k].shape) == tuple(v.shape), (k, full[k].shape, v.shape)
        if k.endswith("pe.pe"):
            assert np.array_equal(full[k].numpy(), v), "positional encoding restatement differs"
        full[k] = torch.from_numpy(np.asarray(v))
    m.load_state_dict(full)
    return m


def trx_case(name, cfg, B, wseed, iseed, kind, affine=False, way=None):
    import torch
    from oracle.synth import make_state_dict, make_episode
    sd = make_state_dict(cfg, wseed, affine_ln=affine)
    m = _ref_model(cfg, sd)
    support, labels, query, planted = make_episode(cfg, B, iseed, kind, way=way)
    with torch.no_grad():
        ssf = m.features_extractor["sk"](torch.from_numpy(support))
        lo, it = [], []
        ch = 256 if support.shape[1] <= 5 else 16
        for s in range(0, B, ch):
            q = torch.from_numpy(query[s:s + ch])
            r = m(None, torch.from_numpy(labels), {"sk": q}, ss_features=ssf.expand(q.shape[0], -1, -1, -1))
            lo.append(r["logits"].numpy())
            it.append(r["is_true"].numpy())
        # uncached path on the first window (ss_data given, model.py:307-317)
        r0 = m({"sk": torch.from_numpy(support)}, torch.from_numpy(labels), {"sk": torch.from_numpy(query[:1])})
    tuples = np.stack([t.numpy() for t in m.transformers[0].tuples])
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        logits=np.concatenate(lo), is_true=np.concatenate(it), support_features=ssf.numpy()[0, :2],
        logits_uncached=r0["logits"].numpy(), is_true_uncached=r0["is_true"].numpy(),
        tuples=tuples, planted=planted,
        meta=np.array([cfg.way if way is None else way, cfg.seq_len, B, wseed, iseed, int(affine)] + list(cfg.temp_set)),
        kind=np.array(kind))
    print(name, "logits", np.concatenate(lo).shape, "argmax==planted", float((np.concatenate(lo).argmax(1) == planted).mean()))


def triple_case(name, cfg, B, wseed, iseed, ti):
    """TemporalCrossTransformer for temp_set[ti] called directly (the reference forward never uses it)."""
    import torch
    from oracle.synth import make_state_dict, make_episode
    sd = make_state_dict(cfg, wseed)
    m = _ref_model(cfg, sd)
    support, labels, query, planted = make_episode(cfg, B, iseed, "structured")
    with torch.no_grad():
        ssf = m.features_extractor["sk"](torch.from_numpy(support))
        qf = m.features_extractor["sk"](torch.from_numpy(query)).unsqueeze(1)
        out = m.transformers[ti](ssf.expand(B, -1, -1, -1), torch.from_numpy(labels), qf)
    tuples = np.stack([t.numpy() for t in m.transformers[ti].tuples])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), logits=out["logits"].numpy(), tuples=tuples,
                        meta=np.array([cfg.way, cfg.seq_len, B, wseed, iseed, ti] + list(cfg.temp_set)))
    print(name, out["logits"].shape)


def ar_case(name):
    """Drive the real modules/ar/ar.py logic on CPU: 3 classes added, 20 frames, one removed, 3 more frames."""
    import torch
    from oracle.synth import Cfg, make_state_dict
    cfg = Cfg()
    sd = make_state_dict(cfg, 0)
    m = _ref_model(cfg, sd)
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # ar.py:41,51,95 -- the only CUDA dependence
    from modules.ar.ar import ActionRecognizer
    ar = ActionRecognizer.__new__(ActionRecognizer)         # skip the checkpoint-loading ctor (ar.py:15-21)
    ar.input_type, ar.device, ar.ar = "skeleton", "cpu", m
    ar.support_set, ar.requires_focus, ar.previous_frames = OrderedDict(), {}, []
    ar.seq_len, ar.way, ar.n_joints = 16, 5, 30
    rng = np.random.default_rng(7)
    poses = (0.17 * rng.standard_normal((3, 16, 90))).astype(np.float32)
    frames = (0.17 * rng.standard_normal((23, 90))).astype(np.float32)
    frames[5:21] = poses[1] + 0.05 * rng.standard_normal((16, 90)).astype(np.float32)
    probs, os_ = [], []
    empties = 0
    for i, nme in enumerate(["wave", "clap", "kick"]):
        ar.train({"flag": nme, "data": {"poses": poses[i]}, "requires_focus": bool(i % 2)})
    for f in range(20):
        res, o, rf = ar.inference({"sk": frames[f]})
        if len(res) == 0:
            empties += 1
            continue
        probs.append([res[k] for k in ["wave", "clap", "kick"]])
        os_.append(np.asarray(o).reshape(-1)[0])
    feats = np.stack([ar.support_set[k]["features"].numpy() for k in ["wave", "clap", "kick"]])
    assert ar.remove("clap") and not ar.remove("nope")
    probs2, os2 = [], []
    for f in range(20, 23):
        res, o, rf = ar.inference({"sk": frames[f]})
        probs2.append([res[k] for k in ["wave", "kick"]])
        os2.append(np.asarray(o).reshape(-1)[0])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), probs=np.array(probs, np.float32), open_set=np.array(os_, np.float32),
                        probs_after_remove=np.array(probs2, np.float32), open_set_after_remove=np.array(os2, np.float32),
                        empties=np.array(empties), features=feats[:, :2])
    print(name, np.array(probs).shape, "empties", empties)


def decode_case(name):
    """Pin oracle/decode_oracle.py helpers against modules/hpe/utils/misc.py, and freeze a decode."""
    sys.path.insert(0, REF)
    import pickle
    from modules.hpe.utils import misc as M
    from oracle import decode_oracle as D
    from oracle.synth import make_heatmaps
    K = D.realsense_K()
    nk_ref, R_ref = M.homography(100, 300, 50, 450, K, 256)
    nk, R = D.homography(100, 300, 50, 450, K, 256)
    assert np.array_equal(nk_ref, nk) and np.array_equal(R_ref, R)
    hm = make_heatmaps(64, seed=2)
    p2, p3 = D.soft_argmax(hm)
    ref_abs = []
    for f in range(hm.shape[0]):
        fov = M.is_within_fov(p2[f:f + 1])
        assert np.array_equal(fov[0], D.is_within_fov(p2[f]))
        a = M.reconstruct_absolute(p2[f:f + 1], p3[f:f + 1], nk[None, ...], fov, weak_perspective=False)
        b = D.reconstruct_absolute_one(p2[f], p3[f], nk, fov[0])
        assert np.allclose(a[0], b, rtol=1e-12, atol=1e-12), np.abs(a[0] - b).max()
        ref_abs.append(a[0])
    E = np.load(os.path.join(REF, "assets/32_to_122.npy"))
    st = pickle.load(open(os.path.join(REF, "assets/skeleton_types.pkl"), "rb"))
    idx = np.array([int(i) for i in st["smpl+head_30"]["indices"]])
    # inline transcription of hpe.py:156-169 + main.py:103-105 on top of the reference's own misc.py results
    poses_ref = []
    for f in range(hm.shape[0]):
        p = ref_abs[f][None] @ R                       # hpe.py:159 (homo_inv is (1,3,3))
        p = (p.swapaxes(1, 2) @ E).swapaxes(1, 2)      # hpe.py:162
        p = p[:, idx][0]                               # hpe.py:164,169
        p = p - p[0, :]                                # main.py:103
        poses_ref.append(p.reshape(-1))
    poses_ref = np.stack(poses_ref)
    poses, valid = D.decode_frames(hm, E, idx, nk, R)
    assert valid.all(), valid.mean()
    assert np.allclose(poses, poses_ref, rtol=1e-12, atol=1e-13), np.abs(poses - poses_ref).max()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pred2d=p2, pred3d=p3, poses=poses_ref, valid=valid,
                        new_K=nk, homo_inv=R, expand30=E[:, idx].astype(np.float32), indices=idx)
    # decoder constants the product needs at run time (32x30 fp32 = column-selected 32_to_122.npy)
    print(name, poses_ref.shape, "valid", valid.mean())


def tta_case(name):
    """Pin the test-time-augmentation helpers against modules/hpe/utils/misc.py (get_augmentations, rotation_mat_zaxis) and
    freeze a per-crop decode: crop k is decoded by the reference's own reconstruct_absolute with crop k's camera."""
    sys.path.insert(0, REF)
    import pickle
    from modules.hpe.utils import misc as M
    from oracle import decode_oracle as D
    from oracle.synth import make_heatmaps
    n = 5
    ref = M.get_augmentations(n)
    got = D.get_augmentations(n)
    for a, b in zip(ref, got):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    K = D.realsense_K()
    nk, R0 = M.homography(100, 300, 50, 450, K, 256)
    # hpe.py:88-93 transcribed on the reference's own outputs
    new_K = np.tile(nk, (n, 1, 1))
    for k in range(n):
        new_K[k, :2, :2] *= ref[3][k]
    homo_inv = ref[1] @ np.tile(R0[0], (n, 1, 1))
    Ks, Rs, flip = D.tta_cameras(nk, R0, n)
    assert np.allclose(Ks, new_K, rtol=0, atol=0) and np.allclose(Rs, homo_inv, rtol=0, atol=0)
    hm = make_heatmaps(n, seed=4)
    p2, p3 = D.soft_argmax(hm)
    E = np.load(os.path.join(REF, "assets/32_to_122.npy"))
    st = pickle.load(open(os.path.join(REF, "assets/skeleton_types.pkl"), "rb"))
    idx = np.array([int(i) for i in st["smpl+head_30"]["indices"]])
    poses_ref = []
    for f in range(n):
        fov = M.is_within_fov(p2[f:f + 1])
        a = M.reconstruct_absolute(p2[f:f + 1], p3[f:f + 1], new_K[f][None, ...], fov, weak_perspective=False)
        p = a @ homo_inv[f]
        p = (p.swapaxes(1, 2) @ E).swapaxes(1, 2)[:, idx][0]
        poses_ref.append((p - p[0, :]).reshape(-1))
    poses_ref = np.stack(poses_ref)
    poses, valid = D.decode_frames_cams(hm, E, idx, Ks, Rs)
    assert valid.all() and np.allclose(poses, poses_ref, rtol=1e-12, atol=1e-13)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), poses=poses_ref, new_K=new_K, homo_inv=homo_inv, flip=ref[0], rotflip=ref[1],
                        gammas=ref[2], scales=ref[3], base_K=nk, base_R=R0)
    print(name, poses_ref.shape)


def support_set_fixture():
    """The reference's own saved support set (assets/saved/support_set.pkl + requires_focus.pkl, written by
    main.py:321-326 `save`): copied byte for byte as the interchange fixture (a data asset, not source).  It is an
    OrderedDict name -> {"poses": (16,90), "features": (16,256)} of CUDA tensors, exactly what `load` (main.py:328-333)
    assigns to `ActionRecognizer.support_set`."""
    import shutil
    for src, dst in [("assets/saved/support_set.pkl", "ref_support_set.pkl"), ("assets/saved/requires_focus.pkl", "ref_requires_focus.pkl")]:
        shutil.copyfile(os.path.join(REF, src), os.path.join(OUT, dst))
    print("support-set fixture copied")


def main():
    os.makedirs(OUT, exist_ok=True)
    support_set_fixture()
    tta_case("tta_5")
    from oracle.synth import Cfg
    trx_case("cfg1_w5_t16_structured", Cfg(), 64, 0, 1, "structured")
    trx_case("cfg1_w5_t16_iid", Cfg(), 64, 0, 3, "iid")
    trx_case("cfg1_w5_t16_affine", Cfg(), 32, 5, 6, "structured", affine=True)
    trx_case("w3_t16_structured", Cfg(), 16, 0, 8, "structured", way=3)
    trx_case("cfg3_w60_t16", Cfg(way=60), 16, 0, 1, "structured")
    trx_case("cfg4_w20_t32_pairs", Cfg(way=20, seq_len=32, temp_set=[2, 3]), 4, 0, 1, "structured")
    triple_case("cfg4_w20_t32_triples", Cfg(way=20, seq_len=32, temp_set=[2, 3]), 1, 0, 1, 1)
    triple_case("w5_t16_triples", Cfg(way=5, seq_len=16, temp_set=[2, 3]), 4, 0, 1, 1)
    ar_case("ar_stream")
    decode_case("decode_64")


if __name__ == "__main__":
    main()
