"""Operand-rounding emulation for the tensor-core path (fp32 accumulate), vs the fp64 oracle.
Usage: python tools/precision_study.py   (CPU only; dev tool, not part of the product path)"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_state_dict, make_episode, tuple_table
from oracle.trx_oracle import TrxOracle

def rnd(x, fmt):
    if fmt == "fp32": return x
    if fmt == "fp16": return x.half().float()
    if fmt == "bf16": return x.bfloat16().float()
    if fmt == "tf32":
        i = x.view(torch.int32); i = (i + 0x1000) & ~0x1FFF; return i.view(torch.float32)
    if fmt == "fp16x2":  # hi+lo split: ~22 bits
        hi = x.half().float(); lo = (x - hi).half().float(); return hi + lo
    raise ValueError(fmt)

def mm(a, b, fa, fb):  # a @ b.T with operand rounding, fp32 accumulate (emulated in fp64 then cast)
    return (rnd(a, fa).double() @ rnd(b, fb).double().T).float()

def emulate(cfg, sd, support, query, fmt):
    w = {k: torch.from_numpy(v) for k, v in sd.items()}
    T, F, D = cfg.seq_len, cfg.trans_linear_in_dim, cfg.trans_linear_out_dim
    tup = torch.from_numpy(tuple_table(T, 2))
    def frames(x):
        sh = x.shape[:-1]; x = x.reshape(-1, x.shape[-1])
        h = torch.relu(mm(x, w["features_extractor.sk.fc1.weight"], fmt["mlp_a"], fmt["mlp_w"]) + w["features_extractor.sk.fc1.bias"])
        f = torch.relu(mm(h, w["features_extractor.sk.fc2.weight"], fmt["mlp_a"], fmt["mlp_w"]) + w["features_extractor.sk.fc2.bias"])
        f = f.reshape(*sh, -1) + w["transformers.0.pe.pe"][0, :T]
        f2 = f.reshape(-1, F)
        Wk, Wv = w["transformers.0.k_linear.weight"], w["transformers.0.v_linear.weight"]
        gk = [mm(f2, Wk[:, p*F:(p+1)*F], fmt["proj_a"], fmt["proj_w"]).reshape(*sh, D) for p in range(2)]
        gv = [mm(f2, Wv[:, p*F:(p+1)*F], fmt["proj_a"], fmt["proj_w"]).reshape(*sh, D) for p in range(2)]
        k = gk[0][..., tup[:, 0], :] + gk[1][..., tup[:, 1], :] + w["transformers.0.k_linear.bias"]
        v = gv[0][..., tup[:, 0], :] + gv[1][..., tup[:, 1], :] + w["transformers.0.v_linear.bias"]
        k = torch.nn.functional.layer_norm(k, (D,), w["transformers.0.norm_k.weight"], w["transformers.0.norm_k.bias"], 1e-5)
        return k, v
    sk, sv = frames(torch.from_numpy(support)[0])      # (W,N,D)
    qk, qv = frames(torch.from_numpy(query))           # (B,N,D)
    B, N = qk.shape[0], qk.shape[1]
    logits = []; diffs = []
    for c in range(sk.shape[0]):
        S = torch.einsum("bqd,sd->bqs", rnd(qk, fmt["qk"]).double(), rnd(sk[c], fmt["qk"]).double()).float() / math.sqrt(D)
        E = torch.exp(S); P = E / E.sum(dim=1, keepdim=True)
        proto = torch.einsum("bqs,sd->bqd", rnd(P, fmt["pv"]).double(), rnd(sv[c], fmt["pv"]).double()).float()
        diff = qv - proto
        logits.append(-(diff.double() ** 2).sum(dim=(1, 2)).float() / N); diffs.append(diff)
    logits = torch.stack(logits, 1)
    ch = logits.argmax(1)
    feat = torch.stack(diffs, 1)[torch.arange(B), ch]
    y = mm(feat.reshape(-1, D), w["discriminator.dimensionality_reduction.weight"], fmt["disc"], fmt["disc"]).reshape(B, -1) \
        + w["discriminator.dimensionality_reduction.bias"].repeat(N)
    y = torch.relu(mm(y, w["discriminator.fc1.weight"], fmt["disc"], fmt["disc"]) + w["discriminator.fc1.bias"])
    y = torch.relu(y @ w["discriminator.fc2.weight"].T + w["discriminator.fc2.bias"])
    it = torch.sigmoid(y @ w["discriminator.fc3.weight"].T + w["discriminator.fc3.bias"])
    return logits.numpy(), it.numpy()

def main():
    B = int(os.environ.get("B", 512))
    # CFG=t32: BASELINE cfg4 pairs (20-way, T=32, N=496) -- does the same operand scheme hold for the N > 128 tiling?
    cfg = Cfg(way=20, seq_len=32, temp_set=[2, 3]) if os.environ.get("CFG") == "t32" else Cfg()
    only = os.environ.get("SCHEMES")
    schemes = {
      "all-fp32":   dict(mlp_a="fp32", mlp_w="fp32", proj_a="fp32", proj_w="fp32", qk="fp32", pv="fp32", disc="fp32"),
      "all-bf16":   dict(mlp_a="bf16", mlp_w="bf16", proj_a="bf16", proj_w="bf16", qk="bf16", pv="bf16", disc="bf16"),
      "all-tf32":   dict(mlp_a="tf32", mlp_w="tf32", proj_a="tf32", proj_w="tf32", qk="tf32", pv="tf32", disc="tf32"),
      "all-fp16":   dict(mlp_a="fp16", mlp_w="fp16", proj_a="fp16", proj_w="fp16", qk="fp16", pv="fp16", disc="fp16"),
      "attn-fp16, frames-fp32": dict(mlp_a="fp32", mlp_w="fp32", proj_a="fp32", proj_w="fp32", qk="fp16", pv="fp16", disc="fp16"),
      "attn-fp16, frames-fp16x2": dict(mlp_a="fp16x2", mlp_w="fp16x2", proj_a="fp16x2", proj_w="fp16x2", qk="fp16", pv="fp16", disc="fp16"),
      "attn-bf16, frames-fp32": dict(mlp_a="fp32", mlp_w="fp32", proj_a="fp32", proj_w="fp32", qk="bf16", pv="bf16", disc="bf16"),
      "all-fp16 but PV bf16":   dict(mlp_a="fp16", mlp_w="fp16", proj_a="fp16", proj_w="fp16", qk="fp16", pv="bf16", disc="fp16"),
      "attn-fp16, mlp-fp16 proj-fp32": dict(mlp_a="fp16", mlp_w="fp16", proj_a="fp32", proj_w="fp32", qk="fp16", pv="fp16", disc="fp16"),
      "attn-fp16, mlp-fp32 proj-fp16": dict(mlp_a="fp32", mlp_w="fp32", proj_a="fp16", proj_w="fp16", qk="fp16", pv="fp16", disc="fp16"),
    }
    for affine, wseed in [(False, 0), (True, 5)]:
      sd = make_state_dict(cfg, wseed, affine_ln=affine)
      for kind in ["structured", "iid"]:
        support, labels, query, planted = make_episode(cfg, B, 1, kind)
        o64 = TrxOracle(cfg, sd, dtype=torch.float64)
        l64, t64 = o64.score(support, labels, query)
        o32 = TrxOracle(cfg, sd)
        l32, t32 = o32.score(support, labels, query)
        srt = np.sort(l64, 1); margin = (srt[:, -1] - srt[:, -2]) / np.abs(srt[:, -1])
        print(f"== affine={affine} kind={kind} B={B}  margin p0.1={np.quantile(margin,0.001):.2e} min={margin.min():.2e}; oracle32 vs 64: {np.abs(l32/l64-1).max():.2e}")
        for name, fmt in schemes.items():
            if only and name not in only.split(";"):
                continue
            l, t = emulate(cfg, sd, support, query, fmt)
            rel = np.abs(l / l64 - 1)
            relw = rel[np.arange(B), l64.argmax(1)]
            print(f"  {name:34s} logits rel max {rel.max():.2e} med {np.median(rel):.2e} (winner max {relw.max():.2e}) | is_true rel max {np.abs(t/t64-1).max():.2e} | argmax agree {np.mean(l.argmax(1)==l64.argmax(1)):.4f} decision agree {np.mean((t>0.5)==(t64>0.5)):.4f}")
main()
