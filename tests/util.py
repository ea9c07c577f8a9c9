import numpy as np
import torch

from oracle.synth import Cfg, make_state_dict


class Args:
    """TRXConfig-like object built from an oracle Cfg."""

    def __init__(self, cfg: Cfg):
        self.model = cfg.model
        self.input_type = "skeleton"
        self.way = cfg.way
        self.shot = 1
        self.device = "cuda"
        self.n_joints = cfg.n_joints
        self.trans_linear_in_dim = cfg.trans_linear_in_dim
        self.trans_linear_out_dim = cfg.trans_linear_out_dim
        self.trans_dropout = 0.0
        self.num_gpus = 1
        self.temp_set = list(cfg.temp_set)
        self.seq_len = cfg.seq_len
        self.final_ckpt_path = None


def torch_sd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def make_model(cfg: Cfg, wseed=0, affine=False, force_path=0, max_chunk=0):
    from isbfsar_b200 import TRXOS
    m = TRXOS(Args(cfg))
    sd = make_state_dict(cfg, wseed, affine_ln=affine)
    missing, unexpected = m.load_state_dict(torch_sd(sd), strict=False)
    assert not unexpected and all(k.startswith("post_resnet.") for k in missing), (missing, unexpected)
    m.force_path = force_path
    m.max_chunk = max_chunk
    return m.cuda(), sd


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
