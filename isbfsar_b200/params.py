"""Configuration mirror of the reference's utils/params.py:50-95 (TRXConfig) --
the fields the scoring path reads, same names and defaults, without the import-time
print or the dataset/training knobs."""
from __future__ import annotations

input_type = "skeleton"          # utils/params.py:4
seq_len = 16                     # utils/params.py:8 (skeleton)


class TRXConfig(object):
    def __init__(self):
        self.model = "DISC"
        self.input_type = input_type
        self.way = 5
        self.shot = 1
        self.device = "cuda"
        self.skeleton_type = "smpl+head_30"
        self.n_joints = 30
        self.trans_linear_in_dim = 256
        self.trans_linear_out_dim = 128
        self.query_per_class = 1
        self.trans_dropout = 0.0
        self.num_gpus = 1
        self.temp_set = [2]
        self.final_ckpt_path = "modules/ar/modules/raws/DISC.pth"
        self.seq_len = seq_len
