"""Host wrapper of the heatmap decoder kernel (arx_decode_heatmaps): MetrABS-style volumetric
heatmaps -> root-centred 30-joint poses that feed the AR scorer (reference modules/hpe/hpe.py:108-169,
main.py:103-105).  The TensorRT engines that produce the heatmaps are outside the path; callers pass
the head output `(B,8,8,288)`."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class HeatmapDecoder:
    def __init__(self, model, expand_joints, indices=None, new_K=None, homo_inv=None):
        """model: a TRXOS on a CUDA device (provides the native handle).
        expand_joints: (32,122) `assets/32_to_122.npy`-style matrix, or already column-selected (32,n_out);
        indices: joint subset (`skeleton_types[...]['indices']`, hpe.py:164) applied to its columns;
        new_K (3,3), homo_inv (3,3)|(1,3,3): outputs of misc.py:homography for the current bounding box."""
        self.model = model
        E = np.asarray(expand_joints, dtype=np.float32)
        if indices is not None:
            E = E[:, np.asarray(indices, dtype=np.int64)]      # selecting columns commutes with the matmul
        self.n_out = E.shape[1]
        self.expand = torch.from_numpy(np.ascontiguousarray(E)).to(model._device())
        self.set_camera(new_K, homo_inv)

    def set_camera(self, new_K, homo_inv):
        self.new_K = None if new_K is None else np.ascontiguousarray(np.asarray(new_K, dtype=np.float32).reshape(3, 3))
        self.homo_inv = None if homo_inv is None else np.ascontiguousarray(np.asarray(homo_inv, dtype=np.float32).reshape(3, 3))

    def decode(self, logits):
        """logits (B,8,8,288) float32 CUDA -> poses (B, 3*n_out) float32, valid (B,) bool."""
        m = self.model
        h = m._ensure()
        dev = m._device()
        x = logits.detach().to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 4 or tuple(x.shape[1:]) != (8, 8, 288):
            raise ValueError("logits must be (B,8,8,288)")
        if self.new_K is None or self.homo_inv is None:
            raise RuntimeError("decode: camera not set")
        B = x.shape[0]
        poses = torch.empty((B, self.n_out * 3), dtype=torch.float32, device=dev)
        valid = torch.empty((B,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().arx_decode_heatmaps(
                h, C.c_void_p(x.data_ptr()), B, C.c_void_p(self.expand.data_ptr()), self.n_out,
                self.new_K.ctypes.data_as(C.c_void_p), self.homo_inv.ctypes.data_as(C.c_void_p),
                C.c_void_p(poses.data_ptr()), C.c_void_p(valid.data_ptr()), m._stream()), h, "arx_decode_heatmaps")
        return poses, valid.bool()
