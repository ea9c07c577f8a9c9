"""Host wrapper of the heatmap decoder kernel (arx_decode_heatmaps): MetrABS-style volumetric
heatmaps -> root-centred 30-joint poses that feed the AR scorer (reference modules/hpe/hpe.py:108-169,
main.py:103-105).  The TensorRT engines that produce the heatmaps are outside the path; callers pass
the head output `(B,8,8,288)`."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def rotation_mat_zaxis(angle):
    """misc.py:299-307"""
    sin, cos = np.sin(angle), np.cos(angle)
    _0, _1 = np.zeros_like(angle), np.ones_like(angle)
    return np.stack([np.stack([cos, -sin, _0], axis=-1), np.stack([sin, cos, _0], axis=-1), np.stack([_0, _0, _1], axis=-1)], axis=-2)


def get_augmentations(num_aug, rot_aug_linspace_noend=True):
    """Test-time-augmentation parameters (misc.py:310-327): -> aug_should_flip, aug_rotflipmat (num_aug,3,3), aug_gammas, aug_scales."""
    aug_gammas = np.linspace(0.6, 1.0, num_aug)
    aug_angle_range = np.float32(np.deg2rad(25))
    if rot_aug_linspace_noend:
        aug_angles = np.linspace(-aug_angle_range, aug_angle_range, num_aug + 1)[:-1]
    else:
        aug_angles = np.linspace(-aug_angle_range, aug_angle_range, num_aug)
    aug_scales = np.concatenate([np.linspace(0.8, 1.0, (num_aug + 1) // 2)[:-1], np.linspace(1.0, 1.1, num_aug - num_aug // 2)], axis=0)
    aug_should_flip = (np.arange(num_aug) - num_aug // 2) % 2 != 0
    aug_flipmat = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
    aug_maybe_flipmat = np.where(aug_should_flip[:, np.newaxis, np.newaxis], aug_flipmat, np.eye(3))
    aug_rotflipmat = aug_maybe_flipmat @ rotation_mat_zaxis(-aug_angles)
    return aug_should_flip, aug_rotflipmat, aug_gammas, aug_scales


def tta_cameras(new_K, homo_inv, num_aug):
    """hpe.py:88-93: the per-augmentation cameras -- intrinsics scaled by aug_scales[k], homography preceded by the k-th
    rotation/flip.  -> new_K (num_aug,3,3), homo_inv (num_aug,3,3), aug_should_flip."""
    flip, rotflip, _, scales = get_augmentations(num_aug)
    K = np.tile(np.asarray(new_K).reshape(3, 3), (num_aug, 1, 1)).astype(np.float64)
    for k in range(num_aug):
        K[k, :2, :2] *= scales[k]
    R = rotflip @ np.tile(np.asarray(homo_inv).reshape(3, 3), (num_aug, 1, 1))
    return K, R, flip


class MetrabsHeads:
    """The heads in front of the decoder: Linear(1280 -> 288) over the (8,8,1280) backbone feature map
    (modules/hpe/setup/4_create_heads_onnx.py:7-16; a TensorRT fp16 engine in the reference, hpe.py:106) as a tcgen05 GEMM."""

    def __init__(self, model, weight, bias):
        """weight (288,1280) (`nn.Linear.weight` layout), bias (288,)."""
        self.model = model
        h = model._ensure()
        w = np.ascontiguousarray(np.asarray(weight, dtype=np.float32))
        b = np.ascontiguousarray(np.asarray(bias, dtype=np.float32))
        if w.shape != (288, 1280) or b.shape != (288,):
            raise ValueError("heads: weight must be (288,1280) and bias (288,)")
        with torch.cuda.device(model._device()):
            _lib.check(_lib.load().arx_heads_load(h, w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 0, model._stream()), h,
                       "arx_heads_load")

    def __call__(self, feats):
        """feats (B,8,8,1280) float32 CUDA -> logits (B,8,8,288) float32."""
        m = self.model
        h = m._ensure()
        dev = m._device()
        x = feats.detach().to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 4 or tuple(x.shape[1:]) != (8, 8, 1280):
            raise ValueError("feats must be (B,8,8,1280)")
        out = torch.empty((x.shape[0], 8, 8, 288), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().arx_heads_forward(h, C.c_void_p(x.data_ptr()), x.shape[0], C.c_void_p(out.data_ptr()), m._stream()), h,
                       "arx_heads_forward")
        return out


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

    def decode_cams(self, logits, new_K, homo_inv):
        """Per-frame cameras: logits (B,8,8,288), new_K (B,3,3), homo_inv (B,3,3) -> poses (B, 3*n_out), valid (B,)."""
        m = self.model
        h = m._ensure()
        dev = m._device()
        x = logits.detach().to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 4 or tuple(x.shape[1:]) != (8, 8, 288):
            raise ValueError("logits must be (B,8,8,288)")
        B = x.shape[0]
        K = torch.as_tensor(np.asarray(new_K, dtype=np.float32).reshape(B, 9)).to(dev).contiguous()
        R = torch.as_tensor(np.asarray(homo_inv, dtype=np.float32).reshape(B, 9)).to(dev).contiguous()
        poses = torch.empty((B, self.n_out * 3), dtype=torch.float32, device=dev)
        valid = torch.empty((B,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().arx_decode_heatmaps_cams(
                h, C.c_void_p(x.data_ptr()), B, C.c_void_p(self.expand.data_ptr()), self.n_out, C.c_void_p(K.data_ptr()),
                C.c_void_p(R.data_ptr()), C.c_void_p(poses.data_ptr()), C.c_void_p(valid.data_ptr()), m._stream()), h,
                "arx_decode_heatmaps_cams")
        return poses, valid.bool()

    def decode_tta(self, logits, num_aug=None):
        """Test-time augmentation (hpe.py:87-93): `logits` (num_aug,8,8,288) are the heads' outputs for the num_aug augmented crops
        of ONE camera frame; crop k is decoded with ITS camera (tta_cameras) in one launch.  Returns (poses (num_aug, 3*n_out),
        valid (num_aug,), mean pose over the valid augmentations).  The reference feeds only crop 0 onward (hpe.py:109 takes
        `logits[0]`); the per-crop poses here are what its decode gives for each crop taken as crop 0, the mean is MetrABS's
        aggregation and has no reference counterpart."""
        n = logits.shape[0] if num_aug is None else num_aug
        K, R, _ = tta_cameras(self.new_K, self.homo_inv, n)
        poses, valid = self.decode_cams(logits, K, R)
        mean = poses[valid].mean(0) if bool(valid.any()) else torch.zeros_like(poses[0])
        return poses, valid, mean
