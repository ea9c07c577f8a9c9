"""CPU restatement (numpy) of the MetrABS-style heatmap decode that feeds the AR path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, frame by frame,
/root/reference/modules/hpe/hpe.py:108-169 (soft-argmax, FOV test, absolute
reconstruction, homography undo, 32->122 joint remap, 30-joint select) and
/root/reference/main.py:103-105 (root-centre + flatten), together with the
helpers of /root/reference/modules/hpe/utils/misc.py it calls:
to_homogeneous :137, reconstruct_ref_fullpersp :141-176, reconstruct_absolute
:183-204, back_project :207-208, is_within_fov :212-220, homography :243-296.

`hpe.py` itself is not importable (tensorrt/pycuda absent), so gen_golden.py
pins these helpers against the importable `misc.py` and the decode block
against an inline transcription check there.
"""
from __future__ import annotations

import numpy as np


def soft_argmax(logits: np.ndarray, n_joints: int = 32):
    """hpe.py:109-146.  logits (b,8,8,32+8*32) f32 -> pred2d (b,32,2) [x<-w,y<-h]*255, pred3d (b,32,3) [x,y,z]."""
    b = logits.shape[0]
    l2 = logits[..., :n_joints]                                   # (b,h,w,j)
    l3 = logits[..., n_joints:].reshape(b, 8, 8, -1, n_joints)    # 'b h w (d j)' -> b h w d j
    # 3-D: softmax over (w,h,d) per joint
    m = l3.max(axis=(1, 2, 3), keepdims=True)
    e = np.exp(l3 - m)
    p = e / e.sum(axis=(1, 2, 3), keepdims=True)                  # float32
    out3 = []
    for ax in (2, 1, 3):                                          # x<-w, y<-h, z<-d
        other = tuple(a for a in (1, 2, 3) if a != ax)
        marg = p.sum(axis=other)                                  # (b, n_ax, j) float32
        coords = np.linspace(0.0, 1.0, p.shape[ax])               # float64
        out3.append(np.tensordot(marg, coords, axes=[[1], [0]]))  # (b,j) float64
    pred3d = np.stack(out3, axis=-1)
    m = l2.max(axis=(1, 2), keepdims=True)
    e = np.exp(l2 - m)
    p = e / e.sum(axis=(1, 2), keepdims=True)
    out2 = []
    for ax in (2, 1):
        other = tuple(a for a in (1, 2) if a != ax)
        marg = p.sum(axis=other)
        coords = np.linspace(0.0, 1.0, p.shape[ax])
        out2.append(np.tensordot(marg, coords, axes=[[1], [0]]))
    pred2d = np.stack(out2, axis=-1) * 255
    return pred2d, pred3d


def is_within_fov(imcoords):                                       # misc.py:212-220
    lower = np.float32(18)
    upper = np.float32(256 - 18)
    return np.all(np.logical_and(imcoords >= lower, imcoords <= upper), axis=-1)


def to_homogeneous(x):                                             # misc.py:137-138
    return np.concatenate([x, np.ones_like(x[..., :1])], axis=-1)


def reconstruct_ref_fullpersp_one(normalized_2d, coords3d_rel, validity_mask):
    """misc.py:141-176 for ONE frame (the reference solves batch element 0 only, :173-174).
    normalized_2d (J,2), coords3d_rel (J,3), validity_mask (J,) -> ref (3,)"""
    J = normalized_2d.shape[0]
    flat2d = normalized_2d.reshape(J * 2)
    scale2d = np.sqrt(np.mean(np.square(flat2d)))
    A = np.concatenate([np.tile(np.eye(2), (J, 1)), -(flat2d / scale2d)[:, None]], axis=1)   # (2J,3)
    rel_backproj = normalized_2d * coords3d_rel[:, 2:] - coords3d_rel[:, :2]
    flatb = rel_backproj.reshape(J * 2)
    scale_b = np.sqrt(np.mean(np.square(flatb)))
    bvec = (flatb / scale_b)[:, None]
    w = validity_mask.astype(np.float32) + np.float32(1e-4)
    w = np.repeat(w, 2)[:, None]
    ref = np.linalg.lstsq(A * w, bvec * w, rcond=None)[0][:, 0]
    return np.array([ref[0], ref[1], ref[2] / scale2d]) * scale_b


def reconstruct_absolute_one(coords2d, coords3d_rel, intrinsics, in_fov):
    """misc.py:183-204 for one frame, weak_perspective=False."""
    inv_k = np.linalg.inv(intrinsics.astype(np.float32))
    n2d = (to_homogeneous(coords2d) @ inv_k.T)[..., :2]
    ref = reconstruct_ref_fullpersp_one(n2d, coords3d_rel, in_fov)
    abs3d_based = coords3d_rel + ref[None]
    abs2d_based = to_homogeneous(n2d) * (coords3d_rel[:, 2] + ref[2])[:, None]      # back_project, misc.py:207-208
    return np.where(in_fov[:, None], abs2d_based, abs3d_based)


def _rotation_to(forward, up):                                     # misc.py:223-236
    z = forward / np.linalg.norm(forward, axis=-1, keepdims=True)
    x = np.cross(z, up)
    x_alt = np.stack([z[:, 2], np.zeros_like(z[:, 2]), -z[:, 0]], axis=1)
    x = np.where(np.linalg.norm(x, axis=-1, keepdims=True) == 0, x_alt, x)
    x = x / np.linalg.norm(x, axis=-1, keepdims=True)
    y = np.cross(z, x)
    return np.stack([x, y, z], axis=1)


def homography(x1, x2, y1, y2, K, out_dim):                        # misc.py:243-296
    pts = to_homogeneous(np.array([[[(x1 + x2) / 2, (y1 + y2) / 2], [(x1 + x2) / 2, y1], [x2, (y1 + y2) / 2],
                                    [(x1 + x2) / 2, y2], [x1, (y1 + y2) / 2]]]))
    cam = pts @ np.linalg.inv(K[None]).transpose((0, 2, 1))
    cam = to_homogeneous(cam[..., :2])
    R = _rotation_to(cam[:, 0], np.array([[0, -1, 0]]))
    side = cam[:, 1:5] @ (K[None] @ R).transpose((0, 2, 1))
    side = side[..., :2] / side[..., 2:3]
    vert = np.linalg.norm(side[:, 0] - side[:, 2], axis=-1)
    horiz = np.linalg.norm(side[:, 1] - side[:, 3], axis=-1)
    scale = out_dim / np.maximum(vert, horiz)
    new_k = np.concatenate([
        np.concatenate([K[:2, :2] * scale, np.full((2, 1), out_dim / 2, dtype=K.dtype)], axis=1),
        np.concatenate([np.zeros((1, 2), np.float32), np.ones((1, 1), np.float32)], axis=1)], axis=0)
    return new_k, R


def realsense_K():
    """hpe.py:28-33 with utils/params.py:40-47."""
    K = np.zeros((3, 3), np.float32)
    K[0][0] = 384.025146484375
    K[0][2] = 319.09661865234375
    K[1][1] = 384.025146484375
    K[1][2] = 237.75723266601562
    K[2][2] = 1
    return K


def decode_frames(logits, expand_joints, indices, new_K, homo_inv, min_in_fov_frac=0.25):
    """hpe.py:108-169 + main.py:103-105 applied independently to every frame.

    logits (B,8,8,288) f32; expand_joints (32,122) f32; indices (30,) ints;
    new_K (3,3); homo_inv (1,3,3) or (3,3).
    Returns poses (B,90) float64 (root-centred, flattened), valid (B,) bool
    (False where the reference returns None, hpe.py:152-153; those rows are 0).
    """
    pred2d, pred3d = soft_argmax(logits)
    B = logits.shape[0]
    R = np.asarray(homo_inv).reshape(3, 3)
    E = np.asarray(expand_joints)
    idx = np.asarray(indices, dtype=np.int64)
    poses = np.zeros((B, len(idx) * 3), np.float64)
    valid = np.zeros((B,), bool)
    for f in range(B):
        fov = is_within_fov(pred2d[f])
        if fov.sum() < fov.size * min_in_fov_frac:                 # hpe.py:152
            continue
        p = reconstruct_absolute_one(pred2d[f], pred3d[f], new_K, fov)
        p = p @ R                                                  # hpe.py:159
        p = (p.T @ E).T                                            # hpe.py:162  (122,3)
        p = p[idx]                                                 # hpe.py:164
        p = p - p[0, :]                                            # main.py:103
        poses[f] = p.reshape(-1)                                   # main.py:105
        valid[f] = True
    return poses, valid


def rotation_mat_zaxis(angle):                                     # misc.py:299-307
    sin, cos = np.sin(angle), np.cos(angle)
    _0, _1 = np.zeros_like(angle), np.ones_like(angle)
    return np.stack([np.stack([cos, -sin, _0], axis=-1), np.stack([sin, cos, _0], axis=-1), np.stack([_0, _0, _1], axis=-1)], axis=-2)


def get_augmentations(num_aug, rot_aug_linspace_noend=True):      # misc.py:310-327
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
    return aug_should_flip, aug_maybe_flipmat @ rotation_mat_zaxis(-aug_angles), aug_gammas, aug_scales


def tta_cameras(new_K, homo_inv, num_aug):
    """hpe.py:88-93: per-augmentation intrinsics (scaled) and homographies (rotation/flip applied first)."""
    flip, rotflip, _, scales = get_augmentations(num_aug)
    K = np.tile(np.asarray(new_K).reshape(3, 3), (num_aug, 1, 1)).astype(np.float64)
    for k in range(num_aug):
        K[k, :2, :2] *= scales[k]
    R = rotflip @ np.tile(np.asarray(homo_inv).reshape(3, 3), (num_aug, 1, 1))
    return K, R, flip


def decode_frames_cams(logits, expand_joints, indices, new_Ks, homo_invs, min_in_fov_frac=0.25):
    """decode_frames with a camera per frame (the call the reference's decode makes for crop k taken as the only crop)."""
    B = logits.shape[0]
    poses = np.zeros((B, len(indices) * 3), np.float64)
    valid = np.zeros((B,), bool)
    for f in range(B):
        p, v = decode_frames(logits[f:f + 1], expand_joints, indices, new_Ks[f], homo_invs[f], min_in_fov_frac)
        poses[f], valid[f] = p[0], v[0]
    return poses, valid


def heads_forward(feats, weight, bias):
    """modules/hpe/setup/4_create_heads_onnx.py:7-16: Linear(1280 -> 288) on (B,8,8,1280)."""
    return feats @ np.asarray(weight).T + np.asarray(bias)
