"""Synthetic inputs of the reference's batch-dict shape: image-like uint8-valued
source/target views and RealEstate10K-style relative poses packed by compose_geometry.
Everything is keyed by the sample SEED so results do not depend on world size (§8(e))."""
import math

import torch

# Geometry normalisation statistics — data constants of the reference's input contract
# (training/utils.py:38-44).
GEOM_MEAN = (9.6681e-01, -1.6038e-04, -3.7034e-05, -1.6904e-03, -8.7718e-05, 9.9869e-01, 3.1288e-03, -1.0794e-03,
             1.0653e-05, 3.0997e-03, 9.6691e-01, 1.2561e-02, 5.7708e+01, 5.7704e+01, 3.2000e+01, 3.2000e+01,
             5.7708e+01, 5.7704e+01, 3.2000e+01, 3.2000e+01)
GEOM_STD = (0.1104, 0.0346, 0.2279, 0.4930, 0.0347, 0.0091, 0.0367, 0.2208, 0.2279, 0.0368, 0.1088, 1.0751, 6.6464,
            6.6511, 0.0, 0.0, 6.6464, 6.6511, 0.0, 0.0)


def compose_K(K):
    """3x3 intrinsic matrices -> [fx, fy, cx, cy] (snapshot experiments/code/training/utils.py:48-52)."""
    return torch.stack((K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2]), -1)


def decompose_K(t):
    """[fx, fy, cx, cy] -> 3x3 intrinsic matrices (snapshot utils.py:55-62)."""
    K = torch.zeros(size=t.shape[:-1] + (3, 3), dtype=t.dtype, device=t.device)
    K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2] = t.unbind(-1)
    K[..., 2, 2] = 1
    return K


def _geom_stats(ref, imsize):
    mean = torch.tensor(GEOM_MEAN, dtype=ref.dtype, device=ref.device).clone()
    std = torch.tensor(GEOM_STD, dtype=ref.dtype, device=ref.device).clone()
    mean[12:] *= imsize / 64
    std[12:] *= (imsize / 64) ** 2
    return mean, std


def compose_geometry(tgt2src, src_k, tgt_k, imsize=64):
    """Pack [R|t] (3x4) + two intrinsics into the normalised 20-vector the networks are conditioned on (zero where the
    statistic's std is 0).  Both call forms of the reference are accepted: [fx,fy,cx,cy] vectors (current tree,
    training/utils.py:64-81) and 3x3 K matrices (snapshot experiments/code/training/utils.py:65-75, via compose_K)."""
    if src_k.shape[-2:] == (3, 3) and src_k.ndim == tgt2src.ndim:
        src_k = compose_K(src_k)
    if tgt_k.shape[-2:] == (3, 3) and tgt_k.ndim == tgt2src.ndim:
        tgt_k = compose_K(tgt_k)
    mean, std = _geom_stats(tgt2src, imsize)
    flat = torch.cat((tgt2src.reshape(*tgt2src.shape[:-2], 12), src_k, tgt_k), dim=-1)
    return torch.where(std > 0, (flat - mean) / std, torch.zeros_like(flat))


def decompose_geometry(t, imsize=64):
    """Inverse of compose_geometry: (tgt2src [.,3,4], src_K [.,3,3], tgt_K [.,3,3]) (snapshot utils.py:78-87)."""
    mean, std = _geom_stats(t, imsize)
    t = t * std + mean
    return t[..., :12].reshape(*t.shape[:-1], 3, 4), decompose_K(t[..., 12:16]), decompose_K(t[..., 16:])


def resize_geometry(geometry, _from, _to):
    """Re-express a pose vector for another image size (snapshot utils.py:90-97)."""
    tgt2src, src_K, tgt_K = decompose_geometry(geometry, _from)
    src_K[..., :2, :] = src_K[..., :2, :] * _to / _from
    tgt_K[..., :2, :] = tgt_K[..., :2, :] * _to / _from
    return compose_geometry(tgt2src, src_K, tgt_K, _to)


def _rot(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    rz = torch.tensor([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    return ry @ rx @ rz


def synth_pose(seed, imsize):
    """One RealEstate10K-style relative pose: small yaw/pitch/roll, translation and focal jitter."""
    g = torch.Generator().manual_seed(0x5EED0000 + int(seed))
    r = torch.randn(8, generator=g).tolist()
    rot = _rot(0.23 * r[0], 0.035 * r[1], 0.035 * r[2])
    t = torch.tensor([0.7 * r[3], 0.47 * r[4], 1.04 * r[5]]).reshape(3, 1)
    ext = torch.cat([rot, t], dim=1)
    f_src = (57.7 + 6.65 * r[6]) * imsize / 64
    f_tgt = (57.7 + 6.65 * r[7]) * imsize / 64
    c = imsize / 2
    return compose_geometry(ext, torch.tensor([f_src, f_src, c, c]), torch.tensor([f_tgt, f_tgt, c, c]), imsize)


def synth_image(seed, res, salt=0):
    """uint8-valued float image [3,res,res]: low-pass filtered noise (image-like spectrum)."""
    g = torch.Generator().manual_seed(0x1A6E0000 + 7919 * salt + int(seed))
    low = max(res // 4, 1)
    x = torch.randint(0, 256, (1, 3, low, low), generator=g).float()
    x = torch.nn.functional.interpolate(x, size=(res, res), mode="bilinear", align_corners=False)
    return x[0].round().clamp(0, 255)


def synth_batch(seeds, res, dual=False):
    """dict(src_image, tgt_image, geometry) for the given seeds; dual-source: 2 sources per target, interleaved."""
    srcs, tgts, geos = [], [], []
    for s in seeds:
        views = 2 if dual else 1
        for v in range(views):
            srcs.append(synth_image(s, res, salt=1 + v))
            tgts.append(synth_image(s, res, salt=0))
            geos.append(synth_pose(2 * int(s) + v if dual else s, res))
    return dict(src_image=torch.stack(srcs), tgt_image=torch.stack(tgts), geometry=torch.stack(geos))
