"""Shared definition of the golden cases: network configs, name-hashed deterministic weights and
seeded inputs.  Used by make_golden.py (which runs the REFERENCE, in the build container only)
and by the tests (which replay the stored outputs against the oracle and the CUDA path).

Weights are not stored: every tensor of a state_dict is regenerated from crc32(name), so a
fixture only carries (name, shape) pairs plus the reference's outputs.
"""
import math
import zlib

import torch

SMALL = dict(img_resolution=16, img_channels=3, model_channels=64, channel_mult=[1, 2], num_blocks=1,
             attn_resolutions=[8])

CASES = {
    # vanilla semantics (snapshot tree experiments/code)
    "v_cond": dict(mode="vanilla", cfg=dict(SMALL, label_dim=20)),
    "v_uncond": dict(mode="vanilla", cfg=dict(SMALL, label_dim=20, uncond=True)),
    "v_sr": dict(mode="vanilla", cfg=dict(SMALL, label_dim=20, super_res=True, noisy_sr=0.25, attn_resolutions=[])),
    "v_tiny": dict(mode="vanilla", cfg=dict(img_resolution=32, img_channels=3, label_dim=20, model_channels=64)),
    # dual-source semantics (current tree)
    "d_cond": dict(mode="dual", cfg=dict(SMALL, source_label_dim=20, target_label_dim=40)),
}
SIGMAS = (80.0, 5.0, 0.5, 0.05)


def synth_tensor(name, shape, salt=0):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (salt * 0x9E3779B1)) & 0x7FFFFFFF)
    if len(shape) == 0:
        # out_gain / emb_gain are zero-initialised in the reference (=> D == c_skip*x, a vacuous test)
        return torch.tensor(0.7 + 0.6 * torch.rand([], generator=g).item())
    if name.endswith("freqs"):
        return 2 * math.pi * torch.randn(shape, generator=g)
    if name.endswith("phases"):
        return 2 * math.pi * torch.rand(shape, generator=g)
    return torch.randn(shape, generator=g)


def synth_state_dict(shapes, salt=0):
    """shapes: iterable of (name, shape) in state_dict order."""
    return {n: synth_tensor(n, tuple(s), salt) for n, s in shapes}


def synth_inputs(case, B, seed=0):
    """Seeded src / tgt / geometry / noise for a case ([-1,1] images, N(0,1) pose vectors)."""
    cfg = CASES[case]["cfg"]
    dual = CASES[case]["mode"] == "dual"
    R = cfg["img_resolution"]
    n = 2 * B if dual else B
    g = torch.Generator().manual_seed(1000 + seed)
    low = torch.rand(n, 3, R // 4, R // 4, generator=g) * 2 - 1
    src = torch.nn.functional.interpolate(low, size=(R, R), mode="bilinear", align_corners=False)
    low = torch.rand(B, 3, R // 4, R // 4, generator=g) * 2 - 1
    tgt = torch.nn.functional.interpolate(low, size=(R, R), mode="bilinear", align_corners=False)
    if dual:
        tgt = tgt.repeat_interleave(2, dim=0)
    geom = torch.randn(n, 20, generator=g)
    noise = torch.randn(B, 3, R, R, generator=g)
    if dual:
        noise = noise.repeat_interleave(2, dim=0)
    return dict(src=src, tgt=tgt, geometry=geom, noise=noise)


# ----------------------------------------------------------------------------- metric statistics (calculate_metrics.py gen)
class FakeDetector:
    """Deterministic stand-in for the downloaded detector networks (calculate_metrics.py:29-83): uint8/float NCHW
    images -> [N, 48] features (4x4 average pooling of a 16x16 image, scaled), so that the statistics code of the
    reference can run offline.  Injected through the reference's own `_detector_cache`."""
    feature_dim = 48

    def __call__(self, x):
        x = torch.as_tensor(x).to(torch.float32)
        f = torch.nn.functional.adaptive_avg_pool2d(x, 4).flatten(1) / 64.0
        return f + 0.25 * torch.sin(f * 3.0)


def synth_metric_batches(num_batches=3, batch=5, res=16, seed=0):
    """[(src, tgt, images)] batches: float src/tgt in [0,255] (as the dataset yields them), uint8 generated images."""
    g = torch.Generator().manual_seed(4242 + seed)
    out = []
    for b in range(num_batches):
        n = batch - (b == num_batches - 1)          # ragged last batch
        src = torch.rand(n, 3, res, res, generator=g) * 255
        tgt = torch.rand(n, 3, res, res, generator=g) * 255
        img = (tgt + 12 * torch.randn(n, 3, res, res, generator=g)).clip(0, 255).to(torch.uint8)
        out.append((src, tgt, img))
    return out
