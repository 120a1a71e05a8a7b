"""Statistics leg of `calculate_metrics.py gen` — drop-in for `calculate_stats_for_iterable_nvs` and
`calculate_metrics_from_stats_nvs` (reference calculate_metrics.py:133-240, :289-322).

The per-batch accumulation (fp64 feature sums, fp64 X^T X with the joint [image | source-view] variants, per-image
PSNR) runs in libvividb200.so (vb_stats_update / vb_psnr_u8, csrc/metrics.cu); the cross-rank reduction is
torch.distributed.all_reduce exactly as in the reference, and the closing Frechet distance is the reference's
numpy/scipy expression on the host (a 2048^2 sqrtm once per evaluation, not part of the per-batch path).

Detector networks (InceptionV3 / DINOv2, calculate_metrics.py:29-83) are downloaded by the reference and are NOT part
of this package: pass them in as `detectors={"fid": callable, ...}` (any callable mapping uint8 NCHW images to
[N, feature_dim] features, with a `feature_dim` attribute).  Metrics whose detector is missing raise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .generate import EasyDict
from .imageops import _require_cuda, _stream, psnr_u8, resize_bilinear  # noqa: F401

STAT_METRICS = ("fid", "fd_dinov2")
_DT = {torch.float32: L.VB_F32, torch.float16: L.VB_F16, torch.bfloat16: L.VB_BF16, torch.float64: L.VB_F64}


def stats_update(cum_mu, cum_sigma, features, features2=None):
    """cum_mu += [f | f2].sum(0); cum_sigma += [f | f2]^T [f | f2] in fp64 (calculate_metrics.py:158-172)."""
    _require_cuda(features, "features")
    if features.dtype not in _DT:
        features = features.to(torch.float32)
    f1 = features if features.stride(-1) == 1 else features.contiguous()
    f2 = None
    if features2 is not None:
        f2 = features2.to(f1.dtype)
        f2 = f2 if f2.stride(-1) == 1 else f2.contiguous()
        assert f2.shape[0] == f1.shape[0]
    F = f1.shape[1] + (f2.shape[1] if f2 is not None else 0)
    assert cum_mu.dtype == torch.float64 and cum_sigma.dtype == torch.float64
    assert cum_mu.shape == (F,) and cum_sigma.shape == (F, F) and cum_sigma.is_contiguous()
    if f1.shape[0] == 0:
        return
    d = L.StatsDesc(feat=f1.data_ptr(), feat2=L.ptr(f2), cum_mu=cum_mu.data_ptr(), cum_sigma=cum_sigma.data_ptr(),
                    ld1=f1.stride(0), ld2=f2.stride(0) if f2 is not None else 0, dtype=_DT[f1.dtype], n=f1.shape[0],
                    f1=f1.shape[1], f2=f2.shape[1] if f2 is not None else 0)
    L.check(L.lib().vb_stats_update(C.byref(d), _stream(f1.device)), "vb_stats_update")


def _all_reduce(x):
    x = x.clone()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(x)
    return x


def finalize_stats(cum_mu, cum_sigma, num_images):
    """mu = sum / N; sigma = (sum_outer - N mu mu^T) / (N - 1) after the cross-rank reduction (:174-183). Host logic."""
    mu = _all_reduce(cum_mu) / num_images
    sigma = (_all_reduce(cum_sigma) - mu.ger(mu) * num_images) / (num_images - 1)
    return dict(mu=mu.cpu().numpy(), sigma=sigma.cpu().numpy())


def calculate_stats_for_iterable_nvs(image_iter, metrics=("psnr",), verbose=True, dest_path=None,
                                     device=torch.device("cuda"), detectors=None):
    """Same contract as the reference (:133-240): an iterable yielding `(r, ref)` per batch; on the last batch
    `r.stats` / `ref.stats` hold `{metric: {mu, sigma}}`, `{'psnr': {val}}` and `num_images`."""
    metrics = list(metrics)
    detectors = dict(detectors or {})
    device = torch.device(device)
    for metric in metrics:
        if "joint_" in metric:
            assert metric.replace("joint_", "") in metrics
        base = metric.replace("joint_", "")
        if base in STAT_METRICS and base not in detectors:
            raise NotImplementedError(
                f"metric '{metric}' needs the '{base}' detector network, which the reference downloads "
                "(calculate_metrics.py:29-83); pass detectors={'%s': callable} — it is not shipped here" % base)
        if base not in STAT_METRICS and metric != "psnr":
            raise ValueError(f"Invalid metric '{metric}'")
    num_batches = len(image_iter)

    class StatsIterable:
        def __len__(self):
            return num_batches

        def __iter__(self):
            def make_state():
                out = []
                for metric in metrics:
                    if metric not in STAT_METRICS:
                        continue
                    det = detectors[metric]
                    fd = int(det.feature_dim)
                    s = EasyDict(metric=metric, detector=det, joint="joint_" + metric in metrics)
                    s.cum_mu = torch.zeros([fd], dtype=torch.float64, device=device)
                    s.cum_sigma = torch.zeros([fd, fd], dtype=torch.float64, device=device)
                    if s.joint:
                        s.j_cum_mu = torch.zeros([2 * fd], dtype=torch.float64, device=device)
                        s.j_cum_sigma = torch.zeros([2 * fd, 2 * fd], dtype=torch.float64, device=device)
                    out.append(s)
                return out

            state, ref_state = make_state(), make_state()
            cum_psnr = torch.zeros([1], dtype=torch.float64, device=device) if "psnr" in metrics else None
            cum_images = torch.zeros([], dtype=torch.int64, device=device)
            cum_tgt = torch.zeros([], dtype=torch.int64, device=device)

            def reduce(st, r):
                for s in st:
                    r.stats[s.metric] = finalize_stats(s.cum_mu, s.cum_sigma, r.num_images)
                    if s.joint:
                        r.stats["joint_" + s.metric] = finalize_stats(s.j_cum_mu, s.j_cum_sigma, r.num_images)

            for batch_idx, data in enumerate(image_iter):
                if isinstance(data, dict):
                    src, tgt, images = (None if data[k] is None else torch.as_tensor(data[k]).to(device)
                                        for k in ("src", "tgt", "images"))
                else:
                    src, tgt, images = (torch.as_tensor(k).to(device) for k in data[:3])
                if images is not None and tgt is not None:
                    with torch.no_grad():
                        for s, sref in zip(state, ref_state):
                            f_img = s.detector(images)
                            f_tgt = s.detector(tgt)
                            stats_update(s.cum_mu, s.cum_sigma, f_img)
                            stats_update(sref.cum_mu, sref.cum_sigma, f_tgt)
                            if s.joint:
                                f_src = s.detector(src)
                                stats_update(s.j_cum_mu, s.j_cum_sigma, f_img, f_src)
                                stats_update(sref.j_cum_mu, sref.j_cum_sigma, f_tgt, f_src)
                    cum_images += images.shape[0]
                    cum_tgt += tgt.shape[0]
                    if cum_psnr is not None:
                        # (dual-source batches carry every target twice, generate_images.py:96-98)
                        psnr_u8(images, tgt[::2] if tgt.shape[0] == 2 * images.shape[0] else tgt, cum_psnr)

                r = EasyDict(stats=None, images=images, batch_idx=batch_idx, num_batches=num_batches)
                r.num_images = int(_all_reduce(cum_images).cpu())
                ref = EasyDict(stats=None, images=images, batch_idx=batch_idx, num_batches=num_batches)
                ref.num_images = int(_all_reduce(cum_tgt).cpu())
                if batch_idx == num_batches - 1:
                    assert r.num_images >= 2
                    r.stats = dict(num_images=r.num_images)
                    reduce(state, r)
                    if cum_psnr is not None:
                        r.stats["psnr"] = dict(val=(_all_reduce(cum_psnr) / r.num_images).cpu().numpy())
                    if dest_path is not None and _rank() == 0:
                        save_stats(r.stats, dest_path)
                    assert ref.num_images >= 2
                    ref.stats = dict(num_images=ref.num_images)
                    reduce(ref_state, ref)
                yield r, ref

    return StatsIterable()


def _rank():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return 0


def save_stats(stats, path):
    """.npz with '<metric>/<key>' entries or a pickle, by extension."""
    import os
    import pickle
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if path.lower().endswith(".npz"):
        flat = {"num_images": np.asarray(stats["num_images"])}
        for m, v in stats.items():
            if isinstance(v, dict):
                for k, a in v.items():
                    flat[f"{m}/{k}"] = a
        np.savez(path, **flat)
    else:
        with open(path, "wb") as f:
            pickle.dump(stats, f)


def frechet_distance(mu, sigma, mu_ref, sigma_ref):
    """|mu - mu_ref|^2 + Tr(sigma + sigma_ref - 2 (sigma sigma_ref)^(1/2)) (calculate_metrics.py:309-311). Host."""
    import scipy.linalg
    m = np.square(mu - mu_ref).sum()
    s = scipy.linalg.sqrtm(np.dot(sigma, sigma_ref))     # (scipy >= 1.16 has no `disp`; older ones return the same matrix)
    s = s[0] if isinstance(s, tuple) else s
    return float(np.real(m + np.trace(sigma + sigma_ref - s * 2)))


def calculate_metrics_from_stats_nvs(stats, ref, metrics=("fid", "fd_dinov2", "joint_fid", "joint_fd_dinov2", "psnr"),
                                     verbose=True):
    results = dict()
    for metric in metrics:
        is_stat = metric.replace("joint_", "") in STAT_METRICS
        if metric not in stats or (is_stat and metric not in ref):
            if verbose:
                print(f"No statistics computed for {metric} -- skipping.")
            continue
        if is_stat:
            value = frechet_distance(stats[metric]["mu"], stats[metric]["sigma"], ref[metric]["mu"], ref[metric]["sigma"])
        else:
            value = float(np.asarray(stats[metric]["val"]).reshape(-1)[0])
        results[metric] = value
        if verbose:
            print(f"{metric} = {value:g}")
    return results
