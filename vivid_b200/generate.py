"""Generation driver — drop-in for the reference's `generate_images_nvs`
(generate_images.py:139-343; snapshot experiments/code/generate_images.py:116-260) and the `gen`
call shape of calculate_metrics.py:406-430.

Differences that are deliberate:
  * runs without litdata / RealEstate files: the default dataset is the seed-keyed synthetic one
    (vivid_b200.synthetic), any object with `batch(seeds) -> dict` honouring the reference's batch
    contract (src_image, tgt_image, geometry, sr_src_image, sr_tgt_image, sr_geometry) can be passed;
  * works with or without an initialised process group (the reference calls barrier() unconditionally);
  * per-sample inputs are keyed by seed, so outputs do not depend on the world size;
  * the SR stage's global-RNG noise is seeded per batch from the batch's first seed.
"""
import os
import pickle

import numpy as np
import torch

from .encoders import StandardRGBEncoder
from .imageops import resize_bilinear
from .precond import NVPrecond
from .sampler import StackedRandomGenerator, edm_sampler
from . import synthetic


class EasyDict(dict):
    """Attribute-style dict (the reference uses dnnlib.EasyDict for the per-batch record)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _rank_world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def _barrier():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()


def resolve_model(model, device, name="model"):
    """Path / reference module / vivid_b200 module / None -> vivid_b200.NVPrecond on `device`
    (reference training/utils.py:219-229 + generate_images.py:164-174)."""
    if model is None:
        return None
    if isinstance(model, str):
        with open(model, "rb") as f:
            data = pickle.load(f)        # needs the reference's torch_utils.persistence importable, as in the reference
        model = data["ema" if "ema" in data else "net"]
    if isinstance(model, NVPrecond):
        return model.to(device).eval()
    if hasattr(model, "unet") and hasattr(model, "state_dict"):      # duck-typed reference NVPrecond
        return NVPrecond.from_reference(model).to(device).eval()
    raise TypeError(f"cannot interpret {name} of type {type(model).__name__}")


class SyntheticDataset:
    """Seed-keyed stand-in for datautils.RealEstate10K / CustomLitDataset (batch-dict contract, datautils.py:99)."""

    def __init__(self, imsize=64, sr_imsize=256, dual=False):
        self.imsize, self.sr_imsize, self.dual = imsize, sr_imsize, dual

    def batch(self, seeds):
        lo = synthetic.synth_batch(seeds, self.imsize, dual=self.dual)
        hi = synthetic.synth_batch(seeds, self.sr_imsize, dual=self.dual)
        return dict(src_image=lo["src_image"], tgt_image=lo["tgt_image"], geometry=lo["geometry"],
                    sr_src_image=hi["src_image"], sr_tgt_image=hi["tgt_image"], sr_geometry=hi["geometry"])


def split_seeds(num_seeds, max_batch_size, rank, world_size):
    """Seed -> batch -> rank partition, identical to generate_images.py:199-200."""
    num_batches = max((num_seeds - 1) // (max_batch_size * world_size) + 1, 1) * world_size
    return np.array_split(np.arange(num_seeds), num_batches)[rank::world_size]


def generate_images_nvs(
    net, gnet=None, encoder=None, outdir=None, subdirs=False, seeds=range(16, 24), class_idx=None, max_batch_size=32,
    encoder_batch_size=None, verbose=True, device=torch.device("cuda"), sampler_fn=edm_sampler, datakwargs=None,
    range_selection=None, sr_model=None, depth_model=None, dataset=None, gather_images=False, shard=None,
    **sampler_kwargs,
):
    """Extras over the reference's signature: `dataset` (see the module docstring), `gather_images=True` (every batch's
    uint8 images, seeds and per-rank counts are collected on rank 0 with ONE NCCL gather per batch: `r.gathered_images`
    [sum_b,3,R,R] uint8 in rank order and `r.gathered_seeds` on rank 0, None elsewhere — SURVEY.md §8(e)), and
    `shard=(rank, world)` to take one rank's share of the seeds without a process group (tests, bench cross-checks)."""
    if depth_model is not None:
        raise NotImplementedError("depth models are outside the B200 hot path (all presets: depth_input=False)")
    device = torch.device(device)
    rank, world = _rank_world() if shard is None else (int(shard[0]), int(shard[1]))
    collective = shard is None and world > 1
    if gather_images and shard is not None and world > 1:
        raise ValueError("gather_images needs the process group; it cannot be combined with shard=")
    if rank != 0 and shard is None:
        _barrier()                      # rank 0 goes first (model files)
    net = resolve_model(net, device, "net")
    assert net is not None
    gnet = resolve_model(gnet, device, "gnet")
    if gnet is None:
        gnet = net
    if encoder is None:
        encoder = StandardRGBEncoder()
    encoder.init(device)
    sr_model = resolve_model(sr_model, device, "sr_model")
    if rank == 0 and shard is None:
        _barrier()

    seeds = list(seeds)
    rank_batches = split_seeds(len(seeds), max_batch_size, rank, world)
    all_batches = [split_seeds(len(seeds), max_batch_size, k, world) for k in range(world)] if gather_images else None
    super_res = net.img_resolution == 256
    dual = bool(getattr(net, "dual", False))
    if dataset is None:
        dataset = SyntheticDataset(imsize=64 if super_res else net.img_resolution,
                                   sr_imsize=sr_model.img_resolution if sr_model is not None else 256, dual=dual,
                                   **(datakwargs or {}))
    sr_sampler_kwargs = {k: v for k, v in sampler_kwargs.items() if k != "guidance"}     # no guidance in the SR stage

    class ImageIterable:
        def __len__(self):
            return len(rank_batches)

        def __iter__(self):
            for batch_idx, indices in enumerate(rank_batches):
                r = EasyDict(images=None, src=None, tgt=None, labels=None, noise=None, batch_idx=batch_idx,
                             num_batches=len(rank_batches), indices=indices, gathered_images=None, gathered_seeds=None)
                r.seeds = [seeds[idx] for idx in indices]
                if len(r.seeds) > 0:
                    data = dataset.batch(r.seeds)
                    pre = "sr_" if super_res else ""
                    r.src, r.tgt, geometry = data[pre + "src_image"], data[pre + "tgt_image"], data[pre + "geometry"]
                    if dual:
                        # dual-source collation (generate_images.py:268-282): the record keeps ONE row per seed
                        # (data[k][::2]); the model is fed every row twice
                        r.src, r.tgt, geometry = r.src[::2], r.tgt[::2], geometry[::2]
                        src_for_model, geometry = r.src.repeat_interleave(2, dim=0), geometry.repeat_interleave(2, dim=0)
                    else:
                        src_for_model = r.src
                    src = encoder.encode_latents(src_for_model.to(device, non_blocking=True))
                    rnd = StackedRandomGenerator(device, r.seeds)
                    r.noise = rnd.randn([len(r.seeds), net.img_channels, net.img_resolution, net.img_resolution], device=device)
                    if dual:
                        r.noise = r.noise.repeat_interleave(2, dim=0)
                    r.labels = geometry.to(device, non_blocking=True)
                    kwargs = dict(sampler_kwargs)
                    if super_res:
                        tgt = encoder.encode_latents(r.tgt.to(device))
                        # low-res conditioning of the SR-only path (generate_images.py:282-283), vb_resize
                        small = resize_bilinear(tgt, tgt.shape[-1] // 4, antialias=True)
                        kwargs["conditioning_image"] = resize_bilinear(small, tgt.shape[-1])
                        torch.manual_seed(int(r.seeds[0]) % (1 << 32))
                    with torch.no_grad():
                        latents = sampler_fn(net=net, src=src, noise=r.noise, labels=r.labels, gnet=gnet,
                                             randn_like=rnd.randn_like, **kwargs)
                        r.images = encoder.decode(latents)
                    if sr_model is not None:
                        r.src, r.tgt, sr_geometry = data["sr_src_image"], data["sr_tgt_image"], data["sr_geometry"]
                        if dual:        # one row per seed, as above (the reference's [:num] slice of the 2B rows cannot run)
                            r.src, r.tgt, sr_geometry = r.src[::2], r.tgt[::2], sr_geometry[::2]
                        sr_dual = bool(getattr(sr_model, "dual", False))
                        twice = (lambda t: t.repeat_interleave(2, dim=0)) if sr_dual else (lambda t: t)
                        sr_src = encoder.encode_latents(twice(r.src).to(device, non_blocking=True))
                        rnd = StackedRandomGenerator(device, r.seeds)
                        r.noise = twice(rnd.randn([len(r.seeds), sr_model.img_channels, sr_model.img_resolution,
                                                   sr_model.img_resolution], device=device))
                        r.labels = twice(sr_geometry).to(device, non_blocking=True)
                        # inter-stage bilinear upscale (generate_images.py:322), vb_resize
                        low_res = resize_bilinear(latents, sr_src.shape[-1])
                        torch.manual_seed(int(r.seeds[0]) % (1 << 32))
                        with torch.no_grad():
                            sr_latents = sampler_fn(net=sr_model, src=sr_src, noise=r.noise, labels=r.labels, gnet=sr_model,
                                                    randn_like=rnd.randn_like, conditioning_image=low_res,
                                                    **sr_sampler_kwargs)
                            r.images = encoder.decode(sr_latents)
                    if outdir is not None:
                        import PIL.Image
                        tiles = zip(r.seeds, r.src.clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy(),
                                    r.tgt.clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy(),
                                    r.images.permute(0, 2, 3, 1).cpu().numpy())
                        for seed, _src, _tgt, image in tiles:
                            image_dir = os.path.join(outdir, f"{seed // 1000 * 1000:06d}") if subdirs else outdir
                            os.makedirs(image_dir, exist_ok=True)
                            PIL.Image.fromarray(_src, "RGB").save(os.path.join(image_dir, f"src_{seed:06d}.png"))
                            PIL.Image.fromarray(_tgt, "RGB").save(os.path.join(image_dir, f"tgt_{seed:06d}.png"))
                            PIL.Image.fromarray(image, "RGB").save(os.path.join(image_dir, f"sample_{seed:06d}.png"))
                if gather_images:
                    out_net = sr_model if sr_model is not None else net
                    r.gathered_images, r.gathered_seeds = gather_batch(
                        r.images, [[seeds[i] for i in all_batches[k][batch_idx]] for k in range(world)],
                        (out_net.img_channels, out_net.img_resolution, out_net.img_resolution), device, rank,
                        world if collective else 1)
                if shard is None:
                    _barrier()          # keep the ranks in step (per-batch metric all_reduce)
                yield r

    return ImageIterable()


def gather_batch(images, seeds_by_rank, chw, device, rank, world):
    """Collect one batch's uint8 images of every rank on rank 0 (the only data-path collective besides
    the metric all_reduces): one NCCL gather of [max_b,3,R,R] uint8 (ranks with a shorter or empty batch pad), sliced
    back to the true per-rank sizes, which every rank knows from the seed split.  Returns (images, seeds) on rank 0 in
    rank order, (None, None) elsewhere."""
    sizes = [len(s) for s in seeds_by_rank]
    flat_seeds = [s for part in seeds_by_rank for s in part]
    if world == 1:
        return images, flat_seeds
    mx = max(max(sizes), 1)
    send = torch.zeros((mx,) + tuple(chw), dtype=torch.uint8, device=device)
    if images is not None and images.shape[0]:
        send[:images.shape[0]].copy_(images)
    recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    torch.distributed.gather(send, recv, dst=0)
    if rank != 0:
        return None, None
    return torch.cat([t[:n] for t, n in zip(recv, sizes)]), flat_seeds


def get_metrics(image_iter, device=torch.device("cuda")):
    """PSNR leg of calculate_metrics.get_metrics / calculate_stats_for_iterable_nvs
    (calculate_metrics.py:148,221-236): per-image PSNR of the uint8 images against tgt (vb_psnr_u8, fp64 statistics),
    one int64 counter and one fp64 sum all_reduced.  The feature-statistics metrics take user-supplied detectors:
    vivid_b200.metrics.calculate_stats_for_iterable_nvs."""
    from . import metrics as M
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("vivid_b200.get_metrics runs on CUDA only; there is no CPU fallback")
    psnr_sum = torch.zeros([1], dtype=torch.float64, device=device)
    count = torch.zeros([], dtype=torch.int64, device=device)
    for r in image_iter:
        if r.images is None:
            continue
        img = r.images.to(device)
        tgt = r.tgt.to(device)
        if tgt.shape[0] != img.shape[0]:        # dual-source: targets are duplicated per source
            tgt = tgt[::2]
        M.psnr_u8(img, tgt, psnr_sum)
        count += img.shape[0]
    return reduce_psnr(psnr_sum, count)


def reduce_psnr(psnr_sum, count):
    """Cross-rank reduction of the PSNR accumulators (calculate_metrics.py:221-236). Host logic, any device."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(psnr_sum)
        torch.distributed.all_reduce(count)
    n = int(count.item())
    return dict(psnr=float(psnr_sum.sum().item() / max(n, 1)), num_images=n)
