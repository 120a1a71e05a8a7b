"""BASELINE.json configs[3]: the `calculate_metrics.py gen --num N --batch 32` shape on real GPUs (run under torchrun):
seeds 0..N-1 split with the reference's formula into batches of 31-32 (generate_images.py:199-200), the two-stage
guided pipeline per batch, every batch's images gathered on rank 0 (one NCCL gather), PSNR + fp64 mu/Sigma statistics
accumulated per batch (vb_psnr_u8 / vb_stats_update) with a synthetic 2048-feature detector standing in for InceptionV3
(so the closing all_reduces have the real sizes: 2048^2 and the joint 4096^2 fp64 Sigma, calculate_metrics.py:176-181),
two int64 counter all_reduces per batch (:225-229).  Prints one JSON line on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sweep_gen.py [num=252] [batch=32]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vivid_b200  # noqa: E402
from vivid_b200.generate import SyntheticDataset, split_seeds  # noqa: E402

num = int(sys.argv[1]) if len(sys.argv) > 1 else 252
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)


class FakeInception:
    """Deterministic 2048-feature stand-in for the downloaded detector (uint8/float NCHW -> [N, 2048], on the GPU)."""
    feature_dim = 2048

    def __init__(self):
        g = torch.Generator().manual_seed(7)
        self.w = (torch.randn(3 * 8 * 8, 2048, generator=g) / 14).to(dev)

    def __call__(self, x):
        f = torch.nn.functional.adaptive_avg_pool2d(torch.as_tensor(x).to(dev, torch.float32) / 255.0, 8).flatten(1)
        return torch.tanh(f @ self.w)


net, gnet, sr = (bench.make_net(n, i, dev) for i, n in enumerate(("vivid-base", "vivid-uncond", "vivid-sr")))
sizes = sorted({len(b) for k in range(world) for b in split_seeds(num, batch, k, world)})
t0 = time.perf_counter()
for b in sizes:                     # plans (tuning, graph capture) for the batch sizes of this split, outside the timed sweep
    for m in (net, gnet, sr):
        m.plan(b, dev)
torch.cuda.synchronize()
t_plans = time.perf_counter() - t0


def sweep(with_stats):
    it = vivid_b200.generate_images_nvs(net, gnet=gnet, sr_model=sr, seeds=range(num), max_batch_size=batch, device=dev,
                                        dataset=SyntheticDataset(64, 256), verbose=False, num_steps=32, guidance=1.5,
                                        gather_images=with_stats)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    out, marks = None, []
    if with_stats:
        for r, ref in vivid_b200.calculate_stats_for_iterable_nvs(it, metrics=["fid", "joint_fid", "psnr"], verbose=False,
                                                                   device=dev, detectors={"fid": FakeInception()}):
            out = (r, ref)
            torch.cuda.synchronize()
            marks.append(time.perf_counter())
    else:
        for r in it:
            out = r
            torch.cuda.synchronize()
            marks.append(time.perf_counter())
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    per_batch = [b - a for a, b in zip([t] + marks[:-1], marks)]
    return time.perf_counter() - t, out, per_batch


sweep(False)                                    # warm-up: clocks, allocator, graphs
sweep(True)                                     # ... and the statistics kernels / NCCL channels
t_plain, _, pb_plain = sweep(False)
t_full, (r, ref), pb_full = sweep(True)
inner_plain = sum(pb_plain[:-1]) / max(len(pb_plain) - 1, 1)
inner_full = sum(pb_full[:-1]) / max(len(pb_full) - 1, 1)
if rank == 0:
    # (the joint 4096^2 statistics are accumulated and reduced above; their host-side sqrtm is skipped here)
    res = vivid_b200.calculate_metrics_from_stats_nvs(r.stats, ref.stats, metrics=["fid", "psnr"], verbose=False)
    print(json.dumps(dict(workload=f"calculate_metrics gen shape: --num {num} --batch {batch}, seeds split into batches of "
                                   f"{sizes} over {world} GPUs; guided vivid-base -> vivid-sr, 32 steps/stage",
                          n_gpus=world, batches_per_rank=len(split_seeds(num, batch, 0, world)), num_images=r.stats["num_images"],
                          plan_build_s=round(t_plans, 1),
                          sampling_only=dict(seconds=round(t_plain, 3), images_per_s=round(num / t_plain, 3)),
                          with_gather_and_statistics=dict(seconds=round(t_full, 3), images_per_s=round(num / t_full, 3)),
                          overhead_frac=round(t_full / t_plain - 1.0, 5),
                          per_batch_s=dict(sampling_only=round(inner_plain, 4), with_gather_and_statistics=round(inner_full, 4),
                                           overhead_frac=round(inner_full / inner_plain - 1.0, 5)),
                          closing_s=round(pb_full[-1] - inner_full, 4),
                          closing_note="last batch only: all_reduce of mu/Sigma (2048, joint 4096; generated + reference sets) and "
                                       "their device->host copies; once per sweep (320 batches in the --num 10000 run)",
                          statistics="PSNR + fp64 mu/Sigma (2048 features, joint 4096): vb_psnr_u8 / vb_stats_update per batch, "
                                     "all_reduce at the end; 2 int64 counter all_reduces per batch; 1 uint8 image gather per batch",
                          results={k: round(v, 4) for k, v in res.items()})), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
