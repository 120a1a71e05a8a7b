#!/usr/bin/env python
"""Headline benchmark: guided NVS images/sec, vivid-base (+ vivid-uncond gnet, w=1.5) -> vivid-sr,
EDM Heun sampler, 32 steps per stage (63 denoiser calls each), random-init weights of the named
architectures, synthetic images/poses (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference ...                   # the reference's algorithm on the host CPU cores

One "step" = one batch of `--batch` images per GPU through the whole two-stage pipeline.
`value` is measured with all inputs resident in HBM; `e2e` goes through the public driver
(vivid_b200.generate_images_nvs) with pinned HOST inputs and a device->host read of the images.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PRESETS = {
    "vivid-base": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1),
    "vivid-uncond": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1, uncond=True),
    "vivid-sr": dict(img_resolution=256, img_channels=3, label_dim=20, model_channels=64, super_res=True, noisy_sr=0.25),
}
# Algorithmic GFLOP per image per denoiser call (uncond with the zero-feature K/V work elided).
ALG_GFLOP = {"vivid-base": 161.43, "vivid-uncond": 93.37, "vivid-sr": 439.72}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], gbs=p["hbm_gbs"], source="measured")
    except Exception:
        return dict(tflops=1400.0, tflops_burst=1590.0, gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


def make_net(name, seed, device):
    import vivid_b200
    torch.manual_seed(seed)
    net = vivid_b200.NVPrecond(**PRESETS[name])
    with torch.no_grad():
        for p in net.parameters():
            if p.ndim == 0:
                p.fill_(1.0)            # zero-init gains would make every net return c_skip*x
    return net.to(device).eval()


# --------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import vivid_b200
    from vivid_b200.generate import SyntheticDataset
    from vivid_b200.imageops import resize_bilinear
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (vivid_b200 has no CPU fallback); use --impl reference for the CPU arm"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    B, T = args.batch, args.num_steps
    net, gnet, sr = make_net("vivid-base", 0, dev), make_net("vivid-uncond", 1, dev), make_net("vivid-sr", 2, dev)
    enc = vivid_b200.StandardRGBEncoder()
    enc.init(dev)
    ds = SyntheticDataset(imsize=64, sr_imsize=256)

    def seeds_of(step):     # every rank and step gets its own seeds (weak scaling: B images per GPU per step)
        base = (step * world + rank) * B
        return list(range(base, base + B))

    def host_batch(seeds):
        d = ds.batch(seeds)
        return {k: v.pin_memory() for k, v in d.items()}

    def resident(seeds):
        d = {k: v.to(dev) for k, v in ds.batch(seeds).items()}
        rnd = vivid_b200.StackedRandomGenerator(dev, seeds)
        return dict(seed0=seeds[0], src=enc.encode_latents(d["src_image"]), geom=d["geometry"],
                    noise=rnd.randn([B, 3, 64, 64], device=dev),
                    sr_src=enc.encode_latents(d["sr_src_image"]), sr_geom=d["sr_geometry"],
                    sr_noise=vivid_b200.StackedRandomGenerator(dev, seeds).randn([B, 3, 256, 256], device=dev))

    def pipeline(r):
        lat = vivid_b200.edm_sampler(net, r["src"], r["noise"], labels=r["geom"], gnet=gnet, num_steps=T, guidance=args.guidance)
        low = resize_bilinear(lat, 256)                 # inter-stage bilinear x4 (generate_images.py:322), vb_resize
        torch.manual_seed(int(r["seed0"]))              # the SR net's per-call noise comes from the global generator (F7)
        sr_lat = vivid_b200.edm_sampler(sr, r["sr_src"], r["sr_noise"], labels=r["sr_geom"], gnet=sr, num_steps=T,
                                        conditioning_image=low)
        return enc.decode(sr_lat)

    def sync():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_warm, n_steps, prepare):
        items = [prepare(seeds_of(i)) for i in range(n_warm + n_steps)]
        for i in range(n_warm):
            fn(items[i])
        sync()
        clk = ClockSampler(local)
        clk.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_warm, n_warm + n_steps):
            fn(items[i])
        e1.record()
        sync()
        clocks = clk.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item(), clocks

    # ---- value: inputs resident in HBM
    ms, clocks = timed(pipeline, args.warmup, args.steps, resident)
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: public driver, pinned host inputs, H2D + D2H inside the timed region; at N>1 every batch's images are
    #      gathered on rank 0 (one NCCL gather per batch) and the PSNR statistics are all_reduced, as calculate_metrics gen does
    host_items = {}

    class HostDataset:
        def batch(self, seeds):
            return host_items[tuple(seeds)]

    pinned_out = torch.empty((world * B, 3, 256, 256), dtype=torch.uint8).pin_memory() if rank == 0 else None

    def e2e_step(step):
        seeds = list(range(step * world * B, (step + 1) * world * B))       # the driver shards these over the ranks itself
        it = vivid_b200.generate_images_nvs(net, gnet=gnet, sr_model=sr, seeds=seeds, max_batch_size=B, device=dev,
                                            dataset=HostDataset(), verbose=False, num_steps=T, guidance=args.guidance,
                                            gather_images=world > 1)
        recs = []
        for r in it:
            recs.append(r)
            imgs = r.gathered_images if world > 1 else r.images
            if rank == 0:                # device -> host read of the step's result (all ranks' images on rank 0)
                pinned_out[:imgs.shape[0]].copy_(imgs, non_blocking=True)
        m = vivid_b200.get_metrics(iter(recs), device=dev)         # PSNR vs tgt: counter + fp64 sum all_reduce, .item()
        return m

    prep_count = [0]

    def prep_host(seeds):
        host_items[tuple(seeds)] = host_batch(seeds)
        prep_count[0] += 1
        return prep_count[0] - 1

    e2e_ms, e2e_clocks = timed(e2e_step, max(1, args.warmup // 3), args.steps, prep_host) if not args.no_e2e else (None, None)
    h2d = B * (3 * 64 * 64 * 4 * 2 + 20 * 4 + 3 * 256 * 256 * 4 * 2 + 20 * 4)
    d2h = B * 3 * 256 * 256 * (world if rank == 0 else 0) + 16

    # ---- N>1: rank-sharded + gathered output == rank 0 generating every seed alone (tiny nets, bitwise), and the cost of the
    #      collectives of the calculate_metrics gen shape (batches of 31-32)
    multi = multi_gpu_checks(vivid_b200, dev, rank, world) if world > 1 else None

    # ---- roofline, regime 1 (in situ): one more step with kernel activity records (CUPTI through torch.profiler) while the
    #      chip is still at its sustained clocks; net and gnet on ONE stream so that kernel durations do not overlap
    insitu = None
    if not args.no_insitu:
        try:
            insitu = insitu_profile(pipeline, resident(seeds_of(args.warmup + args.steps)), dev, (net, gnet, sr))
        except Exception as e:      # the number is an attribution aid: never lose the bench line over it
            insitu = dict(error=f"{type(e).__name__}: {e}"[:200])

    # ---- roofline, regime 2 (isolated): every recorded op timed alone with CUDA events behind a blocker kernel
    calls = 2 * T - 1
    plans = {"vivid-base": net.plan(B, dev), "vivid-uncond": gnet.plan(B, dev), "vivid-sr": sr.plan(B, dev)}
    agg = {}
    pk = peaks()
    for name, p in plans.items():
        for kind, label, fl, by, op_ms in p.profile(repeats=2):
            a = agg.setdefault(kind, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0, bound_ms=0.0))
            a["ms"] += op_ms * calls
            a["flops"] += fl * calls
            a["bytes"] += by * calls
            a["launches"] += calls
            # the op's own combined roofline: the slower of its FLOPs at the tensor peak and its algorithmic bytes at the HBM
            # copy rate (an HBM-bound 1x1 conv at its byte roofline is not "far from the tensor peak")
            a["bound_ms"] += max(fl / (pk["tflops_burst"] * 1e9), by / (pk["gbs"] * 1e6)) * calls
    conv_ms = agg["conv3"]["ms"] + agg["conv1"]["ms"]
    conv_fl = agg["conv3"]["flops"] + agg["conv1"]["flops"]
    conv_n = agg["conv3"]["launches"] + agg["conv1"]["launches"]
    total_ms = sum(a["ms"] for a in agg.values())
    isolated = conv_fl / (conv_ms / 1e3) / 1e12
    conv_by = agg["conv3"]["bytes"] + agg["conv1"]["bytes"]
    traffic, traffic_src = None, None
    for tag in ("r02", "r01"):      # DRAM bytes per conv launch from the committed ncu launch list of one call per net at this batch
        try:
            with open(os.path.join(ROOT, "profiles", f"{tag}_launches_B{B}.json")) as f:
                tj = json.load(f)
            if tj.get("batch") == B:
                traffic, traffic_src = round(tj["dram_bytes_per_launch"]), "profiles/%s_launches_B%d.json: %s" % (tag, B, tj["source"])
                break
        except Exception:
            pass
    step_ms = ms / args.steps
    if insitu and "conv_ms" in insitu:
        achieved, regime = conv_fl / (insitu["conv_ms"] / 1e3) / 1e12, "in situ"
        avg_us, share = insitu["conv_ms"] / max(insitu["conv_launches"], 1) * 1e3, insitu["conv_ms"] / insitu["step_ms"]
        peak, peak_src = pk["tflops"], f"{pk['source']} bf16 sustained (kernel durations recorded inside a running step)"
    else:           # no in-situ record: the isolated timing belongs against the burst peak
        achieved, regime = isolated, "isolated ops (no in-situ record)"
        avg_us, share = conv_ms / conv_n * 1e3, conv_ms / total_ms
        peak, peak_src = pk["tflops_burst"], f"{pk['source']} bf16 burst (ops timed alone at cool clocks)"
    roofline = dict(bound="tensor", kernel="conv_gemm_kernel (tcgen05 implicit-GEMM 3x3/1x1 conv)", achieved=round(achieved, 1),
                    peak=peak, unit="TFLOP/s", frac=round(achieved / peak, 4), regime=regime, peak_source=peak_src,
                    achieved_isolated=round(isolated, 1), peak_burst=pk["tflops_burst"],
                    frac_isolated=round(isolated / pk["tflops_burst"], 4),
                    isolated_note="ops timed alone behind a blocker at cool clocks, against the burst peak",
                    traffic=traffic, traffic_unit="bytes per launch (dram read+write, ncu)", traffic_source=traffic_src,
                    alg_bytes_per_launch=round(conv_by / conv_n),
                    launches_per_step=conv_n, avg_launch_us=round(avg_us, 2), share_of_step=round(share, 4),
                    alg_flops_per_step=conv_fl)
    kernels = {}
    for kind, a in agg.items():
        k = dict(ms_per_step=round(a["ms"], 2), share=round(a["ms"] / total_ms, 4), launches=a["launches"])
        if a["flops"] > 0:
            k["tflops"] = round(a["flops"] / (a["ms"] / 1e3) / 1e12, 1)
            k["frac_of_tensor_peak"] = round(k["tflops"] / pk["tflops"], 4)
        else:
            k["gbs"] = round(a["bytes"] / (a["ms"] / 1e3) / 1e9, 1)
            k["frac_of_hbm_peak"] = round(k["gbs"] / pk["gbs"], 4)
        k["roofline_ms"] = round(a["bound_ms"], 2)
        k["frac_of_roofline"] = round(a["bound_ms"] / a["ms"], 4)
        kernels[kind] = k
    kernels["_note"] = ("isolated regime (each op timed alone behind a blocker): frac_of_tensor_peak / frac_of_hbm_peak against the "
                        "sustained tensor peak / copy rate; frac_of_roofline = sum over ops of max(FLOPs / burst tensor peak, "
                        "algorithmic bytes / copy rate) / measured time")
    alg_tflop_img = calls * (ALG_GFLOP["vivid-base"] + ALG_GFLOP["vivid-uncond"] + ALG_GFLOP["vivid-sr"]) / 1e3
    launches = calls * sum(p.launches for p in plans.values()) + 2 * calls + 1 + calls + 2     # + Heun, SR noise, resize, decode

    out = dict(metric="guided NVS images/sec (vivid-base+SR)", value=round(value, 3), unit="images/s", n_gpus=world,
               steps=args.steps, warmup=args.warmup, ms_per_step=round(step_ms, 2), higher_is_better=True,
               scaling="weak", vs_baseline=None, dtype="fp16", data="synthetic",
               config=dict(workload="vivid-base guided (vivid-uncond gnet, w=%.1f) -> bilinear x4 -> vivid-sr; Heun %d steps/stage "
                           "(%d denoiser calls each); random-init weights" % (args.guidance, T, calls),
                           batch_per_gpu=B, global_batch=B * world, parallelism=f"sample-sharded x{world}, no collective on the sampling path"
                           + ("; e2e leg: one NCCL image gather per batch + metric all_reduces" if world > 1 else ""),
                           l2="per-step working set (weights 0.84 GB + activations) exceeds the 126 MB L2; no flush needed",
                           operands="fp16 (the reference's own reduced precision; tcgen05 kind::f16)", accumulate="fp32",
                           residual_stream="fp16", sampler_state="fp32"),
               clocks=clocks, gpu_launches=int(launches * args.steps),
               e2e=None if e2e_ms is None else dict(value=round(world * B * args.steps / (e2e_ms / 1e3), 3), unit="images/s",
                                                    h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, clocks=e2e_clocks,
                                                    includes="H2D of pinned inputs, per-seed noise, both stages, uint8 decode, "
                                                             + ("NCCL gather of every rank's images to rank 0, " if world > 1 else "")
                                                             + "D2H of the images into pinned memory, PSNR statistics (all_reduced)"),
               roofline=roofline, kernels=kernels, kernels_insitu=insitu,
               model_tflops=round(value / world * alg_tflop_img, 1),
               model_frac_of_peak=round(value / world * alg_tflop_img / pk["tflops"], 4))
    if multi is not None:
        multi["batch_ms_at_32"] = round(32.0 / (value / world) * 1e3, 1)
        per_batch = multi["counters_ms"] + multi["gather_32_ms"]
        multi["frac_of_batch"] = round(per_batch / multi["batch_ms_at_32"], 6)
        multi["note"] = ("configs[3] shape (gen --num 10000 --batch 32): per batch two int64 counter all_reduces + host reads and one "
                         "uint8 image gather; at the end one all_reduce per statistic (largest: the joint 4096^2 fp64 Sigma). "
                         "batch_ms_at_32 is derived from this line's per-GPU images/s")
        out["collective"] = multi
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline()
    if rank == 0:
        emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


def insitu_profile(pipeline, item, dev, nets):
    """Kernel durations of ONE pipeline step recorded in place (CUPTI activity records via torch.profiler; CUDA side only).
    For the records to be per-kernel the step runs as an eager replay of the plans on one stream with programmatic dependent
    launch off (under PDL a kernel is resident, and counted as running, while it still waits for its predecessor).
    Returns per-family totals, the conv kernel's total, and the part of the step no kernel covers."""
    import re
    from torch.profiler import ProfilerActivity, profile
    from vivid_b200 import _lib as L
    prev = os.environ.get("VB_DUAL_STREAM")
    os.environ["VB_DUAL_STREAM"] = "0"
    prev_pdl = L.lib().vb_set_pdl(0)
    for n in nets:
        n.use_graph = False
    try:
        pipeline(item)                          # same code path once without records (sustained clocks, caches warm)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            e0.record()
            pipeline(item)
            e1.record()
            torch.cuda.synchronize(dev)
    finally:
        for n in nets:
            n.use_graph = True
        L.lib().vb_set_pdl(prev_pdl)
        if prev is None:
            os.environ.pop("VB_DUAL_STREAM", None)
        else:
            os.environ["VB_DUAL_STREAM"] = prev
    kname = re.compile(r"([A-Za-z0-9_]+_kernel)")
    fam, spans = {}, []
    for ev in prof.events():
        dur = getattr(ev, "device_time_total", None)
        if dur is None:
            dur = getattr(ev, "cuda_time_total", 0)
        if not dur:
            continue
        name = ev.name
        low = name.lower()
        if "memcpy" in low or "memset" in low:
            key = "memcpy/memset"
        elif "conv_gemm_kernel" in name:
            key = "conv_gemm_kernel"
        elif "attn" in low:
            key = "attention"
        elif "vb::" in name:
            m = kname.search(name)
            key = m.group(1) if m else name[:48]
        else:
            key = "torch (RNG, copies, stack)"
        f = fam.setdefault(key, [0, 0.0])
        f[0] += 1
        f[1] += dur / 1e3
        tr = getattr(ev, "time_range", None)
        if tr is not None:
            spans.append((tr.start, tr.end))
    step_ms = e0.elapsed_time(e1)
    spans.sort()
    busy, cur_s, cur_e = 0.0, None, None
    for s_, e_ in spans:
        if cur_e is None or s_ > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    if cur_e is not None:
        busy += cur_e - cur_s
    conv = fam.get("conv_gemm_kernel", [0, 0.0])
    total = sum(v[1] for v in fam.values())
    return dict(method="torch.profiler CUDA activity records over one extra step (eager replay, one stream, PDL off)",
                step_ms=round(step_ms, 2),
                kernels_ms=round(total, 2), busy_ms=round(busy / 1e3, 2), idle_ms=round(step_ms - busy / 1e3, 2),
                idle_frac=round(1.0 - busy / 1e3 / step_ms, 4), conv_ms=round(conv[1], 2), conv_launches=conv[0],
                families={k: dict(launches=v[0], ms=round(v[1], 2), share=round(v[1] / step_ms, 4))
                          for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])})


def multi_gpu_checks(vivid_b200, dev, rank, world):
    """(1) tiny nets: images generated rank-sharded and gathered on rank 0 are bit-identical to rank 0 generating every
    batch alone; (2) device time (max over ranks) of the collectives of the calculate_metrics gen shape."""
    from vivid_b200.generate import SyntheticDataset, gather_batch, split_seeds
    dist = torch.distributed
    small = dict(img_channels=3, label_dim=20, model_channels=64, channel_mult=[1, 2], num_blocks=1)
    torch.manual_seed(0)
    nets = [vivid_b200.NVPrecond(img_resolution=16, attn_resolutions=[8], **small),
            vivid_b200.NVPrecond(img_resolution=16, attn_resolutions=[8], uncond=True, **small),
            vivid_b200.NVPrecond(img_resolution=64, attn_resolutions=[], super_res=True, **small)]
    for m in nets:
        with torch.no_grad():
            for p in m.parameters():
                if p.ndim == 0:
                    p.fill_(0.5)
        m.to(dev).eval()
    ds = SyntheticDataset(imsize=16, sr_imsize=64)
    seeds = list(range(100, 100 + 5 * world + 3))          # ragged: some ranks get shorter batches
    kw = dict(gnet=nets[1], sr_model=nets[2], seeds=seeds, max_batch_size=3, device=dev, dataset=ds, verbose=False, num_steps=4,
              guidance=1.5)
    gathered = [(r.gathered_images, r.gathered_seeds) for r in vivid_b200.generate_images_nvs(nets[0], gather_images=True, **kw)]
    ok = True
    if rank == 0:
        alone = {}
        for k in range(world):      # rank 0 replays every rank's share on its own
            for r in vivid_b200.generate_images_nvs(nets[0], shard=(k, world), **kw):
                alone.update({s: img for s, img in zip(r.seeds, r.images)})
        seen = []
        for imgs, sds in gathered:
            seen += sds
            ok = ok and all(torch.equal(imgs[i], alone[s]) for i, s in enumerate(sds))
        ok = ok and sorted(seen) == seeds
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    assert flag.item() == 1, "rank-sharded + gathered images differ from the single-rank run"

    def timed(fn, reps=5):
        fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(t.item(), 4)

    cum_a, cum_b = torch.zeros([], dtype=torch.int64, device=dev), torch.zeros([], dtype=torch.int64, device=dev)

    def counters():                  # calculate_metrics.py:225-229: two counter all_reduces + host reads per batch
        for c in (cum_a, cum_b):
            x = c.clone()
            dist.all_reduce(x)
            int(x.cpu())

    sigma = torch.zeros(4096, 4096, dtype=torch.float64, device=dev)
    imgs = torch.zeros(32 if rank % 2 == 0 else 31, 3, 256, 256, dtype=torch.uint8, device=dev)
    by_rank = [list(range(32 if k % 2 == 0 else 31)) for k in range(world)]
    return dict(gathered_equals_single_rank=True, check_seeds=len(seeds),
                counters_ms=timed(counters), joint_sigma_4096_fp64_allreduce_ms=timed(lambda: dist.all_reduce(sigma.clone())),
                gather_32_ms=timed(lambda: gather_batch(imgs, by_rank, (3, 256, 256), dev, rank, world)))


# --------------------------------------------------------------------------------------------- CPU arms
def _fill_gains(m):
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 0:
                p.fill_(1.0)
    return m


def cpu_nets():
    """The reference's CPU implementation of the path: the UNMODIFIED reference staged under oracle/_ref (kind "reference",
    snapshot tree = the semantics in which guidance and SR run), else the oracle port (kind "port")."""
    from oracle import ref_loader
    names = (("base", "vivid-base"), ("uncond", "vivid-uncond"), ("sr", "vivid-sr"))
    tiny = dict(img_resolution=32, img_channels=3, label_dim=20, model_channels=64)
    if ref_loader.available():
        ns = ref_loader.load("snapshot")

        def make(cfg, seed):
            torch.manual_seed(seed)
            return _fill_gains(ns.models.NVPrecond(use_fp16=False, **cfg)).eval().requires_grad_(False)
        return "reference", {k: make(PRESETS[n], i) for i, (k, n) in enumerate(names)}, make(tiny, 3), ns.generate_images.edm_sampler
    import vivid_b200
    from oracle import vivid_oracle as O

    def make(cfg, seed):
        torch.manual_seed(seed)
        return O.OracleNet(_fill_gains(vivid_b200.NVPrecond(**cfg)).state_dict(), cfg)
    return "port", {k: make(PRESETS[n], i) for i, (k, n) in enumerate(names)}, make(tiny, 3), O.edm_sampler


def cpu_sample_seconds(nets, threads, repeats):
    """`repeats` samples; one sample = one denoiser call each of vivid-base, vivid-uncond and vivid-sr at batch 1."""
    from vivid_b200.synthetic import synth_batch
    torch.set_num_threads(threads)
    lo, hi = synth_batch([0], 64), synth_batch([0], 256)
    src, g = lo["src_image"] / 127.5 - 1, lo["geometry"]
    ssrc, sg = hi["src_image"] / 127.5 - 1, hi["geometry"]
    x, sx = torch.randn(1, 3, 64, 64) * 5, torch.randn(1, 3, 256, 256) * 5
    sig = torch.full((1,), 5.0)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            nets["base"](src, x, sig, g)
            nets["uncond"](src, x, sig)
            nets["sr"](ssrc, sx, sig, sg, ssrc)
            times.append(time.perf_counter() - t0)
    return times


def cpu_config1_seconds(tiny, sampler, threads):
    """BASELINE.json configs[0] end to end on the host: tiny net (64 ch, 32x32), Heun 8 steps, batch 2."""
    from vivid_b200.synthetic import synth_batch
    torch.set_num_threads(threads)
    b = synth_batch([0, 1], 32)
    src, geom = b["src_image"] / 127.5 - 1, b["geometry"]
    noise = torch.stack([torch.randn(3, 32, 32, generator=torch.Generator().manual_seed(s)) for s in (0, 1)])
    with torch.no_grad():
        t0 = time.perf_counter()
        sampler(tiny, src, noise, labels=geom, num_steps=8)
        return time.perf_counter() - t0


def cpu_baseline(repeats=3, warmup=1, one_thread=True):
    threads = os.cpu_count() or 1
    kind, nets, tiny, sampler = cpu_nets()
    times = cpu_sample_seconds(nets, threads, warmup + repeats)[warmup:]
    secs = sum(times) / len(times)
    calls = 63
    out = dict(value=round(1.0 / (calls * secs), 6), unit="images/s", cores=threads, kind=kind,
               sample="%d (+%d warm-up) samples of one denoiser call each of vivid-base, vivid-uncond, vivid-sr at batch 1 (%s, fp32 "
                      "PyTorch CPU ops), mean %.2f s (min %.2f, max %.2f); one image needs 63 of each"
                      % (len(times), warmup, "the unmodified reference staged in oracle/_ref" if kind == "reference"
                         else "oracle port of the reference algorithm", secs, min(times), max(times)),
               seconds_per_sample=round(secs, 3), gflops=round(sum(ALG_GFLOP.values()) / secs, 1))
    t_cfg1 = cpu_config1_seconds(tiny, sampler, threads)
    out["config1"] = dict(workload="tiny net (64 ch, 32x32) Heun 8 steps, batch 2, end to end", seconds=round(t_cfg1, 2),
                          images_per_s=round(2.0 / t_cfg1, 3), cores=threads)
    if one_thread:
        t1 = cpu_sample_seconds(nets, 1, 1)[0]
        out["one_thread"] = dict(value=round(1.0 / (calls * t1), 6), unit="images/s", cores=1, seconds_per_sample=round(t1, 2))
        torch.set_num_threads(threads)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(repeats=max(args.steps, 1), warmup=max(args.warmup, 1))
    t_step = base["seconds_per_sample"]
    out = dict(impl="reference", metric="guided NVS images/sec (vivid-base+SR)", value=base["value"], unit="images/s",
               n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=args.steps, warmup=args.warmup,
               ms_per_step=round(t_step * 1e3, 1), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
               data="synthetic",
               config=dict(workload="vivid-base guided (vivid-uncond gnet) -> vivid-sr, Heun 32 steps/stage; bounded sample per step: "
                           "one denoiser call of each net at batch 1 on the host CPU", batch_per_gpu=1),
               cpu_baseline=base, e2e=dict(value=base["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(out)


_JSON_OUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 (NCCL's version banner, library
    printf) was redirected to stderr in main()."""
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="images per GPU per step (the reference's --batch option; its default is 32)")
    ap.add_argument("--num-steps", type=int, default=32, help="Heun steps per stage (reference default)")
    ap.add_argument("--guidance", type=float, default=1.5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-insitu", action="store_true", help="skip the extra step recorded with kernel activity records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
