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
        return dict(src=enc.encode_latents(d["src_image"]), geom=d["geometry"], noise=rnd.randn([B, 3, 64, 64], device=dev),
                    sr_src=enc.encode_latents(d["sr_src_image"]), sr_geom=d["sr_geometry"],
                    sr_noise=vivid_b200.StackedRandomGenerator(dev, seeds).randn([B, 3, 256, 256], device=dev))

    def pipeline(r):
        lat = vivid_b200.edm_sampler(net, r["src"], r["noise"], labels=r["geom"], gnet=gnet, num_steps=T, guidance=args.guidance)
        low = torch.nn.functional.interpolate(lat, size=256, mode="bilinear")
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

    # ---- e2e: public driver, pinned host inputs, H2D + D2H inside the timed region
    host_items = {}

    class HostDataset:
        def batch(self, seeds):
            return host_items[tuple(seeds)]

    def e2e_step(step):
        seeds = list(range(step * world * B, (step + 1) * world * B))       # the driver shards these over the ranks itself
        it = vivid_b200.generate_images_nvs(net, gnet=gnet, sr_model=sr, seeds=seeds, max_batch_size=B, device=dev,
                                            dataset=HostDataset(), verbose=False, num_steps=T, guidance=args.guidance)
        out = None
        for r in it:
            out = r.images.cpu()        # device -> host read of the step's result
        return out

    prep_count = [0]

    def prep_host(seeds):
        host_items[tuple(seeds)] = host_batch(seeds)
        prep_count[0] += 1
        return prep_count[0] - 1

    e2e_ms, e2e_clocks = timed(e2e_step, max(1, args.warmup // 3), args.steps, prep_host) if not args.no_e2e else (None, None)
    h2d = B * (3 * 64 * 64 * 4 * 2 + 20 * 4 + 3 * 256 * 256 * 4 * 2 + 20 * 4)
    d2h = B * 3 * 256 * 256

    # ---- roofline: every recorded op timed alone with CUDA events (eager replay on the launching stream)
    calls = 2 * T - 1
    plans = {"vivid-base": net.plan(B, dev), "vivid-uncond": gnet.plan(B, dev), "vivid-sr": sr.plan(B, dev)}
    agg = {}
    for name, p in plans.items():
        for kind, label, fl, by, op_ms in p.profile(repeats=2):
            a = agg.setdefault(kind, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            a["ms"] += op_ms * calls
            a["flops"] += fl * calls
            a["bytes"] += by * calls
            a["launches"] += calls
    pk = peaks()
    conv_ms = agg["conv3"]["ms"] + agg["conv1"]["ms"]
    conv_fl = agg["conv3"]["flops"] + agg["conv1"]["flops"]
    conv_n = agg["conv3"]["launches"] + agg["conv1"]["launches"]
    total_ms = sum(a["ms"] for a in agg.values())
    achieved = conv_fl / (conv_ms / 1e3) / 1e12
    conv_by = agg["conv3"]["bytes"] + agg["conv1"]["bytes"]
    traffic, traffic_src = None, None
    try:        # DRAM bytes per conv launch from the committed ncu launch list of one call per net at this batch
        with open(os.path.join(ROOT, "profiles", f"r01_launches_B{B}.json")) as f:
            tj = json.load(f)
        if tj.get("batch") == B:
            traffic, traffic_src = round(tj["dram_bytes_per_launch"]), "profiles/r01_launches_B%d.json: %s" % (B, tj["source"])
    except Exception:
        pass
    roofline = dict(bound="tensor", kernel="conv_gemm_kernel (tcgen05 implicit-GEMM 3x3/1x1 conv)", achieved=round(achieved, 1),
                    peak=pk["tflops"], unit="TFLOP/s", frac=round(achieved / pk["tflops"], 4),
                    peak_source=f"{pk['source']} bf16 sustained (kernel timed inside a long step)", traffic=traffic,
                    traffic_unit="bytes per launch (dram read+write, ncu)", traffic_source=traffic_src,
                    alg_bytes_per_launch=round(conv_by / conv_n),
                    launches_per_step=conv_n, avg_launch_us=round(conv_ms / conv_n * 1e3, 2),
                    share_of_step=round(conv_ms / total_ms, 4),
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
        kernels[kind] = k
    alg_tflop_img = calls * (ALG_GFLOP["vivid-base"] + ALG_GFLOP["vivid-uncond"] + ALG_GFLOP["vivid-sr"]) / 1e3
    launches = calls * sum(p.launches for p in plans.values()) + 2 * calls + 1     # + Heun passes + decode

    out = dict(metric="guided NVS images/sec (vivid-base+SR)", value=round(value, 3), unit="images/s", n_gpus=world,
               steps=args.steps, warmup=args.warmup, ms_per_step=round(ms / args.steps, 2), higher_is_better=True,
               scaling="weak", vs_baseline=None, dtype="fp16", data="synthetic",
               config=dict(workload="vivid-base guided (vivid-uncond gnet, w=%.1f) -> bilinear x4 -> vivid-sr; Heun %d steps/stage "
                           "(%d denoiser calls each); random-init weights" % (args.guidance, T, calls),
                           batch_per_gpu=B, global_batch=B * world, parallelism=f"sample-sharded x{world}, no data-path collective",
                           l2="per-step working set (weights 0.84 GB + activations) exceeds the 126 MB L2; no flush needed",
                           operands="fp16 (the reference's own reduced precision; tcgen05 kind::f16)", accumulate="fp32",
                           residual_stream="fp16", sampler_state="fp32"),
               clocks=clocks, gpu_launches=int(launches * args.steps),
               e2e=None if e2e_ms is None else dict(value=round(world * B * args.steps / (e2e_ms / 1e3), 3), unit="images/s",
                                                    h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, clocks=e2e_clocks),
               roofline=roofline, kernels=kernels,
               model_tflops=round(value / world * alg_tflop_img, 1),
               model_frac_of_peak=round(value / world * alg_tflop_img / pk["tflops"], 4))
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(dict(base=net, uncond=gnet, sr=sr))
    if rank == 0:
        emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


# --------------------------------------------------------------------------------------------- CPU arms
def cpu_sample_seconds(state, threads, repeats=1):
    """One denoiser call of vivid-base, vivid-uncond and vivid-sr at B=1 with the oracle (the reference's algorithm,
    fp32, PyTorch CPU ops, all host threads).  Returns seconds for the three calls."""
    from oracle import vivid_oracle as O
    from vivid_b200.synthetic import synth_batch
    torch.set_num_threads(threads)
    times = []
    nets = {k: O.OracleNet(sd, dict(PRESETS[name])) for k, (name, sd) in state.items()}
    lo, hi = synth_batch([0], 64), synth_batch([0], 256)
    src, g = lo["src_image"] / 127.5 - 1, lo["geometry"]
    ssrc, sg = hi["src_image"] / 127.5 - 1, hi["geometry"]
    x, sx = torch.randn(1, 3, 64, 64) * 5, torch.randn(1, 3, 256, 256) * 5
    sig = torch.full((1,), 5.0)
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            nets["base"](src, x, sig, g)
            nets["uncond"](src, x, sig)
            nets["sr"](ssrc, sx, sig, sg, ssrc)
            times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(gpu_nets=None, repeats=1, warmup=0):
    threads = os.cpu_count() or 1
    if gpu_nets is not None:
        state = {k: (n, {kk: v.detach().float().cpu() for kk, v in gpu_nets[k].state_dict().items()})
                 for k, n in (("base", "vivid-base"), ("uncond", "vivid-uncond"), ("sr", "vivid-sr"))}
    else:
        import vivid_b200
        state = {}
        for i, (k, n) in enumerate((("base", "vivid-base"), ("uncond", "vivid-uncond"), ("sr", "vivid-sr"))):
            torch.manual_seed(i)
            m = vivid_b200.NVPrecond(**PRESETS[n])
            with torch.no_grad():
                for p in m.parameters():
                    if p.ndim == 0:
                        p.fill_(1.0)
            state[k] = (n, m.state_dict())
    times = cpu_sample_seconds(state, threads, warmup + repeats)[warmup:]
    secs = sum(times) / len(times)
    calls = 63
    return dict(value=round(1.0 / (calls * secs), 6), unit="images/s", cores=threads, kind="port",
                sample="1 denoiser call each of vivid-base, vivid-uncond, vivid-sr at batch 1 (oracle = reference algorithm "
                       "in fp32 PyTorch CPU ops), %.2f s; one image needs 63 of each" % secs, seconds_per_sample=round(secs, 3))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(repeats=args.steps, warmup=args.warmup)
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
