"""GPU bring-up: vivid_b200.NVPrecond (CUDA plan) vs the oracle (torch fp32, TF32 off) on the same weights/inputs."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vivid_oracle as O  # noqa: E402
from vivid_b200.precond import NVPrecond  # noqa: E402
from vivid_b200.synthetic import synth_batch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def set_gains(net, value=1.0):
    with torch.no_grad():
        for n, p in net.named_parameters():
            if p.ndim == 0:
                p.fill_(value)


def check(name, cfg, B, sigmas=(80.0, 5.0, 0.5, 0.05), with_geom=True, graph=False):
    torch.manual_seed(0)
    net = NVPrecond(**cfg)
    set_gains(net)
    net = net.to(dev).eval()
    net.use_graph = graph
    R = cfg["img_resolution"]
    dual = "source_label_dim" in cfg
    ocfg = dict(cfg, dual=dual)
    onet = O.OracleNet({k: v.detach().clone() for k, v in net.state_dict().items()}, ocfg)
    batch = synth_batch(range(B), R, dual=dual)
    src = (batch["src_image"] / 127.5 - 1).to(dev)
    geom = batch["geometry"].to(dev) if with_geom else None
    n_in = src.shape[0]
    g = torch.Generator().manual_seed(1)
    ok = True
    for sg in sigmas:
        tgt = (batch["tgt_image"] / 127.5 - 1)
        x = (tgt + sg * torch.randn(tgt.shape, generator=g)).to(dev)
        sigma = torch.full((n_in,), sg, device=dev)
        cond = None
        kw = {}
        if cfg.get("super_res"):
            cond = (batch["tgt_image"] / 127.5 - 1).to(dev)
            torch.manual_seed(123)
        t0 = time.time()
        d = net(src, x, sigma, geom, cond)
        torch.cuda.synchronize()
        t1 = time.time()
        if cfg.get("super_res"):
            torch.manual_seed(123)
        with torch.no_grad():
            ref = onet(src, x, sigma, geom, cond)
        e = rel(d, ref)
        # the part that actually exercises the network: F_x = (D - c_skip x)/c_out
        c_skip = 0.25 / (sg ** 2 + 0.25)
        xs = x[::2] if dual else x
        ef = rel(d - c_skip * xs, ref - c_skip * xs)
        good = e <= 1e-2 and e == e
        ok &= good
        print(f"{'PASS' if good else 'FAIL'} {name} B{B} sigma={sg}: D rel_l2={e:.3e}  F-part rel_l2={ef:.3e}  ({(t1-t0)*1e3:.1f} ms)", flush=True)
    p = net.plan(B, dev)
    print(f"   plan: {p.num_ops} ops, {p.launches} launches, alg {p.alg_flops/B/1e9:.2f} GFLOP/img, padded {p.padded_flops/B/1e9:.2f}")
    return ok


ok = True
tiny = dict(img_resolution=32, img_channels=3, label_dim=20, model_channels=64)
ok &= check("tiny-vanilla", tiny, 2)
ok &= check("tiny-vanilla-graph", tiny, 3, sigmas=(5.0,), graph=True)
ok &= check("tiny-nogeom", tiny, 2, sigmas=(5.0,), with_geom=False)
ok &= check("tiny-uncond", dict(tiny, uncond=True), 2, sigmas=(5.0, 0.5), with_geom=False)
ok &= check("tiny-dual", dict(img_resolution=32, img_channels=3, source_label_dim=20, target_label_dim=40, model_channels=64), 2,
            sigmas=(5.0, 0.5))
ok &= check("tiny-sr", dict(img_resolution=32, img_channels=3, label_dim=20, model_channels=64, super_res=True, noisy_sr=0.25,
                            channel_mult=[1, 2], attn_resolutions=[]), 2, sigmas=(5.0, 0.5))
if "--full" in sys.argv:
    base = dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1)
    ok &= check("vivid-base", base, 4, sigmas=(5.0, 0.5))
    ok &= check("vivid-uncond", dict(base, uncond=True), 4, sigmas=(5.0,), with_geom=False)
    ok &= check("vivid-sr", dict(img_resolution=256, img_channels=3, label_dim=20, model_channels=64, super_res=True,
                                 noisy_sr=0.25), 2, sigmas=(5.0,))
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
