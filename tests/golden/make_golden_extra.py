"""Golden vectors for the two optional arguments of the reference's NVPrecond.forward that sit on the guided denoising path
(SURVEY.md §8(f) N4), produced by the UNMODIFIED reference in the build container (reads /root/reference):
  logvar   : NVPrecond(..., return_logvar=True) -> u(sigma) = logvar_linear(logvar_fourier(ln(sigma)/4))
             (snapshot tree training/models.py:746-747; current tree :686-688 reads c_noise[::2])
  features : a no_time_enc net: return_features=True, then edm_sampler, which runs the source-view encoder once and
             injects its maps at every step (generate_images.py:52-57; training/models.py:664-672)
One process per tree because their module names collide.  Writes / updates tests/golden/extra.pt:
    python tests/golden/make_golden_extra.py vanilla && python tests/golden/make_golden_extra.py dual"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

mode = sys.argv[1]
assert mode in ("vanilla", "dual")
for m in ("kornia", "litdata"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path[:0] = ["/root/reference/experiments/code", "/root/reference"] if mode == "vanilla" else ["/root/reference"]
import training.models as M  # noqa: E402
import generate_images as G  # noqa: E402

path = os.path.join(HERE, "extra.pt")
out = torch.load(path) if os.path.exists(path) else {}
case = "v_cond" if mode == "vanilla" else "d_cond"


def build(**extra):
    cfg = dict(cases.CASES[case]["cfg"], **extra)
    net = M.NVPrecond(use_fp16=False, **cfg).eval()
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    net.load_state_dict(cases.synth_state_dict(shapes))
    return net, shapes, cfg


with torch.no_grad():
    # ---- uncertainty head
    net, shapes, cfg = build()
    B = 3
    inp = cases.synth_inputs(case, B)
    sigma = torch.tensor([80.0, 1.7, 0.02])
    if mode == "dual":   # 2B interleaved; the odd entries differ so that the [::2] selection is observable
        sigma = torch.stack([sigma, sigma * 1.5], dim=1).reshape(-1)
    x = inp["tgt"] + sigma.reshape(-1, 1, 1, 1) * inp["noise"]
    d, lv = net(inp["src"], x, sigma, inp["geometry"], return_logvar=True)
    out[f"logvar_{case}"] = dict(B=B, sigma=sigma, logvar=lv.clone(), D=d.clone(), shapes=shapes, cfg=cfg)
    print(case, "logvar", lv.flatten().tolist())

    # ---- no_time_enc: cached source-view features
    net, shapes, cfg = build(no_time_enc=True)
    B = 2
    inp = cases.synth_inputs(case, B)
    n_in = inp["src"].shape[0]
    feats = net(inp["src"], torch.zeros_like(inp["src"]), torch.ones(n_in), inp["geometry"], None, return_features=True)
    sg = 2.5
    x = inp["tgt"] + sg * inp["noise"]
    d_inj = net(torch.zeros_like(inp["src"]), x, torch.full((n_in,), sg), inp["geometry"], inject_features=feats)
    d_full = net(inp["src"], x, torch.full((n_in,), sg), inp["geometry"])
    assert torch.allclose(d_inj, d_full, atol=1e-5)
    lat = G.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=3)
    out[f"features_{case}"] = dict(B=B, shapes=shapes, cfg=cfg, features=[f.clone() for f in feats], sigma=sg, D=d_full.clone(),
                                   latents=lat.clone(), num_steps=3)
    print(case, "features", [tuple(f.shape) for f in feats], "latents", tuple(lat.shape))
with torch.no_grad():
    # ---- stochastic sampler branch (S_churn > 0, generate_images.py:77-84)
    net, shapes, cfg = build()
    B = 2
    inp = cases.synth_inputs(case, B)
    kw = dict(num_steps=4, S_churn=8, S_min=0.05, S_max=50, S_noise=1.003)
    torch.manual_seed(77)                         # randn_like is the default torch.randn_like: global CPU generator
    lat = G.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], **kw)
    out[f"churn_{case}"] = dict(B=B, shapes=shapes, cfg=cfg, latents=lat.clone(), seed=77, kwargs=kw)
    print(case, "churn latents", tuple(lat.shape), float(lat.abs().mean()))
torch.save(out, path)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
