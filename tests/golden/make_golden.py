"""Generate the golden fixtures by running the UNMODIFIED reference (read-only /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py vanilla   # snapshot tree experiments/code  -> tests/golden/vanilla.pt
    python tests/golden/make_golden.py dual      # current tree                    -> tests/golden/dual.pt
The two trees define colliding module names, hence one process per mode.
kornia and litdata are absent here and not touched on this path; they are stubbed.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

mode = sys.argv[1]
assert mode in ("vanilla", "dual")
for m in ("kornia", "litdata"):
    sys.modules[m] = types.ModuleType(m)
if mode == "vanilla":
    sys.path[:0] = ["/root/reference/experiments/code", "/root/reference"]
else:
    sys.path[:0] = ["/root/reference"]

import training.models as M  # noqa: E402
import training.encoders as E  # noqa: E402
import training.utils as U  # noqa: E402
import generate_images as G  # noqa: E402

torch.manual_seed(0)
out = {"torch_version": torch.__version__, "mode": mode}


def build(case, salt=0):
    cfg = cases.CASES[case]["cfg"]
    net = M.NVPrecond(use_fp16=False, **cfg).eval()
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    net.load_state_dict(cases.synth_state_dict(shapes, salt))
    return net, shapes


# ------------------------------------------------------------------ op-level vectors (a1-a7, a17, a18)
g = torch.Generator().manual_seed(7)
ops = {}
x = torch.randn(2, 8, 6, 6, generator=g)
ops["x"] = x
ops["normalize_dim1"] = M.normalize(x, dim=1)
ops["normalize_all"] = M.normalize(x)
ops["resample_down"] = M.resample(x, f=[1, 1], mode="down")
ops["resample_up"] = M.resample(x, f=[1, 1], mode="up")
ops["mp_silu"] = M.mp_silu(x)
y = torch.randn(2, 8, 6, 6, generator=g)
ops["y"] = y
ops["mp_sum_03"] = M.mp_sum(x, y, t=0.3)
ops["mp_cat"] = M.mp_cat(x, y[:, :4], t=0.5)
w3 = torch.randn(5, 8, 3, 3, generator=g)
w1 = torch.randn(5, 8, generator=g)
ops["w3"], ops["w1"] = w3, w1
conv = M.MPConv(8, 5, kernel=[3, 3]).eval()
conv.weight.data.copy_(w3)
ops["mpconv3"] = conv(x, gain=0.7).detach()
lin = M.MPConv(8, 5, kernel=[]).eval()
lin.weight.data.copy_(w1)
ops["mplinear"] = lin(x[:, :, 0, 0]).detach()
four = M.MPFourier(16)
ops["freqs"], ops["phases"] = four.freqs.clone(), four.phases.clone()
ops["fourier_in"] = torch.tensor([0.3, -1.2, 2.5])
ops["fourier"] = four(ops["fourier_in"])
enc = E.StandardRGBEncoder()
u8 = torch.randint(0, 256, (2, 3, 4, 4), generator=g, dtype=torch.uint8)
ops["u8"] = u8
ops["encode_latents"] = enc.encode_latents(u8)
lat = torch.randn(2, 3, 4, 4, generator=g) * 0.8
ops["lat"] = lat
ops["decode"] = enc.decode(lat)
ext = torch.randn(3, 3, 4, generator=g)
k_src = torch.tensor([[60.0, 61.0, 32.0, 32.0]] * 3) + torch.randn(3, 4, generator=g)
k_tgt = torch.tensor([[55.0, 56.0, 32.0, 32.0]] * 3) + torch.randn(3, 4, generator=g)
ops["ext"], ops["k_src"], ops["k_tgt"] = ext, k_src, k_tgt
if mode == "dual":   # current tree takes [fx,fy,cx,cy] vectors
    ops["compose_geometry_64"] = U.compose_geometry(ext, k_src, k_tgt, imsize=64)
    ops["compose_geometry_256"] = U.compose_geometry(ext, k_src * 4, k_tgt * 4, imsize=256)
else:                # snapshot takes 3x3 K matrices
    ops["compose_geometry_64"] = U.compose_geometry(ext, U.decompose_K(k_src), U.decompose_K(k_tgt), imsize=64)
    ops["compose_geometry_256"] = U.compose_geometry(ext, U.decompose_K(k_src * 4), U.decompose_K(k_tgt * 4), imsize=256)
rnd = G.StackedRandomGenerator("cpu", [3, 4, (1 << 32) + 3])
ops["stacked_randn"] = rnd.randn([3, 2, 4])
out["ops"] = ops

# ------------------------------------------------------------------ network-level vectors
nets = {}
with torch.no_grad():
    for case, spec in cases.CASES.items():
        if spec["mode"] != mode:
            continue
        net, shapes = build(case)
        B = 2
        inp = cases.synth_inputs(case, B)
        rec = {"shapes": shapes, "cfg": spec["cfg"], "B": B, "D": {}}
        n_in = inp["src"].shape[0]
        sigmas = cases.SIGMAS if case != "v_tiny" else (5.0,)
        for sg in sigmas:
            x = inp["tgt"] + sg * inp["noise"]
            sigma = torch.full((n_in,), sg)
            kw = {}
            if spec["cfg"].get("super_res"):
                torch.manual_seed(123)                       # the SR forward draws from the global RNG (F7)
                kw["conditioning_image"] = inp["tgt"]
            rec["D"][sg] = net(inp["src"], x, sigma, inp["geometry"], **kw)
            if mode == "vanilla" and not spec["cfg"].get("super_res") and sg == 5.0:
                rec["D_nogeom"] = net(inp["src"], x, sigma)    # gnet-style call: geometry=None
        nets[case] = rec

    # ------------------------------------------------------------------ sampler-level vectors
    def traced(net, log):
        def call(*a, **k):
            d = net(*a, **k)
            log.append(d.clone())
            return d
        for attr in ("no_time_enc", "img_resolution", "img_channels"):
            setattr(call, attr, getattr(net, attr))
        return call

    if mode == "vanilla":
        net, _ = build("v_cond")
        gnet, _ = build("v_uncond")
        inp = cases.synth_inputs("v_cond", 2)
        log = []
        lat = G.edm_sampler(traced(net, log), inp["src"], inp["noise"], labels=inp["geometry"], gnet=gnet, num_steps=4,
                            guidance=1.5)
        nets["sampler_guided"] = dict(latents=lat, net_calls=torch.stack(log), num_steps=4, guidance=1.5,
                                      images=E.StandardRGBEncoder().decode(lat))
        lat1 = G.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=4, guidance=1)
        nets["sampler_unguided"] = dict(latents=lat1, num_steps=4)
        sr, _ = build("v_sr")
        inp = cases.synth_inputs("v_sr", 2)
        torch.manual_seed(321)
        lat2 = G.edm_sampler(sr, inp["src"], inp["noise"], labels=inp["geometry"], gnet=sr, num_steps=3,
                             conditioning_image=inp["tgt"])
        nets["sampler_sr"] = dict(latents=lat2, num_steps=3, seed=321)
        tiny, _ = build("v_tiny")
        inp = cases.synth_inputs("v_tiny", 2)
        lat3 = G.edm_sampler(tiny, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=8)
        nets["sampler_tiny"] = dict(latents=lat3, num_steps=8)
    else:
        net, _ = build("d_cond")
        inp = cases.synth_inputs("d_cond", 2)
        lat = G.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=3)
        nets["sampler_dual"] = dict(latents=lat, num_steps=3)
    step = torch.arange(32, dtype=torch.float32)
    nets["t_steps_32"] = (80 ** (1 / 7) + step / 31 * (0.002 ** (1 / 7) - 80 ** (1 / 7))) ** 7
out["nets"] = nets

path = os.path.join(HERE, f"{mode}.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
