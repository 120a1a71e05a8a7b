"""Small driver that touches the validation (fp32), statistics, resize and logvar kernels plus one tiny production forward
per golden case on ragged shapes — a quick end-to-end sanity run (compute-sanitizer is closed on this pool, so bad accesses
are hunted with small cases and the comparisons in tests/ instead)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import cases  # noqa: E402
import vivid_b200  # noqa: E402
from vivid_b200 import metrics as M  # noqa: E402

dev = torch.device("cuda")
for case in ("v_cond", "d_cond", "v_sr"):
    cfg = cases.CASES[case]["cfg"]
    net = vivid_b200.NVPrecond(**cfg)
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    net.load_state_dict(cases.synth_state_dict(shapes))
    net = net.to(dev).eval()
    net.use_graph = False
    inp = {k: v.to(dev) for k, v in cases.synth_inputs(case, 3).items()}
    n = inp["src"].shape[0]
    x = inp["tgt"] + 2.0 * inp["noise"]
    sigma = torch.full((n,), 2.0, device=dev)
    kw = dict(conditioning_image=inp["tgt"]) if cfg.get("super_res") else {}
    torch.manual_seed(1)          # the SR forward draws its conditioning noise from the global generator
    a = net(inp["src"], x, sigma, inp["geometry"], **kw)
    torch.manual_seed(1)
    b = net(inp["src"], x, sigma, inp["geometry"], force_fp32=True, **kw)
    if not cfg.get("super_res"):
        _, lv = net(inp["src"], x, sigma, inp["geometry"], return_logvar=True)
        feats = net(inp["src"], x, sigma, inp["geometry"], return_features=True)
        net(inp["src"], x, sigma, inp["geometry"], inject_features=feats)
    print(case, "ok", float((a - b).norm() / b.norm()))
g = torch.Generator().manual_seed(0)
mu = torch.zeros(130, dtype=torch.float64, device=dev)
sg = torch.zeros(130, 130, dtype=torch.float64, device=dev)
M.stats_update(mu, sg, torch.randn(37, 100, generator=g).to(dev), torch.randn(37, 30, generator=g).to(dev))
img = torch.randint(0, 256, (5, 3, 17, 19), generator=g, dtype=torch.uint8).to(dev)
M.psnr_u8(img, img.float() + 1.0, torch.zeros(1, dtype=torch.float64, device=dev))
M.resize_bilinear(torch.rand(2, 3, 24, 40, generator=g).to(dev), (96, 100))
M.resize_bilinear(torch.rand(2, 3, 40, 24, generator=g).to(dev), (10, 7), antialias=True)
torch.cuda.synchronize()
print("all ok")
