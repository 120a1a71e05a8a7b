import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vivid_b200
from vivid_b200.generate import SyntheticDataset
dev = torch.device("cuda")
torch.manual_seed(0)
small = dict(img_channels=3, label_dim=20, model_channels=64, channel_mult=[1, 2], num_blocks=1)
net = vivid_b200.NVPrecond(img_resolution=16, attn_resolutions=[8], **small)
gnet = vivid_b200.NVPrecond(img_resolution=16, attn_resolutions=[8], uncond=True, **small)
for m in (net, gnet):
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 0:
                p.fill_(0.5)
ds = SyntheticDataset(imsize=16, sr_imsize=64)
kw = dict(gnet=gnet, device=dev, dataset=ds, num_steps=4, guidance=1.5, verbose=False)
a = list(vivid_b200.generate_images_nvs(net, seeds=[3, 4, 5, 6, 7], max_batch_size=8, **kw))
b = list(vivid_b200.generate_images_nvs(net, seeds=[3, 4, 5, 6, 7], max_batch_size=2, **kw))
d = (a[0].images.int() - torch.cat([r.images for r in b]).int()).abs()
print("max", d.max().item(), "mean", d.float().mean().item(), "count>1", (d > 1).sum().item(), "per-image max", d.flatten(1).max(dim=1).values.tolist())
# per-call denoiser comparison at batch 5 vs batch 2 on the same inputs
from vivid_b200.synthetic import synth_batch
bt = synth_batch([3, 4, 5, 6, 7], 16)
src = (bt["src_image"] / 127.5 - 1).to(dev); g = bt["geometry"].to(dev)
x = torch.randn(5, 3, 16, 16, device=dev) * 5; sig = torch.full((5,), 5.0, device=dev)
net = net.to(dev).eval()
full = net(src, x, sig, g).clone()
part = torch.cat([net(src[i:i+2], x[i:i+2], sig[i:i+2], g[i:i+2]).clone() for i in (0, 2)] + [net(src[4:5], x[4:5], sig[4:5], g[4:5]).clone()])
print("denoiser rel diff batch5 vs split:", ((full - part).norm() / full.norm()).item(), "per-sample", [((full[i]-part[i]).norm()/full[i].norm()).item() for i in range(5)])
