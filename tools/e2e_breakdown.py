"""Where does the end-to-end path (generate_images_nvs with host inputs) spend time beyond the resident pipeline?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, vivid_b200
from vivid_b200.generate import SyntheticDataset
dev = torch.device("cuda"); B = 64; T = 32
net, gnet, sr = bench.make_net("vivid-base", 0, dev), bench.make_net("vivid-uncond", 1, dev), bench.make_net("vivid-sr", 2, dev)
enc = vivid_b200.StandardRGBEncoder(); enc.init(dev)
ds = SyntheticDataset(imsize=64, sr_imsize=256)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
host = {}
class HostDataset:
    def batch(self, seeds): return host[tuple(seeds)]
def prep(step):
    seeds = list(range(step * B, (step + 1) * B)); host[tuple(seeds)] = {k: v.pin_memory() for k, v in ds.batch(seeds).items()}; return seeds
def e2e(seeds):
    it = vivid_b200.generate_images_nvs(net, gnet=gnet, sr_model=sr, seeds=seeds, max_batch_size=B, device=dev, dataset=HostDataset(), verbose=False, num_steps=T, guidance=1.5)
    for r in it: out = r.images.cpu()
    return out
def resident(seeds):
    d = {k: v.to(dev) for k, v in ds.batch(seeds).items()}
    rnd = vivid_b200.StackedRandomGenerator(dev, seeds)
    r = dict(src=enc.encode_latents(d["src_image"]), geom=d["geometry"], noise=rnd.randn([B, 3, 64, 64], device=dev), sr_src=enc.encode_latents(d["sr_src_image"]), sr_geom=d["sr_geometry"], sr_noise=vivid_b200.StackedRandomGenerator(dev, seeds).randn([B, 3, 256, 256], device=dev))
    return r
def pipeline(r):
    lat = vivid_b200.edm_sampler(net, r["src"], r["noise"], labels=r["geom"], gnet=gnet, num_steps=T, guidance=1.5)
    low = torch.nn.functional.interpolate(lat, size=256, mode="bilinear")
    sl = vivid_b200.edm_sampler(sr, r["sr_src"], r["sr_noise"], labels=r["sr_geom"], gnet=sr, num_steps=T, conditioning_image=low)
    return enc.decode(sl)
s0 = prep(0); s1 = prep(1); s2 = prep(2)
e2e(s0)                                  # warm-up (plans, tuning)
r1 = resident(s1); pipeline(r1)
for name, fn, arg in (("e2e", e2e, s1), ("resident", pipeline, r1), ("e2e", e2e, s2), ("resident", pipeline, r1)):
    t0 = sync(); fn(arg); t1 = sync(); print(f"{name}: {t1 - t0:.3f} s", flush=True)
# pieces
t0 = sync(); rnd = vivid_b200.StackedRandomGenerator(dev, s1); a = rnd.randn([B, 3, 64, 64], device=dev); b = vivid_b200.StackedRandomGenerator(dev, s1).randn([B, 3, 256, 256], device=dev); t1 = sync(); print(f"noise (2 x 64 generators): {t1 - t0:.4f} s")
t0 = sync(); d = {k: v.to(dev, non_blocking=True) for k, v in host[tuple(s1)].items()}; t1 = sync(); print(f"H2D of the batch: {t1 - t0:.4f} s")
t0 = sync(); x = r1["sr_noise"]; y = rnd.randn_like(x) if False else None; t1 = sync()
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); e2e(s2); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
