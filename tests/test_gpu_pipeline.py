"""GPU parity tests of the WORKLOAD the headline number times (pytest -m gpu, B200): the SR-stage sampler loop, the
two-stage base -> resize -> SR pipeline at preset size, dual-source vivid-base at full size, fp16-persisted weights,
the dual-source driver path, the rank-0 image gather — and, when the staged reference is present (oracle/_ref, see
oracle/make_ref.py), the CUDA path against the UNMODIFIED reference running its fp32 path on the same GPU.

Tolerances (BASELINE.json north_star): per-call denoiser output rel-L2 <= 1e-2, final image PSNR >= 40 dB.
All product calls go through the public API (NVPrecond / edm_sampler / generate_images_nvs) -> ctypes -> C ABI.
"""
import math
import os

import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (vivid_b200 has no CPU fallback)")
    from vivid_b200 import _lib as L
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    lib = L.lib()
    L.check(lib.vb_device_check(), "vb_device_check")
    return L, lib, torch.device("cuda")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def psnr_u8(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return 10 * math.log10(255.0 ** 2 / max(mse, 1e-12))


PRESETS = {
    "vivid-base": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1),
    "vivid-uncond": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1, uncond=True),
    "vivid-sr": dict(img_resolution=256, img_channels=3, label_dim=20, model_channels=64, super_res=True, noisy_sr=0.25),
    "vivid-base-dual": dict(img_resolution=64, img_channels=3, source_label_dim=20, target_label_dim=40, model_channels=128,
                            extra_attn=1),
}


def make_pair(name, seed, dev, half=False):
    """(product net, oracle net) of a preset with random-init weights and unit gains."""
    import vivid_b200
    from oracle import vivid_oracle as O
    cfg = PRESETS[name]
    torch.manual_seed(seed)
    net = vivid_b200.NVPrecond(**cfg)
    with torch.no_grad():
        for p in net.parameters():
            if p.ndim == 0:
                p.fill_(1.0)
    if half:
        net = net.half()                          # persisted EMA snapshots are fp16
    net = net.to(dev).eval()
    ocfg = dict(cfg, dual=True) if "source_label_dim" in cfg else cfg
    onet = O.OracleNet({k: v.detach().clone() for k, v in net.state_dict().items()}, ocfg)
    return net, onet


def synth(seeds, R, dev, dual=False):
    from vivid_b200.synthetic import synth_batch
    b = synth_batch(seeds, R, dual=dual)
    return (b["src_image"] / 127.5 - 1).to(dev), (b["tgt_image"] / 127.5 - 1).to(dev), b["geometry"].to(dev)


# ------------------------------------------------------------------------------- SR-stage sampler loop
@pytest.mark.parametrize("bound", ["1", "0"])
def test_sr_sampler_loop_vs_oracle_full_size(env, monkeypatch, bound):
    """vivid-sr at 256x256 through edm_sampler(gnet=sr, conditioning_image=...) exactly as the driver's second stage
    calls it (generate_images.py:322-326): every call draws its low-res noise from the GLOBAL generator
    (experiments/code/training/models.py:608-611), so product and oracle are started from the same torch.manual_seed on
    the same device.  Checked per call (rel-L2 <= 1e-2) and on the final image (PSNR >= 40 dB), graphs on, in the
    zero-copy loop (bound=1) and the generic one (bound=0)."""
    import vivid_b200
    from oracle import vivid_oracle as O
    L, lib, dev = env
    monkeypatch.setenv("VB_BOUND_SAMPLER", bound)
    net, onet = make_pair("vivid-sr", 2, dev)
    B, steps = 2, 4
    src, tgt, geom = synth(range(B), 256, dev)
    low = torch.nn.functional.interpolate(torch.nn.functional.interpolate(tgt, size=64, mode="bilinear", antialias=True),
                                          size=256, mode="bilinear")
    noise = vivid_b200.StackedRandomGenerator(dev, range(B)).randn([B, 3, 256, 256], device=dev)
    mine, want = [], []
    torch.manual_seed(1234)
    lat = vivid_b200.edm_sampler(net, src, noise, labels=geom, gnet=net, num_steps=steps, conditioning_image=low, _trace=mine)
    torch.manual_seed(1234)
    with torch.no_grad():
        ref = O.edm_sampler(onet, src, noise, labels=geom, gnet=onet, num_steps=steps, conditioning_image=low, trace=want)
    assert len(mine) == len(want) == 2 * steps - 1
    worst = max(rel(a, b) for a, b in zip(mine, want))
    assert worst <= 1e-2, worst
    assert rel(lat, ref) <= 1e-2
    dec = vivid_b200.StandardRGBEncoder().decode
    assert psnr_u8(dec(lat), O.decode(ref)) >= 40.0
    # same seed, same bits (graph replay, plan buffers reused)
    torch.manual_seed(1234)
    again = vivid_b200.edm_sampler(net, src, noise, labels=geom, gnet=net, num_steps=steps, conditioning_image=low)
    assert torch.equal(again, lat)
    print(f"SR sampler (bound={bound}): worst per-call rel-L2 {worst:.2e}, final {rel(lat, ref):.2e}")


def test_bound_and_generic_sampler_loops_agree_bitwise(env, monkeypatch):
    """The zero-copy loop (vb_heun writes the next call's inputs) and the generic loop (host copies) are the same
    arithmetic: guided base stage, bit for bit."""
    import vivid_b200
    L, lib, dev = env
    net, _ = make_pair("vivid-base", 0, dev)
    gnet, _ = make_pair("vivid-uncond", 1, dev)
    B = 3
    src, tgt, geom = synth(range(B), 64, dev)
    noise = vivid_b200.StackedRandomGenerator(dev, range(B)).randn([B, 3, 64, 64], device=dev)
    out = {}
    for bound in ("1", "0"):
        monkeypatch.setenv("VB_BOUND_SAMPLER", bound)
        out[bound] = vivid_b200.edm_sampler(net, src, noise, labels=geom, gnet=gnet, num_steps=5, guidance=1.5)
    assert torch.equal(out["1"], out["0"])
    # global-generator consumption of the SR noise: normal_() on the plan buffer == torch.randn_like(conditioning_image)
    torch.manual_seed(9)
    a = torch.randn_like(tgt)
    torch.manual_seed(9)
    b = torch.empty_like(tgt).normal_()
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------- the bench workload itself
def test_two_stage_pipeline_vs_oracle_preset_size(env):
    """BASELINE.json configs[1..2] at preset size, the workload bench.py times: vivid-base guided by vivid-uncond (w=1.5),
    32 Heun steps -> bilinear x4 -> vivid-sr, 32 steps, through generate_images_nvs; the oracle replays the same seeds,
    noise, poses, resize and global-generator seeding.  Final uint8 PSNR >= 40 dB for both stages."""
    import vivid_b200
    from oracle import vivid_oracle as O
    from vivid_b200.generate import SyntheticDataset
    L, lib, dev = env
    net, onet = make_pair("vivid-base", 0, dev)
    gnet, ognet = make_pair("vivid-uncond", 1, dev)
    sr, osr = make_pair("vivid-sr", 2, dev)
    seeds, T, w = [0, 1], 32, 1.5
    ds = SyntheticDataset(imsize=64, sr_imsize=256)
    got = list(vivid_b200.generate_images_nvs(net, gnet=gnet, sr_model=sr, seeds=seeds, max_batch_size=2, device=dev,
                                              dataset=ds, verbose=False, num_steps=T, guidance=w))
    base_only = list(vivid_b200.generate_images_nvs(net, gnet=gnet, seeds=seeds, max_batch_size=2, device=dev, dataset=ds,
                                                    verbose=False, num_steps=T, guidance=w))
    assert len(got) == 1 and got[0].images.shape == (2, 3, 256, 256) and got[0].images.dtype == torch.uint8
    # oracle: the same pipeline (generate_images.py:262-326), fp32, TF32 off
    data = ds.batch(seeds)
    src = (data["src_image"] / 127.5 - 1).to(dev)
    noise = O.StackedRandomGenerator(dev, seeds).randn([2, 3, 64, 64], device=dev)
    with torch.no_grad():
        lat = O.edm_sampler(onet, src, noise, labels=data["geometry"].to(dev), gnet=ognet, num_steps=T, guidance=w)
        base_img = O.decode(lat)
        sr_src = (data["sr_src_image"] / 127.5 - 1).to(dev)
        sr_noise = O.StackedRandomGenerator(dev, seeds).randn([2, 3, 256, 256], device=dev)
        low = torch.nn.functional.interpolate(lat, size=256, mode="bilinear")
        torch.manual_seed(seeds[0])
        sr_lat = O.edm_sampler(osr, sr_src, sr_noise, labels=data["sr_geometry"].to(dev), gnet=osr, num_steps=T,
                               conditioning_image=low)
        want = O.decode(sr_lat)
    p_base, p_sr = psnr_u8(base_only[0].images, base_img), psnr_u8(got[0].images, want)
    print(f"two-stage pipeline at preset size: base-stage PSNR {p_base:.1f} dB, final PSNR {p_sr:.1f} dB")
    assert p_base >= 40.0 and p_sr >= 40.0


# ------------------------------------------------------------------------------- dual-source at preset size
def test_dual_source_full_size_vs_oracle(env):
    """Current-tree semantics (training/models.py:628-689) at preset size: 2B interleaved inputs, keys = self | src1 |
    src2 (Sk = 3 Sq = 3072 at 32x32), B outputs; plus a 3-step unguided sampler run."""
    import vivid_b200
    from oracle import vivid_oracle as O
    L, lib, dev = env
    net, onet = make_pair("vivid-base-dual", 4, dev)
    B = 2
    src, tgt, geom = synth(range(B), 64, dev, dual=True)
    g = torch.Generator().manual_seed(3)
    for sg in (20.0, 0.7):
        x = tgt + sg * torch.randn(B, 3, 64, 64, generator=g).to(dev).repeat_interleave(2, dim=0)
        sigma = torch.full((2 * B,), sg, device=dev)
        d = net(src, x, sigma, geom)
        with torch.no_grad():
            ref = onet(src, x, sigma, geom)
        assert d.shape == (B, 3, 64, 64) and rel(d, ref) <= 1e-2, (sg, rel(d, ref))
        c_skip = 0.25 / (sg ** 2 + 0.25)
        assert rel(d - c_skip * x[::2], ref - c_skip * x[::2]) <= 1.5e-2
        d32 = net(src, x, sigma, geom, force_fp32=True)
        assert rel(d32, ref) <= 1e-4
    noise = vivid_b200.StackedRandomGenerator(dev, range(B)).randn([B, 3, 64, 64], device=dev).repeat_interleave(2, dim=0)
    lat = vivid_b200.edm_sampler(net, src, noise, labels=geom, num_steps=3)
    with torch.no_grad():
        want = O.edm_sampler(onet, src, noise, labels=geom, num_steps=3)
    assert lat.shape == want.shape == (B, 3, 64, 64) and rel(lat, want) <= 1e-2


def test_fp16_persisted_weights(env):
    """Persisted EMA snapshots hold fp16 parameters and buffers: the plans read them as they are."""
    L, lib, dev = env
    net, onet = make_pair("vivid-uncond", 1, dev, half=True)
    assert next(net.parameters()).dtype == torch.float16
    src, tgt, geom = synth(range(2), 64, dev)
    x = tgt + 3.0 * torch.randn(tgt.shape, generator=torch.Generator().manual_seed(1)).to(dev)
    sigma = torch.full((2,), 3.0, device=dev)
    d = net(src, x, sigma)
    with torch.no_grad():
        ref = onet(src, x, sigma)
    assert rel(d, ref) <= 1e-2
    lv = net(src, x, sigma, return_logvar=True)[1]
    assert (lv - onet.logvar(sigma)).abs().max() < 2e-3


# ------------------------------------------------------------------------------- dual-source driver path, gather
def test_dual_source_driver_outdir_metrics_and_joint_stats(env, tmp_path):
    """generate_images_nvs with a dual-source net (generate_images.py:268-282): one record row per seed, PNGs of the right
    seed, get_metrics and the joint statistics run, two-stage with a vanilla SR model works."""
    import PIL.Image
    import numpy as np
    import vivid_b200
    from vivid_b200.generate import SyntheticDataset
    L, lib, dev = env
    torch.manual_seed(0)
    small = dict(img_channels=3, model_channels=64, channel_mult=[1, 2], num_blocks=1)
    net = vivid_b200.NVPrecond(img_resolution=16, attn_resolutions=[8], source_label_dim=20, target_label_dim=40, **small)
    sr = vivid_b200.NVPrecond(img_resolution=64, attn_resolutions=[], super_res=True, label_dim=20, **small)
    for m in (net, sr):
        with torch.no_grad():
            for p in m.parameters():
                if p.ndim == 0:
                    p.fill_(0.5)
    ds = SyntheticDataset(imsize=16, sr_imsize=64, dual=True)
    seeds = [11, 12, 13]
    out = str(tmp_path / "png")
    recs = list(vivid_b200.generate_images_nvs(net, seeds=seeds, max_batch_size=8, device=dev, dataset=ds, num_steps=3,
                                               verbose=False, outdir=out, gather_images=True))
    r = recs[0]
    assert r.images.shape == (3, 3, 16, 16) and r.src.shape == (3, 3, 16, 16) and r.tgt.shape == (3, 3, 16, 16)
    assert r.noise.shape[0] == 6 and r.labels.shape == (6, 20)
    assert torch.equal(r.gathered_images, r.images) and r.gathered_seeds == seeds
    data = ds.batch(seeds)
    for k, s in enumerate(seeds):
        png = np.asarray(PIL.Image.open(os.path.join(out, f"tgt_{s:06d}.png")))
        assert np.array_equal(png, data["tgt_image"][2 * k].clip(0, 255).to(torch.uint8).permute(1, 2, 0).numpy())
        png = np.asarray(PIL.Image.open(os.path.join(out, f"sample_{s:06d}.png")))
        assert np.array_equal(png, r.images[k].permute(1, 2, 0).cpu().numpy())
    m = vivid_b200.get_metrics(iter(recs), device=dev)
    assert m["num_images"] == 3 and 3.0 < m["psnr"] < 60.0
    last = None
    for rr, ref in vivid_b200.calculate_stats_for_iterable_nvs(recs, metrics=["fid", "joint_fid", "psnr"], verbose=False,
                                                                device=dev, detectors={"fid": cases.FakeDetector()}):
        last = (rr, ref)
    assert last[0].stats["num_images"] == 3 and last[1].stats["num_images"] == 3
    assert last[0].stats["joint_fid"]["sigma"].shape == (96, 96)
    two = list(vivid_b200.generate_images_nvs(net, seeds=seeds, max_batch_size=8, device=dev, dataset=ds, num_steps=3,
                                              verbose=False, sr_model=sr))
    assert two[0].images.shape == (3, 3, 64, 64) and two[0].tgt.shape == (3, 3, 64, 64)


def test_shard_override_reproduces_rank_shares(env):
    """shard=(rank, world): one rank's share of the seeds without a process group — the union over ranks equals the
    single-process run bit for bit (seed-keyed inputs and noise; bitwise-neutral tuning)."""
    import vivid_b200
    from vivid_b200.generate import SyntheticDataset
    L, lib, dev = env
    torch.manual_seed(0)
    net = vivid_b200.NVPrecond(img_resolution=16, img_channels=3, label_dim=20, model_channels=64, channel_mult=[1, 2],
                               num_blocks=1, attn_resolutions=[8])
    with torch.no_grad():
        for p in net.parameters():
            if p.ndim == 0:
                p.fill_(0.5)
    ds = SyntheticDataset(imsize=16, sr_imsize=64)
    seeds = list(range(40, 47))
    kw = dict(seeds=seeds, max_batch_size=2, device=dev, dataset=ds, num_steps=3, verbose=False)
    whole = {s: img for r in vivid_b200.generate_images_nvs(net, **kw) for s, img in zip(r.seeds, r.images)}
    parts = {}
    for rank in range(2):
        for r in vivid_b200.generate_images_nvs(net, shard=(rank, 2), **kw):
            parts.update({s: img for s, img in zip(r.seeds, r.images)})
    assert sorted(parts) == seeds and all(torch.equal(parts[s], whole[s]) for s in seeds)


# ------------------------------------------------------------------------------- the unmodified reference on the GPU
def _reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py in the build container)")
    return ref_loader.load("snapshot")


def _ref_net(ns, name, seed, dev):
    cfg = PRESETS[name]
    torch.manual_seed(seed)
    ref = ns.models.NVPrecond(use_fp16=False, **cfg)
    with torch.no_grad():
        for p in ref.parameters():
            if p.ndim == 0:
                p.fill_(1.0)
    return ref.to(dev).eval().requires_grad_(False)


def test_against_unmodified_reference_on_gpu(env):
    """The reference's own PyTorch path (snapshot tree, fp32, TF32 off) on the same GPU, identical weights (converted
    with from_reference), noise and poses: per-call rel-L2 <= 1e-2 for vivid-base / vivid-uncond / vivid-sr at preset
    size, <= 1e-4 in fp32 mode, and a guided sampler run through the reference's own edm_sampler: PSNR >= 40 dB."""
    import vivid_b200
    L, lib, dev = env
    ns = _reference()
    refs = {n: _ref_net(ns, n, s, dev) for s, n in enumerate(("vivid-base", "vivid-uncond", "vivid-sr"))}
    mine = {n: vivid_b200.NVPrecond.from_reference(r) for n, r in refs.items()}
    g = torch.Generator().manual_seed(21)
    for name in refs:
        R = PRESETS[name]["img_resolution"]
        B = 2 if R == 64 else 1
        src, tgt, geom = synth(range(B), R, dev)
        for sg in (30.0, 0.4):
            x = tgt + sg * torch.randn(tgt.shape, generator=g).to(dev)
            sigma = torch.full((B,), sg, device=dev)
            kw = dict(conditioning_image=tgt) if name == "vivid-sr" else {}
            torch.manual_seed(5)
            with torch.no_grad():
                want = refs[name](src, x, sigma, geom, **kw)
            torch.manual_seed(5)
            got = mine[name](src, x, sigma, geom, **kw)
            torch.manual_seed(5)
            got32 = mine[name](src, x, sigma, geom, force_fp32=True, **kw)
            assert rel(got, want) <= 1e-2 and rel(got32, want) <= 1e-4, (name, sg, rel(got, want), rel(got32, want))
            print(f"{name} sigma={sg} vs reference on GPU: fp16 path {rel(got, want):.2e}, fp32 mode {rel(got32, want):.2e}")
    B = 2
    src, tgt, geom = synth(range(B), 64, dev)
    noise = vivid_b200.StackedRandomGenerator(dev, range(B)).randn([B, 3, 64, 64], device=dev)
    ref_noise = ns.generate_images.StackedRandomGenerator(dev, range(B)).randn([B, 3, 64, 64], device=dev)
    assert torch.equal(noise, ref_noise)
    with torch.no_grad():
        want = ns.generate_images.edm_sampler(refs["vivid-base"], src, noise, labels=geom, gnet=refs["vivid-uncond"],
                                              num_steps=8, guidance=1.5)
    got = vivid_b200.edm_sampler(mine["vivid-base"], src, noise, labels=geom, gnet=mine["vivid-uncond"], num_steps=8,
                                 guidance=1.5)
    dec = vivid_b200.StandardRGBEncoder().decode
    enc = ns.encoders.StandardRGBEncoder()
    assert psnr_u8(dec(got), enc.decode(want)) >= 40.0 and rel(got, want) <= 1e-2


# ------------------------------------------------------------------------------- whole-call C entry point
@pytest.mark.parametrize("name,B", [("vivid-base", 3), ("vivid-uncond", 2), ("vivid-sr", 1), ("vivid-base-dual", 2)])
def test_vb_denoise_whole_call_entry_point(env, name, B):
    """vb_denoise (include/vivid_b200.h) runs one NVPrecond.forward on a recorded plan from plain device
    pointers: bit-identical to the Python surface's forward on the same inputs, incl. sigma / pose broadcast, a missing
    pose vector, and the SR conditioning noise drawn by the caller (reference: torch.randn_like per call, :608-611)."""
    import ctypes as C
    L, lib, dev = env
    net, _ = make_pair(name, 5, dev)
    dual = "dual" in name
    R = net.img_resolution
    src, tgt, geom = synth(list(range(B)), R, dev, dual=dual)
    n_x = src.shape[0]
    g = torch.Generator(device="cpu").manual_seed(3)
    x = (tgt + 2.0 * torch.randn(tgt.shape, generator=g).to(dev)).contiguous()
    sigma = torch.full((n_x,), 2.0, device=dev)
    cond = torch.randn(B, 3, R, R, generator=g).to(dev) if net.super_res else None
    torch.manual_seed(11)
    ref = net(src, x, sigma, geom, cond).clone()
    plan = net.plan(B, dev)
    assert lib.vb_workspace_bytes(plan.handle) == plan.owned_bytes > 0
    torch.manual_seed(11)
    noise = torch.randn_like(cond) if cond is not None else None       # the draw forward() makes internally
    out = torch.full_like(ref, float("nan"))
    st = torch.cuda.current_stream().cuda_stream

    def call(sig, sig_n, geo, geo_rows):
        L.check(lib.vb_denoise(plan.handle, src.data_ptr(), x.data_ptr(), sig.data_ptr(), sig_n, L.ptr(geo), geo_rows, L.ptr(cond),
                               L.ptr(noise), out.data_ptr(), st), "vb_denoise")
        torch.cuda.synchronize()
        return out.clone()

    gfull = geom.to(torch.float32).reshape(n_x, -1).contiguous()
    assert torch.equal(call(sigma, n_x, gfull, n_x), ref)
    assert torch.equal(call(sigma[:1].contiguous(), 1, gfull, n_x), ref)              # sigma broadcast
    if not dual:                                                                      # no pose vector == zeros (vanilla trees)
        torch.manual_seed(11)
        ref0 = net(src, x, sigma, None, cond).clone()
        assert torch.equal(call(sigma, n_x, None, 0), ref0)
    # argument errors come back as codes, never exceptions from the library
    assert lib.vb_denoise(plan.handle, src.data_ptr(), x.data_ptr(), sigma.data_ptr(), n_x + 1, None, 0, L.ptr(cond), L.ptr(noise),
                          out.data_ptr(), st) != 0
    assert b"sigma_n" in lib.vb_last_error()


# ------------------------------------------------------------------------------- whole-sampler C entry point
def _small_net(case, dev):
    import vivid_b200
    cfg = cases.CASES[case]["cfg"]
    net = vivid_b200.NVPrecond(**cfg)
    net.load_state_dict(cases.synth_state_dict([(k, tuple(v.shape)) for k, v in net.state_dict().items()]))
    return net.to(dev).eval()


@pytest.mark.parametrize("kind", ["guided", "guided_one_stream", "unguided", "sr", "two_steps"])
def test_vb_sample_whole_sampler_entry_point(env, monkeypatch, kind):
    """vb_sample (include/vivid_b200.h; SURVEY.md 8(b) `vb_sample`) enqueues the Heun loop of edm_sampler
    (generate_images.py:72-118) from C over bound plans.  edm_sampler routes through it; it must be bit-identical to the
    Python loop (VB_C_SAMPLER=0: the same replays and vb_heun passes issued from Python), incl. the guiding net on a second
    stream and the SR net's per-call draw from the global generator (experiments/code/training/models.py:608-611) — and
    callable with plain device pointers after vb_plan_set_inputs."""
    import ctypes as C
    import vivid_b200
    L, lib, dev = env
    B, steps = 2, 2 if kind == "two_steps" else 4        # (num_steps = 1 is 0/0 in the reference's schedule, generate_images.py:69)
    sr = kind == "sr"
    net = _small_net("v_sr" if sr else "v_cond", dev)
    gnet = _small_net("v_uncond", dev) if kind.startswith("guided") else None
    inp = {k: v.to(dev) for k, v in cases.synth_inputs("v_sr" if sr else "v_cond", B).items()}
    kw = dict(labels=inp["geometry"], num_steps=steps, guidance=1.7 if gnet is not None else 1,
              conditioning_image=inp["tgt"] if sr else None, gnet=gnet)
    if kind == "guided_one_stream":
        monkeypatch.setenv("VB_DUAL_STREAM", "0")

    def run(c_loop):
        monkeypatch.setenv("VB_C_SAMPLER", "1" if c_loop else "0")
        torch.manual_seed(21)
        return vivid_b200.edm_sampler(net, inp["src"], inp["noise"], **kw).clone()

    want = run(False)
    got = run(True)
    assert torch.isfinite(got).all() and torch.equal(got, want)
    # the traced Python loop (what the per-step parity tests observe) is the same computation
    trace = []
    torch.manual_seed(21)
    assert torch.equal(vivid_b200.edm_sampler(net, inp["src"], inp["noise"], _trace=trace, **kw), want)
    assert len(trace) == 2 * steps - 1

    # raw-pointer call, as a non-Python host would make it
    from vivid_b200.sampler import sigma_steps
    plan = net.plan(B, dev)
    st = torch.cuda.current_stream().cuda_stream
    geom = inp["geometry"].to(torch.float32).contiguous()
    L.check(lib.vb_plan_set_inputs(plan.handle, inp["src"].data_ptr(), geom.data_ptr(), B, inp["tgt"].data_ptr() if sr else None, st),
            "vb_plan_set_inputs")
    gplan = gnet.plan(B, dev) if gnet is not None else None
    if gplan is not None:
        L.check(lib.vb_plan_set_inputs(gplan.handle, inp["src"].data_ptr(), None, 0, None, st), "vb_plan_set_inputs")
    t = sigma_steps(steps, 0.002, 80, 7, dev).tolist()
    ws = torch.empty(lib.vb_sample_workspace_bytes(plan.handle) // 4, device=dev)
    assert ws.numel() == 3 * inp["noise"].numel()
    out = torch.full_like(inp["noise"], float("nan"))
    draws = []

    def draw(user, dst, n, stream):
        draws.append(n)
        plan.in_noise.normal_()
        return 0
    d = L.SampleDesc(net=plan.handle, gnet=gplan.handle if gplan is not None else None, noise=inp["noise"].data_ptr(),
                     t_steps=(C.c_float * len(t))(*t), workspace=ws.data_ptr(), x_out=out.data_ptr(), num_steps=steps,
                     net_first_op=0, guidance=kw["guidance"])
    if sr:
        d.sr_noise = L.NOISE_FN(draw)
    torch.manual_seed(21)
    L.check(lib.vb_sample(C.byref(d), st), "vb_sample")
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    assert len(draws) == (2 * steps - 1 if sr else 0)
    # argument errors come back as codes
    d.num_steps = 0
    assert lib.vb_sample(C.byref(d), st) != 0 and b"num_steps" in lib.vb_last_error()
    if sr:
        d.num_steps, d.sr_noise = steps, L.NOISE_FN(0)
        assert lib.vb_sample(C.byref(d), st) != 0 and b"sr_noise" in lib.vb_last_error()
