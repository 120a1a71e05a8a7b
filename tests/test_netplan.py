"""Plan recording inside the library (include/vivid_b200.h: vb_net_plan_create / vb_net_plan_trace; SURVEY.md 8(b)
`vb_plan_create(net_desc)` + `vb_plan_set_weights(names, ptrs)`).

CPU: the library's recorder (csrc/netplan.cu) and engine.Plan must emit the SAME sequence of buffer allocations, weight
preparations and ops — compared line for line on dry runs, for every golden case and every preset, fp32 and fp16 parameters,
several batch sizes (the heuristic N tile depends on the batch).  GPU: a library-recorded plan run through vb_denoise / vb_sample
is bit-identical to the Python surface on the same inputs.
"""
import ctypes as C
import os

import pytest
import torch

import cases

PRESETS = {
    "vivid-base": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1),
    "vivid-uncond": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1, uncond=True),
    "vivid-sr": dict(img_resolution=256, img_channels=3, label_dim=20, model_channels=64, super_res=True, noisy_sr=0.25),
    "vivid-base-dual": dict(img_resolution=64, img_channels=3, source_label_dim=20, target_label_dim=40, model_channels=128,
                            extra_attn=1),
    "vivid-base-no-time-enc": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1,
                                   no_time_enc=True),
}


def small_net(case, dev="cpu"):
    import vivid_b200
    net = vivid_b200.NVPrecond(**cases.CASES[case]["cfg"])
    net.load_state_dict(cases.synth_state_dict([(k, tuple(v.shape)) for k, v in net.state_dict().items()]))
    return net.to(dev).eval()


def first_difference(a, b):
    for i, (x, y) in enumerate(zip(a.splitlines(), b.splitlines())):
        if x != y:
            return f"line {i}:\n  engine : {x}\n  library: {y}"
    return f"lengths differ: {a.count(chr(10))} vs {b.count(chr(10))} lines"


@pytest.mark.parametrize("case", list(cases.CASES))
def test_library_recorder_matches_engine_on_golden_cases(case):
    from vivid_b200 import netplan
    net = small_net(case)
    for B in (1, 3):
        a, b = netplan.trace_engine(net, B), netplan.trace_library(net, B)
        assert a == b, first_difference(a, b)
        kinds = {l.split()[0] for l in b.splitlines()}
        assert {"alloc", "wprep", "conv", "attn", "embed", "precond_in", "precond_out", "io", "enc_ops", "fill", "to_f32"} <= kinds


@pytest.mark.parametrize("name", list(PRESETS))
def test_library_recorder_matches_engine_on_presets(name):
    import vivid_b200
    from vivid_b200 import netplan
    torch.manual_seed(3)
    net = vivid_b200.NVPrecond(**PRESETS[name]).eval()
    batches = (1, 4) if "sr" in name else (1, 32, 128)
    for half in (False, True):
        if half:
            net = net.half()                    # persisted EMA snapshots are fp16, parameters and buffers
        for B in batches:
            a, b = netplan.trace_engine(net, B), netplan.trace_library(net, B)
            assert a == b, f"{name} fp16={half} B={B}: " + first_difference(a, b)
    ops = [l for l in b.splitlines() if l.split()[0] in ("conv", "attn", "eltwise", "embed", "precond_in", "precond_out")]
    assert len(ops) == {"vivid-base": 315, "vivid-uncond": 159, "vivid-sr": 150}.get(name, len(ops))


def test_library_recorder_matches_engine_on_random_architectures():
    """Beyond the presets: random constructor arguments (resolution, widths, depth, attention placement, balances, embedding
    widths; conditional / uncond / super_res / dual-source / no_time_enc nets), random batch sizes — the two recorders stay
    identical line for line."""
    import random
    import vivid_b200
    from vivid_b200 import netplan
    rng = random.Random(1234)
    done = 0
    while done < 24:
        R = rng.choice([16, 32, 64])
        mult = rng.choice([[1, 2], [1, 1], [1, 2, 2], [1, 2, 3], [2, 2], [1, 2, 3, 4]])
        if R >> (len(mult) - 1) < 4:
            continue
        attn = [R >> i for i in range(len(mult)) if (R >> i) <= 32 and rng.random() < 0.5]
        flavor = rng.choice(["cond", "uncond", "sr", "dual", "no_time_enc"])
        cfg = dict(img_resolution=R, img_channels=3, model_channels=rng.choice([64, 128]), channel_mult=mult,
                   num_blocks=rng.choice([1, 2, 3]), attn_resolutions=attn, extra_attn=rng.choice([None, 0, 1]),
                   res_balance=rng.choice([0.3, 0.5]), attn_balance=rng.choice([0.3, 0.2]), concat_balance=rng.choice([0.5, 0.3]),
                   label_balance=rng.choice([0.5, 0.25]))
        if rng.random() < 0.3:
            cfg["channel_mult_emb"] = rng.choice([2, 4])
        if rng.random() < 0.3:
            cfg["channel_mult_noise"] = rng.choice([1, 2])
        if flavor == "dual":
            cfg.update(source_label_dim=20, target_label_dim=40)
        else:
            cfg.update(label_dim=20)
        if flavor == "uncond":
            cfg["uncond"] = True
        if flavor == "sr":
            cfg.update(super_res=True, noisy_sr=0.25)
        if flavor == "no_time_enc":
            cfg["no_time_enc"] = True
        net = vivid_b200.NVPrecond(**cfg).eval()
        with torch.no_grad():
            for p in net.parameters():
                if p.ndim == 0:
                    p.fill_(rng.random() + 0.5)
        if rng.random() < 0.3:
            net = net.half()
        B = rng.choice([1, 2, 5, 33])
        a, b = netplan.trace_engine(net, B), netplan.trace_library(net, B)
        assert a == b, f"{cfg} B={B}: " + first_difference(a, b)
        done += 1


def test_library_recorder_rejects_bad_descriptions():
    from vivid_b200 import _lib as L
    from vivid_b200 import netplan
    lib = L.lib()
    net = small_net("v_cond")
    desc = netplan.net_desc(net)
    params, keep = netplan.param_table(net)
    n = len(params)
    assert lib.vb_net_plan_trace(C.byref(desc), params, n, 2, None, 0) > 0
    assert lib.vb_net_plan_trace(C.byref(desc), params, n, 0, None, 0) < 0 and b"batch" in lib.vb_last_error()
    assert lib.vb_net_plan_trace(C.byref(desc), params, n - 40, 2, None, 0) < 0 and b"is missing" in lib.vb_last_error()
    assert lib.vb_net_plan_trace(None, params, n, 2, None, 0) < 0
    bad = netplan.net_desc(net)
    bad.unet.cemb += 64                         # the parameter table no longer fits the layer table
    assert lib.vb_net_plan_trace(C.byref(bad), params, n, 2, None, 0) < 0 and b"layer table expects" in lib.vb_last_error()
    bad = netplan.net_desc(net)
    bad.uncond = 1
    assert lib.vb_net_plan_trace(C.byref(bad), params, n, 2, None, 0) < 0 and b"uncond" in lib.vb_last_error()
    bad = netplan.net_desc(net)
    bad.unet.channels_per_head = 4096
    assert lib.vb_net_plan_trace(C.byref(bad), params, n, 2, None, 0) < 0 and b"head" in lib.vb_last_error()
    # without a device the real recorder fails loudly (no CPU fallback), the dry run above is host arithmetic only
    if not torch.cuda.is_available():
        out = C.c_void_p()
        assert lib.vb_net_plan_create(C.byref(desc), params, n, 2, None, C.byref(out)) != 0 and not out.value
    buf = C.create_string_buffer(64)
    full = lib.vb_net_plan_trace(C.byref(desc), params, n, 2, buf, len(buf))
    assert full > 64 and len(buf.value) == 63                   # truncated, NUL-terminated, full length returned
    del keep


# ------------------------------------------------------------------------------------------------ GPU
def _gpu_env():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (vivid_b200 has no CPU fallback)")
    from vivid_b200 import _lib as L
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    lib = L.lib()
    L.check(lib.vb_device_check(), "vb_device_check")
    return L, lib, torch.device("cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("name,B,half", [("vivid-base", 3, False), ("vivid-uncond", 2, True), ("vivid-sr", 1, False),
                                         ("vivid-base-dual", 2, False), ("v_tiny", 2, False)])
def test_library_recorded_plan_is_bit_identical_to_engine_plan(name, B, half):
    """vb_net_plan_create -> vb_denoise against NVPrecond.forward (engine.Plan) on the same weights and inputs."""
    import vivid_b200
    from vivid_b200 import netplan
    from vivid_b200.synthetic import synth_batch
    L, lib, dev = _gpu_env()
    if name in PRESETS:
        torch.manual_seed(5)
        net = vivid_b200.NVPrecond(**PRESETS[name])
        with torch.no_grad():
            for p in net.parameters():
                if p.ndim == 0:
                    p.fill_(1.0)                  # gains are zero-initialised in the reference
        net = (net.half() if half else net).to(dev).eval()
    else:
        net = small_net(name, dev)
    dual = bool(net.dual)
    R = net.img_resolution
    b = synth_batch(list(range(B)), R, dual=dual)
    src, tgt, geom = (b["src_image"] / 127.5 - 1).to(dev), (b["tgt_image"] / 127.5 - 1).to(dev), b["geometry"].to(dev)
    n_x = src.shape[0]
    g = torch.Generator(device="cpu").manual_seed(3)
    x = (tgt + 2.0 * torch.randn(tgt.shape, generator=g).to(dev)).contiguous()
    sigma = torch.full((n_x,), 2.0, device=dev)
    cond = torch.randn(B, 3, R, R, generator=g).to(dev) if net.super_res else None
    torch.manual_seed(11)
    ref = net(src, x, sigma, geom, cond).clone()
    assert torch.isfinite(ref).all() and ref.abs().max() > 0
    plan = net.plan(B, dev)

    lp = netplan.LibPlan(net, B, dev)
    assert (lp.num_ops, lp.enc_ops, lp.launches) == (plan.num_ops, plan.enc_ops, plan.launches)
    assert lib.vb_workspace_bytes(lp.handle) == plan.owned_bytes
    assert (lp.io.n_x, lp.io.n_out, lp.io.img_elems, lp.io.geom_dim) == (n_x, B, 3 * R * R, plan.in_geom.shape[1])
    torch.manual_seed(11)
    noise = torch.randn_like(cond) if cond is not None else None       # the draw forward() makes internally
    out = torch.full_like(ref, float("nan"))
    st = torch.cuda.current_stream().cuda_stream
    gfull = geom.to(torch.float32).reshape(n_x, -1).contiguous()
    for _ in range(2):                                                  # first call captures the graph, second replays it
        out.fill_(float("nan"))
        L.check(lib.vb_denoise(lp.handle, src.data_ptr(), x.data_ptr(), sigma.data_ptr(), n_x, gfull.data_ptr(), n_x, L.ptr(cond),
                               L.ptr(noise), out.data_ptr(), st), "vb_denoise")
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
    del lp                                                              # frees the plan's device buffers


@pytest.mark.gpu
def test_library_recorded_plans_run_the_whole_sampler():
    """Checkpoint -> vb_net_plan_create (net and guiding net) -> vb_plan_set_inputs -> vb_sample: no Python-side plan, same bits
    as vivid_b200.edm_sampler."""
    import vivid_b200
    from vivid_b200 import netplan
    from vivid_b200.sampler import sigma_steps
    L, lib, dev = _gpu_env()
    B, steps, w = 2, 4, 1.7
    net, gnet = small_net("v_cond", dev), small_net("v_uncond", dev)
    inp = {k: v.to(dev) for k, v in cases.synth_inputs("v_cond", B).items()}
    os.environ["VB_DUAL_STREAM"] = "0"
    try:
        want = vivid_b200.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], gnet=gnet, num_steps=steps, guidance=w)
    finally:
        os.environ.pop("VB_DUAL_STREAM")
    lp, lg = netplan.LibPlan(net, B, dev), netplan.LibPlan(gnet, B, dev)
    st = torch.cuda.current_stream().cuda_stream
    geom = inp["geometry"].to(torch.float32).contiguous()
    L.check(lib.vb_plan_set_inputs(lp.handle, inp["src"].data_ptr(), geom.data_ptr(), B, None, st), "vb_plan_set_inputs")
    L.check(lib.vb_plan_set_inputs(lg.handle, None, None, 0, None, st), "vb_plan_set_inputs")
    t = sigma_steps(steps, 0.002, 80, 7, dev).tolist()
    ws = torch.empty(lib.vb_sample_workspace_bytes(lp.handle) // 4, device=dev)
    out = torch.full_like(inp["noise"], float("nan"))
    d = L.SampleDesc(net=lp.handle, gnet=lg.handle, noise=inp["noise"].data_ptr(), t_steps=(C.c_float * len(t))(*t),
                     workspace=ws.data_ptr(), x_out=out.data_ptr(), num_steps=steps, net_first_op=0, guidance=w)
    L.check(lib.vb_sample(C.byref(d), st), "vb_sample")
    torch.cuda.synchronize()
    assert torch.equal(out, want)


@pytest.mark.gpu
def test_python_surface_runs_on_library_recorded_plans(monkeypatch):
    """VB_LIB_PLAN=1: NVPrecond.plan() hands out plans recorded by the library (netplan.LibPlan: torch views of the plan's own
    device buffers) — forward, return_features / inject_features and the guided sampler give the bits of the engine.py plans.
    (The whole `-m gpu` suite passes under this switch; this test keeps the switch itself covered.)"""
    import vivid_b200
    from vivid_b200 import netplan
    L, lib, dev = _gpu_env()
    B = 2
    net, gnet = small_net("v_cond", dev), small_net("v_uncond", dev)
    inp = {k: v.to(dev) for k, v in cases.synth_inputs("v_cond", B).items()}
    x = inp["tgt"] + 2.0 * inp["noise"]
    sigma = torch.full((B,), 2.0, device=dev)

    def everything():
        d = net(inp["src"], x, sigma, inp["geometry"])
        feats = net(inp["src"], x, sigma, inp["geometry"], return_features=True)
        d2 = net(inp["src"], x, sigma, inp["geometry"], inject_features=feats)
        s = vivid_b200.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], gnet=gnet, num_steps=3, guidance=1.5)
        return [d.clone(), d2.clone(), s.clone()] + [f.clone() for f in feats]

    want = everything()
    assert type(net.plan(B, dev)).__name__ == "Plan"
    monkeypatch.setenv("VB_LIB_PLAN", "1")
    net.invalidate_plans()
    gnet.invalidate_plans()
    got = everything()
    assert isinstance(net.plan(B, dev), netplan.LibPlan) and isinstance(gnet.plan(B, dev), netplan.LibPlan)
    assert len(got) == len(want) > 3
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert torch.equal(want[0], want[1])            # injected features == recomputed features
    net.invalidate_plans()
    gnet.invalidate_plans()


def test_weight_refresh_rejects_foreign_plans():
    from vivid_b200 import _lib as L
    from vivid_b200 import netplan
    lib = L.lib()
    net = small_net("v_uncond")
    params, keep = netplan.param_table(net)
    plan = C.c_void_p()
    assert lib.vb_plan_create(C.byref(plan)) == 0
    try:
        assert lib.vb_net_plan_set_weights(plan, params, len(params), None) != 0 and b"not recorded by vb_net_plan_create" in lib.vb_last_error()
        assert lib.vb_net_plan_set_weights(None, params, len(params), None) != 0
        assert lib.vb_plan_num_features(plan) == 0 and lib.vb_plan_num_features(None) == 0
    finally:
        lib.vb_plan_destroy(plan)
    del keep


@pytest.mark.gpu
def test_library_plan_weight_refresh():
    """vb_net_plan_set_weights: a recorded plan takes the next checkpoint's weights in place (fp32, then the fp16-persisted form) and
    gives the bits of a plan recorded from scratch; a bad table changes nothing."""
    from vivid_b200 import netplan
    L, lib, dev = _gpu_env()
    B = 2
    net = small_net("v_cond", dev)
    inp = {k: v.to(dev) for k, v in cases.synth_inputs("v_cond", B).items()}
    x = (inp["tgt"] + 2.0 * inp["noise"]).contiguous()
    sigma = torch.full((B,), 2.0, device=dev)
    geom = inp["geometry"].to(torch.float32).contiguous()
    st = torch.cuda.current_stream().cuda_stream
    lp = netplan.LibPlan(net, B, dev)

    def lib_call():
        out = torch.full_like(x, float("nan"))
        L.check(lib.vb_denoise(lp.handle, inp["src"].data_ptr(), x.data_ptr(), sigma.data_ptr(), B, geom.data_ptr(), B, None, None,
                               out.data_ptr(), st), "vb_denoise")
        torch.cuda.synchronize()
        return out

    ref1 = net(inp["src"], x, sigma, inp["geometry"]).clone()
    assert torch.equal(lib_call(), ref1)
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    net.load_state_dict(cases.synth_state_dict(shapes, salt=1))              # "the next checkpoint"
    ref2 = net(inp["src"], x, sigma, inp["geometry"]).clone()
    assert not torch.equal(ref2, ref1)
    assert torch.equal(lib_call(), ref1)                                     # the plan still holds the old weights
    lp.set_weights(net)
    assert torch.equal(lib_call(), ref2)
    params, keep = netplan.param_table(net)
    assert lib.vb_net_plan_set_weights(lp.handle, params, len(params) - 40, st) != 0 and b"is missing" in lib.vb_last_error()
    assert torch.equal(lib_call(), ref2)                                     # all or nothing
    del keep
    net = net.half()                                                         # fp16-persisted parameters and buffers
    ref3 = net(inp["src"], x, sigma, inp["geometry"]).clone()
    lp.set_weights(net)
    assert torch.equal(lib_call(), ref3)
