"""CPU tests (no GPU): the C-ABI library loads and exports what include/vivid_b200.h declares, the
host-side mirror of the reference interface has the reference's parameter tree / error behaviour,
and the seed sharding + metric reduction work at world_size 2 over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vivid_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "vivid_b200.h")).read()
    declared = set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    raw = ctypes.CDLL(os.path.join(ROOT, "vivid_b200", "libvividb200.so"))
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/vivid_b200.h but not exported"
    from vivid_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_abi_version_and_struct_mirror(lib):
    from vivid_b200 import _lib
    assert lib.vb_abi_version() == 2
    for i, st in enumerate(_lib.STRUCTS):
        assert lib.vb_struct_size(i) == ctypes.sizeof(st)
    assert lib.vb_struct_size(99) == -1


def test_fails_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.vb_device_check() != 0
    assert b"no CPU fallback" in lib.vb_last_error()
    import vivid_b200
    net = vivid_b200.NVPrecond(**cases.CASES["v_cond"]["cfg"])
    x = torch.zeros(1, 3, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, x, torch.ones(1), torch.zeros(1, 20))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vivid_b200.edm_sampler(net, x, x)


def test_argument_validation_returns_error_codes(lib):
    from vivid_b200 import _lib as L
    d = L.ConvDesc(B=1, H=12, W=12, cin_pad=64, cout_pad=64, taps=9, block_n=64)
    assert lib.vb_conv(ctypes.byref(d), None) == -1
    assert b"required" in lib.vb_last_error()
    d = L.ConvDesc(x=1 << 20, w=1 << 20, B=1, H=12, W=12, cin_pad=64, cout_pad=64, taps=9, block_n=64)
    assert lib.vb_conv(ctypes.byref(d), None) == -1 and b"power of two" in lib.vb_last_error()
    d = L.ConvDesc(x=1 << 20, w=1 << 20, B=1, H=16, W=16, cin_pad=60, cout_pad=64, taps=9, block_n=64)
    assert lib.vb_conv(ctypes.byref(d), None) == -1 and b"multiples of 64" in lib.vb_last_error()
    # K-split (tune bit 8) needs its workspace and an even number of 64-channel K blocks per tap
    d = L.ConvDesc(x=1 << 20, w=1 << 20, B=2, H=8, W=8, cin_pad=128, cout_pad=64, taps=9, block_n=64, tune=256)
    assert lib.vb_conv(ctypes.byref(d), None) == -1 and b"K-split" in lib.vb_last_error()
    d = L.ConvDesc(x=1 << 20, w=1 << 20, ks_ws=1 << 20, B=2, H=8, W=8, cin_pad=192, cout_pad=64, taps=9, block_n=64, tune=256)
    assert lib.vb_conv(ctypes.byref(d), None) == -1 and b"K-split" in lib.vb_last_error()
    assert lib.vb_conv_ksplit_ws_bytes(9, 8, 8, 512) == 5 * 128 * 512 * 4          # 2 images per tile -> 5 tiles
    assert lib.vb_conv_ksplit_ws_bytes(2, 16, 16, 128) == 4 * 128 * 128 * 4
    assert lib.vb_conv_ksplit_ws_bytes(0, 8, 8, 64) == 0
    a = L.AttnDesc(q=1, k=1, v=1, y=1, B=1, heads=1, sq=16, sk=16, head_dim=48)
    assert lib.vb_attn(ctypes.byref(a), None) == -1 and b"head_dim" in lib.vb_last_error()
    assert lib.vb_plan_run(None, 0, -1, None) == -1
    # the auxiliary entry points validate before they launch as well
    assert lib.vb_logvar(1, 0, 1, 1, 1, 1, 128, 1, None) == -1 and b"vb_logvar" in lib.vb_last_error()
    s = L.StatsDesc(feat=1, cum_mu=1, cum_sigma=1, n=4, f1=8, f2=4, ld1=8, dtype=L.VB_F32)
    assert lib.vb_stats_update(ctypes.byref(s), None) == -1 and b"second feature block" in lib.vb_last_error()
    s = L.StatsDesc(feat=1, cum_mu=1, cum_sigma=1, n=4, f1=8, ld1=8, dtype=L.VB_U8)
    assert lib.vb_stats_update(ctypes.byref(s), None) == -1 and b"dtype" in lib.vb_last_error()
    assert lib.vb_psnr_u8(1, 1, L.VB_F16, 2, 48, 48, 1, None, None) == -1 and b"uint8 or fp32" in lib.vb_last_error()
    assert lib.vb_resize(1, 1, 3, 256, 256, 16, 16, 1, None) == -1 and b"factor of 8" in lib.vb_last_error()
    c = L.F32ConvDesc(x=1, w=1, out=1, B=1, H=4, W=4, cin=4, cout=8, taps=4, ldo=8)
    assert lib.vb_f32_conv(ctypes.byref(c), None) == -1 and b"taps" in lib.vb_last_error()
    o = L.F32OpDesc(a=1, out=1, kind=L.VB_F32_QKV, B=1, H=4, W=4, ca=100, heads=2, parts=3, head_dim=16, seg_div=1)
    assert lib.vb_f32_op(ctypes.byref(o), None) == -1 and b"channel count" in lib.vb_last_error()
    assert lib.vb_f32_attn(1, 1, 1, 1, 1, 1, 16, 16, 48, 0, None) == -1 and b"head_dim" in lib.vb_last_error()


@pytest.mark.parametrize("case", list(cases.CASES))
def test_parameter_tree_matches_reference(golden, case):
    """state_dict keys, shapes AND order equal the reference's (shapes recorded by make_golden.py)."""
    import vivid_b200
    mode = cases.CASES[case]["mode"]
    ref_shapes = [(k, tuple(s)) for k, s in golden[mode]["nets"][case]["shapes"]]
    with torch.device("meta"):
        net = vivid_b200.NVPrecond(**cases.CASES[case]["cfg"])
    mine = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert mine == ref_shapes
    assert net.dual == (mode == "dual")


def test_preset_parameter_counts():
    """base 250.65 M (enc 119.39 + unet 131.26), uncond 131.26 M, sr 38.20 M, tiny 59.33 M."""
    import vivid_b200

    def count(**kw):
        with torch.device("meta"):
            return sum(p.numel() for p in vivid_b200.NVPrecond(img_channels=3, label_dim=20, **kw).parameters())
    assert count(img_resolution=64, model_channels=128, extra_attn=1) == 250643011
    assert count(img_resolution=64, model_channels=128, extra_attn=1, uncond=True) == 131254309
    assert count(img_resolution=256, model_channels=64, super_res=True) == 38193205
    assert count(img_resolution=32, model_channels=64) == 59326852
    with torch.device("meta"):
        base = vivid_b200.NVPrecond(64, 3, 20, model_channels=128, extra_attn=1)
    feats = base.encoder.feature_specs()
    assert len(feats) == 17                      # 32²x256 x2, 16²x384 x7, 8²x512 x8
    assert sorted((f.res, f.cout) for f in feats).count((8, 512)) == 8
    assert [f.name for f in base.unet.feature_specs()] == [f.name for f in feats]


def test_interface_attributes_and_unsupported_flags():
    import vivid_b200
    with torch.device("meta"):
        net = vivid_b200.NVPrecond(64, 3, 20, model_channels=128, extra_attn=1, noisy_sr=0.25)
    for attr, val in dict(img_resolution=64, img_channels=3, no_time_enc=None, depth_input=False, super_res=False,
                          uncond=None, use_fp16=True, sigma_data=0.5, noisy_sr=0.25, label_dim=20).items():
        assert getattr(net, attr) == val
    assert net.init_kwargs["model_channels"] == 128
    with pytest.raises(NotImplementedError):
        vivid_b200.NVPrecond(64, 3, 20, depth_input=True)
    with pytest.raises(NotImplementedError):
        vivid_b200.NVPrecond(64, 3, 20, warp_depth_coor=True)
    with pytest.raises(NotImplementedError):
        vivid_b200.NVPrecond(64, 3, 20, model_channels=64, epipolar_attention_bias=True)
    with pytest.raises(TypeError):
        vivid_b200.NVPrecond(64, 3)
    with pytest.raises(ValueError):
        vivid_b200.NVPrecond(48, 3, 20)
    x = torch.zeros(1, 3, 64, 64)
    # the optional outputs / modes are served by the CUDA path like everything else
    for kw in (dict(force_fp32=True), dict(return_logvar=True), dict(return_features=True), dict(inject_features=[x])):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            net(x, x, torch.ones(1), None, **kw)
    unc = vivid_b200.NVPrecond(64, 3, 20, model_channels=64, uncond=True)
    with pytest.raises(NotImplementedError):                       # no source-view encoder -> no features to hand out
        unc(x, x, torch.ones(1), None, return_features=True)


def test_seed_sharding_matches_reference_formula():
    from vivid_b200.generate import split_seeds
    for n, mb, w in [(8, 32, 1), (10000, 32, 8), (100, 32, 2), (33, 32, 4), (1, 32, 8)]:
        got = [split_seeds(n, mb, r, w) for r in range(w)]
        # reference: generate_images.py:199-200
        num_batches = max((n - 1) // (mb * w) + 1, 1) * w
        ref = np.array_split(np.arange(n), num_batches)
        flat = sorted(int(i) for part in got for b in part for i in b)
        assert flat == list(range(n))
        assert all(len(b) <= mb for part in got for b in part)
        assert len({len(p) for p in got}) == 1            # every rank sees the same number of batches
        for r in range(w):
            assert all(np.array_equal(a, b) for a, b in zip(got[r], ref[r::w]))
    assert {len(b) for b in split_seeds(10000, 32, 0, 8)} == {31, 32}


def test_synthetic_inputs_are_seed_keyed_and_contract_shaped():
    from vivid_b200 import synthetic
    a = synthetic.synth_batch([5, 9], 64)
    b = synthetic.synth_batch([9], 64)
    assert a["src_image"].shape == (2, 3, 64, 64) and a["geometry"].shape == (2, 20)
    assert torch.equal(a["src_image"][1], b["src_image"][0]) and torch.equal(a["geometry"][1], b["geometry"][0])
    assert a["src_image"].min() >= 0 and a["src_image"].max() <= 255
    assert (a["geometry"][:, [14, 15, 18, 19]] == 0).all()           # std == 0 entries (cx, cy) are zeroed
    d = synthetic.synth_batch([1, 2], 32, dual=True)
    assert d["src_image"].shape[0] == 4 and torch.equal(d["tgt_image"][0], d["tgt_image"][1])


def test_compose_geometry_and_schedule_match_reference(golden):
    from vivid_b200 import synthetic
    from vivid_b200.sampler import sigma_steps
    g = golden["dual"]["ops"]
    assert torch.allclose(synthetic.compose_geometry(g["ext"], g["k_src"], g["k_tgt"], 64), g["compose_geometry_64"], atol=1e-6)
    assert torch.allclose(synthetic.compose_geometry(g["ext"], g["k_src"] * 4, g["k_tgt"] * 4, 256),
                          g["compose_geometry_256"], atol=1e-6)
    # the snapshot tree's call form: 3x3 K matrices (experiments/code/training/utils.py:65-75), against ITS golden
    v = golden["vanilla"]["ops"]
    K = synthetic.decompose_K
    assert torch.allclose(synthetic.compose_geometry(v["ext"], K(v["k_src"]), K(v["k_tgt"]), 64), v["compose_geometry_64"], atol=1e-6)
    assert torch.allclose(synthetic.compose_geometry(v["ext"], K(v["k_src"] * 4), K(v["k_tgt"] * 4), 256),
                          v["compose_geometry_256"], atol=1e-6)
    assert torch.equal(synthetic.compose_K(K(v["k_src"])), v["k_src"])
    ext, ks, kt = synthetic.decompose_geometry(v["compose_geometry_64"], 64)
    assert torch.allclose(ext, v["ext"], atol=1e-4) and torch.allclose(synthetic.compose_K(ks)[..., :2], v["k_src"][..., :2], atol=1e-3)
    assert torch.allclose(synthetic.resize_geometry(v["compose_geometry_64"], 64, 256)[..., :14],
                          v["compose_geometry_256"][..., :14], atol=1e-4)
    t = sigma_steps(32, 0.002, 80, 7, "cpu")
    assert torch.equal(t[:-1], golden["vanilla"]["nets"]["t_steps_32"]) and t[-1] == 0
    from vivid_b200 import StackedRandomGenerator
    assert torch.equal(StackedRandomGenerator("cpu", [3, 4, (1 << 32) + 3]).randn([3, 2, 4]), g["stacked_randn"])


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vivid_b200.generate import split_seeds, reduce_psnr
from vivid_b200.metrics import finalize_stats
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
batches = split_seeds(10, 3, rank, 2)
# the per-batch accumulation is CUDA-only (vb_psnr_u8 / vb_stats_update); here: the cross-rank reduction of the
# accumulators each rank would hold after its shard (calculate_metrics.py:174-183,221-236)
import math
val = 10 * math.log10(255 ** 2 / 4.0)
n_local = sum(len(b) for b in batches)
m = reduce_psnr(torch.full([1], val * n_local, dtype=torch.float64), torch.tensor(n_local))
assert m["num_images"] == 10, m
assert abs(m["psnr"] - val) < 1e-9, m
g = torch.Generator().manual_seed(5)
x = torch.randn(10, 6, generator=g, dtype=torch.float64)
mine = x[[i for b in batches for i in b]]
st = finalize_stats(mine.sum(0), mine.T @ mine, 10)
import numpy as np
assert np.allclose(st["mu"], x.mean(0).numpy()) and np.allclose(st["sigma"], np.cov(x.numpy(), rowvar=False))
# rank-0 image gather (SURVEY.md §8(e)): ragged and empty per-rank batches, one gather per batch
from vivid_b200.generate import gather_batch
for by_rank in ([[1, 2, 3], [4, 5]], [[7], []]):
    mine_n = len(by_rank[rank])
    imgs = torch.full((mine_n, 3, 4, 4), 10 * rank + 1, dtype=torch.uint8) if mine_n else None
    got, gs = gather_batch(imgs, by_rank, (3, 4, 4), torch.device("cpu"), rank, 2)
    if rank == 0:
        n0, n1 = len(by_rank[0]), len(by_rank[1])
        assert got.shape == (n0 + n1, 3, 4, 4) and got.dtype == torch.uint8 and gs == by_rank[0] + by_rank[1]
        assert (got[:n0] == 1).all() and (got[n0:] == 11).all()
    else:
        assert got is None and gs is None
print("rank", rank, "ok")
"""


def test_world_size_2_sharding_and_metric_reduction(tmp_path):
    """N>1 host path on CPU: gloo, two ranks, disjoint seed shards, all_reduced PSNR statistics."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_metrics_host_logic_and_cpu_refusal():
    """calculate_metrics_from_stats_nvs (host numpy/scipy) against the reference's results; CUDA-only accumulation."""
    import vivid_b200
    from vivid_b200 import metrics as M
    g = torch.load(os.path.join(ROOT, "tests", "golden", "metrics.pt"), weights_only=False)
    res = vivid_b200.calculate_metrics_from_stats_nvs(g["stats"], g["ref"], metrics=g["metrics"], verbose=False)
    assert set(res) == set(g["results"])
    for k, v in g["results"].items():
        assert abs(res[k] - v) < 1e-8 * max(1.0, abs(v)), k
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.stats_update(torch.zeros(4, dtype=torch.float64), torch.zeros(4, 4, dtype=torch.float64), torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.psnr_u8(torch.zeros(1, 3, 2, 2, dtype=torch.uint8), torch.zeros(1, 3, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vivid_b200.get_metrics(iter([]), device=torch.device("cpu"))
    with pytest.raises(NotImplementedError):
        vivid_b200.calculate_stats_for_iterable_nvs([], metrics=["fid"])


def test_whole_call_entry_point_argument_validation(lib):
    """vb_plan_bind_io / vb_denoise / vb_workspace_bytes reject bad arguments with an error code (no CUDA call is made)."""
    import ctypes as C
    from vivid_b200 import _lib as L
    plan = C.c_void_p()
    assert lib.vb_plan_create(C.byref(plan)) == 0
    try:
        assert lib.vb_workspace_bytes(plan) == 0
        assert lib.vb_denoise(plan, None, None, None, 1, None, 0, None, None, None, None) != 0
        assert b"bound" in lib.vb_last_error()
        io = L.IoDesc()
        assert lib.vb_plan_bind_io(plan, C.byref(io)) != 0 and b"required" in lib.vb_last_error()
        io = L.IoDesc(in_x=8, in_sigma=8, in_geom=8, out_d=8, n_x=2, n_out=2, img_elems=12, geom_dim=4, in_cond=8, workspace_bytes=99)
        assert lib.vb_plan_bind_io(plan, C.byref(io)) != 0 and b"come together" in lib.vb_last_error()
        io.in_cond = None
        assert lib.vb_plan_bind_io(plan, C.byref(io)) == 0
        assert lib.vb_workspace_bytes(plan) == 99
        assert lib.vb_denoise(plan, None, None, None, 1, None, 0, None, None, None, None) != 0
        assert b"required" in lib.vb_last_error()
        assert lib.vb_plan_bind_io(None, C.byref(io)) != 0
        # whole-sampler entry point: same rules
        assert lib.vb_sample_workspace_bytes(plan) == 3 * 2 * 12 * 4 and lib.vb_sample_workspace_bytes(None) == 0
        unbound = C.c_void_p()
        assert lib.vb_plan_create(C.byref(unbound)) == 0
        d = L.SampleDesc(net=unbound, noise=8, workspace=8, x_out=8, num_steps=2, guidance=1.0)
        assert lib.vb_sample(C.byref(d), None) != 0 and b"bound I/O" in lib.vb_last_error()
        assert lib.vb_plan_set_inputs(unbound, None, None, 0, None, None) != 0 and b"bound" in lib.vb_last_error()
        lib.vb_plan_destroy(unbound)
        d.net = plan
        assert lib.vb_sample(C.byref(d), None) != 0 and b"t_steps" in lib.vb_last_error()
        d.t_steps = (C.c_float * 3)(80.0, 1.0, 0.0)
        d.guidance = 2.0
        assert lib.vb_sample(C.byref(d), None) != 0 and b"gnet" in lib.vb_last_error()
        d.guidance, d.num_steps = 1.0, 0
        assert lib.vb_sample(C.byref(d), None) != 0 and b"num_steps" in lib.vb_last_error()
        assert lib.vb_plan_set_inputs(plan, None, 8, 3, None, None) != 0 and b"geometry_rows" in lib.vb_last_error()
    finally:
        lib.vb_plan_destroy(plan)


def test_c_host_example_compiles_against_the_header():
    """examples/sample_from_c.c (the INTEGRATION.md snippet: checkpoint -> vb_net_plan_create -> vb_plan_set_inputs -> vb_sample)
    is valid C against include/vivid_b200.h — the header is plain C: no torch, no C++ in the signatures."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    cuda_inc = "/usr/local/cuda/include"
    if gcc is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("gcc or the CUDA headers are not installed")
    snippet = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    src = open(os.path.join(ROOT, "examples", "sample_from_c.c")).read()
    body = src[src.index('#include "vivid_b200.h"'):]
    assert body.strip() in snippet, "examples/sample_from_c.c and the INTEGRATION.md snippet have diverged"
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
                        os.path.join(ROOT, "examples", "sample_from_c.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
