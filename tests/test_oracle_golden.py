"""Pin the oracle (oracle/vivid_oracle.py) against outputs of the UNMODIFIED reference.

The reference ships no tests or golden vectors; tests/golden/*.pt were produced by
tests/golden/make_golden.py running /root/reference in the build container.  fp32 on CPU; the oracle
restates the ops with different but equivalent torch calls, so agreement is to rounding (<= 2e-5 rel).
"""
import pytest
import torch

import cases
from oracle import vivid_oracle as O

TOL = 2e-5


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_primitive_ops(golden):
    for mode in ("vanilla", "dual"):
        g = golden[mode]["ops"]
        x, y = g["x"], g["y"]
        assert rel(O.normalize(x, dim=1), g["normalize_dim1"]) < 1e-6
        assert rel(O.normalize(x), g["normalize_all"]) < 1e-6
        assert rel(O.resample(x, "down"), g["resample_down"]) < 1e-6
        assert torch.equal(O.resample(x, "up"), g["resample_up"])
        assert rel(O.mp_silu(x), g["mp_silu"]) < 1e-6
        assert rel(O.mp_sum(x, y, 0.3), g["mp_sum_03"]) < 1e-6
        assert rel(O.mp_cat(x, y[:, :4], 0.5), g["mp_cat"]) < 1e-6
        assert rel(O.mp_conv(x, g["w3"], gain=0.7), g["mpconv3"]) < 1e-6
        assert rel(O.mp_conv(x[:, :, 0, 0], g["w1"]), g["mplinear"]) < 1e-6
        assert rel(O.mp_fourier(g["fourier_in"], g["freqs"], g["phases"]), g["fourier"]) < 1e-6
        assert torch.equal(O.encode_latents(g["u8"]), g["encode_latents"])
        assert torch.equal(O.decode(g["lat"]), g["decode"])
        assert rel(O.compose_geometry(g["ext"], g["k_src"], g["k_tgt"], 64), g["compose_geometry_64"]) < 1e-6
        assert rel(O.compose_geometry(g["ext"], g["k_src"] * 4, g["k_tgt"] * 4, 256), g["compose_geometry_256"]) < 1e-6
        r = O.StackedRandomGenerator("cpu", [3, 4, (1 << 32) + 3]).randn([3, 2, 4])
        assert torch.equal(r, g["stacked_randn"])
        assert torch.equal(r[0], r[2])          # seeds are taken modulo 2**32


def test_sigma_schedule(golden):
    t = O.sigma_schedule(32, 0.002, 80, 7, "cpu")
    assert t.shape == (33,) and t[-1] == 0
    assert rel(t[:-1], golden["vanilla"]["nets"]["t_steps_32"]) < 1e-6


def _oracle(rec, case):
    sd = cases.synth_state_dict(rec["shapes"])
    cfg = dict(rec["cfg"], dual=cases.CASES[case]["mode"] == "dual")
    return O.OracleNet(sd, cfg)


@pytest.mark.parametrize("case", ["v_cond", "v_uncond", "v_sr", "d_cond", "v_tiny"])
def test_denoiser_against_reference(golden, case):
    mode = cases.CASES[case]["mode"]
    rec = golden[mode]["nets"][case]
    net = _oracle(rec, case)
    inp = cases.synth_inputs(case, rec["B"])
    n_in = inp["src"].shape[0]
    with torch.no_grad():
        for sg, ref in rec["D"].items():
            x = inp["tgt"] + sg * inp["noise"]
            kw = {}
            if rec["cfg"].get("super_res"):
                torch.manual_seed(123)
                kw["conditioning_image"] = inp["tgt"]
            d = net(inp["src"], x, torch.full((n_in,), sg), inp["geometry"], **kw)
            assert d.shape == ref.shape
            assert rel(d, ref) < TOL, (case, sg)
            # the test is not vacuous: the network part of D is non-zero
            c_skip = 0.25 / (sg ** 2 + 0.25)
            xs = x[::2] if mode == "dual" else x
            assert (ref - c_skip * xs).abs().max() > 1e-3
        if "D_nogeom" in rec:
            x = inp["tgt"] + 5.0 * inp["noise"]
            d = net(inp["src"], x, torch.full((n_in,), 5.0))
            assert rel(d, rec["D_nogeom"]) < TOL


def _extra(key):
    import os
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "extra.pt"), weights_only=False)[key]


@pytest.mark.parametrize("case", ["v_cond", "d_cond"])
def test_logvar_head_against_reference(case):
    """Uncertainty head (return_logvar=True); fixture from tests/golden/make_golden_extra.py."""
    rec = _extra(f"logvar_{case}")
    net = O.OracleNet(cases.synth_state_dict(rec["shapes"]), dict(rec["cfg"], dual=case == "d_cond"))
    inp = cases.synth_inputs(case, rec["B"])
    x = inp["tgt"] + rec["sigma"].reshape(-1, 1, 1, 1) * inp["noise"]
    with torch.no_grad():
        d, lv = net(inp["src"], x, rec["sigma"], inp["geometry"], return_logvar=True)
    assert lv.shape == rec["logvar"].shape == (rec["B"], 1, 1, 1)
    assert (lv - rec["logvar"]).abs().max() < 1e-5
    assert rel(d, rec["D"]) < TOL


@pytest.mark.parametrize("case", ["v_cond", "d_cond"])
def test_cached_source_features_against_reference(case):
    """no_time_enc nets: return_features / inject_features and the sampler that caches the encoder output."""
    rec = _extra(f"features_{case}")
    net = O.OracleNet(cases.synth_state_dict(rec["shapes"]), dict(rec["cfg"], dual=case == "d_cond"))
    assert net.no_time_enc
    inp = cases.synth_inputs(case, rec["B"])
    n_in = inp["src"].shape[0]
    with torch.no_grad():
        feats = net(inp["src"], torch.zeros_like(inp["src"]), torch.ones(n_in), inp["geometry"], None, return_features=True)
        assert len(feats) == len(rec["features"])
        for f, ref in zip(feats, rec["features"]):
            assert f.shape == ref.shape and rel(f, ref) < TOL
        x = inp["tgt"] + rec["sigma"] * inp["noise"]
        d = net(torch.zeros_like(inp["src"]), x, torch.full((n_in,), rec["sigma"]), inp["geometry"], inject_features=feats)
        assert rel(d, rec["D"]) < TOL
        lat = O.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=rec["num_steps"])
    assert rel(lat, rec["latents"]) < 5 * TOL


@pytest.mark.parametrize("case", ["v_cond", "d_cond"])
def test_stochastic_sampler_against_reference(case):
    """S_churn > 0 branch (generate_images.py:77-84) with the default torch.randn_like and a seeded global generator."""
    rec = _extra(f"churn_{case}")
    net = O.OracleNet(cases.synth_state_dict(rec["shapes"]), dict(rec["cfg"], dual=case == "d_cond"))
    inp = cases.synth_inputs(case, rec["B"])
    torch.manual_seed(rec["seed"])
    with torch.no_grad():
        lat = O.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], **rec["kwargs"])
    assert lat.shape == rec["latents"].shape and rel(lat, rec["latents"]) < 5 * TOL


def test_guided_sampler_against_reference(golden):
    nets = golden["vanilla"]["nets"]
    net = _oracle(nets["v_cond"], "v_cond")
    gnet = _oracle(nets["v_uncond"], "v_uncond")
    inp = cases.synth_inputs("v_cond", 2)
    ref = nets["sampler_guided"]
    trace = []
    with torch.no_grad():
        lat = O.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], gnet=gnet, num_steps=ref["num_steps"],
                            guidance=ref["guidance"], trace=trace)
        lat1 = O.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=4)
    assert rel(lat, ref["latents"]) < 1e-4
    assert len(trace) == 2 * ref["num_steps"] - 1 == ref["net_calls"].shape[0]
    assert (O.decode(lat).int() - ref["images"].int()).abs().max() <= 1
    assert rel(lat1, nets["sampler_unguided"]["latents"]) < 1e-4


def test_sr_sampler_against_reference(golden):
    nets = golden["vanilla"]["nets"]
    sr = _oracle(nets["v_sr"], "v_sr")
    inp = cases.synth_inputs("v_sr", 2)
    torch.manual_seed(nets["sampler_sr"]["seed"])
    with torch.no_grad():
        lat = O.edm_sampler(sr, inp["src"], inp["noise"], labels=inp["geometry"], gnet=sr, num_steps=3,
                            conditioning_image=inp["tgt"])
    assert rel(lat, nets["sampler_sr"]["latents"]) < 1e-4


def test_dual_sampler_against_reference(golden):
    nets = golden["dual"]["nets"]
    net = _oracle(nets["d_cond"], "d_cond")
    inp = cases.synth_inputs("d_cond", 2)
    with torch.no_grad():
        lat = O.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=3)
    assert lat.shape[0] == 2
    assert rel(lat, nets["sampler_dual"]["latents"]) < 1e-4


def test_tiny_config_sampler_against_reference(golden):
    """BASELINE.json configs[0]: tiny EDM2 NVS UNet (64 ch, 32x32), Heun 8 steps, batch 2, CPU."""
    nets = golden["vanilla"]["nets"]
    net = _oracle(nets["v_tiny"], "v_tiny")
    inp = cases.synth_inputs("v_tiny", 2)
    with torch.no_grad():
        lat = O.edm_sampler(net, inp["src"], inp["noise"], labels=inp["geometry"], num_steps=8)
    assert rel(lat, nets["sampler_tiny"]["latents"]) < 1e-4


# ----------------------------------------------------------------------------- metric statistics (calculate_metrics.py gen)
def _metrics_golden():
    import os
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "metrics.pt"), weights_only=False)


def test_metric_statistics_against_reference():
    """Oracle restatement of update_mu_sigma / reduce / psnr / Frechet distance vs the reference's own outputs
    (tests/golden/make_golden_metrics.py: real calculate_stats_for_iterable_nvs with a fake detector)."""
    import numpy as np
    g = _metrics_golden()
    det = cases.FakeDetector()
    st, rf, jst, jrf = O.StatsOracle(48), O.StatsOracle(48), O.StatsOracle(96), O.StatsOracle(96)
    ps = []
    for src, tgt, img in cases.synth_metric_batches():
        f_img, f_tgt, f_src = det(img).numpy(), det(tgt).numpy(), det(src).numpy()
        st.update(f_img)
        rf.update(f_tgt)
        jst.update(f_img, f_src)
        jrf.update(f_tgt, f_src)
        ps.append(O.psnr_u8(img.numpy(), tgt.numpy()))
    assert st.n == g["stats"]["num_images"] == 14
    for name, acc, side in (("fid", st, "stats"), ("fid", rf, "ref"), ("joint_fid", jst, "stats"), ("joint_fid", jrf, "ref")):
        got = acc.finalize()
        assert np.allclose(got["mu"], g[side][name]["mu"], rtol=1e-12, atol=1e-13)
        assert np.allclose(got["sigma"], g[side][name]["sigma"], rtol=1e-9, atol=1e-12)
    # the reference evaluates PSNR in fp32 (:148): agreement to fp32 rounding
    assert abs(np.concatenate(ps).mean() - float(np.asarray(g["stats"]["psnr"]["val"]).reshape(-1)[0])) < 1e-4
    fid = O.frechet_distance(st.finalize()["mu"], st.finalize()["sigma"], rf.finalize()["mu"], rf.finalize()["sigma"])
    assert abs(fid - g["results"]["fid"]) < 1e-8 * max(1.0, abs(g["results"]["fid"]))
    jfid = O.frechet_distance(jst.finalize()["mu"], jst.finalize()["sigma"], jrf.finalize()["mu"], jrf.finalize()["sigma"])
    assert abs(jfid - g["results"]["joint_fid"]) < 1e-8 * max(1.0, abs(g["results"]["joint_fid"]))
    # known answers: unbiased covariance, identical statistics -> distance 0
    x = np.random.default_rng(0).normal(size=(40, 7))
    acc = O.StatsOracle(7)
    acc.update(x[:25])
    acc.update(x[25:])
    assert np.allclose(acc.finalize()["sigma"], np.cov(x, rowvar=False), atol=1e-12)
    s = acc.finalize()
    assert abs(O.frechet_distance(s["mu"], s["sigma"], s["mu"], s["sigma"])) < 1e-6
