"""Pin the oracle: against the committed reference outputs (tests/golden, produced by
tests/golden/make_golden.py running the REAL reference) and, when /root/reference is
present, against the live reference modules."""
import os

import numpy as np
import pytest
import torch

import refshim
from diffsheg_b200 import synth
from oracle import diffusion as odiff
from oracle.denoiser import unidiffuser_forward

have_ref = pytest.mark.skipif(not refshim.have_reference(), reason="/root/reference not present")


def _fp(*tensors):
    return np.array([float(t.double().sum()) for t in tensors] + [float(t.double().abs().sum()) for t in tensors])


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_tables_match_reference(golden_dir):
    tabs = np.load(os.path.join(golden_dir, "tables.npz"))
    d25 = odiff.OracleDiffusion(1000, "ddim25")
    d1000 = odiff.OracleDiffusion(1000, None)
    for tag, d in (("d25", d25), ("d1000", d1000)):
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            np.testing.assert_array_equal(getattr(d, k), tabs[f"{tag}_{k}"], err_msg=f"{tag}_{k}")
    assert d25.timestep_map == list(tabs["d25_timestep_map"]) == list(range(0, 1000, 40))
    assert odiff.schedule_jump_ddim(25, 3, 5) == list(tabs["jump_25_3_5"])
    assert odiff.schedule_jump_ddim(25, 3, 2) == list(tabs["jump_25_3_2"])
    assert odiff.schedule_jump_ddim(25) == list(tabs["jump_25_1_1"])
    assert odiff.schedule_jump_paper() == list(tabs["jump_paper"])
    # 63 denoise + 48 undo with (3,5); 27 + 12 with (3,2); 15 with no_resample
    for args, nd, nu in (((25, 3, 5), 63, 48), ((25, 3, 2), 27, 12), ((25,), 15, 0)):
        ts = odiff.schedule_jump_ddim(*args)
        pairs = list(zip(ts[:-1], ts[1:]))
        assert sum(c < l for l, c in pairs) == nd and sum(c > l for l, c in pairs) == nu


@pytest.mark.parametrize("name,B,T,t_resp", [("show", 2, 88, 12), ("show", 3, 84, 0),
                                              ("beat", 2, 34, 24), ("beat", 1, 30, 3)])
def test_denoiser_matches_golden(golden_dir, name, B, T, t_resp):
    g = np.load(os.path.join(golden_dir, f"denoise_{name}_B{B}_T{T}_t{t_resp}.npz"))
    cfg = synth.make_cfg(name)
    sd = synth.make_state_dict(cfg, seed=1)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    np.testing.assert_allclose(_fp(inp["x_T"], inp["mel"], inp["hubert"]), g["fp"], rtol=1e-9)
    ts = torch.full((B,), int(g["t_orig"]), dtype=torch.long)
    a, b = torch.tensor(float(g["a"])), torch.tensor(float(g["b"]))
    with torch.no_grad():
        eps = unidiffuser_forward(sd, cfg, inp["x_T"], ts, (a, b), inp["mel"], inp["person_id"], inp["hubert"])
    # same op stream, same fp32 kernels: differences are reassociation-level only
    assert relmax(eps.numpy(), g["eps"]) < 2e-5


_VARIANTS = [(cp, cr) for cp in ("mlp_includeX", "linear_includeX", "mlp_excludeX", "linear_excludeX") for cr in (True, False)
             if not (cp == "mlp_includeX" and cr)]


@pytest.mark.parametrize("name", ["show", "beat"])
def test_denoiser_variants_match_golden(golden_dir, name):
    """cond_projection / cond_residual variants (tr:262-263,281-289,300-338) against outputs of the REAL reference
    (tests/golden/make_golden_variants.py)."""
    g = np.load(os.path.join(golden_dir, "denoise_variants.npz"))
    B, T, _, t_orig, a, b = g[name + "_consts"]
    B, T = int(B), int(T)
    for cp, cr in _VARIANTS:
        cfg = synth.make_cfg(name, cond_projection=cp, cond_residual=cr)
        sd = synth.make_state_dict(cfg, seed=1)
        inp = synth.make_inputs(cfg, B, T, seed=2)
        np.testing.assert_allclose(_fp(inp["x_T"], inp["mel"], inp["hubert"]), g[name + "_fp"], rtol=1e-9)
        ts = torch.full((B,), int(t_orig), dtype=torch.long)
        with torch.no_grad():
            eps = unidiffuser_forward(sd, cfg, inp["x_T"], ts, (torch.tensor(float(a)), torch.tensor(float(b))), inp["mel"],
                                      inp["person_id"], inp["hubert"])
        assert relmax(eps.numpy(), g[f"{name}_{cp}_{'res' if cr else 'nores'}"]) < 2e-5, (name, cp, cr)


@have_ref
@pytest.mark.parametrize("cond_projection,cond_residual", _VARIANTS)
def test_denoiser_variants_match_live_reference(cond_projection, cond_residual):
    """Same op stream as the reference module itself (strict state_dict load: the synthetic key layout of every variant IS the
    reference's), two layers per net, CFG pair."""
    cfg = synth.make_cfg("show", cond_projection=cond_projection, cond_residual=cond_residual, num_layers=2)
    sd = synth.make_state_dict(cfg, seed=3)
    opt = refshim.make_opt(cfg, cond_projection=cond_projection, cond_residual=cond_residual)
    model, _ = refshim.build_reference(cfg, sd, opt)
    B, T = 2, 11
    inp = synth.make_inputs(cfg, B, T, seed=4)
    ts = torch.full((B,), 360, dtype=torch.long)
    sa = (torch.tensor(1.7), torch.tensor(1.4))
    with torch.no_grad():
        want = model(inp["x_T"], ts, sqrt_alphas=sa, audio_emb=inp["mel"], length=torch.LongTensor([T] * B),
                     person_id=inp["person_id"], add_cond={"pretrain_aud_feat": inp["hubert"]}, pe_type="pe_sinu", y={})
        got = unidiffuser_forward(sd, cfg, inp["x_T"], ts, sa, inp["mel"], inp["person_id"], inp["hubert"])
    assert relmax(got.numpy(), want.numpy()) < 1e-6


@pytest.mark.parametrize("fn,name,B,T,ov,kw", [
    ("loop_show_B1_T88_ov0_ddim25.npz", "show", 1, 88, 0, {}),
    ("loop_beat_B2_T34_ov0_ddim25.npz", "beat", 2, 34, 0, {}),
    ("loop_beat_B1_T34_ov4_ddim25_jn2.npz", "beat", 1, 34, 4, dict(jump_n_sample=2)),
    ("loop_show_B2_T88_ov10_ddim25.npz", "show", 2, 88, 10, {}),
])
def test_ddim_loop_matches_golden(golden_dir, fn, name, B, T, ov, kw):
    g = np.load(os.path.join(golden_dir, fn))
    cfg = synth.make_cfg(name)
    sd = synth.make_state_dict(cfg, seed=1)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    y = _inpaint(cfg, B, T, ov)
    d = odiff.OracleDiffusion(1000, "ddim25", overlap_len=ov, **kw)
    den = odiff.make_denoise(sd, cfg, inp["mel"], inp["person_id"], inp["hubert"])
    torch.manual_seed(int(g["seed"]))
    with torch.no_grad():
        out = d.ddim_sample_loop(den, (B, T, cfg["net_dim_pose"]), y=y)
    assert relmax(out.numpy(), g["sample"]) < 1e-3


def test_ddpm_loop_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loop_beat_B2_T34_ov0_ddpm40.npz"))
    cfg = synth.make_cfg("beat")
    sd = synth.make_state_dict(cfg, seed=1)
    inp = synth.make_inputs(cfg, 2, 34, seed=2)
    d = odiff.OracleDiffusion(40, None)
    den = odiff.make_denoise(sd, cfg, inp["mel"], inp["person_id"], inp["hubert"])
    torch.manual_seed(int(g["seed"]))
    with torch.no_grad():
        out = d.p_sample_loop(den, (2, 34, cfg["net_dim_pose"]), y={})
    assert relmax(out.numpy(), g["sample"]) < 1e-3


def _inpaint(cfg, B, T, ov):
    if ov == 0:
        return {}
    g = torch.Generator().manual_seed(5)
    gt = torch.zeros(B, T, cfg["net_dim_pose"])
    gt[:, :ov] = torch.randn(B, ov, cfg["net_dim_pose"], generator=g)
    mask = torch.zeros(B, T, cfg["net_dim_pose"], dtype=torch.bool)
    mask[:, :ov] = True
    return {"gt": gt, "outpainting_mask": mask}


@have_ref
def test_synthetic_state_dict_is_the_reference_layout():
    for name in ("show", "beat"):
        cfg = synth.make_cfg(name)
        model, _ = refshim.build_reference(cfg, synth.make_state_dict(cfg))  # strict=True inside
        ref_sd = model.state_dict()
        assert {k: tuple(v.shape) for k, v in ref_sd.items()} == \
            {k: tuple(s) for k, s in synth.state_dict_shapes(cfg).items()}


@have_ref
def test_denoiser_matches_live_reference_nocfg_show():
    """cond_scale == 1 with classifier_free=True: no batch doubling (tr:537)."""
    cfg = synth.make_cfg("show", cond_scale=1.0)
    sd = synth.make_state_dict(cfg, seed=3)
    model, opt = refshim.build_reference(cfg, sd)
    inp = synth.make_inputs(cfg, 2, 88, seed=7)
    ts = torch.full((2,), 200, dtype=torch.long)
    a, b = torch.tensor(1.7), torch.tensor(1.3)
    shp = (2, 88, cfg["expression_dim"])
    with torch.no_grad():
        ref = model(inp["x_T"], ts, sqrt_alphas=[a.expand(shp), b.expand(shp)], audio_emb=inp["mel"],
                    length=torch.LongTensor([88, 88]), person_id=inp["person_id"],
                    add_cond={"pretrain_aud_feat": inp["hubert"]}, pe_type="pe_sinu", y={})
        got = unidiffuser_forward(sd, cfg, inp["x_T"], ts, (a, b), inp["mel"], inp["person_id"], inp["hubert"])
    assert relmax(got.numpy(), ref.numpy()) < 2e-5


@have_ref
def test_cross_attention_restatement_matches_the_live_reference_module():
    """The formula the GPU op test (test_op_cross_attention_with_static_shift_numerators) checks against IS the reference's
    LinearTemporalCrossAttention core (tr:153-164): same einsums, softmax over the head dim for Q and over the N conditioning
    frames for K."""
    import sys
    if refshim.REF not in sys.path:
        sys.path.insert(0, refshim.REF)
    from models.transformer import LinearTemporalCrossAttention
    torch.manual_seed(0)
    B, T, N, D, A, H = 2, 9, 5, 64, 32, 4
    m = LinearTemporalCrossAttention(seq_len=T, latent_dim=D, aud_latent_dim=A, num_head=H, dropout=0.0, time_embed_dim=48).eval()
    for p in m.parameters():
        torch.nn.init.normal_(p, std=0.2)
    x, xf, emb = torch.randn(B, T, D), torch.randn(B, N, A), torch.randn(B, 48)
    with torch.no_grad():
        ref = m(x, xf, emb)
        q = m.query(m.norm(x)).view(B, T, H, -1).softmax(-1)
        k = m.key(m.text_norm(xf)).view(B, N, H, -1).softmax(1)
        v = m.value(m.text_norm(xf)).view(B, N, H, -1)
        y = torch.einsum("bnhd,bhdl->bnhl", q, torch.einsum("bnhd,bnhl->bhdl", k, v)).reshape(B, T, D)
        e = m.proj_out.emb_layers(emb).unsqueeze(1)
        scale, shift = e.chunk(2, dim=2)
        got = x + m.proj_out.out_layers(m.proj_out.norm(y) * (1 + scale) + shift)
    assert relmax(got.numpy(), ref.numpy()) < 1e-6
