"""CPU tests of the host-side mirror of the reference interface (no GPU, no kernels)."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from diffsheg_b200 import (FusedGaussianDiffusion, FusedSpacedDiffusion, build_diffusions, cfg_from_opt, get_windows,
                           patch_trainer, space_timesteps, synth)
from oracle import diffusion as odiff

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg_from_opt_accepts_the_shipped_configs_and_rejects_the_rest():
    for name in ("show", "beat"):
        cfg0 = synth.make_cfg(name)
        cfg = cfg_from_opt(synth.make_opt(cfg0))
        for k in ("dim_pose", "expression_dim", "style_dim", "classifier_free", "cond_scale", "net_dim_pose"):
            assert cfg[k] == cfg0[k], k
    opt = synth.make_opt(synth.make_cfg("show"))
    for field, bad in (("model_base", "transformer_decoder"), ("cond_projection", "none"), ("unidiffuser", False),
                       ("encode_hubert", False), ("model_mean_type", "start_x")):
        o = argparse.Namespace(**vars(opt))
        setattr(o, field, bad)
        with pytest.raises(NotImplementedError):
            cfg_from_opt(o)


def test_cfg_from_opt_carries_the_cond_projection_variants_into_the_engine_config():
    """options/base_options.py:21,95: every projection the reference's UniDiffuser can run, cond_residual on or off."""
    from diffsheg_b200 import _lib
    from diffsheg_b200.engine import engine_config
    opt = synth.make_opt(synth.make_cfg("show"))
    c = engine_config(cfg_from_opt(opt), "bf16", 4, 88)     # an opt without the two fields = the shipped configuration
    assert (c.abi_version, c.cond_projection, c.no_cond_residual) == (_lib.ABI_VERSION, 0, 0)
    for code, cp in enumerate(("mlp_includeX", "linear_includeX", "mlp_excludeX", "linear_excludeX")):
        for cr in (True, False):
            o = argparse.Namespace(**vars(opt), cond_projection=cp, cond_residual=cr)
            cfg = cfg_from_opt(o)
            assert (cfg["cond_projection"], cfg["cond_residual"]) == (cp, cr)
            c = engine_config(cfg, "fp32", 1, 8)
            assert (c.cond_projection, c.no_cond_residual) == (code, int(not cr))
            # the synthetic state_dict follows the projection: P = [512 +] 256 + 128 [+ expression_dim] (tr:260-276)
            shapes = synth.state_dict_shapes(dict(synth.make_cfg("show"), cond_projection=cp))
            P = (512 if cp.endswith("includeX") else 0) + 256 + 128
            assert shapes["encoder_exp.null_cond_emb"] == (1, P)
            assert shapes["encoder_ges.null_cond_emb"] == (1, P + 103)
            key = "encoder_exp.temporal_decoder_blocks.0.feat_proj" + (".weight" if cp.startswith("linear") else ".1.weight")
            assert shapes[key] == ((512 if cp.startswith("linear") else 1024), P)


def test_unsupported_sampler_options_fail_loudly():
    betas = np.linspace(1e-4, 0.02, 1000)
    with pytest.raises(NotImplementedError):
        FusedGaussianDiffusion(betas=betas, model_mean_type="START_X")
    with pytest.raises(NotImplementedError):
        FusedGaussianDiffusion(betas=betas, model_var_type="LEARNED_RANGE")
    with pytest.raises(NotImplementedError):
        FusedGaussianDiffusion(betas=betas, opt=argparse.Namespace(fix_head_var=True))
    d = FusedGaussianDiffusion(betas=betas)
    with pytest.raises(NotImplementedError):  # generate_batch always passes clip_denoised=False (show:173)
        d.ddim_sample_loop(None, (1, 2, 3), clip_denoised=True, model_kwargs={})
    with pytest.raises(NotImplementedError):
        d.p_sample_loop(None, (1, 2, 3), clip_denoised=False, cond_fn=lambda *a: None, model_kwargs={})


def test_patch_trainer_swaps_both_sampler_objects():
    cfg = synth.make_cfg("show")

    class Trainer:  # the two attributes DDPMTrainer_show builds at show:63-80
        def __init__(self, opt):
            self.opt, self.diffusion, self.diffusion_ddim_val = opt, object(), object()

    tr = patch_trainer(Trainer(synth.make_opt(cfg, ddim=True)))
    assert isinstance(tr.diffusion, FusedGaussianDiffusion) and tr.diffusion.num_timesteps == 1000
    assert isinstance(tr.diffusion_ddim_val, FusedSpacedDiffusion) and tr.diffusion_ddim_val.num_timesteps == 25
    assert tr.diffusion_ddim_val.timestep_map == list(range(0, 1000, 40))   # 'ddim25' is hard-coded at show:73
    full, ddim = build_diffusions(synth.make_opt(cfg, ddim=False))
    assert ddim is None and full.num_timesteps == 1000


@pytest.mark.parametrize("n,spec", [(1000, "ddim25"), (1000, "ddim50"), (100, "10,5"), (300, "10,15,20"), (40, "ddim8")])
def test_space_timesteps_matches_oracle(n, spec):
    assert space_timesteps(n, spec) == odiff.space_timesteps(n, spec)


def test_get_windows_mirrors_the_reference_rule():
    import torch
    x = torch.arange(2 * 1800 * 3, dtype=torch.float32).view(2, 1800, 3)
    w = get_windows(x, 88, 78)
    assert len(w) == 23 and [t.shape[1] for t in w[-2:]] == [88, 84]
    assert torch.equal(w[1], x[:, 78:166]) and torch.equal(w[-1], x[:, 22 * 78:])
    assert len(get_windows(x[:, :88], 88, 78)) == 1 and len(get_windows(x[:, :60], 88, 78)) == 1
    d = get_windows({"a": x, "b": x * 2}, 88, 78)
    assert len(d) == 23 and torch.equal(d[3]["b"], w[3] * 2)


def test_reference_arm_prints_one_contract_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-batch", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1


def test_model_protocol_never_reuses_a_freed_windows_conditioning():
    """SEAM #1 caches the step-invariant window work by the IDENTITY of the caller's tensors (strong references), not by
    address: a next window whose freshly allocated tensors land on the freed addresses must be prepared again."""
    import torch
    from diffsheg_b200.engine import FusedUniDiffuser
    eng = FusedUniDiffuser.__new__(FusedUniDiffuser)   # host logic only: no library, no GPU
    eng.device = torch.device("cpu")
    prepared = []
    eng.prepare_window = lambda a, h, p: (prepared.append((a, h, p)), setattr(eng, "_window_key", None))
    eng.denoise = lambda x, t0, a, b: x
    eng._f32 = lambda t: t

    def call(mel, hub, pid):
        return eng(torch.zeros(1, 4, 6), torch.full((1,), 40), sqrt_alphas=(1.5, 1.1), audio_emb=mel, person_id=pid,
                   add_cond={"pretrain_aud_feat": hub})

    mel, hub, pid = torch.zeros(1, 4, 8), torch.zeros(1, 4, 16), torch.zeros(1, 4)
    call(mel, hub, pid)
    call(mel, hub, pid)
    assert len(prepared) == 1                                  # same objects, same versions: one preparation per window
    mel.add_(1.0)
    call(mel, hub, pid)
    assert len(prepared) == 2                                  # in-place update of the conditioning is seen (Tensor._version)
    addr = mel.data_ptr()
    del mel
    for _ in range(64):                                        # same shape, usually the same address, version 0 again
        mel2 = torch.zeros(1, 4, 8)
        if mel2.data_ptr() == addr:
            break
    call(mel2, hub, pid)
    assert len(prepared) == 3 and prepared[-1][0] is mel2      # a NEW tensor is a new window, whatever its address


def test_patched_trainer_keeps_the_training_half_of_the_reference_object():
    cfg = synth.make_cfg("show")

    class RefDiffusion:   # stands for models/gaussian_diffusion.py::GaussianDiffusion (training_losses gd:1319, q_sample gd:423)
        def training_losses(self, *a, **k):
            return "ref-losses"

    class Trainer:
        def __init__(self, opt):
            self.opt, self.diffusion, self.diffusion_ddim_val = opt, RefDiffusion(), RefDiffusion()

    tr = patch_trainer(Trainer(synth.make_opt(cfg, ddim=True)))
    assert tr.diffusion.training_losses() == "ref-losses" and tr.diffusion_ddim_val.training_losses() == "ref-losses"
    assert tr.diffusion.num_timesteps == 1000                   # the fused object's own attributes win
    with pytest.raises(AttributeError):
        tr.diffusion.no_such_thing
    with pytest.raises(AttributeError):                          # without a reference object there is nothing to fall back to
        FusedGaussianDiffusion(betas=np.linspace(1e-4, 0.02, 10)).training_losses


def test_generate_long_rejects_fix_very_first_without_motions():
    import torch
    from diffsheg_b200 import generate_long
    opt = synth.make_opt(synth.make_cfg("show"), ddim=True)
    opt.overlap_len, opt.fix_very_first = 10, True
    with pytest.raises(ValueError):
        generate_long(opt, None, None, torch.zeros(1, 200, 128), torch.zeros(1, 4), 232, {"pretrain_aud_feat": torch.zeros(1, 200, 1024)})
