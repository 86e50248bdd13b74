"""CPU test of the PRODUCT's sampler orchestration (diffsheg_b200/diffusion.py: schedules, host scalars, RNG draw order,
RePaint / undo / DDPM control flow) against the REAL reference's loop outputs (tests/golden).

The CUDA pieces are stubbed: the C-ABI step kernels by torch expressions with the kernels' exact argument contract
(include/diffsheg_b200.h), the engine by the oracle denoiser.  Nothing of the product is modified -- the stubs are
monkeypatched in -- so this pins everything ABOVE the C ABI on a machine without a GPU.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import diffsheg_b200.diffusion as D
from diffsheg_b200 import FusedGaussianDiffusion, FusedSpacedDiffusion, FusedUniDiffuser, get_named_beta_schedule, space_timesteps, synth
from oracle.denoiser import unidiffuser_forward


class FakeLib:
    """Torch restatement of the four sampler-step entry points; tensors arrive in place of device pointers."""

    def dsheg_ddim_step(self, x, eps, out, n, T, Dm, a, b, sacp, s1m, gt, mask, noise2, blend, ov, pred, stream):
        f = lambda v: torch.tensor(np.float32(v))
        px = f(a) * x - f(b) * eps
        e2 = (f(a) * x - px) / f(b)
        s = px * f(sacp) + f(s1m) * e2
        if mask is not None:
            wg = f(sacp) * gt + f(s1m) * noise2 if noise2 is not None else gt.clone()
            if blend:
                lw = torch.linspace(0, 1, ov).view(1, -1, 1)
                wg[:, :ov] = wg[:, :ov] * (1 - lw) + s[:, :ov] * lw
            s = torch.where(mask.bool(), wg, s)
        out.copy_(s)
        return 0

    def dsheg_undo_step(self, x, noise, out, n, c1, c2, stream):
        out.copy_(torch.tensor(np.float32(c1)) * x + torch.tensor(np.float32(c2)) * noise)
        return 0

    def dsheg_ddpm_step(self, x, eps, noise, out, n, a, b, k1, k2, sigma, pred, stream):
        f = lambda v: torch.tensor(np.float32(v))
        px = f(a) * x - f(b) * eps
        out.copy_((f(k1) * px + f(k2) * x) + f(sigma) * noise)
        return 0

    def dsheg_repaint_merge(self, x, gt, mask, noise, out, n, c1, c2, stream):
        f = lambda v: torch.tensor(np.float32(v))
        out.copy_(torch.where(mask.bool(), f(c1) * gt + f(c2) * noise, x))
        return 0


class FakeLibModule:
    _L = FakeLib()

    @classmethod
    def lib(cls):
        return cls._L

    @staticmethod
    def check(rc, handle=None, what=""):
        assert rc == 0, what


class OracleEngine(FusedUniDiffuser):
    """Stands in for the CUDA engine: same methods the sampler uses, arithmetic by the oracle (CPU, fp32)."""

    def __init__(self, sd, cfg):  # no super().__init__: no CUDA library / device
        self.sd, self.cfg = sd, cfg
        self.device = torch.device("cpu")
        self.cond_scale = cfg["cond_scale"]
        self.max_batch, self.max_frames = 1 << 30, 1 << 30
        self.calls = []

    def __del__(self):
        pass

    def prepare_window(self, mel, hubert, person_id):
        self.mel, self.hub, self.pid = mel, hubert, person_id

    def denoise(self, x, t_orig, a, b, cond_scale=None, out=None):
        cfg = dict(self.cfg, cond_scale=self.cond_scale if cond_scale is None else cond_scale)
        ts = torch.full((x.shape[0],), int(t_orig), dtype=torch.long)
        self.calls.append(int(t_orig))
        with torch.no_grad():
            eps = unidiffuser_forward(self.sd, cfg, x, ts, (torch.tensor(np.float32(a)), torch.tensor(np.float32(b))),
                                      self.mel, self.pid, self.hub)
        out.copy_(eps)
        return out


class EmulatedStepLib:
    """The REAL sampler-step entry points -- csrc/engine.cu's argument checks and launches, csrc/sampler.cuh's kernels -- compiled for
    the CPU emulator (tests/emu/emu_engine.cpp).  Tensors arrive in place of device pointers; their host addresses ARE the emulated
    device addresses."""

    def __init__(self):
        import emu
        self.L = emu.engine_lib()

    @staticmethod
    def _p(t, dtype=torch.float32):
        if t is None:
            return None
        assert t.dtype == dtype and t.is_contiguous() and t.device.type == "cpu", (t.dtype, t.is_contiguous())
        return ctypes.c_void_p(t.data_ptr())

    def dsheg_ddim_step(self, x, eps, out, n, T, Dm, a, b, sacp, s1m, gt, mask, noise2, blend, ov, pred, stream):
        p = self._p
        return self.L.dsheg_ddim_step(p(x), p(eps), p(out), n, T, Dm, a, b, sacp, s1m, p(gt), p(mask, torch.uint8), p(noise2), blend, ov, p(pred), None)

    def dsheg_undo_step(self, x, noise, out, n, c1, c2, stream):
        p = self._p
        return self.L.dsheg_undo_step(p(x), p(noise), p(out), n, c1, c2, None)

    def dsheg_ddpm_step(self, x, eps, noise, out, n, a, b, k1, k2, sigma, pred, stream):
        p = self._p
        return self.L.dsheg_ddpm_step(p(x), p(eps), p(noise), p(out), n, a, b, k1, k2, sigma, p(pred), None)

    def dsheg_repaint_merge(self, x, gt, mask, noise, out, n, c1, c2, stream):
        p = self._p
        return self.L.dsheg_repaint_merge(p(x), p(gt), p(mask, torch.uint8), p(noise), p(out), n, c1, c2, None)


class EmulatedLibModule(FakeLibModule):
    _L = None

    @classmethod
    def lib(cls):
        if cls._L is None:
            cls._L = EmulatedStepLib()
        return cls._L


@pytest.fixture(params=["torch-stubs", "emulated-kernels"])
def cpu_sampler(monkeypatch, request):
    """torch-stubs: the step kernels restated in torch (pins the orchestration alone).  emulated-kernels: the product's loops drive the
    real C entry points and kernel sources on the CPU emulator -- everything of the sampler path but the hardware."""
    monkeypatch.setattr(D, "_lib", FakeLibModule if request.param == "torch-stubs" else EmulatedLibModule)
    monkeypatch.setattr(D, "_ptr", lambda t: t)
    monkeypatch.setattr(D, "_stream", lambda device=None: None)


class EmulatedEngine(FusedUniDiffuser):
    """The sampler's view of the engine, served by the WHOLE emulated engine (tests/emu: csrc/engine.cu + every kernel on the CPU,
    behind the production C ABI).  The product's own __call__ / window-cache logic (diffsheg_b200/engine.py) runs unmodified on top."""

    def __init__(self, sd, cfg, precision, max_batch, max_frames):  # no super().__init__: no CUDA library / device
        import emu
        self.e = emu.EmuEngine(sd, cfg, precision=precision, max_batch=max_batch, max_frames=max_frames)
        self.cfg, self.device, self.cond_scale = dict(cfg), torch.device("cpu"), float(cfg["cond_scale"])
        self.max_batch, self.max_frames, self._window = max_batch, max_frames, None

    def __del__(self):
        pass

    def prepare_window(self, mel, hubert, person_id):
        self.e.prepare_window(mel, hubert, person_id)
        self._window, self._window_key = (mel.shape[0], mel.shape[1]), None

    def denoise(self, x, t_orig, a, b, cond_scale=None, out=None):
        res = self.e.denoise(x, t_orig, a, b, self.cond_scale if cond_scale is None else cond_scale)
        if out is None:
            return res
        out.copy_(res)
        return out


def test_product_sampling_loop_over_the_emulated_engine_and_step_kernels(monkeypatch):
    """Everything of the hot path but the hardware: the product's FusedSpacedDiffusion.ddim_sample_loop (Python step loop, host
    scalars, model(x, ts, **kwargs) protocol and window cache of FusedUniDiffuser.__call__) drives the emulated engine (graph replay
    from the third call on) and the emulated step kernels, against the oracle's loop on the same x_T.  SHOW under CFG, one layer per
    net, a 6-step DDIM schedule."""
    from oracle import diffusion as odiff
    monkeypatch.setattr(D, "_lib", EmulatedLibModule)
    monkeypatch.setattr(D, "_ptr", lambda t: t)
    monkeypatch.setattr(D, "_stream", lambda device=None: None)
    B, T = 1, 6
    cfg = synth.make_cfg("show", num_layers=1)
    sd = synth.make_state_dict(cfg, seed=1)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    eng = EmulatedEngine(sd, cfg, "fp32", B, T)
    replays0 = eng.e.graph_launches()
    opt = synth.make_opt(cfg, timestep_respacing="ddim6")
    diff = FusedSpacedDiffusion(space_timesteps(1000, "ddim6"), opt=opt, betas=get_named_beta_schedule("linear", 1000))
    kw = dict(audio_emb=inp["mel"], length=None, person_id=inp["person_id"], add_cond={"pretrain_aud_feat": inp["hubert"]}, y={}, pe_type="pe_sinu")
    out = diff.ddim_sample_loop(eng, (B, T, cfg["net_dim_pose"]), noise=inp["x_T"].clone(), clip_denoised=False, model_kwargs=kw)
    assert diff.last_stats == {"denoise_calls": 6, "undo_steps": 0}
    assert eng.e.graph_launches() - replays0 == 5          # the first call runs eagerly, the second captures, calls 2 .. 6 replay
    eng.e.close()
    with torch.no_grad():
        den = odiff.make_denoise(sd, cfg, inp["mel"], inp["person_id"], inp["hubert"])
        want = odiff.OracleDiffusion(1000, "ddim6").ddim_sample_loop(den, (B, T, cfg["net_dim_pose"]), y={}, noise=inp["x_T"].clone())
    assert relmax(out.numpy(), want.numpy()) < 2e-5


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _inpaint(cfg, B, T, ov):
    if ov == 0:
        return {}
    g = torch.Generator().manual_seed(5)
    gt = torch.zeros(B, T, cfg["net_dim_pose"])
    gt[:, :ov] = torch.randn(B, ov, cfg["net_dim_pose"], generator=g)
    mask = torch.zeros(B, T, cfg["net_dim_pose"], dtype=torch.bool)
    mask[:, :ov] = True
    return {"gt": gt, "outpainting_mask": mask}


@pytest.mark.parametrize("fn,name,B,T,ov,over,calls,undos", [
    ("loop_beat_B2_T34_ov0_ddim25.npz", "beat", 2, 34, 0, {}, 25, 0),
    ("loop_beat_B1_T34_ov4_ddim25_jn2.npz", "beat", 1, 34, 4, dict(jump_n_sample=2), 27, 12),
    ("loop_show_B2_T88_ov10_ddim25.npz", "show", 2, 88, 10, {}, 63, 48),
])
def test_product_ddim_loop_orchestration_matches_reference(cpu_sampler, golden_dir, fn, name, B, T, ov, over, calls, undos):
    g = np.load(os.path.join(golden_dir, fn))
    cfg = synth.make_cfg(name)
    eng = OracleEngine(synth.make_state_dict(cfg, seed=1), cfg)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    opt = synth.make_opt(cfg, overlap_len=ov, **over)
    diff = FusedSpacedDiffusion(space_timesteps(1000, "ddim25"), opt=opt, betas=get_named_beta_schedule("linear", 1000))
    kw = dict(audio_emb=inp["mel"], length=None, person_id=inp["person_id"], add_cond={"pretrain_aud_feat": inp["hubert"]},
              y=_inpaint(cfg, B, T, ov), pe_type="pe_sinu")
    torch.manual_seed(int(g["seed"]))   # the reference drew x_T and every per-step noise from this generator state (F11)
    out = diff.ddim_sample_loop(eng, (B, T, cfg["net_dim_pose"]), clip_denoised=False, model_kwargs=kw)
    assert diff.last_stats == {"denoise_calls": calls, "undo_steps": undos}
    assert eng.calls[0] == (960 if ov == 0 else 560) and eng.calls[-1] == 0   # respaced 24 -> 960; repaint starts at 14 -> 560 (F6)
    assert relmax(out.numpy(), g["sample"]) < 1e-3


def test_product_ddpm_loop_orchestration_matches_reference(cpu_sampler, golden_dir):
    g = np.load(os.path.join(golden_dir, "loop_beat_B2_T34_ov0_ddpm40.npz"))
    cfg = synth.make_cfg("beat")
    eng = OracleEngine(synth.make_state_dict(cfg, seed=1), cfg)
    inp = synth.make_inputs(cfg, 2, 34, seed=2)
    diff = FusedGaussianDiffusion(opt=synth.make_opt(cfg, ddim=False, diffusion_steps=40), betas=get_named_beta_schedule("linear", 40))
    kw = dict(audio_emb=inp["mel"], length=None, person_id=inp["person_id"], add_cond={"pretrain_aud_feat": inp["hubert"]},
              y=None, pe_type="pe_sinu")   # y=None: the reference would crash (gd:810); accepted as {}
    torch.manual_seed(int(g["seed"]))
    out = diff.p_sample_loop(eng, (2, 34, cfg["net_dim_pose"]), clip_denoised=False, model_kwargs=kw)
    assert diff.last_stats == {"denoise_calls": 40, "undo_steps": 0} and eng.calls == list(range(39, -1, -1))
    assert relmax(out.numpy(), g["sample"]) < 1e-3


# ------------------------------------------------------------------------------------------------------------------------
# control flow the network-driven goldens do not reach: the DDPM RePaint loop (gd:843-920) and --same_overlap_noisy
# (gd:1040-1042, :1058-1060, :1155-1159).  Fixtures: the REAL reference sampler driving a closed-form toy denoiser
# (tests/golden/make_golden_sampler.py); product and oracle are both pinned to them here.
# ------------------------------------------------------------------------------------------------------------------------
TB, TT, TD, TOV = 2, 12, 6, 3


def toy_eps(x, t_orig):
    return 0.9 * torch.tanh(0.7 * x.flip(-1) + 0.3 * x.roll(1, 1) + 0.002 * float(t_orig))


class ToyEngine(FusedUniDiffuser):
    """The sampler's view of an engine, arithmetic by the toy denoiser (any device)."""

    def __init__(self, device="cpu"):
        self.device = torch.device(device)
        self.cond_scale, self.max_batch, self.max_frames, self.calls = 1.0, 1 << 30, 1 << 30, []

    def __del__(self):
        pass

    def prepare_window(self, mel, hubert, person_id):
        pass

    def denoise(self, x, t_orig, a, b, cond_scale=None, out=None):
        self.calls.append(int(t_orig))
        out.copy_(toy_eps(x, t_orig))
        return out


def _toy_kwargs(y, device="cpu"):
    z = torch.zeros(TB, TT, 1, device=device)
    return dict(audio_emb=z, length=None, person_id=torch.zeros(TB, 1, device=device), add_cond={"pretrain_aud_feat": z}, y=y, pe_type="pe_sinu")


def _toy_inpaint(seed=5):
    g = torch.Generator().manual_seed(seed)
    gt = torch.zeros(TB, TT, TD)
    gt[:, :TOV] = torch.randn(TB, TOV, TD, generator=g)
    mask = torch.zeros(TB, TT, TD, dtype=torch.bool)
    mask[:, :TOV] = True
    return gt, mask


def test_product_and_oracle_ddpm_repaint_loop_match_reference(cpu_sampler, golden_dir):
    """p_sample_loop with known frames: get_schedule_jump_paper (t_T = 250, jump 10 x 10): 2 410 denoiser calls, 2 160 re-noise
    steps with t_shift = 1 (gd:912-917), the known frames merged BEFORE every call but the first (gd:727-745)."""
    from oracle import diffusion as odiff
    g = np.load(os.path.join(golden_dir, "sampler_ddpm_harmonize_toy.npz"))
    gt, mask = _toy_inpaint()
    eng = ToyEngine()
    opt = synth.make_opt(synth.make_cfg("beat"), ddim=False, overlap_len=TOV)
    diff = FusedGaussianDiffusion(opt=opt, betas=get_named_beta_schedule("linear", 1000))
    torch.manual_seed(int(g["seed"]))
    out = diff.p_sample_loop(eng, (TB, TT, TD), clip_denoised=False, model_kwargs=_toy_kwargs({"gt": gt.clone(), "outpainting_mask": mask}))
    assert diff.last_stats == {"denoise_calls": int(g["calls"]), "undo_steps": 2160}
    assert (eng.calls[0], eng.calls[-1]) == (int(g["first_t"]), int(g["last_t"]))
    assert relmax(out.numpy(), g["sample"]) < 2e-4
    d = odiff.OracleDiffusion(1000, None, overlap_len=TOV)
    torch.manual_seed(int(g["seed"]))
    want = d.p_sample_loop(lambda x, t, a, b: toy_eps(x, t), (TB, TT, TD), y={"gt": gt.clone(), "outpainting_mask": mask})
    assert relmax(want.numpy(), g["sample"]) < 2e-4


def _same_overlap_chain(loop, alias_ok=True):
    """beat:1003-1028: three windows; window ii > 0 repaints its head from the tails window ii-1 saved at every step."""
    samples, prev, tail = [], None, None
    for ii in range(3):
        gtw = torch.zeros(TB, TT, TD)
        maskw = torch.zeros(TB, TT, TD, dtype=torch.bool)
        y = {"gt": gtw, "outpainting_mask": maskw, "clip_idx": ii}
        if ii > 0:
            maskw[:, :TOV] = True
            gtw[:, :TOV] = prev[:, -TOV:]
            y["previous_noisy_tail"] = tail
        out = loop(y)
        prev, tail = out["sample"], out["saved_noisy_tail"]
        samples.append(prev.clone())
    return torch.stack(samples)


def test_product_and_oracle_same_overlap_noisy_chain_match_reference(cpu_sampler, golden_dir):
    from oracle import diffusion as odiff
    g = np.load(os.path.join(golden_dir, "sampler_same_overlap_noisy_toy.npz"))
    eng = ToyEngine()
    opt = synth.make_opt(synth.make_cfg("beat"), overlap_len=TOV, same_overlap_noisy=True)
    diff = FusedSpacedDiffusion(space_timesteps(1000, "ddim25"), opt=opt, betas=get_named_beta_schedule("linear", 1000))
    torch.manual_seed(int(g["seed"]))
    got = _same_overlap_chain(lambda y: diff.ddim_sample_loop(eng, (TB, TT, TD), clip_denoised=False, model_kwargs=_toy_kwargs(y)))
    assert len(eng.calls) == int(g["calls"]) == 25 + 63 + 63
    assert relmax(got.numpy(), g["samples"]) < 2e-4
    d = odiff.OracleDiffusion(1000, "ddim25", overlap_len=TOV, same_overlap_noisy=True)
    torch.manual_seed(int(g["seed"]))
    want = _same_overlap_chain(lambda y: d.ddim_sample_loop(lambda x, t, a, b: toy_eps(x, t), (TB, TT, TD), y=y))
    assert relmax(want.numpy(), g["samples"]) < 2e-4


def test_generate_long_threads_the_noisy_tails_between_windows(cpu_sampler):
    """The product's own window loop (trainer.generate_long) under --same_overlap_noisy equals the hand-written beat:1003-1028 chain."""
    from diffsheg_b200 import generate_long
    frames = TT + 2 * (TT - TOV)
    opt = synth.make_opt(synth.make_cfg("beat"), overlap_len=TOV, same_overlap_noisy=True, n_poses=TT)
    diff = FusedSpacedDiffusion(space_timesteps(1000, "ddim25"), opt=opt, betas=get_named_beta_schedule("linear", 1000))
    z = torch.zeros(TB, frames, 1)
    torch.manual_seed(31)
    got = generate_long(opt, ToyEngine(), diff, z, torch.zeros(TB, 1), TD, {"pretrain_aud_feat": z})
    diff2 = FusedSpacedDiffusion(space_timesteps(1000, "ddim25"), opt=opt, betas=get_named_beta_schedule("linear", 1000))
    eng2 = ToyEngine()
    torch.manual_seed(31)
    chain = _same_overlap_chain(lambda y: diff2.ddim_sample_loop(eng2, (TB, TT, TD), clip_denoised=False, model_kwargs=_toy_kwargs(y)))
    want = torch.cat([chain[0][:, :TT - TOV], chain[1][:, :TT - TOV], chain[2]], 1)
    assert got.shape == (TB, frames, TD) and torch.equal(got, want)
