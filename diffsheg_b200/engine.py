"""FusedUniDiffuser: the reference denoiser's calling protocol over the CUDA engine.

SEAM #1 of SURVEY section 8b: ``model(x, ts, **model_kwargs)`` with
``model_kwargs = {audio_emb, length, person_id, add_cond{'pretrain_aud_feat'}, y, pe_type}``
plus the injected ``sqrt_alphas`` (trainers/ddpm_show_trainer.py:175-182,
models/gaussian_diffusion.py:527-536, models/transformer.py:728).
PyTorch is used for device memory and streams only; the arithmetic runs in
libdiffsheg_b200.so through ctypes with raw device pointers.
"""
import ctypes

import torch

from . import _lib
from .pack import pack_state_dict

_CFG_KEYS = ("dim_pose", "expression_dim", "audio_dim", "hubert_dim", "aud_latent_dim", "latent_dim",
             "num_layers", "num_heads", "ff_size", "style_dim")

_SUPPORTED = dict(unidiffuser=True, model_base="transformer_encoder", addHubert=True, encode_hubert=True, expAddHubert=False,
                  addWav2Vec2=False, addTextCond=False, addEmoCond=False, no_style=False, ExprID_off=False,
                  ExprID_off_uncond=False, fix_head_var=False, separate=None)


def cfg_from_opt(opt, **over):
    """Frozen engine configuration from the reference's mutable ``opt`` Namespace
    (runner.py:124-222, options/base_options.py); rejects unsupported combinations up front."""
    for k, v in _SUPPORTED.items():
        if hasattr(opt, k) and getattr(opt, k) != v:
            raise NotImplementedError(f"diffsheg_b200 supports only opt.{k}={v!r} (got {getattr(opt, k)!r})")
    if getattr(opt, "model_mean_type", "epsilon") != "epsilon":
        raise NotImplementedError("only epsilon prediction is supported")
    cond_projection = getattr(opt, "cond_projection", "mlp_includeX")
    if cond_projection not in _lib.COND_PROJECTION:   # 'none': the reference's own layers raise for it (tr:323-324)
        raise NotImplementedError(f"diffsheg_b200 supports opt.cond_projection in {sorted(_lib.COND_PROJECTION)} (got {cond_projection!r})")
    cfg = dict(dim_pose=opt.dim_pose, expression_dim=opt.expression_dim,
               audio_dim=getattr(opt, "audio_dim", 128), hubert_dim=1024,
               aud_latent_dim=getattr(opt, "audio_latent_dim", 256), latent_dim=getattr(opt, "latent_dim", 512),
               num_layers=getattr(opt, "num_layers", 8), num_heads=8, ff_size=1024,
               style_dim=getattr(opt, "style_dim", 4), hubert_enc_dim=128,
               classifier_free=bool(getattr(opt, "classifier_free", False)),
               cond_scale=float(getattr(opt, "cond_scale", 1.0)), n_poses=getattr(opt, "n_poses", 88),
               cond_projection=cond_projection, cond_residual=bool(getattr(opt, "cond_residual", True)))
    cfg.update(over)
    cfg["net_dim_pose"] = cfg["dim_pose"] + cfg["expression_dim"]
    return cfg


def engine_config(cfg, precision, max_batch, max_frames):
    """The ``dsheg_config`` struct of include/diffsheg_b200.h for a configuration dict (``synth.make_cfg`` / ``cfg_from_opt``)."""
    return _lib.Config(abi_version=_lib.ABI_VERSION, classifier_free=int(bool(cfg["classifier_free"])),
                       precision=_lib.PREC[precision], max_batch=int(max_batch), max_frames=int(max_frames),
                       cond_projection=_lib.COND_PROJECTION[cfg.get("cond_projection", "mlp_includeX")],
                       no_cond_residual=int(not cfg.get("cond_residual", True)),
                       **{k: int(cfg[k]) for k in _CFG_KEYS})


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device=None):
    """The caller's current stream ON THE GIVEN DEVICE (an engine may live on a GPU that is not torch's current one)."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class FusedUniDiffuser:
    """UniDiffuser (transformer.py:590-770) on hand-written sm_100a kernels.

    precision: "bf16" (tcgen05 GEMMs, bf16 activations, fp32 accumulation/statistics) or "fp32"
    (strict parity mode).  The workspace is sized once for (max_batch, max_frames).
    """

    def __init__(self, state_dict, cfg, precision="bf16", max_batch=1, max_frames=None, device=0):
        self.cfg = dict(cfg)
        self.precision = precision
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.max_batch = int(max_batch)
        self.max_frames = int(max_frames or cfg["n_poses"])
        self.cond_scale = float(cfg.get("cond_scale", 1.0))
        L = _lib.lib()
        c = engine_config(cfg, precision, self.max_batch, self.max_frames)
        h = ctypes.c_void_p()
        _lib.check(L.dsheg_create(ctypes.byref(c), self.device.index, ctypes.byref(h)), None, "dsheg_create")
        self._h = h
        self._L = L
        packed = pack_state_dict(state_dict, self.cfg, precision, self.max_frames)
        for name, (t, dt) in packed.items():
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            _lib.check(L.dsheg_load_tensor(h, name.encode(), ctypes.c_void_p(t.data_ptr()), dt, shape, t.dim()),
                       h, f"dsheg_load_tensor({name})")
        _lib.check(L.dsheg_finalize_weights(h), h, "dsheg_finalize_weights")
        self._window = None
        self._keep = None

    @classmethod
    def from_module(cls, module, opt=None, **kw):
        """Harvest weights from a reference UniDiffuser nn.Module (possibly DDP-wrapped)."""
        inner = getattr(module, "module", module)
        opt = opt or inner.opt
        return cls(inner.state_dict(), cfg_from_opt(opt), **kw)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._L.dsheg_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- engine calls ---------------------------------------------------------------------------
    def _f32(self, t):
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def prepare_window(self, mel, hubert, person_id):
        """Step-invariant work for one window: person-id MLP, hubert conv stack, mel staging."""
        mel, hubert, person_id = self._f32(mel), self._f32(hubert), self._f32(person_id)
        if person_id.dim() == 1:
            person_id = person_id.unsqueeze(0)  # tr:502-503
        B, T = mel.shape[0], mel.shape[1]
        assert hubert.shape[:2] == (B, T) and person_id.shape[0] == B, "conditioning shapes disagree"
        with torch.cuda.device(self.device):
            _lib.check(self._L.dsheg_prepare_window(self._h, _ptr(mel), _ptr(hubert), _ptr(person_id), B, T,
                                                    _stream(self.device)), self._h, "dsheg_prepare_window")
        self._keep = (mel, hubert, person_id)  # async kernels read these: keep them alive
        self._window = (B, T)
        self._window_key = None                # an explicit prepare_window supersedes whatever __call__ cached

    def denoise(self, x, t_orig, a, b, cond_scale=None, out=None):
        """eps = UniDiffuser.forward(x, [t_orig]*B, sqrt_alphas=(a, b), <window conditioning>)."""
        assert self._window is not None, "call prepare_window first"
        B, T = self._window
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        assert tuple(x.shape) == (B, T, self.cfg["net_dim_pose"]), (x.shape, self._window)
        if out is None:
            out = torch.empty_like(x)
        s = self.cond_scale if cond_scale is None else float(cond_scale)
        with torch.cuda.device(self.device):
            _lib.check(self._L.dsheg_denoise(self._h, _ptr(x), int(t_orig), float(a), float(b), s, _ptr(out),
                                             _stream(self.device)), self._h, "dsheg_denoise")
        return out

    def launch_count(self):
        return int(self._L.dsheg_launch_count(self._h))

    def profile_begin(self):
        _lib.check(self._L.dsheg_profile_begin(self._h), self._h, "dsheg_profile_begin")

    def profile_end(self):
        """-> {class: dict(ms, work, count)} for gemm (work = FLOPs), attention and rowwise (work = bytes)."""
        ms, work, cnt = (ctypes.c_double * 3)(), (ctypes.c_double * 3)(), (ctypes.c_int64 * 3)()
        _lib.check(self._L.dsheg_profile_end(self._h, ms, work, cnt), self._h, "dsheg_profile_end")
        return {n: dict(ms=ms[i], work=work[i], count=int(cnt[i])) for i, n in enumerate(("gemm", "attention", "rowwise"))}

    # -- reference calling protocol (SEAM #1) ---------------------------------------------------
    def __call__(self, x, ts, sqrt_alphas=None, audio_emb=None, length=None, person_id=None, add_cond=None,
                 pe_type="pe_sinu", y=None, **_):
        if pe_type != "pe_sinu":
            raise NotImplementedError("only --PE pe_sinu is supported")
        hub = (add_cond or {}).get("pretrain_aud_feat")
        if hub is None:
            raise ValueError("add_cond['pretrain_aud_feat'] (HuBERT features) is required (addHubert=True)")
        # the SAME tensor objects with the same in-place version => same window (Tensor._version counts in-place writes,
        # no device sync).  The key holds strong references to the caller's tensors: comparing addresses alone would
        # match a freed-and-reallocated buffer of the next window / clip and silently reuse stale conditioning.
        src = (audio_emb, hub, person_id)
        key = getattr(self, "_window_key", None)
        same = key is not None and all(a is b for a, b in zip(key[0], src)) and \
            key[1] == tuple(t._version for t in src) and key[2] == tuple(tuple(t.shape) for t in src)
        if not same:
            self.prepare_window(audio_emb, hub, person_id)
            self._window_key = (src, tuple(t._version for t in src), tuple(tuple(t.shape) for t in src))
        t0 = int(ts.reshape(-1)[0])
        if ts.numel() > 1 and not bool((ts == t0).all()):
            raise NotImplementedError("per-sample timesteps are not supported (the samplers use t = [i]*B)")
        a, b = sqrt_alphas
        a = float(a.reshape(-1)[0]) if torch.is_tensor(a) else float(a)
        b = float(b.reshape(-1)[0]) if torch.is_tensor(b) else float(b)
        return self.denoise(self._f32(x), t0, a, b)

    def parameters(self):  # `next(model.parameters()).device` idiom of gaussian_diffusion.py:1181
        yield torch.empty(0, device=self.device)

    def eval(self):
        return self
