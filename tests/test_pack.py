"""CPU checks of the host logic: the packer's folds and the engine's dataflow (emulated in torch over the
packed tensors, tests/emulate.py) against the oracle; the C-ABI library loads and exports every symbol."""
import os
import re

import pytest
import torch

from diffsheg_b200 import synth
from diffsheg_b200.pack import pack_state_dict
from emulate import Emu
from oracle.denoiser import unidiffuser_forward


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


@pytest.mark.parametrize("name,B,T,cs", [("show", 2, 20, 1.25), ("show", 1, 9, 1.0), ("beat", 2, 12, 1.0)])
def test_packed_dataflow_matches_oracle_fp64(name, B, T, cs):
    cfg = synth.make_cfg(name, cond_scale=cs)
    sd = synth.make_state_dict(cfg, seed=1)
    inp = synth.make_inputs(cfg, B, T, seed=2)
    packed = pack_state_dict(sd, cfg, "fp32", max_frames=T)
    got = Emu(packed, cfg).denoise(inp["x_T"], 520, 1.7, 1.3, cs, inp["mel"], inp["hubert"], inp["person_id"])
    ts = torch.full((B,), 520, dtype=torch.long)
    with torch.no_grad():
        want = unidiffuser_forward(sd, cfg, inp["x_T"].double(), ts, (torch.tensor(1.7, dtype=torch.float64),
                                   torch.tensor(1.3, dtype=torch.float64)), inp["mel"], inp["person_id"], inp["hubert"],
                                   dtype=torch.float64)
    assert relmax(got, want) < 5e-6  # fp32-stored weights vs fp64 reference weights


def test_bf16_pack_is_rounded_fp32_pack_and_csum_consistent():
    cfg = synth.make_cfg("beat")
    sd = synth.make_state_dict(cfg, seed=1)
    p32 = pack_state_dict(sd, cfg, "fp32", max_frames=34)
    p16 = pack_state_dict(sd, cfg, "bf16", max_frames=34)
    assert p32.keys() == p16.keys()
    w32, w16 = p32["ges.l3.feat1.w"][0], p16["ges.l3.feat1.w"][0]
    assert w16.dtype == torch.bfloat16 and w16.shape == w32.shape == (1024, 512 + 256 + 128 + 64)
    assert torch.equal(w16, w32.to(torch.bfloat16))
    # csum must be the row sums of the STORED (rounded) weight so the LayerNorm fold cancels exactly
    assert torch.allclose(p16["ges.l3.feat1.csum"][0].double(), w16.double().sum(1), rtol=1e-6, atol=1e-6)
    # K padding columns of every segment are zero
    assert float(w32[:, 512 + 256 + 128 + 51:].abs().max()) == 0.0


def test_static_softmax_shifts_bound_every_input():
    """pack.py:expo_shift -- the ACT_EXPO epilogue replaces the running maxima of softmax_d(Q) / softmax_t(K) (tr:122-123) by
    static shifts; the packer must PROVE |value - shift| <= EXPO_LIMIT for every hidden row.  Checked against the worst-case
    rows (LayerNorm output aligned with a weight row: the Cauchy-Schwarz bound is attained up to the 1 % slack) and against
    weights for which no safe shift exists (the optional tensor must then be absent)."""
    from diffsheg_b200.pack import EXPO_LIMIT, expo_shift
    cfg = synth.make_cfg("show")
    sd = synth.make_state_dict(cfg, seed=1)
    packed = pack_state_dict(sd, cfg, "bf16", max_frames=8)
    D = cfg["latent_dim"]
    for name in ("ges.l0", "exp.l7"):
        W = packed[name + ".qkv.w"][0].double()
        b = packed[name + ".qkv.b"][0].double()
        es = packed[name + ".qkv.eshift"][0].double()
        assert es.shape == (2 * D,) and torch.equal(es[D:], b[D:2 * D].float().double())
        # Q: ONE shift per 64-channel head (a row softmax tolerates nothing finer): the head's mean folded bias
        assert torch.equal(es[:D].reshape(-1, 64), es[:D].reshape(-1, 64)[:, :1].expand(-1, 64))
        assert torch.allclose(es[:D].reshape(-1, 64)[:, 0], b[:D].reshape(-1, 64).mean(1), atol=1e-6)
        g = torch.Generator().manual_seed(3)
        worst = 0.0
        for n in list(range(0, 2 * D, 97)) + [2 * D - 1]:
            wc = W[n] - W[n].mean()
            for sign in (1.0, -1.0):
                # the hidden row whose LayerNorm output is +-sqrt(P) * wc / ||wc|| (any offset / scale of h gives the same xhat)
                h = 5.0 + 3.0 * sign * wc / wc.norm()
                xhat = (h - h.mean()) / torch.sqrt(h.var(unbiased=False) + 1e-5)
                v = W[:2 * D] @ xhat + b[:2 * D]
                worst = max(worst, float((v - es).abs().max()))
        for _ in range(4):   # random rows are far inside
            h = torch.randn(D, generator=g).double() * 10 + 3
            xhat = (h - h.mean()) / torch.sqrt(h.var(unbiased=False) + 1e-5)
            assert float((W[:2 * D] @ xhat + b[:2 * D] - es).abs().max()) < worst
        assert worst <= EXPO_LIMIT
        # the proven radius is attained (to the slack) by the aligned rows: the bound is tight, not vacuous
        R = (W[:2 * D] - W[:2 * D].mean(1, keepdim=True)).norm(dim=1) * D ** 0.5
        assert worst > 0.9 * float(R.max())
    # no safe static shift: large K weights, or a large Q bias (Q is shifted by 0) -> no tensor, the engine keeps the max-search kernels
    W = packed["ges.l0.qkv.w"][0].double()
    b = packed["ges.l0.qkv.b"][0].double()
    assert expo_shift(W, b, D) is not None
    assert expo_shift(W * 40.0, b, D) is None
    bq = b.clone()
    bq[5] = 1.1 * EXPO_LIMIT     # one channel far from its head's mean bias: no per-head shift can cover it
    assert expo_shift(W, bq, D) is None
    bq = b.clone()
    bq[:64] += 500.0             # a whole head offset by a constant is harmless: the head's shift absorbs it
    assert expo_shift(W, bq, D) is not None
    bk = b.clone()
    bk[D + 5] = 1e4     # a K bias of any size is harmless: it IS the shift
    assert expo_shift(W, bk, D) is not None


def test_library_exports_every_declared_symbol():
    from diffsheg_b200 import _lib
    _lib.build()
    L = _lib.lib()
    header = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(dsheg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name


def test_product_raises_without_cuda():
    """No CPU fallback: on a machine without a GPU the engine must fail loudly, not compute elsewhere."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffsheg_b200 import FusedUniDiffuser
    cfg = synth.make_cfg("beat")
    with pytest.raises(RuntimeError):
        FusedUniDiffuser(synth.make_state_dict(cfg), cfg, precision="fp32", max_batch=1, max_frames=34)
