"""Pack a reference ``checkpoint['encoder']`` state_dict into the tensors the CUDA engine loads.

Input contract: the UniDiffuser state_dict of the reference (keys as produced by
models/transformer.py:590-699; DDP ``module.`` prefixes are stripped like
trainers/ddpm_show_trainer.py:271-274).  All folds are done in float64 on the host and are
exact up to floating-point reassociation:

* ``*.feat1`` / ``*.qkv``: the LayerNorm in front of the Linear (tr:285-286, tr:119-125) is folded
  into it: W' = W * gamma, b' = b + W beta, csum[n] = sum_k W'[n,k]; the kernel applies
  rstd * (x W'^T - mu * csum) + b' with per-row (mu, rstd).  csum is taken over W' AS STORED
  (after bf16 rounding) so the subtraction cancels exactly.
* ``*.nullc``: under classifier-free guidance the whole feat_proj input row of the uncond half is the
  learned ``null_cond_emb`` (tr:326-332), so feat_proj(null) is one constant vector per layer.
* ``*.qkv.eshift`` (optional): static softmax shifts for the Q | K columns when ``expo_shift`` can prove the exponent range
  for every possible input (enables the ACT_EXPO epilogue of the QKV GEMM + attn_ws; layers without it keep attn_v3).
* ``*.hub``: BatchNorm1d (eval) folded into the first Conv1d of hubert_encoder (tr:436-442).
* K axes of GEMM weights are laid out per A-segment, each padded to a multiple of 64
  (h | audio_proj | hubert | expr for feat1).
* ``cfg['cond_projection']`` (tr:262-263,281-289,304-324; default ``mlp_includeX``): the ``*_excludeX`` projections drop the ``h``
  segment; the ``linear_*`` ones have a single Linear, packed as ``*.featl`` (no LayerNorm, so no csum), and their ``*.nullc`` is
  ``W null + b``.

Packed names (n in aud/exp/ges, i = layer): see ``pack_state_dict``; shapes are validated by
``dsheg_finalize_weights``.
"""
import math

import numpy as np
import torch

F32, BF16 = 0, 1

# ACT_EXPO (gemm_tc.cuh) writes softmax numerators exp(v - shift) with STATIC shifts; |v - shift| must stay below this bound
# (natural-log units) for EVERY possible input, so that neither the numerators (e^+-72 = 2e+-31, normal numbers in bf16 and fp32)
# nor their sums over up to 96 frames / 64 channels, nor the K'^T V and Q' A products with |V| <= EXPO_V_LIMIT
# (96 * e^72 * 1e3 = 2e36 < 3.4e38) leave the fp32 / bf16 exponent range
EXPO_LIMIT = 72.0
EXPO_V_LIMIT = 1.0e3


def expo_shift(stored_w, bias, D, head_dim=64):
    """Static softmax shifts for the Q | K columns of an LN-folded QKV projection, or None when no safe ones exist.

    The GEMM computes v[n] = xhat . W'[n] + b'[n] with xhat = (h - mean) * rstd, the LayerNorm-ed hidden row (tr:119-123):
    sum(xhat) = 0 and ||xhat||_2 <= sqrt(P), whatever h is.  Cauchy-Schwarz on the centred weight row gives
    |v[n] - b'[n]| <= R[n] = sqrt(P) * ||W'[n] - mean(W'[n])||_2  for every input.  softmax is shift-invariant, so
      * K (softmax over time, tr:123, one shift per column): shift = b'[n], exponent within +-R[n];
      * Q (softmax over the 64 channels of a head, tr:122, one shift per row and head): shift = mean of b' over the head,
        exponent within +-(R[n] + |b'[n] - shift|)
    are exact replacements for the running maxima as long as the exponents stay inside EXPO_LIMIT.  `stored_w` are the weights
    AS STORED (bf16-rounded), so the bound is about the numbers the tensor core multiplies.  Returns [2 D] float64."""
    P = stored_w.shape[1]
    w = stored_w[:3 * D]
    R = math.sqrt(P) * (w - w.mean(dim=1, keepdim=True)).norm(dim=1) * 1.01   # 1 %: fp32 accumulation and statistics
    q_shift = bias[:D].to(torch.float64).reshape(-1, head_dim).mean(dim=1, keepdim=True).expand(-1, head_dim).reshape(D)
    q_bound = (R[:D] + (bias[:D] - q_shift).abs()).max()
    k_bound = R[D:2 * D].max()
    v_bound = (R[2 * D:] + bias[2 * D:3 * D].abs()).max()    # |V| itself enters the products linearly
    if not (torch.isfinite(R).all() and float(q_bound) <= EXPO_LIMIT and float(k_bound) <= EXPO_LIMIT and float(v_bound) <= EXPO_V_LIMIT):
        return None
    return torch.cat([q_shift, bias[D:2 * D].to(torch.float64)])


def _r64(k):
    return (k + 63) // 64 * 64


def _pad_segments(W, widths):
    """[N, sum(widths)] -> [N, sum(round_up(w, 64))], zero padded per segment."""
    cols, off = [], 0
    for w in widths:
        seg = W[:, off:off + w]
        pad = _r64(w) - w
        if pad:
            seg = torch.cat([seg, torch.zeros(W.shape[0], pad, dtype=W.dtype)], dim=1)
        cols.append(seg)
        off += w
    assert off == W.shape[1], (off, W.shape)
    return torch.cat(cols, dim=1)


class Packer:
    def __init__(self, sd, cfg, precision):
        self.sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
        self.cfg, self.bf16 = cfg, precision == "bf16"
        self.out = {}

    def g(self, key):
        return self.sd[key].detach().to("cpu", torch.float64)

    def put_f32(self, name, t):
        self.out[name] = (t.to(torch.float32).contiguous(), F32)

    def put_w(self, name, W, widths=None):
        """GEMM weight [N, K] -> stored [N, Kp] in the GEMM element type; returns the stored values (f64)."""
        W = _pad_segments(W, widths or [W.shape[1]])
        if self.bf16:
            st = W.to(torch.bfloat16).contiguous()
            self.out[name] = (st, BF16)
            return st.to(torch.float64)
        st = W.to(torch.float32).contiguous()
        self.out[name] = (st, F32)
        return st.to(torch.float64)

    def lin(self, name, key, widths=None):
        self.put_w(name + ".w", self.g(key + ".weight"), widths)
        self.put_f32(name + ".b", self.g(key + ".bias"))

    def lin_ln_fold(self, name, W, b, gamma, beta, widths=None):
        Wp = W * gamma[None, :]
        stored = self.put_w(name + ".w", Wp, widths)
        self.put_f32(name + ".b", b + W @ beta)
        self.put_f32(name + ".csum", stored.sum(dim=1))
        return stored, b + W @ beta

    def mlp(self, name, key0, key2, pad_k=False):
        w0 = self.g(key0 + ".weight")
        if pad_k:
            w0 = _pad_segments(w0, [w0.shape[1]])
        self.put_f32(name + ".w0", w0)
        self.put_f32(name + ".b0", self.g(key0 + ".bias"))
        self.put_f32(name + ".w2", self.g(key2 + ".weight"))
        self.put_f32(name + ".b2", self.g(key2 + ".bias"))

    def layer(self, name, key, seg_widths, null_row):
        g = self.g
        if seg_widths and self.cfg.get("cond_projection", "mlp_includeX").startswith("linear"):
            W, b = g(key + ".feat_proj.weight"), g(key + ".feat_proj.bias")   # tr:281-282: one Linear(pre_proj_dim, latent_dim)
            self.put_w(name + ".featl.w", W, seg_widths)
            self.put_f32(name + ".featl.b", b)
            if null_row is not None:
                self.put_f32(name + ".nullc", W @ null_row + b)
        elif seg_widths:
            fp = key + ".feat_proj"
            gam, bet = g(fp + ".0.weight"), g(fp + ".0.bias")
            W1, b1 = g(fp + ".1.weight"), g(fp + ".1.bias")
            W2, b2 = g(fp + ".3.weight"), g(fp + ".3.bias")
            self.lin_ln_fold(name + ".feat1", W1, b1, gam, bet, seg_widths)
            self.put_w(name + ".feat2.w", W2)
            self.put_f32(name + ".feat2.b", b2)
            if null_row is not None:
                mu = null_row.mean()
                var = ((null_row - mu) ** 2).mean()
                z = (null_row - mu) / torch.sqrt(var + 1e-5) * gam + bet
                hid = W1 @ z + b1
                hid = hid * torch.sigmoid(hid)
                self.put_f32(name + ".nullc", W2 @ hid + b2)
        sa = key + ".sa_block"
        Wqkv = torch.cat([g(sa + ".query.weight"), g(sa + ".key.weight"), g(sa + ".value.weight")], 0)
        bqkv = torch.cat([g(sa + ".query.bias"), g(sa + ".key.bias"), g(sa + ".value.bias")], 0)
        stored, bfold = self.lin_ln_fold(name + ".qkv", Wqkv, bqkv, g(sa + ".norm.weight"), g(sa + ".norm.bias"))
        Dm = Wqkv.shape[0] // 3
        es = expo_shift(stored, bfold, Dm, Dm // self.cfg["num_heads"])   # optional tensor: present only when static shifts are provably safe
        if es is not None:
            self.put_f32(name + ".qkv.eshift", es)
        self.put_f32(name + ".sa.g", g(sa + ".proj_out.norm.weight"))
        self.put_f32(name + ".sa.b", g(sa + ".proj_out.norm.bias"))
        self.lin(name + ".sa_out", sa + ".proj_out.out_layers.2")
        ff = key + ".ffn"
        self.lin(name + ".ffn1", ff + ".linear1")
        self.lin(name + ".ffn2", ff + ".linear2")
        self.put_f32(name + ".ffn.g", g(ff + ".proj_out.norm.weight"))
        self.put_f32(name + ".ffn.b", g(ff + ".proj_out.norm.bias"))
        self.lin(name + ".ffn_out", ff + ".proj_out.out_layers.2")

    def run(self, max_frames):
        cfg, g = self.cfg, self.g
        D, L = cfg["latent_dim"], cfg["num_layers"]
        half = D // 2
        # tr:52-54, computed with torch exactly as the reference does
        freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
        self.put_f32("freqs", freqs)
        self.mlp("aud.te", "time_embed.0", "time_embed.2")
        a = "encoder_aud"
        self.put_f32("aud.ss.w", torch.cat([g(a + ".sa_block.proj_out.emb_layers.1.weight"),
                                            g(a + ".ffn.proj_out.emb_layers.1.weight")], 0))
        self.put_f32("aud.ss.b", torch.cat([g(a + ".sa_block.proj_out.emb_layers.1.bias"),
                                            g(a + ".ffn.proj_out.emb_layers.1.bias")], 0))
        self.layer("aud.l0", a, None, None)
        for name, key, extra in (("exp", "encoder_exp", 0), ("ges", "encoder_ges", cfg["expression_dim"])):
            self.mlp(name + ".te", key + ".time_embed.0", key + ".time_embed.2")
            self.mlp(name + ".pid", key + ".pid_embed.0", key + ".pid_embed.2", pad_k=True)
            he = key + ".hubert_encoder"
            s = g(he + ".1.weight") / torch.sqrt(g(he + ".1.running_var") + 1e-5)
            w0 = g(he + ".0.weight") * s[:, None, None]           # [128, 1024, 3]
            self.put_f32(name + ".hub.w0", w0.permute(2, 1, 0))   # [3][Cin][Cout]
            self.put_f32(name + ".hub.b0", g(he + ".1.bias") - g(he + ".1.running_mean") * s)
            self.put_f32(name + ".hub.w3", g(he + ".3.weight").permute(2, 1, 0))
            pe = g(key + ".PE.pe")[0]
            assert max_frames <= pe.shape[0]
            self.put_f32(name + ".pe", pe[:max_frames])
            blocks = [key + f".temporal_decoder_blocks.{i}" for i in range(L)]
            self.put_w(name + ".ss.w", torch.cat(
                [g(b + m + ".proj_out.emb_layers.1.weight") for b in blocks for m in (".sa_block", ".ffn")], 0))
            self.put_f32(name + ".ss.b", torch.cat(
                [g(b + m + ".proj_out.emb_layers.1.bias") for b in blocks for m in (".sa_block", ".ffn")], 0))
            self.lin(name + ".joint", key + ".joint_embed")
            self.lin(name + ".audproj", key + ".audio_proj")
            self.lin(name + ".out", key + ".out")
            include_x = cfg.get("cond_projection", "mlp_includeX").endswith("includeX")   # tr:262-263: *_excludeX projects the conditioning only
            widths = ([D] if include_x else []) + [cfg["aud_latent_dim"], cfg["hubert_enc_dim"]] + ([extra] if extra else [])
            null = g(key + ".null_cond_emb")[0] if cfg["classifier_free"] else None
            for i, b in enumerate(blocks):
                self.layer(f"{name}.l{i}", b, widths, null)
        return self.out


def pack_state_dict(sd, cfg, precision, max_frames):
    """-> {packed_name: (cpu tensor, dtype_code)}"""
    return Packer(sd, cfg, precision).run(max_frames)
