"""Torch (CPU, fp64) emulation of the ENGINE'S dataflow over the PACKED tensors -- test infrastructure.

It follows diffsheg_b200/csrc/engine.cu launch by launch (LayerNorm folds, virtual-concat K layout,
null-row constants, scale/shift table layout, CFG row order) so that pack.py and the graph design
can be checked against the oracle on a machine without a GPU.
"""
import torch
import torch.nn.functional as F


def _r64(k):
    return (k + 63) // 64 * 64


class Emu:
    def __init__(self, packed, cfg):
        self.p = {k: v[0].to(torch.float64) for k, v in packed.items()}
        self.cfg = cfg

    def gemm(self, name, segs, mu=None, rstd=None, act=None, res=None):
        W, b = self.p[name + ".w"], self.p[name + ".b"]
        cols = []
        for s in segs:  # virtual concat: every segment padded to 64 along K
            pad = _r64(s.shape[-1]) - s.shape[-1]
            cols.append(F.pad(s, (0, pad)) if pad else s)
        A = torch.cat(cols, -1)
        assert A.shape[-1] == W.shape[1], (name, A.shape, W.shape)
        v = A @ W.T
        if mu is not None:
            v = rstd[:, None] * (v - mu[:, None] * self.p[name + ".csum"][None])
        v = v + b
        if act == "silu":
            v = F.silu(v)
        elif act == "gelu":
            v = F.gelu(v)
        return v if res is None else v + res

    @staticmethod
    def stats(rows):
        mu = rows.mean(-1)
        var = ((rows - mu[:, None]) ** 2).mean(-1)
        return mu, torch.rsqrt(var + 1e-5)

    def lms(self, y, g, b, scale, shift):
        return F.silu(F.layer_norm(y, (y.shape[-1],), g, b, 1e-5) * (1 + scale) + shift)

    def layer(self, name, h, n_uncond, extra, ss, T, H):
        """h [rows, D]; ss [n_samples_mod, 4D] for this layer; returns new h."""
        p, D = self.p, h.shape[-1]
        rows = h.shape[0]
        if name + ".feat1.w" in p:
            hu, hc = h[:n_uncond], h[n_uncond:]
            if n_uncond:
                hu = hu + p[name + ".nullc"]
            mu, rstd = self.stats(torch.cat([hc] + extra, -1))
            f1 = self.gemm(name + ".feat1", [hc] + extra, mu, rstd, "silu")
            hc = self.gemm(name + ".feat2", [f1], res=hc)
            h = torch.cat([hu, hc], 0)
        mu, rstd = self.stats(h)
        qkv = self.gemm(name + ".qkv", [h], mu, rstd)
        n = rows // T
        q, k, v = qkv.view(n, T, 3 * D).split(D, -1)
        q = torch.softmax(q.reshape(n, T, H, -1), -1)
        k = torch.softmax(k.reshape(n, T, H, -1), 1)
        att = torch.einsum("bnhd,bnhl->bhdl", k, v.reshape(n, T, H, -1))
        y = torch.einsum("bnhd,bhdl->bnhl", q, att).reshape(n, T, D)
        nb = ss.shape[0]
        idx = torch.arange(n) % nb
        sc = ss[idx]  # [n, 4D]
        z = self.lms(y, p[name + ".sa.g"], p[name + ".sa.b"], sc[:, None, :D], sc[:, None, D:2 * D]).reshape(rows, D)
        h = self.gemm(name + ".sa_out", [z], res=h)
        f = self.gemm(name + ".ffn1", [h], act="gelu")
        y2 = self.gemm(name + ".ffn2", [f]).view(n, T, D)
        z2 = self.lms(y2, p[name + ".ffn.g"], p[name + ".ffn.b"], sc[:, None, 2 * D:3 * D], sc[:, None, 3 * D:]).reshape(rows, D)
        return self.gemm(name + ".ffn_out", [z2], res=h)

    def mlp(self, name, v, pad=False):
        if pad:
            v = F.pad(v, (0, _r64(v.shape[-1]) - v.shape[-1]))
        hid = F.silu(v @ self.p[name + ".w0"].T + self.p[name + ".b0"])
        return hid @ self.p[name + ".w2"].T + self.p[name + ".b2"]

    def hubconv(self, name, hub):
        p = self.p
        B, T, C = hub.shape

        def conv(x, w, b):  # w [3][Cin][Cout]
            xp = F.pad(x, (0, 0, 1, 1))
            out = sum(xp[:, dk:dk + T] @ w[dk] for dk in range(3))
            return out if b is None else out + b
        mid = F.gelu(conv(hub, p[name + ".hub.w0"], p[name + ".hub.b0"]))
        return conv(mid, p[name + ".hub.w3"], None)

    def denoise(self, x, t_orig, a, b, cond_scale, mel, hubert, pid):
        cfg, p = self.cfg, self.p
        x, mel, hubert, pid = (v.to(torch.float64) for v in (x, mel, hubert, pid))
        B, T, _ = x.shape
        D, L, H, A = cfg["latent_dim"], cfg["num_layers"], cfg["num_heads"], cfg["audio_dim"]
        R1 = B * T
        two = cfg["classifier_free"] and cond_scale != 1.0
        arg = float(t_orig) * p["freqs"]
        sin = torch.cat([torch.cos(arg), torch.sin(arg)])
        temb = {n: self.mlp(n + ".te", sin) for n in ("aud", "exp", "ges")}
        ssa = (F.silu(temb["aud"]) @ p["aud.ss.w"].T + p["aud.ss.b"])[None]  # [1, 4A]
        a0 = (2 * mel).reshape(R1, A)
        a2 = self.layer("aud.l0", a0, 0, [], ssa, T, H)
        aud256 = torch.cat([mel.reshape(R1, A), a2], -1)
        eps = torch.zeros_like(x)
        expr = None
        for n, feats, off in (("exp", cfg["expression_dim"], cfg["dim_pose"]), ("ges", cfg["dim_pose"], 0)):
            embs = F.silu(temb[n][None] + self.mlp(n + ".pid", pid, pad=True))
            ss = embs @ p[n + ".ss.w"].T + p[n + ".ss.b"]  # [B, L*4D]
            xf = self.gemm(n + ".audproj", [aud256])
            hub = self.hubconv(n, hubert).reshape(R1, -1)
            xin = x[..., off:off + feats].reshape(R1, feats)
            pe = p[n + ".pe"][:T].repeat(B, 1)
            h = self.gemm(n + ".joint", [xin], res=pe)
            if two:
                h = torch.cat([h, h], 0)
            extra = [xf, hub] + ([expr] if n == "ges" else [])
            for l in range(L):
                h = self.layer(f"{n}.l{l}", h, R1 if two else 0, extra, ss[:, l * 4 * D:(l + 1) * 4 * D], T, H)
            o = self.gemm(n + ".out", [h])
            e = o[:R1] + cond_scale * (o[R1:] - o[:R1]) if two else o
            eps[..., off:off + feats] = e.view(B, T, feats)
            if n == "exp":
                expr = a * xin - b * e
        return eps
