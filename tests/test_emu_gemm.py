"""The tcgen05 GEMM kernel SOURCE (diffsheg_b200/csrc/gemm_tc.cuh, kernel and host-side launch code) executed on the CPU by
the thread-level emulator with host models of mbarrier / TMA / tcgen05 / TMEM (tests/emu/emu_tc_prims.h), checked against a
float64 evaluation of the same bf16 operands.  Covers what a GPU-less change can break: the producer / MMA / epilogue
pipeline protocol (stage and accumulator parities, early TMEM hand-off, CTA-pair barriers), tensor-map coordinates and
zero-filled ragged edges, the virtual-concat K walk, every epilogue variant the engine uses, the persistent tile walk.
The kernel is hardware-validated (tests/test_gpu_parity.py::test_op_linear on the B200); agreeing with it anchors the models."""
import ctypes

import numpy as np
import pytest
import torch

import emu

ACT_NONE, ACT_SILU, ACT_GELU, ACT_EXPO, ACT_LNMS = 0, 1, 2, 4, 5


class Args(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int32), ("N", ctypes.c_int32), ("nseg", ctypes.c_int32),
                ("seg_k", ctypes.c_int32 * 4), ("seg_ld", ctypes.c_int32 * 4), ("seg_ptr", ctypes.c_void_p * 4),
                ("w", ctypes.c_void_p), ("Kp", ctypes.c_int32),
                ("bias", ctypes.c_void_p), ("csum", ctypes.c_void_p), ("mu", ctypes.c_void_p), ("rstd", ctypes.c_void_p),
                ("act", ctypes.c_int32), ("res", ctypes.c_void_p), ("ldr", ctypes.c_int32), ("res_mod", ctypes.c_int32),
                ("res_f32", ctypes.c_int32), ("out", ctypes.c_void_p), ("ldo", ctypes.c_int32), ("out_f32", ctypes.c_int32),
                ("out2", ctypes.c_void_p), ("ps_out", ctypes.c_void_p), ("nullc", ctypes.c_void_p), ("n_uncond", ctypes.c_int32),
                ("ps_in", ctypes.c_void_p), ("cs_in", ctypes.c_void_p), ("ps_slots", ctypes.c_int32), ("ps_P", ctypes.c_int32),
                ("num_sms", ctypes.c_int32), ("bn_force", ctypes.c_int32), ("cg_force", ctypes.c_int32),
                ("eshift", ctypes.c_void_p), ("expo_cols", ctypes.c_int32),
                ("lnms_g", ctypes.c_void_p), ("lnms_b", ctypes.c_void_p), ("lnms_ss", ctypes.c_void_p),
                ("lnms_ld", ctypes.c_int32), ("lnms_B", ctypes.c_int32), ("lnms_T", ctypes.c_int32), ("rev", ctypes.c_int32)]


def bf16_bits(t):
    return np.ascontiguousarray(t.bfloat16().view(torch.int16).numpy())


def bits_to_f64(a):
    return torch.from_numpy(np.ascontiguousarray(a)).view(torch.bfloat16).double()


def ptr(a):
    return a.ctypes.data if a is not None else None


def r64(k):
    return (k + 63) // 64 * 64


def run_gemm(M, N, seg_ks, *, ln=False, act=ACT_NONE, res=None, out_f32=False, dup=False, stats_out=False, n_uncond=0,
             ps_in=False, num_sms=4, bn=0, cg=0, seed=0, lib=None, expo_cols=0, expo_q_cols=0, lnms_T=0, lnms_B=0, rev=0):
    """Builds operands like the engine does (K laid out per segment padded to 64), runs the emulated kernel, returns
    (got, want, extras)."""
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g)
    Kp = sum(r64(k) for k in seg_ks)
    segs, keep = [], []
    Wfull = torch.zeros(N, Kp)
    off = 0
    A_cat = []
    for k in seg_ks:
        ld = r64(k) + 64                                   # a wider buffer: exercises the leading dimension
        a = torch.zeros(M, ld)
        a[:, :k] = rnd(M, k)
        a[:, k:] = 7.0                                     # garbage beyond the segment width must never be read as data...
        Wfull[:, off:off + k] = rnd(N, k) / np.sqrt(sum(seg_ks))
        ab = bf16_bits(a)
        keep.append(ab)
        segs.append((ab, ld, k))
        A_cat.append(bits_to_f64(ab)[:, :k])
        off += r64(k)                                      # ...W is zero there anyway only up to the 64-padding
    # columns k..r64(k) of a segment ARE read (TMA box = 64 columns) unless clipped by the tensor map's width k -> zero fill
    Wb = bf16_bits(Wfull)
    Wd = bits_to_f64(Wb)
    Wcols = torch.cat([Wd[:, o:o + k] for o, k in zip(np.cumsum([0] + [r64(k) for k in seg_ks[:-1]]), seg_ks)], 1)
    acc = torch.cat(A_cat, 1) @ Wcols.T
    bias = rnd(N).float().numpy()
    a = Args()
    a.M, a.N, a.nseg, a.Kp = M, N, len(seg_ks), Kp
    a.rev = rev
    for i, (ab, ld, k) in enumerate(segs):
        a.seg_k[i], a.seg_ld[i], a.seg_ptr[i] = k, ld, ptr(ab)
    a.w, a.bias, a.act = ptr(Wb), ptr(bias), act
    want = acc.clone()
    extras = {}
    if ln:
        csum = Wcols.sum(1).float().numpy()
        mu, rstd = rnd(M).float().numpy() * 0.1, (rnd(M).abs() + 0.5).float().numpy()
        keep += [csum, mu, rstd]
        a.csum = ptr(csum)
        if ps_in:   # statistics rebuilt from 8 (sum, sumsq) partials per row, plus a conditioning partial
            P = 512 + 37
            tot_s = torch.from_numpy(mu.astype(np.float64)) * P
            var = 1.0 / torch.from_numpy(rstd.astype(np.float64)) ** 2 - 1e-5
            tot_q = (var + torch.from_numpy(mu.astype(np.float64)) ** 2) * P
            w8 = torch.softmax(rnd(M, 9), 1).double()
            parts = torch.stack([tot_s[:, None] * w8, tot_q[:, None] * w8], -1).float()     # [M, 9, 2]
            ps = np.ascontiguousarray(parts[:, :8].numpy())
            cs = np.ascontiguousarray(parts[:, 8].numpy())
            keep += [ps, cs]
            a.ps_in, a.cs_in, a.ps_slots, a.ps_P = ptr(ps), ptr(cs), 8, P
            sx = parts[..., 0].double().sum(1)
            sq = parts[..., 1].double().sum(1)
            mean = sx / P
            rs = 1.0 / torch.sqrt(torch.clamp(sq / P - mean * mean, min=0) + 1e-5)
            want = rs[:, None] * (acc - mean[:, None] * torch.from_numpy(csum.astype(np.float64))[None])
        else:
            a.mu, a.rstd = ptr(mu), ptr(rstd)
            want = torch.from_numpy(rstd.astype(np.float64))[:, None] * (acc - torch.from_numpy(mu.astype(np.float64))[:, None] * torch.from_numpy(csum.astype(np.float64))[None])
    want = want + torch.from_numpy(bias.astype(np.float64))
    if n_uncond:
        nullc = rnd(N).float().numpy()
        keep.append(nullc)
        a.nullc, a.n_uncond = ptr(nullc), n_uncond
        want[:n_uncond] += torch.from_numpy(nullc.astype(np.float64))
    extras["raw"] = want.clone()
    if act == ACT_EXPO:    # leading expo_cols columns: exp(v - eshift[n]) with static per-column shifts
        eshift = 3.0 * rnd(expo_cols)
        if expo_q_cols:    # a row softmax (Q) tolerates one shift per 64-column head only; the column softmax (K) any per-column shift
            eshift[:expo_q_cols] = eshift[:expo_q_cols:64].repeat_interleave(64)
        eshift = eshift.float().numpy()
        keep.append(eshift)
        a.eshift, a.expo_cols = ptr(eshift), expo_cols
        extras["eshift"] = torch.from_numpy(eshift.astype(np.float64))
        want = torch.cat([torch.exp(want[:, :expo_cols] - extras["eshift"]), want[:, expo_cols:]], 1)
    if act == ACT_LNMS:    # StylizationBlock prologue over the full row: SiLU(LN(v) * (1 + scale) + shift), sample = row // T
        lg, lb = (1 + 0.2 * rnd(N)).float().numpy(), (0.2 * rnd(N)).float().numpy()
        ss = np.ascontiguousarray((0.5 * rnd(lnms_B, 2 * N + 4)).float().numpy())        # ld = 2 N + 4: exercises the row stride
        keep += [lg, lb, ss]
        a.lnms_g, a.lnms_b, a.lnms_ss, a.lnms_ld, a.lnms_B, a.lnms_T = ptr(lg), ptr(lb), ptr(ss), 2 * N + 4, lnms_B, lnms_T
        idx = (torch.arange(M) // lnms_T) % lnms_B
        sst = torch.from_numpy(ss.astype(np.float64))[idx]
        yn = torch.nn.functional.layer_norm(want, (N,), torch.from_numpy(lg.astype(np.float64)), torch.from_numpy(lb.astype(np.float64)), 1e-5)
        want = torch.nn.functional.silu(yn * (1 + sst[:, :N]) + sst[:, N:2 * N])
    if act == ACT_SILU:
        want = torch.nn.functional.silu(want)
    elif act == ACT_GELU:
        want = torch.nn.functional.gelu(want)
    if res == "bf16":
        r = bf16_bits(rnd(M, N))
        keep.append(r)
        a.res, a.ldr = ptr(r), N
        want = want + bits_to_f64(r)
    elif res == "f32mod":
        mod = 88
        r = np.ascontiguousarray(rnd(mod, N).float().numpy())
        keep.append(r)
        a.res, a.ldr, a.res_mod, a.res_f32 = ptr(r), N, mod, 1
        want = want + torch.from_numpy(r.astype(np.float64))[torch.arange(M) % mod]
    if out_f32:
        out = np.full((M, N), np.nan, np.float32)
    else:
        out = np.full((M, N), 0x7FC0, np.int16)            # bf16 NaN: unwritten outputs are caught
    a.out, a.ldo, a.out_f32 = ptr(out), N, int(out_f32)
    out2 = None
    if dup:
        out2 = np.full((M, N), 0x7FC0, np.int16)
        a.out2 = ptr(out2)
    if stats_out:
        ps_o = np.full((M, N // 64, 2), np.nan, np.float32)
        a.ps_out = ptr(ps_o)
        extras["ps_out"] = ps_o
    a.num_sms, a.bn_force, a.cg_force = num_sms, bn, cg
    L = lib or emu.gemm_lib()
    rc = L.emu_gemm_tc(ctypes.byref(a))
    assert rc == 0, L.emu_gemm_last_error().decode()
    got = torch.from_numpy(out).double() if out_f32 else bits_to_f64(out)
    if dup:
        assert np.array_equal(out, out2)
    extras["want_f64"] = want
    extras["out_bits"] = None if out_f32 else out
    return got, want, extras


def check(got, want, out_f32=False, tanh_gelu=False):
    assert torch.isfinite(got).all(), "an output element was never written"
    err = float((got - want).abs().max() / want.abs().max())
    # fp32 accumulation of bf16 products + (bf16 output rounding 2^-9 | tanh-form GELU 5e-4)
    assert err < (2e-5 if out_f32 else (8e-3 if not tanh_gelu else 9e-3)), err


def test_single_cta_bias_only_smoke():
    got, want, _ = run_gemm(256, 256, [128], cg=1)
    check(got, want)


# (M, N, segment widths, options): every epilogue variant launch_gemm_tc dispatches for the engine, single CTAs and CTA pairs,
# ragged M / N / K, persistent walks with more tiles than CTAs (accumulator and stage parities wrap several times)
CASES = [
    # single-CTA kernels (BN = 256 and 128)
    dict(M=300, N=512, ks=[192], cg=1, num_sms=2),                                   # 3 x 2 tiles on 2 CTAs
    dict(M=129, N=128, ks=[64], cg=1, res="bf16"),                                   # BN = 128, TMA residual boxes
    dict(M=77, N=103, ks=[129], cg=1, out_f32=True),                                 # ragged everything, fp32 scalar stores
    dict(M=200, N=256, ks=[256], cg=1, ln=True, act=ACT_SILU),
    dict(M=140, N=512, ks=[232], cg=1, res="f32mod", dup=True, num_sms=2),           # joint_embed + PE, dual store (CFG halves)
    dict(M=260, N=384, ks=[128], cg=1, act=ACT_GELU),                                # N % 256 != 0 -> BN = 128
    # CTA pairs (cta_group::2), K = 512 class: 4 stages, wide boxes
    dict(M=600, N=512, ks=[512], cg=2, num_sms=4, res="bf16", stats_out=True, n_uncond=300),   # sa_out / ffn_out (+ LN partials, nullc)
    dict(M=520, N=768, ks=[512], cg=2, num_sms=4, ln=True),                           # qkv-like, 3 n-tiles, ragged last pair
    dict(M=512, N=512, ks=[512], cg=2, num_sms=2, ln=True, ps_in=True),               # statistics rebuilt from producer partials
    dict(M=300, N=1024, ks=[512], cg=2, num_sms=4, act=ACT_GELU),                     # ffn1
    # CTA pairs, K >= 768 class: deep ring, narrow (2 KB) boxes without residual, wide with
    dict(M=520, N=512, ks=[1024], cg=2, num_sms=2),                                   # ffn2: 6 stages, narrow boxes
    dict(M=300, N=512, ks=[1024], cg=2, num_sms=2, res="bf16", stats_out=True, n_uncond=100),   # feat2: 5 stages, wide boxes
    dict(M=270, N=1024, ks=[512, 128, 128, 103], cg=2, num_sms=4, ln=True, act=ACT_SILU),        # feat1: virtual concat, 999 -> 1024
    dict(M=256, N=256, ks=[768], cg=2, num_sms=2, dup=True),
    # the feat_proj GEMMs of the other cond_projection variants (engine.cu: feat_proj_variant; tr:262-263,281-289): operand lists
    # without the 512-wide hidden-state segment, the no-LayerNorm multi-segment Linear with and without the bf16 residual
    dict(M=270, N=1024, ks=[256, 128, 103], cg=2, num_sms=4, ln=True, act=ACT_SILU),    # mlp_excludeX feat1, gesture net: K = 512 class pairs
    dict(M=300, N=1024, ks=[256, 128], cg=1, num_sms=3, ln=True, act=ACT_SILU),         # mlp_excludeX feat1, expression net: K = 384, single CTAs
    dict(M=300, N=512, ks=[512, 256, 128, 103], cg=2, num_sms=2, res="bf16"),           # linear_includeX + cond_residual, gesture net (long K, wide boxes)
    dict(M=300, N=512, ks=[512, 256, 128], cg=2, num_sms=2),                            # linear_includeX without residual, expression net (narrow boxes)
    dict(M=300, N=512, ks=[256, 128, 103], cg=2, num_sms=2, res="bf16"),                # linear_excludeX, gesture net (in place in the engine)
    dict(M=200, N=512, ks=[256, 128], cg=1, num_sms=2, res="bf16"),                     # linear_excludeX, expression net
    dict(M=40, N=512, ks=[256, 128], cg=1, bn=128, num_sms=8, res="bf16"),              # ... in the single-clip regime (128-wide tiles)
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "M{M}-N{N}-K{k}-cg{cg}{x}".format(
    M=c["M"], N=c["N"], k="+".join(map(str, c["ks"])), cg=c["cg"],
    x="".join("-" + k for k in ("ln", "res", "out_f32", "dup", "stats_out", "ps_in") if c.get(k)) + ("-act%d" % c["act"] if c.get("act") else "")))
def test_gemm_kernel_source_on_emulator(case):
    c = dict(case)
    M, N, ks = c.pop("M"), c.pop("N"), c.pop("ks")
    got, want, extras = run_gemm(M, N, ks, **c)
    check(got, want, out_f32=c.get("out_f32", False), tanh_gelu=c.get("act") == ACT_GELU)
    if "ps_out" in extras:   # per-row, per-64-column (sum, sum of squares) of the values the epilogue stored (before bf16 rounding)
        w = extras["want_f64"].reshape(M, N // 64, 64)
        ps = torch.from_numpy(extras["ps_out"]).double()
        assert torch.isfinite(ps).all()
        assert float((ps[..., 0] - w.sum(-1)).abs().max()) < 2e-3 * float(w.abs().sum(-1).max())
        assert float((ps[..., 1] - (w * w).sum(-1)).abs().max()) < 2e-3 * float((w * w).sum(-1).max())


ADVERSARIAL = [dict(M=600, N=512, ks=[512], cg=2, num_sms=2, res="bf16", stats_out=True, n_uncond=300),
               dict(M=520, N=512, ks=[1024], cg=2, num_sms=2),
               dict(M=300, N=512, ks=[192], cg=1, num_sms=2, ln=True)]


@pytest.mark.parametrize("slow", ["EMU_DELAY_TMEM_LD", "EMU_DELAY_TMA", "EMU_DELAY_MMA"])
@pytest.mark.parametrize("case", ADVERSARIAL, ids=["pair-k512-res", "pair-k1024-narrow", "cta1-ln"])
def test_gemm_pipeline_protocol_under_adversarial_timing(case, slow, monkeypatch):
    """The round-robin emulator makes producer, MMA issuer and epilogue equally fast.  Slowing ONE role down (it sits out 40
    scheduler passes before each of its operations) forces the interleavings in which a protocol bug shows -- e.g. handing a
    TMEM accumulator back before its last tcgen05.ld corrupts the output under EMU_DELAY_TMEM_LD (checked by mutation when
    the knob was added).  The real kernel must be insensitive to all of them."""
    monkeypatch.setenv(slow, "40")
    c = dict(case)
    M, N, ks = c.pop("M"), c.pop("N"), c.pop("ks")
    got, want, _ = run_gemm(M, N, ks, **c)
    check(got, want)


@pytest.mark.parametrize("M,N,ec,cg,num_sms,ps", [(300, 768, 512, 1, 2, False), (520, 1536, 1024, 2, 2, True), (300, 512, 512, 2, 2, False),
                                                  (140, 384, 128, 1, 1, True)])
def test_exponential_epilogue(M, N, ec, cg, num_sms, ps):
    """ACT_EXPO (default; DSHEG_EXPO=0 disables): the LN-fold QKV projection writes exp(v - eshift[n]) for its leading columns (Q and K:
    softmax numerators with static shifts, transformer.py:122-123 moved into the producing GEMM) and the rest plain."""
    got, want, ex = run_gemm(M, N, [512], ln=True, act=ACT_EXPO, expo_cols=ec, cg=cg, num_sms=num_sms, ps_in=ps, seed=7)
    assert torch.isfinite(got).all()
    rel = ((got[:, :ec] - want[:, :ec]) / want[:, :ec]).abs().max()
    assert float(rel) < 6e-3, float(rel)        # bf16 rounding (2^-9) + ex2.approx of an fp32 argument, RELATIVE: exponentials span decades
    if ec < N:
        assert float((got[:, ec:] - want[:, ec:]).abs().max() / want[:, ec:].abs().max()) < 8e-3


@pytest.mark.parametrize("M,K,T,B,num_sms,cg", [(600, 1024, 88, 3, 2, 2), (1100, 768, 34, 2, 4, 2), (256, 1024, 7, 40, 2, 2), (700, 1024, 88, 2, 6, 2),
                                                  # single CTAs (small batches): 128 rows x 512 columns of TMEM per CTA, any K
                                                  (176, 1024, 88, 1, 4, 1), (34, 1024, 34, 1, 2, 1), (300, 512, 34, 3, 1, 1), (520, 1024, 88, 2, 2, 1)])
def test_full_row_layernorm_modulate_silu_epilogue(M, K, T, B, num_sms, cg):
    """ACT_LNMS (default for rows >= 4096; DSHEG_FUSE_LNMS=0 disables): ffn.linear2 + the StylizationBlock prologue (transformer.py:178-181 + :92-96) in ONE
    kernel -- a CTA pair owns both 256-column tiles of its row panel (one per TMEM accumulator stage), so LayerNorm statistics
    span the full 512-wide row; persistent walks with more panels than pairs wrap the stage / accumulator parities."""
    got, want, _ = run_gemm(M, 512, [K], act=ACT_LNMS, lnms_T=T, lnms_B=B, cg=cg, num_sms=num_sms, seed=11)
    check(got, want)


@pytest.mark.parametrize("slow,cg", [("EMU_DELAY_TMEM_LD", 2), ("EMU_DELAY_TMA", 2), ("EMU_DELAY_MMA", 2), ("EMU_DELAY_TMEM_LD", 1), ("EMU_DELAY_MMA", 1)])
def test_full_row_layernorm_epilogue_under_adversarial_timing(slow, cg, monkeypatch):
    monkeypatch.setenv(slow, "40")
    got, want, _ = run_gemm(520 if cg == 2 else 270, 512, [768], act=ACT_LNMS, lnms_T=88, lnms_B=5, cg=cg, num_sms=2, seed=12)
    check(got, want)


@pytest.mark.parametrize("sched", ["shuffle"])     # (reverse is subsumed: shuffle re-draws the order every pass and lets warps sit passes out)
@pytest.mark.parametrize("kw", [dict(M=520, N=512, ks=[1024], act=ACT_LNMS, lnms_T=88, lnms_B=3, cg=2, num_sms=2),
                                dict(M=300, N=512, ks=[768], act=ACT_LNMS, lnms_T=34, lnms_B=2, cg=1, num_sms=2),
                                dict(M=520, N=768, ks=[512], ln=True, act=ACT_EXPO, expo_cols=512, cg=2, num_sms=2, ps_in=True),
                                dict(M=600, N=512, ks=[512], cg=2, num_sms=2, res="bf16", stats_out=True, n_uncond=300),
                                dict(M=300, N=1024, ks=[512], cg=2, num_sms=4, act=ACT_GELU),
                                dict(M=270, N=1024, ks=[512, 128, 128, 103], cg=2, num_sms=4, ln=True, act=ACT_SILU),
                                dict(M=520, N=512, ks=[1024], cg=2, num_sms=2),
                                dict(M=140, N=512, ks=[232], cg=1, res="f32mod", dup=True, num_sms=2),
                                dict(M=260, N=384, ks=[128], cg=1, act=ACT_GELU)],
                         ids=["lnms-pair", "lnms-cta1", "expo-pair", "residual-pair", "gelu-pair", "feat1-pair", "longk-narrow-pair",
                              "joint-cta1", "bn128-cta1"])
def test_gemm_is_independent_of_the_thread_schedule(kw, sched, monkeypatch):
    """Emulator stand-in for racecheck (see tests/emu/emu_cuda.h EMU_SCHED): reversed and per-pass shuffled thread orders must give
    bit-identical outputs -- the epilogue's shared-memory exchanges (statistics partials, staged vectors, TMA boxes) included."""
    k = dict(kw)
    M, N, ks = k.pop("M"), k.pop("N"), k.pop("ks")
    _, _, ex0 = run_gemm(M, N, ks, seed=31, **k)
    base = ex0["out_bits"].copy()
    monkeypatch.setenv("EMU_SCHED", sched)
    _, _, ex1 = run_gemm(M, N, ks, seed=31, **k)
    assert np.array_equal(base, ex1["out_bits"])


@pytest.mark.parametrize("kw", [dict(M=520, N=512, ks=[512], cg=2, num_sms=2, res="bf16", stats_out=True),
                                dict(M=700, N=1536, ks=[512], cg=2, num_sms=4, ln=True, ps_in=True),
                                dict(M=520, N=512, ks=[1024], act=ACT_LNMS, lnms_T=88, lnms_B=3, cg=2, num_sms=2),
                                dict(M=300, N=512, ks=[768], cg=1, num_sms=2, res="bf16")])
def test_reverse_row_panel_walk_gives_the_same_result(kw):
    """GemmDesc::rev (DSHEG_ZIGZAG experiment): the persistent tile walk takes the row panels last-to-first; every row still gets
    exactly its own tile (ragged last panel included), so the output is bit-identical to the forward walk."""
    kw = dict(kw)
    M, N, ks = kw.pop("M"), kw.pop("N"), kw.pop("ks")
    fwd, want, _ = run_gemm(M, N, ks, seed=21, **kw)
    rev, _, _ = run_gemm(M, N, ks, seed=21, rev=1, **kw)
    assert torch.equal(fwd, rev)
    assert float((fwd - want).abs().max() / want.abs().max()) < 2e-2


# ------------------------------------------------------------------------------------------------------------------------------
# gemm_tf32.cuh: the GEMM engine of the tf32 precision mode (fp32 activations, TF32 operands, fp32 accumulate / epilogue / output)
# ------------------------------------------------------------------------------------------------------------------------------
class Tf32Args(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int32), ("N", ctypes.c_int32), ("nseg", ctypes.c_int32),
                ("seg_k", ctypes.c_int32 * 4), ("seg_ld", ctypes.c_int32 * 4), ("seg_ptr", ctypes.c_void_p * 4),
                ("w", ctypes.c_void_p), ("Kp", ctypes.c_int32),
                ("bias", ctypes.c_void_p), ("csum", ctypes.c_void_p), ("mu", ctypes.c_void_p), ("rstd", ctypes.c_void_p),
                ("act", ctypes.c_int32), ("res", ctypes.c_void_p), ("ldr", ctypes.c_int32), ("res_mod", ctypes.c_int32),
                ("out", ctypes.c_void_p), ("ldo", ctypes.c_int32), ("out2", ctypes.c_void_p), ("cg_force", ctypes.c_int32), ("num_sms", ctypes.c_int32)]


def _to_tf32(t):
    """fp32 -> TF32 (10-bit mantissa, round to nearest even) as the TFLOAT32 tensor map delivers it; returned as float64."""
    u = t.float().contiguous().view(torch.int32)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & ~0x1FFF
    return u.view(torch.float32).double()


def run_gemm_tf32(M, N, seg_ks, *, ln=False, act=ACT_NONE, res=None, dup=False, cg=0, seed=0, num_sms=2):
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g)
    Kp = sum(r64(k) for k in seg_ks)
    W = torch.zeros(N, Kp)
    keep, A_cat, off = [], [], 0
    a = Tf32Args()
    a.M, a.N, a.nseg, a.Kp, a.cg_force, a.act, a.num_sms = M, N, len(seg_ks), Kp, cg, act, num_sms
    for i, k in enumerate(seg_ks):
        ld = r64(k) + 12                                   # a wider buffer: exercises the leading dimension (multiple of 4)
        seg = torch.full((M, ld), 7.0)                     # garbage beyond the segment width must never reach the product
        seg[:, :k] = rnd(M, k)
        W[:, off:off + k] = rnd(N, k) / np.sqrt(sum(seg_ks))
        buf = np.ascontiguousarray(seg.numpy())
        keep.append(buf)
        a.seg_k[i], a.seg_ld[i], a.seg_ptr[i] = k, ld, buf.ctypes.data
        A_cat.append((_to_tf32(seg[:, :k]), _to_tf32(W[:, off:off + k])))
        off += r64(k)
    acc = sum(x @ w.T for x, w in A_cat)
    Wb = np.ascontiguousarray(W.numpy())
    bias = rnd(N).numpy().copy()
    a.w, a.bias = Wb.ctypes.data, bias.ctypes.data
    want = acc.clone()
    if ln:                                                 # LayerNorm fold: out = rstd * (acc - mu * csum) + bias
        csum, mu, rstd = rnd(N).numpy().copy(), rnd(M).numpy().copy(), (0.5 + torch.rand(M, generator=g)).numpy().copy()
        keep += [csum, mu, rstd]
        a.csum, a.mu, a.rstd = csum.ctypes.data, mu.ctypes.data, rstd.ctypes.data
        want = torch.from_numpy(rstd).double()[:, None] * (acc - torch.from_numpy(mu).double()[:, None] * torch.from_numpy(csum).double()[None, :])
    want = want + torch.from_numpy(bias).double()[None, :]
    if act == ACT_GELU:
        want = torch.nn.functional.gelu(want)
    elif act == ACT_SILU:
        want = torch.nn.functional.silu(want)
    if res == "f32":
        ldr = N + 4
        r = np.ascontiguousarray(rnd(M, ldr).numpy())
        keep.append(r)
        a.res, a.ldr, a.res_mod = r.ctypes.data, ldr, 0
        want = want + torch.from_numpy(r[:, :N]).double()
    elif res == "f32mod":
        rows = 34
        r = np.ascontiguousarray(rnd(rows, N).numpy())
        keep.append(r)
        a.res, a.ldr, a.res_mod = r.ctypes.data, N, rows
        want = want + torch.from_numpy(r).double()[torch.arange(M) % rows]
    ldo = N + 8
    out = np.full((M, ldo), np.float32(-777.0))
    out2 = np.full((M, ldo), np.float32(-777.0)) if dup else None
    a.out, a.ldo = out.ctypes.data, ldo
    a.out2 = out2.ctypes.data if dup else None
    L = emu.gemm_tf32_lib()
    rc = L.emu_gemm_tf32(ctypes.byref(a))
    assert rc == 0, L.emu_gemm_tf32_last_error().decode()
    assert (out[:, N:] == -777.0).all(), "wrote beyond the N columns of a row"
    got = torch.from_numpy(out[:, :N]).double()
    if dup:
        assert np.array_equal(out, out2)
    return got, want


@pytest.mark.parametrize("kw", [
    dict(M=300, N=512, ks=[512], cg=2),                                   # CTA pair, ragged last pair of row panels
    dict(M=700, N=1536, ks=[512], cg=2, ln=True),                         # qkv: LayerNorm fold
    dict(M=260, N=1024, ks=[512], cg=2, act=ACT_GELU),                    # ffn1: exact-erf GELU
    dict(M=260, N=512, ks=[1024], cg=2, res="f32"),                       # ffn_out / sa_out: fp32 residual (prefetched per chunk)
    dict(M=140, N=512, ks=[232], cg=1, res="f32mod", dup=True),           # joint_embed + PE, dual store, single CTAs, ragged K
    dict(M=200, N=512, ks=[512, 256, 128, 103], cg=2, ln=True, act=ACT_SILU),   # feat_proj: 4-segment virtual concat, ragged last segment
    dict(M=129, N=200, ks=[96], cg=1),                                    # ragged N (element-wise store path), one k-block of padding
])
def test_gemm_tf32_kernel_source_on_emulator(kw):
    """The tf32 engine's kernel AND host launch code on the emulator (it had hardware coverage only): TFLOAT32 tensor maps, kind::tf32
    MMAs with K = 8, CTA pairs, the 8-warp epilogue with its turn through shared memory, every epilogue family the tf32 mode uses."""
    kw = dict(kw)
    M, N, ks = kw.pop("M"), kw.pop("N"), kw.pop("ks")
    got, want = run_gemm_tf32(M, N, ks, seed=5, **kw)
    assert torch.isfinite(got).all()
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 2e-5, err        # operands are exactly the TF32 values the reference uses; only fp32 accumulation order differs


@pytest.mark.parametrize("sched", ["reverse", "shuffle"])
def test_gemm_tf32_is_independent_of_the_thread_schedule(sched, monkeypatch):
    base, _ = run_gemm_tf32(260, 512, [512], cg=2, res="f32", seed=6)
    monkeypatch.setenv("EMU_SCHED", sched)
    got, _ = run_gemm_tf32(260, 512, [512], cg=2, res="f32", seed=6)
    assert torch.equal(base, got)


@pytest.mark.parametrize("env", [{"EMU_DELAY_TMEM_LD": "40"}, {"EMU_DELAY_MMA": "40"}, {"EMU_DELAY_TMA": "40"},
                                 {"EMU_SCHED": "shuffle", "EMU_SCHED_SEED": "7", "EMU_DELAY_TMEM_LD": "15"}])
@pytest.mark.parametrize("cg,num_sms", [(2, 2), (1, 1)])
def test_gemm_tf32_persistent_pipeline_under_adversarial_timing(env, cg, num_sms, monkeypatch):
    """The persistent form: ONE CTA (pair) walks 12 / 24 tiles, so ring slots and both accumulator stages wrap many times.  A slow
    epilogue (late tcgen05.ld), a slow MMA issuer or late TMA arrivals must not change a bit: an accumulator stage overwritten before
    its last read, a ring slot refilled under a pending MMA or per-column vectors restaged too early would."""
    kw = dict(ln=True, act=ACT_GELU, cg=cg, num_sms=num_sms, seed=9)
    base, want = run_gemm_tf32(390, 1024, [256], **kw)
    assert float((base - want).abs().max() / want.abs().max()) < 2e-5
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got, _ = run_gemm_tf32(390, 1024, [256], **kw)
    assert torch.equal(base, got)
