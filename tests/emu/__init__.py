"""Host emulation of the SIMT / mma.sync kernel SOURCES (test infrastructure only, never imported by the product).

``lib()`` compiles tests/emu/emu_kernels.cpp -- which #includes the kernel headers of diffsheg_b200/csrc with -DDSHEG_EMU --
with g++ into tests/emu/_build/ and loads it through ctypes."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "diffsheg_b200", "csrc")
_lib = None
_libs = {}


def _stale(so):
    if not os.path.exists(so):
        return True
    m = os.path.getmtime(so)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".h", ".cpp"))]
    srcs += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(s) > m for s in srcs)


def _load(name, defines=()):
    key = (name,) + tuple(defines)
    if key not in _libs:
        tag = name + "".join("_" + d.replace("=", "") for d in defines)
        so = os.path.join(_HERE, "_build", f"lib{tag}.so")
        if _stale(so):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            cmd = ["g++", "-O2", "-std=c++17", "-DDSHEG_EMU", "-Wno-unknown-pragmas", "-Wno-attributes", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared",
                   "-I", _HERE, "-I", CSRC] + ["-D" + d for d in defines] + ["-o", so, os.path.join(_HERE, name + ".cpp")]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("emulator build failed:\n" + res.stderr)
        _libs[key] = ctypes.CDLL(so)
    return _libs[key]


def lib():
    """SIMT / mma.sync kernels (attention, sampler steps, post-processing)."""
    global _lib
    if _lib is None:
        _lib = _load("emu_kernels")
        _lib.emu_last_error.restype = ctypes.c_char_p
    return _lib


def gemm_lib(*defines):
    """The tcgen05 GEMM (gemm_tc.cuh) on the mbarrier / TMA / tcgen05 models of emu_tc_prims.h; `defines` selects an
    experiment build (a -D macro of gemm_tc.cuh)."""
    L = _load("emu_gemm", defines)
    L.emu_gemm_last_error.restype = ctypes.c_char_p
    return L


def attn_ws_lib():
    """attn_ws.cuh (warp-specialised TMA attention: mbarriers, 3-D TMA boxes, mma.sync, tensor-memory parking) on the models of
    emu_prims.h + emu_tc_prims.h."""
    L = _load("emu_attn_ws")
    L.emu_attn_ws_last_error.restype = ctypes.c_char_p
    return L


def gemm_tf32_lib():
    """gemm_tf32.cuh (tcgen05 kind::tf32, TFLOAT32 tensor maps, CTA pairs, two CTAs per SM) on the models of emu_tc_prims.h."""
    L = _load("emu_gemm_tf32")
    L.emu_gemm_tf32_last_error.restype = ctypes.c_char_p
    return L


def engine_lib():
    """The WHOLE engine (diffsheg_b200/csrc/engine.cu and every kernel it launches) on the emulator: the library exports the C ABI of
    include/diffsheg_b200.h itself; "device" pointers are host pointers (emu_runtime.h)."""
    from diffsheg_b200 import _lib as product
    L = _load("emu_engine", ("DSHEG_EMU_RUNTIME",))
    for name, sig in product.SIGNATURES.items():     # everything but the op-level test / bench entry points is in the emulated build
        if name.startswith(("dsheg_op_", "dsheg_bench_")):
            continue
        fn = getattr(L, name)
        fn.restype, fn.argtypes = sig
    L.emu_engine_last_launch_error.restype = ctypes.c_char_p
    L.emu_engine_launches.restype = ctypes.c_longlong
    L.emu_engine_graph_launches.restype = ctypes.c_longlong
    return L


class EmuEngine:
    """FusedUniDiffuser's create / load / finalize / prepare_window / denoise sequence (diffsheg_b200/engine.py) against the emulated
    engine, with CPU tensors.  The packer (diffsheg_b200/pack.py) is the product's own."""

    def __init__(self, state_dict, cfg, precision="fp32", max_batch=1, max_frames=None, sms=8):
        import torch
        from diffsheg_b200 import _lib as product
        from diffsheg_b200.engine import engine_config
        from diffsheg_b200.pack import pack_state_dict
        self.torch, self.cfg = torch, dict(cfg)
        self.L = L = engine_lib()
        L.emu_engine_set_sms(int(sms))
        self.max_frames = int(max_frames or cfg["n_poses"])
        c = engine_config(cfg, precision, int(max_batch), self.max_frames)
        h = ctypes.c_void_p()
        self._check(L.dsheg_create(ctypes.byref(c), 0, ctypes.byref(h)), None, "dsheg_create")
        self.h = h
        self._packed = pack_state_dict(state_dict, self.cfg, precision, self.max_frames)
        for name, (t, dt) in self._packed.items():
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            self._check(L.dsheg_load_tensor(h, name.encode(), ctypes.c_void_p(t.data_ptr()), dt, shape, t.dim()), h, name)
        self._check(L.dsheg_finalize_weights(h), h, "dsheg_finalize_weights")
        self._keep = None

    def _check(self, rc, h, what):
        if rc != 0:
            msg = self.L.dsheg_last_error(h)
            raise RuntimeError(f"emulated engine: {what} failed: {msg.decode() if msg else ''} "
                               f"[{self.L.emu_engine_last_launch_error().decode()}]")

    def prepare_window(self, mel, hubert, person_id):
        f = lambda t: t.to(self.torch.float32).contiguous()   # noqa: E731
        mel, hubert, person_id = f(mel), f(hubert), f(person_id)
        self._keep = (mel, hubert, person_id)
        self.B, self.T = mel.shape[0], mel.shape[1]
        self._check(self.L.dsheg_prepare_window(self.h, mel.data_ptr(), hubert.data_ptr(), person_id.data_ptr(), self.B, self.T, None),
                    self.h, "dsheg_prepare_window")

    def denoise(self, x, t_orig, a, b, cond_scale=None):
        x = x.to(self.torch.float32).contiguous()
        out = self.torch.empty_like(x)
        s = float(self.cfg.get("cond_scale", 1.0)) if cond_scale is None else float(cond_scale)
        self._check(self.L.dsheg_denoise(self.h, x.data_ptr(), int(t_orig), float(a), float(b), s, out.data_ptr(), None), self.h, "dsheg_denoise")
        return out

    def launch_count(self):
        return int(self.L.dsheg_launch_count(self.h))

    def emulated_launches(self):
        """Kernel launches the emulator has actually executed in this process (all engines)."""
        return int(self.L.emu_engine_launches())

    def graph_launches(self):
        """cudaGraphLaunch calls so far in this process: replays of a captured denoiser call (the small-batch regime of dsheg_denoise)."""
        return int(self.L.emu_engine_graph_launches())

    def close(self):
        if self.h:
            self.L.dsheg_destroy(self.h)
            self.h = None
