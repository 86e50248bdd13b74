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
    experiment build (e.g. "DSHEG_K512_DEEP=1")."""
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
