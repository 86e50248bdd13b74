"""Host emulation of the SIMT / mma.sync kernel SOURCES (test infrastructure only, never imported by the product).

``lib()`` compiles tests/emu/emu_kernels.cpp -- which #includes the kernel headers of diffsheg_b200/csrc with -DDSHEG_EMU --
with g++ into tests/emu/_build/ and loads it through ctypes."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "diffsheg_b200", "csrc")
_SO = os.path.join(_HERE, "_build", "libemu_kernels.so")
_lib = None


def _stale():
    if not os.path.exists(_SO):
        return True
    m = os.path.getmtime(_SO)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".h", ".cpp"))]
    srcs += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(s) > m for s in srcs)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            cmd = ["g++", "-O2", "-std=c++17", "-DDSHEG_EMU", "-Wno-unknown-pragmas", "-Wno-attributes", "-ffp-contract=off", "-fPIC", "-shared",
                   "-I", _HERE, "-I", CSRC, "-o", _SO, os.path.join(_HERE, "emu_kernels.cpp")]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("emulator build failed:\n" + res.stderr)
        _lib = ctypes.CDLL(_SO)
        _lib.emu_last_error.restype = ctypes.c_char_p
    return _lib
