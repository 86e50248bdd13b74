"""Build and load libdiffsheg_b200.so (the C-ABI CUDA library) through ctypes.

There is no CPU fallback: if the library is missing or fails to load, every product entry
point raises.  ``build()`` cross-compiles for sm_100a with nvcc (works without a GPU).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("DSHEG_LIB") or os.path.join(_HERE, "libdiffsheg_b200.so")   # DSHEG_LIB: experiment builds
HEADER = os.path.join(ROOT, "include", "diffsheg_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

PREC = {"fp32": 0, "bf16": 1, "tf32": 2}   # tf32: fp32 activations + tcgen05 kind::tf32 GEMMs
COND_PROJECTION = {"mlp_includeX": 0, "linear_includeX": 1, "mlp_excludeX": 2, "linear_excludeX": 3}   # DSHEG_COND_*
ABI_VERSION = 2


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [HEADER]


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    so_m = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(s) > so_m for s in _sources())


def build(force=False, verbose=False):
    """Compile diffsheg_b200/csrc/engine.cu (which includes every kernel) in-tree."""
    if os.environ.get("DSHEG_LIB"):   # an experiment build (scripts/build_variants.sh): never overwrite it with a default build
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"DSHEG_LIB={SO_PATH} does not exist (build it with scripts/build_variants.sh)")
        return SO_PATH
    if not force and not needs_build():
        return SO_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", SO_PATH, os.path.join(CSRC, "engine.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO_PATH


class Config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "abi_version", "dim_pose", "expression_dim", "audio_dim", "hubert_dim", "aud_latent_dim",
        "latent_dim", "num_layers", "num_heads", "ff_size", "style_dim", "classifier_free",
        "precision", "max_batch", "max_frames", "cond_projection", "no_cond_residual")]


_lib = None

_P = ctypes.c_void_p
_I32, _I64, _F = ctypes.c_int32, ctypes.c_int64, ctypes.c_float

SIGNATURES = {
    "dsheg_last_error": (ctypes.c_char_p, [_P]),
    "dsheg_create": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.c_int, ctypes.POINTER(_P)]),
    "dsheg_destroy": (None, [_P]),
    "dsheg_load_tensor": (ctypes.c_int, [_P, ctypes.c_char_p, _P, _I32, ctypes.POINTER(_I64), _I32]),
    "dsheg_finalize_weights": (ctypes.c_int, [_P]),
    "dsheg_prepare_window": (ctypes.c_int, [_P, _P, _P, _P, _I32, _I32, _P]),
    "dsheg_denoise": (ctypes.c_int, [_P, _P, _I32, _F, _F, _F, _P, _P]),
    "dsheg_launch_count": (_I64, [_P]),
    "dsheg_profile_begin": (ctypes.c_int, [_P]),
    "dsheg_profile_end": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(_I64)]),
    "dsheg_ddim_step": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _I32, _F, _F, _F, _F, _P, _P, _P, _I32, _I32, _P, _P]),
    "dsheg_undo_step": (ctypes.c_int, [_P, _P, _P, _I64, _F, _F, _P]),
    "dsheg_ddpm_step": (ctypes.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _P, _P]),
    "dsheg_repaint_merge": (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, _F, _F, _P]),
    "dsheg_inv_standardize": (ctypes.c_int, [_P, _I32, _P, _P, _P, _I32, _I64, _I32, _P]),
    "dsheg_beat_axis_angle_to_euler": (ctypes.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "dsheg_resample_linear": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "dsheg_mel_spectrogram": (ctypes.c_int, [_P, _I64, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _I32, _P]),
    "dsheg_op_linear": (ctypes.c_int, [_I32, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "dsheg_op_linear_fused": (ctypes.c_int, [_I32, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "dsheg_bench_gemm": (ctypes.c_int, [_I32, _I32, _I32, _I32, _I32, _I32, ctypes.POINTER(ctypes.c_float)]),
    "dsheg_op_attention": (ctypes.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "dsheg_op_attention_tf32": (ctypes.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "dsheg_op_attention_bf16": (ctypes.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "dsheg_op_cross_attention_bf16": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
}


def lib():
    """The loaded library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            try:  # the CUDA library itself is the only implementation: build it if a toolchain is present
                build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    f"{SO_PATH} not found and could not be built ({e}); build it with "
                    "`python -c 'import __graft_entry__ as g; g.build()'` (diffsheg_b200 has no CPU / PyTorch fallback)")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def last_error(handle=None):
    msg = lib().dsheg_last_error(handle)
    return msg.decode() if msg else ""


def check(rc, handle=None, what=""):
    if rc != 0:
        raise RuntimeError(f"diffsheg_b200 {what} failed: {last_error(handle)}")
