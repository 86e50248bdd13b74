"""Dry run of the GPU test bodies of tests/test_variants_gpu.py on the CPU: the emulated engine (tests/emu: csrc/engine.cu + every
kernel behind the production C ABI, graph replay included) stands in for FusedUniDiffuser, the product's sampler classes run over the
emulated step-kernel entry points, `.cuda()` is the identity.  Test infrastructure, offline (the pair and loop cases take 8 - 10
minutes each): it executes the TEST code -- names, keyword arguments, fixtures, gates -- at the GPU sizes before its first hardware run.

    python scripts/emu_dry_run_gpu_tests.py [golden|protocol|pair|loop|mel ...] > profiles/r02/emu/gpu_test_bodies_dry_run.txt
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import diffsheg_b200  # noqa: E402
import diffsheg_b200.diffusion as D  # noqa: E402
import test_sampler_host_logic as H  # noqa: E402


class StandIn(H.EmulatedEngine):
    def __init__(self, sd, cfg, precision="bf16", max_batch=1, max_frames=None, device=0):
        super().__init__(sd, cfg, precision, max_batch, int(max_frames or cfg["n_poses"]))


diffsheg_b200.FusedUniDiffuser = StandIn
D._lib, D._ptr, D._stream = H.EmulatedLibModule, (lambda t: t), (lambda device=None: None)
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.synchronize = lambda *a, **k: None
_full, _tensor = torch.full, torch.tensor
torch.full = lambda *a, **k: _full(*a, **{kk: v for kk, v in k.items() if kk != "device"})
torch.tensor = lambda *a, **k: _tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"})

import test_variants_gpu as t  # noqa: E402


def _mel_cases():
    """tests/test_wave_frontend.py's GPU tests: diffsheg_b200/frontend.py itself over the emulated dsheg_mel_spectrogram (a Tensor
    subclass answers is_cuda, torch.cuda.device is a null context)."""
    import contextlib

    import emu
    import test_wave_frontend as w
    from diffsheg_b200 import frontend as fe

    class FakeCuda(torch.Tensor):
        is_cuda = property(lambda self: True)

    L = emu.engine_lib()

    class LibMod:
        lib = staticmethod(lambda: L)

        @staticmethod
        def check(rc, handle=None, what=""):
            if rc != 0:
                raise RuntimeError(what + ": " + L.dsheg_last_error(None).decode())

    fe._lib, fe._stream = LibMod, (lambda device=None: None)
    torch.cuda.device = lambda d: contextlib.nullcontext()
    torch.Tensor.cuda = lambda self, *a, **k: self.as_subclass(FakeCuda)
    torch.Tensor.cpu = lambda self, *a, **k: self.as_subclass(torch.Tensor)
    return [(w.test_mel_spectrogram_matches_oracle, a) for a in ((60.0, "constant"), (7.3, "reflect"), (0.2, "constant"))] + \
        [(w.test_audio_embedding_is_the_trainers_tensor_and_bad_calls_fail_loudly, ())]


CASES = {
    "golden": [(t.test_variant_denoise_matches_reference_golden, (os.path.join(ROOT, "tests", "golden"),) + a) for a in (
        ("beat", "linear_excludeX", False, "bf16"), ("beat", "mlp_includeX", False, "fp32"), ("show", "mlp_excludeX", True, "tf32"))],
    "protocol": [(t.test_from_module_protocol_picks_the_variant_up_from_opt, ())],
    "pair": [(t.test_variant_denoise_at_a_cta_pair_row_count_matches_oracle, a) for a in (("mlp_excludeX", True), ("linear_includeX", True))],
    "loop": [(t.test_variant_ddim25_loop_matches_oracle, ("beat", "linear_excludeX", True, "bf16"))],
}

if __name__ == "__main__":
    for kind in (sys.argv[1:] or ["golden", "protocol"]):
        for fn, args in (_mel_cases() if kind == "mel" else CASES[kind]):
            t0 = time.time()
            fn(*args)
            print(f"PASSED on the emulated stack: {fn.__name__}{args[1:] if kind == 'golden' else args}  ({time.time() - t0:.0f} s)", flush=True)
