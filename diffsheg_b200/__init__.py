"""diffsheg_b200: B200-native (sm_100a) DDIM/DDPM sampling hot path of DiffSHEG.

Host-side mirror of the reference's sampling seams over libdiffsheg_b200.so (hand-written CUDA,
C ABI in include/diffsheg_b200.h).  There is no CPU or PyTorch fallback.
"""
from .diffusion import (FusedGaussianDiffusion, FusedSpacedDiffusion, get_named_beta_schedule,  # noqa: F401
                        get_schedule_jump_cjm_ddim, get_schedule_jump_paper, space_timesteps)
from .engine import FusedUniDiffuser, cfg_from_opt  # noqa: F401
from .frontend import audio_embedding, mel_spectrogram  # noqa: F401
from .postprocess import axis_angle_to_euler, finish_beat, finish_show, inv_standardize, resample_features  # noqa: F401
from .trainer import build_diffusions, generate_batch, generate_long, get_windows, patch_trainer  # noqa: F401

__all__ = ["FusedGaussianDiffusion", "FusedSpacedDiffusion", "FusedUniDiffuser", "cfg_from_opt", "space_timesteps",
           "get_named_beta_schedule", "get_schedule_jump_cjm_ddim", "get_schedule_jump_paper", "build_diffusions",
           "generate_batch", "generate_long", "get_windows", "patch_trainer", "inv_standardize", "axis_angle_to_euler",
           "finish_show", "finish_beat", "resample_features", "mel_spectrogram", "audio_embedding"]
