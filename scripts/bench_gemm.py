"""Device-time sweep of the tcgen05 GEMM over the denoiser's shapes (B=950 SHOW/CFG), BN=128 vs 256."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401

from diffsheg_b200 import _lib  # noqa: E402

L = _lib.lib()
R, R1 = 167200, 83600
shapes = [("qkv      LN", R, 1536, 512, 1), ("feat1 LN+silu", R1, 1024, 896, 2), ("feat1g LN+silu", R1, 1024, 1024, 2),
          ("feat2 +res", R1, 512, 1024, 3), ("sa_out +res", R, 512, 512, 3), ("ffn1 gelu", R, 1024, 512, 4),
          ("ffn2 bias", R, 512, 1024, 0), ("qkv LN+expo", R, 1536, 512, 5), ("ffn2 +lnms", R, 512, 1024, 6), ("aud qkv", R1, 384, 128, 1), ("aud ffn1", R1, 1024, 128, 4),
          ("aud ffn2", R1, 128, 1024, 0), ("audproj", R1, 256, 256, 0)]
ms = ctypes.c_float()
for name, M, N, K, mode in shapes:
    line = f"{name:16s} M={M:7d} N={N:5d} K={K:5d}"
    for bn in (256, 2256):   # 256 = single-CTA 128x256 tiles, 2256 = CTA pair (cta_group::2) 256x256 tiles
        if N % 256:
            line += "   n/a      "
            continue
        rc = L.dsheg_bench_gemm(M, N, K, mode, bn, 10, ctypes.byref(ms))
        if rc:
            line += f"   bn{bn}: ERR {L.dsheg_last_error(None)}"
            continue
        tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12
        line += f"   {'pair' if bn > 1000 else 'cta1'}: {ms.value * 1e3:8.1f} us {tf:7.1f} TF/s"
    print(line, flush=True)
