"""Three representative GEMM shapes for an ncu capture (qkv LN-fold, sa_out +residual, ffn2 bias)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa
from diffsheg_b200 import _lib
L = _lib.lib()
ms = ctypes.c_float()
R = 167200
for name, M, N, K, mode in (("qkv", R, 1536, 512, 1), ("sa_out", R, 512, 512, 3), ("ffn2", R, 512, 1024, 0)):
    rc = L.dsheg_bench_gemm(M, N, K, mode, 256, 1, ctypes.byref(ms))
    print(name, rc, ms.value)
