"""Summary of the round-2 hardware session: one line per gpurun_out/[rc][0-9]_bench*.json (frames/s, GEMM TF/s, row-wise ms, attention GB/s, clock)."""
import glob
import json
import os

for f in sorted(glob.glob("gpurun_out/[rc][0-9]_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{os.path.basename(f):44s} {d['value']:10.0f} frames/s  gemm {d['roofline']['achieved']:7.1f} TF/s ({d['roofline']['ms_per_step']:6.1f} ms)"
              f"  rowwise {d.get('rowwise', dict()).get('ms_per_step', 0):5.1f} ms  attention {d['roofline_attention']['achieved']:6.0f} GB/s"
              f" ({d['roofline_attention']['ms_per_step']:6.1f} ms)  sm {d['clocks']['sm_mhz']}")
    except Exception as e:  # noqa: BLE001
        print(os.path.basename(f), "failed:", e)
