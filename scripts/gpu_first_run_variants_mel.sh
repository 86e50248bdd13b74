#!/bin/bash
# First hardware run of what was built after round 2's GPU budget was spent: the cond_projection / cond_residual variants
# (host orchestration on the shipped kernels) and the mel front-end kernel.  One B200, about 6 minutes:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/gpu_first_run_variants_mel.sh'
mkdir -p gpurun_out; O=gpurun_out
# 1. the two new GPU test files (they sort after the validated suite in a plain `pytest -m gpu`)
timeout 300 python -m pytest tests/test_variants_gpu.py tests/test_wave_frontend.py -m gpu -q -p no:cacheprovider -s > $O/fr_pytest.log 2>&1; echo "pytest rc=$?" > $O/fr_rc.txt
grep "\[parity\]" $O/fr_pytest.log > $O/fr_parity_lines.txt       # measured relmax / per_channel / rel_rms of every case: set the gates at 2x of these
# 2. memcheck of one variant call with a memset node and a 2-D copy (linear_includeX without cond_residual), and of the mel kernel
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_variants_gpu.py -m gpu -q -p no:cacheprovider \
    -k "golden and beat and linear_includeX and False and bf16" > $O/fr_memcheck_variant.log 2>&1; echo "memcheck variant rc=$?" >> $O/fr_rc.txt
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_wave_frontend.py -m gpu -q -p no:cacheprovider \
    -k "7.3" > $O/fr_memcheck_mel.log 2>&1; echo "memcheck mel rc=$?" >> $O/fr_rc.txt
# 3. what the generic (unfused-statistics) layer path costs at the headline size: one variant line next to the shipped one
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --cond-projection mlp_excludeX > $O/fr_bench_mlp_excludeX.json 2> $O/fr_bench_mlp_excludeX.err; echo "bench variant rc=$?" >> $O/fr_rc.txt
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/fr_bench_shipped.json 2> $O/fr_bench_shipped.err; echo "bench shipped rc=$?" >> $O/fr_rc.txt
# 4. the mel kernel's launch time (60 s clip = 901 CTAs)
timeout 60 python - > $O/fr_mel_time.txt 2>&1 <<'PY'
import torch
from diffsheg_b200 import mel_spectrogram
y = torch.randn(18000 * 60, device="cuda") * 0.1
for _ in range(3): mel_spectrogram(y)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): mel_spectrogram(y)
e1.record(); torch.cuda.synchronize()
print("mel_spectrogram, 60 s clip (901 frames): %.1f us per call incl. the output allocation" % (e0.elapsed_time(e1) / 20 * 1e3))
PY
cat $O/fr_rc.txt; tail -3 $O/fr_pytest.log; cat $O/fr_mel_time.txt
