#!/bin/bash
# What the driver runs at round end, plus the reference arm: GPU tests, smoke, both bench arms.  Run under gpurun.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/ci_pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/ci_rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/ci_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ci_rc.txt
timeout 600 python bench.py --impl reference > gpurun_out/ci_bench_reference.json 2> gpurun_out/ci_bench_reference.err; echo "ref rc=$?" >> gpurun_out/ci_rc.txt
timeout 900 python bench.py > gpurun_out/ci_bench.json 2> gpurun_out/ci_bench.err; echo "bench rc=$?" >> gpurun_out/ci_rc.txt
cat gpurun_out/ci_rc.txt; grep -E "passed|failed|rror" gpurun_out/ci_pytest_gpu.log | tail -3; tail -2 gpurun_out/ci_smoke.log
