mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/c6_pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/c6_rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 30 -c 6 -o $O/c6_gemm_tf32 python scripts/prof_denoise.py --batch 950 --calls 1 --precision tf32 > $O/c6_prof_tf32.log 2>&1
cat $O/c6_rc.txt; grep -E "passed|failed|rror" $O/c6_pytest_gpu.log | tail -5
