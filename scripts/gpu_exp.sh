mkdir -p gpurun_out
for st in 2 3; do
  DSHEG_LIB=$PWD/diffsheg_b200/libdiffsheg_b200_st$st.so timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep_st$st.log 2>&1
done
timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep_st4.log 2>&1
for st in 2 3 4; do echo "--- pair ring stages = $st"; cut -c1-150 gpurun_out/gemm_sweep_st$st.log | head -7; done
