#!/bin/bash
# Experiment builds of the library (selected at run time with DSHEG_LIB=<path>, see diffsheg_b200/_lib.py): compile-time
# knobs of the tcgen05 GEMM that cannot be switched by an environment variable.  Run here (no GPU needed); the .so files
# travel to the GPU box with the snapshot.  scripts/gpu_round2_first.sh times them back to back on one box.
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
build() { echo "building $1: $2"; $NVCC $FLAGS $2 -o build_variants/libdiffsheg_b200_$1.so diffsheg_b200/csrc/engine.cu; }
build k512deep "-DDSHEG_K512_DEEP=1" &     # K = 512 pair GEMMs: 5 stages with wide epilogue boxes
build stages3 "-DDSHEG_PAIR_STAGES=3" &    # control: one stage less everywhere (reproduces the round-1 sweep direction)
wait
build split73 "-DDSHEG_SPLIT_RINGS=1" &                                    # K = 512 pair GEMMs: A ring 7 deep, W ring 3 deep, two producers
build split64 "-DDSHEG_SPLIT_RINGS=1 -DDSHEG_SPLIT_A=6 -DDSHEG_SPLIT_W=4" &
wait
build epipacked "-DDSHEG_EPI_PACKED=1" &                                   # GEMM epilogue math on packed fp32 (FFMA2 / FMUL2 / FADD2)
build pdl "-DDSHEG_PDL=1"                                                  # programmatic dependent launch in every bf16 hot-path kernel (griddepcontrol)
wait
ls -la build_variants
