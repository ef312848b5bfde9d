#!/bin/bash
# ncu source-level capture of the short-K attention projection GEMMs (epilogue-bound): q|k projection (mask 129 = bias + bf16 out)
# and the PV product (mask 136 = 1/rowsum + bf16 out) of one DDPM++ evaluation at B=512.
set -e
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:gemm_kernel<\(int\)128, \(int\)(129|136), \(int\)1>' -s 2 -c 2 -o gpurun_out/prof_shortk -f \
    python tests/gpu_profile_eval.py 512 1 > gpurun_out/ncu_shortk.log 2>&1
tail -2 gpurun_out/ncu_shortk.log
