#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn_${TAG} -f python scripts/bench_ops.py attn > gpurun_out/ncu_attn_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2cta_kernel -s 3 -c 2 -o gpurun_out/prof_gemm_${TAG} -f python scripts/bench_ops.py gemm > gpurun_out/ncu_gemm_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
