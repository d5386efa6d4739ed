#!/bin/bash
# ncu --set full captures of the self-attention kernels (attention2.cu, attention6.cu), with source:  gpurun -- 'bash scripts/gpu_ncu_attn.sh <tag>'
TAG=${1:-attn}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn2_${TAG} -f python scripts/bench_ops.py attnself > gpurun_out/ncu_attn2_${TAG}.log 2>&1
CE_ATTN_V2=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention6_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn6_${TAG} -f python scripts/bench_ops.py attnself > gpurun_out/ncu_attn6_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
