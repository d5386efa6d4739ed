#!/bin/bash
# In-loop A/B of the self-attention kernels (attention2.cu vs attention6.cu) inside the 14B / 720p step
mkdir -p gpurun_out
OUT=gpurun_out/bench_ab_attn6_${1:-r2u}.log
: > $OUT
S="--steps 6 --warmup 3 --no-vae --no-cpu-baseline --no-library-bar"
for V in 2 6 2 6; do
  echo "=== CE_ATTN_V2=$V" | tee -a $OUT
  CE_ATTN_V2=$V timeout 400 python bench.py $S 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k: d[k] for k in ('value','ms_per_step','gpu_launches')}), 'gemm_ms', d['roofline']['ms_total'], 'attn_ms', d['roofline']['attention']['ms_total'], 'attn_tflops', d['roofline']['attention'].get('achieved'), 'sm_mhz', d['clocks']['sm_mhz'])" | tee -a $OUT
done
