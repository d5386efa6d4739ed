mkdir -p gpurun_out
echo "=== tests" | tee gpurun_out/tests_r2j.log
timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | tee -a gpurun_out/tests_r2j.log
S="--steps 6 --warmup 3 --no-vae --no-cpu-baseline --no-library-bar"
for V in 1 0 1 0; do
  echo "=== CE_DIT_STATS=$V" | tee -a gpurun_out/bench_ab_r2j.log
  CE_DIT_STATS=$V timeout 400 python bench.py $S 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k: d[k] for k in ('value','ms_per_step','gpu_launches')}), d['roofline']['rows_ms_total'], d['roofline']['ms_total'], d['roofline']['attention']['ms_total'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/bench_ab_r2j.log
done
echo "=== cuda graph" | tee -a gpurun_out/bench_ab_r2j.log
timeout 400 python bench.py $S --cuda-graph 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k: d[k] for k in ('value','ms_per_step','gpu_launches')}), d['clocks']['sm_mhz'])" | tee -a gpurun_out/bench_ab_r2j.log
