# Round 2: a last look at the freshly built tree (every .so rebuilt): drop-in + early-return + parity subset, smoke, the driver's bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_early_return.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h_bench_c2.json').read().strip().splitlines()[-1])
print('c2', round(d['value'], 1), d['ms_per_step'], 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'], 3), d['build'])
PY
