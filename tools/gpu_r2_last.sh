# Round 2, last check of the shipped tree on one B200: the whole GPU suite, smoke, both bench arms with the driver's flags.
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2g_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_c2_ref.json 2> gpurun_out/r2g_bench_c2_ref.err; echo "ref rc=$?"
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2g_bench_c2.json', 'gpurun_out/r2g_bench_c2_ref.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ('impl', 'value', 'ms_per_step', 'gpu_launches', 'vs_baseline', 'dtype', 'scaling')}, 'e2e', (d.get('e2e') or {}).get('value'),
          'roofline', {k: (d.get('roofline') or {}).get(k) for k in ('achieved', 'peak', 'frac', 'traffic', 'dram_frac')}, 'cpu', d.get('cpu_baseline'), 'clocks', d.get('clocks'))
PY
