# Round 2, call AA: the lattice kernel takes part in the shared zero fill (alignment band: c5)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_upload.py tests/test_gpu_unaligned.py -m gpu -q --maxfail=5 > gpurun_out/r2aa_pytest.txt 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2aa_pytest.txt
timeout 300 python tools/share_sweep.py c5 --shares 100 --steps 50 > gpurun_out/r2aa_c5.txt 2>&1
timeout 300 python tools/kernel_times.py c5 --iters 20 2>&1 | grep -v cost-only >> gpurun_out/r2aa_c5.txt
cat gpurun_out/r2aa_c5.txt
timeout 500 python bench.py --workload c5 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2aa_bench_c5.json 2> gpurun_out/r2aa_bench_c5.err; echo "bench c5 rc=$?"
tail -c 500 gpurun_out/r2aa_bench_c5.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2aa_bench_c5.json').read().strip().splitlines()[-1])
print('c5 value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'kernels', d['kernels_ms'], 'parity', d.get('parity'))
PY
