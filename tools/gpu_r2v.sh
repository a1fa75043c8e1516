# Round 2, call V: early return of the synchronous call (costs on the host, gradient kernel in stream order) and K1 as a
# programmatic dependent of the previous call's K3: tests, then c2 / c3 / c5 through bench.py.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_early_return.py tests/test_gpu_peer.py tests/test_gpu_concurrent.py tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_upload.py -m gpu -q --maxfail=10 > gpurun_out/r2v_pytest.txt 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2v_pytest.txt
for wl in c2 c3 c5; do
  timeout 500 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench_$wl.json 2> gpurun_out/r2v_bench_$wl.err; echo "bench $wl rc=$?"
  tail -c 600 gpurun_out/r2v_bench_$wl.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2v_bench_*.json')):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'blocks', [round(x, 4) for x in d['timing']['ms_per_step_blocks']],
          'full wait', d['full_wait'] and round(d['full_wait']['ms_per_step'], 4), 'async', d['async_enqueue'] and round(d['async_enqueue']['ms_per_step'], 4),
          'kernels', {k: round(v, 4) for k, v in d['kernels_ms'].items() if k.startswith('k') and 'GBps' not in k}, 'e2e', round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches'])
PY
