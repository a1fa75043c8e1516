# Round 2, call AG: free_workspace without a wait of the host (the cached block carries the event behind its last user)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_early_return.py tests/test_gpu_dropin.py tests/test_gpu_parity.py tests/test_gpu_peer.py tests/test_gpu_concurrent.py -m gpu -q --maxfail=5 > gpurun_out/r2ag_pytest.txt 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2ag_pytest.txt
for i in 1 2; do timeout 300 python tools/dropin_time.py c2 2>&1 | tail -1; done
timeout 300 python tools/dropin_time.py c5 2>&1 | tail -1
timeout 300 python tools/dropin_time.py c3 2>&1 | tail -1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2ag_bench_c2.json 2> gpurun_out/r2ag_bench_c2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ag_bench_c2.json').read().strip().splitlines()[-1])
print('c2', round(d['value'], 1), d['ms_per_step'], 'per_call_workspace', d['per_call_workspace'])
PY
