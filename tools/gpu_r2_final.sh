# Round 2, final pass on one B200 for the shipped binary: the whole GPU suite (with the drop-in build), smoke, both bench
# arms on c2 as the driver runs them, the other named shapes and the unaligned-vocabulary variants.
set -x
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 --durations=6 > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r2f_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_c2_driver.json 2> gpurun_out/r2f_bench_c2_driver.err; echo "bench c2 (driver's flags) rc=$?"
timeout 500 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_c2_ref.json 2> gpurun_out/r2f_bench_c2_ref.err; echo "ref rc=$?"
for wl in c3 c5 c4; do
  timeout 700 python bench.py --workload $wl --steps 30 --warmup 5 > gpurun_out/r2f_bench_$wl.json 2> gpurun_out/r2f_bench_$wl.err; echo "bench $wl rc=$?"
done
for wl in c2v1025 c4v5001; do
  timeout 500 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_$wl.json 2> gpurun_out/r2f_bench_$wl.err; echo "bench $wl rc=$?"
done
timeout 300 python tools/dropin_time.py 2>&1 | tail -3
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2f_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r = d.get('roofline') or {}
    c = d.get('call_roofline') or {}
    print(f.split('/')[-1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'full wait', (d.get('full_wait') or {}).get('ms_per_step'),
          'e2e', round((d.get('e2e') or {}).get('value', 0), 1), (d.get('e2e') or {}).get('ms_per_step'),
          'k3 frac', round(r.get('frac', 0), 3), 'dram_frac', r.get('dram_frac'), 'call req', c.get('required_frac_of_measured_peak'), 'call dram', c.get('dram_frac_of_measured_peak'),
          'kernels', {k: round(v, 4) for k, v in (d.get('kernels_ms') or {}).items() if k != 'k1_GBps_of_live_logits' and k != 'k1_GBps_of_4N'},
          'parity', {k: v for k, v in (d.get('parity') or {}).items() if k.startswith(('cost_max', 'grad_max'))}, 'launches', d.get('gpu_launches'),
          'cpu', (d.get('cpu_baseline') or {}).get('value'), 'alloc', (d.get('per_call_workspace') or {}).get('ms_per_step'),
          'refcuda', (d.get('reference_cuda_same_gpu') or {}).get('ms_per_call'))
PY
