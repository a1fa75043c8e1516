# Round 2, call L: the round-end pass on one B200 for the current binary: the whole GPU suite (with the drop-in build), smoke,
# both bench arms on c2, the other named shapes and the two unaligned-vocabulary variants, and the round-1 bench on the same box.
set -x
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 --durations=6 > gpurun_out/r2l_pytest.txt 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r2l_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
timeout 500 python bench.py --steps 100 --warmup 5 > gpurun_out/r2l_bench_c2.json 2> gpurun_out/r2l_bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_c2_ref.json 2> gpurun_out/r2l_bench_c2_ref.err; echo "ref rc=$?"
(cd tools/_r1 && timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > ../../gpurun_out/r2l_bench_c2_round1_binary.json 2> ../../gpurun_out/r2l_bench_c2_round1_binary.err); echo "r1 bench rc=$?"
for wl in c3 c5 c4; do
  timeout 700 python bench.py --workload $wl --steps 30 --warmup 5 > gpurun_out/r2l_bench_$wl.json 2> gpurun_out/r2l_bench_$wl.err; echo "bench $wl rc=$?"
done
for wl in c2v1025 c4v5001; do
  timeout 500 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_$wl.json 2> gpurun_out/r2l_bench_$wl.err; echo "bench $wl rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2l_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r = d.get('roofline') or {}
    print(f.split('/')[-1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round((d.get('e2e') or {}).get('value', 0), 1),
          'k3 frac', round(r.get('frac', 0), 3), 'dram_frac', r.get('dram_frac'), 'kernels', {k: round(v, 4) for k, v in (d.get('kernels_ms') or {}).items() if k != 'k1_GBps_of_live_logits' and k != 'k1_GBps_of_4N'},
          'parity', {k: v for k, v in (d.get('parity') or {}).items() if k.startswith(('cost_max', 'grad_max'))}, 'launches', d.get('gpu_launches'))
PY
