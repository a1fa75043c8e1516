# Round 2, call N: the upload with its copy-engine part (tests + the e2e legs of c2, c3, c5), one B200.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_upload.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2n_pytest.txt
for wl in c2 c3 c5; do
  timeout 500 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_$wl.json 2> gpurun_out/r2n_bench_$wl.err; echo "bench $wl rc=$?"
  tail -c 800 gpurun_out/r2n_bench_$wl.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2n_bench_*.json')):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'k3', d['roofline']['frac'], d['roofline']['dram_frac'],
          'call required frac', round(d['call_roofline']['required_frac_of_measured_peak'], 3), 'dram', d['call_roofline']['dram_frac_of_measured_peak'])
    for k, v in d['e2e_paths'].items():
        print('   ', k, round(v['ms_per_step'], 3), 'ms', round(v['value'], 1), 'utt/s', round(v['host_read_GBps_all_ranks'], 2), 'GB/s', v['h2d_bytes_per_step'])
PY
