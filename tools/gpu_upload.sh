# The live-row upload: parity, then the e2e leg of bench.py both ways on c2 and c5.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_upload.py -m gpu -x -q 2>&1 | tail -8
for w in c2 c5; do
  timeout 400 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench20_$w.json 2> gpurun_out/bench20_$w.err; echo "$w rc=$?"
  tail -3 gpurun_out/bench20_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench20_$w.json").read().strip().splitlines()[-1])
print("$w", d["ms_per_step"], d["value"]); print(json.dumps(d["e2e_paths"], indent=1))
PY
done
