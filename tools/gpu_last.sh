set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/bench24.json 2> gpurun_out/bench24.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench24.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernels_ms"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["dram_frac"], d["clocks"])
PY
