# 8 GPUs of one box: the headline shape with the cost sum over peer memory.
set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench22_n8.json 2> gpurun_out/bench22_n8.err; echo "n8 rc=$?"
tail -3 gpurun_out/bench22_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench22_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernels_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("collective"), d.get("collective_check"), d["clocks"])
PY
