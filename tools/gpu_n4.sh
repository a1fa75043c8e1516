# 4 GPUs: the cost sum over peer memory at a world above two, against the NCCL step; the ragged shape at N = 4.
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 200 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -3
run() { # name, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 100 --warmup 5 $2 > gpurun_out/bench19_$1.json 2> gpurun_out/bench19_$1.err; echo "$1 rc=$?"
  tail -3 gpurun_out/bench19_$1.err
}
run n4_fused ""
run n4_nccl "--collective nccl"
run n4_c3 "--workload c3"
python - <<'PY'
import json
for f in ("n4_fused","n4_nccl","n4_c3"):
    try:
        d=json.loads(open(f"gpurun_out/bench19_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["kernels_ms"], d["e2e"]["value"], d["config"].get("collective"), d.get("collective_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
