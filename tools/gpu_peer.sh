set -x
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench16_n1.json 2> gpurun_out/bench16_n1.err; echo "rc=$?"
for c in fused nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 5 --collective $c > gpurun_out/bench16_n2_$c.json 2> gpurun_out/bench16_n2_$c.err; echo "rc=$?"
tail -5 gpurun_out/bench16_n2_$c.err
done
python - <<'PY'
import json
for f in ("n1","n2_fused","n2_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/bench16_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["kernels_ms"], d["e2e"]["value"], d["config"].get("collective"), d.get("collective_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
