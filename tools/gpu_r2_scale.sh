# Round 2: BASELINE.json configs[2] batch-sharded over N GPUs (strong scaling) + the weak-scaling c2 line at the same N.
#   gpurun --gpus N -- 'bash tools/gpu_r2_scale.sh N'
set -x
N=${1:-1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" = "1" ]; then
  RUN="python"
else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
fi
for part in lpt contiguous; do
  timeout 600 $RUN bench.py --gpus $N --workload c3 --scaling strong --partition $part --steps 30 --warmup 5 --no-cpu-baseline \
      > gpurun_out/r2_c3_strong_${part}_n$N.json 2> gpurun_out/r2_c3_strong_${part}_n$N.err; echo "c3 strong $part rc=$?"
  tail -c 600 gpurun_out/r2_c3_strong_${part}_n$N.err
  [ "$N" = "1" ] && break
done
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_c2_weak_n$N.json 2> gpurun_out/r2_c2_weak_n$N.err; echo "c2 weak rc=$?"
# the other named shapes, the named batch on every GPU (weak): utt/s, GB/s and the roofline fraction per N
for wl in c3 c5 c4; do
  EXTRA=""; [ "$wl" = "c4" ] && [ "$N" != "1" ] && EXTRA="--no-e2e"   # (15.5 GB of pinned host memory per rank)
  timeout 600 $RUN bench.py --gpus $N --workload $wl --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2_${wl}_weak_n$N.json 2> gpurun_out/r2_${wl}_weak_n$N.err; echo "$wl weak rc=$?"
done
if [ "$N" = "2" ]; then
  # the CUDA-IPC boards between two processes (skipped on one GPU), and the NCCL fallback of the step
  timeout 400 python -m pytest tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -3
  timeout 600 $RUN bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --collective nccl > gpurun_out/r2_c2_weak_nccl_n2.json 2> gpurun_out/r2_c2_weak_nccl_n2.err; echo "nccl fallback rc=$?"
fi
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_c*_n$N.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'blocks', [round(x, 4) for x in d['timing']['ms_per_step_blocks']],
          'shards', (d.get('shards') or {}).get('imbalance_max_over_mean'), 'bit-equal', (d.get('shards') or {}).get('costs_bit_identical_to_the_whole_batch_on_one_gpu'),
          'e2e', round((d.get('e2e') or {}).get('value', 0)), 'host GB/s', round((d.get('e2e') or {}).get('host_read_GBps_all_ranks', 0), 1), 'coll', d['config']['collective'][:30])
PY
