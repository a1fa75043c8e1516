set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -2
for i in 1 2; do
timeout 300 python tools/dropin_time.py c2 2>&1 | tail -1
done
timeout 300 python tools/dropin_time.py c5 2>&1 | tail -1
timeout 300 python tools/dropin_time.py c3 2>&1 | tail -1
