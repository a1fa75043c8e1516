set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r2ad_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2ad_pytest.txt
timeout 300 python tools/dropin_time.py c5 2>&1 | tail -2
timeout 300 python tools/dropin_time.py c2 2>&1 | tail -2
