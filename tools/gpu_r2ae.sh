# Round 2, call AE: the plan in one launch
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py tests/test_gpu_upload.py tests/test_gpu_unaligned.py tests/test_gpu_shard.py -m gpu -q --maxfail=5 > gpurun_out/r2ae_pytest.txt 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2ae_pytest.txt
timeout 300 python tools/dropin_time.py c2 2>&1 | tail -1
timeout 300 python tools/dropin_time.py c5 2>&1 | tail -1
timeout 300 python tools/dropin_time.py c3 2>&1 | tail -1
