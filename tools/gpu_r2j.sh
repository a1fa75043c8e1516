# Round 2, call J: unaligned rows as interior + one edge element per lane: tests, times.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unaligned.py tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_gpu_envelope.py tests/test_gpu_fuzz.py tests/test_gpu_upload.py -m gpu -q --maxfail=10 > gpurun_out/r2j_pytest.txt 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2j_pytest.txt
{
(cd tools/_r1 && timeout 200 python tools/kernel_times.py c2 --iters 30 2>&1 | grep -v cost-only | sed 's/^/R1   /')
timeout 400 python tools/kernel_times.py c2 c3 c2v1025 c4 c4v5001 --iters 20 2>&1 | grep -v cost-only | sed "s/^/NEW  /"
} > gpurun_out/r2j_times.txt 2>&1
grep -v "^+" gpurun_out/r2j_times.txt | cut -c1-235
