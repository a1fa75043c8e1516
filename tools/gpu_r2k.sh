# Round 2, call K: tile-size rule near a power-of-two boundary (V = 1025 back to 32 KB tiles): unaligned tests, times.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unaligned.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q --maxfail=10 > gpurun_out/r2k_pytest.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2k_pytest.txt
timeout 400 python tools/kernel_times.py c2 c2v1025 c4v5001 --iters 20 2>&1 | grep -v cost-only > gpurun_out/r2k_times.txt 2>&1
cut -c1-235 gpurun_out/r2k_times.txt
