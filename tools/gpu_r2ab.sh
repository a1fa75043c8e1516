# Round 2, call AB: the gradient kernel's producer issues its first two tiles before the wait for the lattice kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_early_return.py tests/test_gpu_shard.py tests/test_gpu_peer.py -m gpu -q --maxfail=5 > gpurun_out/r2ab_pytest.txt 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2ab_pytest.txt
{
for rep in 1 2; do
timeout 300 python tools/share_sweep.py c2 --shares 100 --opt 15=0
timeout 300 python tools/share_sweep.py c2 --shares 100 --opt 15=1
done
timeout 300 python tools/share_sweep.py c3 --shard 0/8 --shares 100 --opt 15=0
timeout 300 python tools/share_sweep.py c3 --shard 0/8 --shares 100 --opt 15=1
timeout 300 python tools/kernel_times.py c2 --iters 20 | grep -v cost-only
} > gpurun_out/r2ab_times.txt 2>&1
cat gpurun_out/r2ab_times.txt
