# Round 2, call Y: the zero fill split between the lattice kernel (front of the batch) and the gradient kernel's zero-fill warp
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "zero or fill or dead" --maxfail=5 > gpurun_out/r2y_pytest.txt 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2y_pytest.txt
{
timeout 300 python tools/share_sweep.py c2 --shares 100,90,80,70,60,50,40,100
timeout 300 python tools/share_sweep.py c3 --shares 100,85,70,55,40 --steps 30
timeout 300 python tools/share_sweep.py c3 --shard 0/8 --shares 100,80,65,50,35
timeout 300 python tools/share_sweep.py c3 --shard 0/4 --shares 100,80,65,50
} > gpurun_out/r2y_share.txt 2>&1
cat gpurun_out/r2y_share.txt
