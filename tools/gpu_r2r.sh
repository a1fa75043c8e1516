# Round 2, call R: where the lattice kernel's time goes on one rank's shard of c3 at N = 8 (8 utterances, T up to 400).
set -x
mkdir -p gpurun_out
{
timeout 300 python tools/kernel_times.py c3 --shard 0/8 --iters 20 --zero=-1,0,1,2,4
timeout 300 python tools/kernel_times.py c3 --shard 0/8 --iters 20 --parts 4
timeout 300 python tools/kernel_times.py c3 --shard 0/8 --iters 20 --parts 2
timeout 300 python tools/kernel_times.py c3 --shard 0/4 --iters 20 --zero=-1,0
timeout 300 python tools/kernel_times.py c3 --shard 0/2 --iters 20 --zero=-1,0
echo "== k2_probe T=400 S=80 B=8 parts=8 K=1 zero=2 V=1024"; ./tools/k2_probe 400 80 8 8 1 2 1024 | grep -v "per chunk"
echo "== k2_probe T=400 S=80 B=8 parts=8 K=1 zero=0 V=1024"; ./tools/k2_probe 400 80 8 8 1 0 1024 | grep -v "per chunk"
echo "== k2_probe T=400 S=80 B=8 parts=4 K=1 zero=0 V=1024"; ./tools/k2_probe 400 80 8 4 1 0 1024 | grep -v "per chunk"
} > gpurun_out/r2r_shard_times.txt 2>&1
grep -v "^+" gpurun_out/r2r_shard_times.txt | grep -v cost-only | cut -c1-260
