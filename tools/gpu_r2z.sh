set -x
mkdir -p gpurun_out
{
timeout 300 python tools/kernel_times.py c2 --iters 20 --share 100,80,60,40,20,1
timeout 300 python tools/kernel_times.py c2 --iters 20 --zero 0
} 2>&1 | grep -v cost-only > gpurun_out/r2z_times.txt
cat gpurun_out/r2z_times.txt
