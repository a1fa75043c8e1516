set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "zeroing" 2>&1 | tail -3
timeout 300 python tools/kernel_times.py c2 --iters 30 --dyn 0,1,2,3,0,2 2>&1 | grep -v cost-only
timeout 300 python tools/kernel_times.py c3 --iters 10 --dyn 0,1,2 2>&1 | grep -v cost-only
timeout 300 python tools/kernel_times.py c4 --iters 5 --dyn 0,1,2 2>&1 | grep -v cost-only
timeout 300 python tools/kernel_times.py c2 --bf16 --iters 20 --dyn 0,1,2 2>&1 | grep -v cost-only
