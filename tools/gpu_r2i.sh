# Round 2, call I: unaligned rows with interior / edge split; the suite without the drop-in build; times against round 1.
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 --ignore=tests/test_gpu_dropin.py > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2i_pytest.txt
{
(cd tools/_r1 && timeout 200 python tools/kernel_times.py c2 c3 --iters 30 2>&1 | grep -v cost-only | sed 's/^/R1   /')
timeout 400 python tools/kernel_times.py c2 c3 c2v1025 c4 c4v5001 --iters 20 2>&1 | grep -v cost-only | sed "s/^/NEW  /"
} > gpurun_out/r2i_times.txt 2>&1
grep -v "^+" gpurun_out/r2i_times.txt | cut -c1-235
