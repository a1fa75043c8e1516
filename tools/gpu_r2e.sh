# Round 2, call E: same-box A/B of the round-1 binary (tools/_r1) against the current one, the whole GPU suite (exact patch
# exponents, unaligned vocabularies), and the unaligned-vocabulary shapes next to their aligned twins.
set -x
mkdir -p gpurun_out
for rep in 1 2; do
  (cd tools/_r1 && timeout 200 python tools/kernel_times.py c2 c3 --iters 30 2>&1 | grep -v cost-only | sed 's/^/R1  /')
  timeout 200 python tools/kernel_times.py c2 c3 --iters 30 2>&1 | grep -v cost-only | sed 's/^/NEW /'
done > gpurun_out/r2e_ab.txt 2>&1
cat gpurun_out/r2e_ab.txt
timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 --durations=8 > gpurun_out/r2e_pytest.txt 2>&1; echo "pytest rc=$?"
grep -n "^\[" gpurun_out/r2e_pytest.txt | cut -c1-220; tail -30 gpurun_out/r2e_pytest.txt
timeout 400 python tools/kernel_times.py c2 c2v1025 c4 c4v5001 --iters 10 > gpurun_out/r2e_unaligned_times.txt 2>&1; echo "kt rc=$?"
grep -v cost-only gpurun_out/r2e_unaligned_times.txt
