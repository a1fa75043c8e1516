# Round 2, call F: where do c2's +4 us in K2 and +4 us in K3 (round-1 binary -> current) come from?  Variants: who zeroes the
# dead rows (K2's fill / K3's consumers), dynamic tiles on / off; the lattice kernel's own time line (k2_probe); then the
# tests that failed in call E and the unaligned-vocabulary shapes again (9-vector K1 variant, zero fill on unaligned rows).
set -x
mkdir -p gpurun_out
{
for rep in 1 2; do
  (cd tools/_r1 && timeout 200 python tools/kernel_times.py c2 --iters 30 --zero -1,0 --dyn -1,0 2>&1 | grep -v cost-only | sed 's/^/R1  /')
  timeout 200 python tools/kernel_times.py c2 --iters 30 --zero -1,0 --dyn -1,0 2>&1 | grep -v cost-only | sed 's/^/NEW /'
done
echo "== k2_probe r1"; ./tools/_r1/tools/k2_probe 150 40 32 4 1 2 1000
echo "== k2_probe new"; ./tools/k2_probe 150 40 32 4 1 2 1000
} > gpurun_out/r2f_ab.txt 2>&1
cut -c1-230 gpurun_out/r2f_ab.txt
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_shard.py tests/test_gpu_unaligned.py tests/test_gpu_parity.py tests/test_gpu_concurrent.py tests/test_gpu_fuzz.py -m gpu -q --maxfail=10 > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2f_pytest.txt
timeout 400 python tools/kernel_times.py c2 c2v1025 c4 c4v5001 --iters 10 > gpurun_out/r2f_unaligned_times.txt 2>&1; echo "kt rc=$?"
grep -v cost-only gpurun_out/r2f_unaligned_times.txt | cut -c1-260
