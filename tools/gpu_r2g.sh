# Round 2, call G: A/B against the round-1 binary again (K2: one block per launched CTA, claimed; K3: the two patched
# elements out of the vector loop), the lattice kernel's time line, the GPU suite (without the drop-in build), unaligned shapes.
set -x
mkdir -p gpurun_out
{
for rep in 1 2; do
  (cd tools/_r1 && timeout 200 python tools/kernel_times.py c2 c3 --iters 30 2>&1 | grep -v cost-only | sed 's/^/R1  /')
  timeout 200 python tools/kernel_times.py c2 c3 --iters 30 2>&1 | grep -v cost-only | sed 's/^/NEW /'
done
(cd tools/_r1 && timeout 200 python tools/kernel_times.py c2 --iters 30 --zero=0 --dyn=-1,0 2>&1 | grep -v cost-only | sed 's/^/R1  /')
timeout 200 python tools/kernel_times.py c2 --iters 30 --zero=0 --dyn=-1,0 2>&1 | grep -v cost-only | sed 's/^/NEW /'
echo "== k2_probe r1"; ./tools/_r1/tools/k2_probe 150 40 32 4 1 2 1000 | grep -v "per chunk"
echo "== k2_probe new"; ./tools/k2_probe 150 40 32 4 1 2 1000 | grep -v "per chunk"
} > gpurun_out/r2g_ab.txt 2>&1
grep -v "^+" gpurun_out/r2g_ab.txt | cut -c1-230
timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 --ignore=tests/test_gpu_dropin.py > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2g_pytest.txt
timeout 400 python tools/kernel_times.py c2v1025 c4v5001 c5 --iters 10 > gpurun_out/r2g_unaligned_times.txt 2>&1; echo "kt rc=$?"
grep -v cost-only gpurun_out/r2g_unaligned_times.txt | cut -c1-260
