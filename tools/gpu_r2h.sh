# Round 2, call H: the gradient kernel's inner loop, four builds on one box against the round-1 binary (c2, c3), and the lattice
# kernel's phase A under the probe's stamps.
#   p0k1: patches in the vector loop behind branches, packed FFMA2/FADD2     p0k0: the same, scalar FFMA + FADD
#   p2k1: patches in the loop, predicated (precomputed)                       p1k1: patched elements rewritten after the loop
set -x
mkdir -p gpurun_out
L=monotonic-rnnt_b200/lib
cp $L/libmonotonic_rnnt.so $L/keep_libmonotonic_rnnt.bin
{
for rep in 1 2; do
  (cd tools/_r1 && timeout 200 python tools/kernel_times.py c2 c3 --iters 30 2>&1 | grep -v cost-only | sed 's/^/R1   /')
  for v in p0k1 p0k0 p2k1 p1k1; do
    cp $L/${v}_libmonotonic_rnnt.bin $L/libmonotonic_rnnt.so
    timeout 200 python tools/kernel_times.py c2 c3 --iters 30 2>&1 | grep -v cost-only | sed "s/^/$v /"
  done
done
cp $L/keep_libmonotonic_rnnt.bin $L/libmonotonic_rnnt.so
echo "== k2_probe r1"; ./tools/_r1/tools/k2_probe 150 40 32 4 1 2 1000 | grep -v "per chunk"
echo "== k2_probe new"; ./tools/k2_probe 150 40 32 4 1 2 1000 | grep -v "per chunk"
echo "== k2_probe new, no zero fill"; ./tools/k2_probe 150 40 32 4 1 0 1000 | grep -v "per chunk"
echo "== k2_probe r1, no zero fill"; ./tools/_r1/tools/k2_probe 150 40 32 4 1 0 1000 | grep -v "per chunk"
} > gpurun_out/r2h_ab.txt 2>&1
grep -v "^+" gpurun_out/r2h_ab.txt | cut -c1-235
