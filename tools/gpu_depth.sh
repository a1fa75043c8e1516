# zero-fill pipeline depth 2 / 3 / 4: kernel times on c2, c3, c5 and a parity run of the zero-fill tests for each.
# Needs lib/d3_libmonotonic_rnnt.bin and d4_...: monotonic_rnnt_b200.build.build(force=True,
# extra_flags=['-DMRNNT_ZERO_FILL_DEPTH=3']) copied aside, likewise 4, then the default build again.
set -x
L=monotonic-rnnt_b200/lib
cp $L/libmonotonic_rnnt.so $L/d2_libmonotonic_rnnt.bin
for d in 2 3 4; do
  cp $L/d${d}_libmonotonic_rnnt.bin $L/libmonotonic_rnnt.so
  echo "== depth $d"
  timeout 200 python tools/kernel_times.py c2 c3 c5 --iters 20 2>&1 | grep -v cost-only
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "zeroing or automatic_choice or masked" 2>&1 | tail -1
done
cp $L/d2_libmonotonic_rnnt.bin $L/libmonotonic_rnnt.so
