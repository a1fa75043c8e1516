# K3's tile counter: requests in flight per producer (MRNNT_GRAB_DEPTH 4 / 8 / 16 / 32), c2 in the stream.
# Needs lib/g<depth>_libmonotonic_rnnt.bin (build with extra_flags=['-DMRNNT_GRAB_DEPTH=<depth>'], copied aside).
L=monotonic-rnnt_b200/lib
for d in 8 4 3 2 1; do
  cp $L/g${d}_libmonotonic_rnnt.bin $L/libmonotonic_rnnt.so
  echo "== grab depth $d"
  timeout 100 python tools/kernel_times.py c2 --iters 20 2>&1 | grep -v cost-only
done
cp $L/g8_libmonotonic_rnnt.bin $L/libmonotonic_rnnt.so
