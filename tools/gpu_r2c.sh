# Round 2, call C: the L2 probe again with part of the L2 set aside for evict_last lines (cudaLimitPersistingL2CacheSize).
set -x
mkdir -p gpurun_out
for mb in 48 80 126; do
  L2P_QUICK=1 L2P_PERSIST_MB=$mb timeout 120 ./tools/l2_probe > gpurun_out/r2c_l2_probe_persist$mb.txt 2>&1; echo "rc=$?"
  cat gpurun_out/r2c_l2_probe_persist$mb.txt
done
