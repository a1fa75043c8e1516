# Round 2, call B: L2 residency probe, single persistent kernel (plain, then DRAM bytes per variant under ncu).
set -x
mkdir -p gpurun_out
timeout 300 ./tools/l2_probe > gpurun_out/r2b_l2_probe.txt 2>&1; echo "probe rc=$?"
cat gpurun_out/r2b_l2_probe.txt
L2P_QUICK=1 timeout 120 ./tools/l2_probe > gpurun_out/r2b_l2_probe_quick.txt 2>&1 && \
L2P_QUICK=1 timeout 600 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:pipeline --csv --log-file gpurun_out/r2b_l2_probe_ncu.csv ./tools/l2_probe > gpurun_out/r2b_l2_probe_ncu.log 2>&1
echo "ncu rc=$?"
