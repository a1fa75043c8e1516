# Round 2, call A: L2 residency probe (plain + DRAM bytes of the second reads under ncu) and the per-kernel times of the
# round-1 binary on the four named shapes (the baseline this round's changes are measured against).
set -x
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
timeout 300 ./tools/l2_probe > gpurun_out/r2a_l2_probe.txt 2>&1; echo "probe rc=$?"
tail -60 gpurun_out/r2a_l2_probe.txt
timeout 300 python tools/kernel_times.py c2 c3 c5 c4 --iters 10 > gpurun_out/r2a_kernel_times.txt 2>&1; echo "kt rc=$?"
cat gpurun_out/r2a_kernel_times.txt
L2P_QUICK=1 timeout 120 ./tools/l2_probe > gpurun_out/r2a_l2_probe_quick.txt 2>&1 && \
L2P_QUICK=1 timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:reread --csv --log-file gpurun_out/r2a_l2_probe_ncu.csv ./tools/l2_probe > gpurun_out/r2a_l2_probe_ncu.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/r2a_l2_probe_ncu.csv
