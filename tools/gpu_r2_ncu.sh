# Round 2: ncu evidence for the current binary.  Per named workload: the launch list of bench.py (gpu__time_duration, every
# launch) and one --set full capture of K1, K2, K3 of the first timed step.  One GPU.  Usage: bash tools/gpu_r2_ncu.sh "c2 c3 c5 c4"
set -x
mkdir -p gpurun_out
for wl in ${1:-c2}; do
  ARGS="--workload $wl --steps 3 --warmup 3 --blocks 1 --no-cpu-baseline --no-e2e"
  timeout 300 python bench.py $ARGS > gpurun_out/r2_ncu_plain_$wl.json 2> gpurun_out/r2_ncu_plain_$wl.err && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$wl.csv \
      python bench.py $ARGS > gpurun_out/r2_ncu_l_$wl.log 2>&1
  echo "ncu launches $wl rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_lse_tma|k2_lattice|k3_grad_tma' --launch-skip 9 \
      --launch-count 3 -o gpurun_out/r2_full_$wl python bench.py $ARGS > gpurun_out/r2_ncu_f_$wl.log 2>&1
  echo "ncu full $wl rc=$?"
  # the summary is made on the box (gpurun_out/ travels back only below 64 MiB in all); the capture itself is kept for c2 only
  python tools/ncu_summary.py gpurun_out/r2_full_$wl.ncu-rep > gpurun_out/r2_ncu_full_$wl.txt 2>&1
  python tools/ncu_launch_summary.py gpurun_out/r2_launches_$wl.csv > gpurun_out/r2_ncu_launches_${wl}_summary.txt 2>&1
  [ "$wl" = "c2" ] || rm -f gpurun_out/r2_full_$wl.ncu-rep
done
ls -la gpurun_out/ | grep r2_ | tail -20
