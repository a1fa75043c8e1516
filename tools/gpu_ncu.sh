# ncu launch list + full capture of the three kernels for the current binary (bench.py has already exited 0 on it).
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1j_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:'k1_lse_tma|k2_lattice|k3_grad_tma' --launch-skip 9 --launch-count 3 -o gpurun_out/r1j_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
echo "ncu full rc=$?"
