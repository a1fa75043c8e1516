set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 5 > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"
tail -1 gpurun_out/bench14.json | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k1_lse_tma|k2_lattice|k3_grad_tma' --launch-skip 9 --launch-count 3 -o gpurun_out/r1g_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/ | tail -8
