# Round-end pass on one B200: tests, smoke, both bench arms, ncu launch list + full capture (r1i).
set -x
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench rc=$?"
tail -1 gpurun_out/bench21.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench21_ref.json 2> gpurun_out/bench21_ref.err; echo "ref rc=$?"
tail -1 gpurun_out/bench21_ref.json
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k1_lse_tma|k2_lattice|k3_grad_tma' --launch-skip 9 --launch-count 3 -o gpurun_out/r1i_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
echo "ncu full rc=$?"
timeout 200 python tools/kernel_times.py c2 2>&1 | tail -6
ls -la gpurun_out/ | tail -12
