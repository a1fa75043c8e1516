# The other named shapes of BASELINE.json through bench.py's contract (N = 1), plus K3 static/dynamic alone under ncu.
set -x
mkdir -p gpurun_out
for w in c3 c5 c4; do
  timeout 500 python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/bench18_$w.json 2> gpurun_out/bench18_$w.err; echo "$w rc=$?"
  tail -3 gpurun_out/bench18_$w.err
  tail -1 gpurun_out/bench18_$w.json | cut -c1-400
done
timeout 200 python tools/kernel_times.py c2 --dyn 0,1 --iters 10 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k3_grad' -c 60 --csv --log-file gpurun_out/k3_dyn_alone.csv python tools/kernel_times.py c2 --dyn 0,1 --iters 5 > gpurun_out/k3_dyn_alone.log 2>&1
echo "ncu rc=$?"
