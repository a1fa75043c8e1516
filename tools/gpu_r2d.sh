# Round 2, call D: the whole GPU suite on the new binary (block hand-out in the lattice kernel, centred gradient
# coefficients, sticky peer failure), smoke, the c2 bench line and the per-kernel times of the four shapes.
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 --durations=12 > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2d_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2d_bench_c2.json; tail -5 gpurun_out/r2d_bench_c2.err
timeout 300 python tools/kernel_times.py c2 c3 c5 c4 --iters 10 > gpurun_out/r2d_kernel_times.txt 2>&1; echo "kt rc=$?"
cat gpurun_out/r2d_kernel_times.txt
