set -x
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"
tail -1 gpurun_out/bench15.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench15_ref.json 2> gpurun_out/bench15_ref.err; echo "ref rc=$?"
tail -1 gpurun_out/bench15_ref.json
