set -x
timeout 300 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_early_return.py tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
