timeout 600 python -m pytest tests/test_gpu_parity.py -k "one_cluster_counting or golden" -x -q 2>&1 | tail -15
