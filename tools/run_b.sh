timeout 900 python -m pytest tests/test_cli_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/cli_wall.py 2>&1 | tail -4
