timeout 600 python tools/cli_wall.py 2>&1 | tail -5
