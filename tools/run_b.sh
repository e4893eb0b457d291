python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_cli_gpu.py -x -q -k "golden" 2>&1 | tail -2
