cd /root/repo
timeout 400 python -m pytest tests/test_multigpu.py -x -q -k "reference_host" 2>&1 | tail -15
