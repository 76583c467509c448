timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=line 2>&1 | tail -2
timeout 90 python tools/stage_times.py 303104 2 2>&1 | tail -1
timeout 90 python tools/stage_times.py 75776 2 2>&1 | tail -1
timeout 60 python tools/stage_times.py 8192 2 2>&1 | tail -1
