timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=line 2>&1 | tail -2
for t in 256 192 320 384 512; do echo "TPSM=$t"; HBLS_TPSM=$t timeout 90 python tools/stage_times.py 303104 2 2>&1 | tail -1; done
