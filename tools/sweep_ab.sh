export HBLS_LIB=$PWD/harmony_b200/lib/libhbls_lst.so
timeout 100 python tools/stage_times.py 303104 2 2>&1 | tail -1
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=line 2>&1 | tail -2
HBLS_TCTA=512 timeout 60 python tools/stage_times.py 303104 2 2>&1 | tail -1
HBLS_TCTA=256 timeout 60 python tools/stage_times.py 303104 2 2>&1 | tail -1
