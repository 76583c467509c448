run() { HBLS_LIB=$PWD/variants_$1.so timeout 120 python tools/stage_times.py 75776 2 2>&1 | tail -1; }
run i0; run i1; run i1r; run i0r; run i0
HBLS_LIB=$PWD/variants_i0.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
