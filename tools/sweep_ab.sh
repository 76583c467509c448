timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
run() { HBLS_TPSM=$2 HBLS_LIB=$PWD/variants_$1.so timeout 60 python tools/stage_times.py 75776 2 2>&1 | tail -1 | sed "s/^/tpsm=$2 /"; }
run base 256; run base 512; run ofp 256; run ofp 512
