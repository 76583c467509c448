run() { HBLS_TPSM_SPLIT=$3 HBLS_SPLIT=1 HBLS_LIB=$PWD/variants_$1.so timeout 120 python tools/stage_times.py $2 1 2>&1 | tail -1; }
run l1 75776 512; run l2 75776 512; run l3 75776 512; run l3b1024 75776 1024; run l3b1024 151552 1024
