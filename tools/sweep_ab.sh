run() { HBLS_SPLIT=$2 HBLS_LIB=$PWD/variants_$1.so timeout 120 python tools/stage_times.py 75776 2 2>&1 | tail -1 | sed "s/^/split=$2 /"; }
run c1 1; run c2 1; run c1 0; run c2 0; run c1 1; run c2 1
