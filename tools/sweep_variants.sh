run() { HBLS_FUSE=$2 HBLS_TPSM=$3 HBLS_LIB=$PWD/variants_$1.so timeout 120 python tools/stage_times.py 37888 1 2>&1 | tail -1 | sed "s/^/fuse=$2 tpsm=$3 /"; }
run f0 0 512; run f0 1 256; run f0 1 512; run f2 0 256; run f2 0 512; run f2 1 256; run f2 1 512; run f0lb8 1 512; run f0lb8 0 512; run f2lb6 1 384; run f2lb6 0 384
