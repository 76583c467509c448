python - <<'PY'
import sys; sys.path.insert(0, ".")
from harmony_b200 import bls
bls.Init(device=0)
print("selftest_split mismatches:", bls.SelfTestSplit(8))
PY
HBLS_SPLIT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
run() { HBLS_SPLIT=$1 HBLS_TPSM_SPLIT=$2 timeout 120 python tools/stage_times.py $3 1 2>&1 | tail -1 | sed "s/^/split=$1 tpsm_split=$2 /"; }
run 0 512 37888; run 1 256 37888; run 1 512 37888; run 1 768 37888; run 1 1024 37888; run 1 512 75776; run 1 1024 75776; run 0 512 75776
