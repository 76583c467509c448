# tools/sweep_ab.sh -- A/B harness used during tuning (run on the GPU box): parity tests on the default library, then per-kernel
# stage times of the default library and of every variant library given as argument (built with extra -D flags).
#   gpurun -- 'bash tools/sweep_ab.sh harmony_b200/lib/libhbls_variant.so'
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=line 2>&1 | tail -2
timeout 120 python tools/stage_times.py ${HBLS_SWEEP_ROUNDS:-303104} 2 2>&1 | tail -1
for lib in "$@"; do HBLS_LIB=$PWD/$lib timeout 120 python tools/stage_times.py ${HBLS_SWEEP_ROUNDS:-303104} 2 2>&1 | tail -1; done
