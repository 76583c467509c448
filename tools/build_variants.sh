# tools/build_variants.sh -- build the experimental device-code variants next to the default library (run HERE, no GPU needed);
# the .so files travel to the GPU box with gpurun and are timed there with
#   gpurun -- 'bash tools/sweep_ab.sh harmony_b200/lib/variants/libhbls_*.so'
# Every variant is kept correct on the CPU by tests/test_emu_kernels.py (same kernels on host threads against the oracle).
set -e
cd "$(dirname "$0")/.."
mkdir -p harmony_b200/lib/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Iinclude"
build() { name=$1; shift; echo "building $name: $*"; $NVCC $FLAGS "$@" -o harmony_b200/lib/variants/libhbls_$name.so harmony_b200/csrc/hbls.cu & }
build noinv -DHB_BATCH_INV=0                          # one inversion per item (round-1 behaviour)
build batchinv4 -DHB_BATCH_INV=1 -DHB_BATCH_K=4      # one shared Fp inversion per 4 items of a persistent thread (hash-to-G2: 2 per item -> 2 per 4)
build batchinv8 -DHB_BATCH_INV=1 -DHB_BATCH_K=8
for extra in "$@"; do build "x$(echo "$extra" | tr -c 'A-Za-z0-9' '_')" $extra; done     # ad-hoc: tools/build_variants.sh "-DHB_X=1 -DHB_Y=2"
wait
ls -la harmony_b200/lib/variants
