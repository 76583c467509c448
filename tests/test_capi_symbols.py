"""CPU test: libhbls.so loads without a GPU and exports every symbol include/hbls.h declares; without a device the
library refuses to work (no CPU fallback)."""
import ctypes, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def _declared():
    src = open(os.path.join(ROOT, "include", "hbls.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:bls|hbls_)[A-Za-z0-9_]+)\s*\(", src)))

def test_header_symbols_exported():
    from harmony_b200 import build
    path = build.build_cuda()
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/hbls.h but not exported"

def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from harmony_b200 import bls
    with pytest.raises(bls.HblsError):
        bls.Init()
    L = bls.lib()
    out = ctypes.create_string_buffer(96)
    assert L.hbls_map_to_g2(b"abc", 3, out) == bls.ERR_CUDA
    assert L.hbls_kernel_launch_count() == 0

def test_struct_sizes_match_herumi():
    from harmony_b200 import bls
    assert ctypes.sizeof(bls._Sec) == 32 and ctypes.sizeof(bls._Pub) == 144 and ctypes.sizeof(bls._Sig) == 288

def test_cpp_host_mirror_cpu_logic():
    """Pure-host parts of the C++ mirror (payload bytes, sig||bitmap parsing, quorum threshold, LRU) on the CPU; the binary
    links libhbls.so but never initialises it."""
    import subprocess
    from harmony_b200 import build
    build.build_cuda()
    hd = os.path.join(ROOT, "harmony_b200", "host"); exe = os.path.join(build.LIBDIR, "hbls_host_cputest")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(hd, "hbls_host_cputest.cpp"),
                           "-L" + build.LIBDIR, "-lhbls", "-Wl,-rpath,$ORIGIN", "-pthread", "-ldl", "-lrt"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr
