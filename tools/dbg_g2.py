import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from harmony_b200 import bls
bls.Init(device=0)
L = bls.lib()
sv = json.load(open(os.path.join(ROOT, "tests/golden/ref_fixtures.json")))["sig_vectors"][0]
out = ctypes.create_string_buffer(512)
rc = L.hbls_debug_g2(bytes.fromhex(sv["sig"]), out)
r = out.raw
print(os.path.basename(os.environ.get("HBLS_LIB", "default")), "rc", rc, "flags", list(r[:8]))
for i, n in enumerate("ABCDP"):
    print(" ", n, r[32 + 96 * i: 32 + 96 * i + 96].hex()[:40], "...", r[32 + 96 * i + 90: 32 + 96 * i + 96].hex())
