"""tools/profile_target.py -- short target for ncu: warm-up pass + one aggregate-verify pipeline pass at B rounds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from harmony_b200 import bls, workload as wl

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    bls.Init(device=0)
    L = bls.lib()
    sks = bench.make_committee_sks()
    pks_blob = bls.GetPublicKeyBatch(b"".join(wl.sk_bytes(k) for k in sks))
    com = bls.Committee([pks_blob[48 * i:48 * i + 48] for i in range(250)])
    bitmaps, agg_sk, msgs, nsig = bench.make_rounds(sks, B, seed=2024)
    sigs, ok = bls.SignHashBatch(agg_sk, msgs, 48)
    dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    d_bm, d_sig, d_msg = dev(bitmaps), dev(sigs), dev(msgs)
    d_res = torch.zeros(B, dtype=torch.uint8, device="cuda")
    for _ in range(passes):
        rc = L.hbls_aggregate_verify_batch_device(com.h, B, d_bm.data_ptr(), 32, d_sig.data_ptr(), d_msg.data_ptr(), 48, d_res.data_ptr(), None)
        assert rc == 0
    torch.cuda.synchronize()
    assert int(d_res.sum().item()) == B
    print("profile target done", B)

if __name__ == "__main__":
    main()
