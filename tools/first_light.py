"""tools/first_light.py -- quick on-GPU timing of the aggregate-verify pipeline (device-resident inputs)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from harmony_b200 import bls, workload as wl

def main():
    bls.Init()
    L = bls.lib()
    for it in (2048, 8192):
        print("probe MAC32/s", it, bls.ProbeMac32PerS(it) / 1e12, "T")
    n = 250
    sks = [wl.seeded_sk("c2", i) for i in range(n)]
    pks = bls.GetPublicKeyBatch(b"".join(wl.sk_bytes(k) for k in sks))
    com = bls.Committee([pks[48 * i:48 * i + 48] for i in range(n)])
    bms = [wl.bitmap_with_k("fl", j, n, [167, 200, 250][j % 3]) for j in range(16)]
    agg = [wl.sk_bytes(wl.round_signer_sum(sks, bm)) for bm in bms]
    for B in [int(x) for x in (sys.argv[1:] or ["1024", "8192", "32768"])]:
        msgs = b"".join(wl.commit_payload("fl", j) for j in range(B))
        t = time.time()
        sigs, ok = bls.SignHashBatch(b"".join(agg[j % 16] for j in range(B)), msgs, 48)
        t_sign = time.time() - t
        bitmaps = b"".join(bms[j % 16] for j in range(B))
        d_bm = torch.frombuffer(bytearray(bitmaps), dtype=torch.uint8).cuda()
        d_sig = torch.frombuffer(bytearray(sigs), dtype=torch.uint8).cuda()
        d_msg = torch.frombuffer(bytearray(msgs), dtype=torch.uint8).cuda()
        d_res = torch.zeros(B, dtype=torch.uint8, device="cuda")
        s = torch.cuda.current_stream()
        def run():
            rc = L.hbls_aggregate_verify_batch_device(com.h, B, d_bm.data_ptr(), com.blen(), d_sig.data_ptr(), d_msg.data_ptr(), 48, d_res.data_ptr(), s.cuda_stream)
            assert rc == 0, rc
        run(); torch.cuda.synchronize()
        assert int(d_res.sum()) == B, int(d_res.sum())
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        nsig = sum(bin(int.from_bytes(bms[j % 16], "little")).count("1") for j in range(B))
        print(f"B={B} verify {ms:.2f} ms  rounds/s={B / ms * 1e3:.0f}  sigs/s={nsig / ms * 1e3:.3e}  (sign batch {t_sign:.2f}s)")
        t = time.time(); res = com.AggregateVerifyBatch(bitmaps, sigs, msgs, 48); dt = time.time() - t
        print(f"   e2e host API {dt * 1e3:.2f} ms ok={sum(res)}")

if __name__ == "__main__":
    main()
