"""tools/stage_times.py -- per-kernel device times of the aggregate-verify pipeline for the library in $HBLS_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from harmony_b200 import bls, workload as wl

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
    if os.environ.get("HBLS_TOTAL_ONLY"):
        st = torch.cuda.Stream()
        def run():
            rc = L.hbls_aggregate_verify_batch_device(com.h, B, d_bm.data_ptr(), 32, d_sig.data_ptr(), d_msg.data_ptr(), 48, d_res.data_ptr(), st.cuda_stream)
            assert rc == 0
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): run()
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"TOTAL overlap={os.environ.get('HBLS_OVERLAP', '1')} B={B} ok={int(d_res.sum().item()) == B} {ms:8.2f} ms/step  sigs/s={nsig / ms * 1e3:.3e}", flush=True)
        return
    bls.StageTimingEnable(True)
    acc = np.zeros(6)
    for r in range(reps + 1):
        rc = L.hbls_aggregate_verify_batch_device(com.h, B, d_bm.data_ptr(), 32, d_sig.data_ptr(), d_msg.data_ptr(), 48, d_res.data_ptr(), None)
        assert rc == 0
        st = np.array(bls.StageTimingGet())
        if r > 0: acc += st
    torch.cuda.synchronize()
    good = int(d_res.sum().item()) == B
    acc /= reps
    tot = acc.sum()
    print(f"{os.path.basename(os.environ.get('HBLS_LIB', 'default')):18s} B={B} ok={good} total={tot:8.2f} ms  rounds/s={B / tot * 1e3:9.0f} sigs/s={nsig / tot * 1e3:.3e} | " +
          " ".join(f"{n[2:8]}={m:.2f}" for n, m in zip(bls.STAGE_NAMES, acc)), flush=True)

if __name__ == "__main__":
    main()
