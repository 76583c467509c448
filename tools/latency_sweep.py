"""tools/latency_sweep.py -- single-call latency of hbls_aggregate_verify_batch at small batch sizes (host buffers, blocking call),
per-kernel stage times included: the numbers behind `single_round_latency_ms` / `batch_sweep_e2e` of bench.py."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from harmony_b200 import bls, workload as wl

def main():
    sizes = [int(x) for x in sys.argv[1:]] or [1, 16, 64, 256, 1024, 4096, 16384]
    bls.Init(device=0)
    sks = bench.make_committee_sks()
    pks_blob = bls.GetPublicKeyBatch(b"".join(wl.sk_bytes(k) for k in sks))
    com = bls.Committee([pks_blob[48 * i:48 * i + 48] for i in range(250)])
    B = max(sizes)
    bitmaps, agg_sk, msgs, nsig = bench.make_rounds(sks, B, seed=11)
    sigs, ok = bls.SignHashBatch(agg_sk, msgs, 48)
    spr = nsig / B
    for coop in (1, 0):
        old = (bls.GetParam("coop_max"), bls.GetParam("rlc_min"))
        if not coop: bls.SetParam("coop_max", 0); bls.SetParam("rlc_min", 1024)      # round-1 behaviour: lane pair per round, batched groups from 1 024 rounds
        for n in sizes:
            lat = []
            bls.StageTimingEnable(True)
            for _ in range(4):
                t0 = time.perf_counter()
                res = com.AggregateVerifyBatch(bitmaps[:32 * n], sigs[:96 * n], msgs[:48 * n], 48)
                lat.append((time.perf_counter() - t0) * 1e3)
            st = bls.StageTimingGet(); bls.StageTimingEnable(False)
            assert res == b"\x01" * n
            ms = float(np.median(lat[1:]))
            info = bls.LastBatchInfo()
            print(f"latency={'warp-per-round' if coop else 'lane-pair/round1'} B={n:6d} {ms:9.3f} ms  {n * spr / ms * 1e3:12.4e} sigs/s  mode={info['mode']} cta={info['cta_threads']} stages(ms)=" +
                  " ".join(f"{x:.3f}" for x in st), flush=True)
        bls.SetParam("coop_max", old[0]); bls.SetParam("rlc_min", old[1])

if __name__ == "__main__":
    main()
