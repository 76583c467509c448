"""tools/cpu_view_change_baseline.py -- how long the reference's SEQUENTIAL view-change handlers take on one host core, for the storm
bench.py times on the GPU (other_configs.c4_view_change_storm.view_change_handlers): 250 VIEWCHANGE messages of a 250-validator
committee, 167 with an embedded PREPARED proof.  Uses the CPU oracle through the sequential restatement of the handlers in
tests/test_consensus.py (test infrastructure; one oracle call per cgo call of the reference).  CPU only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from harmony_b200 import workload as wl, consensus as cs
from test_consensus import RefViewChange, le64

def main():
    orc = oracle_lib.load()
    n, vid = 250, 7
    sks = [wl.seeded_sk("bench-vc", i) for i in range(n)]
    pks = [orc.get_public_key(wl.sk_bytes(k)) for k in sks]
    bh = wl.seeded_bytes("bench-vc/hash", 0, 32); bmq = wl.bitmap_with_k("bench-vc/prep", 0, n, wl.quorum_k(n))
    payload = bh + orc.sign_hash(wl.sk_bytes(wl.round_signer_sum(sks, bmq)), bh) + bmq
    msgs = []
    for i in range(n):
        m1 = i % 3 != 2
        msgs.append(cs.FBFTMessage(ViewID=vid, BlockNum=1, SenderPubkey=pks[i], LeaderPubkey=pks[0], Payload=payload if m1 else b"", Block=b"\xc0" if m1 else b"",
                                   ViewchangeSig=orc.sign_hash(wl.sk_bytes(sks[i]), payload if m1 else cs.NIL), ViewidSig=orc.sign_hash(wl.sk_bytes(sks[i]), le64(vid))))
    ts = []
    for _ in range(3):
        ref = RefViewChange(orc, pks)
        t0 = time.perf_counter(); res = [ref.on_view_change(m) for m in msgs]; ts.append((time.perf_counter() - t0) * 1e3)
        assert res == [None] * n
    ms = sorted(ts)[1]
    print(f"sequential view-change handlers, restated CPU path (oracle/hbls_oracle.c), 1 thread: {n} messages (167 M1 + 83 M2, 667 signature checks) "
          f"in {ms:.0f} ms = {ms / n:.2f} ms per message; host: {os.cpu_count()} logical CPUs")

if __name__ == "__main__":
    main()
