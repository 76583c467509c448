"""View-change storm and consensus-message signatures (harmony_b200/consensus.py) against a SEQUENTIAL restatement of the reference's
handlers that calls the CPU oracle once per VerifyHash -- i.e. what consensus/checks.go:139-193, consensus/view_change_construct.go:
154-375 and consensus/view_change.go:445-500 do today, one cgo call at a time.  The batched mirror must give every message the same
error and leave the same state.

CPU tests run the mirror over an oracle-backed backend (host logic only); the `gpu` tests run it over libhbls.so (two device calls
per storm) and compare with the same sequential restatement.
"""
import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harmony_b200 import workload as wl
from harmony_b200 import consensus as cs
from harmony_b200.consensus import FBFTMessage, NIL, ValidPayloadLength

BAD_SIG = b"\xff" * 96          # x >= p: does not decode
BAD_KEY = b"\xff" * 48

def le64(v): return int(v).to_bytes(8, "little")
def popcount(bm, n): return sum(1 for i in range(n) if bm[i >> 3] >> (i & 7) & 1)

# ------------------------------------------------------------------ sequential restatement of the reference handlers (oracle = libbls)
class RefViewChange:
    """Transcription of ProcessViewChangeMsg & co. with one oracle call where the reference makes one cgo call."""
    def __init__(self, orc, members):
        self.o = orc; self.members = list(members); self.n = len(members); self.blen = (self.n + 7) >> 3
        self.och = orc.committee(self.members)
        self.index = {m: i for i, m in enumerate(self.members)}
        self.bhpSigs, self.nilSigs, self.viewIDSigs = {}, {}, {}
        self.bhpBitmap, self.nilBitmap, self.viewIDBitmap = {}, {}, {}
        self.m1Payload = b""
        self.quorum = 2 * self.n // 3 + 1
    def _setkey(self, table, vid, pk):
        bm = table.setdefault(vid, bytearray(self.blen)); i = self.index.get(pk)
        if i is not None: bm[i >> 3] |= 1 << (i & 7)
    def on_view_change(self, m):
        o = self.o
        # ParseViewChangeMessage (view_change_msg.go:159-179)
        if len(m.SenderPubkey) != 48 or not o.pk_check(m.SenderPubkey): return cs.errKeyDeserialize
        if len(m.ViewchangeSig) != 96 or not o.sig_check(m.ViewchangeSig): return cs.errSigDeserialize
        if len(m.ViewidSig) != 96 or not o.sig_check(m.ViewidSig): return cs.errSigDeserialize
        # onViewChangeSanityCheck (checks.go:184-191)
        if not o.verify_hash(m.ViewidSig, m.SenderPubkey, le64(m.ViewID)): return cs.errViewIDSig
        # ProcessViewChangeMsg (view_change_construct.go:237-375)
        sender = m.SenderPubkey.hex()
        if sender in self.viewIDSigs.get(m.ViewID, {}): return cs.errDupM3
        if len(m.Payload) >= ValidPayloadLength and len(m.Block) != 0:
            if sender in self.bhpSigs.get(m.ViewID, {}): return cs.errDupM1
            if not o.verify_hash(m.ViewchangeSig, m.SenderPubkey, m.Payload): return cs.errVerifyM1
            blockHash = m.Payload[:32]; body = m.Payload[32:]
            sig, bitmap = body[:96], body[96:]
            if not o.sig_check(sig): return cs.errMultiSigDeserialize                        # sig.go:39-43
            if len(bitmap) != self.blen: return cs.errSetMask                                # sig.go:44-48
            if popcount(bitmap, self.n) < self.quorum: return cs.errNoQuorum
            if o.committee_aggregate_verify(self.och, bitmap, sig, blockHash) != 1: return cs.errM1Payload
            self.bhpSigs.setdefault(m.ViewID, {})[sender] = m.ViewchangeSig
            self._setkey(self.bhpBitmap, m.ViewID, m.SenderPubkey)
            self.viewIDSigs.setdefault(m.ViewID, {})[sender] = m.ViewidSig
            self._setkey(self.viewIDBitmap, m.ViewID, m.SenderPubkey)
            if not self.m1Payload: self.m1Payload = m.Payload
            return None
        if sender in self.nilSigs.get(m.ViewID, {}): return cs.errDupM2
        if not o.verify_hash(m.ViewchangeSig, m.SenderPubkey, NIL): return cs.errVerifyM2
        self.nilSigs.setdefault(m.ViewID, {})[sender] = m.ViewchangeSig
        self._setkey(self.nilBitmap, m.ViewID, m.SenderPubkey)
        self.viewIDSigs.setdefault(m.ViewID, {})[sender] = m.ViewidSig
        self._setkey(self.viewIDBitmap, m.ViewID, m.SenderPubkey)
        return None
    def on_new_view(self, m):
        o = self.o
        def mask_of(bm): return bytes(bm) if bm is not None and len(bm) == self.blen else bytes(self.blen)
        has_m3 = bool(m.M3AggSig); has_m2 = bool(m.M2AggSig)
        if has_m3 and (len(m.M3AggSig) != 96 or not o.sig_check(m.M3AggSig)): return cs.errSigDeserialize     # ParseNewViewMessage
        if has_m2 and (len(m.M2AggSig) != 96 or not o.sig_check(m.M2AggSig)): return cs.errSigDeserialize
        if not has_m3 or m.M3Bitmap is None: return cs.errM3Nil
        m3 = mask_of(m.M3Bitmap); m2 = mask_of(m.M2Bitmap) if has_m2 else None
        if o.committee_aggregate_verify(self.och, m3, m.M3AggSig, le64(m.ViewID)) != 1: return cs.errM3Verify
        if has_m2 and o.committee_aggregate_verify(self.och, m2, m.M2AggSig, NIL) != 1: return cs.errM2Verify
        if popcount(m3, self.n) < self.quorum: return cs.errNewViewQuorum
        if m2 is None or popcount(m3, self.n) > popcount(m2, self.n):
            if 32 + 96 > len(m.Payload): return cs.errPayloadLength
            body = m.Payload[32:]
            if not o.sig_check(body[:96]): return cs.errMultiSigDeserialize
            if len(body) - 96 != self.blen: return cs.errSetMask
            if o.committee_aggregate_verify(self.och, body[96:], body[:96], m.Payload[:32]) != 1: return cs.errNewViewM1
        return None

class OracleBackend:
    """consensus.DeviceBackend's interface over the CPU oracle: lets the host logic of the mirror run without a GPU (tests only)."""
    def __init__(self, orc): self.o = orc; self.calls = {"verify_status": 0, "verify_headers": 0, "aggregate_sigs": 0}
    def committee(self, pks): return {"pks": list(pks), "h": self.o.committee(list(pks))}
    def verify_status(self, pks, sigs, msgs):
        self.calls["verify_status"] += 1
        out = bytearray()
        for pk, sg, m in zip(pks, sigs, msgs):
            if not self.o.pk_check(pk): out.append(4)
            elif not self.o.sig_check(sg): out.append(3)
            else: out.append(1 if self.o.verify_hash(sg, pk, m) else 0)
        return bytes(out)
    def verify_headers(self, com, sigs, bitmaps, msgs, quorum):
        self.calls["verify_headers"] += 1
        n = len(com["pks"]); out = bytearray()
        for sg, bm, m in zip(sigs, bitmaps, msgs):
            if not self.o.sig_check(sg): out.append(3)
            elif quorum and popcount(bm, n) < quorum: out.append(2)
            else: out.append(1 if self.o.committee_aggregate_verify(com["h"], bm, sg, m) == 1 else 0)
        return bytes(out)
    def aggregate_sigs(self, sigs):
        self.calls["aggregate_sigs"] += 1
        return self.o.aggregate_sigs(list(sigs))

# ------------------------------------------------------------------ the storm (shared by the CPU and the GPU test)
def build_storm(n, n_msgs, sign, pk_of, viewID=77):
    """n-validator committee; n_msgs VIEWCHANGE messages from distinct senders (2/3 M1 with an embedded PREPARED proof, 1/3 M2)
    with one fault of every kind the handlers distinguish.  sign(sk_int, msg) -> sig96, pk_of(list of sk) -> list of pk48."""
    sks = [wl.seeded_sk("vc", i) for i in range(n)]
    pks = pk_of(sks)
    q = wl.quorum_k(n); blen = (n + 7) >> 3
    blockHash = wl.seeded_bytes("vc/hash", 0, 32)
    def proof(tag, k, bh=blockHash, signed_bh=None):
        bm = wl.bitmap_with_k(tag, 0, n, k)
        return bh + sign(wl.round_signer_sum(sks, bm), signed_bh or bh) + bm
    payload = proof("vc/prep", q)
    assert len(payload) >= ValidPayloadLength + blen and len(payload) > 48
    msgs = []
    for i in range(n_msgs):
        if i % 3 != 2:
            msgs.append(FBFTMessage(ViewID=viewID, BlockNum=9, SenderPubkey=pks[i], LeaderPubkey=pks[0], Payload=payload, Block=b"\xc0rlp",
                                    ViewchangeSig=sign(sks[i], payload), ViewidSig=sign(sks[i], le64(viewID))))
        else:
            msgs.append(FBFTMessage(ViewID=viewID, BlockNum=9, SenderPubkey=pks[i], LeaderPubkey=pks[0],
                                    ViewchangeSig=sign(sks[i], NIL), ViewidSig=sign(sks[i], le64(viewID))))
    f = {}
    def m1(i): assert i % 3 != 2; return msgs[i]
    def m2(i): assert i % 3 == 2; return msgs[i]
    m1(0).ViewidSig = sign(sks[0], le64(viewID + 1)); f[0] = cs.errViewIDSig                       # signed another view
    m1(1).ViewchangeSig = sign(sks[1], b"\x55" * 48 + payload[48:]); f[1] = cs.errVerifyM1         # differs inside the first 48 bytes
    m1(3).Payload = proof("vc/low", q - 1); m1(3).ViewchangeSig = sign(sks[3], m1(3).Payload); f[3] = cs.errNoQuorum
    m1(4).Payload = proof("vc/prep", q, signed_bh=wl.seeded_bytes("vc/hash", 1, 32)); m1(4).ViewchangeSig = sign(sks[4], m1(4).Payload); f[4] = cs.errM1Payload
    m1(6).Payload = blockHash + BAD_SIG + payload[128:]; m1(6).ViewchangeSig = sign(sks[6], m1(6).Payload); f[6] = cs.errMultiSigDeserialize
    m1(7).Payload = payload[:-1]; m1(7).ViewchangeSig = sign(sks[7], m1(7).Payload); f[7] = cs.errSetMask
    m2(8).ViewchangeSig = sign(sks[8], b"\x02"); f[8] = cs.errVerifyM2                             # signed 0x02, not NIL
    m1(9).ViewchangeSig = BAD_SIG; f[9] = cs.errSigDeserialize
    m1(10).SenderPubkey = BAD_KEY; f[10] = cs.errKeyDeserialize
    m2(11).ViewidSig = BAD_SIG; f[11] = cs.errSigDeserialize
    # the hash reads only the first 48 bytes: a signature over payload[:48] verifies against the whole M1 payload
    m1(12).ViewchangeSig = sign(sks[12], payload[:48])
    # M1 payload but no block attached: handled as an M2 message, whose signature must then be over NIL
    msgs[13].Payload = payload; msgs[13].Block = b""; msgs[13].ViewchangeSig = sign(sks[13], payload); f[13] = cs.errVerifyM2
    # repeats: an accepted sender again (M3 duplicate), a rejected sender again with a good message (accepted this time)
    dup = FBFTMessage(**{**msgs[15].__dict__}); msgs.append(dup); f[len(msgs) - 1] = cs.errDupM3
    retry = FBFTMessage(ViewID=viewID, BlockNum=9, SenderPubkey=pks[1], LeaderPubkey=pks[0], Payload=payload, Block=b"\xc0rlp",
                        ViewchangeSig=sign(sks[1], payload), ViewidSig=sign(sks[1], le64(viewID)))
    msgs.append(retry)
    # another view id in the same batch keeps its own tables
    other = FBFTMessage(ViewID=viewID + 5, BlockNum=9, SenderPubkey=pks[15], LeaderPubkey=pks[0],
                        ViewchangeSig=sign(sks[15], NIL), ViewidSig=sign(sks[15], le64(viewID + 5)))
    msgs.append(other)
    return sks, pks, msgs, f, payload

def check_storm(vc, ref, msgs, faults):
    got = vc.ProcessViewChangeMsgs(msgs)
    exp = [ref.on_view_change(m) for m in msgs]
    assert got == exp
    for i, e in faults.items(): assert exp[i] == e, (i, exp[i], e)
    assert sum(1 for e in exp if e is None) == len(msgs) - len(faults)
    for mine, theirs in ((vc.bhpSigs, ref.bhpSigs), (vc.nilSigs, ref.nilSigs), (vc.viewIDSigs, ref.viewIDSigs)): assert mine == theirs
    for mine, theirs in ((vc.bhpBitmap, ref.bhpBitmap), (vc.nilBitmap, ref.nilBitmap), (vc.viewIDBitmap, ref.viewIDBitmap)):
        assert {k: bytes(v) for k, v in mine.items()} == {k: bytes(v) for k, v in theirs.items()}
    assert vc.GetM1Payload() == ref.m1Payload and not vc.IsM1PayloadEmpty()

def new_view_cases(vc, ref, orc, sks, pks, payload, sign, viewID=77):
    """The new leader's NEWVIEW (aggregates of what the storm collected) and what a validator answers to it and to damaged copies."""
    n = len(pks)
    m2sig, m2bm = vc.GetM2Bitmap(viewID); m3sig, m3bm = vc.GetM3Bitmap(viewID)
    assert m2sig == orc.aggregate_sigs(list(ref.nilSigs[viewID].values())) and m2bm == bytes(ref.nilBitmap[viewID])
    assert m3sig == orc.aggregate_sigs(list(ref.viewIDSigs[viewID].values())) and m3bm == bytes(ref.viewIDBitmap[viewID])
    assert vc.GetM2Bitmap(viewID + 99) == (None, None)
    # the storm alone does not reach 2/3: complete M3 with the remaining validators so that the NEWVIEW carries a quorum
    bm3 = bytearray(m3bm); extra = []
    for i in range(n):
        if popcount(bm3, n) >= wl.quorum_k(n): break
        if not bm3[i >> 3] >> (i & 7) & 1: bm3[i >> 3] |= 1 << (i & 7); extra.append(sign(sks[i], le64(viewID)))
    m3full = orc.aggregate_sigs([m3sig] + extra) if extra else m3sig
    base = dict(ViewID=viewID, BlockNum=9, SenderPubkey=pks[0], Payload=payload, Block=b"\xc0rlp",
                M2AggSig=m2sig, M2Bitmap=m2bm, M3AggSig=m3full, M3Bitmap=bytes(bm3))
    cases = [FBFTMessage(**base)]
    cases.append(FBFTMessage(**{**base, "M3AggSig": m3sig}))                                    # aggregate does not match the completed bitmap
    cases.append(FBFTMessage(**{**base, "M2AggSig": m3full}))                                   # M2 aggregate over the wrong message
    cases.append(FBFTMessage(**{**base, "M3AggSig": m3sig, "M3Bitmap": m3bm}))                  # valid aggregate, below quorum
    cases.append(FBFTMessage(**{**base, "Payload": payload[:5] + bytes([payload[5] ^ 1]) + payload[6:]}))         # M1 proof over another block hash
    cases.append(FBFTMessage(**{**base, "Payload": payload[:32] + BAD_SIG + payload[128:]}))
    cases.append(FBFTMessage(**{**base, "Payload": payload[:100]}))
    cases.append(FBFTMessage(**{**base, "Payload": payload + b"\x00"}))
    cases.append(FBFTMessage(**{**base, "M3AggSig": None}))
    cases.append(FBFTMessage(**{**base, "M3AggSig": BAD_SIG}))
    cases.append(FBFTMessage(**{**base, "M2AggSig": None, "M2Bitmap": None}))                  # no M2 at all: M1 must be there
    cases.append(FBFTMessage(**{**base, "M3Bitmap": bytes(bm3)[:-1]}))                          # parser ignores SetMask's error: empty mask
    exp = [ref.on_new_view(m) for m in cases]
    got = [vc.OnNewViewChecks(m) for m in cases]
    assert got == exp
    assert exp[0] is None and exp[1] == cs.errM3Verify and exp[2] == cs.errM2Verify and exp[3] == cs.errNewViewQuorum
    assert exp[4] == cs.errNewViewM1 and exp[5] == cs.errMultiSigDeserialize and exp[6] == cs.errPayloadLength and exp[7] == cs.errSetMask
    assert exp[8] == cs.errM3Nil and exp[9] == cs.errSigDeserialize and exp[10] is None and exp[11] == cs.errM3Verify

# ------------------------------------------------------------------ CPU: host logic over the oracle backend
def test_keccak256_vectors(fixtures):
    assert cs.Keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert cs.Keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert cs.Keccak256(b"harmony-one").hex() == fixtures["sig_vectors"][0]["msg"]          # staking/types/validator.go:30
    assert cs.Keccak256(b"har", b"mony", b"-one") == cs.Keccak256(b"harmony-one")            # variadic like hash.Keccak256(data ...[]byte)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden                                                                       # independent restatement used for the fixtures
    for ln in (1, 55, 135, 136, 137, 271, 272, 273, 1000):
        m = wl.seeded_bytes("keccak", ln, ln)
        assert cs.Keccak256(m) == make_golden.keccak256(m)

def test_message_padding_is_invisible_to_the_hash(oracle):
    """Why NIL (1 byte), the view id (8) and an M1 payload (>= 128) can share one 48-byte batch (consensus._m48)."""
    for m in (NIL, le64(77), wl.seeded_bytes("pad", 0, 32), wl.seeded_bytes("pad", 1, 200)):
        assert oracle.map_to_g2(m) == oracle.map_to_g2(cs._m48(m))

def test_view_change_storm_host_logic(oracle):
    n = 24
    sign = lambda sk, m: oracle.sign_hash(wl.sk_bytes(sk), m)
    sks, pks, msgs, faults, payload = build_storm(n, 18, sign, lambda ks: [oracle.get_public_key(wl.sk_bytes(k)) for k in ks])
    be = OracleBackend(oracle)
    vc = cs.viewChange(pks, backend=be)
    ref = RefViewChange(oracle, pks)
    check_storm(vc, ref, msgs, faults)
    assert be.calls["verify_status"] == 1 and be.calls["verify_headers"] == 1               # the whole storm: two backend calls
    new_view_cases(vc, ref, oracle, sks, pks, payload, sign)
    # a staked-vote decider plugs in its own predicate (the device then skips its popcount gate)
    vc2 = cs.viewChange(pks, backend=be, isQuorumAchievedByMask=lambda bm: popcount(bm, n) >= 1)
    assert vc2.ProcessViewChangeMsg(msgs[3]) is None                                         # below 2/3 but fine for this decider
    vc3 = cs.viewChange(pks, backend=be, verifyBlock=lambda blk: "block rejected")
    assert vc3.ProcessViewChangeMsg(msgs[15]) == "block rejected" and vc3.ProcessViewChangeMsg(msgs[14]) is None      # M2 has no block
    assert vc.ProcessViewChangeMsgs([]) == []

def test_message_signature_batch_host_logic(oracle):
    k = 9
    sks = [wl.seeded_sk("msgsig", i) for i in range(k)]
    pks = [oracle.get_public_key(wl.sk_bytes(s)) for s in sks]
    bodies = [wl.seeded_bytes("msgsig/body", i, 40 + 37 * i) for i in range(k)]
    sigs = [oracle.sign_hash(wl.sk_bytes(sks[i]), cs.Keccak256(bodies[i])) for i in range(k)]
    bodies[2] = bodies[2][:-1] + bytes([bodies[2][-1] ^ 1]); sigs[4] = BAD_SIG; pks[6] = BAD_KEY; sigs[7] = sigs[7][:50]
    got = cs.verifyMessageSigBatch(pks, bodies, sigs, backend=OracleBackend(oracle))
    assert got == [None, None, cs.errMsgSig, None, cs.errSigDeserialize, None, cs.errKeyDeserialize, cs.errSigDeserialize, None]

# ------------------------------------------------------------------ GPU: the same scenarios through libhbls.so
@pytest.mark.gpu
def test_view_change_storm_250_validators(gbls, oracle):
    """BASELINE configs[3] as the reference meets it: a new leader receives a VIEWCHANGE message from (almost) every validator of a
    250-key committee.  Two device calls for the storm; errors and state equal the sequential handlers run over the oracle."""
    n = 250
    def pk_of(ks):
        blob = gbls.GetPublicKeyBatch(b"".join(wl.sk_bytes(k) for k in ks)); return [blob[48 * i:48 * i + 48] for i in range(len(ks))]
    def sign(sk, m):
        s = oracle.sign_hash(wl.sk_bytes(sk), m); assert s is not None; return s
    sks, pks, msgs, faults, payload = build_storm(n, 160, sign, pk_of)
    assert pks[5] == oracle.get_public_key(wl.sk_bytes(sks[5]))
    l0 = gbls.KernelLaunchCount()
    vc = cs.viewChange(pks)
    ref = RefViewChange(oracle, pks)
    check_storm(vc, ref, msgs, faults)
    assert gbls.KernelLaunchCount() > l0
    new_view_cases(vc, ref, oracle, sks, pks, payload, sign)
    # one message on its own (latency path) gives the same answer as inside the batch
    fresh = cs.viewChange(pks)
    assert fresh.ProcessViewChangeMsg(msgs[14]) is None and fresh.ProcessViewChangeMsg(msgs[1]) == cs.errVerifyM1

@pytest.mark.gpu
def test_consensus_message_signatures(gbls, oracle, fixtures):
    """consensus/consensus_service.go:115-119 signMessage and consensus/checks.go:20-56: the reference's only (sk, message) ->
    signature vector IS signMessage(\"harmony-one\") (staking/types/validator.go:30,525-527), then a queue of messages in one call."""
    v = fixtures["sig_vectors"][0]
    sk = gbls.SecretKey(); sk.DeserializeHexStr(v["sk"])
    assert cs.signMessage(b"harmony-one", sk).hex() == v["sig"]
    pk = sk.GetPublicKey()
    assert cs.verifyMessageSig(pk, b"harmony-one", bytes.fromhex(v["sig"])) is None
    assert cs.verifyMessageSig(pk, b"harmony-two", bytes.fromhex(v["sig"])) == cs.errMsgSig
    assert cs.verifyMessageSig(pk, b"harmony-one", BAD_SIG) == cs.errSigDeserialize
    k = 300
    sks = [wl.seeded_sk("msgsig", i) for i in range(k)]
    blob = gbls.GetPublicKeyBatch(b"".join(wl.sk_bytes(s) for s in sks)); pks = [blob[48 * i:48 * i + 48] for i in range(k)]
    bodies = [wl.seeded_bytes("msgsig/body", i, 30 + (i * 7) % 400) for i in range(k)]
    sig_blob, ok = gbls.SignHashBatch(b"".join(wl.sk_bytes(s) for s in sks), b"".join(cs.Keccak256(b) for b in bodies), 32)
    sigs = [sig_blob[96 * i:96 * i + 96] for i in range(k)]
    assert sigs[0] == oracle.sign_hash(wl.sk_bytes(sks[0]), cs.Keccak256(bodies[0]))
    for i in range(0, k, 37): bodies[i] = bodies[i] + b"!"
    sigs[5] = BAD_SIG; pks[9] = BAD_KEY; sigs[11], sigs[12] = sigs[12], sigs[11]
    got = cs.verifyMessageSigBatch(pks, bodies, sigs)
    exp = []
    for i in range(k):
        if not oracle.pk_check(pks[i]): exp.append(cs.errKeyDeserialize)
        elif not oracle.sig_check(sigs[i]): exp.append(cs.errSigDeserialize)
        else: exp.append(None if oracle.verify_hash(sigs[i], pks[i], cs.Keccak256(bodies[i])) else cs.errMsgSig)
    assert got == exp
    assert exp[5] == cs.errSigDeserialize and exp[9] == cs.errKeyDeserialize and exp[11] == exp[12] == cs.errMsgSig and exp[37] == cs.errMsgSig
    assert sum(1 for e in exp if e is None) == k - len(set(range(0, k, 37)) | {5, 9, 11, 12})
    # raw status entry point: statuses in the order the reference meets the errors
    st = gbls.VerifyBatchStatus(b"".join(pks[:16]), b"".join(sigs[:16]), b"".join(cs.Keccak256(b) for b in bodies[:16]), 32)
    assert st[5] == gbls.VB_BAD_SIG_ENCODING and st[9] == gbls.VB_BAD_KEY_ENCODING and st[11] == gbls.VB_BAD_SIG and st[0] == gbls.VB_BAD_SIG and st[1] == gbls.VB_OK

# ------------------------------------------------------------------ the leader's vote collection (consensus/leader.go:110-345, quorum.go:164-196,354-377)
class OracleVoteBackend(OracleBackend):
    def aggregate_keys(self, pks):
        acc = bytes(48)
        for k in pks:
            if len(k) != 48 or not self.o.pk_check(k): raise ValueError("err blsPublicKeyDeserialize")
            acc = self.o.pk_add(acc, k)
        return acc

class RefLeader:
    """onPrepare / onCommit vote by vote, one oracle call per cgo call."""
    def __init__(self, orc, members, message):
        self.o = orc; self.members = list(members); self.index = {m: i for i, m in enumerate(members)}; self.message = message
        self.box = {}; self.bitmap = bytearray((len(members) + 7) >> 3); self.quorum = 2 * len(members) // 3 + 1
    def on_vote(self, v):
        o = self.o
        for k in v.SenderPubkeys:
            if len(k) != 48 or not o.pk_check(k): return cs.errKeyDeserialize                # parser: BytesToBLSPublicKey
        if not v.SenderPubkeys: return cs.errKeyDeserialize
        if any(k in self.box for k in v.SenderPubkeys): return cs.errAlreadyReceived
        if len(v.Payload) != 96 or not o.sig_check(v.Payload): return cs.errSigDeserialize
        apk = v.SenderPubkeys[0]
        if len(v.SenderPubkeys) > 1:
            apk = bytes(48)
            for k in v.SenderPubkeys: apk = o.pk_add(apk, k)
        if not o.verify_hash(v.Payload, apk, self.message): return cs.errVoteSig
        if len(set(v.SenderPubkeys)) != len(v.SenderPubkeys): return cs.errDuplicateKey
        for k in v.SenderPubkeys: self.box[k] = (list(v.SenderPubkeys), v.Payload)
        if any(k not in self.index for k in v.SenderPubkeys): return cs.errKeyNotFound
        for k in v.SenderPubkeys: i = self.index[k]; self.bitmap[i >> 3] |= 1 << (i & 7)
        return None
    def aggregate(self):
        sigs, seen = [], set()
        for key, (keys, sig) in self.box.items():
            if any(k in seen for k in keys): continue
            seen.update(keys); sigs.append(sig)
        return self.o.aggregate_sigs(sigs), bytes(self.bitmap)

def build_votes(n, sign, pk_of, message, n_votes):
    """single-key votes, multi-key votes (3 keys each, one signature with the key sum), and one fault of every kind"""
    sks = [wl.seeded_sk("votes", i) for i in range(n + 2)]
    pks = pk_of(sks)
    members, outsider = pks[:n], pks[n]
    votes, i = [], 0
    while len(votes) < n_votes and i + 3 <= n:
        if len(votes) % 5 == 4:
            idx = [i, i + 1, i + 2]; i += 3
            votes.append(cs.Vote([pks[j] for j in idx], sign(sum(sks[j] for j in idx) % wl.R_ORDER, message)))
        else:
            votes.append(cs.Vote([pks[i]], sign(sks[i], message))); i += 1
    f = {}
    votes[1].Payload = sign(sks[n + 1], message); f[1] = cs.errVoteSig                        # somebody else's signature
    votes[2].Payload = BAD_SIG; f[2] = cs.errSigDeserialize
    votes[3].SenderPubkeys = [BAD_KEY]; f[3] = cs.errKeyDeserialize
    votes.append(cs.Vote(list(votes[0].SenderPubkeys), votes[0].Payload)); f[len(votes) - 1] = cs.errAlreadyReceived
    votes.append(cs.Vote([outsider], sign(sks[n], message))); f[len(votes) - 1] = cs.errKeyNotFound      # valid signature, not in the committee
    k = votes[5].SenderPubkeys[0]; sk5 = sks[pks.index(k)]
    votes[5] = cs.Vote([k, k], sign(2 * sk5 % wl.R_ORDER, message)); f[5] = cs.errDuplicateKey
    mk = votes[4]                                                                             # multi-key vote whose signature misses one key
    votes[4] = cs.Vote(list(mk.SenderPubkeys), sign(sum(sks[pks.index(x)] for x in mk.SenderPubkeys[:2]) % wl.R_ORDER, message)); f[4] = cs.errVoteSig
    votes.append(cs.Vote([mk.SenderPubkeys[0]], sign(sks[pks.index(mk.SenderPubkeys[0])], message)))      # one of its keys alone: fine
    votes.append(cs.Vote(list(mk.SenderPubkeys), mk.Payload)); f[len(votes) - 1] = cs.errAlreadyReceived   # the full set again: one key has voted
    return members, votes, f

def check_votes(col, ref, votes, faults, com_verify):
    got, quorum_at = col.onVotes(votes)
    exp, q_exp = [], None
    for i, v in enumerate(votes):
        was = len(ref.box) >= ref.quorum
        e = ref.on_vote(v); exp.append(e)
        if e is None and not was and len(ref.box) >= ref.quorum and q_exp is None: q_exp = i
    assert got == exp and quorum_at == q_exp
    for i, e in faults.items(): assert exp[i] == e, (i, exp[i], e)
    assert col.BallotBox == ref.box and bytes(col.bitmap) == bytes(ref.bitmap)
    sig, bm = col.AggregateVotes(); rsig, rbm = ref.aggregate()
    assert (sig, bm) == (rsig, rbm)
    # the PREPARED / COMMITTED payload verifies against the bitmap -- unless a non-member's vote got in: the reference records the
    # ballot (decider.AddNewVote) BEFORE bitmap.SetKeysAtomic rejects the key (leader.go:185-196), so its signature is aggregated
    # without a bit; the mirror reproduces that (callers drop non-members at message validation)
    assert com_verify(bm, sig) == (cs.errKeyNotFound not in exp)
    return quorum_at

def test_leader_vote_collection_host_logic(oracle):
    n = 30; message = wl.commit_payload("votes", 0)
    sign = lambda sk, m: oracle.sign_hash(wl.sk_bytes(sk), m)
    members, votes, faults = build_votes(n, sign, lambda ks: [oracle.get_public_key(wl.sk_bytes(k)) for k in ks], message, 22)
    be = OracleVoteBackend(oracle); och = oracle.committee(members)
    col = cs.VoteCollector(members, message, backend=be); ref = RefLeader(oracle, members, message)
    q = check_votes(col, ref, votes, faults, lambda bm, sig: oracle.committee_aggregate_verify(och, bm, sig, message) == 1)
    assert be.calls["verify_status"] == 1 and q is not None and col.IsQuorumAchieved()
    members_only = [v for i, v in enumerate(votes) if faults.get(i) != cs.errKeyNotFound]
    col2 = cs.VoteCollector(members, message, backend=be); col2.onVotes(members_only)
    sig2, bm2 = col2.AggregateVotes(); assert oracle.committee_aggregate_verify(och, bm2, sig2, message) == 1
    # a second queue continues from the state of the first
    more = [cs.Vote(list(votes[0].SenderPubkeys), votes[0].Payload)]
    assert col.onVotes(more) == ([cs.errAlreadyReceived], None) and col.onVotes([]) == ([], None)

@pytest.mark.gpu
def test_leader_vote_collection_250(gbls, oracle):
    """The leader's PREPARE / COMMIT collection for a 250-validator committee (consensus/leader.go:110-345): one device call for the
    queue, H(m) hashed once; counted votes, drop reasons, quorum index, ballot box, bitmap and the aggregate equal the sequential
    handlers run over the oracle."""
    n = 250; message = wl.commit_payload("votes", 1)
    def pk_of(ks):
        blob = gbls.GetPublicKeyBatch(b"".join(wl.sk_bytes(k) for k in ks)); return [blob[48 * i:48 * i + 48] for i in range(len(ks))]
    sign = lambda sk, m: oracle.sign_hash(wl.sk_bytes(sk), m)
    members, votes, faults = build_votes(n, sign, pk_of, message, 175)
    com = gbls.Committee(members)
    col = cs.VoteCollector(members, message); ref = RefLeader(oracle, members, message)
    q = check_votes(col, ref, votes, faults, lambda bm, sig: com.AggregateVerify(bm, sig, message))
    assert q is not None and col.SignersCount() >= col.quorum
    members_only = [v for i, v in enumerate(votes) if faults.get(i) != cs.errKeyNotFound]
    col2 = cs.VoteCollector(members, message); col2.onVotes(members_only)
    sig2, bm2 = col2.AggregateVotes(); assert com.AggregateVerify(bm2, sig2, message)

# ------------------------------------------------------------------ randomized storms: batch == sequential for every arrival order
def test_view_change_random_storms_host_logic(oracle):
    """Seeded random storms on a small committee: random mix of M1 / M2 messages, two view ids, every fault of the menu at random
    places, duplicates and retries in random order.  The batched handler must agree with the sequential restatement message by message
    and in the final state, for ONE call over the whole queue and for the queue cut into several calls at random places."""
    import random
    n = 9; q = wl.quorum_k(n)
    sks = [wl.seeded_sk("rnd", i) for i in range(n)]
    pks = [oracle.get_public_key(wl.sk_bytes(k)) for k in sks]
    cache = {}
    def sign(sk, m):
        key = (sk, bytes(m))
        if key not in cache: cache[key] = oracle.sign_hash(wl.sk_bytes(sk), m)
        return cache[key]
    bh = wl.seeded_bytes("rnd/hash", 0, 32)
    def proof(k, signed=None):
        bm = wl.bitmap_with_k("rnd/bm", k, n, k)
        return bh + sign(wl.round_signer_sum(sks, bm), signed or bh) + bm
    good, low, wrong = proof(q), proof(q - 1), proof(q, signed=wl.seeded_bytes("rnd/hash", 1, 32))
    be = OracleBackend(oracle)
    for seed in range(12):
        rng = random.Random(seed)
        msgs = []
        for _ in range(rng.randrange(4, 14)):
            i = rng.randrange(n); vid = rng.choice((5, 6)); kind = rng.choice(("m1", "m1", "m2"))
            payload = rng.choice((good, good, good, low, wrong, bh + BAD_SIG + good[128:], good[:-1])) if kind == "m1" else b""
            m = FBFTMessage(ViewID=vid, BlockNum=3, SenderPubkey=pks[i], LeaderPubkey=pks[0], Payload=payload, Block=b"\xc0" if kind == "m1" else b"",
                            ViewchangeSig=sign(sks[i], payload if kind == "m1" else NIL), ViewidSig=sign(sks[i], le64(vid)))
            fault = rng.randrange(10)
            if fault == 0: m.ViewidSig = sign(sks[i], le64(vid + 1))
            elif fault == 1: m.ViewchangeSig = sign(sks[(i + 1) % n], payload if kind == "m1" else NIL)
            elif fault == 2: m.ViewchangeSig = BAD_SIG
            elif fault == 3: m.SenderPubkey = BAD_KEY
            elif fault == 4 and kind == "m1": m.Block = b""                      # M1 payload without a block: judged as M2
            msgs.append(m)
            if rng.randrange(4) == 0: msgs.append(FBFTMessage(**{**m.__dict__}))   # immediate duplicate
        rng.shuffle(msgs)
        ref = RefViewChange(oracle, pks); exp = [ref.on_view_change(m) for m in msgs]
        one = cs.viewChange(pks, backend=be)
        assert one.ProcessViewChangeMsgs(msgs) == exp, seed
        cut = cs.viewChange(pks, backend=be); got = []; pos = 0
        while pos < len(msgs):
            step = rng.randrange(1, 5); got += cut.ProcessViewChangeMsgs(msgs[pos:pos + step]); pos += step
        assert got == exp, seed
        for vc in (one, cut):
            assert vc.bhpSigs == ref.bhpSigs and vc.nilSigs == ref.nilSigs and vc.viewIDSigs == ref.viewIDSigs and vc.GetM1Payload() == ref.m1Payload
            assert {k: bytes(v) for k, v in vc.viewIDBitmap.items()} == {k: bytes(v) for k, v in ref.viewIDBitmap.items()}
