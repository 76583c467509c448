"""Double-sign slash verification (harmony_b200/slash.py) against a sequential restatement of staking/slash/double-sign.go:139-168,
215-262 over the CPU oracle; the records follow staking/slash/double-sign_test.go:79 (a validator with several BLS keys signing two
conflicting blocks at one height / view)."""
import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harmony_b200 import workload as wl, slash, bls as hb
from test_consensus import OracleBackend, BAD_SIG, BAD_KEY

class OracleSlashBackend(OracleBackend):
    def aggregate_keys(self, pks):
        acc = bytes(48)
        for k in pks:
            if not self.o.pk_check(k): raise ValueError("err blsPublicKeyDeserialize")
            acc = self.o.pk_add(acc, k)
        return acc
    def sig_decodes(self, sig): return self.o.sig_check(sig)

def ref_verify(orc, rec):
    """double-sign.go, ballot checks only, one oracle call per cgo call."""
    ev = rec.Evidence; first, second = ev.FirstVote, ev.SecondVote
    for k in list(first.SignerPubKeys) + list(second.SignerPubKeys):
        if len(k) != 48: return slash.errSignerKeyNotRightSize
    if first.BlockHeaderHash == second.BlockHeaderHash: return slash.errSlashBlockNoConflict
    if not [a for a in first.SignerPubKeys if a in second.SignerPubKeys]: return slash.errNoMatchingDoubleSignKeys
    for ballot in (first, second):
        if len(ballot.Signature) != 96 or not orc.sig_check(ballot.Signature): return slash.errSigDeserialize
        apk = bytes(48)
        for k in ballot.SignerPubKeys:
            if not orc.pk_check(k): return slash.errKeyDeserialize
            apk = orc.pk_add(apk, k)
        payload = hb.ConstructCommitPayload(True, ballot.BlockHeaderHash, ev.Height, ev.ViewID)
        if not orc.verify_hash(ballot.Signature, apk, payload): return slash.errFailVerifySlash
    return None

def build_records(sign, pk_of):
    sks = [wl.seeded_sk("slash", i) for i in range(6)]
    pks = pk_of(sks)
    h1, h2 = wl.seeded_bytes("slash/h", 1, 32), wl.seeded_bytes("slash/h", 2, 32)
    height, view = 37, 38
    def vote(idx, h, signed=None):
        s = sum(sks[i] for i in idx) % wl.R_ORDER
        return slash.Vote([pks[i] for i in idx], h, sign(s, hb.ConstructCommitPayload(True, signed or h, height, view)))
    def rec(a, b): return slash.Record(slash.Evidence(Epoch=3, ShardID=0, Height=height, ViewID=view, FirstVote=a, SecondVote=b))
    recs = [rec(vote([0, 1], h1), vote([0, 1], h2)),                    # valid: both keys of the validator signed both blocks
            rec(vote([0], h1), vote([0, 1, 2], h2)),                    # valid: overlapping key sets
            rec(vote([0, 1], h1), vote([0, 1], h1)),                    # same block
            rec(vote([0, 1], h1), vote([2, 3], h2)),                    # no key signed both
            rec(vote([0, 1], h1), vote([0, 1], h2, signed=h1)),         # second signature is over the first block
            rec(vote([4], h1, signed=h2), vote([4], h2)),               # first signature does not verify
            rec(vote([0, 1], h1), vote([0, 1], h2)),
            rec(vote([0, 1], h1), vote([0, 1], h2)),
            rec(vote([0, 1], h1), vote([0, 1], h2)),
            rec(vote([5], h1), vote([5], h2))]
    recs[6].Evidence.FirstVote.Signature = BAD_SIG
    recs[7].Evidence.SecondVote.SignerPubKeys = [pks[0], BAD_KEY]; recs[7].Evidence.FirstVote.SignerPubKeys = [pks[0], pks[1]]
    recs[8].Evidence.SecondVote.SignerPubKeys = [pks[0], pks[1][:40]]
    recs[9].Evidence.SecondVote.Signature = recs[9].Evidence.SecondVote.Signature[:60]
    exp = [None, None, slash.errSlashBlockNoConflict, slash.errNoMatchingDoubleSignKeys, slash.errFailVerifySlash, slash.errFailVerifySlash,
           slash.errSigDeserialize, slash.errKeyDeserialize, slash.errSignerKeyNotRightSize, slash.errSigDeserialize]
    return recs, exp

def test_slash_ballots_host_logic(oracle):
    sign = lambda sk, m: oracle.sign_hash(wl.sk_bytes(sk), m)
    recs, exp = build_records(sign, lambda ks: [oracle.get_public_key(wl.sk_bytes(k)) for k in ks])
    be = OracleSlashBackend(oracle)
    assert [ref_verify(oracle, r) for r in recs] == exp
    assert slash.VerifyBallots(recs, backend=be) == exp
    assert be.calls["verify_status"] == 1 and slash.VerifyBallots([], backend=be) == []

@pytest.mark.gpu
def test_slash_ballots_on_device(gbls, oracle):
    def pk_of(ks):
        blob = gbls.GetPublicKeyBatch(b"".join(wl.sk_bytes(k) for k in ks)); return [blob[48 * i:48 * i + 48] for i in range(len(ks))]
    sign = lambda sk, m: oracle.sign_hash(wl.sk_bytes(sk), m)
    recs, exp = build_records(sign, pk_of)
    assert [ref_verify(oracle, r) for r in recs] == exp
    assert slash.VerifyBallots(recs) == exp
    # many records in one call
    many = [recs[i % 6] for i in range(300)]
    assert slash.VerifyBallots(many) == [exp[i % 6] for i in range(300)]
