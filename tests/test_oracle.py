"""CPU tests (no GPU): pin the oracle.  oracle/pyref.py (big-int) and oracle/hbls_oracle.c (6x64 Montgomery) are checked
against every byte-level fixture the reference holds for the path (tests/golden/ref_fixtures.json) and
against each other; the reference's functional pins (quorom_test.go, mask_test.go) are restated on the oracle."""
import os, random, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref as o
from harmony_b200 import workload as wl

def test_pyref_golden_sk_pk(fixtures):
    assert len(fixtures["sk_pk"]) >= 35
    odd = 0
    for v in fixtures["sk_pk"]:
        sk = o.fr_from_bytes(bytes.fromhex(v["sk"]))
        pk = o.g1_serialize(o.get_public_key(sk))
        assert pk.hex() == v["pk"], v["src"]
        odd += pk[47] >> 7
    assert odd >= 5          # the y-parity flag is exercised both ways

def test_pyref_golden_signature(fixtures):
    sv = fixtures["sig_vectors"][0]
    msg = bytes.fromhex(sv["msg"])
    sig = o.sign_hash(o.fr_from_bytes(bytes.fromhex(sv["sk"])), msg)
    assert o.g2_serialize(sig).hex() == sv["sig"]
    pk = o.g1_deserialize(bytes.fromhex(sv["pk"]))
    s2 = o.g2_deserialize(bytes.fromhex(sv["sig"]))
    assert o.verify_hash(s2, pk, msg)
    assert not o.verify_hash(s2, pk, msg[:-1] + bytes([msg[-1] ^ 1]))

def test_c_oracle_golden(oracle, fixtures):
    for v in fixtures["sk_pk"]:
        assert oracle.get_public_key(bytes.fromhex(v["sk"])).hex() == v["pk"], v["src"]
    sv = fixtures["sig_vectors"][0]
    msg = bytes.fromhex(sv["msg"])
    assert oracle.sign_hash(bytes.fromhex(sv["sk"]), msg).hex() == sv["sig"]
    assert oracle.verify_hash(bytes.fromhex(sv["sig"]), bytes.fromhex(sv["pk"]), msg)
    assert not oracle.verify_hash(bytes.fromhex(sv["sig"]), bytes.fromhex(sv["pk"]), b"\x00" + msg[1:])
    for h in fixtures["genesis_pubkeys_sample"][:40]:
        assert oracle.pk_check(bytes.fromhex(h)), h

def test_c_vs_pyref_field_and_map(oracle):
    rng = random.Random(5)
    for _ in range(200):
        a, b = rng.randrange(o.P), rng.randrange(o.P)
        assert int.from_bytes(oracle.fp_mul(a.to_bytes(48, "little"), b.to_bytes(48, "little")), "little") == a * b % o.P
    for m in [b"\x01", bytes(8), rng.randbytes(32), rng.randbytes(48), rng.randbytes(64), b"\xff" * 48, rng.randbytes(5)]:
        h = o.map_to_g2(m)          # bytes(8) (viewID 0) maps to t = 0: undefined in both
        assert oracle.map_to_g2(m) == (o.g2_serialize(h) if h is not None else None), m.hex()
    assert oracle.map_to_g2(bytes(32)) is None and o.map_to_g2(bytes(32)) is None

def test_c_vs_pyref_sign_aggregate_verify(oracle):
    sks = [wl.seeded_sk("orc", i) for i in range(5)]
    msg = wl.commit_payload("orc", 0)
    sigs_c = [oracle.sign_hash(wl.sk_bytes(k), msg) for k in sks]
    sigs_p = [o.sign_hash(k, msg) for k in sks]
    for c, p in zip(sigs_c, sigs_p): assert c == o.g2_serialize(p)
    pks_c = [oracle.get_public_key(wl.sk_bytes(k)) for k in sks]
    for c, k in zip(pks_c, sks): assert c == o.g1_serialize(o.get_public_key(k))
    agg_c = oracle.aggregate_sigs(sigs_c[:4])
    assert agg_c == o.g2_serialize(o.aggregate_sigs(sigs_p[:4]))
    assert agg_c == oracle.sign_hash(wl.sk_bytes(sum(sks[:4]) % o.R), msg)      # (sum sk) H == sum (sk H)
    bm = b"\x0f"
    apk_c = oracle.mask_aggregate(pks_c, bm)
    assert apk_c == o.g1_serialize(o.mask_aggregate([o.get_public_key(k) for k in sks], bm))
    assert oracle.fast_aggregate_verify(pks_c, bm, agg_c, msg) == 1
    assert o.fast_aggregate_verify([o.get_public_key(k) for k in sks], bm, agg_c, msg)
    assert oracle.fast_aggregate_verify(pks_c, b"\x1f", agg_c, msg) == 0
    assert oracle.fast_aggregate_verify(pks_c, b"\x0f\x00", agg_c, msg) == -1      # mismatching bitmap length (mask.go:114-120)
    # add / sub
    assert oracle.pk_add(pks_c[0], pks_c[1]) == o.g1_serialize(o.pt_add(o.FP, o.get_public_key(sks[0]), o.get_public_key(sks[1])))
    assert oracle.pk_add(oracle.pk_add(pks_c[0], pks_c[1]), pks_c[1], sub=True) == pks_c[0]
    assert oracle.pk_add(pks_c[0], pks_c[0], sub=True) == bytes(48)

def test_quorum_pins_on_oracle(oracle):
    """consensus/quorum/quorom_test.go:381-552: 4-of-8 aggregate verifies; duplicated signer fails against the deduped keys."""
    sks = [wl.seeded_sk("q", i) for i in range(8)]
    pks = [oracle.get_public_key(wl.sk_bytes(k)) for k in sks]
    msg = wl.seeded_bytes("q/h", 0, 32)
    sigs = [oracle.sign_hash(wl.sk_bytes(k), msg) for k in sks]
    agg4 = oracle.aggregate_sigs(sigs[:4])
    assert oracle.fast_aggregate_verify(pks, b"\x0f", agg4, msg) == 1
    assert oracle.fast_aggregate_verify(pks, b"\xf0", agg4, msg) == 0
    dup = oracle.aggregate_sigs([sigs[0], sigs[1], sigs[1]])
    assert oracle.fast_aggregate_verify(pks, b"\x03", dup, msg) == 0
    assert oracle.fast_aggregate_verify(pks, b"\x03", oracle.aggregate_sigs(sigs[:2]), msg) == 1
    # aggregation order is free (quorum.go:165-195 iterates a Go map)
    assert oracle.aggregate_sigs(sigs[:4][::-1]) == agg4

def test_pairing_bilinearity_pyref():
    a, b = 0x1234567, 0x89abcde
    Q = o.map_to_g2(b"bilinear")
    Pg = o.G1_GEN
    # e(aP, bQ) * e(-abP, Q) == 1
    assert o.pairing_product_is_one([(o.pt_mul(o.FP, Pg, a), o.pt_mul(o.FP2, Q, b)), (o.pt_neg(o.FP, o.pt_mul(o.FP, Pg, a * b)), Q)])
    assert not o.pairing_product_is_one([(o.pt_mul(o.FP, Pg, a), o.pt_mul(o.FP2, Q, b)), (o.pt_neg(o.FP, o.pt_mul(o.FP, Pg, a * b + 1)), Q)])

def test_deserialize_semantics(oracle):
    assert o.g1_deserialize(bytes(48)) == o.pt_inf(o.FP) and oracle.pk_check(bytes(48))
    assert o.g1_deserialize(b"\xff" * 48) is None and not oracle.pk_check(b"\xff" * 48)
    rng = random.Random(9); rej = 0
    for _ in range(10):
        b = bytearray(rng.randbytes(48)); b[47] &= 0x99
        exp = o.g1_deserialize(bytes(b)) is not None
        assert oracle.pk_check(bytes(b)) == exp; rej += not exp
    assert rej > 0
    for _ in range(4):
        b = bytearray(rng.randbytes(96)); b[95] &= 0x99; b[47] &= 0x19
        assert oracle.sig_check(bytes(b)) == (o.g2_deserialize(bytes(b)) is not None)

def test_truncation_to_48_bytes(oracle):
    """inputs longer than 48 bytes are silently truncated (view-change M1 payloads)."""
    m = wl.seeded_bytes("trunc", 0, 128)
    assert oracle.map_to_g2(m) == oracle.map_to_g2(m[:48])
    assert oracle.map_to_g2(m) != oracle.map_to_g2(m[:47])

def test_commit_payload():
    """consensus/signature/signature_test.go restated."""
    h = bytes(range(32))
    p = o.construct_commit_payload(0x0102030405060708, h, 0x1112131415161718, True)
    assert p == bytes([8, 7, 6, 5, 4, 3, 2, 1]) + h + bytes([0x18, 0x17, 0x16, 0x15, 0x14, 0x13, 0x12, 0x11])
    assert o.construct_commit_payload(1, h, 2, False) == (1).to_bytes(8, "little") + h
    from harmony_b200 import bls          # pure-Python part of the binding: no library load here
    assert bls.ConstructCommitPayload(True, h, 0x0102030405060708, 0x1112131415161718) == p

def test_op_counter(oracle):
    oracle.counters_reset()
    sk = wl.sk_bytes(wl.seeded_sk("cnt", 0)); msg = wl.commit_payload("cnt", 0)
    sig = oracle.sign_hash(sk, msg); pk = oracle.get_public_key(sk)
    oracle.counters_reset()
    assert oracle.verify_hash(sig, pk, msg)
    mul, sqr = oracle.counters()
    assert 15000 < mul < 60000 and 3000 < sqr < 30000
