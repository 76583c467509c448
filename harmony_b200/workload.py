"""harmony_b200/workload.py -- deterministic synthetic committees / rounds, byte-level only.

Seeds every value with SHA-256 counter mode over "hbls-bench" || tag || index, never a CSPRNG, so the oracle (tests)
and the product (bench) see identical inputs.  No group arithmetic happens here: secret keys are scalars, and the
aggregate signature of a round is produced by whoever signs with sum(sk_i) mod r  ((sum sk_i) * H(m) == sum (sk_i * H(m))).
"""
import hashlib, random

R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001

def seeded_bytes(tag: str, index: int, n: int) -> bytes:
    out = b""; ctr = 0
    while len(out) < n:
        out += hashlib.sha256(b"hbls-bench" + tag.encode() + index.to_bytes(8, "little") + ctr.to_bytes(4, "little")).digest()
        ctr += 1
    return out[:n]

def seeded_sk(tag: str, index: int) -> int:
    return int.from_bytes(seeded_bytes(tag, index, 32), "little") % R_ORDER

def sk_bytes(k: int) -> bytes:
    return (k % R_ORDER).to_bytes(32, "little")

def commit_payload(tag: str, j: int) -> bytes:
    """consensus/signature/signature.go:12-24, staking era: LE64(blockNum) || hash32 || LE64(viewID); blockNum, viewID < 2^32."""
    h = seeded_bytes(tag + "/hash", j, 32)
    bn = int.from_bytes(seeded_bytes(tag + "/bn", j, 4), "little")
    vid = int.from_bytes(seeded_bytes(tag + "/vid", j, 4), "little")
    return bn.to_bytes(8, "little") + h + vid.to_bytes(8, "little")

def bitmap_with_k(tag: str, j: int, n: int, k: int) -> bytes:
    rng = random.Random(int.from_bytes(seeded_bytes(tag + "/bm", j, 8), "little"))
    idx = list(range(n)); rng.shuffle(idx)
    bm = bytearray((n + 7) >> 3)
    for i in idx[:k]: bm[i >> 3] |= 1 << (i & 7)
    return bytes(bm)

def round_signer_sum(sks, bitmap: bytes) -> int:
    s = 0
    for i, k in enumerate(sks):
        if bitmap[i >> 3] & (1 << (i & 7)): s += k
    return s % R_ORDER

def quorum_k(n: int) -> int:
    """one-node-one-vote quorum (consensus/quorum/one-node-one-vote.go:57-72): floor(2n/3)+1."""
    return 2 * n // 3 + 1
