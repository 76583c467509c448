"""
tests/golden/make_golden.py -- regenerates tests/golden/ref_fixtures.json from the reference tree.

Run in the build container only (needs /root/reference; the GPU box never reads it):
    python tests/golden/make_golden.py

Collects every byte-level pin the reference holds for the BLS path:
  * sk->pk pairs in Go test sources (core/tx_pool_test.go:52-53, internal/blsgen/utils_test.go:30-43)
  * .hmy/**/*.key files: file name = pk hex, content = hex(nonce12 || AES-256-GCM(key = hex(md5(pass))))
    of hex(sk) (internal/blsgen/lib.go:101-159); those that open with the empty passphrase.
  * the (sk, keccak256("harmony-one")) -> signature vector
    (rosetta/services/construction_create_test.go:460-467, staking/types/validator.go:30,525-527)
  * valid pubkeys for decode tests (internal/genesis account tables)
No reference source is copied: only hex test data.
"""
import glob, hashlib, json, os, re, sys
from cryptography.hazmat.primitives.ciphers.aead import AESGCM

REF = "/root/reference"

def decrypt_keyfile(blob_hex: str, passphrase: str):
    raw = bytes.fromhex(blob_hex.strip())
    key = hashlib.md5(passphrase.encode()).hexdigest().encode()
    try:
        pt = AESGCM(key).decrypt(raw[:12], raw[12:], None)
    except Exception:
        return None
    return pt.decode().strip()

def keccak256(data: bytes) -> bytes:
    # minimal Keccak-f[1600] (legacy 0x01 padding) -- only used to build the message of the sig vector
    RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
          0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
          0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
          0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
          0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
    ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
    M = (1 << 64) - 1
    rol = lambda v, n: ((v << n) | (v >> (64 - n))) & M if n else v
    rate = 136
    msg = bytearray(data); msg.append(0x01)
    while len(msg) % rate: msg.append(0)
    msg[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], 'little')
        for rnd in range(24):
            C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
            D = [C[(x - 1) % 5] ^ rol(C[(x + 1) % 5], 1) for x in range(5)]
            A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
            Bm = [[0] * 5 for _ in range(5)]
            for x in range(5):
                for y in range(5):
                    Bm[y][(2 * x + 3 * y) % 5] = rol(A[x][y], ROT[x][y])
            A = [[Bm[x][y] ^ ((~Bm[(x + 1) % 5][y]) & Bm[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
            A[0][0] ^= RC[rnd]
    out = b''.join(A[i % 5][i // 5].to_bytes(8, 'little') for i in range(4))
    return out

def main():
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    sk_pk = []
    keyfiles = []          # encrypted blobs themselves: pins of the key-file codec (internal/blsgen/lib.go:20-159)
    seen = set()
    def add(sk, pk, src):
        if (sk, pk) in seen: return
        seen.add((sk, pk)); sk_pk.append({"sk": sk, "pk": pk, "src": src})
    add("c6d7603520311f7a4e6aac0b26701fc433b75b38df504cd416ef2b900cd66205",
        "30b2c38b1316da91e068ac3bd8751c0901ef6c02a1d58bc712104918302c6ed03d5894671d0c816dad2b4d303320f202",
        "core/tx_pool_test.go:52-53")
    src = open(f"{REF}/internal/blsgen/utils_test.go").read()
    for m in re.finditer(r'publicKey:\s+"([0-9a-f]{96})",\s+privateKey:\s+"([0-9a-f]{64})",\s+passphrase:\s+"([^"]*)",\s+keyFileData:\s+"([0-9a-f]+)"', src):
        pk, sk, pw, blob = m.groups()
        add(sk, pk, "internal/blsgen/utils_test.go:30-43")
        assert decrypt_keyfile(blob, pw) == sk
        keyfiles.append({"pk": pk, "pass": pw, "blob": blob, "src": "internal/blsgen/utils_test.go:30-43"})
    nfiles = 0
    for path in sorted(glob.glob(f"{REF}/.hmy/**/*.key", recursive=True)):
        nfiles += 1
        pk = os.path.basename(path)[:-4]
        if not re.fullmatch(r"[0-9a-f]{96}", pk): continue
        pw = ""
        if os.path.exists(path[:-4] + ".pass"):
            pw = open(path[:-4] + ".pass").read().strip("\r\n")
        sk = decrypt_keyfile(open(path).read(), pw)
        if sk is None or not re.fullmatch(r"[0-9a-f]{64}", sk): continue
        add(sk, pk, os.path.relpath(path, REF))
        if len(keyfiles) < 10: keyfiles.append({"pk": pk, "pass": pw, "blob": open(path).read().strip(), "src": os.path.relpath(path, REF)})
    # valid pubkeys (decode-only pins): genesis account tables
    pks = set()
    for path in sorted(glob.glob(f"{REF}/internal/genesis/*.go")):
        for m in re.finditer(r'BLSPublicKey:\s*"([0-9a-f]{96})"', open(path).read()):
            pks.add(m.group(1))
    pks = sorted(pks)
    msg = keccak256(b"harmony-one")
    out = {
        "generated_by": "tests/golden/make_golden.py",
        "sk_pk": sk_pk,
        "keyfiles": keyfiles,
        "sig_vectors": [{
            "sk": "c6d7603520311f7a4e6aac0b26701fc433b75b38df504cd416ef2b900cd66205",
            "pk": "30b2c38b1316da91e068ac3bd8751c0901ef6c02a1d58bc712104918302c6ed03d5894671d0c816dad2b4d303320f202",
            "msg": msg.hex(),
            "msg_note": "keccak256('harmony-one') staking/types/validator.go:30,525",
            "sig": "68f800b6adf657b674903e04708060912b893b7c7b500788808247550ab3e186e56a44ebf3ca488f8ed1a42f6cef3a04bd5d2b2b7eb5a767848d3135b362e668ce6bba42c7b9d5666d8e3a83be707b5708e722c58939fe9b07c170f3b7062414",
            "src": "rosetta/services/construction_create_test.go:460-467"}],
        "genesis_pubkeys_sample": pks[:: max(1, len(pks) // 200)][:200],
        "genesis_pubkeys_total": len(pks),
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_fixtures.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(f"key files: {nfiles}, sk->pk vectors: {len(sk_pk)}, genesis pubkeys: {len(pks)} (sampled {len(out['genesis_pubkeys_sample'])})")

if __name__ == "__main__":
    main()
