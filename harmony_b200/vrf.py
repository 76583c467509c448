"""harmony_b200/vrf.py -- BLS-VRF over the CUDA backend (next-tier caller of the same primitives).

Mirror of reference crypto/vrf/bls/bls_vrf.go:63-101: pi = SignHash(sha256(alpha)); beta = sha256(pi);
ProofToHash = Deserialize(pi) + VerifyHash(pk, sha256(alpha)).  `ProofToHashBatch` pushes many (pk, alpha, pi)
triples through ONE device call (hbls_verify_batch), the shape `internal/chain/engine.go:137-189` would feed per block range."""
import hashlib
from . import bls

class ErrInvalidVRF(Exception):
    pass

class PrivateKey:
    def __init__(self, sk: "bls.SecretKey"): self.sk = sk
    def Public(self): return self.sk.GetPublicKey()
    def Evaluate(self, alpha: bytes):
        """([32]byte beta, pi bytes); ([0]*32, None) when the message maps to no point (SignHash returns nil)."""
        pi = self.sk.SignHash(hashlib.sha256(alpha).digest())
        if pi is None:
            return bytes(32), None
        ser = pi.Serialize()
        return hashlib.sha256(ser).digest(), ser

class PublicKey:
    def __init__(self, pk: "bls.PublicKey"): self.pk = pk
    def ProofToHash(self, alpha: bytes, pi: bytes) -> bytes:
        if len(pi) == 0:
            raise ErrInvalidVRF("invalid VRF proof")
        sig = bls.Sign()
        sig.Deserialize(pi)                         # ValueError == the Go deserialize error
        if not sig.VerifyHash(self.pk, hashlib.sha256(alpha).digest()):
            raise ErrInvalidVRF("invalid VRF proof")
        return hashlib.sha256(pi).digest()

def NewVRFSigner(sk): return PrivateKey(sk)
def NewVRFVerifier(pk): return PublicKey(pk)

def ProofToHashBatch(pks48, alphas, pis):
    """List of beta (bytes) or None per item; one hbls_verify_batch call."""
    n = len(pis)
    ok_len = [len(p) == 96 for p in pis]
    sigs = b"".join(p if ok else bytes(96) for p, ok in zip(pis, ok_len))
    msgs = b"".join(hashlib.sha256(a).digest() for a in alphas)
    res = bls.VerifyBatch(b"".join(pks48), sigs, msgs, 32) if n else b""
    return [hashlib.sha256(pis[i]).digest() if (ok_len[i] and res[i] == 1) else None for i in range(n)]
