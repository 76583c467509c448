"""tests/oracle_lib.py -- ctypes handle on oracle/libhbls_oracle.so (the CPU restatement; test infrastructure only)."""
import ctypes, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

class Oracle:
    def __init__(self, L):
        self.L = L
        c = ctypes
        L.ho_committee_new.restype = c.c_void_p
        L.ho_committee_new.argtypes = [c.c_size_t, c.c_char_p]
        L.ho_committee_free.argtypes = [c.c_void_p]
        L.ho_committee_mask_aggregate.argtypes = [c.c_void_p, c.c_char_p, c.c_size_t, c.c_char_p]
        L.ho_committee_aggregate_verify.argtypes = [c.c_void_p, c.c_char_p, c.c_size_t, c.c_char_p, c.c_char_p, c.c_size_t]
        for name in ("ho_sign_hash", "ho_verify_hash", "ho_map_to_g2", "ho_fast_aggregate_verify", "ho_mask_aggregate", "ho_aggregate_sigs"):
            getattr(L, name).restype = c.c_int
    def get_public_key(self, sk32):
        out = ctypes.create_string_buffer(48)
        if self.L.ho_get_public_key(bytes(sk32), out) != 0: return None
        return out.raw
    def sign_hash(self, sk32, msg):
        out = ctypes.create_string_buffer(96)
        if self.L.ho_sign_hash(bytes(sk32), bytes(msg), ctypes.c_size_t(len(msg)), out) != 0: return None
        return out.raw
    def verify_hash(self, sig96, pk48, msg):
        return self.L.ho_verify_hash(bytes(sig96), bytes(pk48), bytes(msg), ctypes.c_size_t(len(msg))) == 1
    def map_to_g2(self, msg):
        out = ctypes.create_string_buffer(96)
        if self.L.ho_map_to_g2(bytes(msg), ctypes.c_size_t(len(msg)), out) != 0: return None
        return out.raw
    def aggregate_sigs(self, sigs):
        blob = b"".join(sigs); out = ctypes.create_string_buffer(96)
        rc = self.L.ho_aggregate_sigs(ctypes.c_size_t(len(sigs)), blob, out)
        return out.raw if rc == 0 else None
    def mask_aggregate(self, pks, bitmap):
        blob = b"".join(pks); out = ctypes.create_string_buffer(48)
        rc = self.L.ho_mask_aggregate(ctypes.c_size_t(len(pks)), blob, bytes(bitmap), ctypes.c_size_t(len(bitmap)), out)
        if rc != 0: raise ValueError(f"mask_aggregate rc={rc}")
        return out.raw
    def fast_aggregate_verify(self, pks, bitmap, sig96, msg):
        blob = b"".join(pks)
        return self.L.ho_fast_aggregate_verify(ctypes.c_size_t(len(pks)), blob, bytes(bitmap), ctypes.c_size_t(len(bitmap)), bytes(sig96), bytes(msg), ctypes.c_size_t(len(msg)))
    def committee(self, pks):
        h = self.L.ho_committee_new(len(pks), b"".join(pks))
        assert h, "oracle: invalid public key in committee"
        return h
    def committee_aggregate_verify(self, h, bitmap, sig96, msg):
        return self.L.ho_committee_aggregate_verify(h, bytes(bitmap), len(bitmap), bytes(sig96), bytes(msg), len(msg))
    def committee_mask_aggregate(self, h, bitmap):
        out = ctypes.create_string_buffer(48)
        rc = self.L.ho_committee_mask_aggregate(h, bytes(bitmap), len(bitmap), out)
        if rc != 0: raise ValueError("bitmap length")
        return out.raw
    def pk_add(self, a, b, sub=False):
        out = ctypes.create_string_buffer(48); assert self.L.ho_pk_add(bytes(a), bytes(b), int(sub), out) == 0; return out.raw
    def sig_add(self, a, b):
        out = ctypes.create_string_buffer(96); assert self.L.ho_sig_add(bytes(a), bytes(b), out) == 0; return out.raw
    def pk_check(self, pk48): return self.L.ho_pk_deserialize_check(bytes(pk48)) == 1
    def sig_check(self, sig96): return self.L.ho_sig_deserialize_check(bytes(sig96)) == 1
    def fp_mul(self, a, b):
        out = ctypes.create_string_buffer(48); assert self.L.ho_fp_mul(bytes(a), bytes(b), out) == 0; return out.raw
    def counters(self):
        c = (ctypes.c_uint64 * 2)(); self.L.ho_counters_get(c); return int(c[0]), int(c[1])
    def counters_reset(self): self.L.ho_counters_reset()

_cached = None
def load():
    global _cached
    if _cached is None:
        from harmony_b200 import build
        path = build.build_oracle()
        _cached = Oracle(ctypes.CDLL(path))
    return _cached
