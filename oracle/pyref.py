"""
oracle/pyref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Plain-Python big-integer restatement of the BLS12-381 arithmetic that Harmony's
`crypto/bls` reaches through cgo (`github.com/harmony-one/bls v0.0.6`, reference
go.mod:27 -> herumi libbls384_256/libmcl built with BLS_SWAP_G=1, reference
Makefile:68-70).  Neither library source is under /root/reference, so this file
restates the published herumi/mcl algorithm and is PINNED
against every byte-level fixture the reference holds for the path
(tests/golden/*.json, produced by tests/golden/make_golden.py):

  * 65 sk -> pk vectors     (core/tx_pool_test.go:52-53, internal/blsgen/utils_test.go:30-43,
                             .hmy/**/*.key with empty passphrase)
  * 1 (sk, msg) -> sig      (rosetta/services/construction_create_test.go:460-467,
                             staking/types/validator.go:30,525-527)

Reference call sites restated here:
  SecretKey.GetPublicKey / SignHash / Sign.VerifyHash / Add / Sub / Serialize /
  Deserialize  -- the surface used by crypto/bls/mask.go:58-134,
  consensus/quorum/quorum.go:164-196, internal/chain/engine.go:619-642.

Pure-Python loops: use only for small cases (a pairing takes ~0.3 s).
"""
import hashlib

# ---------------------------------------------------------------- parameters (A.1)
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
Z_ABS = 0xd201000000010000          # z = -Z_ABS
Z = -Z_ABS
B1 = 4                               # G1: y^2 = x^3 + 4
B2 = (4, 4)                          # G2: y^2 = x^3 + 4(1+i)
XI = (1, 1)

# ---------------------------------------------------------------- Fp
def fp_inv(a):
    return pow(a, P - 2, P)

def fp_sqrt(a):
    """mcl Fp::squareRoot for p = 3 mod 4: candidate a^((p+1)/4), accepted iff it squares back."""
    a %= P
    y = pow(a, (P + 1) // 4, P)
    if y * y % P != a:
        return None
    return y

def fp_legendre(a):
    a %= P
    if a == 0:
        return 0
    return 1 if pow(a, (P - 1) // 2, P) == 1 else -1

# ---------------------------------------------------------------- Fp2 = Fp[i]/(i^2+1)
def f2(a, b=0):
    return (a % P, b % P)

F2_ZERO = (0, 0)
F2_ONE = (1, 0)

def f2_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def f2_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def f2_neg(x): return ((-x[0]) % P, (-x[1]) % P)
def f2_conj(x): return (x[0], (-x[1]) % P)
def f2_mul(x, y):
    a, b = x; c, d = y
    return ((a * c - b * d) % P, (a * d + b * c) % P)
def f2_sqr(x):
    a, b = x
    return ((a + b) * (a - b) % P, 2 * a * b % P)
def f2_mulfp(x, k): return (x[0] * k % P, x[1] * k % P)
def f2_norm(x): return (x[0] * x[0] + x[1] * x[1]) % P
def f2_inv(x):
    n = fp_inv(f2_norm(x))
    return (x[0] * n % P, (-x[1]) * n % P)
def f2_is_zero(x): return x[0] == 0 and x[1] == 0
def f2_mul_xi(x):
    a, b = x
    return ((a - b) % P, (a + b) % P)
def f2_pow(x, e):
    r = F2_ONE
    for bit in bin(e)[2:]:
        r = f2_sqr(r)
        if bit == '1':
            r = f2_mul(r, x)
    return r

def f2_sqrt(x):
    """mcl Fp2::squareRoot: fixes WHICH root is produced."""
    a, b = x
    if b == 0:
        t = fp_sqrt(a)
        if t is not None:
            return (t, 0)
        t = fp_sqrt(-a)
        assert t is not None
        return (0, t)
    n = fp_sqrt((a * a + b * b) % P)
    if n is None:
        return None
    inv2 = (P + 1) // 2
    c = fp_sqrt((a + n) * inv2 % P)
    if c is None:
        c = fp_sqrt((a - n) * inv2 % P)
        assert c is not None
    return (c, b * fp_inv(2 * c % P) % P)

# ---------------------------------------------------------------- Fp12 = Fp2[w]/(w^6 - xi), dense 6-coefficient form
F12_ONE = (F2_ONE,) + (F2_ZERO,) * 5

def f12_mul(x, y):
    acc = [F2_ZERO] * 11
    for i in range(6):
        if f2_is_zero(x[i]):
            continue
        for j in range(6):
            acc[i + j] = f2_add(acc[i + j], f2_mul(x[i], y[j]))
    for k in range(10, 5, -1):
        acc[k - 6] = f2_add(acc[k - 6], f2_mul_xi(acc[k]))
    return tuple(acc[:6])

def f12_sqr(x): return f12_mul(x, x)
def f12_conj(x):  # w -> -w  (= Frobenius p^6)
    return tuple(c if (k % 2 == 0) else f2_neg(c) for k, c in enumerate(x))

# Fp6 = Fp2[v]/(v^3 - xi) helpers for inversion (v = w^2)
def _f6_mul(x, y):
    a0, a1, a2 = x; b0, b1, b2 = y
    c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)
def _f6_sub(x, y): return tuple(f2_sub(a, b) for a, b in zip(x, y))
def _f6_neg(x): return tuple(f2_neg(a) for a in x)
def _f6_mul_v(x): return (f2_mul_xi(x[2]), x[0], x[1])
def _f6_inv(x):
    a0, a1, a2 = x
    t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))

def f12_inv(x):
    a0 = (x[0], x[2], x[4]); a1 = (x[1], x[3], x[5])
    d = _f6_sub(_f6_mul(a0, a0), _f6_mul_v(_f6_mul(a1, a1)))
    di = _f6_inv(d)
    r0 = _f6_mul(a0, di); r1 = _f6_neg(_f6_mul(a1, di))
    return (r0[0], r1[0], r0[1], r1[1], r0[2], r1[2])

def f12_pow(x, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == '1':
            r = f12_mul(r, x)
    return r

_GAMMA_P2 = [pow(f2_pow(XI, (P * P - 1) // 6)[0], k, P) for k in range(6)]  # xi^((p^2-1)/6) lies in Fp
assert f2_pow(XI, (P * P - 1) // 6)[1] == 0

def f12_frob2(x):
    return tuple(f2_mulfp(c, _GAMMA_P2[k]) for k, c in enumerate(x))

# ---------------------------------------------------------------- curves (Jacobian, generic over the field)
class _Field:
    pass

class _FpOps(_Field):
    zero = 0; one = 1
    add = staticmethod(lambda x, y: (x + y) % P)
    sub = staticmethod(lambda x, y: (x - y) % P)
    neg = staticmethod(lambda x: (-x) % P)
    mul = staticmethod(lambda x, y: x * y % P)
    sqr = staticmethod(lambda x: x * x % P)
    inv = staticmethod(fp_inv)
    is_zero = staticmethod(lambda x: x % P == 0)
    b = B1

class _Fp2Ops(_Field):
    zero = F2_ZERO; one = F2_ONE
    add = staticmethod(f2_add); sub = staticmethod(f2_sub); neg = staticmethod(f2_neg)
    mul = staticmethod(f2_mul); sqr = staticmethod(f2_sqr); inv = staticmethod(f2_inv)
    is_zero = staticmethod(f2_is_zero)
    b = B2

FP = _FpOps; FP2 = _Fp2Ops

def pt_inf(F): return (F.one, F.one, F.zero)
def pt_is_inf(F, p): return F.is_zero(p[2])
def pt_from_affine(F, x, y): return (x, y, F.one)

def pt_dbl(F, p):
    X, Y, Zc = p
    if F.is_zero(Zc) or F.is_zero(Y):
        return pt_inf(F)
    A = F.sqr(X); Bq = F.sqr(Y); C = F.sqr(Bq)
    D = F.sub(F.sqr(F.add(X, Bq)), F.add(A, C)); D = F.add(D, D)
    E = F.add(F.add(A, A), A)
    Fq = F.sqr(E)
    X3 = F.sub(Fq, F.add(D, D))
    C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    Z3 = F.mul(F.add(Y, Y), Zc)
    return (X3, Y3, Z3)

def pt_add(F, p, q):
    if pt_is_inf(F, p): return q
    if pt_is_inf(F, q): return p
    X1, Y1, Z1 = p; X2, Y2, Z2 = q
    Z1Z1 = F.sqr(Z1); Z2Z2 = F.sqr(Z2)
    U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2); S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    H = F.sub(U2, U1); Rr = F.sub(S2, S1)
    if F.is_zero(H):
        if F.is_zero(Rr):
            return pt_dbl(F, p)
        return pt_inf(F)
    HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)

def pt_neg(F, p): return (p[0], F.neg(p[1]), p[2])
def pt_sub(F, p, q): return pt_add(F, p, pt_neg(F, q))

def pt_mul(F, p, k):
    if k < 0:
        return pt_mul(F, pt_neg(F, p), -k)
    r = pt_inf(F)
    for bit in bin(k)[2:] if k else '':
        r = pt_dbl(F, r)
        if bit == '1':
            r = pt_add(F, r, p)
    return r

def pt_affine(F, p):
    if pt_is_inf(F, p):
        return None
    zi = F.inv(p[2]); zi2 = F.sqr(zi)
    return (F.mul(p[0], zi2), F.mul(p[1], F.mul(zi2, zi)))

def pt_eq(F, p, q):
    if pt_is_inf(F, p) or pt_is_inf(F, q):
        return pt_is_inf(F, p) and pt_is_inf(F, q)
    Z1Z1 = F.sqr(p[2]); Z2Z2 = F.sqr(q[2])
    if not F.is_zero(F.sub(F.mul(p[0], Z2Z2), F.mul(q[0], Z1Z1))): return False
    return F.is_zero(F.sub(F.mul(F.mul(p[1], q[2]), Z2Z2), F.mul(F.mul(q[1], p[2]), Z1Z1)))

def pt_on_curve(F, p):
    a = pt_affine(F, p)
    if a is None: return True
    x, y = a
    return F.is_zero(F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), F.b)))

def pt_in_subgroup(F, p):
    return pt_is_inf(F, pt_mul(F, p, R))

# ---------------------------------------------------------------- SW map (mcl MapTo::calcBN, A.3(3))
C1 = fp_sqrt(-3)                       # mcl: c1 = sqrt(-3) as produced by Fp::squareRoot
assert C1 == 0xbe32ce5fbeed9ca374d38c0ed41eefd5bb675277cdf12d11bc2fb026c41400045c03fffffffdfffd
C2 = (C1 - 1) * fp_inv(2) % P

def _sw_map_fp(t):
    """G1 flavour (field Fp, b=4); only used to derive the generator (A.2)."""
    t %= P
    if t == 0: return None
    negative = fp_legendre(t) < 0
    w = (t * t + B1 + 1) % P
    if w == 0: return None
    w = fp_inv(w) * C1 % P * t % P
    x = None
    for i in range(3):
        if i == 0: x = (C2 - t * w) % P
        elif i == 1: x = (-x - 1) % P
        else: x = (fp_inv(w * w % P) + 1) % P
        y = fp_sqrt((x * x * x + B1) % P)
        if y is not None:
            if negative: y = (-y) % P
            return (x, y)
    return None

def sw_map_fp2(t):
    """G2 flavour: t in Fp2 -> affine point on E'(Fp2), NOT yet in the r-torsion."""
    if f2_is_zero(t): return None
    negative = fp_legendre(f2_norm(t)) < 0
    w = f2_add(f2_sqr(t), B2); w = ((w[0] + 1) % P, w[1])
    if f2_is_zero(w): return None
    w = f2_mul(f2_mulfp(f2_inv(w), C1), t)
    x = None
    for i in range(3):
        if i == 0:
            x = f2_neg(f2_mul(t, w)); x = ((x[0] + C2) % P, x[1])
        elif i == 1:
            x = f2_neg(x); x = ((x[0] - 1) % P, x[1])
        else:
            x = f2_inv(f2_sqr(w)); x = ((x[0] + 1) % P, x[1])
        y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B2))
        if y is not None:
            if negative: y = f2_neg(y)
            return (x, y)
    return None

# ---------------------------------------------------------------- generator (A.2): B = cofactor * SWmap_G1(1)
G1_COFACTOR = (Z - 1) ** 2 // 3
_g = _sw_map_fp(1)
G1_GEN = pt_mul(FP, pt_from_affine(FP, *_g), G1_COFACTOR)
_ga = pt_affine(FP, G1_GEN)
assert _ga[0] == 0x04f58f3d9ee829f9a853f80b0e32c2981be883a537f0c21ad4af17be22e6e9959915ec21b7f9d8cc4c7315f31f3600e5
assert _ga[1] == 0x1212110eb10dbc575bccc44dcd77400f38282c4728b5efac69c0b4c9011bd27b8ed608acd81f027039216a291ac636a8
G1_GEN = pt_from_affine(FP, *_ga)

# ---------------------------------------------------------------- psi endomorphism + Budroni-Pintore cofactor clearing (A.3(4))
PSI_CX = f2_inv(f2_pow(XI, (P - 1) // 3))
PSI_CY = f2_inv(f2_pow(XI, (P - 1) // 2))
assert PSI_CX == (0, 0x1a0111ea397fe699ec02408663d4de85aa0d857d89759ad4897d29650fb85f9b409427eb4f49fffd8bfd00000000aaad)

def g2_psi(p):
    a = pt_affine(FP2, p)
    if a is None: return pt_inf(FP2)
    return pt_from_affine(FP2, f2_mul(f2_conj(a[0]), PSI_CX), f2_mul(f2_conj(a[1]), PSI_CY))

def g2_clear_cofactor(p):
    t1 = pt_mul(FP2, p, Z * Z - Z - 1)
    t2 = g2_psi(pt_mul(FP2, p, Z - 1))
    t3 = g2_psi(g2_psi(pt_dbl(FP2, p)))
    return pt_add(FP2, pt_add(FP2, t1, t2), t3)

# ---------------------------------------------------------------- message -> G2 (A.3(1-2))
def hash_to_fp(msg: bytes):
    """mcl Fp::setArrayMask on the first min(len,48) bytes, little-endian."""
    v = int.from_bytes(msg[:48], 'little')
    v &= (1 << 381) - 1
    if v >= P:
        v &= (1 << 380) - 1      # unpinned branch (A.7)
    return v

def map_to_g2(msg: bytes):
    t0 = hash_to_fp(msg)
    a = sw_map_fp2((t0, 0))
    if a is None:
        return None
    return g2_clear_cofactor(pt_from_affine(FP2, *a))

# ---------------------------------------------------------------- serialisation (A.5)
def fr_from_bytes(b: bytes):
    assert len(b) == 32
    v = int.from_bytes(b, 'little')
    if v >= R: return None
    return v

def g1_serialize(p) -> bytes:
    a = pt_affine(FP, p)
    if a is None: return bytes(48)
    out = bytearray(a[0].to_bytes(48, 'little'))
    if a[1] & 1: out[47] |= 0x80
    return bytes(out)

def g1_deserialize(b: bytes, check_order=True):
    if len(b) != 48: return None
    if b == bytes(48): return pt_inf(FP)
    odd = (b[47] & 0x80) != 0
    x = int.from_bytes(b, 'little') & ((1 << 383) - 1)
    if x >= P: return None
    y = fp_sqrt((x * x * x + B1) % P)
    if y is None: return None
    if (y & 1) != odd: y = (-y) % P
    p = pt_from_affine(FP, x, y)
    if check_order and not pt_in_subgroup(FP, p): return None
    return p

def g2_serialize(p) -> bytes:
    a = pt_affine(FP2, p)
    if a is None: return bytes(96)
    out = bytearray(a[0][0].to_bytes(48, 'little') + a[0][1].to_bytes(48, 'little'))
    if a[1][0] & 1: out[95] |= 0x80
    return bytes(out)

def g2_deserialize(b: bytes, check_order=True):
    if len(b) != 96: return None
    if b == bytes(96): return pt_inf(FP2)
    odd = (b[95] & 0x80) != 0
    xa = int.from_bytes(b[:48], 'little')
    xb = int.from_bytes(b[48:], 'little') & ((1 << 383) - 1)
    if xa >= P or xb >= P: return None
    x = (xa, xb)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B2))
    if y is None: return None
    if (y[0] & 1) != odd: y = f2_neg(y)
    p = pt_from_affine(FP2, x, y)
    if check_order and not pt_in_subgroup(FP2, p): return None
    return p

# ---------------------------------------------------------------- pairing (textbook affine Miller loop; values never exposed, only == 1)
def _line(T, Q2, Pa):
    """Line through twist points T, Q2 (affine Fp2) evaluated at G1 affine Pa, scaled by w^3 (killed by FE).
    l = yP*w^3 - lam*xP*w^2 + (lam*xT - yT)."""
    (xt, yt) = T
    if T == Q2:
        lam = f2_mul(f2_mulfp(f2_sqr(xt), 3), f2_inv(f2_add(yt, yt)))
    else:
        lam = f2_mul(f2_sub(Q2[1], yt), f2_inv(f2_sub(Q2[0], xt)))
    c0 = f2_sub(f2_mul(lam, xt), yt)
    c2 = f2_neg(f2_mulfp(lam, Pa[0]))
    c3 = (Pa[1], 0)
    x3 = f2_sub(f2_sub(f2_sqr(lam), xt), Q2[0])
    y3 = f2_sub(f2_mul(lam, f2_sub(xt, x3)), yt)
    return (c0, F2_ZERO, c2, c3, F2_ZERO, F2_ZERO), (x3, y3)

def miller_loop(pairs):
    """pairs: list of (G1 jacobian, G2 jacobian); identity members contribute 1."""
    aff = []
    for p1, q2 in pairs:
        a = pt_affine(FP, p1); b = pt_affine(FP2, q2)
        if a is None or b is None: continue
        aff.append((a, b))
    f = F12_ONE
    Ts = [b for (_, b) in aff]
    for bit in bin(Z_ABS)[3:]:
        f = f12_sqr(f)
        for k, (a, b) in enumerate(aff):
            l, Ts[k] = _line(Ts[k], Ts[k], a)
            f = f12_mul(l, f)
        if bit == '1':
            for k, (a, b) in enumerate(aff):
                l, Ts[k] = _line(Ts[k], b, a)
                f = f12_mul(l, f)
    return f12_conj(f)    # z < 0

_HARD = (P ** 4 - P ** 2 + 1) // R

def final_exp(f):
    f = f12_mul(f12_conj(f), f12_inv(f))        # ^(p^6-1)
    f = f12_mul(f12_frob2(f), f)                # ^(p^2+1)
    return f12_pow(f, _HARD)

def pairing_product_is_one(pairs):
    return final_exp(miller_loop(pairs)) == F12_ONE

# ---------------------------------------------------------------- BLS API (mirrors ffi/go/bls, SWAP_G)
def get_public_key(sk: int):
    return pt_mul(FP, G1_GEN, sk % R)

def sign_hash(sk: int, h: bytes):
    """SecretKey.SignHash: None when the map fails (Go wrapper returns nil)."""
    Hm = map_to_g2(h)
    if Hm is None: return None
    return pt_mul(FP2, Hm, sk % R)

def verify_hash(sig, pk, h: bytes) -> bool:
    """Sign.VerifyHash(pk, h): e(B, sig) == e(pk, H(h))."""
    if pt_is_inf(FP, pk): return False       # identity public key never verifies (include/hbls.h; unpinned: SURVEY A.7)
    Hm = map_to_g2(h)
    if Hm is None: return False
    return pairing_product_is_one([(G1_GEN, sig), (pt_neg(FP, pk), Hm)])

def aggregate_sigs(sigs):
    """crypto/bls/mask.go:58-64 AggregateSig: fold Sign.Add from the zero value."""
    acc = pt_inf(FP2)
    for s in sigs: acc = pt_add(FP2, acc, s)
    return acc

def mask_aggregate(pubkeys, bitmap: bytes):
    """crypto/bls/mask.go:113-134 SetMask on a fresh mask: sum of pubkeys whose bit (LSB-first) is set."""
    if len(bitmap) != (len(pubkeys) + 7) >> 3:
        raise ValueError("mismatching bitmap lengths")
    acc = pt_inf(FP)
    for i, pk in enumerate(pubkeys):
        if bitmap[i >> 3] & (1 << (i & 7)):
            acc = pt_add(FP, acc, pk)
    return acc

def construct_commit_payload(block_num: int, block_hash: bytes, view_id: int, is_staking: bool) -> bytes:
    """consensus/signature/signature.go:12-24."""
    out = block_num.to_bytes(8, 'little') + block_hash
    if is_staking:
        out += view_id.to_bytes(8, 'little')
    return out

def fast_aggregate_verify(pubkeys, bitmap, sig96: bytes, msg: bytes) -> bool:
    """internal/chain/engine.go:619-642 minus the quorum predicate: decode sig, SetMask, VerifyHash."""
    sig = g2_deserialize(sig96)
    if sig is None: return False
    apk = mask_aggregate(pubkeys, bitmap)
    return verify_hash(sig, apk, msg)

def seeded_bytes(tag: str, index: int, n: int) -> bytes:
    """SHA-256 counter mode over "hbls-bench"||tag||index: deterministic synthetic inputs."""
    out = b''
    ctr = 0
    while len(out) < n:
        out += hashlib.sha256(b"hbls-bench" + tag.encode() + index.to_bytes(8, 'little') + ctr.to_bytes(4, 'little')).digest()
        ctr += 1
    return out[:n]

def seeded_sk(tag: str, index: int) -> int:
    return int.from_bytes(seeded_bytes(tag, index, 32), 'little') % R
