#!/usr/bin/env python
"""tools/vmgen.py -- build-time generator of harmony_b200/csrc/vm_programs.cuh: the warp-cooperative ("latency mode") pairing.

One warp = 16 lane pairs verifies ONE round.  The Miller loop and the final exponentiation are straight-line programs over Fp2
values that live in shared-memory slots; a program is a list of STEPS, and in every step each lane pair executes at most one Fp2
operation (pair p runs instruction p of the step), so up to 16 independent Fp2 products run side by side:
    MUL   dst = a * b                     (lane-pair product: 2 wide products + 1 reduction per lane, tower.cuh fp2h)
    SQR   dst = a * a                     (1 wide product + 1 reduction per lane)
    LIN   dst = sum_k M_k * src_k         (M_k = 2x2 small-integer matrices over (re, im): add, sub, conj, xi*, i*, 2x, 3x ...;
                                           stored per lane role as a list of <= 8 ATOMS (slot, half, sign) that the lane adds up, so
                                           that all 32 lanes of a step run the same branch-free loop)
This file holds the formulas (the same tower / Miller-loop / final-exponentiation formulas as tower.cuh and pairing.cuh, written once
over an abstract Fp2 type), a tracer that turns them into a dependency graph, a list scheduler (<= 16 operations of one class per
step, linear combinations folded so that at most one LIN level sits between two MUL levels), a slot allocator, the C emitter and
a reference interpreter (`run_program`) that tests/test_vm_programs.py uses to check every program against the CPU oracle.

Run:  python tools/vmgen.py            (rewrites harmony_b200/csrc/vm_programs.cuh; the generated header is committed)
"""
import os, sys

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
Z_ABS = 0xd201000000010000
NPAIR = 16
MAXA = 8            # atoms per row (re / im) of a LIN instruction: an atom adds or subtracts m x ONE half of ONE slot, m in 1..4
MAXW = 24           # total weight (sum of the m) of a row: the lane accumulates plain integers < 32 p (13 limbs) before ONE reduction
CMAX = 7            # |matrix entry| limit of a LIN term (5, 6, 7 take two atoms)

# ------------------------------------------------------------------ concrete Fp2 (reference semantics of the VM)
def f2(a, b=0): return (a % P, b % P)
def f2_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def f2_mul(x, y): return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def f2_apply(M, x): return ((M[0] * x[0] + M[1] * x[1]) % P, (M[2] * x[0] + M[3] * x[1]) % P)
def f2_pow(x, e):
    r = (1, 0)
    while e:
        if e & 1: r = f2_mul(r, x)
        x = f2_mul(x, x); e >>= 1
    return r
XI = (1, 1)
CONSTS = {"ZERO": (0, 0), "ONE": (1, 0), "TWIST3B": (12, 12), "INV2": (pow(2, -1, P), 0)}       # TWIST3B = 3 b' = 3 * 4(1 + i)
# psi endomorphism of E'(Fp2) (oracle/pyref.py PSI_CX / PSI_CY; curve.cuh g2_psi): psi(x, y) = (conj(x) cx, conj(y) cy)
def f2_inv(x): return f2_pow(x, P * P - 2)
CONSTS["PSI_CX"] = f2_inv(f2_pow(XI, (P - 1) // 3)); CONSTS["PSI_CY"] = f2_inv(f2_pow(XI, (P - 1) // 2))
CONSTS["PSI2_CX"] = f2_mul(CONSTS["PSI_CX"], (CONSTS["PSI_CX"][0], -CONSTS["PSI_CX"][1])); assert CONSTS["PSI2_CX"][1] == 0
for k in range(6):
    CONSTS[f"FROB1_{k}"] = f2_pow(XI, k * (P - 1) // 6)
    g2 = f2_pow(XI, k * (P * P - 1) // 6); assert g2[1] == 0
    CONSTS[f"FROB2_{k}"] = g2

# ------------------------------------------------------------------ tracer
I2 = (1, 0, 0, 1)
def m_mul(A, B): return (A[0] * B[0] + A[1] * B[2], A[0] * B[1] + A[1] * B[3], A[2] * B[0] + A[3] * B[2], A[2] * B[1] + A[3] * B[3])
def m_add(A, B): return tuple(a + b for a, b in zip(A, B))
M_CONJ, M_XI, M_NEG = (1, 0, 0, -1), (1, -1, 1, 1), (-1, 0, 0, -1)

class Graph:
    """nodes: ('in', name) | ('const', name) | ('mul', a, b) | ('sqr', a) | ('lin', ((node, M), ...))"""
    def __init__(self, name):
        self.name = name; self.nodes = []; self.cse = {}; self.outputs = []; self.inputs = {}
    def add(self, node):
        key = node
        if key in self.cse: return self.cse[key]
        self.nodes.append(node); self.cse[key] = len(self.nodes) - 1
        return len(self.nodes) - 1
    def inp(self, reg):
        i = self.add(("in", reg)); self.inputs[reg] = i
        return V(self, {i: I2})
    def const(self, name): return V(self, {self.add(("const", name)): I2})
    def out(self, reg, v): self.outputs.append((reg, v.node()))

class V:
    """symbolic Fp2 value = sum of M_n * node_n"""
    def __init__(self, g, terms): self.g = g; self.t = {n: M for n, M in terms.items() if M != (0, 0, 0, 0)}
    def _lin(self, M): return V(self.g, {n: m_mul(M, X) for n, X in self.t.items()})
    def __add__(self, o):
        t = dict(self.t)
        for n, M in o.t.items(): t[n] = m_add(t[n], M) if n in t else M
        return V(self.g, t)
    def __neg__(self): return self._lin(M_NEG)
    def __sub__(self, o): return self + (-o)
    def scale(self, k): return self._lin((k, 0, 0, k))
    def dbl(self): return self.scale(2)
    def conj(self): return self._lin(M_CONJ)
    def xi(self): return self._lin(M_XI)
    def node(self):
        """materialise: a node index holding exactly this value"""
        g = self.g
        if len(self.t) == 1:
            (n, M), = self.t.items()
            if M == I2: return n
        items = sorted(self.t.items())
        if not items: items = [(g.add(("const", "ONE")), (0, 0, 0, 0))]
        # split so that both rows (re: m00, m01; im: m10, m11) of every LIN instruction stay within MAXA atoms
        def fits(M): return all(abs(c) <= CMAX for c in M)
        for n, M in items:
            if not fits(M): raise ValueError(f"{g.name}: coefficient out of range {M}")
        def na(c): return 0 if c == 0 else (1 if abs(c) <= 4 else 2)
        def cost(ts):      # (atoms, weight) of the heavier row
            return (max(sum(na(M[0]) + na(M[1]) for _, M in ts), sum(na(M[2]) + na(M[3]) for _, M in ts)),
                    max(sum(abs(M[0]) + abs(M[1]) for _, M in ts), sum(abs(M[2]) + abs(M[3]) for _, M in ts)))
        def ok(ts): a, w = cost(ts); return a <= MAXA and w <= MAXW
        while not ok(items):
            head = []
            while items and ok(head + [items[0]]): head.append(items.pop(0))
            if not head or (len(head) == 1 and head[0][1] == I2): raise ValueError(f"{g.name}: cannot split a linear combination into LIN instructions")
            items.insert(0, (g.add(("lin", tuple(head))), I2))
        return g.add(("lin", tuple(items)))
    def __mul__(self, o):
        a, b = self.node(), o.node()
        if a == b: return V(self.g, {self.g.add(("sqr", a)): I2})
        if a > b: a, b = b, a
        return V(self.g, {self.g.add(("mul", a, b)): I2})
    def sqr(self): return self * self

class C:
    """concrete twin of V (same interface) used to validate formulas and the interpreter"""
    def __init__(self, v): self.v = (v[0] % P, v[1] % P)
    def __add__(self, o): return C(f2_add(self.v, o.v))
    def __neg__(self): return C((-self.v[0], -self.v[1]))
    def __sub__(self, o): return self + (-o)
    def scale(self, k): return C((self.v[0] * k, self.v[1] * k))
    def dbl(self): return self.scale(2)
    def conj(self): return C((self.v[0], -self.v[1]))
    def xi(self): return C(f2_apply(M_XI, self.v))
    def __mul__(self, o): return C(f2_mul(self.v, o.v))
    def sqr(self): return self * self

# ------------------------------------------------------------------ formulas over the abstract Fp2 type (V or C).  k(name) = constant
# Fp6 = (c0, c1, c2) over v^3 = xi; Fp12 = (c0, c1) over w^2 = v; fp12 as a 6-tuple in tower order (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2)
def fp6_add(x, y): return tuple(a + b for a, b in zip(x, y))
def fp6_sub(x, y): return tuple(a - b for a, b in zip(x, y))
def fp6_neg(x): return tuple(-a for a in x)
def fp6_mul_v(x): return (x[2].xi(), x[0], x[1])
def fp6_mul(x, y):
    v0, v1, v2 = x[0] * y[0], x[1] * y[1], x[2] * y[2]
    t0 = ((x[1] + x[2]) * (y[1] + y[2]) - v1 - v2).xi() + v0
    t1 = (x[0] + x[1]) * (y[0] + y[1]) - v0 - v1 + v2.xi()
    t2 = (x[0] + x[2]) * (y[0] + y[2]) - v0 - v2 + v1
    return (t0, t1, t2)
def fp6_mul_by_01(x, b0, b1):
    v0, v1 = x[0] * b0, x[1] * b1
    t0 = (x[2] * b1).xi() + v0
    t1 = (x[0] + x[1]) * (b0 + b1) - v0 - v1
    t2 = x[2] * b0 + v1
    return (t0, t1, t2)
def fp6_mul_by_1(x, b1): return ((x[2] * b1).xi(), x[0] * b1, x[1] * b1)
def fp6_inv_parts(x):
    """(t0, t1, t2, d): x^-1 = (t0, t1, t2) / d with d in Fp2"""
    t0 = x[0].sqr() - (x[1] * x[2]).xi()
    t1 = x[2].sqr().xi() - x[0] * x[1]
    t2 = x[1].sqr() - x[0] * x[2]
    d = (x[2] * t1 + x[1] * t2).xi() + x[0] * t0
    return t0, t1, t2, d
def fp12_split(f): return f[0:3], f[3:6]
def fp12_mul(x, y):
    x0, x1 = fp12_split(x); y0, y1 = fp12_split(y)
    v0, v1 = fp6_mul(x0, y0), fp6_mul(x1, y1)
    s = fp6_sub(fp6_sub(fp6_mul(fp6_add(x0, x1), fp6_add(y0, y1)), v0), v1)
    return fp6_add(v0, fp6_mul_v(v1)) + s
def fp12_sqr(x):
    x0, x1 = fp12_split(x)
    ab = fp6_mul(x0, x1)
    s = fp6_mul(fp6_add(x0, x1), fp6_add(fp6_mul_v(x1), x0))
    c0 = fp6_sub(fp6_sub(s, ab), fp6_mul_v(ab))
    return c0 + fp6_add(ab, ab)
def fp12_conj(x): return tuple(x[0:3]) + fp6_neg(x[3:6])
def fp12_mul_by_014(x, o0, o1, o4):
    x0, x1 = fp12_split(x)
    aa = fp6_mul_by_01(x0, o0, o1)
    bb = fp6_mul_by_1(x1, o4)
    s = fp6_sub(fp6_sub(fp6_mul_by_01(fp6_add(x0, x1), o0, o1 + o4), aa), bb)
    return fp6_add(aa, fp6_mul_v(bb)) + s
# coefficient of w^k (k = 2i + j) <-> tower slot index
SLOT_OF_W = [0, 3, 1, 4, 2, 5]
def fp12_frob(x, k):
    r = list(x)
    for w in range(6):
        i = SLOT_OF_W[w]; r[i] = x[i].conj() * k(f"FROB1_{w}")
    return tuple(r)
def fp12_frob2(x, k):
    r = list(x)
    for w in range(6):
        i = SLOT_OF_W[w]; r[i] = x[i] * k(f"FROB2_{w}")
    return tuple(r)
def fp4_sqr(a, b):
    t0, t1 = a.sqr(), b.sqr()
    return t1.xi() + t0, (a + b).sqr() - t0 - t1
def fp12_cyc_sqr(x, xn=None):
    """xn (optional): the same six values as single materialised nodes.  Squarings and the (a + b) sums use x (expansions in terms
    of the previous squarings, so that they sit ONE linear level after them); the 3 t -+ 2 z outputs use xn (small coefficients)."""
    z0, z4, z3, z2, z1, z5 = x
    if xn is not None:
        n0, n4, n3, n2, n1, n5 = xn
        t0, t1 = fp4_sqr(z0, z1)
        r0 = (t0 - n0).dbl() + t0; r1 = (t1 + n1).dbl() + t1
        t0, t1 = fp4_sqr(z2, z3); t2, t3 = fp4_sqr(z4, z5)
        r4 = (t0 - n4).dbl() + t0; r5 = (t1 + n5).dbl() + t1
        t0 = t3.xi()
        r2 = (t0 + n2).dbl() + t0; r3 = (t2 - n3).dbl() + t2
        return (r0, r4, r3, r2, r1, r5)
    t0, t1 = fp4_sqr(z0, z1)
    z0 = (t0 - z0).dbl() + t0; z1 = (t1 + z1).dbl() + t1
    t0, t1 = fp4_sqr(z2, z3); t2, t3 = fp4_sqr(z4, z5)
    z4 = (t0 - z4).dbl() + t0; z5 = (t1 + z5).dbl() + t1
    t0 = t3.xi()
    z2 = (t0 + z2).dbl() + t0; z3 = (t2 - z3).dbl() + t2
    return (z0, z4, z3, z2, z1, z5)
def ml_dbl(T, k):
    x, y, z = T
    A = (x * y) * k("INV2")
    B, Cc = y.sqr(), z.sqr()
    E = k("TWIST3B") * Cc
    F = E.scale(3)
    H = (y + z).sqr() - B - Cc
    l0 = B - E
    l2 = -(x.sqr().scale(3))
    l3 = H
    x3 = (B - F) * A
    y3 = ((B + F) * k("INV2")).sqr() - E.sqr().scale(3)
    return (x3, y3, B * H), (l0, l2, l3)
def ml_add(T, qx, qy):
    x, y, z = T
    th = y - qy * z
    mu = x - qx * z
    l0 = th * qx - mu * qy
    Cc, D = th.sqr(), mu.sqr()
    E = mu * D; F = z * Cc; G = x * D
    H = E + F - G.dbl()
    x3 = mu * H
    y3 = (G - H) * th - E * y
    return (x3, y3, z * E), (l0, -th, mu)
def line_at(l, px, py): return l[0], l[1] * px, l[2] * py
# ---- G2 group law in Jacobian coordinates (a = 0): the cofactor clearing of hash-to-G2 on the VM (curve.cuh g2_clear_cofactor).
# Generic formulas: a degenerate addition (equal / opposite operands, an identity) yields Z = 0, which stays 0 through every later
# operation -- the kernel then redoes the message with the complete lane-pair code.
def g2_dbl(T):
    x, y, z = T
    A, B = x.sqr(), y.sqr()
    Cc, C4 = B.sqr(), B.dbl().sqr()                  # Y^4 and 4 Y^4 (coefficients of a linear combination stay <= 7)
    S = (x + B).sqr()
    E = A.scale(3); F = E.sqr()
    x3 = F - S.scale(4) + A.scale(4) + C4              # F - 2 D,  D = 2 (S - A - C)
    W = S.scale(6) - A.scale(6) - Cc.scale(6) - F      # D - x3 = 3 D - F
    return (x3, E * W - C4.dbl(), (y * z).dbl())
def g2_add(T, Q):
    x1, y1, z1 = T; x2, y2, z2 = Q
    z1z1, z2z2 = z1.sqr(), z2.sqr()
    u1, u2 = x1 * z2z2, x2 * z1z1
    s1, s2 = (y1 * z2) * z2z2, (y2 * z1) * z1z1
    h = u2 - u1
    i = h.dbl().sqr(); j = h * i
    r = (s2 - s1).dbl()
    v = u1 * i
    x3 = r.sqr() - j - v.dbl()
    return (x3, r * (v - x3) - (s1 * j).dbl(), ((z1 + z2).sqr() - z1z1 - z2z2) * h)
def g2_neg(T): return (T[0], -T[1], T[2])
def g2_psi(T, k): return (T[0].conj() * k("PSI_CX"), T[1].conj() * k("PSI_CY"), T[2].conj())
def g2_psi2(T, k): return (T[0] * k("PSI2_CX"), -T[1], T[2])

# ------------------------------------------------------------------ programs.  Persistent registers (fixed slots, shared by all programs)
REG_CONST = ["ZERO", "ONE", "TWIST3B", "INV2"] + [f"FROB1_{k}" for k in range(6)] + [f"FROB2_{k}" for k in range(6)] + ["PSI_CX", "PSI_CY", "PSI2_CX"]
REG_IN = ["P1X", "P1Y", "Q1X", "Q1Y", "P2X", "P2Y", "Q2X", "Q2Y"]
def r6(n): return [f"{n}{i}" for i in range(6)]
REG_STATE = ["T1X", "T1Y", "T1Z", "T2X", "T2Y", "T2Z"] + r6("F") + r6("M") + r6("X") + r6("ACC") + r6("A") + r6("B") + r6("C") + \
            ["IT0", "IT1", "IT2", "ID", "NORM", "NINV", "HX", "HY"]
REGS = REG_CONST + REG_IN + REG_STATE
assert len(set(REGS)) == len(REGS), "duplicate register name"
SLOT = {r: i for i, r in enumerate(REGS)}
NSLOTS = 150        # slots per warp (100 B each): 83 registers + temporaries; lowest-free-first allocation keeps every program under it (15 KB per warp)

def build_programs(make_graph):
    """every program as a function of a fresh graph; returns {name: graph}"""
    progs = {}
    def prog(fn):
        g = make_graph(fn.__name__[2:].upper()); k = g.const
        fn(g, k); progs[g.name] = g
        return fn
    def get6(g, n): return tuple(g.inp(r) for r in r6(n))
    def put6(g, n, v):
        for r, x in zip(r6(n), v): g.out(r, x)
    @prog
    def p_ml_init(g, k):
        one = k("ONE")
        for t, q in (("T1", "Q1"), ("T2", "Q2")):
            g.out(t + "X", g.inp(q + "X")); g.out(t + "Y", g.inp(q + "Y")); g.out(t + "Z", one)
        put6(g, "F", (one,) + tuple(one.scale(0) for _ in range(5)))
    @prog
    def p_ml_dbl(g, k):
        f = fp12_sqr(get6(g, "F"))
        for t, pp in (("T1", "P1"), ("T2", "P2")):
            T, l = ml_dbl((g.inp(t + "X"), g.inp(t + "Y"), g.inp(t + "Z")), k)
            f = fp12_mul_by_014(f, *line_at(l, g.inp(pp + "X"), g.inp(pp + "Y")))
            g.out(t + "X", T[0]); g.out(t + "Y", T[1]); g.out(t + "Z", T[2])
        put6(g, "F", f)
    @prog
    def p_ml_dbl2(g, k):          # two doubling iterations in one program (no add step between them)
        f = get6(g, "F"); T = {t: (g.inp(t + "X"), g.inp(t + "Y"), g.inp(t + "Z")) for t in ("T1", "T2")}
        for _ in range(2):
            f = fp12_sqr(f)
            for t, pp in (("T1", "P1"), ("T2", "P2")):
                T[t], l = ml_dbl(T[t], k)
                f = fp12_mul_by_014(f, *line_at(l, g.inp(pp + "X"), g.inp(pp + "Y")))
            f = tuple(V(g, {x.node(): I2}) for x in f); T = {t: tuple(V(g, {x.node(): I2}) for x in T[t]) for t in T}
        for t in ("T1", "T2"):
            g.out(t + "X", T[t][0]); g.out(t + "Y", T[t][1]); g.out(t + "Z", T[t][2])
        put6(g, "F", f)
    @prog
    def p_ml_add(g, k):
        f = get6(g, "F")
        for t, pp, q in (("T1", "P1", "Q1"), ("T2", "P2", "Q2")):
            T, l = ml_add((g.inp(t + "X"), g.inp(t + "Y"), g.inp(t + "Z")), g.inp(q + "X"), g.inp(q + "Y"))
            f = fp12_mul_by_014(f, *line_at(l, g.inp(pp + "X"), g.inp(pp + "Y")))
            g.out(t + "X", T[0]); g.out(t + "Y", T[1]); g.out(t + "Z", T[2])
        put6(g, "F", f)
    # single-pair forms on the pair-2 registers and a running Fp12 product in A: the multi-GPU split of ONE batch lets
    # every warp multiply the Miller values of its items into A (no final exponentiation per item); the fold multiplies the gathered
    # partial products into F before ONE final exponentiation.
    @prog
    def p_ml1_init(g, k):
        one = k("ONE")
        g.out("T2X", g.inp("Q2X")); g.out("T2Y", g.inp("Q2Y")); g.out("T2Z", one)
        put6(g, "F", (one,) + tuple(one.scale(0) for _ in range(5)))
    @prog
    def p_ml1_dbl(g, k):
        f = fp12_sqr(get6(g, "F"))
        T, l = ml_dbl((g.inp("T2X"), g.inp("T2Y"), g.inp("T2Z")), k)
        f = fp12_mul_by_014(f, *line_at(l, g.inp("P2X"), g.inp("P2Y")))
        g.out("T2X", T[0]); g.out("T2Y", T[1]); g.out("T2Z", T[2]); put6(g, "F", f)
    @prog
    def p_ml1_add(g, k):
        T, l = ml_add((g.inp("T2X"), g.inp("T2Y"), g.inp("T2Z")), g.inp("Q2X"), g.inp("Q2Y"))
        f = fp12_mul_by_014(get6(g, "F"), *line_at(l, g.inp("P2X"), g.inp("P2Y")))
        g.out("T2X", T[0]); g.out("T2Y", T[1]); g.out("T2Z", T[2]); put6(g, "F", f)
    @prog
    def p_a_one(g, k):
        one = k("ONE"); put6(g, "A", (one,) + tuple(one.scale(0) for _ in range(5)))
    @prog
    def p_amulf(g, k): put6(g, "A", fp12_mul(get6(g, "A"), get6(g, "F")))
    @prog
    def p_fmula(g, k): put6(g, "F", fp12_mul(get6(g, "F"), get6(g, "A")))
    # final exponentiation, easy part: m = conj(f) * f^-1, m = frob2(m) * m.  The Fp12 inverse goes down to ONE Fp inversion
    # (of NORM = N(d), d in Fp2), which a single lane computes between FE_INV_A and FE_INV_B.
    @prog
    def p_fe_inv_a(g, k):
        f = get6(g, "F"); f0, f1 = fp12_split(f)
        t = fp6_sub(fp6_mul(f0, f0), fp6_mul_v(fp6_mul(f1, f1)))        # f0^2 - v f1^2 in Fp6
        t0, t1, t2, d = fp6_inv_parts(t)
        g.out("IT0", t0); g.out("IT1", t1); g.out("IT2", t2); g.out("ID", d)
        g.out("NORM", d * d.conj())                                      # (re^2 + im^2, 0)
    @prog
    def p_fe_inv_b(g, k):
        f = get6(g, "F"); f0, f1 = fp12_split(f)
        dinv = g.inp("ID").conj() * g.inp("NINV")                        # d^-1 = conj(d) / N(d)
        tinv = tuple(g.inp(r) * dinv for r in ("IT0", "IT1", "IT2"))     # (f0^2 - v f1^2)^-1
        finv = fp6_mul(f0, tinv) + fp6_neg(fp6_mul(f1, tinv))
        m = fp12_mul(fp12_conj(f), finv)
        m = fp12_mul(fp12_frob2(m, k), m)
        put6(g, "M", m); put6(g, "X", m); put6(g, "ACC", m)
    @prog
    def p_cycsqr(g, k): put6(g, "ACC", fp12_cyc_sqr(get6(g, "ACC")))
    # runs of squarings between the set bits of |z| as ONE program each: the linear steps between consecutive squarings fold into one
    def cyc_run(n):
        def f(g, k):
            a = get6(g, "ACC")
            an = a
            for _ in range(n):
                a = fp12_cyc_sqr(a, an)
                an = tuple(V(g, {x.node(): I2}) for x in a)          # the materialised twins keep the coefficients small
            put6(g, "ACC", a)
        f.__name__ = f"p_cycsqr{n}"; return f
    for n in (2, 4, 8, 16): prog(cyc_run(n))
    @prog
    def p_mulx(g, k): put6(g, "ACC", fp12_mul(get6(g, "ACC"), get6(g, "X")))
    # hard part (pairing.cuh final_exp): a = m^(z-1), b = a^(z-1), c = b^(z+p), d = c^(z^2+p^2-1), result d * m^3.
    # exp_z(x) = conj(x^|z|); the kernel runs the 63-step square-and-multiply loop on (X, ACC) between these glue programs.
    @prog
    def p_glue1(g, k):
        a = fp12_mul(fp12_conj(get6(g, "ACC")), fp12_conj(get6(g, "M")))
        put6(g, "A", a); put6(g, "X", a); put6(g, "ACC", a)
    @prog
    def p_glue2(g, k):
        b = fp12_mul(fp12_conj(get6(g, "ACC")), fp12_conj(get6(g, "A")))
        put6(g, "B", b); put6(g, "X", b); put6(g, "ACC", b)
    @prog
    def p_glue3(g, k):
        c = fp12_mul(fp12_conj(get6(g, "ACC")), fp12_frob(get6(g, "B"), k))
        put6(g, "C", c); put6(g, "X", c); put6(g, "ACC", c)
    @prog
    def p_glue4(g, k):
        t = fp12_conj(get6(g, "ACC"))
        put6(g, "X", t); put6(g, "ACC", t)
    @prog
    def p_glue5(g, k):            # d = c^(z^2) * frob2(c) * conj(c), kept in B
        c = get6(g, "C")
        put6(g, "B", fp12_mul(fp12_mul(fp12_conj(get6(g, "ACC")), fp12_frob2(c, k)), fp12_conj(c)))
    @prog
    def p_glue6(g, k):            # result = d * m^3  (two short programs instead of one: the slot working set stays under 128)
        m = get6(g, "M")
        put6(g, "ACC", fp12_mul(get6(g, "B"), fp12_mul(fp12_cyc_sqr(m), m)))
    # ---- hash-to-G2 cofactor clearing (Budroni-Pintore, curve.cuh g2_clear_cofactor): h(P) = [z^2 - z - 1]P + psi([z - 1]P) + psi^2([2]P)
    # with P = (Q2X, Q2Y) the affine Shallue-van de Woestijne point.  [|z|] = 63 doublings of T1 (run programs like the cyclotomic
    # squarings) + 5 additions of T2.  Result affine in (HX, HY) around ONE Fp inversion (NORM -> NINV, like the final exponentiation).
    def getT(g, t): return (g.inp(t + "X"), g.inp(t + "Y"), g.inp(t + "Z"))
    def putT(g, t, v):
        for c, x in zip("XYZ", v): g.out(t + c, x)
    @prog
    def p_g2_init(g, k):
        pt = (g.inp("Q2X"), g.inp("Q2Y"), k("ONE")); putT(g, "T1", pt); putT(g, "T2", pt)
    def g2_dbl_run(n):
        def f(g, k):
            T = getT(g, "T1")
            for _ in range(n):
                T = g2_dbl(T); T = tuple(V(g, {x.node(): I2}) for x in T)
            putT(g, "T1", T)
        f.__name__ = "p_g2dbl" + (str(n) if n > 1 else ""); return f
    for n in (1, 2, 4, 8, 16): prog(g2_dbl_run(n))
    @prog
    def p_g2_add(g, k): putT(g, "T1", g2_add(getT(g, "T1"), getT(g, "T2")))
    @prog
    def p_hc_mid(g, k):           # T1 = [|z|]P  ->  zp = [z]P = -T1 (z < 0), kept in M0..2; next chain starts from zp
        zp = g2_neg(getT(g, "T1"))
        for r, x in zip(("M0", "M1", "M2"), zp): g.out(r, x)
        putT(g, "T1", zp); putT(g, "T2", zp)
    @prog
    def p_hc_fin(g, k):           # T1 = [|z|]zp -> z2p = -T1
        z2p = g2_neg(getT(g, "T1"))
        zp = (g.inp("M0"), g.inp("M1"), g.inp("M2"))
        pt = (g.inp("Q2X"), g.inp("Q2Y"), k("ONE")); npt = g2_neg(pt)
        def mat(T): return tuple(V(g, {x.node(): I2}) for x in T)
        t1 = mat(g2_add(mat(g2_add(z2p, npt)), g2_neg(zp)))
        t2 = mat(g2_psi(mat(g2_add(zp, npt)), k))
        t3 = mat(g2_psi2(mat(g2_dbl(pt)), k))
        putT(g, "T1", g2_add(mat(g2_add(t1, t2)), t3))
    @prog
    def p_g2_norm(g, k):
        z = g.inp("T1Z"); g.out("NORM", z * z.conj())
    @prog
    def p_g2_aff(g, k):
        zi = g.inp("T1Z").conj() * g.inp("NINV"); zi2 = zi.sqr()
        g.out("HX", g.inp("T1X") * zi2); g.out("HY", g.inp("T1Y") * (zi2 * zi))
    return progs

# ------------------------------------------------------------------ scheduling + slot allocation
OP_MUL, OP_SQR, OP_LIN = 1, 2, 3
class Program:
    def __init__(self, name): self.name = name; self.steps = []      # step = (cls, [instr...]); instr = (dst, a, b) | (dst, ((slot, M), ...))

def compile_graph(g):
    nodes = g.nodes
    # live nodes (reachable from the outputs)
    need = set(); stack = [n for _, n in g.outputs]
    def deps(n):
        nd = nodes[n]
        if nd[0] == "mul": return [nd[1], nd[2]]
        if nd[0] == "sqr": return [nd[1]]
        if nd[0] == "lin": return [s for s, _ in nd[1]]
        return []
    while stack:
        n = stack.pop()
        if n in need: continue
        need.add(n); stack += deps(n)
    ops = [n for n in sorted(need) if nodes[n][0] in ("mul", "sqr", "lin")]
    users = {n: [] for n in need}
    for n in ops:
        for d in deps(n): users[d].append(n)
    # critical-path priority (MUL/SQR weigh 8, LIN 1)
    w = {n: (8 if nodes[n][0] != "lin" else 1) for n in ops}
    prio = {}
    for n in reversed(ops): prio[n] = w[n] + max([prio[u] for u in users[n]] + [0])
    def schedule(mul_first):
        done = set(n for n in need if nodes[n][0] in ("in", "const"))
        pending = set(ops); steps = []
        while pending:
            ready = [n for n in pending if all(d in done for d in deps(n))]
            lin = sorted([n for n in ready if nodes[n][0] == "lin"], key=lambda n: -prio[n])
            mul = sorted([n for n in ready if nodes[n][0] != "lin"], key=lambda n: -prio[n])
            # mul_first: products as soon as any is ready, so that the linear work in between piles up into few steps;
            # otherwise cheap steps first (they unlock products and fill the product steps better)
            if mul and (mul_first or not lin):
                chosen = mul[:NPAIR]
                cls = OP_SQR if all(nodes[n][0] == "sqr" for n in chosen) else OP_MUL
            else:
                chosen = lin[:NPAIR]; cls = OP_LIN
            steps.append((cls, chosen)); done |= set(chosen); pending -= set(chosen)
        return steps
    def cost(steps):        # rough step times on the device (us): product 1.45, squaring 1.0, linear 0.6
        return sum({OP_MUL: 1.45, OP_SQR: 1.0, OP_LIN: 0.6}[c] for c, _ in steps)
    steps = min((schedule(False), schedule(True)), key=cost)
    done = set(n for n in need if nodes[n][0] in ("in", "const"))
    # final move step(s): outputs into their registers (a register may still be read as an input until the very end)
    # slot allocation over steps: value of node n lives from its defining step to its last use
    nstep = len(steps)
    defstep = {n: -1 for n in done if nodes[n][0] in ("in", "const")}
    for si, (_, ch) in enumerate(steps):
        for n in ch: defstep[n] = si
    lastuse = {n: defstep[n] for n in need}
    for n in ops:
        for d in deps(n): lastuse[d] = max(lastuse[d], defstep[n])
    for _, n in g.outputs: lastuse[n] = nstep            # read by the move step
    slot = {}
    for n in need:
        if nodes[n][0] in ("in", "const"): slot[n] = SLOT[nodes[n][1]]
    import heapq
    free = list(range(len(REGS), NSLOTS)); heapq.heapify(free); release = {}   # step -> slots that become free AFTER that step
    for si, (_, ch) in enumerate(steps):
        for s in release.pop(si - 1, []): heapq.heappush(free, s)
        for n in ch:
            if not free: raise RuntimeError(f"{g.name}: out of slots")
            slot[n] = heapq.heappop(free)                   # lowest free slot: the working set stays compact
            release.setdefault(lastuse[n], []).append(slot[n])
    prog = Program(g.name)
    for cls, ch in steps:
        ins = []
        for n in ch:
            nd = nodes[n]
            if nd[0] == "mul": ins.append((slot[n], slot[nd[1]], slot[nd[2]]))
            elif nd[0] == "sqr": ins.append((slot[n], slot[nd[1]], slot[nd[1]]))
            else: ins.append((slot[n], tuple((slot[s], M) for s, M in nd[1])))
        prog.steps.append((cls, ins))
    moves = [(SLOT[reg], ((slot[n], I2),)) for reg, n in g.outputs if SLOT[reg] != slot[n]]
    for i in range(0, len(moves), NPAIR): prog.steps.append((OP_LIN, moves[i:i + NPAIR]))
    # the moves read temporaries (or registers that no move writes) and write registers: no read-after-write hazard inside or across
    # the move steps (pairs of one step are not ordered against each other on the device)
    srcs = {t[0][0] for _, t in moves}; dsts = {d for d, _ in moves}
    if srcs & dsts: raise RuntimeError(f"{g.name}: move hazard on slots {sorted(srcs & dsts)}")
    # and no step may write a slot that the same step reads
    for cls, ins in prog.steps:
        rd = set()
        for i in ins: rd |= ({s_ for s_, _ in i[1]} if cls == OP_LIN else {i[1], i[2]})
        if rd & {i[0] for i in ins}: raise RuntimeError(f"{g.name}: a step writes a slot it reads")
    return prog

def run_program(prog, slots):
    """reference interpreter: slots = list of (re, im) ints mod P (plain, not Montgomery).  All instructions of a step read
    before any writes (the device separates steps by __syncwarp and never lets a step write a slot it reads)."""
    for cls, ins in prog.steps:
        res = []
        for i in ins:
            if cls == OP_LIN:
                acc = (0, 0)
                for s, M in i[1]: acc = f2_add(acc, f2_apply(M, slots[s]))
                res.append((i[0], acc))
            else: res.append((i[0], f2_mul(slots[i[1]], slots[i[2]])))
        for d, v in res: slots[d] = v
    return slots

def compile_all():
    graphs = build_programs(Graph)
    return {n: compile_graph(g) for n, g in graphs.items()}

# ------------------------------------------------------------------ whole pairing check with the reference interpreter (tests)
def fresh_slots():
    s = [(0, 0)] * NSLOTS
    for r in REG_CONST: s[SLOT[r]] = CONSTS[r]
    return s
def fp_inv(a): return pow(a, P - 2, P)
def vm_pairing_is_one(progs, p1, q1, p2, q2):
    """e(p1, q1) e(p2, q2) == 1 via the VM programs; p = (x, y) in Fp, q = ((x0, x1), (y0, y1)) in Fp2 (affine, not identity)"""
    s = fresh_slots()
    for reg, v in (("P1X", (p1[0], 0)), ("P1Y", (p1[1], 0)), ("Q1X", q1[0]), ("Q1Y", q1[1]), ("P2X", (p2[0], 0)), ("P2Y", (p2[1], 0)), ("Q2X", q2[0]), ("Q2Y", q2[1])):
        s[SLOT[reg]] = (v[0] % P, v[1] % P)
    miller2_vm(progs, s)
    final_exp_vm(progs, s)
    return [s[SLOT[r]] for r in r6("ACC")] == [(1, 0)] + [(0, 0)] * 5
def miller2_vm(progs, s):
    """the device's loop (vm.cuh vm_pairing_check): doubling iterations in pairs where no addition step separates them"""
    run_program(progs["ML_INIT"], s)
    i = 62
    while i >= 0:
        if (Z_ABS >> i) & 1: run_program(progs["ML_DBL"], s); run_program(progs["ML_ADD"], s); i -= 1
        elif i >= 1 and not (Z_ABS >> (i - 1)) & 1: run_program(progs["ML_DBL2"], s); i -= 2
        else: run_program(progs["ML_DBL"], s); i -= 1
def vm_miller1(progs, s, p, q):
    """F <- f_{|z|,q}(p) (single pair on the pair-2 registers)"""
    for reg, v in (("P2X", (p[0], 0)), ("P2Y", (p[1], 0)), ("Q2X", q[0]), ("Q2Y", q[1])): s[SLOT[reg]] = (v[0] % P, v[1] % P)
    run_program(progs["ML1_INIT"], s)
    for i in range(62, -1, -1):
        run_program(progs["ML1_DBL"], s)
        if (Z_ABS >> i) & 1: run_program(progs["ML1_ADD"], s)
def vm_product_is_one(progs, pairs_per_part):
    """the split protocol: every part multiplies the Miller values of its pairs into A; the fold multiplies the parts' A into F"""
    parts = []
    for pairs in pairs_per_part:
        s = fresh_slots(); run_program(progs["A_ONE"], s)
        for p, q in pairs: vm_miller1(progs, s, p, q); run_program(progs["AMULF"], s)
        parts.append([s[SLOT[r]] for r in r6("A")])
    s = fresh_slots()
    for i, r in enumerate(r6("F")): s[SLOT[r]] = (1, 0) if i == 0 else (0, 0)
    for a in parts:
        for r, v in zip(r6("A"), a): s[SLOT[r]] = v
        run_program(progs["FMULA"], s)
    final_exp_vm(progs, s)
    return [s[SLOT[r]] for r in r6("ACC")] == [(1, 0)] + [(0, 0)] * 5
def vm_clear_cofactor(progs, pt):
    """hash-to-G2 cofactor clearing on the VM programs: pt = affine ((x0, x1), (y0, y1)) -> affine h(pt), or None when the generic
    formulas degenerate (Z = 0)"""
    s = fresh_slots()
    s[SLOT["Q2X"]], s[SLOT["Q2Y"]] = pt
    def zmul():
        for run, add in expz_schedule():
            for n in (16, 8, 4, 2, 1):
                while run >= n: run_program(progs["G2DBL" if n == 1 else f"G2DBL{n}"], s); run -= n
            if add: run_program(progs["G2_ADD"], s)
    run_program(progs["G2_INIT"], s); zmul()
    run_program(progs["HC_MID"], s); zmul()
    run_program(progs["HC_FIN"], s); run_program(progs["G2_NORM"], s)
    n = s[SLOT["NORM"]]; assert n[1] == 0
    if n[0] == 0: return None
    s[SLOT["NINV"]] = (fp_inv(n[0]), 0)
    run_program(progs["G2_AFF"], s)
    return s[SLOT["HX"]], s[SLOT["HY"]]
def expz_schedule():
    """x^|z| by square-and-multiply from bit 62 down: [(number of squarings, multiply afterwards?), ...]"""
    out = []; run = 0
    for i in range(62, -1, -1):
        run += 1
        if (Z_ABS >> i) & 1: out.append((run, True)); run = 0
    if run: out.append((run, False))
    return out
def final_exp_vm(progs, s):
    run_program(progs["FE_INV_A"], s)
    n = s[SLOT["NORM"]]; assert n[1] == 0
    s[SLOT["NINV"]] = (fp_inv(n[0]), 0)
    run_program(progs["FE_INV_B"], s)
    def expz():
        for run, mul in expz_schedule():
            for n in (16, 8, 4, 2, 1):
                while run >= n: run_program(progs["CYCSQR" if n == 1 else f"CYCSQR{n}"], s); run -= n
            if mul: run_program(progs["MULX"], s)
    expz(); run_program(progs["GLUE1"], s)
    expz(); run_program(progs["GLUE2"], s)
    expz(); run_program(progs["GLUE3"], s)
    expz(); run_program(progs["GLUE4"], s)
    expz(); run_program(progs["GLUE5"], s); run_program(progs["GLUE6"], s)

# ------------------------------------------------------------------ emitter
# Instruction = 12 x u32 per (step, pair):
#   w0        MUL / SQR: dst | a << 8 | b << 16          LIN: dst | (atoms in the re row) << 8 | (atoms in the im row) << 12
#   w4 .. w7  LIN: the re row, 8 atoms of 16 bits (slot | half << 8 | negate << 9 | (m - 1) << 10: adds (+-) m x that half, m = 1..4),
#             unused atoms = the ZERO slot
#   w8 .. w11 LIN: the im row
# dst = 0xff: no operation.  A lane reads w0 and the uint4 of its own row.  Step header = class | (longest row of the step) << 8.
INS_WORDS = 12
def lin_rows(terms):
    rows = ([], [])
    for s_, M in terms:
        for r in (0, 1):
            for half in (0, 1):
                c = M[2 * r + half]; m = abs(c)
                for part in ([m] if m <= 4 else [4, m - 4]):
                    if part: rows[r].append(s_ | (half << 8) | ((1 if c < 0 else 0) << 9) | ((part - 1) << 10))
    return rows
def enc_lin(dst, terms):
    rows = lin_rows(terms)
    assert len(rows[0]) <= MAXA and len(rows[1]) <= MAXA
    w = [dst | (len(rows[0]) << 8) | (len(rows[1]) << 12), 0, 0, 0]
    for r in (0, 1):
        at = rows[r] + [SLOT["ZERO"]] * (MAXA - len(rows[r]))
        w += [at[2 * i] | (at[2 * i + 1] << 16) for i in range(4)]
    return w
def enc_mul(dst, a, b): return [dst | (a << 8) | (b << 16)] + [0] * (INS_WORDS - 1)
NOP = [0xff] + [0] * (INS_WORDS - 1)          # dst 0xff = no operation
PROGRAM_ORDER = ["ML_INIT", "ML_DBL", "ML_DBL2", "ML_ADD", "FE_INV_A", "FE_INV_B", "CYCSQR", "CYCSQR2", "CYCSQR4", "CYCSQR8", "CYCSQR16", "MULX",
                 "GLUE1", "GLUE2", "GLUE3", "GLUE4", "GLUE5", "GLUE6", "ML1_INIT", "ML1_DBL", "ML1_ADD", "A_ONE", "AMULF", "FMULA",
                 "G2_INIT", "G2DBL", "G2DBL2", "G2DBL4", "G2DBL8", "G2DBL16", "G2_ADD", "HC_MID", "HC_FIN", "G2_NORM", "G2_AFF"]

def emit(progs, path):
    order = PROGRAM_ORDER
    assert set(order) == set(progs)
    L = []
    A = L.append
    A("// GENERATED by tools/vmgen.py -- do not edit.  Step programs of the warp-cooperative pairing (see tools/vmgen.py for the format).")
    A("#ifndef HBLS_VM_PROGRAMS_CUH\n#define HBLS_VM_PROGRAMS_CUH\n#include <stdint.h>")
    A(f"#define VM_NSLOTS {NSLOTS}\n#define VM_OP_MUL {OP_MUL}\n#define VM_OP_SQR {OP_SQR}\n#define VM_OP_LIN {OP_LIN}\n#define VM_INS_WORDS {INS_WORDS}")
    for r in REGS: A(f"#define VM_R_{r} {SLOT[r]}")
    hdr = []; ins = []; table = []
    for pi, name in enumerate(order):
        p = progs[name]; first = len(hdr)
        for cls, instrs in p.steps:
            longest = max([max(len(r) for r in lin_rows(i[1])) for i in instrs]) if cls == OP_LIN else 0
            hdr.append(cls | (longest << 8))
            row = []
            for i in instrs: row += enc_lin(i[0], i[1]) if cls == OP_LIN else enc_mul(*i)
            row += NOP * (NPAIR - len(instrs))
            ins += row
        table.append((name, first, len(p.steps)))
        A(f"#define VM_P_{name} {pi}")
    A(f"#define VM_NPROG {len(order)}")
    A("static __device__ const uint16_t VM_PROG_FIRST[VM_NPROG] = {" + ", ".join(str(t[1]) for t in table) + "};")
    A("static __device__ const uint16_t VM_PROG_STEPS[VM_NPROG] = {" + ", ".join(str(t[2]) for t in table) + "};")
    A(f"// steps per program: " + ", ".join(f"{t[0]}={t[2]}" for t in table))
    A(f"static __device__ const uint16_t VM_STEP_HDR[{len(hdr)}] = {{" + ", ".join(str(h) for h in hdr) + "};")
    A(f"static __device__ const uint4 VM_INS[{len(ins) // 4}] = {{")
    for i in range(0, len(ins), 12):
        A("    " + " ".join("{0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu}," % tuple(ins[j:j + 4]) for j in range(i, i + 12, 4)))
    A("};\n#endif")
    open(path, "w").write("\n".join(L) + "\n")
    return table, len(hdr)

def stats(progs):
    out = {}
    for n, p in progs.items():
        c = {OP_MUL: 0, OP_SQR: 0, OP_LIN: 0}; fill = {OP_MUL: 0, OP_SQR: 0, OP_LIN: 0}
        for cls, ins in p.steps: c[cls] += 1; fill[cls] += len(ins)
        out[n] = {"steps": len(p.steps), "mul_steps": c[OP_MUL], "sqr_steps": c[OP_SQR], "lin_steps": c[OP_LIN],
                  "mul_ops": fill[OP_MUL], "sqr_ops": fill[OP_SQR], "lin_ops": fill[OP_LIN]}
    return out

if __name__ == "__main__":
    progs = compile_all()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    table, nsteps = emit(progs, os.path.join(root, "harmony_b200", "csrc", "vm_programs.cuh"))
    for n, s in stats(progs).items(): print(f"{n:10s} {s}")
    print("total steps stored:", nsteps)
