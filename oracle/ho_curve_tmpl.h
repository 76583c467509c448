/* oracle/ho_curve_tmpl.h -- Jacobian short-Weierstrass (a = 0) arithmetic, included twice by hbls_oracle.c
 * (FT = fp for G1, FT = fp2 for G2).  Test infrastructure only.  The all-zero struct is the identity,
 * exactly like the zero-value Go structs of the reference (crypto/bls/mask.go:59,88). */
typedef struct { FT x, y, z; } PT;

static inline int C_(is_inf)(const PT *p) { return F_(is_zero)(&p->z); }
static inline void C_(set_inf)(PT *p) { memset(p, 0, sizeof *p); }
static inline void C_(neg)(PT *r, const PT *p) { r->x = p->x; F_(neg)(&r->y, &p->y); r->z = p->z; }

static void C_(dbl)(PT *r, const PT *p) {
    if (C_(is_inf)(p)) { C_(set_inf)(r); return; }
    FT A, B, C, D, E, F, t;
    F_(sqr)(&A, &p->x); F_(sqr)(&B, &p->y); F_(sqr)(&C, &B);
    F_(add)(&D, &p->x, &B); F_(sqr)(&D, &D); F_(sub)(&D, &D, &A); F_(sub)(&D, &D, &C); F_(dbl)(&D, &D);
    F_(dbl)(&E, &A); F_(add)(&E, &E, &A);
    F_(sqr)(&F, &E);
    F_(mul)(&t, &p->y, &p->z); F_(dbl)(&r->z, &t);
    F_(dbl)(&t, &D); F_(sub)(&r->x, &F, &t);
    F_(sub)(&t, &D, &r->x); F_(mul)(&t, &t, &E);
    F_(dbl)(&C, &C); F_(dbl)(&C, &C); F_(dbl)(&C, &C);
    F_(sub)(&r->y, &t, &C);
}

static void C_(add)(PT *r, const PT *p, const PT *q) {
    if (C_(is_inf)(p)) { *r = *q; return; }
    if (C_(is_inf)(q)) { *r = *p; return; }
    FT Z1Z1, Z2Z2, U1, U2, S1, S2, H, R, HH, HHH, V, t;
    F_(sqr)(&Z1Z1, &p->z); F_(sqr)(&Z2Z2, &q->z);
    F_(mul)(&U1, &p->x, &Z2Z2); F_(mul)(&U2, &q->x, &Z1Z1);
    F_(mul)(&S1, &p->y, &q->z); F_(mul)(&S1, &S1, &Z2Z2);
    F_(mul)(&S2, &q->y, &p->z); F_(mul)(&S2, &S2, &Z1Z1);
    F_(sub)(&H, &U2, &U1); F_(sub)(&R, &S2, &S1);
    if (F_(is_zero)(&H)) {
        if (F_(is_zero)(&R)) { C_(dbl)(r, p); return; }
        C_(set_inf)(r); return;
    }
    F_(sqr)(&HH, &H); F_(mul)(&HHH, &H, &HH); F_(mul)(&V, &U1, &HH);
    F_(mul)(&t, &p->z, &q->z); F_(mul)(&r->z, &t, &H);
    F_(sqr)(&t, &R); F_(sub)(&t, &t, &HHH); F_(sub)(&t, &t, &V); F_(sub)(&r->x, &t, &V);
    F_(sub)(&t, &V, &r->x); F_(mul)(&t, &t, &R); F_(mul)(&S1, &S1, &HHH); F_(sub)(&r->y, &t, &S1);
}

static int C_(eq)(const PT *p, const PT *q) {
    if (C_(is_inf)(p) || C_(is_inf)(q)) return C_(is_inf)(p) && C_(is_inf)(q);
    FT Z1Z1, Z2Z2, a, b;
    F_(sqr)(&Z1Z1, &p->z); F_(sqr)(&Z2Z2, &q->z);
    F_(mul)(&a, &p->x, &Z2Z2); F_(mul)(&b, &q->x, &Z1Z1);
    if (!F_(eq)(&a, &b)) return 0;
    F_(mul)(&a, &p->y, &q->z); F_(mul)(&a, &a, &Z2Z2);
    F_(mul)(&b, &q->y, &p->z); F_(mul)(&b, &b, &Z1Z1);
    return F_(eq)(&a, &b);
}

/* to affine with z = 1 (identity stays all-zero) */
static void C_(normalize)(PT *r, const PT *p) {
    if (C_(is_inf)(p)) { C_(set_inf)(r); return; }
    FT zi, zi2;
    F_(inv)(&zi, &p->z); F_(sqr)(&zi2, &zi);
    F_(mul)(&r->x, &p->x, &zi2); F_(mul)(&zi2, &zi2, &zi); F_(mul)(&r->y, &p->y, &zi2);
    memset(&r->z, 0, sizeof r->z); memcpy(&r->z, K_ONE, 48);
}

/* r = [k]p, k = nl little-endian u64 limbs; 4-bit fixed window */
static void C_(mul)(PT *r, const PT *p, const u64 *k, int nl) {
    PT tbl[16]; C_(set_inf)(&tbl[0]); tbl[1] = *p;
    for (int i = 2; i < 16; i++) C_(add)(&tbl[i], &tbl[i - 1], p);
    PT acc; C_(set_inf)(&acc);
    for (int i = nl * 16 - 1; i >= 0; i--) {
        unsigned w = (unsigned)(k[i / 16] >> (4 * (i % 16))) & 15;
        C_(dbl)(&acc, &acc); C_(dbl)(&acc, &acc); C_(dbl)(&acc, &acc); C_(dbl)(&acc, &acc);
        if (w) C_(add)(&acc, &acc, &tbl[w]);
    }
    *r = acc;
}
