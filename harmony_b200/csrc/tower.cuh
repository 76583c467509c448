// harmony_b200/csrc/tower.cuh -- Fp2 = Fp[i]/(i^2+1), Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v), xi = 1 + i.
// Replaces mcl's Fp2/Fp6/Fp12 tower used by Sign.VerifyHash (reference call sites: consensus/leader.go:173,287,
// internal/chain/engine.go:638).  Pairing values are never exposed by the reference API (only booleans), so the
// tower basis and formulas are free; results are checked as booleans against oracle/.
#pragma once
#include "fp.cuh"
#include "fp_wide.cuh"
#include "hbls_constants.cuh"

namespace hb {
// per-helper outlining switches (code size vs call overhead); -DHB_OL_<NAME>=0/1, default = HB_OUTLINE_FP2
#ifndef HB_OL_ADD
#define HB_OL_ADD HB_OUTLINE_FP2
#endif
#if HB_OL_ADD
#define HB_ATTR_ADD __device__ __noinline__
#else
#define HB_ATTR_ADD __device__ __forceinline__
#endif
#ifndef HB_OL_SUB
#define HB_OL_SUB HB_OUTLINE_FP2
#endif
#if HB_OL_SUB
#define HB_ATTR_SUB __device__ __noinline__
#else
#define HB_ATTR_SUB __device__ __forceinline__
#endif
#ifndef HB_OL_NEG
#define HB_OL_NEG HB_OUTLINE_FP2
#endif
#if HB_OL_NEG
#define HB_ATTR_NEG __device__ __noinline__
#else
#define HB_ATTR_NEG __device__ __forceinline__
#endif
#ifndef HB_OL_CONJ
#define HB_OL_CONJ HB_OUTLINE_FP2
#endif
#if HB_OL_CONJ
#define HB_ATTR_CONJ __device__ __noinline__
#else
#define HB_ATTR_CONJ __device__ __forceinline__
#endif
#ifndef HB_OL_DBL
#define HB_OL_DBL HB_OUTLINE_FP2
#endif
#if HB_OL_DBL
#define HB_ATTR_DBL __device__ __noinline__
#else
#define HB_ATTR_DBL __device__ __forceinline__
#endif
#ifndef HB_OL_MULFP
#define HB_OL_MULFP HB_OUTLINE_FP2
#endif
#if HB_OL_MULFP
#define HB_ATTR_MULFP __device__ __noinline__
#else
#define HB_ATTR_MULFP __device__ __forceinline__
#endif
#ifndef HB_OL_MULXI
#define HB_OL_MULXI HB_OUTLINE_FP2
#endif
#if HB_OL_MULXI
#define HB_ATTR_MULXI __device__ __noinline__
#else
#define HB_ATTR_MULXI __device__ __forceinline__
#endif

struct fp2 { fp a, b; };
struct fp6 { fp2 c0, c1, c2; };
struct fp12 { fp6 c0, c1; };

HB_DEV void fp_const(fp& r, const uint32_t* k) { fp_set(r, k); }
HB_DEV void fp_one(fp& r) { fp_set(r, K_ONE); }

// ------------------------------------------------------------------ Fp exponentiation (fixed public exponents)
// r = a^e, e = 12 x u32 plain integer in constant memory; 4-bit fixed window (uniform control flow across the warp)
HB_NOINLINE void fp_pow(fp& r, const fp& a, const uint32_t* e) {
    fp tbl[16];
    fp_one(tbl[0]); tbl[1] = a;
    for (int i = 2; i < 16; i++) fp_mul(tbl[i], tbl[i - 1], a);
    fp acc; fp_one(acc);
    bool started = false;
    for (int i = 95; i >= 0; i--) {
        uint32_t w = (e[i >> 3] >> (4 * (i & 7))) & 15u;
        if (started) { fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc); }
        if (w) { if (started) fp_mul(acc, acc, tbl[w]); else { acc = tbl[w]; started = true; } }
    }
    r = acc;
}
HB_DEV void fp_inv(fp& r, const fp& a) { fp_pow(r, a, K_P_MINUS_2); }
// mcl Fp::squareRoot (p = 3 mod 4): candidate a^((p+1)/4), accepted iff it squares back
HB_DEV bool fp_sqrt(fp& r, const fp& a) {
    fp y, y2; fp_pow(y, a, K_P_PLUS_1_DIV_4); fp_sqr(y2, y);
    if (!fp_eq(y2, a)) return false;
    r = y; return true;
}
HB_DEV int fp_legendre(const fp& a) {
    if (fp_is_zero(a)) return 0;
    fp t, one; fp_pow(t, a, K_P_MINUS_1_DIV_2); fp_one(one);
    return fp_eq(t, one) ? 1 : -1;
}
// Montgomery <-> canonical integer limbs
HB_DEV void fp_from_int(fp& r, const fp& v) { fp r2; fp_set(r2, K_R2); fp_mul(r, v, r2); }
HB_DEV void fp_to_int(fp& v, const fp& a) { fp one; fp_zero(one); one.l[0] = 1; fp_mul(v, a, one); }
HB_DEV bool fp_int_geq_p(const fp& v) {
    uint32_t s, borrow;
    sub_cc(s, v.l[0], HB_P0);
#pragma unroll
    for (int j = 1; j < 12; j++) subc_cc(s, v.l[j], p_limb(j));
    subc(borrow, 0, 0);
    return borrow == 0;
}

// ------------------------------------------------------------------ Fp2
HB_DEV void fp2_zero(fp2& r) { fp_zero(r.a); fp_zero(r.b); }
HB_DEV void fp2_one(fp2& r) { fp_one(r.a); fp_zero(r.b); }
HB_DEV bool fp2_is_zero(const fp2& x) { return fp_is_zero(x.a) && fp_is_zero(x.b); }
HB_DEV bool fp2_eq(const fp2& x, const fp2& y) { return fp_eq(x.a, y.a) && fp_eq(x.b, y.b); }
HB_ATTR_ADD void fp2_add(fp2& r, const fp2& x, const fp2& y) { fp_add(r.a, x.a, y.a); fp_add(r.b, x.b, y.b); }
HB_ATTR_SUB void fp2_sub(fp2& r, const fp2& x, const fp2& y) { fp_sub(r.a, x.a, y.a); fp_sub(r.b, x.b, y.b); }
HB_ATTR_NEG void fp2_neg(fp2& r, const fp2& x) { fp_neg(r.a, x.a); fp_neg(r.b, x.b); }
HB_ATTR_CONJ void fp2_conj(fp2& r, const fp2& x) { r.a = x.a; fp_neg(r.b, x.b); }
HB_ATTR_DBL void fp2_dbl(fp2& r, const fp2& x) { fp_dbl(r.a, x.a); fp_dbl(r.b, x.b); }
HB_DEV void fp2_const(fp2& r, const uint32_t k[2][12]) { fp_set(r.a, k[0]); fp_set(r.b, k[1]); }
HB_DEV void fp2_cmov(fp2& r, const fp2& x, bool c) { fp_cmov(r.a, x.a, c); fp_cmov(r.b, x.b, c); }

#ifndef HB_FUSED_FP2
#define HB_FUSED_FP2 0      // 0: three out-of-line CIOS products (fastest measured); 1: lazy-reduction fused; 2: fused CIOS (see DESIGN.md results log)
#endif
#if HB_FUSED_FP2
// register-resident Karatsuba with lazy reduction (fp_wide.cuh): 744 IMAD.WIDE, operands touched once
HB_NOINLINE void fp2_mul(fp2& r, const fp2& x, const fp2& y) {
    uint32_t xa[12], xb[12], ya[12], yb[12], ra[12], rb[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { xa[j] = x.a.l[j]; xb[j] = x.b.l[j]; ya[j] = y.a.l[j]; yb[j] = y.b.l[j]; }
#if HB_FUSED_FP2 == 2
    fp2_mul_regs_cios(ra, rb, xa, xb, ya, yb);
#else
    fp2_mul_regs(ra, rb, xa, xb, ya, yb);
#endif
#pragma unroll
    for (int j = 0; j < 12; j++) { r.a.l[j] = ra[j]; r.b.l[j] = rb[j]; }
}
HB_NOINLINE void fp2_sqr(fp2& r, const fp2& x) {
    uint32_t xa[12], xb[12], ra[12], rb[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { xa[j] = x.a.l[j]; xb[j] = x.b.l[j]; }
#if HB_FUSED_FP2 == 2
    fp2_sqr_regs_cios(ra, rb, xa, xb);
#else
    fp2_sqr_regs(ra, rb, xa, xb);
#endif
#pragma unroll
    for (int j = 0; j < 12; j++) { r.a.l[j] = ra[j]; r.b.l[j] = rb[j]; }
}
#else
HB_NOINLINE void fp2_mul(fp2& r, const fp2& x, const fp2& y) {
    fp t0, t1, t2, s0, s1;
    fp_mul(t0, x.a, y.a); fp_mul(t1, x.b, y.b);
    fp_add(s0, x.a, x.b); fp_add(s1, y.a, y.b); fp_mul(t2, s0, s1);
    fp_sub(r.a, t0, t1); fp_sub(t2, t2, t0); fp_sub(r.b, t2, t1);
}
HB_NOINLINE void fp2_sqr(fp2& r, const fp2& x) {
    fp s, d, m;
    fp_add(s, x.a, x.b); fp_sub(d, x.a, x.b); fp_mul(m, x.a, x.b);
    fp_mul(r.a, s, d); fp_dbl(r.b, m);
}
#endif
HB_ATTR_MULFP void fp2_mul_fp(fp2& r, const fp2& x, const fp& k) { fp_mul(r.a, x.a, k); fp_mul(r.b, x.b, k); }
HB_ATTR_MULXI void fp2_mul_xi(fp2& r, const fp2& x) { fp t; fp_sub(t, x.a, x.b); fp_add(r.b, x.a, x.b); r.a = t; }
HB_DEV void fp2_norm(fp& r, const fp2& x) { fp t; fp_sqr(r, x.a); fp_sqr(t, x.b); fp_add(r, r, t); }
HB_NOINLINE void fp2_inv(fp2& r, const fp2& x) {
    fp n; fp2_norm(n, x); fp_inv(n, n);
    fp_mul(r.a, x.a, n); fp_mul(r.b, x.b, n); fp_neg(r.b, r.b);
}
// mcl Fp2::squareRoot: fixes WHICH root is produced -- bit-exactness of hash-to-G2 and of
// signature decompression depends on it
HB_NOINLINE bool fp2_sqrt(fp2& r, const fp2& x) {
    fp t1, t2, inv2;
    if (fp_is_zero(x.b)) {
        if (fp_sqrt(t1, x.a)) { r.a = t1; fp_zero(r.b); }
        else { fp_neg(t2, x.a); if (!fp_sqrt(t1, t2)) return false; fp_zero(r.a); r.b = t1; }
        return true;
    }
    fp2_norm(t1, x);
    if (!fp_sqrt(t1, t1)) return false;
    fp_set(inv2, K_INV2);
    fp_add(t2, x.a, t1); fp_mul(t2, t2, inv2);
    if (!fp_sqrt(t2, t2)) {
        fp_sub(t2, x.a, t1); fp_mul(t2, t2, inv2);
        if (!fp_sqrt(t2, t2)) return false;
    }
    fp c = t2;
    fp_dbl(t2, t2); fp_inv(t2, t2);
    fp_mul(r.b, x.b, t2); r.a = c;
    return true;
}

// ------------------------------------------------------------------ Fp6
HB_DEV void fp6_add(fp6& r, const fp6& x, const fp6& y) { fp2_add(r.c0, x.c0, y.c0); fp2_add(r.c1, x.c1, y.c1); fp2_add(r.c2, x.c2, y.c2); }
HB_DEV void fp6_sub(fp6& r, const fp6& x, const fp6& y) { fp2_sub(r.c0, x.c0, y.c0); fp2_sub(r.c1, x.c1, y.c1); fp2_sub(r.c2, x.c2, y.c2); }
HB_DEV void fp6_neg(fp6& r, const fp6& x) { fp2_neg(r.c0, x.c0); fp2_neg(r.c1, x.c1); fp2_neg(r.c2, x.c2); }
HB_DEV void fp6_mul_v(fp6& r, const fp6& x) { fp2 t; fp2_mul_xi(t, x.c2); r.c2 = x.c1; r.c1 = x.c0; r.c0 = t; }
HB_NOINLINE void fp6_mul(fp6& r, const fp6& x, const fp6& y) {
    fp2 v0, v1, v2, t0, t1, t2, s;
    fp2_mul(v0, x.c0, y.c0); fp2_mul(v1, x.c1, y.c1); fp2_mul(v2, x.c2, y.c2);
    fp2_add(t0, x.c1, x.c2); fp2_add(s, y.c1, y.c2); fp2_mul(t0, t0, s);
    fp2_sub(t0, t0, v1); fp2_sub(t0, t0, v2); fp2_mul_xi(t0, t0); fp2_add(t0, t0, v0);
    fp2_add(t1, x.c0, x.c1); fp2_add(s, y.c0, y.c1); fp2_mul(t1, t1, s);
    fp2_sub(t1, t1, v0); fp2_sub(t1, t1, v1); fp2_mul_xi(s, v2); fp2_add(t1, t1, s);
    fp2_add(t2, x.c0, x.c2); fp2_add(s, y.c0, y.c2); fp2_mul(t2, t2, s);
    fp2_sub(t2, t2, v0); fp2_sub(t2, t2, v2); fp2_add(t2, t2, v1);
    r.c0 = t0; r.c1 = t1; r.c2 = t2;
}
// x * (b0 + b1 v)
HB_NOINLINE void fp6_mul_by_01(fp6& r, const fp6& x, const fp2& b0, const fp2& b1) {
    fp2 v0, v1, t0, t1, t2, s;
    fp2_mul(v0, x.c0, b0); fp2_mul(v1, x.c1, b1);
    fp2_mul(t0, x.c2, b1); fp2_mul_xi(t0, t0); fp2_add(t0, t0, v0);
    fp2_add(t1, x.c0, x.c1); fp2_add(s, b0, b1); fp2_mul(t1, t1, s); fp2_sub(t1, t1, v0); fp2_sub(t1, t1, v1);
    fp2_mul(t2, x.c2, b0); fp2_add(t2, t2, v1);
    r.c0 = t0; r.c1 = t1; r.c2 = t2;
}
// x * (b1 v)
HB_NOINLINE void fp6_mul_by_1(fp6& r, const fp6& x, const fp2& b1) {
    fp2 t0, t1, t2;
    fp2_mul(t0, x.c2, b1); fp2_mul_xi(t0, t0);
    fp2_mul(t1, x.c0, b1); fp2_mul(t2, x.c1, b1);
    r.c0 = t0; r.c1 = t1; r.c2 = t2;
}
HB_NOINLINE void fp6_inv(fp6& r, const fp6& x) {
    fp2 t0, t1, t2, s, d;
    fp2_sqr(t0, x.c0); fp2_mul(s, x.c1, x.c2); fp2_mul_xi(s, s); fp2_sub(t0, t0, s);
    fp2_sqr(t1, x.c2); fp2_mul_xi(t1, t1); fp2_mul(s, x.c0, x.c1); fp2_sub(t1, t1, s);
    fp2_sqr(t2, x.c1); fp2_mul(s, x.c0, x.c2); fp2_sub(t2, t2, s);
    fp2_mul(d, x.c2, t1); fp2_mul(s, x.c1, t2); fp2_add(d, d, s); fp2_mul_xi(d, d);
    fp2_mul(s, x.c0, t0); fp2_add(d, d, s);
    fp2_inv(d, d);
    fp2_mul(r.c0, t0, d); fp2_mul(r.c1, t1, d); fp2_mul(r.c2, t2, d);
}

// ------------------------------------------------------------------ Fp12
HB_DEV void fp12_one(fp12& r) {
    fp2_one(r.c0.c0); fp2_zero(r.c0.c1); fp2_zero(r.c0.c2); fp2_zero(r.c1.c0); fp2_zero(r.c1.c1); fp2_zero(r.c1.c2);
}
HB_DEV bool fp12_is_one(const fp12& x) {
    fp one; fp_one(one);
    return fp_eq(x.c0.c0.a, one) && fp_is_zero(x.c0.c0.b) && fp2_is_zero(x.c0.c1) && fp2_is_zero(x.c0.c2) &&
           fp2_is_zero(x.c1.c0) && fp2_is_zero(x.c1.c1) && fp2_is_zero(x.c1.c2);
}
HB_NOINLINE void fp12_mul(fp12& r, const fp12& x, const fp12& y) {
    fp6 v0, v1, s, t;
    fp6_mul(v0, x.c0, y.c0); fp6_mul(v1, x.c1, y.c1);
    fp6_add(s, x.c0, x.c1); fp6_add(t, y.c0, y.c1); fp6_mul(s, s, t);
    fp6_sub(s, s, v0); fp6_sub(s, s, v1);
    fp6_mul_v(t, v1); fp6_add(r.c0, v0, t); r.c1 = s;
}
HB_NOINLINE void fp12_sqr(fp12& r, const fp12& x) {
    fp6 ab, s, t;
    fp6_mul(ab, x.c0, x.c1);
    fp6_add(s, x.c0, x.c1); fp6_mul_v(t, x.c1); fp6_add(t, t, x.c0); fp6_mul(s, s, t);
    fp6_sub(s, s, ab); fp6_mul_v(t, ab); fp6_sub(r.c0, s, t);
    fp6_add(r.c1, ab, ab);
}
HB_DEV void fp12_conj(fp12& r, const fp12& x) { r.c0 = x.c0; fp6_neg(r.c1, x.c1); }
HB_NOINLINE void fp12_inv(fp12& r, const fp12& x) {
    fp6 t0, t1;
    fp6_mul(t0, x.c0, x.c0); fp6_mul(t1, x.c1, x.c1); fp6_mul_v(t1, t1); fp6_sub(t0, t0, t1);
    fp6_inv(t0, t0);
    fp6_mul(r.c0, x.c0, t0); fp6_mul(t1, x.c1, t0); fp6_neg(r.c1, t1);
}
// sparse multiply by a Miller line (o0 + o1 v) + (o4 v) w : non-zero coefficients at w^0, w^2, w^3
HB_NOINLINE void fp12_mul_by_014(fp12& r, const fp12& x, const fp2& o0, const fp2& o1, const fp2& o4) {
    fp6 aa, bb, s; fp2 o14;
    fp6_mul_by_01(aa, x.c0, o0, o1);
    fp6_mul_by_1(bb, x.c1, o4);
    fp2_add(o14, o1, o4);
    fp6_add(s, x.c0, x.c1); fp6_mul_by_01(s, s, o0, o14);
    fp6_sub(s, s, aa); fp6_sub(s, s, bb);
    fp6_mul_v(bb, bb); fp6_add(r.c0, aa, bb); r.c1 = s;
}
// coefficient of w^k (k = 2i + j) <-> tower slot
HB_DEV fp2& fp12_slot(fp12& x, int k) {
    fp6& h = (k & 1) ? x.c1 : x.c0;
    int i = k >> 1;
    return i == 0 ? h.c0 : (i == 1 ? h.c1 : h.c2);
}
HB_NOINLINE void fp12_frob(fp12& r, const fp12& x) {
    fp12 t = x;
    for (int k = 0; k < 6; k++) {
        fp2& s = fp12_slot(t, k); fp2 g;
        fp2_conj(s, s); fp2_const(g, K_FROB1[k]); fp2_mul(s, s, g);
    }
    r = t;
}
HB_NOINLINE void fp12_frob2(fp12& r, const fp12& x) {
    fp12 t = x;
    for (int k = 0; k < 6; k++) { fp2& s = fp12_slot(t, k); fp g; fp_set(g, K_FROB2[k]); fp2_mul_fp(s, s, g); }
    r = t;
}
// Granger-Scott squaring in the cyclotomic subgroup (valid after the easy part of the final exponentiation)
HB_DEV void fp4_sqr(fp2& c0, fp2& c1, const fp2& a, const fp2& b) {
    fp2 t0, t1, t2;
    fp2_sqr(t0, a); fp2_sqr(t1, b);
    fp2_mul_xi(t2, t1); fp2_add(c0, t2, t0);
    fp2_add(t2, a, b); fp2_sqr(t2, t2); fp2_sub(t2, t2, t0); fp2_sub(c1, t2, t1);
}
HB_NOINLINE void fp12_cyc_sqr(fp12& r, const fp12& x) {
    fp2 z0 = x.c0.c0, z4 = x.c0.c1, z3 = x.c0.c2, z2 = x.c1.c0, z1 = x.c1.c1, z5 = x.c1.c2;
    fp2 t0, t1, t2, t3;
    fp4_sqr(t0, t1, z0, z1);
    fp2_sub(z0, t0, z0); fp2_dbl(z0, z0); fp2_add(z0, z0, t0);
    fp2_add(z1, t1, z1); fp2_dbl(z1, z1); fp2_add(z1, z1, t1);
    fp4_sqr(t0, t1, z2, z3);
    fp4_sqr(t2, t3, z4, z5);
    fp2_sub(z4, t0, z4); fp2_dbl(z4, z4); fp2_add(z4, z4, t0);
    fp2_add(z5, t1, z5); fp2_dbl(z5, z5); fp2_add(z5, z5, t1);
    fp2_mul_xi(t0, t3);
    fp2_add(z2, t0, z2); fp2_dbl(z2, z2); fp2_add(z2, z2, t0);
    fp2_sub(z3, t2, z3); fp2_dbl(z3, z3); fp2_add(z3, z3, t2);
    r.c0.c0 = z0; r.c0.c1 = z4; r.c0.c2 = z3; r.c1.c0 = z2; r.c1.c1 = z1; r.c1.c2 = z5;
}
// r = x^z, z = -0xd201000000010000, x in the cyclotomic subgroup
HB_NOINLINE void fp12_cyc_exp_z(fp12& r, const fp12& x) {
    fp12 acc = x;
    for (int i = 62; i >= 0; i--) {
        fp12_cyc_sqr(acc, acc);
        if ((K_Z_ABS >> i) & 1) fp12_mul(acc, acc, x);
    }
    fp12_conj(r, acc);
}

}  // namespace hb
