// harmony_b200/csrc/pairing.cuh -- optimal-ate Miller loop and final exponentiation on BLS12-381.
// Replaces the pairing inside Sign.VerifyHash (reference consensus/leader.go:173,287; internal/chain/engine.go:638;
// consensus/validator.go:228).  The reference only exposes e(B, sig) == e(pk, H(m)) as a boolean, so the loop
// shape (projective twist coordinates, lines scaled by subfield elements, exponent 3*(p^12-1)/r) is free.
#pragma once
#include "curve.cuh"

namespace hb {

template <class E> struct g2proj_t { E x, y, z; };    // homogeneous projective point on the twist
typedef g2proj_t<fp2> g2proj;

// doubling step; line = l0 + (l2 * xP) w^2 + (l3 * yP) w^3 up to an Fp2 factor
template <class E> HB_NOINLINE void ml_dbl(g2proj_t<E>& t, E& l0, E& l2, E& l3) {
    hb_lockstep2<E>();
    // 4 products + 6 squarings of Fp2 (24 Fp products): the halvings and the product by the twist constant 3 b' = 12 (1 + i) are
    // shifts / additions (fp2_half, fp2_mul_twist3b), not multiplications by constants (7 Fp products fewer per step than round 1)
    E A, B, C, Ee, F, H, s;
    fp2_mul(A, t.x, t.y); fp2_half(A, A);
    fp2_sqr(B, t.y); fp2_sqr(C, t.z);
    fp2_mul_twist3b(Ee, C);
    fp2_dbl(F, Ee); fp2_add(F, F, Ee);
    fp2_add(H, t.y, t.z); fp2_sqr(H, H); fp2_sub(H, H, B); fp2_sub(H, H, C);     // 2YZ
    fp2_sub(l0, B, Ee);                                                           // Y^2 - 3b'Z^2
    fp2_sqr(s, t.x); fp2_dbl(l2, s); fp2_add(l2, l2, s); fp2_neg(l2, l2);         // -3X^2
    l3 = H;
    E x3, y3, e2;
    fp2_sub(x3, B, F); fp2_mul(x3, x3, A);
    fp2_add(y3, B, F); fp2_half(y3, y3); fp2_sqr(y3, y3);
    fp2_sqr(e2, Ee); fp2_dbl(s, e2); fp2_add(s, s, e2); fp2_sub(y3, y3, s);
    fp2_mul(t.z, B, H); t.x = x3; t.y = y3;
}
// addition step T += Q, Q affine
template <class E> HB_NOINLINE void ml_add(g2proj_t<E>& t, const E& qx, const E& qy, E& l0, E& l2, E& l3) {
    hb_lockstep2<E>();
    E th, mu, C, D, Ee, F, G, H, s;
    fp2_mul(th, qy, t.z); fp2_sub(th, t.y, th);
    fp2_mul(mu, qx, t.z); fp2_sub(mu, t.x, mu);
    fp2_mul(l0, th, qx); fp2_mul(s, mu, qy); fp2_sub(l0, l0, s);
    fp2_neg(l2, th); l3 = mu;
    fp2_sqr(C, th); fp2_sqr(D, mu); fp2_mul(Ee, mu, D); fp2_mul(F, t.z, C); fp2_mul(G, t.x, D);
    fp2_add(H, Ee, F); fp2_sub(H, H, G); fp2_sub(H, H, G);
    E x3, y3;
    fp2_mul(x3, mu, H);
    fp2_sub(y3, G, H); fp2_mul(y3, y3, th); fp2_mul(s, Ee, t.y); fp2_sub(y3, y3, s);
    fp2_mul(t.z, t.z, Ee); t.x = x3; t.y = y3;
}
// f = f_{|z|,Q}(P) (conjugation for z < 0 omitted: f == 1 after final exp  <=>  conj(f) == 1 after final exp,
// and a product of such values is conjugated as a whole).  Identity inputs give f = 1.
HB_NOINLINE void miller_loop(fp12& f, const g1a& p, const g2a& q) {
    fp12_one(f);
    if (aff_is_inf(p) || aff_is_inf(q)) return;
    g2proj T; T.x = q.x; T.y = q.y; fp2_one(T.z);
    fp2 l0, l2, l3;
    for (int i = 62; i >= 0; i--) {
        fp12_sqr(f, f);
        ml_dbl(T, l0, l2, l3);
        fp2_mul_fp(l2, l2, p.x); fp2_mul_fp(l3, l3, p.y);
        fp12_mul_by_014(f, f, l0, l2, l3);
        if ((K_Z_ABS >> i) & 1) {
            ml_add(T, q.x, q.y, l0, l2, l3);
            fp2_mul_fp(l2, l2, p.x); fp2_mul_fp(l3, l3, p.y);
            fp12_mul_by_014(f, f, l0, l2, l3);
        }
    }
}
// two pairs at once: f = f_{|z|,Q1}(P1) * f_{|z|,Q2}(P2) sharing the 63 Fp12 squarings (-18% vs two single loops).
// This is the shape of every verification: (B, sig) and (-pk, H(m)).  on1/on2 = pair is not an identity pair.
template <class E> HB_NOINLINE void miller_loop2(fp12_t<E>& f, const g1a& p1, const E& q1x, const E& q1y, const g1a& p2, const E& q2x, const E& q2y,
                                                 bool on1, bool on2) {
    fp12_one(f);
    g2proj_t<E> T1, T2;
    T1.x = q1x; T1.y = q1y; fp2_one(T1.z);
    T2.x = q2x; T2.y = q2y; fp2_one(T2.z);
    E l0, l2, l3;
    for (int i = 62; i >= 0; i--) {
        hb_lockstep<E>();
        fp12_sqr(f, f);
        if (on1) { ml_dbl(T1, l0, l2, l3); fp2_mul_fp(l2, l2, p1.x); fp2_mul_fp(l3, l3, p1.y); fp12_mul_by_014(f, f, l0, l2, l3); }
        if (on2) { ml_dbl(T2, l0, l2, l3); fp2_mul_fp(l2, l2, p2.x); fp2_mul_fp(l3, l3, p2.y); fp12_mul_by_014(f, f, l0, l2, l3); }
        if ((K_Z_ABS >> i) & 1) {
            if (on1) { ml_add(T1, q1x, q1y, l0, l2, l3); fp2_mul_fp(l2, l2, p1.x); fp2_mul_fp(l3, l3, p1.y); fp12_mul_by_014(f, f, l0, l2, l3); }
            if (on2) { ml_add(T2, q2x, q2y, l0, l2, l3); fp2_mul_fp(l2, l2, p2.x); fp2_mul_fp(l3, l3, p2.y); fp12_mul_by_014(f, f, l0, l2, l3); }
        }
    }
}
HB_DEV void miller_loop2(fp12& f, const g1a& p1, const g2a& q1, const g1a& p2, const g2a& q2) {
    miller_loop2<fp2>(f, p1, q1.x, q1.y, p2, q2.x, q2.y, !(aff_is_inf(p1) || aff_is_inf(q1)), !(aff_is_inf(p2) || aff_is_inf(q2)));
}
// NP pairs at once with ONE shared accumulator: f = prod_k f_{|z|,Q_k}(P_k).  Used by the random-linear-combination
// batch: rounds of a group contribute (-r_j apk_j, H(m_j)), the last pair is (B, sum_j r_j sigma_j).  P_k / Q_k stay in
// global memory (read again at the five addition steps); only the running points T_k live in the thread.
#ifndef HB_LOCKSTEP_PAIR
#define HB_LOCKSTEP_PAIR 1
#endif
template <class E, int NP> HB_NOINLINE void miller_loop_multi(fp12_t<E>& f, const g1a* const* ps, const E* qx, const E* qy) {
    fp12_one(f);
    g2proj_t<E> T[NP];
    for (int k = 0; k < NP; k++) { T[k].x = qx[k]; T[k].y = qy[k]; fp2_one(T[k].z); }
    E l0, l2, l3;
    for (int i = 62; i >= 0; i--) {
        hb_lockstep<E>();
        fp12_sqr(f, f);
#pragma unroll 1
        for (int k = 0; k < NP; k++) {
            // long pair lists: re-align the CTA's warps inside the iteration too (instruction cache); measured on B200,
            // 9 pairs: none 110 ms, every two pairs 97 ms
#if HB_LOCKSTEP_PAIR == 1
            if (NP > 4 && (k & 1) == 0 && k) hb_lockstep<E>();
#elif HB_LOCKSTEP_PAIR >= 2
            if (k) hb_lockstep<E>();
#endif
            const fp px = ps[k]->x, py = ps[k]->y;          // staged into thread-local storage: field routines take local operands
            ml_dbl(T[k], l0, l2, l3); fp2_mul_fp(l2, l2, px); fp2_mul_fp(l3, l3, py);
#if HB_LOCKSTEP_PAIR >= 3
            hb_lockstep<E>();
#endif
            fp12_mul_by_014(f, f, l0, l2, l3);
        }
        if ((K_Z_ABS >> i) & 1) {
#pragma unroll 1
            for (int k = 0; k < NP; k++) {
                const fp px = ps[k]->x, py = ps[k]->y;
                ml_add(T[k], qx[k], qy[k], l0, l2, l3); fp2_mul_fp(l2, l2, px); fp2_mul_fp(l3, l3, py); fp12_mul_by_014(f, f, l0, l2, l3);
            }
        }
    }
}
// r = f^(3 (p^12 - 1) / r): easy part, then (z-1)^2 (z+p) (z^2+p^2-1) + 3
template <class E> HB_NOINLINE void final_exp(fp12_t<E>& r, const fp12_t<E>& f) {
    fp12_t<E> t0, t1, t2, m;
    fp12_conj(t0, f); fp12_inv(t1, f); fp12_mul(m, t0, t1);
    fp12_frob2(t0, m); fp12_mul(m, t0, m);
    fp12_cyc_exp_z(t0, m); fp12_conj(t1, m); fp12_mul(t0, t0, t1);               // a = m^(z-1)
    fp12_cyc_exp_z(t1, t0); fp12_conj(t2, t0); fp12_mul(t1, t1, t2);             // b = a^(z-1)
    fp12_cyc_exp_z(t0, t1); fp12_frob(t2, t1); fp12_mul(t0, t0, t2);             // c = b^(z+p)
    fp12_cyc_exp_z(t1, t0); fp12_cyc_exp_z(t1, t1); fp12_frob2(t2, t0); fp12_mul(t1, t1, t2);
    fp12_conj(t2, t0); fp12_mul(t1, t1, t2);                                     // d = c^(z^2+p^2-1)
    fp12_cyc_sqr(t2, m); fp12_mul(t2, t2, m); fp12_mul(r, t1, t2);               // d * m^3
}

}  // namespace hb
