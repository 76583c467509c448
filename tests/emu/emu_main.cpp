// tests/emu/emu_main.cpp -- TEST HARNESS ONLY: compiles the device headers (harmony_b200/csrc/*.cuh) for the host with
// software carry flags (HB_HOST_EMU) so tests can diff the device LOGIC against the oracle without a GPU.
// Never linked into libhbls.so; the product has no CPU path.
#define HB_HOST_EMU 1
#include <cstring>
#include "../../harmony_b200/csrc/pairing.cuh"
using namespace hb;
static void ld(fp& r, const uint8_t* b) { fp v; load_words(v.l, b, 12); fp_from_int(r, v); }
static void st(uint8_t* b, const fp& a) { fp v; fp_to_int(v, a); store_words(b, v.l, 12); }
extern "C" {
void emu_fp_mul(const uint8_t* a, const uint8_t* b, uint8_t* o) { fp x, y; ld(x, a); ld(y, b); fp_mul(x, x, y); st(o, x); }
void emu_fp_sqr(const uint8_t* a, uint8_t* o) { fp x; ld(x, a); fp_sqr(x, x); st(o, x); }
void emu_fp_addsubneg(const uint8_t* a, const uint8_t* b, uint8_t* o) { fp x, y, r; ld(x, a); ld(y, b); fp_add(r, x, y); st(o, r); fp_sub(r, x, y); st(o + 48, r); fp_neg(r, x); st(o + 96, r); }
void emu_fp2_mul(const uint8_t* a, const uint8_t* b, uint8_t* o) { fp2 x, y; ld(x.a, a); ld(x.b, a + 48); ld(y.a, b); ld(y.b, b + 48); fp2_mul(x, x, y); st(o, x.a); st(o + 48, x.b); }
void emu_fp2_sqr(const uint8_t* a, uint8_t* o) { fp2 x; ld(x.a, a); ld(x.b, a + 48); fp2_sqr(x, x); st(o, x.a); st(o + 48, x.b); }
int emu_map_to_g2(const uint8_t* msg, uint32_t len, uint8_t* out96) { g2 h; if (!map_to_g2(h, msg, len)) return -1; g2_serialize(out96, h); return 0; }
int emu_g2_check(const uint8_t* in96) { g2 p; return g2_deserialize(p, in96, true) ? 1 : 0; }
int emu_g1_check(const uint8_t* in48) { g1 p; return g1_deserialize(p, in48, true) ? 1 : 0; }
int emu_g1_mul_gen(const uint8_t* sk32, uint8_t* out48) { uint32_t k[8]; load_words(k, sk32, 8); g1 g, r; g1_generator(g); pt_mul(r, g, k, 8); g1_serialize(out48, r); return 0; }
// e(B, sig) * e(-pk, H(m)) == 1 on serialized inputs (the VerifyHash composition of the kernels)
int emu_verify(const uint8_t* sig96, const uint8_t* pk48, const uint8_t* msg, uint32_t len) {
    g2 s, h; g1 p;
    if (!g2_deserialize(s, sig96, true) || !g1_deserialize(p, pk48, true) || !map_to_g2(h, msg, len)) return 0;
    g2a sa, ha; g1a pa, ga; pt_to_aff(sa, s); pt_to_aff(ha, h); pt_to_aff(pa, p); fp_neg(pa.y, pa.y);
    fp_set(ga.x, K_G1_X); fp_set(ga.y, K_G1_Y);
    fp12 f1, f2, m; miller_loop(f1, ga, sa); miller_loop(f2, pa, ha); fp12_mul(m, f1, f2); final_exp(m, m);
    fp12 m2; miller_loop2(m2, ga, sa, pa, ha); final_exp(m2, m2);
    int a = fp12_is_one(m) ? 1 : 0, b = fp12_is_one(m2) ? 1 : 0;
    return a == b ? a : -7;            // split and fused pairing forms must agree
}
}
extern "C" void emu_mul_wide2(const uint8_t* a1, const uint8_t* b1, const uint8_t* a2, const uint8_t* b2, uint8_t* out96) {
    uint32_t A1[12], B1[12], A2[12], B2[12], T[24];
    load_words(A1, a1, 12); load_words(B1, b1, 12); load_words(A2, a2, 12); load_words(B2, b2, 12);
    mul_wide2(T, A1, B1, A2, B2); store_words(out96, T, 24);
}
extern "C" void emu_sqr_wide_redc(const uint8_t* a, uint8_t* out96) { uint32_t A[12], T[24]; load_words(A, a, 12); sqr_wide(T, A); store_words(out96, T, 24); }
// random-linear-combination group check (the math of k_rlc_scale / k_rlc_group_sum / k_rlc_pairing_split with the fp2 carrier):
// 7 (pk, sig, msg) rounds + coefficients r -> prod_j e(-r_j pk_j, H_j) * e(B, sum r_j sig_j) == 1 ?
extern "C" int emu_rlc_group(const uint8_t* pks48, const uint8_t* sigs96, const uint8_t* msgs, uint32_t len, const uint64_t* r) {
    const int G = 7;
    static g1a P[G + 1]; fp2 qx[G + 1], qy[G + 1]; const g1a* ps[G + 1];
    g2 acc; pt_set_inf(acc);
    for (int k = 0; k < G; k++) {
        g1 pk; g2 sg, h;
        if (!g1_deserialize(pk, pks48 + 48 * k, true) || !g2_deserialize(sg, sigs96 + 96 * k, true) || !map_to_g2(h, msgs + len * k, len)) return -1;
        g2a sa; pt_to_aff(sa, sg); g1 rp; g2 rs; rlc_scale_pair(rp, rs, pk, sa, r[k]); pt_add(acc, acc, rs);
        pt_to_aff(P[k], rp); fp_neg(P[k].y, P[k].y);
        g2a ha; pt_to_aff(ha, h); qx[k] = ha.x; qy[k] = ha.y; ps[k] = &P[k];
    }
    g2a sga; pt_to_aff(sga, acc); qx[G] = sga.x; qy[G] = sga.y;
    fp_set(P[G].x, K_G1_X); fp_set(P[G].y, K_G1_Y); ps[G] = &P[G];
    fp12 m; miller_loop_multi<fp2, G + 1>(m, ps, qx, qy); final_exp(m, m);
    return fp12_is_one(m) ? 1 : 0;
}
// executed Fp mul / sqr counts of the batched pairing stage for one group of G rounds (bench.py: "executed" roofline):
// out[0..1] = scaling of the 4 rounds (r * apk -> affine, r * sigma), out[2..3] = group sum + 5-pair Miller loop + final exp
template <int G> static int rlc_stage_counts(const uint8_t* pks48, const uint8_t* sigs96, const uint8_t* msgs, uint32_t len, uint64_t* out) {
    static g1a P[G + 1]; fp2 qx[G + 1], qy[G + 1]; const g1a* ps[G + 1];
    g1 pk[G]; g2a sa[G]; g2a ha[G];
    for (int k = 0; k < G; k++) {
        g2 sg, h;
        if (!g1_deserialize(pk[k], pks48 + 48 * k, true) || !g2_deserialize(sg, sigs96 + 96 * k, true) || !map_to_g2(h, msgs + len * k, len)) return -1;
        pt_to_aff(sa[k], sg); pt_to_aff(ha[k], h);
    }
    uint64_t m0 = hb_emu_cnt_mul, s0 = hb_emu_cnt_sqr;
    g2 S[G];
    for (int k = 0; k < G; k++) {
        const uint64_t r = 0x9e3779b97f4a7c15ull * (k + 1) | 1ull;
        g1 rp; rlc_scale_pair(rp, S[k], pk[k], sa[k], r); pt_to_aff(P[k], rp); fp_neg(P[k].y, P[k].y);
        qx[k] = ha[k].x; qy[k] = ha[k].y; ps[k] = &P[k];
    }
    out[0] = hb_emu_cnt_mul - m0; out[1] = hb_emu_cnt_sqr - s0; m0 = hb_emu_cnt_mul; s0 = hb_emu_cnt_sqr;
    g2 acc; pt_set_inf(acc);
    for (int k = 0; k < G; k++) pt_add(acc, acc, S[k]);
    g2a sga; pt_to_aff(sga, acc); qx[G] = sga.x; qy[G] = sga.y;
    fp_set(P[G].x, K_G1_X); fp_set(P[G].y, K_G1_Y); ps[G] = &P[G];
    fp12 m; miller_loop_multi<fp2, G + 1>(m, ps, qx, qy); final_exp(m, m);
    out[2] = hb_emu_cnt_mul - m0; out[3] = hb_emu_cnt_sqr - s0;
    // the share of the line kernel of the two-kernel form (k_rlc_lines_split: running points + lines evaluated at P, G + 1 pairs):
    // out[4..5]; the accumulator kernel (k_rlc_accum_split) executes the rest of out[2..3] minus the group sum
    m0 = hb_emu_cnt_mul; s0 = hb_emu_cnt_sqr;
    for (int k = 0; k <= G; k++) {
        g2proj T; T.x = qx[k]; T.y = qy[k]; fp2_one(T.z);
        fp2 l0, l2, l3;
        for (int i = 62; i >= 0; i--) {
            ml_dbl(T, l0, l2, l3); fp2_mul_fp(l2, l2, ps[k]->x); fp2_mul_fp(l3, l3, ps[k]->y);
            if ((K_Z_ABS >> i) & 1) { ml_add(T, qx[k], qy[k], l0, l2, l3); fp2_mul_fp(l2, l2, ps[k]->x); fp2_mul_fp(l3, l3, ps[k]->y); }
        }
    }
    out[4] = hb_emu_cnt_mul - m0; out[5] = hb_emu_cnt_sqr - s0;
    return fp12_is_one(m) ? 1 : 0;
}
extern "C" int emu_rlc_stage_counts(int G, const uint8_t* pks48, const uint8_t* sigs96, const uint8_t* msgs, uint32_t len, uint64_t* out) {
    return G == 8 ? rlc_stage_counts<8>(pks48, sigs96, msgs, len, out) : rlc_stage_counts<4>(pks48, sigs96, msgs, len, out);
}
// the two-base ladder against the plain ladder on the full scalar s = a + b z^2 mod r (8 LE words from the test)
extern "C" int emu_rlc_scale_check(const uint8_t* pk48, const uint8_t* sig96, uint64_t c, const uint32_t* s_words) {
    g1 pk; g2 sg;
    if (!g1_deserialize(pk, pk48, true) || !g2_deserialize(sg, sig96, true)) return -1;
    g2a sa; pt_to_aff(sa, sg);
    g1 ra, ea; g2 rs, es;
    rlc_scale_pair(ra, rs, pk, sa, c);
    pt_mul(ea, pk, s_words, 8); pt_mul(es, sg, s_words, 8);
    return (pt_eq(ra, ea) ? 1 : 0) | (pt_eq(rs, es) ? 2 : 0);
}
// Jacobi-symbol Legendre against the exponentiation on n values (Montgomery limbs in, 12 words each); returns mismatches
extern "C" int emu_legendre_check(const uint32_t* vals, int n) {
    int bad = 0;
    for (int i = 0; i < n; i++) { fp a; for (int j = 0; j < 12; j++) a.l[j] = vals[12 * i + j]; if (fp_legendre(a) != fp_legendre_pow(a)) bad++; }
    return bad;
}

// Karatsuba wide product against the schoolbook one on n operand pairs (12 words each, any values); returns mismatches
extern "C" int emu_mul_wide_k_check(const uint32_t* a, const uint32_t* b, int n) {
    int bad = 0;
    for (int i = 0; i < n; i++) {
        uint32_t T1[24], T2[24];
        mul_wide(T1, a + 12 * i, b + 12 * i); mul_wide_k(T2, a + 12 * i, b + 12 * i);
        for (int j = 0; j < 24; j++) if (T1[j] != T2[j]) { bad++; break; }
    }
    return bad;
}
// fp_half (conditional add of p + shift) on n values (12 words each, canonical): h + h == a and h == a * (1/2); the twist-constant
// product by additions == the product by K_B2_3; and the doubling step built from them == the round-1 form that multiplied by
// the constants 1/2 and 3b' (same point, same line).  Returns mismatches.
extern "C" int emu_half_and_dbl_check(const uint32_t* vals, int n) {
    int bad = 0;
    fp inv2; fp_set(inv2, K_INV2);
    for (int i = 0; i < n; i++) {
        fp a, h, d, m; for (int j = 0; j < 12; j++) a.l[j] = vals[12 * i + j];
        fp_half(h, a); fp_add(d, h, h); fp_mul(m, a, inv2);
        if (!fp_eq(d, a) || !fp_eq(h, m)) bad++;
    }
    for (int i = 0; i + 5 < n; i += 6) {
        fp2 c[3]; for (int k = 0; k < 3; k++) for (int j = 0; j < 12; j++) { c[k].a.l[j] = vals[12 * (i + 2 * k) + j]; c[k].b.l[j] = vals[12 * (i + 2 * k + 1) + j]; }
        fp2 b3, e1, e2; fp2_const(b3, K_B2_3); fp2_mul(e1, b3, c[0]); fp2_mul_twist3b(e2, c[0]);
        if (!fp2_eq(e1, e2)) bad++;
        // round-1 doubling step (multiplications by the constants), restated
        g2proj t; t.x = c[0]; t.y = c[1]; t.z = c[2];
        fp2 A, B, C, Ee, F, H, s, l0, l2, l3, x3, y3, e2s, z3;
        fp2_mul(A, t.x, t.y); fp2_mul_fp(A, A, inv2);
        fp2_sqr(B, t.y); fp2_sqr(C, t.z); fp2_mul(Ee, b3, C);
        fp2_dbl(F, Ee); fp2_add(F, F, Ee);
        fp2_add(H, t.y, t.z); fp2_sqr(H, H); fp2_sub(H, H, B); fp2_sub(H, H, C);
        fp2_sub(l0, B, Ee); fp2_sqr(s, t.x); fp2_dbl(l2, s); fp2_add(l2, l2, s); fp2_neg(l2, l2); l3 = H;
        fp2_sub(x3, B, F); fp2_mul(x3, x3, A);
        fp2_add(y3, B, F); fp2_mul_fp(y3, y3, inv2); fp2_sqr(y3, y3);
        fp2_sqr(e2s, Ee); fp2_dbl(s, e2s); fp2_add(s, s, e2s); fp2_sub(y3, y3, s);
        fp2_mul(z3, B, H);
        fp2 n0, n2, n3; ml_dbl(t, n0, n2, n3);
        if (!fp2_eq(t.x, x3) || !fp2_eq(t.y, y3) || !fp2_eq(t.z, z3) || !fp2_eq(n0, l0) || !fp2_eq(n2, l2) || !fp2_eq(n3, l3)) bad++;
    }
    return bad;
}
// binary-GCD inversion against the Fermat exponentiation on n values (Montgomery limbs in, 12 words each); returns mismatches
extern "C" int emu_inv_gcd_check(const uint32_t* vals, int n) {
    int bad = 0;
    for (int i = 0; i < n; i++) {
        fp a, x, y; for (int j = 0; j < 12; j++) a.l[j] = vals[12 * i + j];
        fp_inv_gcd(x, a); fp_inv(y, a);
        if (!fp_eq(x, y)) bad++;
        fp one, t; fp_one(one); fp_mul(t, x, a);
        if (!fp_is_zero(a) && !fp_eq(t, one)) bad++;
    }
    return bad;
}

// ---- lane-pair (fp2h) code on two host threads: the same templates the split kernels instantiate, shuffles by rendezvous.
// kind 0: exact round  -- miller_loop2<fp2h>(B, sig, -pk, H) + final_exp, compared with the single-thread fp2 form
// kind 1: batched group of 4 -- miller_loop_multi<fp2h, 5> on the pairs produced by rlc_scale_pair, compared likewise
// returns (split verdict) | (single-thread verdict << 1) | (Fp12 values identical << 2), or -1 on undecodable input
extern "C" int emu_split_pairing(int kind, const uint8_t* pks48, const uint8_t* sigs96, const uint8_t* msgs, uint32_t len) {
    const int G = kind ? 4 : 1, NPmax = 5;
    g1a P[NPmax]; fp2 qx[NPmax], qy[NPmax]; const g1a* ps[NPmax];
    int np;
    if (kind == 0) {
        g1 pk; g2 sg, h;
        if (!g1_deserialize(pk, pks48, true) || !g2_deserialize(sg, sigs96, true) || !map_to_g2(h, msgs, len)) return -1;
        g2a sa, ha; pt_to_aff(sa, sg); pt_to_aff(ha, h);
        fp_set(P[0].x, K_G1_X); fp_set(P[0].y, K_G1_Y); qx[0] = sa.x; qy[0] = sa.y;
        pt_to_aff(P[1], pk); fp_neg(P[1].y, P[1].y); qx[1] = ha.x; qy[1] = ha.y;
        np = 2;
    } else {
        g2 acc; pt_set_inf(acc);
        for (int k = 0; k < G; k++) {
            g1 pk; g2 sg, h;
            if (!g1_deserialize(pk, pks48 + 48 * k, true) || !g2_deserialize(sg, sigs96 + 96 * k, true) || !map_to_g2(h, msgs + len * k, len)) return -1;
            g2a sa, ha; pt_to_aff(sa, sg); pt_to_aff(ha, h);
            g1 rp; g2 rs; rlc_scale_pair(rp, rs, pk, sa, 0x9e3779b97f4a7c15ull * (k + 3));
            pt_to_aff(P[k], rp); fp_neg(P[k].y, P[k].y); qx[k] = ha.x; qy[k] = ha.y; pt_add(acc, acc, rs);
        }
        g2a sga; pt_to_aff(sga, acc); qx[G] = sga.x; qy[G] = sga.y;
        fp_set(P[G].x, K_G1_X); fp_set(P[G].y, K_G1_Y);
        np = G + 1;
    }
    for (int k = 0; k < np; k++) ps[k] = &P[k];
    // single-thread reference with the same templates over the plain Fp2 carrier
    fp12 ref;
    if (np == 2) { g2a q0, q1; q0.x = qx[0]; q0.y = qy[0]; q1.x = qx[1]; q1.y = qy[1]; miller_loop2(ref, P[0], q0, P[1], q1); }
    else miller_loop_multi<fp2, 5>(ref, ps, qx, qy);
    final_exp(ref, ref);
    const int ref_one = fp12_is_one(ref) ? 1 : 0;
    // the two lanes
    fp12_t<fp2h> out[2]; int verdict[2] = {0, 0};
    hb_emu_pair_reset();
    auto lane = [&](int role) {
        hb_emu.role = role; hb_emu.seq = 0;
        fp2h hx[NPmax], hy[NPmax];
        for (int k = 0; k < np; k++) { hx[k].c = role ? qx[k].b : qx[k].a; hy[k].c = role ? qy[k].b : qy[k].a; }
        fp12_t<fp2h> m;
        if (np == 2) miller_loop2<fp2h>(m, P[0], hx[0], hy[0], P[1], hx[1], hy[1], true, true);
        else miller_loop_multi<fp2h, 5>(m, ps, hx, hy);
        final_exp(m, m);
        verdict[role] = fp12_is_one(m) ? 1 : 0;
        out[role] = m;
    };
    std::thread t1(lane, 1); lane(0); t1.join();
    hb_emu.role = 0;
    // lane 0 holds the real parts, lane 1 the imaginary parts of the 6 Fp2 coefficients
    const fp2* rc[6] = {&ref.c0.c0, &ref.c0.c1, &ref.c0.c2, &ref.c1.c0, &ref.c1.c1, &ref.c1.c2};
    const fp2h* l0[6] = {&out[0].c0.c0, &out[0].c0.c1, &out[0].c0.c2, &out[0].c1.c0, &out[0].c1.c1, &out[0].c1.c2};
    const fp2h* l1[6] = {&out[1].c0.c0, &out[1].c0.c1, &out[1].c0.c2, &out[1].c1.c0, &out[1].c1.c1, &out[1].c1.c2};
    bool same = true;
    for (int i = 0; i < 6; i++) same = same && fp_eq(rc[i]->a, l0[i]->c) && fp_eq(rc[i]->b, l1[i]->c);
    if (verdict[0] != verdict[1]) return -2;
    return verdict[0] | (ref_one << 1) | ((same ? 1 : 0) << 2);
}
