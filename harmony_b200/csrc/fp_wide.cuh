// harmony_b200/csrc/fp_wide.cuh -- unreduced 768-bit products + one Montgomery reduction ("lazy reduction").
//
// Fp2 products are the unit of work of the pairing (Miller loop / final exponentiation are ~100% Fp2 mul/sqr), so
// they are computed register-resident in one piece:
//     (a0 + a1 i)(b0 + b1 i):  t0 = a0 b0, t1 = a1 b1, t2 = (a0+a1)(b0+b1) as 24-limb integers,
//                              c0 = redc(t0 + (p^2 - t1)),  c1 = redc(t2 - t0 - t1)
// = 3 x 144 + 2 x 156 = 744 IMAD.WIDE instead of 3 x 300, no intermediate trips through local memory, and three
// independent carry-chain families for the scheduler to interleave.
#pragma once
#include "fp.cuh"

namespace hb {

template <int N> HB_DEV void lane_mad_n(uint32_t* acc, const uint32_t* a, uint32_t b);
// T[0..23] = a * b (plain integer product of two 12-limb values, any a, b < 2^384)
HB_DEV void mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t x[26], y[26];          // x: 64-bit lanes at even limb positions, y: lanes at odd positions (absolute)
#pragma unroll
    for (int i = 0; i < 26; i++) { x[i] = 0; y[i] = 0; }
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
        lane_mad(x + i, a, b[i]);               // a_even * b_i   -> even positions
        lane_mad(y + i + 1, a + 1, b[i]);       // a_odd  * b_i   -> odd positions
        lane_mad(y + i + 1, a, b[i + 1]);       // a_even * b_i+1 -> odd positions
        lane_mad(x + i + 2, a + 1, b[i + 1]);   // a_odd  * b_i+1 -> even positions
    }
    T[0] = x[0];
    add_cc(T[1], x[1], y[1]);
#pragma unroll
    for (int j = 2; j < 23; j++) addc_cc(T[j], x[j], y[j]);
    addc(T[23], x[23], y[23]);
}

// T[0..11] = a * b for 6-limb operands (same even / odd lane scheme as mul_wide): 36 IMAD.WIDE
HB_DEV void mul_wide6(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t x[14], y[14];
#pragma unroll
    for (int i = 0; i < 14; i++) { x[i] = 0; y[i] = 0; }
#pragma unroll
    for (int i = 0; i < 6; i += 2) {
        lane_mad_n<3>(x + i, a, b[i]);
        lane_mad_n<3>(y + i + 1, a + 1, b[i]);
        lane_mad_n<3>(y + i + 1, a, b[i + 1]);
        lane_mad_n<3>(x + i + 2, a + 1, b[i + 1]);
    }
    T[0] = x[0];
    add_cc(T[1], x[1], y[1]);
#pragma unroll
    for (int j = 2; j < 11; j++) addc_cc(T[j], x[j], y[j]);
    addc(T[11], x[11], y[11]);
}
// T[0..23] = a * b by one level of Karatsuba over the 6-limb halves: 3 x 36 = 108 IMAD.WIDE instead of 144, paid with ~85 adds /
// selects on the ALU pipe (which idles at ~1/3 while the FMA-heavy pipe is the bound: profiles/r2_probe_int.txt).
HB_DEV void mul_wide_k(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t z0[12], z2[12], s[6], t[6], m[12], z1[13];
    mul_wide6(z0, a, b);
    mul_wide6(z2, a + 6, b + 6);
    uint32_t cs, ct;
    add_cc(s[0], a[0], a[6]);
#pragma unroll
    for (int j = 1; j < 6; j++) addc_cc(s[j], a[j], a[6 + j]);
    addc(cs, 0, 0);
    add_cc(t[0], b[0], b[6]);
#pragma unroll
    for (int j = 1; j < 6; j++) addc_cc(t[j], b[j], b[6 + j]);
    addc(ct, 0, 0);
    mul_wide6(m, s, t);
    // z1 = (s + cs 2^192)(t + ct 2^192) - z0 - z2 = a0 b1 + a1 b0 < 2^385: 13 limbs
    const uint32_t ms = 0u - cs, mt = 0u - ct;
#pragma unroll
    for (int j = 0; j < 6; j++) z1[j] = m[j];
    add_cc(z1[6], m[6], t[0] & ms);
#pragma unroll
    for (int j = 1; j < 6; j++) addc_cc(z1[6 + j], m[6 + j], t[j] & ms);
    addc(z1[12], cs & ct, 0);
    add_cc(z1[6], z1[6], s[0] & mt);
#pragma unroll
    for (int j = 1; j < 6; j++) addc_cc(z1[6 + j], z1[6 + j], s[j] & mt);
    addc(z1[12], z1[12], 0);
    sub_cc(z1[0], z1[0], z0[0]);
#pragma unroll
    for (int j = 1; j < 12; j++) subc_cc(z1[j], z1[j], z0[j]);
    subc(z1[12], z1[12], 0);
    sub_cc(z1[0], z1[0], z2[0]);
#pragma unroll
    for (int j = 1; j < 12; j++) subc_cc(z1[j], z1[j], z2[j]);
    subc(z1[12], z1[12], 0);
    // T = z0 + z1 2^192 + z2 2^384
#pragma unroll
    for (int j = 0; j < 6; j++) T[j] = z0[j];
    add_cc(T[6], z0[6], z1[0]);
#pragma unroll
    for (int j = 1; j < 6; j++) addc_cc(T[6 + j], z0[6 + j], z1[j]);
#pragma unroll
    for (int j = 0; j < 6; j++) addc_cc(T[12 + j], z2[j], z1[6 + j]);
    addc_cc(T[18], z2[6], z1[12]);
#pragma unroll
    for (int j = 7; j < 11; j++) addc_cc(T[12 + j], z2[j], 0);
    addc(T[23], z2[11], 0);
}

// T[0..23] = a1 * b1 + a2 * b2 accumulated in ONE pair of lane accumulators (one merge instead of two + a wide add);
// caller guarantees the sum < 2^768
HB_DEV void mul_wide2(uint32_t* T, const uint32_t* a1, const uint32_t* b1, const uint32_t* a2, const uint32_t* b2) {
    uint32_t x[26], y[26];
#pragma unroll
    for (int i = 0; i < 26; i++) { x[i] = 0; y[i] = 0; }
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
        lane_mad(x + i, a1, b1[i]);
        lane_mad(y + i + 1, a1 + 1, b1[i]);
        lane_mad(y + i + 1, a1, b1[i + 1]);
        lane_mad(x + i + 2, a1 + 1, b1[i + 1]);
        lane_mad(x + i, a2, b2[i]);
        lane_mad(y + i + 1, a2 + 1, b2[i]);
        lane_mad(y + i + 1, a2, b2[i + 1]);
        lane_mad(x + i + 2, a2 + 1, b2[i + 1]);
    }
    T[0] = x[0];
    add_cc(T[1], x[1], y[1]);
#pragma unroll
    for (int j = 2; j < 23; j++) addc_cc(T[j], x[j], y[j]);
    addc(T[23], x[23], y[23]);
}

#ifndef HB_KARATSUBA
#define HB_KARATSUBA 0      // 1: products through mul_wide_k (108 IMAD.WIDE) + a separate reduction instead of the interleaved 288 + 12 form
#endif
// T = a1 * b1 + a2 * b2 from two Karatsuba products (sum < 2^768 by the caller's contract)
HB_DEV void mul_wide2_k(uint32_t* T, const uint32_t* a1, const uint32_t* b1, const uint32_t* a2, const uint32_t* b2) {
    uint32_t U[24];
    mul_wide_k(T, a1, b1); mul_wide_k(U, a2, b2);
    add_cc(T[0], T[0], U[0]);
#pragma unroll
    for (int j = 1; j < 23; j++) addc_cc(T[j], T[j], U[j]);
    addc(T[23], T[23], U[23]);
}
#if HB_KARATSUBA
#define HB_MUL_WIDE mul_wide_k
#define HB_MUL_WIDE2 mul_wide2_k
#else
#define HB_MUL_WIDE mul_wide
#define HB_MUL_WIDE2 mul_wide2
#endif
// r[0..11] = T / 2^384 mod p for T < p * 2^384; result canonical in [0, p).  T is consumed.
HB_DEV void redc_wide(uint32_t* r, const uint32_t* T) {
    uint32_t x[28], y[28];
#pragma unroll
    for (int i = 0; i < 28; i++) { x[i] = i < 12 ? T[i] : 0; y[i] = 0; }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int q = 2 * k;
        const uint32_t m0 = x[q] * HB_N0;
        lane_mad_p<0>(x + q, m0);
        lane_mad_p<1>(y + q + 1, m0);
        const uint32_t m1 = (x[q + 1] + y[q + 1]) * HB_N0;
        lane_mad_p<0>(y + q + 1, m1);
        lane_mad_p<1>(x + q + 2, m1);
        uint32_t dead;
        add_cc(dead, x[q + 1], y[q + 1]);
        addc_cc(x[q + 2], x[q + 2], y[q + 2]);
#pragma unroll
        for (int j = q + 3; j < q + 15; j++) addc_cc(x[j], x[j], 0);
        addc(x[q + 15], x[q + 15], 0);
        (void)dead;
    }
    // (T_low + m p) / R = X[12..] + Y[13..] ; then add T_high
    add_cc(x[13], x[13], y[13]);
#pragma unroll
    for (int j = 14; j < 24; j++) addc_cc(x[j], x[j], y[j]);
    add_cc(x[12], x[12], T[12]);
#pragma unroll
    for (int j = 13; j < 23; j++) addc_cc(x[j], x[j], T[j]);
    addc(x[23], x[23], T[23]);
    uint32_t s[12];
    sub_cc(s[0], x[12], HB_P0);
#pragma unroll
    for (int j = 1; j < 12; j++) subc_cc(s[j], x[12 + j], p_limb(j));
    uint32_t borrow;
    subc(borrow, 0, 0);
#pragma unroll
    for (int j = 0; j < 12; j++) r[j] = borrow ? x[12 + j] : s[j];
}

// acc[0..2N-1] (N 64-bit lanes) += {a[0], a[2], ..., a[2N-2]} * b ; carry rippled into acc[2N], acc[2N+1]
template <int N> HB_DEV void lane_mad_n(uint32_t* acc, const uint32_t* a, uint32_t b) {
    if (N == 0) return;
    mad_lo_cc(acc[0], a[0], b, acc[0]);
    madc_hi_cc(acc[1], a[0], b, acc[1]);
#pragma unroll
    for (int j = 1; j < N; j++) {
        madc_lo_cc(acc[2 * j], a[2 * j], b, acc[2 * j]);
        madc_hi_cc(acc[2 * j + 1], a[2 * j], b, acc[2 * j + 1]);
    }
    addc(acc[2 * N], acc[2 * N], 0);
}
template <int I> HB_DEV void sqr_row(uint32_t* x, uint32_t* y, const uint32_t* a) {
    // off-diagonal products a_I * a_j, j > I: odd distance -> Y lanes at 2I+1, even distance -> X lanes at 2I+2
    lane_mad_n<(12 - I) / 2>(y + 2 * I + 1, a + I + 1, a[I]);
    lane_mad_n<(11 - I) / 2>(x + 2 * I + 2, a + I + 2, a[I]);
}
// T[0..23] = a^2: 66 off-diagonal products once, doubled by a 1-bit shift, plus 12 diagonal squares (78 IMAD.WIDE)
HB_DEV void sqr_wide(uint32_t* T, const uint32_t* a) {
    uint32_t x[28], y[28];
#pragma unroll
    for (int i = 0; i < 28; i++) { x[i] = 0; y[i] = 0; }
    sqr_row<0>(x, y, a); sqr_row<1>(x, y, a); sqr_row<2>(x, y, a); sqr_row<3>(x, y, a);
    sqr_row<4>(x, y, a); sqr_row<5>(x, y, a); sqr_row<6>(x, y, a); sqr_row<7>(x, y, a);
    sqr_row<8>(x, y, a); sqr_row<9>(x, y, a); sqr_row<10>(x, y, a);
    uint32_t s[24];
    s[0] = 0;
    add_cc(s[1], x[1], y[1]);
#pragma unroll
    for (int j = 2; j < 23; j++) addc_cc(s[j], x[j], y[j]);
    addc(s[23], x[23], y[23]);
    uint32_t d[24];                       // 2 * S
    d[0] = 0;
#pragma unroll
    for (int j = 1; j < 24; j++) d[j] = (s[j] << 1) | (s[j - 1] >> 31);
    mad_lo_cc(T[0], a[0], a[0], d[0]);
    madc_hi_cc(T[1], a[0], a[0], d[1]);
#pragma unroll
    for (int j = 1; j < 12; j++) {
        madc_lo_cc(T[2 * j], a[j], a[j], d[2 * j]);
        madc_hi_cc(T[2 * j + 1], a[j], a[j], d[2 * j + 1]);
    }
}

// p^2 as 24 little-endian 32-bit limbs (immediates)
HB_DEV uint32_t p2_limb(int i) {
    switch (i) {
    case 0: return 0x1c718e39u; case 1: return 0x26aa0000u; case 2: return 0x76382eabu; case 3: return 0x7ced6b1du;
    case 4: return 0x62113cfdu; case 5: return 0x162c3383u; case 6: return 0x3e71b743u; case 7: return 0x66bf91edu;
    case 8: return 0x7091a049u; case 9: return 0x292e85a8u; case 10: return 0x86185c7bu; case 11: return 0x1d68619cu;
    case 12: return 0x0978ef01u; case 13: return 0xf5314933u; case 14: return 0x16ddca6eu; case 15: return 0x50a62cfdu;
    case 16: return 0x349e8bd0u; case 17: return 0x66e59e49u; case 18: return 0x0e7046b4u; case 19: return 0xe2dc90e5u;
    case 20: return 0xa22f25e9u; case 21: return 0x4bd278eau; case 22: return 0xb8c35fc7u; default: return 0x02a437a4u;
    }
}

// 24-limb helpers (no reduction)
HB_DEV void wide_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {     // r = a - b (caller guarantees a >= b)
    sub_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int j = 1; j < 23; j++) subc_cc(r[j], a[j], b[j]);
    subc(r[23], a[23], b[23]);
}
HB_DEV void wide_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    add_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int j = 1; j < 23; j++) addc_cc(r[j], a[j], b[j]);
    addc(r[23], a[23], b[23]);
}
HB_DEV void wide_p2_minus(uint32_t* r, const uint32_t* b) {                    // r = p^2 - b, b <= p^2
    sub_cc(r[0], p2_limb(0), b[0]);
#pragma unroll
    for (int j = 1; j < 23; j++) subc_cc(r[j], p2_limb(j), b[j]);
    subc(r[23], p2_limb(23), b[23]);
}
HB_DEV void limbs_add12(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // r = a + b, no reduction (< 2^384 by contract)
    add_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int j = 1; j < 11; j++) addc_cc(r[j], a[j], b[j]);
    addc(r[11], a[11], b[11]);
}
// r = a - b + p (in (0, 2p)), no reduction
HB_DEV void limbs_sub12_plus_p(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t t[12];
    sub_cc(t[0], a[0], b[0]);
#pragma unroll
    for (int j = 1; j < 11; j++) subc_cc(t[j], a[j], b[j]);
    subc(t[11], a[11], b[11]);
    add_cc(r[0], t[0], HB_P0);
#pragma unroll
    for (int j = 1; j < 11; j++) addc_cc(r[j], t[j], p_limb(j));
    addc(r[11], t[11], HB_P11);
}

#if HB_KARATSUBA
HB_NOINLINE void fp_mul(fp& r, const fp& a, const fp& b) {
    HB_EMU_COUNT(hb_emu_cnt_mul);
    uint32_t ra[12], rb[12], T[24], rr[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { ra[j] = a.l[j]; rb[j] = b.l[j]; }
    mul_wide_k(T, ra, rb);
    redc_wide(rr, T);
#pragma unroll
    for (int j = 0; j < 12; j++) r.l[j] = rr[j];
}
#endif
// r = a^2 / R mod p: 78 + 156 = 234 IMAD.WIDE (the exponentiation chains of sqrt / inverse / Legendre are ~80% squarings)
HB_NOINLINE void fp_sqr(fp& r, const fp& a) {
    HB_EMU_COUNT(hb_emu_cnt_sqr);
    uint32_t ra[12], T[24], rr[12];
#pragma unroll
    for (int j = 0; j < 12; j++) ra[j] = a.l[j];
    sqr_wide(T, ra);
    redc_wide(rr, T);
#pragma unroll
    for (int j = 0; j < 12; j++) r.l[j] = rr[j];
}

// (ra + rb i) = (xa + xb i)(ya + yb i), all canonical
HB_DEV void fp2_mul_regs(uint32_t* ra, uint32_t* rb, const uint32_t* xa, const uint32_t* xb, const uint32_t* ya, const uint32_t* yb) {
    uint32_t t0[24], t1[24], t2[24], s0[12], s1[12];
    mul_wide(t0, xa, ya);
    mul_wide(t1, xb, yb);
    limbs_add12(s0, xa, xb);
    limbs_add12(s1, ya, yb);
    mul_wide(t2, s0, s1);
    wide_sub(t2, t2, t0);
    wide_sub(t2, t2, t1);                 // a0 b1 + a1 b0 < 2 p^2
    redc_wide(rb, t2);
    wide_p2_minus(t1, t1);
    wide_add(t0, t0, t1);                 // a0 b0 - a1 b1 + p^2 in (0, 2 p^2)
    redc_wide(ra, t0);
}
// (ra + rb i) = (xa + xb i)^2
HB_DEV void fp2_sqr_regs(uint32_t* ra, uint32_t* rb, const uint32_t* xa, const uint32_t* xb) {
    uint32_t t0[24], t1[24], s[12], d[12];
    limbs_add12(s, xa, xb);               // < 2p
    limbs_sub12_plus_p(d, xa, xb);        // in (0, 2p)
    mul_wide(t0, s, d);                   // < 4 p^2 < p R
    mul_wide(t1, xa, xb);
    wide_add(t1, t1, t1);                 // 2 a0 a1 < 2 p^2
    redc_wide(ra, t0);
    redc_wide(rb, t1);
}

// variant 2: register-resident Karatsuba built from three interleaved-reduction (CIOS) products, canonical add/sub
HB_DEV void fp2_mul_regs_cios(uint32_t* ra, uint32_t* rb, const uint32_t* xa, const uint32_t* xb, const uint32_t* ya, const uint32_t* yb) {
    uint32_t t0[12], t1[12], t2[12], s0[12], s1[12];
    fp_mul_regs(t0, xa, ya);
    fp_mul_regs(t1, xb, yb);
    fp_add_regs(s0, xa, xb);
    fp_add_regs(s1, ya, yb);
    fp_mul_regs(t2, s0, s1);
    fp_sub_regs(ra, t0, t1);
    fp_sub_regs(t2, t2, t0);
    fp_sub_regs(rb, t2, t1);
}
HB_DEV void fp2_sqr_regs_cios(uint32_t* ra, uint32_t* rb, const uint32_t* xa, const uint32_t* xb) {
    uint32_t s[12], d[12], m[12];
    fp_add_regs(s, xa, xb);
    fp_sub_regs(d, xa, xb);
    fp_mul_regs(m, xa, xb);
    fp_mul_regs(ra, s, d);
    fp_add_regs(rb, m, m);
}

}  // namespace hb
