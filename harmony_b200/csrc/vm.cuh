// harmony_b200/csrc/vm.cuh -- warp-cooperative ("latency mode") pairing: ONE warp = 16 lane pairs verifies one round.
//
// The reference verifies one aggregate per block on the consensus path (consensus/validator.go:219-236, internal/chain/engine.go:
// 619-642): there the figure of merit is the latency of a single check, and a thread-per-round (or lane-pair-per-round) pairing is
// ~10^4 dependent field products long.  Here the Miller loop and the final exponentiation run as straight-line STEP PROGRAMS over
// Fp2 values held in shared memory (vm_programs.cuh, generated and CPU-verified by tools/vmgen.py): in every step each of the 16
// lane pairs of the warp executes one Fp2 operation -- a product, a squaring, or a small-integer linear combination -- so up to 16
// independent Fp2 products are in flight.  The whole working set (150 slots x 100 B) stays in shared memory: no local-memory stack.
// Linear combinations are stored per lane role as lists of atoms (+- m x one half of one slot) so that all 32 lanes of a step run
// the same branch-free loop and reduce once.
//
// Slot layout: slot s = 25 words at s * 25: real part (12 Montgomery limbs), imaginary part (12), 1 pad word (odd stride: the 16
// pairs of a step read 16 different slots without systematic bank conflicts).  Lane 2k is the "real" lane of pair k, lane 2k+1 the
// "imaginary" lane: both read the full operands from shared memory (no shuffles) and each writes its half of the result.
#pragma once
#include "curve.cuh"
#include "vm_programs.cuh"

namespace hb {

#define VM_SLOT_WORDS 25
#define VM_SMEM_WORDS (VM_NSLOTS * VM_SLOT_WORDS)

HB_DEV void vm_ld(uint32_t* r, const uint32_t* slots, int s, int half) {
    const uint32_t* p = slots + s * VM_SLOT_WORDS + half * 12;
#pragma unroll
    for (int j = 0; j < 12; j++) r[j] = p[j];
}
HB_DEV void vm_st(uint32_t* slots, int s, int half, const uint32_t* r) {
    uint32_t* p = slots + s * VM_SLOT_WORDS + half * 12;
#pragma unroll
    for (int j = 0; j < 12; j++) p[j] = r[j];
}
// dst = a * b: real lane a.re b.re + a.im (p - b.im), imaginary lane a.re b.im + a.im b.re -- two wide products in one accumulator
// pair and one reduction per lane (the lane-pair product of tower.cuh without the shuffles)
HB_NOINLINE void vm_mul(uint32_t* slots, int dst, int a, int b, int im) {
    uint32_t ar[12], ai[12], br[12], bi[12], y1[12], y2[12], T[24], rr[12];
    vm_ld(ar, slots, a, 0); vm_ld(ai, slots, a, 1); vm_ld(br, slots, b, 0); vm_ld(bi, slots, b, 1);
    uint32_t nbi[12];                                  // p - b.im in (0, p]
    sub_cc(nbi[0], HB_P0, bi[0]);
#pragma unroll
    for (int j = 1; j < 11; j++) subc_cc(nbi[j], p_limb(j), bi[j]);
    subc(nbi[11], HB_P11, bi[11]);
#pragma unroll
    for (int j = 0; j < 12; j++) { y1[j] = im ? bi[j] : br[j]; y2[j] = im ? br[j] : nbi[j]; }
    HB_MUL_WIDE2(T, ar, y1, ai, y2);                   // < 2 p^2 < p R
    redc_wide(rr, T);
    vm_st(slots, dst, im, rr);
}
// dst = a^2: real lane (a.re + a.im)(a.re - a.im), imaginary lane 2 a.re a.im
HB_NOINLINE void vm_sqr(uint32_t* slots, int dst, int a, int im) {
    uint32_t ar[12], ai[12], s[12], d[12], t[12], A[12], B[12], T[24], rr[12];
    vm_ld(ar, slots, a, 0); vm_ld(ai, slots, a, 1);
    limbs_add12(s, ar, ai);            // < 2p
    limbs_sub12_plus_p(d, ar, ai);     // in (0, 2p)
    limbs_add12(t, ai, ai);            // < 2p
#pragma unroll
    for (int j = 0; j < 12; j++) { A[j] = im ? ar[j] : s[j]; B[j] = im ? t[j] : d[j]; }
    HB_MUL_WIDE(T, A, B);              // < 4 p^2 < p R
    redc_wide(rr, T);
    vm_st(slots, dst, im, rr);
}
// k * p as 13 little-endian words, k = 1, 2, 4, 8, 16 (reduction ladder of vm_lin)
__device__ __constant__ const uint32_t VM_KP[5][13] = {
    {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau, 0x00000000u},
    {0xffff5556u, 0x73fdffffu, 0x62a7ffffu, 0x3d57fffdu, 0xed61ec48u, 0xce61a541u, 0xe70a257eu, 0xc8ee9709u, 0x869759aeu, 0x96374f6cu, 0x72ffcd34u, 0x340223d4u, 0x00000000u},
    {0xfffeaaacu, 0xe7fbffffu, 0xc54ffffeu, 0x7aaffffau, 0xdac3d890u, 0x9cc34a83u, 0xce144afdu, 0x91dd2e13u, 0x0d2eb35du, 0x2c6e9ed9u, 0xe5ff9a69u, 0x680447a8u, 0x00000000u},
    {0xfffd5558u, 0xcff7ffffu, 0x8a9ffffdu, 0xf55ffff5u, 0xb587b120u, 0x39869507u, 0x9c2895fbu, 0x23ba5c27u, 0x1a5d66bbu, 0x58dd3db2u, 0xcbff34d2u, 0xd0088f51u, 0x00000000u},
    {0xfffaaab0u, 0x9fefffffu, 0x153ffffbu, 0xeabfffebu, 0x6b0f6241u, 0x730d2a0fu, 0x38512bf6u, 0x4774b84fu, 0x34bacd76u, 0xb1ba7b64u, 0x97fe69a4u, 0xa0111ea3u, 0x00000001u}};
// dst = sum_k M_k src_k, stored per lane role as <= 8 atoms (slot, half, sign, multiplier 1..4): every lane of the step runs the same
// `natoms` iterations (its unused atoms point at the ZERO slot).
//
// Latency form (one warp is alone on its scheduler, so DEPENDENT instruction chains are what costs): the atoms are accumulated
// WITHOUT carry chains -- each 32-bit limb of a source is split into its 16-bit halves and multiplied into 24 independent 32-bit
// accumulators (IMAD by the signed multiplier) -- and only then folded into 13 limbs with one carry chain.  The accumulators start
// at INIT = 2^22 + (16-bit digits of X), X = 631 p - 2^22 * sum_k 2^(16 k): 2^22 absorbs the negative atoms (|sum| <= 8 * 4 * 65535
// < 2^21 per accumulator) and the whole offset is exactly 631 p, so the folded value is (sum + 631 p) in (599 p, 663 p).  It is
// reduced by one quotient estimate from the top 42 bits (floor(top / (p_11 + 1)) never overshoots and leaves < 2 p) and two trial
// subtractions.  Measured (tools/lat_probe.cu, profiles/r2_lat_probe.txt): 4-atom step 1 759 -> 1 332 cycles, 8-atom step 2 871 -> 1 750; a Miller doubling
// iteration (14 steps) 42.5k -> 35.6k cycles.
#ifndef HB_VM_LIN_V1
#define HB_VM_LIN_V1 0
#endif
#if !HB_VM_LIN_V1
HB_NOINLINE void vm_lin(uint32_t* slots, int dst, const uint4 row, int natoms, int im) {
    uint32_t lo[12] = {0x40ab7du, 0x40ffbfu, 0x40ff13u, 0x40fc87u, 0x40b2c7u, 0x402a80u, 0x403587u, 0x402474u, 0x4006a8u, 0x4061ffu, 0x40660fu, 0x402813u};
    uint32_t hi[12] = {0x40feedu, 0x407348u, 0x4015cbu, 0x4099b3u, 0x400deeu, 0x405917u, 0x403cc1u, 0x40a1cbu, 0x40df47u, 0x4020eau, 0x40ba01u, 0x401863u};
    const uint32_t rw[4] = {row.x, row.y, row.z, row.w};
#pragma unroll 2
    for (int a = 0; a < natoms; a++) {
        const uint32_t at = (rw[a >> 1] >> (16 * (a & 1))) & 0xffffu;
        const int s = at & 0xff, half = (at >> 8) & 1;
        const uint32_t m = ((at >> 10) & 3u) + 1u;
        const uint32_t sm = ((at >> 9) & 1u) ? 0u - m : m;              // signed multiplier, two's complement (wrap-around arithmetic)
        uint32_t v[12];
        vm_ld(v, slots, s, half);
#pragma unroll
        for (int j = 0; j < 12; j++) { lo[j] += sm * (v[j] & 0xffffu); hi[j] += sm * (v[j] >> 16); }
    }
    // fold: u_j = lo_j + 2^16 hi_j (< 2^40), value = sum_j u_j 2^(32 j)
    uint32_t low[12], high[12], acc[13];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        low[j] = lo[j] + (hi[j] << 16);
        high[j] = (hi[j] >> 16) + (low[j] < lo[j] ? 1u : 0u);
    }
    acc[0] = low[0];
    add_cc(acc[1], low[1], high[0]);
#pragma unroll
    for (int j = 2; j < 12; j++) addc_cc(acc[j], low[j], high[j - 1]);
    addc(acc[12], high[11], 0u);
    // quotient estimate from the top 42 bits: q <= floor(value / p), value - q p < 2 p
    const uint64_t top = ((uint64_t)acc[12] << 32) | acc[11];
    const uint32_t q = (uint32_t)(top / 0x1a0111ebull);               // p_11 + 1 (constant divisor: multiply-high + shift)
    uint32_t pl[12], ph[12], qp[13];
#pragma unroll
    for (int j = 0; j < 12; j++) { pl[j] = q * p_limb(j); ph[j] = (uint32_t)(((uint64_t)q * p_limb(j)) >> 32); }
    qp[0] = pl[0];
    add_cc(qp[1], pl[1], ph[0]);
#pragma unroll
    for (int j = 2; j < 12; j++) addc_cc(qp[j], pl[j], ph[j - 1]);
    addc(qp[12], ph[11], 0u);
    sub_cc(acc[0], acc[0], qp[0]);
#pragma unroll
    for (int j = 1; j < 12; j++) subc_cc(acc[j], acc[j], qp[j]);
    subc(acc[12], acc[12], qp[12]);
#pragma unroll
    for (int k = 0; k < 2; k++) {                                     // < 2 p (+ rounding of the estimate): p, p
        uint32_t d[13], bw;
        sub_cc(d[0], acc[0], VM_KP[0][0]);
#pragma unroll
        for (int j = 1; j < 13; j++) subc_cc(d[j], acc[j], VM_KP[0][j]);
        subc(bw, 0, 0);                                               // all-ones when acc < p
#pragma unroll
        for (int j = 0; j < 13; j++) acc[j] = (acc[j] & bw) | (d[j] & ~bw);
    }
    vm_st(slots, dst, im, acc);
}
#else
HB_NOINLINE void vm_lin(uint32_t* slots, int dst, const uint4 row, int natoms, int im) {
    uint32_t acc[13];
#pragma unroll
    for (int j = 0; j < 13; j++) acc[j] = 0;
    const uint32_t rw[4] = {row.x, row.y, row.z, row.w};
#pragma unroll 1
    for (int a = 0; a < natoms; a++) {
        const uint32_t at = (rw[a >> 1] >> (16 * (a & 1))) & 0xffffu;
        const int s = at & 0xff, half = (at >> 8) & 1, sh = (at >> 10) & 3;             // multiplier m = sh + 1
        const uint32_t neg = 0u - ((at >> 9) & 1u);
        uint32_t v[12], w[12];
        vm_ld(v, slots, s, half);
        // w = neg ? p - v : v   (in [0, p])
        sub_cc(w[0], HB_P0, v[0]);
#pragma unroll
        for (int j = 1; j < 11; j++) subc_cc(w[j], p_limb(j), v[j]);
        subc(w[11], HB_P11, v[11]);
#pragma unroll
        for (int j = 0; j < 12; j++) w[j] = (w[j] & neg) | (v[j] & ~neg);
        // acc += m * w:  m = 1, 2, 4 -> one shifted addend; m = 3 -> 2 w + w
        const int shift = sh == 3 ? 2 : (sh == 0 ? 0 : 1);                               // sh: 0 -> x1, 1 -> x2, 2 -> x3 (= x2 + x1), 3 -> x4
        uint32_t t[13];
        t[0] = w[0] << shift;
#pragma unroll
        for (int j = 1; j < 12; j++) t[j] = __funnelshift_l(w[j - 1], w[j], shift);
        t[12] = __funnelshift_l(w[11], 0u, shift);
        add_cc(acc[0], acc[0], t[0]);
#pragma unroll
        for (int j = 1; j < 12; j++) addc_cc(acc[j], acc[j], t[j]);
        addc(acc[12], acc[12], t[12]);
        const uint32_t extra = sh == 2 ? 0xffffffffu : 0u;                               // x3: one more w
        add_cc(acc[0], acc[0], w[0] & extra);
#pragma unroll
        for (int j = 1; j < 12; j++) addc_cc(acc[j], acc[j], w[j] & extra);
        addc(acc[12], acc[12], 0u);
    }
    // acc < 32 p: subtract 16 p, 8 p, 4 p, 2 p, p, p where it fits
#pragma unroll
    for (int k = 5; k >= 0; k--) {
        const int row_k = k == 0 ? 0 : k - 1;                                              // ladder 16p, 8p, 4p, 2p, p, p
        uint32_t d[13], bw;
        sub_cc(d[0], acc[0], VM_KP[row_k][0]);
#pragma unroll
        for (int j = 1; j < 13; j++) subc_cc(d[j], acc[j], VM_KP[row_k][j]);
        subc(bw, 0, 0);                                                                    // all-ones when acc < k p
#pragma unroll
        for (int j = 0; j < 13; j++) acc[j] = (acc[j] & bw) | (d[j] & ~bw);
    }
    vm_st(slots, dst, im, acc);
}
#endif
// run one step program; every lane of the warp must call it (steps end in __syncwarp).  The instruction words of the next step
// are fetched while the current one executes.
HB_NOINLINE void vm_run(int prog, uint32_t* slots) {
    const int lane = threadIdx.x & 31, pair = lane >> 1, im = lane & 1;
    const int first = VM_PROG_FIRST[prog], n = VM_PROG_STEPS[prog];
    const uint4* ip = VM_INS + ((size_t)first * 16 + pair) * (VM_INS_WORDS / 4);
    uint32_t hdr = VM_STEP_HDR[first]; uint32_t w0 = ip[0].x; uint4 row = ip[1 + im];
    for (int st = 0; st < n; st++) {
        const uint32_t c_hdr = hdr, c_w0 = w0; const uint4 c_row = row;
        if (st + 1 < n) {
            ip += 16 * (VM_INS_WORDS / 4);
            hdr = VM_STEP_HDR[first + st + 1]; w0 = ip[0].x; row = ip[1 + im];
        }
        const int cls = c_hdr & 0xff, dst = c_w0 & 0xff;
        if (dst != 0xff) {
            if (cls == VM_OP_LIN) vm_lin(slots, dst, c_row, (int)(c_hdr >> 8), im);
            else if (cls == VM_OP_SQR) vm_sqr(slots, dst, (c_w0 >> 8) & 0xff, im);
            else vm_mul(slots, dst, (c_w0 >> 8) & 0xff, (c_w0 >> 16) & 0xff, im);
        }
        __syncwarp();
    }
}
HB_DEV void vm_set_fp(uint32_t* slots, int s, const fp& re) {        // slot = (re, 0); called by one lane
    uint32_t z[12];
#pragma unroll
    for (int j = 0; j < 12; j++) z[j] = 0;
    vm_st(slots, s, 0, re.l); vm_st(slots, s, 1, z);
}
HB_DEV void vm_set_fp2(uint32_t* slots, int s, const fp2& v) { vm_st(slots, s, 0, v.a.l); vm_st(slots, s, 1, v.b.l); }
// constants of the programs (0, 1, 3 b', 1/2, Frobenius coefficients): lanes share the 16 slots
HB_DEV void vm_load_consts(uint32_t* slots) {
    const int lane = threadIdx.x & 31;
    if (lane == 15) { fp z; fp_zero(z); vm_set_fp(slots, VM_R_ZERO, z); }
    if (lane == 0) { fp o; fp_one(o); vm_set_fp(slots, VM_R_ONE, o); }
    if (lane == 1) { fp2 b; fp2_const(b, K_B2_3); vm_set_fp2(slots, VM_R_TWIST3B, b); }
    if (lane == 2) { fp h; fp_set(h, K_INV2); vm_set_fp(slots, VM_R_INV2, h); }
    if (lane >= 3 && lane < 9) { fp2 g; fp2_const(g, K_FROB1[lane - 3]); vm_set_fp2(slots, VM_R_FROB1_0 + (lane - 3), g); }
    if (lane >= 9 && lane < 15) { fp g; fp_set(g, K_FROB2[lane - 9]); vm_set_fp(slots, VM_R_FROB2_0 + (lane - 9), g); }
    if (lane == 16) { fp2 c; fp2_const(c, K_PSI_CX); vm_set_fp2(slots, VM_R_PSI_CX, c); }
    if (lane == 17) { fp2 c; fp2_const(c, K_PSI_CY); vm_set_fp2(slots, VM_R_PSI_CY, c); }
    if (lane == 18) { fp c; fp_set(c, K_PSI2_CX); vm_set_fp(slots, VM_R_PSI2_CX, c); }
    __syncwarp();
}
// x^|z| on (X, ACC) by square-and-multiply, |z| = 0xd201000000010000 (ACC = X on entry = bit 63).  The runs of squarings between
// the set bits (1, 2, 3, 9, 32, then 16 trailing) go through the run programs CYCSQR16 / 8 / 4 / 2 / 1.
HB_DEV void vm_cycsqr_run(uint32_t* slots, int run) {
    while (run >= 16) { vm_run(VM_P_CYCSQR16, slots); run -= 16; }
    if (run & 8) vm_run(VM_P_CYCSQR8, slots);
    if (run & 4) vm_run(VM_P_CYCSQR4, slots);
    if (run & 2) vm_run(VM_P_CYCSQR2, slots);
    if (run & 1) vm_run(VM_P_CYCSQR, slots);
}
HB_DEV void vm_expz(uint32_t* slots) {
    int run = 0;
    for (int i = 62; i >= 0; i--) {
        run++;
        if ((K_Z_ABS >> i) & 1) { vm_cycsqr_run(slots, run); run = 0; vm_run(VM_P_MULX, slots); }
    }
    vm_cycsqr_run(slots, run);
}
// ---- hash-to-G2 cofactor clearing on the VM (tools/vmgen.py: G2_INIT .. G2_AFF).  T1 = [|z|] T2 (T1 == T2 on entry): 63 doublings in
// runs of 16 / 8 / 4 / 2 / 1 + an addition of T2 at the five lower set bits of |z|.
HB_DEV void vm_g2_dbl_run(uint32_t* slots, int run) {
    while (run >= 16) { vm_run(VM_P_G2DBL16, slots); run -= 16; }
    if (run & 8) vm_run(VM_P_G2DBL8, slots);
    if (run & 4) vm_run(VM_P_G2DBL4, slots);
    if (run & 2) vm_run(VM_P_G2DBL2, slots);
    if (run & 1) vm_run(VM_P_G2DBL, slots);
}
HB_DEV void vm_g2_zmul(uint32_t* slots) {
    int run = 0;
    for (int i = 62; i >= 0; i--) {
        run++;
        if ((K_Z_ABS >> i) & 1) { vm_g2_dbl_run(slots, run); run = 0; vm_run(VM_P_G2_ADD, slots); }
    }
    vm_g2_dbl_run(slots, run);
}
// in: (Q2X, Q2Y) = the affine Shallue-van de Woestijne point.  out: (HX, HY) = h(P) affine (Budroni-Pintore, the bytes of
// curve.cuh g2_clear_cofactor).  false: the generic group law degenerated somewhere (Z = 0: equal / opposite operands or an
// identity -- never for an honest map output); the caller redoes the message with the complete lane-pair code.
HB_NOINLINE bool vm_hash_cofactor(uint32_t* slots) {
    const int lane = threadIdx.x & 31;
    vm_run(VM_P_G2_INIT, slots); vm_g2_zmul(slots);
    vm_run(VM_P_HC_MID, slots); vm_g2_zmul(slots);
    vm_run(VM_P_HC_FIN, slots); vm_run(VM_P_G2_NORM, slots);
    bool zero = false;
    if (lane == 0) {
        fp n, ni; vm_ld(n.l, slots, VM_R_NORM, 0);
        zero = fp_is_zero(n);
        fp_inv_gcd(ni, n);
        vm_set_fp(slots, VM_R_NINV, ni);
    }
    if (__ballot_sync(0xffffffffu, zero) & 1u) return false;
    vm_run(VM_P_G2_AFF, slots);
    return true;
}
// F^(3 (p^12 - 1) / r) == 1 ?  Verdict to every lane.  Easy part around ONE Fp inversion (binary GCD on one lane), hard part = five
// x^|z| chains.
HB_NOINLINE bool vm_final_exp_is_one(uint32_t* slots) {
    const int lane = threadIdx.x & 31;
    vm_run(VM_P_FE_INV_A, slots);
    if (lane == 0) {
        fp n, ni; vm_ld(n.l, slots, VM_R_NORM, 0);
        fp_inv_gcd(ni, n);
        vm_set_fp(slots, VM_R_NINV, ni);
    }
    __syncwarp();
    vm_run(VM_P_FE_INV_B, slots);
    vm_expz(slots); vm_run(VM_P_GLUE1, slots);
    vm_expz(slots); vm_run(VM_P_GLUE2, slots);
    vm_expz(slots); vm_run(VM_P_GLUE3, slots);
    vm_expz(slots); vm_run(VM_P_GLUE4, slots);
    vm_expz(slots); vm_run(VM_P_GLUE5, slots); vm_run(VM_P_GLUE6, slots);
    // result in ACC0..5: lanes 0..11 compare one half each against 1
    uint32_t diff = 0;
    if (lane < 12) {
        fp v, one; vm_ld(v.l, slots, VM_R_ACC0 + (lane >> 1), lane & 1);
        fp_one(one);
#pragma unroll
        for (int j = 0; j < 12; j++) diff |= v.l[j] ^ (lane == 0 ? one.l[j] : 0u);
    }
    return __ballot_sync(0xffffffffu, diff != 0) == 0;
}
// e(P1, Q1) e(P2, Q2) == 1 ?  Inputs already in the slots P1X .. Q2Y (affine, none the identity); returns the verdict to every lane.
HB_NOINLINE bool vm_pairing_check(uint32_t* slots) {
    vm_run(VM_P_ML_INIT, slots);
    for (int i = 62; i >= 0; ) {
        if ((K_Z_ABS >> i) & 1) { vm_run(VM_P_ML_DBL, slots); vm_run(VM_P_ML_ADD, slots); i--; }
        else if (i >= 1 && !((K_Z_ABS >> (i - 1)) & 1)) { vm_run(VM_P_ML_DBL2, slots); i -= 2; }      // two doubling iterations in one program
        else { vm_run(VM_P_ML_DBL, slots); i--; }
    }
    return vm_final_exp_is_one(slots);
}

}  // namespace hb
