// harmony_b200/csrc/vm.cuh -- warp-cooperative ("latency mode") pairing: ONE warp = 16 lane pairs verifies one round.
//
// The reference verifies one aggregate per block on the consensus path (consensus/validator.go:219-236, internal/chain/engine.go:
// 619-642): there the figure of merit is the latency of a single check, and a thread-per-round (or lane-pair-per-round) pairing is
// ~10^4 dependent field products long.  Here the Miller loop and the final exponentiation run as straight-line STEP PROGRAMS over
// Fp2 values held in shared memory (vm_programs.cuh, generated and CPU-verified by tools/vmgen.py): in every step each of the 16
// lane pairs of the warp executes one Fp2 operation -- a product, a squaring, or a small-integer linear combination -- so up to 16
// independent Fp2 products are in flight.  The whole working set (232 slots x 100 B) stays in shared memory: no local-memory stack.
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
    mul_wide2(T, ar, y1, ai, y2);                      // < 2 p^2 < p R
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
    mul_wide(T, A, B);                 // < 4 p^2 < p R
    redc_wide(rr, T);
    vm_st(slots, dst, im, rr);
}
// acc += c * v for a small signed c (|c| <= 7), all values canonical
HB_DEV void vm_axpy(fp& acc, int c, const fp& v) {
    if (c == 0) return;
    const int m = c < 0 ? -c : c;
    fp t = v, sum; bool have = false;
    for (int bit = 0; bit < 3; bit++) {
        if (m & (1 << bit)) { if (have) fp_add(sum, sum, t); else { sum = t; have = true; } }
        if ((m >> (bit + 1)) == 0) break;
        fp_dbl(t, t);
    }
    if (c > 0) fp_add(acc, acc, sum); else fp_sub(acc, acc, sum);
}
// dst = sum_k M_k src_k, M_k = 2 x 2 integer matrix on (re, im): this lane evaluates its own row
HB_NOINLINE void vm_lin(uint32_t* slots, const uint4 ins, int im) {
    const int dst = ins.x & 0xff, nt = (ins.x >> 8) & 0xff;
    uint32_t tw[4];
    tw[0] = (ins.x >> 16) | ((ins.y & 0xffu) << 16); tw[1] = ins.y >> 8; tw[2] = ins.z & 0xffffffu; tw[3] = (ins.z >> 24) | ((ins.w & 0xffffu) << 8);
    fp acc; fp_zero(acc);
    for (int k = 0; k < nt; k++) {
        const int s = tw[k] & 0xff;
        const uint32_t e = tw[k] >> 8;                                  // four 4-bit two's-complement entries m00 m01 m10 m11
        const int sh = im ? 8 : 0;
        const int c_re = ((int)((e >> sh) & 0xf) ^ 8) - 8, c_im = ((int)((e >> (sh + 4)) & 0xf) ^ 8) - 8;
        fp v;
        if (c_re) { vm_ld(v.l, slots, s, 0); vm_axpy(acc, c_re, v); }
        if (c_im) { vm_ld(v.l, slots, s, 1); vm_axpy(acc, c_im, v); }
    }
    vm_st(slots, dst, im, acc.l);
}
// run one step program; every lane of the warp must call it (steps end in __syncwarp)
HB_NOINLINE void vm_run(int prog, uint32_t* slots) {
    const int lane = threadIdx.x & 31, pair = lane >> 1, im = lane & 1;
    const int first = VM_PROG_FIRST[prog], n = VM_PROG_STEPS[prog];
    for (int st = first; st < first + n; st++) {
        const int cls = VM_STEP_CLASS[st];
        const uint4 ins = VM_INS[st * 16 + pair];
        const int dst = ins.x & 0xff;
        if (dst != 0xff) {
            if (cls == VM_OP_LIN) vm_lin(slots, ins, im);
            else if (cls == VM_OP_SQR) vm_sqr(slots, dst, (ins.x >> 8) & 0xff, im);
            else vm_mul(slots, dst, (ins.x >> 8) & 0xff, (ins.x >> 16) & 0xff, im);
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
// constants of the programs (ONE, 3 b', 1/2, Frobenius coefficients): lanes share the 15 slots
HB_DEV void vm_load_consts(uint32_t* slots) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) { fp o; fp_one(o); vm_set_fp(slots, VM_R_ONE, o); }
    if (lane == 1) { fp2 b; fp2_const(b, K_B2_3); vm_set_fp2(slots, VM_R_TWIST3B, b); }
    if (lane == 2) { fp h; fp_set(h, K_INV2); vm_set_fp(slots, VM_R_INV2, h); }
    if (lane >= 3 && lane < 9) { fp2 g; fp2_const(g, K_FROB1[lane - 3]); vm_set_fp2(slots, VM_R_FROB1_0 + (lane - 3), g); }
    if (lane >= 9 && lane < 15) { fp g; fp_set(g, K_FROB2[lane - 9]); vm_set_fp(slots, VM_R_FROB2_0 + (lane - 9), g); }
    __syncwarp();
}
// x^|z| on (X, ACC) by square-and-multiply, |z| = 0xd201000000010000 (ACC = X on entry = bit 63)
HB_DEV void vm_expz(uint32_t* slots) {
    for (int i = 62; i >= 0; i--) {
        vm_run(VM_P_CYCSQR, slots);
        if ((K_Z_ABS >> i) & 1) vm_run(VM_P_MULX, slots);
    }
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
    vm_expz(slots); vm_run(VM_P_GLUE5, slots);
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
    for (int i = 62; i >= 0; i--) {
        vm_run(VM_P_ML_DBL, slots);
        if ((K_Z_ABS >> i) & 1) vm_run(VM_P_ML_ADD, slots);
    }
    return vm_final_exp_is_one(slots);
}

}  // namespace hb
