// tools/lat_probe.cu -- single-warp latency of the building blocks of the latency path (clock64 around dependent chains).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/lat_probe tools/lat_probe.cu
// Run (GPU box): tools/lat_probe > gpurun_out/lat_probe.txt
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../harmony_b200/csrc/vm.cuh"

using namespace hb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

enum { T_FP_MUL, T_FP_SQR, T_FP_POW, T_FP_INVGCD, T_FP_LEG, T_FP2_SQRT, T_F2H_MUL, T_F2H_SQR, T_G2H_DBL, T_G2H_ZMUL, T_SUBGROUP, T_COFACTOR, T_SWMAP,
       T_VM_MUL, T_VM_SQR, T_VM_LIN4, T_VM_LIN8, T_VM_DBL, T_VM_CYC16, T_N };
static const char* NAMES[T_N] = {"fp_mul x64 (1 thread)", "fp_sqr x64", "fp_pow (p-3)/4", "fp_inv_gcd", "fp_legendre", "fp2_sqrt", "fp2h mul x64 (lane pair)",
    "fp2h sqr x64", "g2 dbl x16 (lane pair)", "g2 [|z|]P (lane pair)", "g2_in_subgroup (lane pair)", "g2_clear_cofactor (lane pair)", "sw_map_g2<GCD>",
    "vm_mul x32 (warp)", "vm_sqr x32", "vm_lin 4 atoms x32", "vm_lin 8 atoms x32", "VM ML_DBL program", "VM CYCSQR16 program"};
static const int REPS[T_N] = {64, 64, 1, 1, 1, 1, 64, 64, 16, 1, 1, 1, 1, 32, 32, 32, 32, 1, 1};

__global__ void __launch_bounds__(32) k_probe(int active, long long* out, uint32_t seed, uint32_t* sink) {
    __shared__ uint32_t slots[VM_SMEM_WORDS];
    const int lane = threadIdx.x;
    for (int i = lane; i < VM_SMEM_WORDS; i += 32) slots[i] = (seed * 2654435761u + i * 40503u) & 0x0fffffffu;      // < p per limb pattern
    __syncwarp();
    vm_load_consts(slots);
    fp a, b; for (int j = 0; j < 12; j++) { a.l[j] = slots[25 * 20 + j]; b.l[j] = slots[25 * 21 + j]; }
    a.l[11] &= 0x0fffffffu; b.l[11] &= 0x0fffffffu;
    uint32_t acc = 0;
    long long t0, t1;
    const bool on = lane < active;
    // ---- single-thread Fp chains
    if (on) {
        fp x = a;
        t0 = clock64(); for (int i = 0; i < 64; i++) fp_mul(x, x, b); t1 = clock64(); if (lane == 0) out[T_FP_MUL] = t1 - t0; acc ^= x.l[0];
        t0 = clock64(); for (int i = 0; i < 64; i++) fp_sqr(x, x); t1 = clock64(); if (lane == 0) out[T_FP_SQR] = t1 - t0; acc ^= x.l[0];
        fp y;
        t0 = clock64(); fp_pow(y, x, K_P_MINUS_3_DIV_4); t1 = clock64(); if (lane == 0) out[T_FP_POW] = t1 - t0; acc ^= y.l[0];
        t0 = clock64(); fp_inv_gcd(y, x); t1 = clock64(); if (lane == 0) out[T_FP_INVGCD] = t1 - t0; acc ^= y.l[0];
        t0 = clock64(); int lg = fp_legendre(x); t1 = clock64(); if (lane == 0) out[T_FP_LEG] = t1 - t0; acc ^= (uint32_t)lg;
        fp2 u, r; u.a = x; u.b = b; fp2_sqr(u, u);
        t0 = clock64(); bool okr = fp2_sqrt(r, u); t1 = clock64(); if (lane == 0) out[T_FP2_SQRT] = t1 - t0; acc ^= r.a.l[0] ^ (uint32_t)okr;
        fp2 tt; tt.a = x; fp_zero(tt.b); g2 pt;
        t0 = clock64(); bool okm = sw_map_g2<true>(pt, tt); t1 = clock64(); if (lane == 0) out[T_SWMAP] = t1 - t0; acc ^= pt.x.a.l[0] ^ (uint32_t)okm;
    }
    __syncwarp();
    // ---- lane-pair chains (lanes 0,1 at least)
    if (lane < (active < 2 ? 2 : active)) {
        fp2h x, y; x.c = a; y.c = b;
        t0 = clock64(); for (int i = 0; i < 64; i++) fp2_mul(x, x, y); t1 = clock64(); if (lane == 0) out[T_F2H_MUL] = t1 - t0; acc ^= x.c.l[0];
        t0 = clock64(); for (int i = 0; i < 64; i++) fp2_sqr(x, x); t1 = clock64(); if (lane == 0) out[T_F2H_SQR] = t1 - t0; acc ^= x.c.l[0];
        // a real curve point: the generator image is not needed -- any (x, y, 1) exercises the same formulas
        jac<fp2h> P, Q; P.x = x; P.y = y; fp2_one(P.z);
        t0 = clock64(); for (int i = 0; i < 16; i++) pt_dbl(P, P); t1 = clock64(); if (lane == 0) out[T_G2H_DBL] = t1 - t0; acc ^= P.x.c.l[0];
        t0 = clock64(); pt_mul_zabs(Q, P); t1 = clock64(); if (lane == 0) out[T_G2H_ZMUL] = t1 - t0; acc ^= Q.x.c.l[0];
        t0 = clock64(); bool sg = g2_in_subgroup(P); t1 = clock64(); if (lane == 0) out[T_SUBGROUP] = t1 - t0; acc ^= (uint32_t)sg;
        t0 = clock64(); g2_clear_cofactor(Q, P); t1 = clock64(); if (lane == 0) out[T_COFACTOR] = t1 - t0; acc ^= Q.x.c.l[0];
    }
    __syncwarp();
    // ---- VM primitives, full warp (16 lane pairs, each its own slots)
    {
        const int pair = lane >> 1, im = lane & 1;
        t0 = clock64(); for (int i = 0; i < 32; i++) { vm_mul(slots, 80 + pair, 24 + pair, 40 + ((pair + i) & 15), im); __syncwarp(); } t1 = clock64(); if (lane == 0) out[T_VM_MUL] = t1 - t0;
        t0 = clock64(); for (int i = 0; i < 32; i++) { vm_sqr(slots, 80 + pair, 24 + ((pair + i) & 15), im); __syncwarp(); } t1 = clock64(); if (lane == 0) out[T_VM_SQR] = t1 - t0;
        // atoms: slot, half, sign, multiplier
        auto atom = [](int s, int half, int neg, int sh) { return (uint32_t)(s | (half << 8) | (neg << 9) | (sh << 10)); };
        uint4 row;
        row.x = atom(24 + pair, 0, 0, 0) | (atom(40 + pair, 1, 1, 1) << 16); row.y = atom(25 + pair, 0, 0, 2) | (atom(41 + pair, 1, 1, 3) << 16);
        row.z = atom(26 + pair, 1, 1, 0) | (atom(42 + pair, 0, 0, 1) << 16); row.w = atom(27 + pair, 1, 0, 2) | (atom(43 + pair, 0, 1, 0) << 16);
        t0 = clock64(); for (int i = 0; i < 32; i++) { vm_lin(slots, 80 + pair, row, 4, im); __syncwarp(); } t1 = clock64(); if (lane == 0) out[T_VM_LIN4] = t1 - t0;
        t0 = clock64(); for (int i = 0; i < 32; i++) { vm_lin(slots, 80 + pair, row, 8, im); __syncwarp(); } t1 = clock64(); if (lane == 0) out[T_VM_LIN8] = t1 - t0;
        t0 = clock64(); vm_run(VM_P_ML_DBL, slots); t1 = clock64(); if (lane == 0) out[T_VM_DBL] = t1 - t0;
        t0 = clock64(); vm_run(VM_P_CYCSQR16, slots); t1 = clock64(); if (lane == 0) out[T_VM_CYC16] = t1 - t0;
        acc ^= slots[25 * 80 + lane];
    }
    if (acc == 0x12345678u) sink[lane] = acc;
}

int main() {
    long long* d_out; uint32_t* d_sink;
    CK(cudaMalloc(&d_out, T_N * sizeof(long long))); CK(cudaMalloc(&d_sink, 128));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    for (int active : {1, 2, 32}) {
        CK(cudaMemset(d_out, 0, T_N * sizeof(long long)));
        k_probe<<<1, 32>>>(active, d_out, 12345u, d_sink);       // warm-up (instruction cache)
        k_probe<<<1, 32>>>(active, d_out, 12345u, d_sink);
        CK(cudaDeviceSynchronize());
        long long h[T_N]; CK(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost));
        printf("--- lanes active in the single-thread / lane-pair sections: %d   (clock %d kHz nominal)\n", active, clk_khz);
        for (int t = 0; t < T_N; t++)
            printf("%-36s %10lld cycles  = %9.1f per op   (%8.2f us total at 1.965 GHz)\n", NAMES[t], h[t], (double)h[t] / REPS[t], h[t] / 1965.0);
    }
    return 0;
}
