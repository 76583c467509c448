// tools/probe_int.cu -- integer-pipe calibration probes for the roofline denominator.
//
// Measures, on the whole chip, the issue rate of the instruction forms the field arithmetic is built from:
//   wide+add64 mad.wide.u32 with a 64-bit addend, no carry flags (ptxas: IMAD.WIDE.U32 Rd, Ra, Rb, RZ + IADD3 / IADD3.X)
//   carry      IMAD.WIDE.U32(.X) chains = mad.lo.cc / madc.hi.cc exactly as fp.cuh: lane_mad emits them (6 IMAD.WIDE per row)
//   imad32 / iadd   32-bit IMAD and IADD3 alone (the FMA-heavy and ALU pipes' instruction rates)
//   mul32x12   the shipped 12 x 32-bit Montgomery product (fp_mul_regs), register resident, dependent products
//   mul28x14   a carry-free 14 x 28-bit Montgomery product (64-bit column accumulators, plain IMAD.WIDE only)
// and cross-checks mul28x14 against mul32x12 on the same canonical operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_int tools/probe_int.cu
// Run (GPU box): tools/probe_int > gpurun_out/probe_int.json
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../harmony_b200/csrc/fp.cuh"
#include "../harmony_b200/csrc/fp_wide.cuh"
#include "../harmony_b200/csrc/hbls_constants.cuh"

using namespace hb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ------------------------------------------------------------------ instruction-form probes
template <int ILP> __global__ void k_plain(int iters, uint32_t seed, uint64_t* sink) {
    // acc_k += lo32(acc_{k+1}) * b: the multiplicand changes every iteration, so ptxas cannot hoist the product out of the loop
    // (the round-1 probe multiplied two loop invariants; ptxas turned it into IADD3 pairs and the "IMAD.WIDE peak" it reported
    // was half the ALU-pipe add rate)
    uint64_t acc[ILP];
    uint32_t a = seed ^ (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x), b = seed * 2654435761u + threadIdx.x;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = (uint64_t)(k + 1) * 0x9e3779b97f4a7c15ull + a;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((uint32_t)acc[(k + 1) % ILP]), "r"(b));
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    if (s == 0x1234567ull) sink[0] = s;
}
// K independent accumulators, each fed by lane_mad (6 lo/hi pairs = 6 IMAD.WIDE.U32(.X) under one carry chain + ripple)
template <int K> __global__ void k_carry(int iters, uint32_t seed, uint32_t* sink) {
    uint32_t acc[K][14], a[12];
    uint32_t t = seed ^ (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x);
#pragma unroll
    for (int j = 0; j < 12; j++) { t = t * 1664525u + 1013904223u; a[j] = t; }
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 14; j++) { t = t * 1664525u + 1013904223u; acc[k][j] = t; }
    uint32_t b = t | 1u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < K; k++) lane_mad(acc[k], a, b);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 14; j++) s ^= acc[k][j];
    if (s == 0x12345u) sink[0] = s;
}
// plain-form twin of k_carry: same number of MACs per iteration (6 per accumulator set), no carry flags
template <int K> __global__ void k_plain_rows(int iters, uint32_t seed, uint64_t* sink) {
    uint64_t acc[K][6]; uint32_t a[6];
    uint32_t t = seed ^ (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x);
#pragma unroll
    for (int j = 0; j < 6; j++) { t = t * 1664525u + 1013904223u; a[j] = t; }
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 6; j++) { t = t * 1664525u + 1013904223u; acc[k][j] = t; }
    uint32_t b = t | 1u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
            for (int j = 0; j < 6; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k][j]) : "r"(a[j]), "r"(b));
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 6; j++) s ^= acc[k][j];
    if (s == 0x12345ull) sink[0] = s;
}
// plain IMAD.WIDE interleaved 1:1 with IADD3 (ALU pipe): do the two pipes issue side by side?
template <int ILP> __global__ void k_plain_plus_alu(int iters, uint32_t seed, uint64_t* sink) {
    uint64_t acc[ILP]; uint32_t x[ILP];
    uint32_t a = seed ^ (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x), b = seed * 2654435761u + threadIdx.x;
#pragma unroll
    for (int k = 0; k < ILP; k++) { acc[k] = (uint64_t)(k + 1) * 0x9e3779b97f4a7c15ull + a; x[k] = a + k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(b));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k] + x[k];
    if (s == 0x1234567ull) sink[0] = s;
}

// ALU pipe alone: ILP independent 32-bit adds (IADD3)
template <int ILP> __global__ void k_alu(int iters, uint32_t seed, uint64_t* sink) {
    uint32_t x[ILP];
    uint32_t a = seed ^ (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x), b = seed * 2654435761u + threadIdx.x;
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = a + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= x[k];
    if (s == 0x1234567u) sink[0] = s;
}
// 32-bit IMAD (mad.lo) alone
template <int ILP> __global__ void k_imad32(int iters, uint32_t seed, uint64_t* sink) {
    uint32_t x[ILP];
    uint32_t a = seed ^ (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x), b = seed * 2654435761u + threadIdx.x;
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = a + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[k]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= x[k];
    if (s == 0x1234567u) sink[0] = s;
}

// ------------------------------------------------------------------ 14 x 28-bit carry-free Montgomery product (R = 2^392)
#define M28 0x0fffffffu
__device__ __forceinline__ uint32_t p28(int i) {
    switch (i) {
    case 0: return 0xfffaaabu; case 1: return 0xfefffffu; case 2: return 0x3ffffb9u; case 3: return 0xfffeb15u;
    case 4: return 0x6241eabu; case 5: return 0xa0f6b0fu; case 6: return 0xf6730d2u; case 7: return 0xf38512bu;
    case 8: return 0x4774b84u; case 9: return 0x4bacd76u; case 10: return 0xba7b643u; case 11: return 0xe69a4b1u;
    case 12: return 0x1ea397fu; default: return 0x1a011u;
    }
}
#define N0_28 0xffcfffdu
__device__ __constant__ const uint32_t K28_R2[14] = {0x10370edu, 0x6d1c345u, 0xe243d62u, 0xec45c53u, 0x3b1d65au, 0x93317du, 0xb4f36a0u, 0x5d74088u, 0xc10ea72u, 0x865d118u, 0x7320a75u, 0xfd5cd50u, 0xcc8a759u, 0xc8d4u};

// r = a b / 2^392 mod p; a, b: limbs < 2^29; r: limbs < 2^28, value < 2p when a b < 2^392 p
// V = 1: asm volatile (ptxas keeps one IMAD.WIDE with 64-bit accumulate per MAC); V = 0: plain asm (ptxas is free to split the
// accumulation into IMAD.WIDE + 3-input IADD3 trees, trading FMA-pipe dependencies for ALU-pipe work)
template <int V> __device__ __forceinline__ void madw(uint64_t& acc, uint32_t a, uint32_t b) {
    if (V) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
    else asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
}
template <int V> __device__ __forceinline__ void fp28_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t t[14];
#pragma unroll
    for (int j = 0; j < 14; j++) t[j] = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
#pragma unroll
        for (int j = 0; j < 14; j++) madw<V>(t[j], a[j], b[i]);
        const uint32_t m = ((uint32_t)t[0] * N0_28) & M28;
#pragma unroll
        for (int j = 0; j < 14; j++) madw<V>(t[j], m, p28(j));
        const uint64_t c = t[0] >> 28;
        t[0] = t[1] + c;
#pragma unroll
        for (int j = 1; j < 13; j++) t[j] = t[j + 1];
        t[13] = 0;
    }
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 14; j++) { c += t[j]; r[j] = (uint32_t)c & M28; c >>= 28; }
}
// canonical 12 x 32 words <-> 14 x 28 limbs
__device__ __forceinline__ void to28(uint32_t* l, const uint32_t* w) {
#pragma unroll
    for (int i = 0; i < 14; i++) {
        const int bit = 28 * i, q = bit >> 5, s = bit & 31;
        uint64_t v = w[q];
        if (q + 1 < 12) v |= (uint64_t)w[q + 1] << 32;
        l[i] = (uint32_t)(v >> s) & M28;
    }
}
__device__ __forceinline__ void from28(uint32_t* w, const uint32_t* l) {
#pragma unroll
    for (int q = 0; q < 12; q++) w[q] = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
        const int bit = 28 * i, q = bit >> 5, s = bit & 31;
        const uint64_t v = (uint64_t)l[i] << s;
        w[q] |= (uint32_t)v;
        if (q + 1 < 12) w[q + 1] |= (uint32_t)(v >> 32);
    }
}
// value < 2p with normalized limbs -> [0, p)
__device__ __forceinline__ void fp28_canon(uint32_t* r) {
    uint32_t s[14]; uint32_t borrow = 0;
#pragma unroll
    for (int j = 0; j < 14; j++) { const uint32_t d = r[j] - p28(j) - borrow; borrow = d >> 31; s[j] = d & M28; }
#pragma unroll
    for (int j = 0; j < 14; j++) r[j] = borrow ? r[j] : s[j];
}

// x <- x*y, y <- y*x, iters times; operands canonical 48-byte little-endian, result canonical
__global__ void k_chain32(int iters, const uint32_t* in, uint32_t* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x[12], y[12], r2[12], one[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { x[j] = in[24 * i + j]; y[j] = in[24 * i + 12 + j]; r2[j] = K_R2[j]; one[j] = j == 0; }
    fp_mul_regs(x, x, r2); fp_mul_regs(y, y, r2);
#pragma unroll 1
    for (int k = 0; k < iters; k++) { fp_mul_regs(x, x, y); fp_mul_regs(y, y, x); }
    fp_mul_regs(x, x, one);
#pragma unroll
    for (int j = 0; j < 12; j++) out[12 * i + j] = x[j];
}
// the shipped limbs with ONE Karatsuba level in the product (fp_wide.cuh mul_wide_k) + a separate reduction: 108 + 156 IMAD instead of 300
__device__ __forceinline__ void fp_mul_k(uint32_t* r, const uint32_t* a, const uint32_t* b) { uint32_t T[24]; mul_wide_k(T, a, b); redc_wide(r, T); }
__device__ __forceinline__ void fp_mul_w(uint32_t* r, const uint32_t* a, const uint32_t* b) { uint32_t T[24]; mul_wide(T, a, b); redc_wide(r, T); }
template <int K> __global__ void k_chain32w(int iters, const uint32_t* in, uint32_t* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x[12], y[12], r2[12], one[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { x[j] = in[24 * i + j]; y[j] = in[24 * i + 12 + j]; r2[j] = K_R2[j]; one[j] = j == 0; }
    fp_mul_regs(x, x, r2); fp_mul_regs(y, y, r2);
#pragma unroll 1
    for (int k = 0; k < iters; k++) { if (K) { fp_mul_k(x, x, y); fp_mul_k(y, y, x); } else { fp_mul_w(x, x, y); fp_mul_w(y, y, x); } }
    fp_mul_regs(x, x, one);
#pragma unroll
    for (int j = 0; j < 12; j++) out[12 * i + j] = x[j];
}
template <int V> __global__ void k_chain28(int iters, const uint32_t* in, uint32_t* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w[12], x[14], y[14], r2[14], one[14];
#pragma unroll
    for (int j = 0; j < 12; j++) w[j] = in[24 * i + j];
    to28(x, w);
#pragma unroll
    for (int j = 0; j < 12; j++) w[j] = in[24 * i + 12 + j];
    to28(y, w);
#pragma unroll
    for (int j = 0; j < 14; j++) { r2[j] = K28_R2[j]; one[j] = j == 0; }
    fp28_mul<V>(x, x, r2); fp28_mul<V>(y, y, r2);
#pragma unroll 1
    for (int k = 0; k < iters; k++) { fp28_mul<V>(x, x, y); fp28_mul<V>(y, y, x); }
    fp28_mul<V>(x, x, one);
    fp28_canon(x);
    from28(w, x);
#pragma unroll
    for (int j = 0; j < 12; j++) out[12 * i + j] = w[j];
}

static float time_ms(cudaStream_t s, void (*launch)(void*), void* ctx) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(ctx); CK(cudaStreamSynchronize(s));
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0, s)); launch(ctx); CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return best;
}
struct L { int blocks, threads, iters; void* p0; void* p1; int which; };
static void do_launch(void* c) {
    L* l = (L*)c;
    switch (l->which) {
    case 0: k_plain<8><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint64_t*)l->p0); break;
    case 1: k_carry<1><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint32_t*)l->p0); break;
    case 2: k_carry<2><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint32_t*)l->p0); break;
    case 3: k_carry<4><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint32_t*)l->p0); break;
    case 4: k_plain_rows<4><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint64_t*)l->p0); break;
    case 5: k_plain_plus_alu<8><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint64_t*)l->p0); break;
    case 9: k_alu<8><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint64_t*)l->p0); break;
    case 10: k_imad32<8><<<l->blocks, l->threads>>>(l->iters, 12345u, (uint64_t*)l->p0); break;
    case 6: k_chain32<<<l->blocks, l->threads>>>(l->iters, (const uint32_t*)l->p0, (uint32_t*)l->p1); break;
    case 7: k_chain28<0><<<l->blocks, l->threads>>>(l->iters, (const uint32_t*)l->p0, (uint32_t*)l->p1); break;
    case 8: k_chain28<1><<<l->blocks, l->threads>>>(l->iters, (const uint32_t*)l->p0, (uint32_t*)l->p1); break;
    case 11: k_chain32w<0><<<l->blocks, l->threads>>>(l->iters, (const uint32_t*)l->p0, (uint32_t*)l->p1); break;
    case 12: k_chain32w<1><<<l->blocks, l->threads>>>(l->iters, (const uint32_t*)l->p0, (uint32_t*)l->p1); break;
    }
}

int main(int argc, char** argv) {
    int only = argc > 1 ? atoi(argv[1]) : -1;      // ncu: run one variant only
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    void* sink; CK(cudaMalloc(&sink, 4096));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
    // wide_plus_add64: ptxas lowers a plain (carry-free) mad.wide.u32 with a 64-bit addend to IMAD.WIDE(RZ) + IADD3 / IADD3.X
    const char* names[11] = {"wide_plus_add64_ilp8", "carry_k1", "carry_k2", "carry_k4", "", "wide_plus_add64_plus_iadd_ilp8", "", "", "", "iadd_only_ilp8", "imad32_ilp8"};
    const double mac_per_iter[11] = {8, 6, 12, 24, 24, 8, 0, 0, 0, 8, 8};      // lane_mad = 6 wide MACs (each mad.lo.cc/madc.hi.cc pair is ONE IMAD.WIDE)
    for (int v = 0; v < 11; v++) {
        if ((v >= 6 && v <= 8) || v == 4) continue;
        if (only >= 0 && only != v) continue;
        for (int tpsm = 256; tpsm <= 1024; tpsm *= 2) {
            L l{sms * (tpsm / 256), 256, 4096, sink, nullptr, v};
            float ms = time_ms(0, do_launch, &l);
            double macs = (double)l.blocks * l.threads * l.iters * mac_per_iter[v];
            printf(", \"%s_t%d_TMACps\": %.3f", names[v], tpsm, macs / (ms * 1e-3) / 1e12);
        }
    }
    // products: operands < p, pseudo-random
    const int maxthreads = sms * 1024;
    std::vector<uint32_t> h(24 * (size_t)maxthreads);
    uint32_t st = 20240923u;
    for (size_t i = 0; i < h.size(); i++) { st = st * 1664525u + 1013904223u; h[i] = st ^ (st >> 15); if (i % 12 == 11) h[i] &= 0x0fffffffu; }
    uint32_t *din, *d32, *d28, *d28v;
    CK(cudaMalloc(&din, h.size() * 4)); CK(cudaMalloc(&d32, 12 * (size_t)maxthreads * 4)); CK(cudaMalloc(&d28, 12 * (size_t)maxthreads * 4)); CK(cudaMalloc(&d28v, 12 * (size_t)maxthreads * 4));
    CK(cudaMemcpy(din, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const int iters = 512;
    uint32_t *dw, *dk; CK(cudaMalloc(&dw, 12 * (size_t)maxthreads * 4)); CK(cudaMalloc(&dk, 12 * (size_t)maxthreads * 4));
    for (int v : {6, 7, 8, 11, 12}) {
        if (only >= 0 && only != v) continue;
        for (int tpsm = 128; tpsm <= 1024; tpsm *= 2) {
            L l{sms * (tpsm / 128), 128, iters, din, v == 6 ? d32 : (v == 7 ? d28 : (v == 8 ? d28v : (v == 11 ? dw : dk))), v};
            float ms = time_ms(0, do_launch, &l);
            double muls = (double)l.blocks * l.threads * (2.0 * iters + 3);
            printf(", \"%s_t%d_Gmulps\": %.3f", v == 6 ? "mul32x12" : (v == 7 ? "mul28x14" : (v == 8 ? "mul28x14_volatile" : (v == 11 ? "mul32x12_wide_plus_redc" : "mul32x12_karatsuba_plus_redc"))), tpsm, muls / (ms * 1e-3) / 1e9);
        }
    }
    if (only < 0) {
        std::vector<uint32_t> a(12 * (size_t)maxthreads), b(12 * (size_t)maxthreads);
        CK(cudaMemcpy(a.data(), d32, a.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), d28, b.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0, nz = 0, badv = 0;
        for (size_t i = 0; i < a.size(); i++) { bad += a[i] != b[i]; nz += a[i] != 0; }
        CK(cudaMemcpy(b.data(), d28v, b.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < a.size(); i++) badv += a[i] != b[i];
        size_t badw = 0, badk = 0;
        CK(cudaMemcpy(b.data(), dw, b.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < a.size(); i++) badw += a[i] != b[i];
        CK(cudaMemcpy(b.data(), dk, b.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < a.size(); i++) badk += a[i] != b[i];
        printf(", \"mul28_vs_mul32_mismatching_words\": %zu, \"mul28v_vs_mul32_mismatching_words\": %zu, \"wide_vs_mul32_mismatching_words\": %zu, \"karatsuba_vs_mul32_mismatching_words\": %zu, \"nonzero_words\": %zu", bad, badv, badw, badk, nz);
    }
    printf("}\n");
    return 0;
}
