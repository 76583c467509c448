// harmony_b200/csrc/hbls.cu -- libhbls.so: C ABI (include/hbls.h) over the sm_100a kernels.
// Host side = plumbing only (buffers, one stream, launches); every group/field operation runs on the GPU.
// No CPU fallback exists: without a usable CUDA device blsInit fails and every entry point returns HBLS_ERR_CUDA.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "../../include/hbls.h"
#include "kernels.cuh"

using namespace hb;

namespace {

struct Ctx {
    bool ready = false;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    uint8_t* scratch = nullptr; size_t scratch_cap = 0;     // device bump arena
    std::atomic<uint64_t> launches{0};
    bool stage_timing = false; bool stage_valid = false;
    int batch_mode = 1;                                     // 1: random-linear-combination groups + exact fallback, 0: exact per round
    uint64_t rlc_seed[2] = {0, 0}; uint64_t rlc_calls = 0;
    cudaEvent_t ev[8] = {};
};
Ctx g;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "[hbls] CUDA error %s at %s:%d: %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); return HBLS_ERR_CUDA; } } while (0)

int ensure_init() {
    if (g.ready) return 0;
    fprintf(stderr, "[hbls] not initialised: call blsInit / hbls_init_device first (no CPU fallback)\n");
    return HBLS_ERR_CUDA;
}

struct Arena {
    uint8_t* base; size_t off = 0, cap;
    template <class T> T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + off); off += n * sizeof(T);
        return p;
    }
};
int reserve(size_t bytes) {
    bytes += 4096;
    if (bytes <= g.scratch_cap) return 0;
    if (g.scratch) CK(cudaFree(g.scratch));
    g.scratch = nullptr; g.scratch_cap = 0;
    size_t cap = bytes + bytes / 4;
    CK(cudaMalloc(&g.scratch, cap));
    g.scratch_cap = cap;
    return 0;
}
unsigned split_blocks(size_t nthreads);
unsigned light_blocks(size_t n);
inline unsigned blocks_for(size_t n, unsigned tpb) { return (unsigned)((n + tpb - 1) / tpb); }
#define LAUNCH(kern, grid, block, strm, ...) do { kern<<<(grid), (block), 0, (strm)>>>(__VA_ARGS__); g.launches++; } while (0)

constexpr unsigned TPB = 64;      // heavy kernels: 64-thread CTAs
// persistent launch geometry for the grid-stride kernels: at most `tpsm` resident threads per SM (HBLS_TPSM, default 384: measured 256: 267 ms, 384: 263 ms, 512: 280 ms per 303 104 rounds)
// one large CTA per SM for the three big thread-per-round kernels when the library is built with HB_LOCKSTEP_T
#if HB_LOCKSTEP_T
static unsigned big_tpb() { static unsigned v = [] { const char* e = getenv("HBLS_TCTA"); return e ? (unsigned)atoi(e) : 384u; }(); return v; }
static unsigned big_blocks(size_t n) { size_t need = (n + big_tpb() - 1) / big_tpb(); return (unsigned)(need < (size_t)g.sm_count ? need : (size_t)g.sm_count); }
#else
static unsigned big_tpb() { return TPB; }
unsigned heavy_blocks(size_t n);
static unsigned big_blocks(size_t n) { return heavy_blocks(n); }
#endif
unsigned heavy_blocks(size_t n) {
    static int tpsm = [] { const char* e = getenv("HBLS_TPSM"); int v = e ? atoi(e) : 384; return v < 64 ? 64 : v; }();
    size_t cap = (size_t)g.sm_count * (size_t)(tpsm / TPB);
    size_t need = (n + TPB - 1) / TPB;
    return (unsigned)(need < cap ? need : cap);
}

// small-state kernels (decode, hash): working set fits L1/L2 at any occupancy -> let the register count decide
unsigned light_blocks(size_t n) {
    static int tpsm = [] { const char* e = getenv("HBLS_TPSM_LIGHT"); int v = e ? atoi(e) : 1024; return v < 64 ? 64 : v; }();
    size_t cap = (size_t)g.sm_count * (size_t)(tpsm / TPB);
    size_t need = (n + TPB - 1) / TPB;
    return (unsigned)(need < cap ? need : cap);
}
// lane-pair kernels: resident threads per SM from HBLS_TPSM_SPLIT (default 512)
unsigned split_blocks(size_t nthreads) {
    static int tpsm = [] { const char* e = getenv("HBLS_TPSM_SPLIT"); int v = e ? atoi(e) : 512; return v < 64 ? 64 : v; }();
    size_t cap = (size_t)g.sm_count * (size_t)(tpsm / HB_TPB_SPLIT);
    size_t need = (nthreads + HB_TPB_SPLIT - 1) / HB_TPB_SPLIT;
    return (unsigned)(need < cap ? need : cap);
}
// ------------------------------------------------------------------ one verification pass over device-resident inputs.
// pk_neg: affine -apk (or -pk) per round; sig/hm decoded inside.  arena must hold verify_scratch_bytes(B).
size_t verify_scratch_bytes(size_t B) {
#if HB_FALLBACK_LIST
    return B * (sizeof(g2a) * 2 + sizeof(g1a) * 2 + sizeof(g1) + sizeof(g2) + sizeof(fp12) * 2 + 16 + 4) + (B / HB_RLC_G + 1) * (sizeof(g2a) + 8) + 33 * 256;
#endif
    return B * (sizeof(g2a) * 2 + sizeof(g1a) * 2 + sizeof(g1) + sizeof(g2) + sizeof(fp12) * 2 + 16) + (B / HB_RLC_G + 1) * (sizeof(g2a) + 8) + 32 * 256;
}
struct VerifyBufs { g2a* sig; g2a* hm; g1a* pkneg; g1* apk; fp12* f; uint8_t* ok_sig; uint8_t* ok_hm; uint8_t* ok_pk;
                    g1a* pk_scaled; g2* S; uint8_t* bad; g2a* Sg; uint8_t* group_ok; int* any_fail;
#if HB_FALLBACK_LIST
                    uint32_t* fail_list; unsigned* fail_count;
#endif
};
VerifyBufs carve_verify(Arena& ar, size_t B) {
    VerifyBufs v;
    v.sig = ar.take<g2a>(B); v.hm = ar.take<g2a>(B); v.pkneg = ar.take<g1a>(B); v.apk = ar.take<g1>(B);
    v.f = ar.take<fp12>(2 * B); v.ok_sig = ar.take<uint8_t>(B); v.ok_hm = ar.take<uint8_t>(B); v.ok_pk = ar.take<uint8_t>(B);
    v.pk_scaled = ar.take<g1a>(B); v.S = ar.take<g2>(B); v.bad = ar.take<uint8_t>(B);
    v.Sg = ar.take<g2a>(B / HB_RLC_G + 1); v.group_ok = ar.take<uint8_t>(B / HB_RLC_G + 1); v.any_fail = ar.take<int>(1);
#if HB_FALLBACK_LIST
    v.fail_list = ar.take<uint32_t>(B); v.fail_count = ar.take<unsigned>(1);
#endif
    return v;
}
// batched (random-linear-combination) form applies: default mode, lane-pair kernels, Jacobian apk at hand, batch large enough
static bool rlc_applies(size_t B, bool have_apk_jac) {
    static const int split_mode = [] { const char* e = getenv("HBLS_SPLIT"); return e ? atoi(e) : 1; }();
    static const int rlc_env = [] { const char* e = getenv("HBLS_RLC"); return e ? atoi(e) : 1; }();
    static const size_t rlc_min = [] { const char* e = getenv("HBLS_RLC_MIN"); return e ? (size_t)atol(e) : (size_t)1024; }();
    return split_mode && rlc_env && g.batch_mode == 1 && have_apk_jac && B >= rlc_min;
}
#define STAGE_EV(i, strm) do { if (g.stage_timing) cudaEventRecord(g.ev[i], (strm)); } while (0)
void launch_verify_tail(size_t B, const VerifyBufs& v, const uint8_t* d_sig96, const uint8_t* d_msgs, uint32_t msg_len,
                        const uint8_t* ok_pk, uint8_t* d_results, cudaStream_t s, bool same_msg = false, const g1* apk_jac = nullptr) {
    STAGE_EV(2, s);
    LAUNCH(k_g2_decode, big_blocks(B), big_tpb(), s, B, d_sig96, v.sig, v.ok_sig, 1);
    STAGE_EV(3, s);
    if (same_msg && B > 1) {
        LAUNCH(k_hash_to_g2, 1, TPB, s, (size_t)1, d_msgs, msg_len, v.hm, v.ok_hm);
        LAUNCH(k_broadcast_hm, blocks_for(B, 256), 256, s, B, v.hm, v.ok_hm);
    } else
    LAUNCH(k_hash_to_g2, big_blocks(B), big_tpb(), s, B, d_msgs, msg_len, v.hm, v.ok_hm);
    STAGE_EV(4, s);
    static const int fuse_mode = [] { const char* e = getenv("HBLS_FUSE"); return e ? atoi(e) : -1; }();   // -1 auto, 0 split, 1 fused
    const bool fused = fuse_mode == 1 || (fuse_mode == -1 && B >= (size_t)g.sm_count * 256);
    static const int split_mode = [] { const char* e = getenv("HBLS_SPLIT"); return e ? atoi(e) : 1; }();                // lane-pair pairing kernel
    if (rlc_applies(B, apk_jac != nullptr)) {
        // batched form (north-star "batched Miller loop + shared final exponentiation"): groups of HB_RLC_G rounds
        // group size: 8 once that still gives every SM a full CTA of lane pairs (fewer Miller-loop pairs and final
        // exponentiations per round), else 4
        static const int g_env = [] { const char* e = getenv("HBLS_RLC_G"); return e ? atoi(e) : 0; }();
        const size_t G = g_env == 4 || g_env == 8 ? (size_t)g_env : (2 * (B / 8) >= (size_t)g.sm_count * HB_TPB_SPLIT ? 8 : 4);
        const size_t ng = B / G, nr = ng * G, tail = B - nr;
        const uint64_t s0 = g.rlc_seed[0] + 0x9e3779b97f4a7c15ull * (++g.rlc_calls), s1 = g.rlc_seed[1] ^ (g.rlc_calls << 32);
        cudaMemsetAsync(v.any_fail, 0, sizeof(int), s);
        LAUNCH(k_rlc_scale, big_blocks(nr), big_tpb(), s, nr, ng, apk_jac, v.sig, v.hm, v.ok_sig, v.ok_hm, s0, s1, v.pk_scaled, v.S, v.bad);
        const bool full = 2 * ng >= (size_t)g.sm_count * HB_TPB_SPLIT;
        const unsigned pb = full ? split_blocks(2 * ng) : blocks_for(2 * ng, 64), pt = full ? HB_TPB_SPLIT : 64;
        if (G == 8) {
            LAUNCH(k_rlc_group_sum<8>, heavy_blocks(ng), TPB, s, ng, v.S, v.Sg);
            STAGE_EV(5, s);
            LAUNCH(k_rlc_pairing_split<8>, pb, pt, s, ng, v.pk_scaled, v.hm, v.Sg, v.bad, v.group_ok);
        } else {
            LAUNCH(k_rlc_group_sum<4>, heavy_blocks(ng), TPB, s, ng, v.S, v.Sg);
            STAGE_EV(5, s);
            LAUNCH(k_rlc_pairing_split<4>, pb, pt, s, ng, v.pk_scaled, v.hm, v.Sg, v.bad, v.group_ok);
        }
        LAUNCH(k_rlc_finish, blocks_for(nr, 256), 256, s, nr, ng, v.group_ok, d_results, v.any_fail);
#if HB_FALLBACK_LIST
        // exact pass over the rounds of failed groups only (compacted on the device; the launches are sized for the worst case
        // and return at once when the list is short or empty)
        cudaMemsetAsync(v.fail_count, 0, sizeof(unsigned), s);
        LAUNCH(k_rlc_collect_failed, blocks_for(nr, 256), 256, s, nr, ng, v.group_ok, v.fail_list, v.fail_count);
        LAUNCH(k_g1_normalize_list, heavy_blocks(B), TPB, s, v.fail_count, v.fail_list, apk_jac, v.pkneg, 1);
        if (2 * B >= (size_t)g.sm_count * HB_TPB_SPLIT)
            LAUNCH(k_pairing_verify_split_list, split_blocks(2 * B), HB_TPB_SPLIT, s, v.fail_count, v.fail_list, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results);
        else
            LAUNCH(k_pairing_verify_split_list, blocks_for(2 * B, 64), 64, s, v.fail_count, v.fail_list, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results);
        LAUNCH(k_pairing_fixup_list, heavy_blocks(B), TPB, s, v.fail_count, v.fail_list, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results);
#else
        // exact per-round pass: returns immediately unless a group failed (then every round is recomputed exactly)
        LAUNCH(k_g1_normalize, blocks_for(B, TPB), TPB, s, B, apk_jac, v.pkneg, 1, (const int*)v.any_fail);
        if (2 * B >= (size_t)g.sm_count * HB_TPB_SPLIT)
            LAUNCH(k_pairing_verify_split, split_blocks(2 * B), HB_TPB_SPLIT, s, B, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results, v.any_fail);
        else
            LAUNCH(k_pairing_verify_split, blocks_for(2 * B, 64), 64, s, B, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results, v.any_fail);
        LAUNCH(k_pairing_fixup, heavy_blocks(B), TPB, s, B, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results, v.any_fail);
#endif
        if (tail) {       // the < G rounds that do not fill a group are always verified exactly
            LAUNCH(k_g1_normalize, 1, TPB, s, tail, apk_jac + nr, v.pkneg + nr, 1, (const int*)nullptr);
            LAUNCH(k_pairing_verify_split, blocks_for(2 * tail, 64), 64, s, tail, v.sig + nr, v.pkneg + nr, v.hm + nr, v.ok_sig + nr, v.ok_hm + nr,
                   (const uint8_t*)nullptr, d_results + nr, (const int*)nullptr);
            LAUNCH(k_pairing_fixup, 1, TPB, s, tail, v.sig + nr, v.pkneg + nr, v.hm + nr, v.ok_sig + nr, v.ok_hm + nr, (const uint8_t*)nullptr, d_results + nr, (const int*)nullptr);
        }
    } else if (split_mode) {
        // default at every batch size: a lane pair per round (half the per-thread state, half the single-round latency).
        // Large batches use 512-thread lock-stepped CTAs (one per SM); small ones 64-thread CTAs spread over the SMs.
        STAGE_EV(5, s);
        if (2 * B >= (size_t)g.sm_count * HB_TPB_SPLIT)
            LAUNCH(k_pairing_verify_split, split_blocks(2 * B), HB_TPB_SPLIT, s, B, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results, (const int*)nullptr);
        else
            LAUNCH(k_pairing_verify_split, blocks_for(2 * B, 64), 64, s, B, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results, (const int*)nullptr);
        LAUNCH(k_pairing_fixup, heavy_blocks(B), TPB, s, B, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results, (const int*)nullptr);
    } else if (fused) {
        // batch alone fills the chip: one thread per round, 2-pair loop with shared squarings + final exponentiation
        STAGE_EV(5, s);
        LAUNCH(k_pairing_verify, heavy_blocks(B), TPB, s, B, v.sig, v.pkneg, v.hm, v.ok_sig, v.ok_hm, ok_pk, d_results);
    } else {
        // small batch: two threads per round for the Miller loops (more parallelism), then one for the exponentiation
        LAUNCH(k_miller_verify, heavy_blocks(2 * B), TPB, s, B, v.sig, v.pkneg, v.hm, v.f);
        STAGE_EV(5, s);
        LAUNCH(k_final_verify, heavy_blocks(B), TPB, s, B, v.f, v.ok_sig, v.ok_hm, ok_pk, d_results);
    }
    STAGE_EV(6, s);
}

int single_op(int op, const void* a, size_t an, const void* b, size_t bn, void* out, size_t on, int* rc_out, uint32_t len = 0) {
    if (int e = ensure_init()) return e;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(an + bn + on + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* da = ar.take<uint8_t>(an ? an : 1); uint8_t* db = ar.take<uint8_t>(bn ? bn : 1);
    uint8_t* dout = ar.take<uint8_t>(on ? on : 1); int* drc = ar.take<int>(1);
    if (an) CK(cudaMemcpyAsync(da, a, an, cudaMemcpyHostToDevice, g.stream));
    if (bn) CK(cudaMemcpyAsync(db, b, bn, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_single, 1, 32, g.stream, op, da, db, dout, drc, len);
    int rc = 0;
    CK(cudaMemcpyAsync(&rc, drc, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    if (on && rc >= 0 && !((op == OP_G1_DES || op == OP_G2_DES) && rc == 0) && !(op == OP_MAP_SER && rc != 0))
        CK(cudaMemcpy(out, dout, on, cudaMemcpyDeviceToHost));
    *rc_out = rc;
    return 0;
}

// r (BLS12-381 group order), little-endian u64
const uint64_t R_ORDER[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
bool scalar_lt_r(const uint64_t k[4]) {
    for (int i = 3; i >= 0; i--) { if (k[i] < R_ORDER[i]) return true; if (k[i] > R_ORDER[i]) return false; }
    return false;
}

// SHA-512 (host; only for the test-only blsSign/blsVerify string API: mcl Fp::setHashOf)
struct Sha512 {
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static void digest(const uint8_t* msg, size_t len, uint8_t out[64]) {
        static const uint64_t K[80] = {
            0x428a2f98d728ae22ull,0x7137449123ef65cdull,0xb5c0fbcfec4d3b2full,0xe9b5dba58189dbbcull,0x3956c25bf348b538ull,0x59f111f1b605d019ull,0x923f82a4af194f9bull,0xab1c5ed5da6d8118ull,
            0xd807aa98a3030242ull,0x12835b0145706fbeull,0x243185be4ee4b28cull,0x550c7dc3d5ffb4e2ull,0x72be5d74f27b896full,0x80deb1fe3b1696b1ull,0x9bdc06a725c71235ull,0xc19bf174cf692694ull,
            0xe49b69c19ef14ad2ull,0xefbe4786384f25e3ull,0x0fc19dc68b8cd5b5ull,0x240ca1cc77ac9c65ull,0x2de92c6f592b0275ull,0x4a7484aa6ea6e483ull,0x5cb0a9dcbd41fbd4ull,0x76f988da831153b5ull,
            0x983e5152ee66dfabull,0xa831c66d2db43210ull,0xb00327c898fb213full,0xbf597fc7beef0ee4ull,0xc6e00bf33da88fc2ull,0xd5a79147930aa725ull,0x06ca6351e003826full,0x142929670a0e6e70ull,
            0x27b70a8546d22ffcull,0x2e1b21385c26c926ull,0x4d2c6dfc5ac42aedull,0x53380d139d95b3dfull,0x650a73548baf63deull,0x766a0abb3c77b2a8ull,0x81c2c92e47edaee6ull,0x92722c851482353bull,
            0xa2bfe8a14cf10364ull,0xa81a664bbc423001ull,0xc24b8b70d0f89791ull,0xc76c51a30654be30ull,0xd192e819d6ef5218ull,0xd69906245565a910ull,0xf40e35855771202aull,0x106aa07032bbd1b8ull,
            0x19a4c116b8d2d0c8ull,0x1e376c085141ab53ull,0x2748774cdf8eeb99ull,0x34b0bcb5e19b48a8ull,0x391c0cb3c5c95a63ull,0x4ed8aa4ae3418acbull,0x5b9cca4f7763e373ull,0x682e6ff3d6b2b8a3ull,
            0x748f82ee5defb2fcull,0x78a5636f43172f60ull,0x84c87814a1f0ab72ull,0x8cc702081a6439ecull,0x90befffa23631e28ull,0xa4506cebde82bde9ull,0xbef9a3f7b2c67915ull,0xc67178f2e372532bull,
            0xca273eceea26619cull,0xd186b8c721c0c207ull,0xeada7dd6cde0eb1eull,0xf57d4f7fee6ed178ull,0x06f067aa72176fbaull,0x0a637dc5a2c898a6ull,0x113f9804bef90daeull,0x1b710b35131c471bull,
            0x28db77f523047d84ull,0x32caab7b40c72493ull,0x3c9ebe0a15c9bebcull,0x431d67c49c100d4cull,0x4cc5d4becb3e42b6ull,0x597f299cfc657e2aull,0x5fcb6fab3ad6faecull,0x6c44198c4a475817ull};
        uint64_t h[8] = {0x6a09e667f3bcc908ull,0xbb67ae8584caa73bull,0x3c6ef372fe94f82bull,0xa54ff53a5f1d36f1ull,0x510e527fade682d1ull,0x9b05688c2b3e6c1full,0x1f83d9abfb41bd6bull,0x5be0cd19137e2179ull};
        std::vector<uint8_t> m(msg, msg + len);
        m.push_back(0x80);
        while (m.size() % 128 != 112) m.push_back(0);
        for (int i = 0; i < 8; i++) m.push_back(0);
        uint64_t bits = (uint64_t)len * 8;
        for (int i = 7; i >= 0; i--) m.push_back((uint8_t)(bits >> (8 * i)));
        for (size_t off = 0; off < m.size(); off += 128) {
            uint64_t w[80];
            for (int i = 0; i < 16; i++) { w[i] = 0; for (int j = 0; j < 8; j++) w[i] = (w[i] << 8) | m[off + 8 * i + j]; }
            for (int i = 16; i < 80; i++) {
                uint64_t s0 = rotr(w[i - 15], 1) ^ rotr(w[i - 15], 8) ^ (w[i - 15] >> 7);
                uint64_t s1 = rotr(w[i - 2], 19) ^ rotr(w[i - 2], 61) ^ (w[i - 2] >> 6);
                w[i] = w[i - 16] + s0 + w[i - 7] + s1;
            }
            uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], gg = h[6], hh = h[7];
            for (int i = 0; i < 80; i++) {
                uint64_t S1 = rotr(e, 14) ^ rotr(e, 18) ^ rotr(e, 41), ch = (e & f) ^ (~e & gg);
                uint64_t t1 = hh + S1 + ch + K[i] + w[i];
                uint64_t S0 = rotr(a, 28) ^ rotr(a, 34) ^ rotr(a, 39), mj = (a & b) ^ (a & c) ^ (b & c);
                uint64_t t2 = S0 + mj;
                hh = gg; gg = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
            }
            h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += gg; h[7] += hh;
        }
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
    }
};

}  // namespace

struct hbls_committee {
    size_t n = 0;
    g1a* table = nullptr;       // device, affine, Montgomery
    g1* total = nullptr;        // device, sum of all rows (lets dense bitmaps be aggregated from their complement)
};

extern "C" {

int hbls_init_device(int device) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.ready) return 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fprintf(stderr, "[hbls] no CUDA device (%s): this backend has no CPU fallback\n", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return HBLS_ERR_CUDA;
    }
    if (device < 0 || device >= count) device = 0;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    g.device = device; g.sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    { FILE* f = fopen("/dev/urandom", "rb"); if (!f || fread(g.rlc_seed, 1, 16, f) != 16) { if (f) fclose(f); fprintf(stderr, "[hbls] cannot read /dev/urandom\n"); return HBLS_ERR_CUDA; } fclose(f); }
    cudaFuncSetAttribute(k_rlc_pairing_split<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_rlc_pairing_split<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    // the heavy kernels keep their Fp12 temporaries in per-thread local memory: give L1 the whole 228 KB
    cudaFuncSetAttribute(k_miller_verify, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_final_verify, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_pairing_verify, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_pairing_verify_split, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_hash_to_g2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_g2_decode, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_rlc_scale, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_mask_aggregate_serial, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaFuncSetAttribute(k_mask_aggregate, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    g.ready = true;
    return 0;
}
int blsInit(int curve, int compiledTimeVar) {
    if (curve != HBLS_BLS12_381 || compiledTimeVar != HBLS_COMPILED_TIME_VAR) return -1;
    const char* d = getenv("HBLS_DEVICE");
    if (!d) d = getenv("LOCAL_RANK");
    return hbls_init_device(d ? atoi(d) : 0);
}
uint64_t hbls_kernel_launch_count(void) { return g.launches.load(); }
void hbls_set_batch_mode(int mode) { std::lock_guard<std::mutex> lk(g.mu); g.batch_mode = mode ? 1 : 0; }
int hbls_get_batch_mode(void) { std::lock_guard<std::mutex> lk(g.mu); return g.batch_mode; }

// ------------------------------------------------------------------ secret keys (host bytes; no group arithmetic)
int blsSecretKeySetByCSPRNG(blsSecretKey* sec) {
    FILE* f = fopen("/dev/urandom", "rb");
    if (!f) return -1;
    size_t got = fread(sec->d, 1, 32, f); fclose(f);
    if (got != 32) return -1;
    sec->d[3] &= 0x3fffffffffffffffull;          // < 2^254 < r
    return 0;
}
size_t blsSecretKeySerialize(void* buf, size_t maxBufSize, const blsSecretKey* sec) { if (maxBufSize < 32) return 0; memcpy(buf, sec->d, 32); return 32; }
size_t blsSecretKeyDeserialize(blsSecretKey* sec, const void* buf, size_t bufSize) {
    if (bufSize < 32) return 0;
    uint64_t k[4]; memcpy(k, buf, 32);
    if (!scalar_lt_r(k)) return 0;
    memcpy(sec->d, k, 32); return 32;
}
int blsSecretKeyIsEqual(const blsSecretKey* l, const blsSecretKey* r) { return memcmp(l, r, 32) == 0; }

// ------------------------------------------------------------------ single-element group ops
void blsPublicKeyAdd(blsPublicKey* pub, const blsPublicKey* rhs) { int rc; single_op(OP_G1_ADD, pub, 144, rhs, 144, pub, 144, &rc); }
void blsPublicKeySub(blsPublicKey* pub, const blsPublicKey* rhs) { int rc; single_op(OP_G1_SUB, pub, 144, rhs, 144, pub, 144, &rc); }
void blsSignatureAdd(blsSignature* sig, const blsSignature* rhs) { int rc; single_op(OP_G2_ADD, sig, 288, rhs, 288, sig, 288, &rc); }
int blsPublicKeyIsEqual(const blsPublicKey* l, const blsPublicKey* r) { int rc = 0; if (single_op(OP_G1_EQ, l, 144, r, 144, nullptr, 0, &rc)) return 0; return rc; }
int blsSignatureIsEqual(const blsSignature* l, const blsSignature* r) { int rc = 0; if (single_op(OP_G2_EQ, l, 288, r, 288, nullptr, 0, &rc)) return 0; return rc; }
size_t blsPublicKeySerialize(void* buf, size_t maxBufSize, const blsPublicKey* pub) {
    if (maxBufSize < 48) return 0; int rc = 0; if (single_op(OP_G1_SER, pub, 144, nullptr, 0, buf, 48, &rc)) return 0; return rc == 48 ? 48 : 0; }
size_t blsSignatureSerialize(void* buf, size_t maxBufSize, const blsSignature* sig) {
    if (maxBufSize < 96) return 0; int rc = 0; if (single_op(OP_G2_SER, sig, 288, nullptr, 0, buf, 96, &rc)) return 0; return rc == 96 ? 96 : 0; }
size_t blsPublicKeyDeserialize(blsPublicKey* pub, const void* buf, size_t bufSize) {
    if (bufSize < 48) return 0; int rc = 0; if (single_op(OP_G1_DES, buf, 48, nullptr, 0, pub, 144, &rc)) return 0; return rc == 48 ? 48 : 0; }
size_t blsSignatureDeserialize(blsSignature* sig, const void* buf, size_t bufSize) {
    if (bufSize < 96) return 0; int rc = 0; if (single_op(OP_G2_DES, buf, 96, nullptr, 0, sig, 288, &rc)) return 0; return rc == 96 ? 96 : 0; }
int hbls_map_to_g2(const void* msg, size_t msg_len, uint8_t out96[96]) {
    if (msg_len > 64) msg_len = 64;      // only the first 48 bytes matter
    int rc = -1; if (int e = single_op(OP_MAP_SER, msg, msg_len, nullptr, 0, out96, 96, &rc, (uint32_t)msg_len)) return e; return rc; }

void blsGetPublicKey(blsPublicKey* pub, const blsSecretKey* sec) {
    if (ensure_init()) { memset(pub, 0, sizeof *pub); return; }
    std::lock_guard<std::mutex> lk(g.mu);
    if (reserve(4096)) return;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* dsk = ar.take<uint8_t>(32); g1* dout = ar.take<g1>(1);
    cudaMemcpyAsync(dsk, sec->d, 32, cudaMemcpyHostToDevice, g.stream);
    LAUNCH(k_g1_mul_gen, 1, 32, g.stream, (size_t)1, dsk, dout);
    cudaMemcpyAsync(pub, dout, 144, cudaMemcpyDeviceToHost, g.stream);
    cudaStreamSynchronize(g.stream);
}
int blsSignHash(blsSignature* sig, const blsSecretKey* sec, const void* h, size_t size) {
    if (int e = ensure_init()) return e;
    std::lock_guard<std::mutex> lk(g.mu);
    if (size > 64) size = 64;
    if (int e = reserve(4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* dsk = ar.take<uint8_t>(32); uint8_t* dmsg = ar.take<uint8_t>(64); g2* dout = ar.take<g2>(1); uint8_t* dok = ar.take<uint8_t>(1);
    CK(cudaMemcpyAsync(dsk, sec->d, 32, cudaMemcpyHostToDevice, g.stream));
    if (size) CK(cudaMemcpyAsync(dmsg, h, size, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_sign_hash, 1, 32, g.stream, (size_t)1, dsk, dmsg, (uint32_t)size, dout, dok);
    uint8_t ok = 0;
    CK(cudaMemcpyAsync(sig, dout, 288, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaMemcpyAsync(&ok, dok, 1, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return ok ? 0 : -1;
}
int blsVerifyHash(const blsSignature* sig, const blsPublicKey* pub, const void* h, size_t size) {
    if (ensure_init()) return 0;
    std::lock_guard<std::mutex> lk(g.mu);
    if (size > 64) size = 64;
    if (reserve(verify_scratch_bytes(1) + 4096)) return 0;
    Arena ar{g.scratch, 0, g.scratch_cap};
    VerifyBufs v = carve_verify(ar, 1);
    g2* dsig = ar.take<g2>(1); uint8_t* dmsg = ar.take<uint8_t>(64); uint8_t* dres = ar.take<uint8_t>(1);
    cudaMemcpyAsync(v.apk, pub, 144, cudaMemcpyHostToDevice, g.stream);
    cudaMemcpyAsync(dsig, sig, 288, cudaMemcpyHostToDevice, g.stream);
    if (size) cudaMemcpyAsync(dmsg, h, size, cudaMemcpyHostToDevice, g.stream);
    // struct inputs are already-decoded Jacobian points: normalise instead of decoding
    LAUNCH(k_g1_normalize, 1, 32, g.stream, (size_t)1, v.apk, v.pkneg, 1, (const int*)nullptr);
    LAUNCH(k_g2_normalize, 1, 32, g.stream, (size_t)1, dsig, v.sig);
    LAUNCH(k_hash_to_g2, 1, 32, g.stream, (size_t)1, dmsg, (uint32_t)size, v.hm, v.ok_hm);
    LAUNCH(k_pairing_verify_split, 1, 64, g.stream, (size_t)1, v.sig, v.pkneg, v.hm, (const uint8_t*)v.ok_hm, (const uint8_t*)nullptr, (const uint8_t*)nullptr, dres, (const int*)nullptr);
    LAUNCH(k_pairing_fixup, 1, 32, g.stream, (size_t)1, v.sig, v.pkneg, v.hm, (const uint8_t*)v.ok_hm, (const uint8_t*)nullptr, (const uint8_t*)nullptr, dres, (const int*)nullptr);
    uint8_t res = 0;
    cudaMemcpyAsync(&res, dres, 1, cudaMemcpyDeviceToHost, g.stream);
    if (cudaStreamSynchronize(g.stream) != cudaSuccess) return 0;
    return res ? 1 : 0;
}
void blsSign(blsSignature* sig, const blsSecretKey* sec, const void* m, size_t size) {
    uint8_t dg[64]; Sha512::digest((const uint8_t*)m, size, dg);
    if (blsSignHash(sig, sec, dg, 64) != 0) memset(sig, 0, sizeof *sig);
}
int blsVerify(const blsSignature* sig, const blsPublicKey* pub, const void* m, size_t size) {
    uint8_t dg[64]; Sha512::digest((const uint8_t*)m, size, dg);
    return blsVerifyHash(sig, pub, dg, 64);
}

// ------------------------------------------------------------------ committee table
int hbls_committee_create(hbls_committee** out, const uint8_t* pk48, size_t n, size_t* bad_index) {
    if (int e = ensure_init()) return e;
    if (!out) return HBLS_ERR_ARG;
    std::lock_guard<std::mutex> lk(g.mu);
    hbls_committee* c = new hbls_committee; c->n = n;
    size_t nn = n ? n : 1;
    CK(cudaMalloc(&c->table, nn * sizeof(g1a)));
    if (int e = reserve(nn * 49 + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* din = ar.take<uint8_t>(nn * 48); uint8_t* dok = ar.take<uint8_t>(nn);
    std::vector<uint8_t> ok(nn, 1);
    if (n) {
        CK(cudaMemcpyAsync(din, pk48, n * 48, cudaMemcpyHostToDevice, g.stream));
        LAUNCH(k_g1_decode, blocks_for(n, TPB), TPB, g.stream, n, din, c->table, dok, 1, 0);
        CK(cudaMalloc(&c->total, sizeof(g1)));
        LAUNCH(k_g1_sum, 1, 128, g.stream, n, c->table, c->total);
        CK(cudaMemcpyAsync(ok.data(), dok, n, cudaMemcpyDeviceToHost, g.stream));
        CK(cudaStreamSynchronize(g.stream));
    }
    for (size_t i = 0; i < n; i++) if (!ok[i]) {
        if (bad_index) *bad_index = i;
        cudaFree(c->table); cudaFree(c->total); delete c; return HBLS_ERR_DECODE;
    }
    *out = c; return 0;
}
void hbls_committee_destroy(hbls_committee* c) { if (!c) return; std::lock_guard<std::mutex> lk(g.mu); cudaFree(c->table); cudaFree(c->total); delete c; }
size_t hbls_committee_size(const hbls_committee* c) { return c ? c->n : 0; }

int hbls_mask_aggregate(const hbls_committee* c, const uint8_t* bitmap, size_t blen, uint8_t out_pk48[48]) {
    if (int e = ensure_init()) return e;
    if (!c || blen != ((c->n + 7) >> 3)) return HBLS_ERR_ARG;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(blen + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* dbm = ar.take<uint8_t>(blen ? blen : 1); g1* dacc = ar.take<g1>(1); uint8_t* dout = ar.take<uint8_t>(48);
    if (blen) CK(cudaMemcpyAsync(dbm, bitmap, blen, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_mask_aggregate, 1, 32, g.stream, (size_t)1, c->n, c->table, dbm, blen, dacc);
    LAUNCH(k_g1_serialize, 1, 32, g.stream, (size_t)1, dacc, dout);
    CK(cudaMemcpyAsync(out_pk48, dout, 48, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int hbls_aggregate_sigs(const uint8_t* sig96, size_t n, uint8_t out96[96]) {
    if (int e = ensure_init()) return e;
    std::lock_guard<std::mutex> lk(g.mu);
    size_t nn = n ? n : 1;
    if (int e = reserve(nn * (96 + sizeof(g2a) + 1) + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* din = ar.take<uint8_t>(nn * 96); g2a* dpts = ar.take<g2a>(nn); uint8_t* dok = ar.take<uint8_t>(nn);
    g2* dsum = ar.take<g2>(1); uint8_t* dout = ar.take<uint8_t>(96);
    std::vector<uint8_t> ok(nn, 1);
    if (n) {
        CK(cudaMemcpyAsync(din, sig96, n * 96, cudaMemcpyHostToDevice, g.stream));
        LAUNCH(k_g2_decode, heavy_blocks(n), TPB, g.stream, n, din, dpts, dok, 1);
    }
    LAUNCH(k_g2_sum, 1, HB_SUM_THREADS, g.stream, n, dpts, dsum);
    LAUNCH(k_g2_serialize, 1, 32, g.stream, (size_t)1, dsum, dout);
    if (n) CK(cudaMemcpyAsync(ok.data(), dok, n, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaMemcpyAsync(out96, dout, 96, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    for (size_t i = 0; i < n; i++) if (!ok[i]) return HBLS_ERR_DECODE;
    return 0;
}

// ------------------------------------------------------------------ aggregate verification
static bool all_messages_equal(const uint8_t* msgs, size_t n, size_t msg_len) {
    for (size_t i = 1; i < n; i++) if (memcmp(msgs, msgs + i * msg_len, msg_len) != 0) return false;
    return n > 1;
}
static int agg_verify_device_locked(const hbls_committee* c, size_t B, const uint8_t* d_bitmaps, size_t blen, const uint8_t* d_sigs,
                                    const uint8_t* d_msgs, size_t msg_len, uint8_t* d_results, cudaStream_t s, Arena& ar, bool same_msg = false) {
    VerifyBufs v = carve_verify(ar, B);
    STAGE_EV(0, s);
    if (B >= (size_t)g.sm_count * 256)
        LAUNCH(k_mask_aggregate_serial, light_blocks(B), TPB, s, B, c->n, c->table, c->total, d_bitmaps, blen, v.apk);
    else
        LAUNCH(k_mask_aggregate, blocks_for(B * 32, 128), 128, s, B, c->n, c->table, d_bitmaps, blen, v.apk);
    STAGE_EV(1, s);
    // the batched check consumes the Jacobian sums directly; -apk in affine form is then only needed if a group fails
    if (!rlc_applies(B, true)) LAUNCH(k_g1_normalize, blocks_for(B, TPB), TPB, s, B, v.apk, v.pkneg, 1, (const int*)nullptr);
    launch_verify_tail(B, v, d_sigs, d_msgs, (uint32_t)msg_len, nullptr, d_results, s, same_msg, v.apk);
    if (g.stage_timing) g.stage_valid = true;
    return 0;
}
int hbls_aggregate_verify_batch_device(const hbls_committee* c, size_t B, const void* d_bitmaps, size_t blen, const void* d_sigs96,
                                       const void* d_msgs, size_t msg_len, void* d_results, void* stream) {
    if (int e = ensure_init()) return e;
    if (!c || blen != ((c->n + 7) >> 3) || msg_len > 64) return HBLS_ERR_ARG;
    if (B == 0) return 0;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(verify_scratch_bytes(B))) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    cudaStream_t s = stream ? (cudaStream_t)stream : g.stream;
    int rc = agg_verify_device_locked(c, B, (const uint8_t*)d_bitmaps, blen, (const uint8_t*)d_sigs96, (const uint8_t*)d_msgs, msg_len, (uint8_t*)d_results, s, ar);
    CK(cudaGetLastError());
    return rc;
}
int hbls_aggregate_verify_batch(const hbls_committee* c, size_t B, const uint8_t* bitmaps, size_t blen, const uint8_t* sigs96,
                                const uint8_t* msgs, size_t msg_len, uint8_t* results) {
    if (int e = ensure_init()) return e;
    if (!c || blen != ((c->n + 7) >> 3) || msg_len > 64) return HBLS_ERR_ARG;
    if (B == 0) return 0;
    std::lock_guard<std::mutex> lk(g.mu);
    size_t in_bytes = B * (blen + 96 + msg_len);
    if (int e = reserve(verify_scratch_bytes(B) + in_bytes + B + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* dbm = ar.take<uint8_t>(B * blen + 1); uint8_t* dsig = ar.take<uint8_t>(B * 96); uint8_t* dmsg = ar.take<uint8_t>(B * msg_len + 1);
    uint8_t* dres = ar.take<uint8_t>(B);
    if (blen) CK(cudaMemcpyAsync(dbm, bitmaps, B * blen, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(dsig, sigs96, B * 96, cudaMemcpyHostToDevice, g.stream));
    if (msg_len) CK(cudaMemcpyAsync(dmsg, msgs, B * msg_len, cudaMemcpyHostToDevice, g.stream));
    agg_verify_device_locked(c, B, dbm, blen, dsig, dmsg, msg_len, dres, g.stream, ar, all_messages_equal(msgs, B, msg_len));
    CK(cudaMemcpyAsync(results, dres, B, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}
int hbls_aggregate_verify(const hbls_committee* c, const uint8_t* bitmap, size_t blen, const uint8_t sig96[96], const void* msg, size_t msg_len) {
    if (msg_len > 64) msg_len = 64;       // only the first 48 bytes enter the map
    uint8_t res = 0;
    int rc = hbls_aggregate_verify_batch(c, 1, bitmap, blen, sig96, (const uint8_t*)msg, msg_len, &res);
    if (rc) return rc;
    return res ? 1 : 0;
}

int hbls_verify_batch(size_t k, const uint8_t* pk48, const uint8_t* sig96, const uint8_t* msgs, size_t msg_len, uint8_t* results) {
    if (int e = ensure_init()) return e;
    if (msg_len > 64) return HBLS_ERR_ARG;
    if (k == 0) return 0;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(verify_scratch_bytes(k) + k * (48 + 96 + msg_len + 1) + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    VerifyBufs v = carve_verify(ar, k);
    uint8_t* dpk = ar.take<uint8_t>(k * 48); uint8_t* dsig = ar.take<uint8_t>(k * 96); uint8_t* dmsg = ar.take<uint8_t>(k * msg_len + 1);
    uint8_t* dres = ar.take<uint8_t>(k);
    CK(cudaMemcpyAsync(dpk, pk48, k * 48, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(dsig, sig96, k * 96, cudaMemcpyHostToDevice, g.stream));
    if (msg_len) CK(cudaMemcpyAsync(dmsg, msgs, k * msg_len, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_g1_decode, blocks_for(k, TPB), TPB, g.stream, k, dpk, v.pkneg, v.ok_pk, 1, 1);
    launch_verify_tail(k, v, dsig, dmsg, (uint32_t)msg_len, v.ok_pk, dres, g.stream, all_messages_equal(msgs, k, msg_len));
    CK(cudaMemcpyAsync(results, dres, k, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int hbls_sign_hash_batch(size_t k, const uint8_t* sk32, const uint8_t* msgs, size_t msg_len, uint8_t* sig96_out, uint8_t* ok) {
    if (int e = ensure_init()) return e;
    if (msg_len > 64) return HBLS_ERR_ARG;
    if (k == 0) return 0;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(k * (32 + msg_len + sizeof(g2) + 96 + 1) + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* dsk = ar.take<uint8_t>(k * 32); uint8_t* dmsg = ar.take<uint8_t>(k * msg_len + 1); g2* dpts = ar.take<g2>(k);
    uint8_t* dout = ar.take<uint8_t>(k * 96); uint8_t* dok = ar.take<uint8_t>(k);
    CK(cudaMemcpyAsync(dsk, sk32, k * 32, cudaMemcpyHostToDevice, g.stream));
    if (msg_len) CK(cudaMemcpyAsync(dmsg, msgs, k * msg_len, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_sign_hash, blocks_for(k, TPB), TPB, g.stream, k, dsk, dmsg, (uint32_t)msg_len, dpts, dok);
    LAUNCH(k_g2_serialize, blocks_for(k, TPB), TPB, g.stream, k, dpts, dout);
    CK(cudaMemcpyAsync(sig96_out, dout, k * 96, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaMemcpyAsync(ok, dok, k, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}
int hbls_get_public_key_batch(size_t k, const uint8_t* sk32, uint8_t* pk48_out) {
    if (int e = ensure_init()) return e;
    if (k == 0) return 0;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(k * (32 + sizeof(g1) + 48) + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* dsk = ar.take<uint8_t>(k * 32); g1* dpts = ar.take<g1>(k); uint8_t* dout = ar.take<uint8_t>(k * 48);
    CK(cudaMemcpyAsync(dsk, sk32, k * 32, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_g1_mul_gen, blocks_for(k, TPB), TPB, g.stream, k, dsk, dpts);
    LAUNCH(k_g1_serialize, blocks_for(k, TPB), TPB, g.stream, k, dpts, dout);
    CK(cudaMemcpyAsync(pk48_out, dout, k * 48, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int hbls_debug_g2(const uint8_t sig96[96], uint8_t out512[512]) {
    int rc = -1; if (int e = single_op(OP_DBG_G2, sig96, 96, nullptr, 0, out512, 512, &rc)) return e; return rc; }
int hbls_fp_mul_batch(size_t n, const uint8_t* a48, const uint8_t* b48, uint8_t* out48) {
    if (int e = ensure_init()) return e;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(n * 144 + 4096)) return e;
    Arena ar{g.scratch, 0, g.scratch_cap};
    uint8_t* da = ar.take<uint8_t>(n * 48); uint8_t* db = ar.take<uint8_t>(n * 48); uint8_t* dout = ar.take<uint8_t>(n * 48);
    CK(cudaMemcpyAsync(da, a48, n * 48, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(db, b48, n * 48, cudaMemcpyHostToDevice, g.stream));
    LAUNCH(k_fp_mul, blocks_for(n, 128), 128, g.stream, n, da, db, dout);
    CK(cudaMemcpyAsync(out48, dout, n * 48, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

void hbls_stage_timing_enable(int on) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (on && !g.ev[0]) for (int i = 0; i < 8; i++) cudaEventCreate(&g.ev[i]);
    g.stage_timing = on != 0; g.stage_valid = false;
}
int hbls_stage_timing_get(float* ms_out, int max_stages) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.stage_valid) return 0;
    if (cudaEventSynchronize(g.ev[6]) != cudaSuccess) return 0;
    int n = max_stages < 6 ? max_stages : 6;
    for (int i = 0; i < n; i++) { float ms = 0; cudaEventElapsedTime(&ms, g.ev[i], g.ev[i + 1]); ms_out[i] = ms; }
    return n;
}
int hbls_selftest_split(uint32_t iters) {
    if (int e = ensure_init()) return e;
    std::lock_guard<std::mutex> lk(g.mu);
    if (int e = reserve(4096)) return e;
    uint32_t* d = reinterpret_cast<uint32_t*>(g.scratch);
    CK(cudaMemsetAsync(d, 0, 4, g.stream));
    LAUNCH(k_selftest_fp2h, 8, 64, g.stream, iters, 20240922u, d);
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, d, 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return (int)bad;
}
double hbls_probe_mac32_per_s(int iters) {
    if (ensure_init()) return -1.0;
    std::lock_guard<std::mutex> lk(g.mu);
    if (reserve(4096)) return -1.0;
    uint64_t* sink = reinterpret_cast<uint64_t*>(g.scratch);
    constexpr int ILP = 8;
    const int threads = 256, blocks = g.sm_count * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    LAUNCH(k_probe_imad<ILP>, blocks, threads, g.stream, iters / 8 + 1, 12345u, sink);     // warm-up
    cudaEventRecord(e0, g.stream);
    LAUNCH(k_probe_imad<ILP>, blocks, threads, g.stream, iters, 12345u, sink);
    cudaEventRecord(e1, g.stream);
    if (cudaStreamSynchronize(g.stream) != cudaSuccess) return -1.0;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    double macs = (double)blocks * threads * (double)iters * ILP;
    return macs / (ms * 1e-3);
}

}  // extern "C"
