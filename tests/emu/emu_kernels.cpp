// tests/emu/emu_kernels.cpp -- TEST HARNESS ONLY: runs the __global__ kernels of harmony_b200/csrc/kernels.cuh on the host
// (HB_HOST_EMU: software carry flags; thread-per-item kernels one "thread" after the other, the lane-pair kernels as a
// 2-thread CTA on two host threads with shuffles / barriers by rendezvous) in the launch order of hbls.cu's
// aggregate-verify pipeline, so the device LOGIC -- mask complement sums, strided batch groups, fallback flags, result
// codes -- is diffed against the oracle on the GPU-less build box.  Never linked into libhbls.so; the product has no CPU path.
#define HB_HOST_EMU 1
#include <cstring>
#include <cstdint>
#include <cstddef>
#include <vector>
struct hb_dim3 { unsigned x, y, z; };
static thread_local hb_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#define __global__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#include "../../harmony_b200/csrc/pairing.cuh"
static inline void __syncthreads() { if (blockDim.x == 2) hb::hb_emu_exchange(0); }       // 2-thread CTA: rendezvous; 1-thread: nothing
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }       // warp-cooperative kernels are not run here
#include "../../harmony_b200/csrc/kernels.cuh"
using namespace hb;

template <class F> static void run_seq(unsigned grid, unsigned block, F f) {
    gridDim = {grid, 1, 1}; blockDim = {block, 1, 1};
    for (unsigned b = 0; b < grid; b++) for (unsigned t = 0; t < block; t++) { blockIdx = {b, 0, 0}; threadIdx = {t, 0, 0}; f(); }
    gridDim = {1, 1, 1}; blockDim = {1, 1, 1}; blockIdx = {0, 0, 0}; threadIdx = {0, 0, 0};
}
template <class F> static void run_pair(F f) {                 // one CTA of two threads = one lane pair, persistent over all items
    hb_emu_pair_reset();
    auto lane = [&](unsigned t) {
        gridDim = {1, 1, 1}; blockDim = {2, 1, 1}; blockIdx = {0, 0, 0}; threadIdx = {t, 0, 0};
        hb_emu.role = (int)t; hb_emu.seq = 0;
        f();
        blockDim = {1, 1, 1}; threadIdx = {0, 0, 0}; hb_emu.role = 0;
    };
    std::thread th(lane, 1u); lane(0u); th.join();
}

// The aggregate-verify pipeline of hbls.cu (agg_verify_device_locked + launch_verify_tail) with groups of G rounds.
// mode 1: batched groups + exact fallback, mode 0: exact only.  Returns 0, or -3 if a committee key does not decode.
template <int G> static int aggregate_verify_batch(int mode, uint32_t n, const uint8_t* pks48, size_t B, const uint8_t* bitmaps, size_t blen,
                                          const uint8_t* sigs96, const uint8_t* msgs, uint32_t msg_len, uint64_t s0, uint64_t s1,
                                          uint8_t* results, int* any_fail_out, uint8_t* group_ok_out) {
    std::vector<g1a> table(n), pkneg(B), pk_scaled(B); std::vector<uint8_t> okk(n), ok_sig(B), ok_hm(B), bad(B), group_ok(B / G + 1);
    std::vector<g1> apk(B); std::vector<g2a> sig(B), hm(B), Sg(B / G + 1); std::vector<g2> S(B);
    run_seq(1, n, [&] { k_g1_decode(n, pks48, table.data(), okk.data(), 1, 0); });
    for (uint32_t i = 0; i < n; i++) if (!okk[i]) return -3;
    g1 total; pt_set_inf(total);
    for (uint32_t i = 0; i < n; i++) pt_add_mixed(total, total, table[i]);
    run_seq(2, 3, [&] { k_mask_aggregate_serial(B, n, table.data(), &total, bitmaps, blen, apk.data()); });     // grid-stride: 6 "threads"
    run_seq(2, 2, [&] { k_g2_decode(B, sigs96, sig.data(), ok_sig.data(), 1); });
    run_seq(3, 1, [&] { k_hash_to_g2(B, msgs, msg_len, hm.data(), ok_hm.data()); });
    int any_fail = 0;
    auto exact = [&](size_t off, size_t cnt, const int* run_if) {
        run_seq(1, (unsigned)cnt, [&] { k_g1_normalize(cnt, apk.data() + off, pkneg.data() + off, 1, run_if); });
        run_pair([&] { k_pairing_verify_split(cnt, sig.data() + off, pkneg.data() + off, hm.data() + off, ok_sig.data() + off, ok_hm.data() + off,
                                              (const uint8_t*)nullptr, results + off, run_if); });
        run_seq(1, 2, [&] { k_pairing_fixup(cnt, sig.data() + off, pkneg.data() + off, hm.data() + off, ok_sig.data() + off, ok_hm.data() + off,
                                            (const uint8_t*)nullptr, results + off, run_if); });
    };
    if (mode == 1 && B >= (size_t)G) {
        const size_t ng = B / G, nr = ng * G, tail = B - nr;
        run_seq(2, 2, [&] { k_rlc_scale(nr, ng, apk.data(), sig.data(), hm.data(), ok_sig.data(), ok_hm.data(), s0, s1, pk_scaled.data(), S.data(), bad.data()); });
        run_seq(1, 2, [&] { k_rlc_group_sum<G>(ng, S.data(), Sg.data()); });
        run_pair([&] { k_rlc_pairing_split<G>(ng, pk_scaled.data(), hm.data(), Sg.data(), bad.data(), group_ok.data()); });
        run_seq(1, (unsigned)nr, [&] { k_rlc_finish(nr, ng, group_ok.data(), results, &any_fail); });
#if HB_FALLBACK_LIST
        std::vector<uint32_t> list(B); unsigned count = 0;
        run_seq(2, (unsigned)((nr + 1) / 2), [&] { k_rlc_collect_failed(nr, ng, group_ok.data(), list.data(), &count); });
        run_seq(1, 3, [&] { k_g1_normalize_list(&count, list.data(), apk.data(), pkneg.data(), 1); });
        run_pair([&] { k_pairing_verify_split_list(&count, list.data(), sig.data(), pkneg.data(), hm.data(), ok_sig.data(), ok_hm.data(), (const uint8_t*)nullptr, results); });
        run_seq(1, 2, [&] { k_pairing_fixup_list(&count, list.data(), sig.data(), pkneg.data(), hm.data(), ok_sig.data(), ok_hm.data(), (const uint8_t*)nullptr, results); });
#else
        exact(0, B, &any_fail);
#endif
        if (tail) exact(nr, tail, nullptr);
        if (group_ok_out) std::memcpy(group_ok_out, group_ok.data(), ng);
    } else {
        exact(0, B, nullptr);
    }
    if (any_fail_out) *any_fail_out = any_fail;
    return 0;
}
extern "C" int emu_aggregate_verify_batch(int mode, uint32_t n, const uint8_t* pks48, size_t B, const uint8_t* bitmaps, size_t blen,
                                          const uint8_t* sigs96, const uint8_t* msgs, uint32_t msg_len, uint64_t s0, uint64_t s1,
                                          uint8_t* results, int* any_fail_out, uint8_t* group_ok_out, int G) {
    return G == 8 ? aggregate_verify_batch<8>(mode, n, pks48, B, bitmaps, blen, sigs96, msgs, msg_len, s0, s1, results, any_fail_out, group_ok_out)
                  : aggregate_verify_batch<4>(mode, n, pks48, B, bitmaps, blen, sigs96, msgs, msg_len, s0, s1, results, any_fail_out, group_ok_out);
}
