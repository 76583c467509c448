// harmony_b200/host/hbls_host_cputest.cpp -- CPU-only checks of the host mirror's non-arithmetic logic (no GPU needed):
// hex codecs, commit payload bytes (consensus/signature/signature_test.go), sig||bitmap parsing (internal/chain/sig.go,
// crypto/bls/bls.go:120-136), quorum threshold / popcount (quorum.go:409-411, one-node-one-vote.go:57-72), the 1024-entry
// public-key LRU (crypto/bls/mask.go:35-55), and that every group operation refuses to run without blsInit (no CPU fallback).
#include <cstdio>
#include "hbls_host.hpp"
#include "hbls_consensus.hpp"
using namespace harmony;
static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); g_fail++; } } while (0)
int main() {
    std::vector<uint8_t> b; CHECK(bls_core::unhex("00ff10Ab", b) && b == std::vector<uint8_t>({0x00, 0xff, 0x10, 0xab}));
    CHECK(!bls_core::unhex("0", b) && !bls_core::unhex("zz", b));
    CHECK(bls_core::hex(b.data(), 4) == "00ff10ab");
    std::array<uint8_t, 32> h; for (int i = 0; i < 32; i++) h[i] = (uint8_t)(0xa0 + i);
    auto p = signature::ConstructCommitPayload(true, h, 0x0102030405060708ull, 0x1112131415161718ull);
    CHECK(p.size() == 48 && p[0] == 0x08 && p[7] == 0x01 && p[8] == 0xa0 && p[39] == 0xbf && p[40] == 0x18 && p[47] == 0x11);
    CHECK(signature::ConstructCommitPayload(false, h, 7, 9).size() == 40);
    std::vector<uint8_t> payload(96 + 32, 0); payload[0] = 1; payload[96] = 0xfe; payload[127] = 0x03;
    bls::SerializedSignature s96; std::vector<uint8_t> bm;
    CHECK(chain::ParseCommitSigAndBitmap(payload, s96, bm) && s96[0] == 1 && bm.size() == 32 && bm[0] == 0xfe && bm[31] == 0x03);
    CHECK(!chain::ParseCommitSigAndBitmap(std::vector<uint8_t>(95), s96, bm));
    std::vector<uint8_t> a, m; CHECK(bls::SeparateSigAndMask(payload, a, m) && a.size() == 96 && m == bm);
    CHECK(!bls::SeparateSigAndMask(std::vector<uint8_t>(10), a, m));
    CHECK(quorum::TwoThirdsSignersCount(250) == 167 && quorum::TwoThirdsSignersCount(4) == 3 && quorum::TwoThirdsSignersCount(1000) == 667);
    CHECK(quorum::CountOneBits({0xff, 0x01, 0x80}) == 10);
    // only the committee's slots count: 250 slots = 31 full bytes + 2 bits; the 6 padding bits of byte 31 are not votes
    { std::vector<uint8_t> bm(32, 0); bm[31] = 0xff; CHECK(quorum::CountSlotBits(bm, 250) == 2 && quorum::CountSlotBits(bm, 256) == 8 && quorum::CountOneBits(bm) == 8);
      bm.assign(32, 0xff); CHECK(quorum::CountSlotBits(bm, 250) == 250 && quorum::CountSlotBits(bm, 6) == 6); }
    bls::PubKeyCache cache(3); bls_core::PublicKey pk{}; pk.v.d[0] = 1;
    cache.Add("a", pk); pk.v.d[0] = 2; cache.Add("b", pk); pk.v.d[0] = 3; cache.Add("c", pk);
    bls_core::PublicKey out{}; CHECK(cache.Get("a", out) && out.v.d[0] == 1);          // "a" becomes most recent
    pk.v.d[0] = 4; cache.Add("d", pk);                                                   // evicts "b"
    CHECK(cache.Len() == 3 && !cache.Get("b", out) && cache.Get("c", out) && cache.Get("d", out) && out.v.d[0] == 4);
    std::vector<uint8_t> o; CHECK(bls::AggregateMasks({1, 2}, {4, 2}, o) && o == std::vector<uint8_t>({5, 2}) && !bls::AggregateMasks({1}, {1, 2}, o));
    bls::SerializedPublicKey z{}; CHECK(bls::IsEmpty(z) && bls::Hex(z).size() == 96);
    // range form: malformed records are rejected on the host, the rest is packed by payload length; quorum and signature checks run
    // in hbls_verify_headers in the reference's order (engine.go:630-640)
    {
        std::vector<chain::HeaderSig> hs(5);
        for (auto& x : hs) { x.commitBitmap.assign(1, 0x07); x.commitPayload.assign(48, 0x11); x.commitSig.fill(0x22); }   // committee of 4: quorum = 3 bits
        hs[1].commitBitmap[0] = 0x03;                       // 2 of 4: below quorum
        hs[2].commitBitmap.assign(2, 0xff);                 // wrong bitmap length
        hs[3].commitPayload.assign(40, 0x33);               // pre-staking payload: its own batch
        hs[4].commitSig.fill(0x44);
        std::vector<std::string> errs;
        auto batches = chain::AssembleHeaderBatches(4, hs, errs);
        CHECK(errs[0].empty() && errs[1].empty() && errs[2] == "deserialize signature and bitmap: mask.SetMask failed" && errs[3].empty() && errs[4].empty());
        CHECK(batches.size() == 2 && batches[0].msgLen == 48 && batches[0].index == std::vector<size_t>({0, 1, 4}) && batches[1].msgLen == 40 && batches[1].index == std::vector<size_t>({3}));
        CHECK(batches[0].sigs.size() == 288 && batches[0].sigs[0] == 0x22 && batches[0].sigs[192] == 0x44 && batches[0].msgs.size() == 144 && batches[0].bitmaps.size() == 3);
        CHECK(std::string(chain::headerStatusError(HBLS_HDR_NO_QUORUM)) == "not enough signature collected" && std::string(chain::headerStatusError(HBLS_HDR_OK)).empty());
        CHECK(batches[1].msgs.size() == 40 && batches[1].msgs[0] == 0x33);
    }
    // key files (internal/blsgen/lib.go): the reference's own vectors, internal/blsgen/utils_test.go:30-43
    {
        CHECK(blsgen::hex_of(blsgen::md5((const uint8_t*)"", 0).data(), 16) == "d41d8cd98f00b204e9800998ecf8427e");
        CHECK(blsgen::hex_of(blsgen::md5((const uint8_t*)"harmony", 7).data(), 16).size() == 32);
        const uint8_t k0[32] = {0}, z[16] = {0}; uint8_t ct[16]; blsgen::Aes256(k0).encrypt(z, ct);
        CHECK(blsgen::hex_of(ct, 16) == "dc95c078a2408989ad48a21492842087");            // FIPS 197 / SP 800-38A all-zero AES-256 block
        struct { const char *sk, *pass, *blob; } v[2] = {
            {"78c88c331195591b396e3205830071901a7a79e14fd0ede7f06bfb4c5e9f3473", "",
             "1d97f32175d8875f251e15805fd08f0cda794d827cb02d2de7b10d10f36f951d68347bef1e7a3018bd865c6966219cd9c4d20b055c50f8e09a6a3a1666b7c112450f643cc3c175f541fae75da8a843d47993fe89ec85788fd6ea2e98"},
            {"c20fa8de733d08e27e3101436d41f6a3207b8bedad7525c6e91a77ae2a49cf56", "harmony",
             "194a2d68c37f037f36b28a560402d64ab007f949313b63d9a08f5adb55a061681c70d9119df2d2cdcae5da6e484550c03bad63aae7c1332a3647ce633999ac4ddbb4a40e213c7e88e604784fef40da9d2f28b392c9fb2462f5e51e9c"}};
        for (auto& t : v) {
            std::string sk, err;
            CHECK(blsgen::LoadBLSKeyHexWithPassPhrase(t.blob, std::string(" ") + t.pass + "\n", sk, &err) && sk == t.sk);     // passphrase is trimmed
            CHECK(!blsgen::LoadBLSKeyHexWithPassPhrase(t.blob, std::string(t.pass) + "x", sk, &err) && err == "cipher: message authentication failed");
            std::vector<uint8_t> raw; blsgen::unhex_to(t.blob, raw);
            CHECK(blsgen::encrypt(t.sk, t.pass, raw.data()) == t.blob);                                                       // same nonce -> same bytes
            CHECK(blsgen::LoadBLSKeyHexWithPassPhrase(std::string(raw.begin(), raw.end()), t.pass, sk, &err) && sk == t.sk);   // binary form fall-back
        }
        std::string sk, err; CHECK(!blsgen::LoadBLSKeyHexWithPassPhrase("", "", sk, &err) && !blsgen::LoadBLSKeyHexWithPassPhrase("zz", "", sk, &err));
    }
    // crypto/hash/hash.go Keccak256 (legacy padding) and the 48-byte view of a message the hash-to-G2 map reads
    {
        auto kh = [](const std::string& m) { auto d = hash::Keccak256((const uint8_t*)m.data(), m.size()); return bls_core::hex(d.data(), d.size()); };
        CHECK(kh("") == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470");
        CHECK(kh("abc") == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45");
        CHECK(kh("harmony-one") == "1eaafe555c82d1cb51afd79875683437858a9fb883c07487717921757d6717f9");      // staking/types/validator.go:30
        CHECK(kh(std::string(135, 'a')) == "34367dc248bbd832f4e3e69dfaac2f92638bd0bbd18f2912ba4ef454919cf446");      // one-byte 0x81 padding
        CHECK(kh(std::string(136, 'a')).substr(0, 16) == "a6c4d403279fe3e0" && kh(std::string(300, 'a')).substr(0, 16) == "5b7e0e47a96f32a8");
        consensus::Bytes blob; consensus::put48(blob, consensus::NIL()); consensus::put48(blob, consensus::le64(0x0102));
        consensus::put48(blob, consensus::Bytes(200, 7));
        CHECK(blob.size() == 144 && blob[0] == 1 && blob[1] == 0 && blob[47] == 0 && blob[48] == 2 && blob[49] == 1 && blob[50] == 0 && blob[96] == 7 && blob[143] == 7);
        CHECK(consensus::ValidPayloadLength == 128);
    }
    // no CPU fallback: without blsInit (no device here) group operations fail instead of computing on the host
    bls_core::PublicKey q{}; std::vector<uint8_t> k48(48, 0); k48[0] = 1;
    CHECK(!q.Deserialize(k48));
    CHECK(hbls_kernel_launch_count() == 0);
    if (g_fail) { fprintf(stderr, "%d check(s) failed\n", g_fail); return 1; }
    printf("hbls_host_cputest: all checks passed\n"); return 0;
}
