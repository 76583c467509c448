// harmony_b200/host/hbls_host_test.cpp -- the reference's own host-level tests restated on the C++ mirror
// (hbls_host.hpp) over libhbls.so.  Needs a GPU (every group operation is a kernel).  Exit code 0 = all passed.
//   crypto/bls/mask_test.go            TestNewMask, TestSetMask/SetBit/SetKey/SetKeysAtomic, TestCountEnabled, policies
//   consensus/quorum/quorom_test.go    TestSubmitVote (:73-125, AggregateVotes == manual Add chain),
//                                      TestAddNewVoteInvalidAggregateSig (:381-501), TestInvalidAggregateSig (:503-552)
//   consensus/signature/signature_test.go, internal/chain/sig.go + engine.go:619-642
#include <cstdio>
#include <cstdlib>
#include "hbls_host.hpp"

using namespace harmony;
static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); g_fail++; } } while (0)

static std::vector<bls::PrivateKeyWrapper> make_keys(int n) {
    std::vector<bls::PrivateKeyWrapper> v;
    for (int i = 0; i < n; i++) v.push_back(bls::WrapperFromPrivateKey(bls::RandPrivateKey()));
    return v;
}

static void test_mask() {
    auto keys = make_keys(9);
    std::vector<bls::PublicKeyWrapper> pubs; for (auto& k : keys) pubs.push_back(*k.Pub);
    auto m = bls::NewMask(pubs);
    CHECK(m->Len() == 2 && m->CountTotal() == 9 && m->CountEnabled() == 0);
    CHECK(!m->SetMask({0x01}).empty());                                   // mismatching bitmap lengths
    CHECK(m->SetBit(0, true).empty()); CHECK(m->SetKey(pubs[8].Bytes, true).empty());
    CHECK(m->GetMask() == std::vector<uint8_t>({0x01, 0x01}) && m->CountEnabled() == 2);
    bls_core::PublicKey exp; exp.Add(pubs[0].Object.get()); exp.Add(pubs[8].Object.get());
    CHECK(m->AggregatePublic->IsEqual(&exp));
    CHECK(m->SetBit(9, true) == "index out of range");
    bls::SerializedPublicKey zero{}; CHECK(m->SetKey(zero, true) == "key not found");
    CHECK(m->SetKeysAtomic({&pubs[1], &pubs[2]}, true).empty());
    bool on = false; CHECK(m->IndexEnabled(1, on).empty() && on); CHECK(m->KeyEnabled(pubs[2].Bytes, on).empty() && on);
    CHECK(m->IndexEnabled(3, on).empty() && !on);
    m->SetBit(0, false); m->SetBit(8, false); m->SetKeysAtomic({&pubs[1], &pubs[2]}, false);
    bls_core::PublicKey id; CHECK(m->AggregatePublic->IsEqual(&id) && m->CountEnabled() == 0);
    CHECK(m->AggregatePublic->Serialize() == std::vector<uint8_t>(48, 0));
    CHECK(!bls::ThresholdPolicy{1}.Check(*m) && !bls::CompletePolicy{}.Check(*m));
    CHECK(m->SetMask({0xff, 0x01}).empty() && bls::CompletePolicy{}.Check(*m));
    CHECK(m->GetPubKeyFromMask(true).size() == 9 && m->GetPubKeyFromMask(false).empty());
    std::vector<uint8_t> o; CHECK(bls::AggregateMasks({1, 2}, {0x10, 2}, o) && o == std::vector<uint8_t>({0x11, 2}));
}

static void test_quorum_votes() {
    auto keys = make_keys(8);
    std::vector<bls::PublicKeyWrapper> pubs; for (auto& k : keys) pubs.push_back(*k.Pub);
    std::vector<uint8_t> hash(32); for (int i = 0; i < 32; i++) hash[i] = (uint8_t)(i * 7 + 1);
    std::vector<std::unique_ptr<bls_core::Sign>> sigs;
    for (auto& k : keys) { sigs.emplace_back(k.Pri->SignHash(hash)); CHECK(sigs.back() != nullptr); }
    // TestSubmitVote: AggregateVotes == manual Add chain (serialised equality)
    std::vector<quorum::Ballot> ballots;
    for (int i = 0; i < 3; i++) ballots.push_back({{pubs[i].Bytes}, sigs[i]->Serialize()});
    ballots.push_back({{pubs[1].Bytes}, sigs[1]->Serialize()});           // duplicate signer ballot is skipped
    auto agg = quorum::AggregateVotes(ballots);
    bls_core::Sign manual; for (int i = 0; i < 3; i++) manual.Add(sigs[i].get());
    CHECK(agg->SerializeToHexStr() == manual.SerializeToHexStr());
    // 4-of-8 aggregate verifies against the 4-signer mask (quorom_test.go:465-472)
    auto a4 = bls::AggregateSig({sigs[0].get(), sigs[1].get(), sigs[2].get(), sigs[3].get()});
    auto mask = bls::NewMask(pubs); CHECK(mask->SetMask({0x0f}).empty());
    CHECK(a4->VerifyHash(mask->AggregatePublic.get(), hash));
    CHECK(!quorum::IsQuorumAchievedByMask(mask.get(), 8));                // 4 < 8*2/3+1 = 6
    CHECK(mask->SetMask({0x3f}).empty() && quorum::IsQuorumAchievedByMask(mask.get(), 8));
    CHECK(!a4->VerifyHash(mask->AggregatePublic.get(), hash));
    // TestInvalidAggregateSig: duplicated signer fails against the de-duplicated key, correct set verifies
    auto dup = bls::AggregateSig({sigs[0].get(), sigs[1].get(), sigs[1].get()});
    bls_core::PublicKey apk; apk.Add(pubs[0].Object.get()); apk.Add(pubs[1].Object.get());
    CHECK(!dup->VerifyHash(&apk, hash));
    auto okk = bls::AggregateSig({sigs[0].get(), sigs[1].get()});
    CHECK(okk->VerifyHash(&apk, hash));
    // engine.go:619-642 through the device-resident committee; sig.go helpers
    bls::Committee com; CHECK(com.Load(pubs).empty());
    auto a6 = bls::AggregateSig({sigs[0].get(), sigs[1].get(), sigs[2].get(), sigs[3].get(), sigs[4].get(), sigs[5].get()});
    std::vector<uint8_t> payload = a6->Serialize(); payload.push_back(0x3f);
    bls::SerializedSignature s96; std::vector<uint8_t> bitmap;
    CHECK(chain::ParseCommitSigAndBitmap(payload, s96, bitmap) && bitmap == std::vector<uint8_t>({0x3f}));
    chain::SignatureVerifier eng;
    std::array<uint8_t, 32> bh; std::copy(hash.begin(), hash.end(), bh.begin());
    CHECK(eng.verifySignatureCached(com, bh, s96, bitmap, hash).empty());
    CHECK(eng.verifySignatureCached(com, bh, s96, bitmap, hash).empty());                       // cache hit
    CHECK(eng.verifySignature(com, s96, {0x0f}, hash) == "not enough signature collected");
    CHECK(eng.verifySignature(com, s96, {0x7e}, hash) == "Unable to verify aggregated signature for block");
    CHECK(!eng.verifySignature(com, s96, {0x3f, 0x00}, hash).empty());
    std::shared_ptr<bls_core::Sign> dsig; std::unique_ptr<bls::Mask> dmask;
    CHECK(chain::DecodeSigBitmap(s96, bitmap, pubs, dsig, dmask).empty() && dsig->VerifyHash(dmask->AggregatePublic.get(), hash));
    CHECK(chain::DecodeSigBitmap(s96, {0x3f, 0x00}, pubs, dsig, dmask) == "mask.SetMask failed");
    CHECK(bls::FastAggregateVerify(com, bitmap, s96, hash) == 1 && bls::VerifyAggregateSig(com, {0x1f}, s96, hash) == 0);
    // padding bits are not votes: 6-key committee (1-byte bitmap, quorum 5).  4 real signers + the two padding bits give a raw
    // popcount of 6, and the aggregate of the 4 verifies -- the reference counts mask.Bitmap after SetMask (4 < 5) and rejects.
    {
        std::vector<bls::PublicKeyWrapper> six(pubs.begin(), pubs.begin() + 6);
        bls::Committee c6; CHECK(c6.Load(six).empty());
        bls::SerializedSignature s4; auto b4 = a4->Serialize(); std::copy(b4.begin(), b4.end(), s4.begin());
        CHECK(bls::FastAggregateVerify(c6, {0xcf}, s4, hash) == 1);                               // the device ignores the padding bits
        CHECK(eng.verifySignature(c6, s4, {0xcf}, hash) == "not enough signature collected");
        CHECK(eng.verifySignature(c6, s4, {0x0f}, hash) == "not enough signature collected");
        // header range: valid / below quorum (padding attack) / wrong payload / undecodable signature / valid
        bls::SerializedSignature s6; auto b6 = a6->Serialize(); std::copy(b6.begin(), b6.end(), s6.begin());
        std::vector<chain::HeaderSig> hs(5);
        for (auto& x : hs) { x.commitSig = s6; x.commitBitmap = {0x3f}; x.commitPayload = hash; }
        hs[1].commitSig = s4; hs[1].commitBitmap = {0xcf};
        hs[2].commitPayload[3] ^= 1;
        hs[3].commitSig.fill(0xff);
        auto errs = chain::VerifyHeaderSignatures(c6, hs);
        CHECK(errs[0].empty() && errs[4].empty());
        CHECK(errs[1] == "not enough signature collected");
        CHECK(errs[2] == "Unable to verify aggregated signature for block");
        CHECK(errs[3] == "deserialize signature and bitmap: unable to deserialize multi-signature from payload");
    }
    // GetAddress (internal/utils/utils.go:77): 20 bytes, deterministic, differs between keys
    CHECK(pubs[0].Object->GetAddress() == pubs[0].Object->GetAddress() && pubs[0].Object->GetAddress() != pubs[1].Object->GetAddress());
}

static void test_codecs_multibls_payload() {
    // core/tx_pool_test.go:52-53 golden pair
    bls_core::SecretKey sk; CHECK(sk.DeserializeHexStr("c6d7603520311f7a4e6aac0b26701fc433b75b38df504cd416ef2b900cd66205"));
    std::unique_ptr<bls_core::PublicKey> pk(sk.GetPublicKey());
    CHECK(pk->SerializeToHexStr() == "30b2c38b1316da91e068ac3bd8751c0901ef6c02a1d58bc712104918302c6ed03d5894671d0c816dad2b4d303320f202");
    std::string err; auto viaCache = bls::BytesToBLSPublicKey(pk->Serialize(), &err);
    CHECK(viaCache && viaCache->IsEqual(pk.get()) && bls::BLSPubKeyCache().Len() >= 1);
    CHECK(bls::BytesToBLSPublicKey({}, &err) == nullptr && err == "BytesToBLSPublicKey: empty input");
    CHECK(bls::BytesToBLSPublicKey(std::vector<uint8_t>(48, 0xff), &err) == nullptr);
    auto s1 = bls::RandPrivateKey(), s2 = bls::RandPrivateKey();
    auto mk = multibls::GetPrivateKeys({s1, s2, s1});
    CHECK(mk.size() == 3 && mk.Dedup().size() == 2);
    auto pubs = mk.GetPublicKeys(); CHECK(pubs.Contains(mk[1].Pub->Object.get()));
    std::unique_ptr<bls_core::PublicKey> other(bls::RandPrivateKey()->GetPublicKey()); CHECK(!pubs.Contains(other.get()));
    CHECK(pubs.SerializeToHexStr() == mk[0].Pub->Hex() + ";" + mk[1].Pub->Hex() + ";" + mk[2].Pub->Hex() + ";");
    std::array<uint8_t, 32> h; for (int i = 0; i < 32; i++) h[i] = (uint8_t)i;
    auto p = signature::ConstructCommitPayload(true, h, 0x0102030405060708ull, 0x1112131415161718ull);
    CHECK(p.size() == 48 && p[0] == 8 && p[7] == 1 && p[8] == 0 && p[39] == 31 && p[40] == 0x18 && p[47] == 0x11);
    CHECK(signature::ConstructCommitPayload(false, h, 1, 2).size() == 40);
    std::vector<uint8_t> a, b; CHECK(!bls::SeparateSigAndMask(std::vector<uint8_t>(95), a, b));
    CHECK(bls::SeparateSigAndMask(std::vector<uint8_t>(100, 7), a, b) && a.size() == 96 && b.size() == 4);
    // string Sign/Verify round trip (consensus/quorum/quorom_test.go:97-121 style)
    std::unique_ptr<bls_core::Sign> sg(s1->SignMsg("test")); std::unique_ptr<bls_core::PublicKey> p1(s1->GetPublicKey());
    CHECK(sg->Verify(p1.get(), "test") && !sg->Verify(p1.get(), "tesT"));
    // SignHash returns nil when the message maps to no point (t = 0)
    CHECK(s1->SignHash(std::vector<uint8_t>(8, 0)) == nullptr);
}

int main() {
    if (bls_core::Init(bls_core::BLS12_381) != 0) { fprintf(stderr, "bls.Init failed: CUDA device required (no CPU fallback)\n"); return 2; }
    test_mask(); test_quorum_votes(); test_codecs_multibls_payload();
    if (g_fail) { fprintf(stderr, "%d check(s) failed\n", g_fail); return 1; }
    printf("hbls_host_test: all checks passed\n");
    return 0;
}
