// harmony_b200/host/hbls_keyfile.hpp -- host mirror of the reference's BLS key files (reference internal/blsgen/lib.go:20-159).
//
// A key file holds  hex( nonce12 || AES-256-GCM(key, nonce12, hex(sk)) )  with  key = the 32 ASCII characters of hex(md5(passphrase))
// (lib.go:101-118: createHash / encrypt); its name is hex(pk) + ".key" (lib.go:24).  LoadBLSKeyWithPassPhrase trims the passphrase,
// accepts the hex form or -- as a fall-back -- the raw binary form (lib.go:120-137), and hands the plaintext to
// SecretKey.DeserializeHexStr (lib.go:63-68).  Pure host code (no group arithmetic, nothing on the GPU): MD5 (RFC 1321), AES-256
// (FIPS 197, encryption direction only) and GCM (SP 800-38D, 96-bit nonce, no AAD) written out here because the build image has no
// crypto library headers.  Pinned by the reference's own vectors (internal/blsgen/utils_test.go:30-43) and key files
// (tests/golden/ref_fixtures.json "keyfiles"): hbls_host_cputest.cpp, tests/test_keyfile.py.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace harmony {
namespace blsgen {

inline std::string hex_of(const uint8_t* p, size_t n) {
    static const char* d = "0123456789abcdef"; std::string s(2 * n, '0');
    for (size_t i = 0; i < n; i++) { s[2 * i] = d[p[i] >> 4]; s[2 * i + 1] = d[p[i] & 15]; }
    return s;
}
inline bool unhex_to(const std::string& s, std::vector<uint8_t>& out) {
    if (s.size() & 1) return false;
    auto nib = [](char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; };
    out.resize(s.size() / 2);
    for (size_t i = 0; i < out.size(); i++) { int a = nib(s[2 * i]), b = nib(s[2 * i + 1]); if (a < 0 || b < 0) return false; out[i] = (uint8_t)(a << 4 | b); }
    return true;
}

// ---- MD5
inline std::array<uint8_t, 16> md5(const uint8_t* msg, size_t len) {
    static const uint32_t S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                   4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t K[64];
    for (int i = 0; i < 64; i++) { double v = __builtin_fabs(__builtin_sin((double)(i + 1))) * 4294967296.0; K[i] = (uint32_t)v; }
    uint32_t a0 = 0x67452301u, b0 = 0xefcdab89u, c0 = 0x98badcfeu, d0 = 0x10325476u;
    std::vector<uint8_t> m(msg, msg + len); m.push_back(0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    const uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) m.push_back((uint8_t)(bits >> (8 * i)));
    for (size_t off = 0; off < m.size(); off += 64) {
        uint32_t M[16];
        for (int i = 0; i < 16; i++) M[i] = (uint32_t)m[off + 4 * i] | (uint32_t)m[off + 4 * i + 1] << 8 | (uint32_t)m[off + 4 * i + 2] << 16 | (uint32_t)m[off + 4 * i + 3] << 24;
        uint32_t A = a0, B = b0, C = c0, D = d0;
        for (int i = 0; i < 64; i++) {
            uint32_t F; int g;
            if (i < 16) { F = (B & C) | (~B & D); g = i; }
            else if (i < 32) { F = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
            else if (i < 48) { F = B ^ C ^ D; g = (3 * i + 5) & 15; }
            else { F = C ^ (B | ~D); g = (7 * i) & 15; }
            F += A + K[i] + M[g]; A = D; D = C; C = B; B += (F << S[i]) | (F >> (32 - S[i]));
        }
        a0 += A; b0 += B; c0 += C; d0 += D;
    }
    std::array<uint8_t, 16> out; const uint32_t w[4] = {a0, b0, c0, d0};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(w[i] >> (8 * j));
    return out;
}
// createHash (lib.go:101-105): the 32 hex characters of md5(passphrase), used AS the AES-256 key bytes
inline std::array<uint8_t, 32> createHash(const std::string& passphrase) {
    const auto d = md5(reinterpret_cast<const uint8_t*>(passphrase.data()), passphrase.size());
    const std::string h = hex_of(d.data(), 16);
    std::array<uint8_t, 32> k; std::memcpy(k.data(), h.data(), 32);
    return k;
}

// ---- AES-256, encryption direction (GCM needs nothing else)
struct Aes256 {
    uint8_t rk[15][16];
    static uint8_t xt(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }
    static const uint8_t* sbox() {
        static uint8_t S[256]; static bool done = false;
        if (!done) {                                      // S-box from its definition: multiplicative inverse in GF(2^8), then the affine map
            uint8_t p = 1, q = 1;
            do {
                p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1b : 0));
                q ^= (uint8_t)(q << 1); q ^= (uint8_t)(q << 2); q ^= (uint8_t)(q << 4); if (q & 0x80) q ^= 0x09;
                const uint8_t x = (uint8_t)(q ^ (uint8_t)(q << 1 | q >> 7) ^ (uint8_t)(q << 2 | q >> 6) ^ (uint8_t)(q << 3 | q >> 5) ^ (uint8_t)(q << 4 | q >> 4));
                S[p] = (uint8_t)(x ^ 0x63);
            } while (p != 1);
            S[0] = 0x63; done = true;
        }
        return S;
    }
    explicit Aes256(const uint8_t key[32]) {
        const uint8_t* S = sbox();
        uint8_t w[60][4];
        for (int i = 0; i < 8; i++) std::memcpy(w[i], key + 4 * i, 4);
        uint8_t rcon = 1;
        for (int i = 8; i < 60; i++) {
            uint8_t t[4]; std::memcpy(t, w[i - 1], 4);
            if (i % 8 == 0) { const uint8_t t0 = t[0]; t[0] = (uint8_t)(S[t[1]] ^ rcon); t[1] = S[t[2]]; t[2] = S[t[3]]; t[3] = S[t0]; rcon = xt(rcon); }
            else if (i % 8 == 4) for (int j = 0; j < 4; j++) t[j] = S[t[j]];
            for (int j = 0; j < 4; j++) w[i][j] = (uint8_t)(w[i - 8][j] ^ t[j]);
        }
        for (int r = 0; r < 15; r++) for (int c = 0; c < 4; c++) std::memcpy(rk[r] + 4 * c, w[4 * r + c], 4);
    }
    void encrypt(const uint8_t in[16], uint8_t out[16]) const {
        const uint8_t* S = sbox();
        uint8_t s[16]; for (int i = 0; i < 16; i++) s[i] = (uint8_t)(in[i] ^ rk[0][i]);
        for (int r = 1; r <= 14; r++) {
            uint8_t t[16];
            for (int c = 0; c < 4; c++) for (int row = 0; row < 4; row++) t[4 * c + row] = S[s[4 * ((c + row) & 3) + row]];      // SubBytes + ShiftRows
            if (r < 14) for (int c = 0; c < 4; c++) {                                                                                // MixColumns
                const uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3], x = (uint8_t)(a0 ^ a1 ^ a2 ^ a3);
                t[4 * c] = (uint8_t)(a0 ^ x ^ xt((uint8_t)(a0 ^ a1))); t[4 * c + 1] = (uint8_t)(a1 ^ x ^ xt((uint8_t)(a1 ^ a2)));
                t[4 * c + 2] = (uint8_t)(a2 ^ x ^ xt((uint8_t)(a2 ^ a3))); t[4 * c + 3] = (uint8_t)(a3 ^ x ^ xt((uint8_t)(a3 ^ a0)));
            }
            for (int i = 0; i < 16; i++) s[i] = (uint8_t)(t[i] ^ rk[r][i]);
        }
        std::memcpy(out, s, 16);
    }
};

// ---- GCM with a 96-bit nonce and no additional data (Go's cipher.NewGCM defaults, lib.go:107-118,139-159)
struct Gcm {
    Aes256 aes; uint8_t H[16];
    explicit Gcm(const uint8_t key[32]) : aes(key) { const uint8_t z[16] = {0}; aes.encrypt(z, H); }
    void gmul(uint8_t x[16]) const {                       // x <- x * H in GF(2^128) (bit-reflected convention of SP 800-38D)
        uint8_t z[16] = {0}, v[16]; std::memcpy(v, H, 16);
        for (int i = 0; i < 128; i++) {
            if ((x[i >> 3] >> (7 - (i & 7))) & 1) for (int j = 0; j < 16; j++) z[j] ^= v[j];
            const bool lsb = v[15] & 1;
            for (int j = 15; j > 0; j--) v[j] = (uint8_t)(v[j] >> 1 | v[j - 1] << 7);
            v[0] >>= 1; if (lsb) v[0] ^= 0xe1;
        }
        std::memcpy(x, z, 16);
    }
    void ctr(const uint8_t nonce[12], uint32_t first, const uint8_t* in, size_t n, uint8_t* out) const {
        uint8_t cb[16], ks[16]; std::memcpy(cb, nonce, 12);
        for (size_t off = 0; off < n; off += 16, first++) {
            cb[12] = (uint8_t)(first >> 24); cb[13] = (uint8_t)(first >> 16); cb[14] = (uint8_t)(first >> 8); cb[15] = (uint8_t)first;
            aes.encrypt(cb, ks);
            for (size_t j = 0; j < 16 && off + j < n; j++) out[off + j] = (uint8_t)(in[off + j] ^ ks[j]);
        }
    }
    void tag(const uint8_t nonce[12], const uint8_t* ct, size_t n, uint8_t out[16]) const {
        uint8_t y[16] = {0};
        for (size_t off = 0; off < n; off += 16) { for (size_t j = 0; j < 16 && off + j < n; j++) y[j] ^= ct[off + j]; gmul(y); }
        const uint64_t cbits = (uint64_t)n * 8;
        for (int j = 0; j < 8; j++) y[8 + j] ^= (uint8_t)(cbits >> (56 - 8 * j));      // len(A) = 0 || len(C)
        gmul(y);
        uint8_t j0[16], e[16]; std::memcpy(j0, nonce, 12); j0[12] = j0[13] = j0[14] = 0; j0[15] = 1;
        aes.encrypt(j0, e);
        for (int j = 0; j < 16; j++) out[j] = (uint8_t)(y[j] ^ e[j]);
    }
    // Seal(nonce, nonce, data, nil): nonce || ciphertext || tag
    std::vector<uint8_t> seal(const uint8_t nonce[12], const uint8_t* data, size_t n) const {
        std::vector<uint8_t> out(12 + n + 16); std::memcpy(out.data(), nonce, 12);
        ctr(nonce, 2, data, n, out.data() + 12); tag(nonce, out.data() + 12, n, out.data() + 12 + n);
        return out;
    }
    bool open(const uint8_t* blob, size_t len, std::vector<uint8_t>& plain) const {
        if (len < 12 + 16) return false;
        const size_t n = len - 28; uint8_t t[16]; tag(blob, blob + 12, n, t);
        uint8_t diff = 0; for (int j = 0; j < 16; j++) diff |= (uint8_t)(t[j] ^ blob[12 + n + j]);
        if (diff) return false;                             // "cipher: message authentication failed"
        plain.resize(n); ctr(blob, 2, blob + 12, n, plain.data());
        return true;
    }
};

// decryptRaw / decrypt (lib.go:120-159): hex form first, then the raw binary form; the error of the hex form wins
inline bool decryptRaw(const std::vector<uint8_t>& data, const std::string& passphrase, std::vector<uint8_t>& plain, std::string* err) {
    if (data.empty()) { if (err) *err = "unable to decrypt raw data with the provided passphrase; the data is empty"; return false; }
    if (data.size() < 12) { if (err) *err = "failed to decrypt raw data with the provided passphrase; the data size is invalid"; return false; }
    const auto key = createHash(passphrase);
    if (!Gcm(key.data()).open(data.data(), data.size(), plain)) { if (err) *err = "cipher: message authentication failed"; return false; }
    return true;
}
inline bool decrypt(const std::string& encrypted, const std::string& passphrase, std::vector<uint8_t>& plain, std::string* err) {
    std::vector<uint8_t> raw; std::string e1;
    if (unhex_to(encrypted, raw)) { if (decryptRaw(raw, passphrase, plain, &e1)) return true; }
    else e1 = "encoding/hex: invalid byte";
    std::string e2;
    if (decryptRaw(std::vector<uint8_t>(encrypted.begin(), encrypted.end()), passphrase, plain, &e2)) return true;
    if (err) *err = e1;
    return false;
}
// encrypt (lib.go:107-118) with a caller-supplied nonce (the reference draws it from crypto/rand)
inline std::string encrypt(const std::string& data, const std::string& passphrase, const uint8_t nonce[12]) {
    const auto key = createHash(passphrase);
    const auto ct = Gcm(key.data()).seal(nonce, reinterpret_cast<const uint8_t*>(data.data()), data.size());
    return hex_of(ct.data(), ct.size());
}
// the secret key (hex, as DeserializeHexStr takes it) inside a key file's content; passphrase is trimmed like strings.TrimSpace
inline bool LoadBLSKeyHexWithPassPhrase(const std::string& fileContent, std::string passphrase, std::string& skHex, std::string* err = nullptr) {
    const char* ws = " \t\r\n\v\f";
    const size_t b = passphrase.find_first_not_of(ws); passphrase = b == std::string::npos ? "" : passphrase.substr(b, passphrase.find_last_not_of(ws) - b + 1);
    std::vector<uint8_t> plain;
    if (!decrypt(fileContent, passphrase, plain, err)) return false;
    skHex.assign(plain.begin(), plain.end());
    return true;
}

}  // namespace blsgen
}  // namespace harmony
