/*
 * include/hbls.h -- C ABI of libhbls.so, the B200-native (CUDA sm_100a) BLS12-381 backend for Harmony's
 * signature aggregation / aggregate-verification hot path.
 *
 * This is exactly what the reference's FFI for this path binds:
 *   - Part 1 mirrors the herumi `bls.h` subset that `github.com/harmony-one/bls/ffi/go/bls` (reference go.mod:27)
 *     calls through cgo, built with BLS_SWAP_G=1 (reference Makefile:68-70): public keys in G1 (48 B),
 *     signatures in G2 (96 B).  Every Go identifier the reference uses maps to one function here.
 *   - Part 2 are additive batch entry points that `crypto/bls` wrappers call so that a whole committee round /
 *     block range crosses cgo once (mask.go:113-134 SetMask, mask.go:58-64 AggregateSig,
 *     internal/chain/engine.go:619-642 verifySignature, consensus/leader.go:227-290 onCommit loop).
 *
 * Conventions: plain pointers and sizes only; all memory caller-owned; no callbacks; int error codes.
 * Every function is re-entrant after blsInit (internally serialised on one CUDA stream).
 * There is NO CPU fallback: if no CUDA device is usable blsInit returns HBLS_ERR_CUDA and every other call fails.
 */
#ifndef HBLS_H
#define HBLS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HBLS_BLS12_381 5            /* bls.BLS12_381 (crypto/bls/mask.go:18-20) */
#define HBLS_COMPILED_TIME_VAR 46   /* MCLBN_FR_UNIT_SIZE*10 + MCLBN_FP_UNIT_SIZE = 4*10 + 6 */

#define HBLS_ERR_CUDA   (-100)      /* no device / CUDA runtime failure (message on stderr) */
#define HBLS_ERR_ARG    (-2)        /* bad argument (e.g. bitmap length mismatch: mask.go:114-120) */
#define HBLS_ERR_DECODE (-3)        /* an input failed to decode (not on curve / >= p / not in subgroup) */

/* ------------------------------------------------------------------ Part 1: herumi bls.h shape (SWAP_G) */
typedef struct { uint64_t d[4]; }  blsSecretKey;   /* 32 B: scalar < r, little-endian */
typedef struct { uint64_t d[18]; } blsPublicKey;   /* 144 B: G1 Jacobian, Montgomery limbs; all-zero = identity */
typedef struct { uint64_t d[36]; } blsSignature;   /* 288 B: G2 Jacobian, Montgomery limbs; all-zero = identity */

/* bls.Init(bls.BLS12_381): crypto/bls/mask.go:18-20, internal/utils/utils.go:33.  0 = ok. */
int blsInit(int curve, int compiledTimeVar);
/* like blsInit but pins the CUDA device (one process per GPU: device = LOCAL_RANK) */
int hbls_init_device(int device);

/* SecretKey.SetByCSPRNG (crypto/bls/mask.go:23-27 RandPrivateKey) */
int blsSecretKeySetByCSPRNG(blsSecretKey* sec);
/* SecretKey.GetPublicKey: pk = sk * B, B = herumi SWAP_G generator */
void blsGetPublicKey(blsPublicKey* pub, const blsSecretKey* sec);
/* SecretKey.SignHash (consensus/construct.go:101,110): 0 ok, -1 when the message maps to no point */
int blsSignHash(blsSignature* sig, const blsSecretKey* sec, const void* h, size_t size);
/* Sign.VerifyHash (consensus/leader.go:173,287; internal/chain/engine.go:638): 1 valid, 0 invalid */
int blsVerifyHash(const blsSignature* sig, const blsPublicKey* pub, const void* h, size_t size);
/* SecretKey.Sign(string) / Sign.Verify(string): tests only, bytes unpinned by the reference */
void blsSign(blsSignature* sig, const blsSecretKey* sec, const void* m, size_t size);
int blsVerify(const blsSignature* sig, const blsPublicKey* pub, const void* m, size_t size);

/* PublicKey.Add / Sub (crypto/bls/mask.go:126,130), Sign.Add (mask.go:61) */
void blsPublicKeyAdd(blsPublicKey* pub, const blsPublicKey* rhs);
void blsPublicKeySub(blsPublicKey* pub, const blsPublicKey* rhs);
void blsSignatureAdd(blsSignature* sig, const blsSignature* rhs);

/* Serialize: bytes written (32/48/96) or 0 on error.  Deserialize: bytes read or 0 on error
 * (>= p, not on curve, not in the r-torsion; scalar >= r). */
size_t blsSecretKeySerialize(void* buf, size_t maxBufSize, const blsSecretKey* sec);
size_t blsPublicKeySerialize(void* buf, size_t maxBufSize, const blsPublicKey* pub);
size_t blsSignatureSerialize(void* buf, size_t maxBufSize, const blsSignature* sig);
size_t blsSecretKeyDeserialize(blsSecretKey* sec, const void* buf, size_t bufSize);
size_t blsPublicKeyDeserialize(blsPublicKey* pub, const void* buf, size_t bufSize);
size_t blsSignatureDeserialize(blsSignature* sig, const void* buf, size_t bufSize);

int blsSecretKeyIsEqual(const blsSecretKey* lhs, const blsSecretKey* rhs);
int blsPublicKeyIsEqual(const blsPublicKey* lhs, const blsPublicKey* rhs);
int blsSignatureIsEqual(const blsSignature* lhs, const blsSignature* rhs);

/* ------------------------------------------------------------------ Part 2: batch entry points (device-resident committee) */
typedef struct hbls_committee hbls_committee;

/* Decode + subgroup-check n public keys once and keep the table in HBM (mirrors the epochCtx cache of
 * internal/chain/engine.go:644-659,727-761 and BLSPubKeyCache of crypto/bls/mask.go:35-55).
 * 0 ok; HBLS_ERR_DECODE if any key is invalid (*bad_index, if non-NULL, receives the first offender). */
int hbls_committee_create(hbls_committee** out, const uint8_t* pk48, size_t n, size_t* bad_index);
void hbls_committee_destroy(hbls_committee* c);
size_t hbls_committee_size(const hbls_committee* c);

/* Mask.SetMask on a fresh mask + AggregatePublic.Serialize (crypto/bls/mask.go:113-134).
 * bitmap: LSB-first within each byte; blen must equal (n+7)>>3 else HBLS_ERR_ARG. */
int hbls_mask_aggregate(const hbls_committee* c, const uint8_t* bitmap, size_t blen, uint8_t out_pk48[48]);

/* AggregateSig (crypto/bls/mask.go:58-64; consensus/quorum/quorum.go:164-196 AggregateVotes):
 * decode n signatures (with subgroup check) and sum them.  0 ok, HBLS_ERR_DECODE if one fails. */
int hbls_aggregate_sigs(const uint8_t* sig96, size_t n, uint8_t out96[96]);

/* FastAggregateVerify / VerifyAggregateSig of BASELINE.json == the reference composition
 * Deserialize(sig) ; SetMask(bitmap) ; aggSig.VerifyHash(mask.AggregatePublic, msg)
 * (internal/chain/engine.go:630-640, consensus/validator.go:219-236).  1 valid, 0 invalid, <0 error. */
int hbls_aggregate_verify(const hbls_committee* c, const uint8_t* bitmap, size_t blen,
                          const uint8_t sig96[96], const void* msg, size_t msg_len);

/* B independent rounds against one committee (block-range sync: api/service/stagedstreamsync/sig_verify.go:23-58).
 * bitmaps: B*blen bytes; sigs96: B*96; msgs: B*msg_len (msg_len <= 64); results[j] = 1/0. */
int hbls_aggregate_verify_batch(const hbls_committee* c, size_t B, const uint8_t* bitmaps, size_t blen,
                                const uint8_t* sigs96, const uint8_t* msgs, size_t msg_len, uint8_t* results);
/* How the two batch entries above check the pairing equations.
 * mode 1 (default): random-linear-combination groups -- 4 or 8 rounds share one Miller accumulator and one final exponentiation
 *   (prod_j [e(B, sigma_j) e(-apk_j, H_j)]^{r_j} == 1, fresh 64-bit r_j per call); if any group fails or holds an undecodable
 *   round, every round is recomputed exactly, so results are the exact booleans (a bad round survives the batched test
 *   with probability <= 2^-63).  Batches under 1024 rounds always use mode 0.
 * mode 0: the exact per-round check only (identical semantics to N calls of hbls_aggregate_verify). */
void hbls_set_batch_mode(int mode);
int  hbls_get_batch_mode(void);
/* same, every pointer already in device memory (HBM); stream = cudaStream_t or NULL; asynchronous on that stream */
int hbls_aggregate_verify_batch_device(const hbls_committee* c, size_t B, const void* d_bitmaps, size_t blen,
                                       const void* d_sigs96, const void* d_msgs, size_t msg_len,
                                       void* d_results, void* stream);

/* k independent (pk, msg, sig) triples (leader onPrepare/onCommit loop consensus/leader.go:127-182,227-290;
 * view-change storm consensus/view_change_construct.go:237-375): results[j] = Deserialize ok && VerifyHash. */
int hbls_verify_batch(size_t k, const uint8_t* pk48, const uint8_t* sig96, const uint8_t* msgs, size_t msg_len,
                      uint8_t* results);

/* batched SignHash / GetPublicKey (consensus/construct.go:97-114 with multibls keys) ; ok[j] = 1/0 */
int hbls_sign_hash_batch(size_t k, const uint8_t* sk32, const uint8_t* msgs, size_t msg_len, uint8_t* sig96_out, uint8_t* ok);
int hbls_get_public_key_batch(size_t k, const uint8_t* sk32, uint8_t* pk48_out);

/* ------------------------------------------------------------------ probes used by tests / bench */
/* message -> G2 point, serialized (the H(m) of SignHash/VerifyHash): 0 ok, -1 undefined */
int hbls_map_to_g2(const void* msg, size_t msg_len, uint8_t out96[96]);
/* n field products on canonical little-endian 48-byte operands (kernel parity probe) */
int hbls_fp_mul_batch(size_t n, const uint8_t* a48, const uint8_t* b48, uint8_t* out48);
/* number of kernels this library has launched so far */
uint64_t hbls_kernel_launch_count(void);
/* device self-test of the lane-pair (split Fp2) primitives used by k_pairing_verify_split against the single-thread
 * primitives on pseudo-random operands: returns the number of mismatches (0 = pass), <0 on error */
int hbls_selftest_split(uint32_t iters);
/* decode-step probe used while chasing a toolchain miscompile (tools/dbg_g2.py); 0 ok */
int hbls_debug_g2(const uint8_t sig96[96], uint8_t out512[512]);
/* per-kernel device timing of the aggregate-verify pipeline (CUDA events on the launching stream).
 * enable(1) makes every following hbls_aggregate_verify_batch[_device] call record events between its kernels;
 * get() waits for the last recorded pipeline and returns the number of stages written to ms_out, in launch order:
 * 0 k_mask_aggregate, 1 k_g1_normalize, 2 k_g2_decode, 3 k_hash_to_g2, 4 k_miller_verify, 5 k_final_verify */
void hbls_stage_timing_enable(int on);
int hbls_stage_timing_get(float* ms_out, int max_stages);
/* integer-pipe probe: runs `iters` dependent-free IMAD.WIDE.U32 MACs per thread on the whole chip and returns
 * the achieved MAC32/s (roofline denominator measured on this box), <0 on error */
double hbls_probe_mac32_per_s(int iters);

#ifdef __cplusplus
}
#endif
#endif
