/* raiko_kzg.h -- C ABI of the B200-native blob-KZG path (libraiko_kzg.so).
 *
 * Drop-in boundary for the functions raiko-lib drives in
 * /root/reference/lib/src/primitives/eip4844.rs (cited per entry point below).
 * The reference has no FFI layer for this path (plain Rust calls into the
 * rust-kzg crate); these are the symbols a `raiko-kzg-sys` crate binds instead
 * (see INTEGRATION.md for the Rust / ctypes stubs).
 *
 * Conventions: status code returned, caller-allocated outputs, plain pointers and
 * sizes, no exceptions cross the ABI, thread-safe and re-entrant (the reference
 * runs up to 16 requests concurrently, host/src/proof.rs:121).  There is NO CPU
 * fallback: without a usable CUDA device every compute entry point returns
 * RK_ERR_CUDA.  Field elements are 32-byte big-endian, group elements 48-byte
 * zcash-compressed G1, exactly the reference's KzgField / KzgGroup
 * (eip4844.rs:28-30).
 */
#ifndef RAIKO_KZG_H
#define RAIKO_KZG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RK_BYTES_PER_BLOB 131072
#define RK_FIELD_ELEMENTS_PER_BLOB 4096
#define RK_BYTES_PER_FIELD_ELEMENT 32
#define RK_BYTES_PER_G1 48
#define RK_VERSIONED_HASH_VERSION_KZG 0x01 /* eip4844.rs:26 */

typedef struct rk_kzg_ctx rk_kzg_ctx; /* opaque: device tables + streams for 1..8 GPUs */

typedef enum {
    RK_OK = 0,
    RK_ERR_BAD_LENGTH = 1,       /* Blob::from_bytes length check                         */
    RK_ERR_NONCANONICAL_FE = 2,  /* Eip4844Error::DeserializeBlob (eip4844.rs:34-35)      */
    RK_ERR_BAD_SETTINGS = 3,     /* settings image not recognised / fails validation      */
    RK_ERR_BAD_POINT = 4,        /* G1 encoding invalid (verification inputs)             */
    RK_ERR_CUDA = 5,             /* no device, launch or allocation failure               */
    RK_ERR_ARG = 6               /* NULL pointer, bad device index, ...                   */
} rk_status;

/* ---- settings: replaces KZG_SETTINGS_BIN / KZG_SETTINGS (eip4844.rs:14-24) -------------
 * `settings` may be any of: kzg_settings_raw.bin (739 624 B), the bincode image
 * lib/kzg_settings/zkcrypto_kzg_settings.bin (1 001 905 B) that eip4844.rs:14 embeds, or
 * this repo's compact image ("RKZGTS02": compressed G1, affine G2).  `devices` lists CUDA device
 * ordinals (NULL/0 = current device only).  `window_bits` selects the fixed-base table
 * width c in 4..15 (0 = RAIKO_KZG_WINDOW_BITS from the environment, else the largest
 * table that fits comfortably in free HBM; c = 15 needs ~115 GB per GPU).              */
rk_status rk_kzg_ctx_create(const uint8_t* settings, size_t len, const int* devices, int ndev,
                            rk_kzg_ctx** out);
rk_status rk_kzg_ctx_create_ex(const uint8_t* settings, size_t len, const int* devices, int ndev,
                               int window_bits, rk_kzg_ctx** out);
void rk_kzg_ctx_destroy(rk_kzg_ctx* ctx);
int rk_kzg_ctx_window_bits(const rk_kzg_ctx* ctx);
int rk_kzg_ctx_num_devices(const rk_kzg_ctx* ctx);
/* bytes of HBM held by the window table on each device */
uint64_t rk_kzg_ctx_table_bytes(const rk_kzg_ctx* ctx);
/* 1 when window_bits was left to the library and a device lacked the free HBM for c = 15, so a
 * narrower (slower) table was built; the choice is also reported on stderr at creation time.  */
int rk_kzg_ctx_window_reduced(const rk_kzg_ctx* ctx);

/* Re-serialise the loaded setup in the reference's own on-disk layouts (what
 * host/src/bin/gen_kzg_settings.rs:8-22 produces).  kind 0 = raw (739 624 B),
 * 1 = bincode (1 001 905 B).  *len in: capacity, out: bytes written.                  */
rk_status rk_kzg_ctx_export_settings(rk_kzg_ctx* ctx, int kind, uint8_t* out, size_t* len);

/* ---- single-blob drop-ins (thread-safe; calls that arrive while the device is busy are
 * merged into one batch per kind, so 16 concurrent requests cost about one) --------------- */
/* calc_kzg_proof_commitment (eip4844.rs:80-89) / blob_to_kzg_commitment_rust             */
rk_status rk_blob_to_kzg_commitment(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len,
                                    uint8_t out_commitment[48]);
/* commitment_to_version_hash (eip4844.rs:91-95); host-only, no ctx needed               */
rk_status rk_kzg_to_versioned_hash(const uint8_t commitment[48], uint8_t out_hash[32]);
/* get_evaluation_point (eip4844.rs:44-48): sha256(sha256(blob) || vh) mod r             */
rk_status rk_get_evaluation_point(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len,
                                  const uint8_t versioned_hash[32], uint8_t out_x[32]);
/* proof_of_equivalence (eip4844.rs:50-65) -> (x, y)                                     */
rk_status rk_proof_of_equivalence(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len,
                                  const uint8_t versioned_hash[32], uint8_t out_x[32],
                                  uint8_t out_y[32]);
/* calc_kzg_proof_with_point (eip4844.rs:71-78) / compute_kzg_proof_rust; z is reduced
 * mod r like Fr::from_bytes_unchecked; out_y may be NULL                                */
rk_status rk_compute_kzg_proof(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len,
                               const uint8_t z[32], uint8_t out_proof[48], uint8_t out_y[32]);
/* calc_kzg_proof (eip4844.rs:67-69)                                                     */
rk_status rk_calc_kzg_proof(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len,
                            const uint8_t versioned_hash[32], uint8_t out_proof[48]);

/* ---- batch entry points (the measured path) -------------------------------------------
 * `blobs` holds n contiguous 131 072-byte blobs and may be a host pointer (pageable or
 * pinned) or a device pointer on ctx device 0 (then outputs must be device pointers on
 * the same device as well, or host pointers).  Blobs are sharded contiguously over the
 * ctx's devices, streamed in chunks, and the outputs are gathered at their blob index.
 * A blob that fails to deserialize sets per_blob_status[i] = RK_ERR_NONCANONICAL_FE and
 * leaves its outputs zeroed without failing the batch.  Any output pointer except the
 * first may be NULL.  Device inputs must be complete on entry: work queued on the legacy
 * default stream is ordered before the batch automatically, producers on other streams
 * must be synchronised by the caller.  All outputs are complete when the call returns.   */
rk_status rk_commit_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, size_t n,
                          uint8_t* out_commitments /* n*48 */,
                          uint8_t* out_versioned_hashes /* n*32 */,
                          uint8_t* per_blob_status /* n */);
/* commitment -> versioned hash -> raiko challenge x -> y = p(x) -> proof, per blob:
 * what preflight + run_prover compute per blob (core/src/preflight.rs:260,
 * core/src/interfaces.rs:207-219).                                                       */
rk_status rk_commit_prove_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, size_t n,
                                uint8_t* out_commitments /* n*48 */,
                                uint8_t* out_versioned_hashes /* n*32 */,
                                uint8_t* out_x /* n*32 */, uint8_t* out_y /* n*32 */,
                                uint8_t* out_proofs /* n*48 */,
                                uint8_t* per_blob_status /* n */);
/* compute_kzg_proof for n (blob, z) pairs                                                */
rk_status rk_compute_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, const uint8_t* zs /* n*32 */,
                                     size_t n, uint8_t* out_proofs /* n*48 */,
                                     uint8_t* out_y /* n*32 */, uint8_t* per_blob_status);

/* compute_blob_kzg_proof (Deneb spec; upstream compute_blob_kzg_proof_rust) for n blobs: the
 * EIP-4844 Fiat-Shamir challenge z_i = hash_to_bls_field(sha256("FSBLOBVERIFY_V1_" | be128(4096)
 * | blob_i | commitment_i)) -- NOT raiko's evaluation point -- then the proof at z_i.  Produces the
 * proofs rk_verify_blob_kzg_proof_batch checks (BASELINE.json configs[4]).  commitments: n*48 bytes,
 * host or device.                                                                             */
rk_status rk_compute_blob_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, const uint8_t* commitments,
                                          size_t n, uint8_t* out_proofs /* n*48 */, uint8_t* per_blob_status);

/* ---- verification (BASELINE.json configs[4]) -------------------------
 * verify_kzg_proof_rust as the reference's tests use it (eip4844.rs:176-183):
 * e(C - [y]G1 + [z]proof, G2) == e(proof, [s]G2).  *out_ok = 1 accept, 0 reject.  Invalid point
 * encodings (not on the curve, not in G1) return RK_ERR_BAD_POINT, non-canonical z / y
 * RK_ERR_NONCANONICAL_FE, like upstream's Err results.                                        */
rk_status rk_verify_kzg_proof(rk_kzg_ctx* ctx, const uint8_t commitment[48], const uint8_t z[32],
                              const uint8_t y[32], const uint8_t proof[48], int* out_ok);
/* verify_kzg_proof_batch (Deneb spec; the routine upstream's verify_blob_kzg_proof_batch ends in):
 * n tuples (C_i, z_i, y_i, proof_i) checked with ONE random-linear-combination transcript and two
 * pairings per <= 16384 tuples.  Inputs host or device.  This is also the whole-batch self-check of
 * rk_commit_prove_batch (z = x: every (C, x, y, proof) of a 65,536-blob batch in under a second). */
rk_status rk_verify_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* commitments /* n*48 */,
                                    const uint8_t* zs /* n*32 */, const uint8_t* ys /* n*32 */,
                                    const uint8_t* proofs /* n*48 */, size_t n, int* out_ok);
/* verify_blob_kzg_proof_batch (Deneb spec): per blob the EIP-4844 Fiat-Shamir challenge
 * (NOT raiko's evaluation point), y = p(z), then one random-linear-combination check with two
 * pairings, all on ctx device 0.  blobs may be host or device memory; commitments / proofs
 * n*48 bytes each.                                                                            */
rk_status rk_verify_blob_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, const uint8_t* commitments,
                                         const uint8_t* proofs, size_t n, int* out_ok);

/* ---- blob -> tx-list byte codec (lib/src/utils.rs:85-144 decode_blob_data; "next" rank 4) ----
 * n blobs in, per blob up to RK_BLOB_DATA_MAX decoded bytes written at out + i*RK_BLOB_DATA_STRIDE
 * and their count in out_len[i]; an invalid blob yields length 0 (the reference returns an empty
 * Vec).  blobs: host or device; out / out_len: both host or both device.                          */
#define RK_BLOB_DATA_MAX 130044
#define RK_BLOB_DATA_STRIDE 130048
rk_status rk_decode_blob_data_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, size_t n, uint8_t* out,
                                    uint32_t* out_len);

/* ---- synthetic input of the measurement contract (bench / test tooling) ------
 * Writes blobs first_blob .. first_blob + n - 1 of the deterministic family
 *   fe(b, i) = BE_int(sha256("raiko-kzg-bench-v1" | LE64(seed) | LE32(b) | LE32(i))) mod r
 * (32-byte big-endian field elements, i = 0..4095) to `out` (n * 131072 bytes, host or device).
 * The oracle generates the same bytes on the host (tests/kzg_testlib.py::synthetic_blob).        */
rk_status rk_synth_blobs(rk_kzg_ctx* ctx, uint64_t seed, uint32_t first_blob, size_t n, uint8_t* out);

/* ---- instrumentation ------------------------------------------------------------------ */
typedef struct {
    double msm_ms;        /* sum of MSM kernel durations (CUDA events on their stream)    */
    double fr_ms;         /* evaluation / quotient kernel                                 */
    double sha_ms;        /* SHA-256 kernels                                              */
    double finalize_ms;   /* partial-sum reduction, inversion, compression                */
    uint64_t msm_launches;
    uint64_t total_launches; /* every kernel this library launched                        */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t msm_point_adds; /* table additions issued (non-zero digits are data dependent:
                                this is the nominal count n * 4096 * windows)             */
    uint64_t msm_affine_launches; /* MSM launches that used the batched-affine kernel
                                (k_msm_affine); the rest used the XYZZ kernel (k_msm)     */
    uint64_t msm_affine_point_adds; /* the part of msm_point_adds those launches issued   */
} rk_kzg_stats;
/* enable = 1 records a CUDA-event pair around every kernel (adds a little host time)     */
void rk_kzg_stats_enable(rk_kzg_ctx* ctx, int enable);
void rk_kzg_stats_reset(rk_kzg_ctx* ctx);
void rk_kzg_stats_get(rk_kzg_ctx* ctx, rk_kzg_stats* out); /* summed over devices         */

/* Integer-multiply roofline microbenchmark on `device`: dependency-free mad.wide.u32
 * chains on all SMs; returns 32x32->64 multiply-accumulates per second.                  */
rk_status rk_measure_imad_peak(int device, double* out_macs_per_sec, double* out_sm_clock_mhz);

/* thread-local message for the last non-OK status (mirrors upstream's String errors)     */
const char* rk_last_error(void);
const char* rk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RAIKO_KZG_H */
