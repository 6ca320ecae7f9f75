//! `extern "C"` surface of include/raiko_kzg.h (every function the header declares;
//! tests/test_rust_boundary.py keeps the two lists identical).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int};

#[repr(C)]
pub struct rk_kzg_ctx {
    _private: [u8; 0],
}
/// 0 OK, 1 BAD_LENGTH, 2 NONCANONICAL_FE, 3 BAD_SETTINGS, 4 BAD_POINT, 5 CUDA, 6 ARG
pub type rk_status = c_int;

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct rk_kzg_stats {
    pub msm_ms: c_double,
    pub fr_ms: c_double,
    pub sha_ms: c_double,
    pub finalize_ms: c_double,
    pub msm_launches: u64,
    pub total_launches: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub msm_point_adds: u64,
    pub msm_affine_launches: u64,
    pub msm_affine_point_adds: u64,
}

extern "C" {
    pub fn rk_kzg_ctx_create(settings: *const u8, len: usize, devices: *const c_int, ndev: c_int, out: *mut *mut rk_kzg_ctx) -> rk_status;
    pub fn rk_kzg_ctx_create_ex(settings: *const u8, len: usize, devices: *const c_int, ndev: c_int, window_bits: c_int, out: *mut *mut rk_kzg_ctx) -> rk_status;
    pub fn rk_kzg_ctx_destroy(ctx: *mut rk_kzg_ctx);
    pub fn rk_kzg_ctx_window_bits(ctx: *const rk_kzg_ctx) -> c_int;
    pub fn rk_kzg_ctx_num_devices(ctx: *const rk_kzg_ctx) -> c_int;
    pub fn rk_kzg_ctx_table_bytes(ctx: *const rk_kzg_ctx) -> u64;
    pub fn rk_kzg_ctx_window_reduced(ctx: *const rk_kzg_ctx) -> c_int;
    pub fn rk_kzg_ctx_export_settings(ctx: *mut rk_kzg_ctx, kind: c_int, out: *mut u8, len: *mut usize) -> rk_status;
    pub fn rk_blob_to_kzg_commitment(ctx: *mut rk_kzg_ctx, blob: *const u8, blob_len: usize, out: *mut u8) -> rk_status;
    pub fn rk_kzg_to_versioned_hash(commitment: *const u8, out: *mut u8) -> rk_status;
    pub fn rk_get_evaluation_point(ctx: *mut rk_kzg_ctx, blob: *const u8, blob_len: usize, vh: *const u8, out_x: *mut u8) -> rk_status;
    pub fn rk_proof_of_equivalence(ctx: *mut rk_kzg_ctx, blob: *const u8, blob_len: usize, vh: *const u8, out_x: *mut u8, out_y: *mut u8) -> rk_status;
    pub fn rk_compute_kzg_proof(ctx: *mut rk_kzg_ctx, blob: *const u8, blob_len: usize, z: *const u8, out_proof: *mut u8, out_y: *mut u8) -> rk_status;
    pub fn rk_calc_kzg_proof(ctx: *mut rk_kzg_ctx, blob: *const u8, blob_len: usize, vh: *const u8, out_proof: *mut u8) -> rk_status;
    pub fn rk_commit_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, n: usize, out_c: *mut u8, out_vh: *mut u8, status: *mut u8) -> rk_status;
    pub fn rk_commit_prove_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, n: usize, out_c: *mut u8, out_vh: *mut u8, out_x: *mut u8, out_y: *mut u8, out_proofs: *mut u8, status: *mut u8) -> rk_status;
    pub fn rk_compute_kzg_proof_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, zs: *const u8, n: usize, out_proofs: *mut u8, out_y: *mut u8, status: *mut u8) -> rk_status;
    pub fn rk_compute_blob_kzg_proof_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, commitments: *const u8, n: usize, out_proofs: *mut u8, status: *mut u8) -> rk_status;
    pub fn rk_verify_kzg_proof(ctx: *mut rk_kzg_ctx, commitment: *const u8, z: *const u8, y: *const u8, proof: *const u8, out_ok: *mut c_int) -> rk_status;
    pub fn rk_verify_kzg_proof_batch(ctx: *mut rk_kzg_ctx, commitments: *const u8, zs: *const u8, ys: *const u8, proofs: *const u8, n: usize, out_ok: *mut c_int) -> rk_status;
    pub fn rk_verify_blob_kzg_proof_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, commitments: *const u8, proofs: *const u8, n: usize, out_ok: *mut c_int) -> rk_status;
    pub fn rk_decode_blob_data_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, n: usize, out: *mut u8, out_len: *mut u32) -> rk_status;
    pub fn rk_synth_blobs(ctx: *mut rk_kzg_ctx, seed: u64, first_blob: u32, n: usize, out: *mut u8) -> rk_status;
    pub fn rk_kzg_stats_enable(ctx: *mut rk_kzg_ctx, enable: c_int);
    pub fn rk_kzg_stats_reset(ctx: *mut rk_kzg_ctx);
    pub fn rk_kzg_stats_get(ctx: *mut rk_kzg_ctx, out: *mut rk_kzg_stats);
    pub fn rk_measure_imad_peak(device: c_int, out_macs_per_sec: *mut c_double, out_sm_clock_mhz: *mut c_double) -> rk_status;
    pub fn rk_last_error() -> *const c_char;
    pub fn rk_version() -> *const c_char;
}
