//! `extern "C"` surface of include/raiko_kzg.h.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct rk_kzg_ctx {
    _private: [u8; 0],
}
/// 0 OK, 1 BAD_LENGTH, 2 NONCANONICAL_FE, 3 BAD_SETTINGS, 4 BAD_POINT, 5 CUDA, 6 ARG
pub type rk_status = c_int;

extern "C" {
    pub fn rk_kzg_ctx_create(settings: *const u8, len: usize, devices: *const c_int, ndev: c_int, out: *mut *mut rk_kzg_ctx) -> rk_status;
    pub fn rk_kzg_ctx_create_ex(settings: *const u8, len: usize, devices: *const c_int, ndev: c_int, window_bits: c_int, out: *mut *mut rk_kzg_ctx) -> rk_status;
    pub fn rk_kzg_ctx_destroy(ctx: *mut rk_kzg_ctx);
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
    pub fn rk_verify_kzg_proof(ctx: *mut rk_kzg_ctx, commitment: *const u8, z: *const u8, y: *const u8, proof: *const u8, out_ok: *mut c_int) -> rk_status;
    pub fn rk_verify_kzg_proof_batch(ctx: *mut rk_kzg_ctx, commitments: *const u8, zs: *const u8, ys: *const u8, proofs: *const u8, n: usize, out_ok: *mut c_int) -> rk_status;
    pub fn rk_verify_blob_kzg_proof_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, commitments: *const u8, proofs: *const u8, n: usize, out_ok: *mut c_int) -> rk_status;
    pub fn rk_decode_blob_data_batch(ctx: *mut rk_kzg_ctx, blobs: *const u8, n: usize, out: *mut u8, out_len: *mut u32) -> rk_status;
    pub fn rk_last_error() -> *const c_char;
    pub fn rk_version() -> *const c_char;
}
