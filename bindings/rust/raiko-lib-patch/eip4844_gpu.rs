//! Drop-in replacement for lib/src/primitives/eip4844.rs over raiko-kzg-sys.
//! Every public item of the original keeps its name; ZFr / ZG1 travel as their byte encodings.
use once_cell::sync::Lazy;
use raiko_kzg_sys as sys;
use reth_primitives::B256;

pub static KZG_SETTINGS_BIN: &[u8] = include_bytes!("../../kzg_settings/zkcrypto_kzg_settings.bin");

pub struct KZGSettings(*mut sys::rk_kzg_ctx);
unsafe impl Send for KZGSettings {}
unsafe impl Sync for KZGSettings {}
pub type KzgSettings = KZGSettings;
impl Drop for KZGSettings {
    fn drop(&mut self) {
        unsafe { sys::rk_kzg_ctx_destroy(self.0) }
    }
}

pub static KZG_SETTINGS: Lazy<KZGSettings> = Lazy::new(|| {
    let mut ctx = std::ptr::null_mut();
    let st = unsafe { sys::rk_kzg_ctx_create(KZG_SETTINGS_BIN.as_ptr(), KZG_SETTINGS_BIN.len(), std::ptr::null(), 0, &mut ctx) };
    assert_eq!(st, 0, "failed to load trusted setup, please run `cargo run --bin gen_kzg_settings`");
    KZGSettings(ctx)
});

pub const VERSIONED_HASH_VERSION_KZG: u8 = 0x01;
pub type KzgGroup = [u8; 48];
pub type KzgField = [u8; 32];
pub type KzgCommitment = KzgGroup;

#[derive(Debug, thiserror::Error)]
pub enum Eip4844Error {
    #[error("Failed to deserialize blob to field elements")]
    DeserializeBlob,
    #[error("Failed to evaluate polynomial at hashed point: {0}")]
    EvaluatePolynomial(String),
    #[error("Failed to compute KZG proof")]
    ComputeKzgProof(String),
    #[error("Failed set commitment proof")]
    KzgDataPoison(String),
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::rk_last_error()) }.to_string_lossy().into_owned()
}
fn map(st: i32, other: fn(String) -> Eip4844Error) -> Result<(), Eip4844Error> {
    match st {
        0 => Ok(()),
        1 | 2 => Err(Eip4844Error::DeserializeBlob),
        _ => Err(other(last_error())),
    }
}

pub fn get_evaluation_point(blob: &[u8], versioned_hash: &B256) -> KzgField {
    let mut x = [0u8; 32];
    let st = unsafe { sys::rk_get_evaluation_point(KZG_SETTINGS.0, blob.as_ptr(), blob.len(), versioned_hash.as_ptr(), x.as_mut_ptr()) };
    assert_eq!(st, 0, "{}", last_error());
    x
}

pub fn proof_of_equivalence(blob: &[u8], versioned_hash: &B256) -> Result<(KzgField, KzgField), Eip4844Error> {
    let (mut x, mut y) = ([0u8; 32], [0u8; 32]);
    map(unsafe { sys::rk_proof_of_equivalence(KZG_SETTINGS.0, blob.as_ptr(), blob.len(), versioned_hash.as_ptr(), x.as_mut_ptr(), y.as_mut_ptr()) },
        Eip4844Error::EvaluatePolynomial)?;
    Ok((x, y))
}

pub fn calc_kzg_proof(blob: &[u8], versioned_hash: &B256) -> Result<KzgGroup, Eip4844Error> {
    let mut p = [0u8; 48];
    map(unsafe { sys::rk_calc_kzg_proof(KZG_SETTINGS.0, blob.as_ptr(), blob.len(), versioned_hash.as_ptr(), p.as_mut_ptr()) },
        Eip4844Error::ComputeKzgProof)?;
    Ok(p)
}

pub fn calc_kzg_proof_with_point(blob: &[u8], z: KzgField) -> Result<KzgGroup, Eip4844Error> {
    let mut p = [0u8; 48];
    map(unsafe { sys::rk_compute_kzg_proof(KZG_SETTINGS.0, blob.as_ptr(), blob.len(), z.as_ptr(), p.as_mut_ptr(), std::ptr::null_mut()) },
        Eip4844Error::ComputeKzgProof)?;
    Ok(p)
}

pub fn calc_kzg_proof_commitment(blob: &[u8]) -> Result<KzgGroup, Eip4844Error> {
    let mut c = [0u8; 48];
    map(unsafe { sys::rk_blob_to_kzg_commitment(KZG_SETTINGS.0, blob.as_ptr(), blob.len(), c.as_mut_ptr()) },
        Eip4844Error::ComputeKzgProof)?;
    Ok(c)
}

pub fn commitment_to_version_hash(commitment: &[u8; 48]) -> B256 {
    let mut h = [0u8; 32];
    unsafe { sys::rk_kzg_to_versioned_hash(commitment.as_ptr(), h.as_mut_ptr()) };
    B256::new(h)
}

pub fn kzg_proof_to_bytes(proof: &KzgGroup) -> KzgGroup {
    *proof
}

// c-kzg era spellings (BASELINE.json north_star)
pub use calc_kzg_proof_commitment as blob_to_kzg_commitment;
pub use calc_kzg_proof_with_point as compute_kzg_proof;
pub use commitment_to_version_hash as kzg_to_versioned_hash;

/// One launch over all sidecar blobs of a block (core/src/preflight.rs:369-376 + :260).
pub fn commit_batch(blobs: &[u8]) -> Result<(Vec<KzgGroup>, Vec<B256>, Vec<u8>), Eip4844Error> {
    let n = blobs.len() / 131072;
    let (mut c, mut vh, mut st) = (vec![0u8; 48 * n], vec![0u8; 32 * n], vec![0u8; n]);
    map(unsafe { sys::rk_commit_batch(KZG_SETTINGS.0, blobs.as_ptr(), n, c.as_mut_ptr(), vh.as_mut_ptr(), st.as_mut_ptr()) },
        Eip4844Error::ComputeKzgProof)?;
    Ok((c.chunks(48).map(|x| x.try_into().unwrap()).collect(), vh.chunks(32).map(|x| B256::from_slice(x)).collect(), st))
}
