//! Drop-in replacement for lib/src/primitives/eip4844.rs over raiko-kzg-sys.
//! Every public item of the original keeps its name; ZFr / ZG1 travel as their byte encodings.
use once_cell::sync::Lazy;
use raiko_kzg_sys as sys;
use reth_primitives::B256;

pub static KZG_SETTINGS_BIN: &[u8] = include_bytes!("../../kzg_settings/zkcrypto_kzg_settings.bin");

/// Owner of the device context: the window tables live on the GPU(s) until the last clone drops.
struct Ctx(*mut sys::rk_kzg_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::rk_kzg_ctx_destroy(self.0) }
    }
}

/// Same name as the item eip4844.rs:13 re-exports (`kzg::kzg_proofs::KZGSettings`).  The reference
/// clones its settings on every call (eip4844.rs:60,75,85; core/src/preflight.rs:311): `Clone` is
/// kept, and is a reference-count bump here instead of a 1 MB deep copy.
#[derive(Clone)]
pub struct KZGSettings(std::sync::Arc<Ctx>);
pub type KzgSettings = KZGSettings;
impl KZGSettings {
    pub fn from_image(image: &[u8]) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { sys::rk_kzg_ctx_create(image.as_ptr(), image.len(), std::ptr::null(), 0, &mut ctx) };
        if st != 0 {
            return Err(last_error());
        }
        Ok(KZGSettings(std::sync::Arc::new(Ctx(ctx))))
    }
    fn raw(&self) -> *mut sys::rk_kzg_ctx {
        (self.0).0
    }
}

pub static KZG_SETTINGS: Lazy<KZGSettings> = Lazy::new(|| {
    KZGSettings::from_image(KZG_SETTINGS_BIN)
        .expect("failed to load trusted setup, please run `cargo run --bin gen_kzg_settings`")
});

// ---- the two upstream items core/src/preflight.rs:305-316 uses through this module -------------
// (`pub use kzg::{eip_4844::deserialize_blob_rust, ...}` at eip4844.rs:13 and
//  `kzg_traits::eip_4844::{blob_to_kzg_commitment_rust, Blob}` at core/src/preflight.rs:11-14),
// kept call-compatible so `calc_blob_versioned_hash` compiles unchanged against the drop-in.
pub const BYTES_PER_BLOB: usize = 131072;
/// r, big-endian (Appendix A of SURVEY.md)
const FR_MODULUS_BE: [u8; 32] = [
    0x73, 0xed, 0xa7, 0x53, 0x29, 0x9d, 0x7d, 0x48, 0x33, 0x39, 0xd8, 0x08, 0x09, 0xa1, 0xd8, 0x05,
    0x53, 0xbd, 0xa4, 0x02, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0xff, 0xff, 0xff, 0x00, 0x00, 0x00, 0x01,
];
pub struct Blob {
    pub bytes: Vec<u8>,
}
impl Blob {
    /// `Blob::from_bytes`: length check only
    pub fn from_bytes(bytes: &[u8]) -> Result<Self, String> {
        if bytes.len() != BYTES_PER_BLOB {
            return Err(format!("Invalid byte length. Expected {} got {}", BYTES_PER_BLOB, bytes.len()));
        }
        Ok(Blob { bytes: bytes.to_vec() })
    }
}
/// The blob's field elements as the library consumes them: 32-byte big-endian, canonical.
pub struct BlobFields(pub Vec<u8>);
/// `deserialize_blob_rust`: every 32-byte chunk must be a canonical field element (< r).  A byte
/// comparison, no arithmetic; the GPU repeats the check on its own copy of the blob.
pub fn deserialize_blob_rust(blob: &Blob) -> Result<BlobFields, String> {
    for chunk in blob.bytes.chunks(32) {
        if chunk >= &FR_MODULUS_BE[..] {
            return Err("Invalid scalar".to_owned());
        }
    }
    Ok(BlobFields(blob.bytes.clone()))
}
/// 48-byte compressed G1 with the `to_bytes()` the call site expects (`G1::to_bytes`).
pub struct G1Bytes(pub KzgGroup);
impl G1Bytes {
    pub fn to_bytes(&self) -> KzgGroup {
        self.0
    }
}
/// `blob_to_kzg_commitment_rust(&fields, &settings)` (core/src/preflight.rs:309-312)
pub fn blob_to_kzg_commitment_rust(fields: &BlobFields, settings: &KZGSettings) -> Result<G1Bytes, String> {
    let mut c = [0u8; 48];
    let st = unsafe { sys::rk_blob_to_kzg_commitment(settings.raw(), fields.0.as_ptr(), fields.0.len(), c.as_mut_ptr()) };
    if st != 0 {
        return Err(last_error());
    }
    Ok(G1Bytes(c))
}

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
    let st = unsafe { sys::rk_get_evaluation_point(KZG_SETTINGS.raw(), blob.as_ptr(), blob.len(), versioned_hash.as_ptr(), x.as_mut_ptr()) };
    assert_eq!(st, 0, "{}", last_error());
    x
}

pub fn proof_of_equivalence(blob: &[u8], versioned_hash: &B256) -> Result<(KzgField, KzgField), Eip4844Error> {
    let (mut x, mut y) = ([0u8; 32], [0u8; 32]);
    map(unsafe { sys::rk_proof_of_equivalence(KZG_SETTINGS.raw(), blob.as_ptr(), blob.len(), versioned_hash.as_ptr(), x.as_mut_ptr(), y.as_mut_ptr()) },
        Eip4844Error::EvaluatePolynomial)?;
    Ok((x, y))
}

pub fn calc_kzg_proof(blob: &[u8], versioned_hash: &B256) -> Result<KzgGroup, Eip4844Error> {
    let mut p = [0u8; 48];
    map(unsafe { sys::rk_calc_kzg_proof(KZG_SETTINGS.raw(), blob.as_ptr(), blob.len(), versioned_hash.as_ptr(), p.as_mut_ptr()) },
        Eip4844Error::ComputeKzgProof)?;
    Ok(p)
}

pub fn calc_kzg_proof_with_point(blob: &[u8], z: KzgField) -> Result<KzgGroup, Eip4844Error> {
    let mut p = [0u8; 48];
    map(unsafe { sys::rk_compute_kzg_proof(KZG_SETTINGS.raw(), blob.as_ptr(), blob.len(), z.as_ptr(), p.as_mut_ptr(), std::ptr::null_mut()) },
        Eip4844Error::ComputeKzgProof)?;
    Ok(p)
}

pub fn calc_kzg_proof_commitment(blob: &[u8]) -> Result<KzgGroup, Eip4844Error> {
    let mut c = [0u8; 48];
    map(unsafe { sys::rk_blob_to_kzg_commitment(KZG_SETTINGS.raw(), blob.as_ptr(), blob.len(), c.as_mut_ptr()) },
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
    map(unsafe { sys::rk_commit_batch(KZG_SETTINGS.raw(), blobs.as_ptr(), n, c.as_mut_ptr(), vh.as_mut_ptr(), st.as_mut_ptr()) },
        Eip4844Error::ComputeKzgProof)?;
    Ok((c.chunks(48).map(|x| x.try_into().unwrap()).collect(), vh.chunks(32).map(|x| B256::from_slice(x)).collect(), st))
}
