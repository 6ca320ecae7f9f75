//! What changes in core/src/preflight.rs when lib/src/primitives/eip4844.rs is replaced by
//! eip4844_gpu.rs.  Two options; either compiles against the drop-in.
//!
//! (a) Minimal: keep `calc_blob_versioned_hash` (core/src/preflight.rs:305-316) as it is and change
//!     only the import at :11-14 from
//!         use kzg_traits::{eip_4844::{blob_to_kzg_commitment_rust, Blob}, G1};
//!     to
//!         use raiko_lib::primitives::eip4844::{blob_to_kzg_commitment_rust, Blob};
//!     The replacement module provides `Blob::from_bytes`, `deserialize_blob_rust`,
//!     `blob_to_kzg_commitment_rust(&fields, &KZG_SETTINGS.clone())` and a `to_bytes()` on the result
//!     with the same call shapes, so lines 305-316 compile unchanged.
//!
//! (b) Batched: one GPU launch for all sidecars of the block instead of one MSM per sidecar, and the
//!     matched commitment is reused instead of being recomputed at :260.  Replaces :305-316 and the
//!     `.find(...)` at :366-378.
use anyhow::{ensure, Result};
use raiko_lib::primitives::eip4844;

fn blob_to_bytes(blob_str: &str) -> Vec<u8> {
    hex::decode(blob_str.to_lowercase().trim_start_matches("0x")).unwrap_or_default()
}

/// Returns (blob bytes, commitment) of the sidecar whose versioned hash is `blob_hash`.
pub fn find_tx_blob(blob_hash: [u8; 32], sidecar_blobs: &[String]) -> Result<(Vec<u8>, eip4844::KzgGroup)> {
    let mut all = Vec::with_capacity(sidecar_blobs.len() * eip4844::BYTES_PER_BLOB);
    for b in sidecar_blobs {
        let bytes = blob_to_bytes(b);
        ensure!(bytes.len() == eip4844::BYTES_PER_BLOB, "Could not create blob");
        all.extend_from_slice(&bytes);
    }
    // ONE rk_commit_batch over the <= 6 sidecars (Cancun maximum per block)
    let (commitments, hashes, status) = eip4844::commit_batch(&all).map_err(|e| anyhow::anyhow!(e))?;
    for i in 0..sidecar_blobs.len() {
        if status[i] == 0 && hashes[i].0 == blob_hash {
            let lo = i * eip4844::BYTES_PER_BLOB;
            return Ok((all[lo..lo + eip4844::BYTES_PER_BLOB].to_vec(), commitments[i]));
        }
    }
    anyhow::bail!("no sidecar blob matches the versioned hash of the proposal")
}
